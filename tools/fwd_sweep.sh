#!/bin/bash
# forward-kernel occupancy sweep: warps per SM (per-thread path)
for w in 1 2 3 4 6; do
  echo "== FATODE_FWD_WARPS=$w"
  FATODE_FWD_WARPS=$w python tools/prof_case.py 4736 72
done
