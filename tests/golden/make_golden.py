#!/usr/bin/env python3
"""Collects the reference's own golden outputs for the hot path into small fixtures
(tests/golden/*.json).  The reference tree is not available on the GPU box, so the numbers are
copied here verbatim (they are run outputs committed upstream, not source code):
  EXAMPLES/cbm4/CBM4_ROS_ADJ/cbm4_ros_sol.txt          Y(Tend), 32 x E24.16
  EXAMPLES/cbm4/CBM4_ROS_ADJ/cbm4_output_sens.txt      Lambda(32,32) then Mu(29,32)
  EXAMPLES/small_strato/small_ros_adj/rkadj_ADJ_results.m   Lambda(5,5) (RosenbrockADJ2, Rodas3)
  EXAMPLES/small_strato/small_ros/small_strato_dr.F90:23-24  ANS(5) (REAL(4) literals)
  EXAMPLES/cbm4/CBM4_SDIRK/cbm4_sdirk_sol.txt          Y(Tend), 32 x E24.16 (FWD/SDIRK, Sdirk4a, rtol 1e-6, atol 1e-8)
  EXAMPLES/cbm4/CBM4_RK/cbm4_rk_sol.txt                Y(Tend), 32 x E24.16 (FWD/RK, Radau-2A, rtol 1e-6, atol 1e-8)
  EXAMPLES/cbm4/CBM4_RK_ADJ/{cbm4_rk_sol.txt,cbm4_output_sens.txt}   ADJ/RK_ADJ Radau-2A: Y(Tend), Lambda(32,32), Mu(29,32)
  EXAMPLES/cbm4/CBM4_SDIRK_ADJ/{cbm4_sdirk_sol.txt,cbm4_output_sens.txt}   ADJ/SDIRK_ADJ Sdirk4a, direct adjoint solve
  EXAMPLES/swe/SWE_ERK/swe_lsode_sol.txt               Y(50 dt), 4800 x E24.16 (LSODE solution the ERK driver compares with)
Usage: python tests/golden/make_golden.py [/root/reference]
"""
import json
import os
import re
import shutil
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    d = os.path.join(REF, "EXAMPLES/cbm4/CBM4_ROS_ADJ")
    y = [float(l) for l in open(os.path.join(d, "cbm4_ros_sol.txt"))]
    vals = [float(l) for l in open(os.path.join(d, "cbm4_output_sens.txt")) if not l.startswith(("L", "M"))]
    assert len(y) == 32 and len(vals) == 32 * 32 + 32 * 29
    json.dump({"source": "EXAMPLES/cbm4/CBM4_ROS_ADJ/{cbm4_ros_sol.txt,cbm4_output_sens.txt}",
               "y": y, "lambda": vals[:1024], "mu": vals[1024:]},
              open(os.path.join(HERE, "cbm4_ros_adj.json"), "w"))
    m = os.path.join(REF, "EXAMPLES/small_strato/small_ros_adj/rkadj_ADJ_results.m")
    lam = [float(x) for x in open(m).read().split()]
    assert len(lam) == 25
    drv = open(os.path.join(REF, "EXAMPLES/small_strato/small_ros/small_strato_dr.F90")).read()
    mm = re.search(r"ANS\s*=\s*\(/(.*?)/\)", drv, re.S)
    ans = [float(np.float32(x)) for x in re.sub(r"&", "", mm.group(1)).replace("\n", "").split(",")]
    assert len(ans) == 5
    json.dump({"source": "EXAMPLES/small_strato/small_ros_adj/rkadj_ADJ_results.m; small_ros/small_strato_dr.F90:23-24",
               "lambda": lam, "ans": ans}, open(os.path.join(HERE, "small_strato.json"), "w"))
    ys = [float(l) for l in open(os.path.join(REF, "EXAMPLES/cbm4/CBM4_SDIRK/cbm4_sdirk_sol.txt"))]
    assert len(ys) == 32
    json.dump({"source": "EXAMPLES/cbm4/CBM4_SDIRK/cbm4_sdirk_sol.txt", "y": ys}, open(os.path.join(HERE, "cbm4_sdirk.json"), "w"))
    yr = [float(l) for l in open(os.path.join(REF, "EXAMPLES/cbm4/CBM4_RK/cbm4_rk_sol.txt"))]
    assert len(yr) == 32
    json.dump({"source": "EXAMPLES/cbm4/CBM4_RK/cbm4_rk_sol.txt", "y": yr}, open(os.path.join(HERE, "cbm4_rk.json"), "w"))
    d = os.path.join(REF, "EXAMPLES/cbm4/CBM4_RK_ADJ")
    y = [float(l) for l in open(os.path.join(d, "cbm4_rk_sol.txt"))]
    vals = [float(l) for l in open(os.path.join(d, "cbm4_output_sens.txt")) if not l.startswith(("L", "M"))]
    assert len(y) == 32 and len(vals) == 32 * 32 + 32 * 29
    json.dump({"source": "EXAMPLES/cbm4/CBM4_RK_ADJ/{cbm4_rk_sol.txt,cbm4_output_sens.txt}",
               "y": y, "lambda": vals[:1024], "mu": vals[1024:]},
              open(os.path.join(HERE, "cbm4_rk_adj.json"), "w"))
    d = os.path.join(REF, "EXAMPLES/cbm4/CBM4_SDIRK_ADJ")
    ffloat = lambda t: float(re.sub(r"(\d)([+-]\d{3})\s*$", r"\1E\2", t.strip()))   # E24.16 drops the E for 3-digit exponents
    y = [ffloat(l) for l in open(os.path.join(d, "cbm4_sdirk_sol.txt"))]
    vals = [ffloat(l) for l in open(os.path.join(d, "cbm4_output_sens.txt")) if not l.startswith(("L", "M"))]
    assert len(y) == 32 and len(vals) == 32 * 32 + 32 * 29
    json.dump({"source": "EXAMPLES/cbm4/CBM4_SDIRK_ADJ/{cbm4_sdirk_sol.txt,cbm4_output_sens.txt}",
               "y": y, "lambda": vals[:1024], "mu": vals[1024:]},
              open(os.path.join(HERE, "cbm4_sdirk_adj.json"), "w"))
    shutil.copy(os.path.join(REF, "EXAMPLES/swe/SWE_ERK/swe_lsode_sol.txt"), os.path.join(HERE, "swe_lsode_sol.txt"))
    print("wrote fixtures")


if __name__ == "__main__":
    main()
