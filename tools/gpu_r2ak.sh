# round 2, call ak: ERK_ADJ / dense-Jacobian ERK_TLM side lines (+ the FD-mode TLM and FWD/ERK lines again: the TLM kernel grew), parity campaign
set -x
mkdir -p gpurun_out
for w in swe_erk_adj swe_erk_tlm_dense swe_erk_tlm swe_erk; do
  timeout 600 python bench.py --workload $w --steps 2 --warmup 1 > gpurun_out/r2ak_side_$w.json 2> gpurun_out/r2ak_side_$w.err; tail -c 400 gpurun_out/r2ak_side_$w.json; tail -3 gpurun_out/r2ak_side_$w.err
done
timeout 1500 python tools/parity_campaign.py 128 > gpurun_out/r2ak_parity_campaign.log 2>&1; echo "rc=$?" >> gpurun_out/r2ak_parity_campaign.log; grep -n "ERK\|MISMATCH\|rc=" gpurun_out/r2ak_parity_campaign.log
