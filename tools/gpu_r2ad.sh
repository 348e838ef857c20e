# round 2, call ad: SDIRK_TLM lock-step modes: side bench lines (direct solve, truncation-error control) and the parity campaign
set -x
mkdir -p gpurun_out
for w in cbm4_sdirk_tlm_direct cbm4_sdirk_tlm_trunc; do
  timeout 600 python bench.py --workload $w --steps 2 --warmup 1 > gpurun_out/r2ad_side_$w.json 2> gpurun_out/r2ad_side_$w.err; tail -c 900 gpurun_out/r2ad_side_$w.json; tail -3 gpurun_out/r2ad_side_$w.err
done
timeout 1500 python tools/parity_campaign.py 128 > gpurun_out/r2ad_parity_campaign.log 2>&1; echo "rc=$?" >> gpurun_out/r2ad_parity_campaign.log; grep -n "lock-step\|MISMATCH\|rc=" gpurun_out/r2ad_parity_campaign.log
