# round 2, call w: RK_ADJ direct mode + general-pivoting cliff side lines, parity campaign with the two-warps-per-cell backward kernel
set -x
mkdir -p gpurun_out
for w in cbm4_rk_adj cbm4_rk_adj_direct cbm4_rk_adj_rtol1e-5 cbm4_rk_adj_rtol1e-4; do
  timeout 600 python bench.py --workload $w --steps 2 --warmup 1 > gpurun_out/r2w_side_$w.json 2> gpurun_out/r2w_side_$w.err; tail -c 700 gpurun_out/r2w_side_$w.json
done
timeout 1200 python tools/parity_campaign.py 128 > gpurun_out/r2w_parity_campaign.log 2>&1; echo "rc=$?" >> gpurun_out/r2w_parity_campaign.log; grep -n "ROS_ADJ\|MISMATCH\|rc=" gpurun_out/r2w_parity_campaign.log
