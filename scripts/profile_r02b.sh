#!/bin/bash
# Round-2 closing profile pass (one GPU): launch list of the final tree and ncu --set full of the kernels added late in the round.
set -u
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-c3 --no-big-index --no-hbm --min-ms 1"
$B > gpurun_out/r02b_plain_bench.log 2>&1 || { echo "plain bench failed"; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02b_launches_bench.csv $B > gpurun_out/r02b_ncu_bench.log 2>&1
echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dense_fwd_fused64|dense_bwd_fused64|sort_hist|sort_scan_flat|sort_scatter|stage_columns|sparse_block" -s 60 -c 16 -o gpurun_out/r02b_step_kernels -f $B > gpurun_out/r02b_ncu_step.log 2>&1
echo "step kernels rc=$?"
ls -la gpurun_out/r02b_* | awk '{print $5, $9}'
