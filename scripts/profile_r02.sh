#!/bin/bash
# Round-2 profiling pass (one GPU).  Every program first runs plain (must exit 0), then under ncu.  Outputs under gpurun_out/r02_*.
set -u
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-c3 --no-big-index --min-ms 1"
$B > gpurun_out/r02_plain_bench.log 2>&1 || { echo "plain bench failed"; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_bench.csv $B > gpurun_out/r02_ncu_bench.log 2>&1
echo "launch list rc=$?"
python scripts/flash_dev.py once 8192 > gpurun_out/r02_plain_flash.log 2>&1 || { echo "plain flash failed"; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"flash_kernel|fl_" -s 12 -c 6 -o gpurun_out/r02_flash_b8192 -f python scripts/flash_dev.py once 8192 > gpurun_out/r02_ncu_flash8k.log 2>&1
echo "flash 8192 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"flash_kernel" -s 4 -c 2 -o gpurun_out/r02_flash_b65536 -f python scripts/flash_dev.py once 65536 > gpurun_out/r02_ncu_flash64k.log 2>&1
echo "flash 65536 rc=$?"
python scripts/index_time.py 105542 > gpurun_out/r02_plain_index.log 2>&1 || { echo "plain index failed"; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"rowpanel|select_threshold|column_test|exact_score|sort_topk|prep_queries|index_exact" -s 20 -c 10 -o gpurun_out/r02_index_105k -f python scripts/index_time.py 105542 > gpurun_out/r02_ncu_index.log 2>&1
echo "index rc=$?"
timeout 900 ncu --set full --clock-control none -k regex:"rowpanel" -s 4 -c 2 -o gpurun_out/r02_index_10m -f python scripts/index_time.py 10000000 > gpurun_out/r02_ncu_index10m.log 2>&1
echo "index 10M rc=$?"
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python scripts/racecheck_unit.py > gpurun_out/r02_racecheck.log 2>&1
echo "racecheck rc=$?"
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python scripts/racecheck_unit.py > gpurun_out/r02_memcheck.log 2>&1
echo "memcheck rc=$?"
ls -la gpurun_out/r02_* | awk '{print $5, $9}'
