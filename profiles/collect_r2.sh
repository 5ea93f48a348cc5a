#!/bin/bash
# Round-2 ncu evidence, run under gpurun on ONE GPU:  gpurun -- 'bash profiles/collect_r2.sh'
# Writes the raw outputs to gpurun_out/; profiles/summarize.py turns them into the text summaries kept under profiles/.
set -x
B="python bench.py --no-e2e --no-cpu --no-secondary --no-parity --steps 2 --warmup 3"
# 1. launch lists (every kernel with its device time; cold-cache and serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_cfg3.csv $B --workload cfg3 > gpurun_out/r2_launches_cfg3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_cfg2.csv $B --workload cfg2 > gpurun_out/r2_launches_cfg2.log 2>&1
# 2. --set full of the two streaming kernels of each workload (third launch of each = steady state)
ncu --set full --clock-control none --import-source on -k regex:"k_h_update_tc|k_xht_tc" -s 6 -c 3 -f -o gpurun_out/r2_full_cfg3 $B --workload cfg3 > gpurun_out/r2_full_cfg3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_h_update_ts|k_xht_ts" -s 4 -c 2 -f -o gpurun_out/r2_full_cfg2 $B --workload cfg2 > gpurun_out/r2_full_cfg2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_h_update_ts|k_xht_ts" -s 4 -c 2 -f -o gpurun_out/r2_full_cfg4k64 $B --workload cfg4k64 > gpurun_out/r2_full_cfg4k64.log 2>&1
ls -la gpurun_out/r2_full_* gpurun_out/r2_launches_*
