#!/bin/bash
# Round-2 (second half) evidence, run under gpurun on ONE GPU:  gpurun -- 'bash profiles/collect_r2b.sh'
# Raw outputs go to gpurun_out/; profiles/summarize.py turns the ncu ones into the text summaries kept under profiles/.
set -x
B="python bench.py --no-e2e --no-cpu --no-secondary --no-parity --steps 2 --warmup 3"
# 0. the default bench line (what the driver runs) and the per-config lines, device-timed
( time python bench.py > gpurun_out/r2b_bench_default.log 2> gpurun_out/r2b_bench_default.err ) 2> gpurun_out/r2b_bench_default.time
for w in cfg1 cfg2 cfg4k16 cfg4k32 cfg4k64 cfg4k128 cfg4k256 cfg4k512 cfg5; do
  python bench.py --no-e2e --no-cpu --no-secondary --no-parity --workload $w > gpurun_out/r2b_bench_$w.log 2>&1
done
python bench.py --no-e2e --no-cpu --no-secondary --no-parity --workload cfg2 --mode h_only > gpurun_out/r2b_bench_cfg2_honly.log 2>&1
python bench.py --no-e2e --no-cpu --no-secondary --no-parity --workload cfg2 --mode w_only > gpurun_out/r2b_bench_cfg2_wonly.log 2>&1
# 1. launch lists (every kernel with its device time; cold-cache and serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches_cfg3.csv $B --workload cfg3 > gpurun_out/r2b_launches_cfg3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches_cfg2.csv $B --workload cfg2 > gpurun_out/r2b_launches_cfg2.log 2>&1
# 2. --set full of the two streaming kernels of each workload (steady-state launches)
ncu --set full --clock-control none --import-source on -k regex:"k_h_update_tc|k_xht_tc" -s 6 -c 3 -f -o gpurun_out/r2b_full_cfg3 $B --workload cfg3 > gpurun_out/r2b_full_cfg3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_h_update_ts|k_xht_ts" -s 4 -c 2 -f -o gpurun_out/r2b_full_cfg2 $B --workload cfg2 > gpurun_out/r2b_full_cfg2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_h_update_ts|k_xht_ts" -s 4 -c 2 -f -o gpurun_out/r2b_full_cfg4k64 $B --workload cfg4k64 > gpurun_out/r2b_full_cfg4k64.log 2>&1
ls -la gpurun_out/r2b_full_* gpurun_out/r2b_launches_*
