#!/bin/bash
# GPU experiment: DRAM bytes of the fused kernel for a few schedules (is the second read of X served from L2?)
# usage: tests/_fused_ncu.sh "depth,hint;..."   -> gpurun_out/fused_ncu.txt
out=gpurun_out/fused_ncu.txt
: > $out
IFS=';' read -ra CFGS <<< "$1"
for c in "${CFGS[@]}"; do
  echo "== $c" >> $out
  SWEEP="$c" ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
     --clock-control none -k regex:k_fused -s 3 -c 1 python tests/_fused_sweep.py 2>&1 | grep -E "gpu__time|dram__bytes|hit_rate|tensor|fused depth" >> $out
done
cat $out
