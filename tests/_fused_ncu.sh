#!/bin/bash
# GPU experiment: DRAM bytes of the fused kernel for a few schedules (is the second read of X served from L2?)
# usage: tests/_fused_ncu.sh "sbc,slab,bcols,lag,hint;..."   -> gpurun_out/fused_ncu.txt
out=gpurun_out/fused_ncu.txt
: > $out
IFS=';' read -ra CFGS <<< "$1"
for c in "${CFGS[@]}"; do
  echo "== $c" >> $out
  SWEEP="$c" ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,lts__t_sectors_srcunit_tex_op_read.sum \
     --clock-control none -k regex:k_fused -s 3 -c 1 python tests/_fused_sweep.py 2>&1 | grep -E "k_fused|gpu__time|dram__bytes|hit_rate|srcunit|ms/iter" >> $out
done
cat $out
