#!/bin/bash
# where does L2 retention of the row panels break as N grows?  (one pair per cluster, region_rows 16384)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
L=gpurun_out/probe46.log
: > $L
for n in 200000 400000; do
  echo "=== N=$n pairs=1 rr=16384" >> $L
  timeout 200 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:gram_kernel -s 1 -c 1 \
    python scripts/gpu_probe.py bench fp16f8 2 $n 512 2 1 16384 2>&1 | grep -E "dram__bytes_read|gpu__time_duration|hit_rate" >> $L
done
cat $L | cut -c1-200
