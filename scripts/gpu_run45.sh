#!/bin/bash
# DRAM bytes of one 1M Gram launch WITHOUT multicast (one pair per cluster) against the super-row height
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
L=gpurun_out/probe45.log
: > $L
for rr in 16384 32768; do
  echo "=== pairs=1 rr=$rr" >> $L
  timeout 250 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:gram_kernel -s 1 -c 1 \
    python scripts/gpu_probe.py bench fp16f8 2 1000000 512 2 1 $rr 2>&1 | grep -E "dram__bytes_read|gpu__time_duration|hit_rate" >> $L
done
cat $L | cut -c1-200
