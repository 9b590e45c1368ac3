#!/bin/bash
# DRAM bytes of one 1M Gram launch against the super-row height (single-pass ncu metrics)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
L=gpurun_out/probe44.log
: > $L
for rr in 16384 24576 32768; do
  echo "=== rr=$rr" >> $L
  timeout 250 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum --clock-control none -k regex:gram_kernel -s 1 -c 1 \
    python scripts/gpu_probe.py bench fp16f8 2 1000000 512 2 2 $rr 2>&1 | grep -E "dram__bytes_read|gpu__time_duration|bench mode" >> $L
done
cat $L | cut -c1-200
