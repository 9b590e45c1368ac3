#!/bin/bash
# fixed row-block ownership per cluster (row blocks per super-row = a multiple of the number of clusters): DRAM bytes and time
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 200 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum --clock-control none -k regex:gram_kernel -c 4 \
  python scripts/probe_rr_pair.py 1000000 2 32768 33792 2>&1 | grep -E "dram__bytes_read|gpu__time_duration|rr=" > gpurun_out/probe47.log
cat gpurun_out/probe47.log | cut -c1-160
