#!/bin/bash
# cluster-progress window at 1M (DRAM bytes + time per launch; first launch is the cold one), then smoke() of the same build
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 60 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum --clock-control none -k regex:gram_kernel -c 7 \
  python scripts/probe_sync_window.py 1000000 2 32768 0 0 1 2 4 0 2 2>&1 | grep -E "dram__bytes_read|gpu__time_duration|window=|rror" > gpurun_out/probe48.log
cut -c1-160 gpurun_out/probe48.log
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke48.log 2>&1; echo "smoke exit=$?" >> gpurun_out/smoke48.log
tail -3 gpurun_out/smoke48.log
