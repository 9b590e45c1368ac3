#!/bin/bash
# ncu on the DEFAULT bench command (1M x 512): launch list + full-set capture of one Gram launch
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity"
$CMD > gpurun_out/plain_1m.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_1m.csv $CMD > gpurun_out/ncu_list_1m.log 2>&1
echo "ncu list exit=$?"
$CMD > gpurun_out/plain2_1m.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gram_kernel -s 1 -c 1 -o gpurun_out/prof_gram_1m $CMD > gpurun_out/ncu_full_1m.log 2>&1
echo "ncu full exit=$?"; tail -3 gpurun_out/ncu_full_1m.log; tail -1 gpurun_out/plain_1m.log | cut -c1-300
