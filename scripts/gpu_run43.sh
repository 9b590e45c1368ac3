#!/bin/bash
# ncu of the 100k bench command (BASELINE config 2): launch list + full-set capture of one Gram launch
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
CMD="python bench.py --workload 100k --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity"
$CMD > gpurun_out/plain_100k.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_100k.csv $CMD > gpurun_out/ncu_list_100k.log 2>&1
echo "ncu list exit=$?"
ncu --set full --clock-control none --import-source on -k regex:gram_kernel -s 1 -c 1 -o gpurun_out/prof_gram_100k -f $CMD > gpurun_out/ncu_full_100k.log 2>&1
echo "ncu full exit=$?"
