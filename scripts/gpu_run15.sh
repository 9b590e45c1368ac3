#!/bin/bash
# ncu: launch list of one bench command + full-set capture of the Gram kernel (same command, after it exited 0 without ncu)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
CMD="python bench.py --workload 100k --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit=$?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gram_kernel -s 1 -c 2 -o gpurun_out/prof_gram $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit=$?"; tail -3 gpurun_out/ncu_full.log; cat gpurun_out/plain.log | tail -2 | cut -c1-300
