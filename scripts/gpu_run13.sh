#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
bash scripts/gpu_tests.sh
python bench.py --workload 100k --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_100k_auto.json 2> gpurun_out/bench_100k_auto.err; echo "exit=$?"; cat gpurun_out/bench_100k_auto.json; tail -3 gpurun_out/bench_100k_auto.err
L=gpurun_out/probe13.log
: > $L
run() { echo "=== FNB_DEBUG=$FNB_DEBUG $*" >> $L; timeout 120 python scripts/gpu_probe.py "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
export FNB_DEBUG=0
for mode in fp16x3 fp16f8 bf16; do
  run bench $mode 2 100000 512 4 1
done
for mode in fp16x3 fp16f8; do
run bench $mode 2 400000 512 5 1
done
grep -E "===|bench|exit=[1-9]" $L | awk '/===/{h=$0; c=0; print} /exit/{print} /bench/{c++; if (c>=3) print}' | cut -c1-200
