#!/bin/bash
# auto tile order / auto cluster pairs: tests, decomposition with representative data, bench lines
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
bash scripts/gpu_tests.sh > gpurun_out/tests21.log 2>&1
tail -4 gpurun_out/tests21.log
L=gpurun_out/probe21.log
: > $L
run() { echo "=== FNB_DEBUG=$FNB_DEBUG $*" >> $L; timeout 200 python scripts/gpu_probe.py "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
export FNB_DEBUG=3
run bench fp16f8 2 1000000 512 3 1 32768
export FNB_DEBUG=1
run bench fp16f8 2 1000000 512 3 1 32768
export FNB_DEBUG=0
run bench fp16f8 2 1000000 512 3 0 0
run bench fp16f8 2 100000 512 4 0 0
run bench fp16x3 2 100000 512 4 0 0
grep -E "===|bench|exit=[1-9]" $L | awk '/===/{h=$0; c=0; print} /exit/{print} /bench/{c++; if (c>=2) print}' | cut -c1-200
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_1m_r01c.json 2> gpurun_out/bench_1m_r01c.err; tail -c 3000 gpurun_out/bench_1m_r01c.json
python bench.py --workload 100k --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_100k_r01c.json 2> gpurun_out/bench_100k_r01c.err; tail -c 1500 gpurun_out/bench_100k_r01c.json
