#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
bash scripts/gpu_tests.sh
L=gpurun_out/probe10.log
: > $L
run() { echo "=== FNB_DEBUG=$FNB_DEBUG $*" >> $L; timeout 120 python scripts/gpu_probe.py "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
export FNB_DEBUG=0
run hist fp16f8 2 300 512 5 2
run accuracy fp16f8
for mode in fp16x3 fp16f8 bf16 tf32x3; do
  run bench $mode 2 100000 512 4 1
done
run bench fp16x3 2 100000 512 4 2
run bench fp16x3 1 100000 512 4 1
for mode in fp16x3 fp16f8; do
run bench $mode 2 400000 512 5 1
done
FNB_DEBUG=1 run bench fp16f8 2 100000 512 4 1
FNB_DEBUG=1 run bench fp16x3 2 100000 512 4 1
FNB_DEBUG=1 run bench bf16 2 100000 512 4 1
FNB_DEBUG=3 run bench fp16x3 2 100000 512 4 1
grep -E "===|bench|accuracy|identical|exit=[1-9]" $L | awk '/===/{h=$0; c=0; print} /accuracy|identical|exit/{print} /bench/{c++; if (c>=3) print}' | cut -c1-200
