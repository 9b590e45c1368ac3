#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
bash scripts/gpu_tests.sh
L=gpurun_out/probe17.log
: > $L
run() { echo "=== FNB_DEBUG=$FNB_DEBUG $*" >> $L; timeout 120 python scripts/gpu_probe.py "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
export FNB_DEBUG=0
for mode in fp16x3 fp16f8 bf16; do
  run bench $mode 2 100000 512 4 1
done
run bench fp16f8 2 100000 512 4 1 4096
run bench fp16f8 2 100000 512 4 1 8192
run bench bf16 2 100000 512 4 1 8192
run bench fp16f8 2 400000 512 5 1
run bench fp16f8 2 400000 512 5 1 8192
run bench fp16x3 2 400000 512 5 1
grep -E "===|bench|exit=[1-9]" $L | awk '/===/{h=$0; c=0; print} /exit/{print} /bench/{c++; if (c>=3) print}' | cut -c1-200
