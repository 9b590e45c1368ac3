#!/bin/bash
# cluster-pairs (A multicast) bring-up: parity vs the single-pair kernel, then timings
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
L=gpurun_out/probe8.log
: > $L
run() { echo "=== $*" >> $L; timeout 120 python scripts/gpu_probe.py "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
run hist fp16x3 2 300 512 5 2
run hist fp16f8 2 300 512 5 2
run hist bf16 2 300 512 5 2
run accuracy fp16f8
for mode in fp16x3 fp16f8 bf16; do
  run bench $mode 2 100000 512 4 1
  run bench $mode 2 100000 512 4 2
done
run bench fp16x3 2 400000 512 4 1
run bench fp16x3 2 400000 512 4 2
run bench fp16f8 2 400000 512 4 1
run bench fp16f8 2 400000 512 4 2
grep -v "^   range" $L
