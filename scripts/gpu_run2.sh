#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/probe2.log
: > $L
run() { echo "=== $*" >> $L; timeout 300 python scripts/gpu_probe.py "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
run hist fp16x3 2 300 512 5
run hist fp16x3 1 300 512 5
for mode in fp16x3 tf32 bf16 tf32x3; do
  run bench $mode 2 20000
done
run bench fp16x3 1 20000
for mode in fp16x3 tf32 bf16; do
  run bench $mode 2 100000
done
cat $L
