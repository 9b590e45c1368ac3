#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
bash scripts/gpu_tests.sh
L=gpurun_out/probe6.log
: > $L
run() { echo "=== $*" >> $L; timeout 300 python scripts/gpu_probe.py "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
run accuracy fp16f8
run accuracy fp16x3
run hist fp16f8 2 300 512 5
run bench fp16f8 2 100000
run bench fp16x3 2 100000
run bench fp16f8 2 1000000
run bench fp16x3 2 1000000
run bench bf16 2 1000000
cat $L
