#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
L=gpurun_out/probe18.log
: > $L
run() { echo "=== $*" >> $L; timeout 200 python scripts/gpu_probe.py "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
for rr in 2048 8192 16384 32768; do
  run bench fp16f8 2 100000 512 4 1 $rr
done
for rr in 2048 8192 16384 32768; do
  run bench fp16f8 2 1000000 512 3 1 $rr
done
run bench fp16x3 2 1000000 512 3 1 2048
run bench fp16x3 2 1000000 512 3 1 16384
grep -E "===|bench|exit=[1-9]" $L | awk '/===/{h=$0; c=0; print} /exit/{print} /bench/{c++; if (c>=2) print}' | cut -c1-200
