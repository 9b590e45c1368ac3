#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
L=gpurun_out/probe11.log
: > $L
run() { echo "=== FNB_DEBUG=$FNB_DEBUG $*" >> $L; timeout 120 python scripts/gpu_probe.py "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
for dbg in 3 0; do
export FNB_DEBUG=$dbg
for d in 256 512 1024 2048; do
  run bench bf16 2 60000 $d 3 1
  run bench fp16x3 2 60000 $d 3 1
done
done
FNB_DEBUG=0 run accuracy fp16f8
grep -E "===|bench|accuracy|bias|d<0.5|exit=[1-9]" $L | awk '/===/{h=$0; c=0; print} /accuracy|bias|d<0.5|exit/{print} /bench/{c++; if (c==3) print}' | cut -c1-200
