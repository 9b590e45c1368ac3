#!/bin/bash
# region-rows x L2-prefetch sweep (FNB_DEBUG bits 2-3 = l2_prefetch), power-capped 1M regime and 100k
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
L=gpurun_out/probe19.log
: > $L
run() { echo "=== FNB_DEBUG=$FNB_DEBUG $*" >> $L; timeout 200 python scripts/gpu_probe.py "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
for dbg in 0 4 12; do
  export FNB_DEBUG=$dbg
  for rr in 2048 8192 32768 65536; do
    run bench fp16f8 2 1000000 512 3 1 $rr
  done
done
for dbg in 0 4 12; do
  export FNB_DEBUG=$dbg
  for rr in 2048 16384; do
    run bench fp16f8 2 100000 512 4 1 $rr
  done
done
export FNB_DEBUG=0
run bench fp16f8 2 1000000 512 3 2 32768
export FNB_DEBUG=4
run bench fp16x3 2 1000000 512 3 1 32768
run bench bf16 2 1000000 512 3 1 32768
grep -E "===|bench|exit=[1-9]" $L | awk '/===/{h=$0; c=0; print} /exit/{print} /bench/{c++; if (c>=2) print}' | cut -c1-200
