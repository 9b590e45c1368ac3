#!/bin/bash
# power-capped regime (1M): where does the energy go?  debug 1 = no epilogue, 2 = no operand loads, 3 = both
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
L=gpurun_out/probe20.log
: > $L
run() { echo "=== FNB_DEBUG=$FNB_DEBUG $*" >> $L; timeout 200 python scripts/gpu_probe.py "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
for dbg in 0 1 2 3; do
  export FNB_DEBUG=$dbg
  run bench fp16f8 2 1000000 512 3 1 32768
done
export FNB_DEBUG=1
run bench fp16f8 2 1000000 512 3 2 32768
export FNB_DEBUG=0
run bench fp16f8 2 400000 512 4 1 32768
run bench fp16f8 2 400000 512 4 2 32768
run bench fp16f8 2 200000 512 4 1 16384
run bench fp16f8 2 200000 512 4 2 16384
run bench fp16x3 2 1000000 512 3 2 32768
run bench fp16x3 2 1000000 512 3 1 32768
grep -E "===|bench|exit=[1-9]" $L | awk '/===/{h=$0; c=0; print} /exit/{print} /bench/{c++; if (c>=2) print}' | cut -c1-200
