#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
L=gpurun_out/probe9.log
: > $L
run() { echo "=== FNB_DEBUG=$FNB_DEBUG $*" >> $L; timeout 120 python scripts/gpu_probe.py "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
export FNB_DEBUG=0
for mode in fp16x3 fp16f8 bf16; do
  run bench $mode 2 100000 512 4 1
  run bench $mode 2 100000 512 4 2
done
FNB_DEBUG=4 run bench fp16f8 2 100000 512 4 1
FNB_DEBUG=4 run bench fp16f8 2 100000 512 4 2
for mode in fp16x3 fp16f8; do
run bench $mode 2 400000 512 5 1
run bench $mode 2 400000 512 5 2
done
FNB_DEBUG=4 run bench fp16f8 2 400000 512 5 1
FNB_DEBUG=4 run bench fp16f8 2 400000 512 5 2
FNB_DEBUG=1 run bench fp16f8 2 100000 512 4 2
FNB_DEBUG=1 run bench fp16x3 2 100000 512 4 2
FNB_DEBUG=1 run bench bf16 2 100000 512 4 2
grep -v "^   range" $L
