#!/bin/bash
# bring-up: each case in its own process with a timeout; logs -> gpurun_out/
mkdir -p gpurun_out
L=gpurun_out/probe.log
: > $L
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm --format=csv >> $L 2>&1
run() { echo "=== $*" >> $L; timeout 180 python scripts/gpu_probe.py "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
for cg in 1 2; do
  for mode in fp16 fp16x3 bf16 tf32 tf32x3; do
    run pairwise $mode $cg
  done
done
run pairwise fp16x3 1 300 260 512
run pairwise fp16x3 2 300 260 512
for cg in 1 2; do
  run hist fp16x3 $cg
  run hist tf32x3 $cg
done
run hist fp16x3 2 300 512 5
for cg in 1 2; do
  run bench fp16x3 $cg
  run bench tf32 $cg
  run bench bf16 $cg
done
tail -100 $L
