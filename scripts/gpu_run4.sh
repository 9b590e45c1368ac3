#!/bin/bash
# parity tests, then per-mode kernel timings (probe), logs under gpurun_out/
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
L=gpurun_out/probe4.log
: > $L
run() { echo "=== $*" >> $L; timeout 300 python scripts/gpu_probe.py "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
for mode in fp16x3 tf32 bf16 tf32x3; do
  run bench $mode 2 100000
done
run bench fp16x3 1 100000
run bench bf16 1 100000
grep -E "===|bench|exit=[1-9]|rror" $L
