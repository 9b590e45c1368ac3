#!/bin/bash
# classifier-path tests + L2 eviction hints (debug bit 4) x region rows at 1M
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_faceclass.py -m gpu -x -q > gpurun_out/tests24.log 2>&1; echo "pytest exit=$?" >> gpurun_out/tests24.log
tail -30 gpurun_out/tests24.log
L=gpurun_out/probe24.log
: > $L
run() { echo "=== FNB_DEBUG=$FNB_DEBUG $*" >> $L; timeout 200 python scripts/gpu_probe.py "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
for dbg in 0 16; do
  export FNB_DEBUG=$dbg
  for rr in 16384 24576 32768; do
    run bench fp16f8 2 1000000 512 3 2 $rr
  done
  run bench fp16f8 2 1000000 512 3 1 32768
  run bench fp16f8 2 100000 512 4 1 16384
done
grep -E "===|bench|exit=[1-9]" $L | awk '/===/{h=$0; c=0; print} /exit/{print} /bench/{c++; if (c>=2) print}' | cut -c1-200
