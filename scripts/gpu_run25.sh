#!/bin/bash
# cross-entropy epilogue tests + serpentine tile order (debug bit 5) at 1M / 100k; full GPU suite at the end
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
python -m pytest tests/test_gpu_faceclass.py -m gpu -x -q > gpurun_out/tests25.log 2>&1; echo "pytest exit=$?" >> gpurun_out/tests25.log
tail -30 gpurun_out/tests25.log
L=gpurun_out/probe25.log
: > $L
run() { echo "=== FNB_DEBUG=$FNB_DEBUG $*" >> $L; timeout 200 python scripts/gpu_probe.py "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
for dbg in 0 32 0 32; do
  export FNB_DEBUG=$dbg
  run bench fp16f8 2 1000000 512 3 2 32768
done
for dbg in 0 32; do
  export FNB_DEBUG=$dbg
  run bench fp16f8 2 1000000 512 3 2 49152
  run bench fp16f8 2 100000 512 4 1 16384
done
export FNB_DEBUG=0
grep -E "===|bench|exit=[1-9]" $L | awk '/===/{h=$0; c=0; print} /exit/{print} /bench/{c++; if (c>=2) print}' | cut -c1-200
bash scripts/gpu_tests.sh > gpurun_out/tests25_all.log 2>&1; tail -4 gpurun_out/tests25_all.log
