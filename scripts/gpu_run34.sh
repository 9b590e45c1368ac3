#!/bin/bash
# compute-sanitizer memcheck over a small slice of the GPU tests (every kernel family once)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 \
  python -m pytest tests/test_gpu_faceclass.py tests/test_gpu_parity.py -m gpu -x -q \
  -k "golden or cluster_pairs or mining or stream_ordered or cross_entropy_and_grads" > gpurun_out/sanitizer.log 2>&1
echo "sanitizer exit=$?" >> gpurun_out/sanitizer.log
tail -25 gpurun_out/sanitizer.log | cut -c1-250
