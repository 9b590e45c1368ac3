#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "shard or rank or cluster_pairs" > gpurun_out/tests36.log 2>&1; echo "pytest exit=$?" >> gpurun_out/tests36.log
tail -12 gpurun_out/tests36.log
timeout 300 python scripts/check_world_emulated.py 1000000 8 0 0 2>&1 | cut -c1-700
