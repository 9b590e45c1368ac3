#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
bash scripts/gpu_tests.sh > gpurun_out/tests41.log 2>&1; tail -5 gpurun_out/tests41.log
timeout 300 python scripts/ab_kernel.py _abtest/libfacenet_b200_head.so facenet_b200/_lib/libfacenet_b200.so > gpurun_out/ab41.log 2>&1; echo "exit=$?" >> gpurun_out/ab41.log
tail -8 gpurun_out/ab41.log | cut -c1-300
