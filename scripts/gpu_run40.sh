#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 300 python scripts/ab_kernel.py _abtest/libfacenet_b200_3880091.so facenet_b200/_lib/libfacenet_b200.so > gpurun_out/ab40.log 2>&1; echo "exit=$?" >> gpurun_out/ab40.log
tail -12 gpurun_out/ab40.log | cut -c1-300
