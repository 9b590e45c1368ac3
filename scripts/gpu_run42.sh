#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gpu_resident or golden or confidence" > gpurun_out/tests42.log 2>&1; echo "pytest exit=$?" >> gpurun_out/tests42.log
tail -12 gpurun_out/tests42.log
timeout 300 python scripts/bench_configs.py c1 > gpurun_out/configs_c1.jsonl 2> gpurun_out/configs_c1.err; echo "c1 exit=$?"; cut -c1-700 gpurun_out/configs_c1.jsonl; tail -3 gpurun_out/configs_c1.err
