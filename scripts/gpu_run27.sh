#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
bash scripts/gpu_tests.sh > gpurun_out/tests27.log 2>&1; tail -6 gpurun_out/tests27.log
python scripts/bench_configs.py f1 f2 --steps 200 > gpurun_out/configs_f1_f2.jsonl 2> gpurun_out/configs_f1_f2.err; echo "configs exit=$?"; cut -c1-900 gpurun_out/configs_f1_f2.jsonl; tail -3 gpurun_out/configs_f1_f2.err
