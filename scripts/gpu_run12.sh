#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
python bench.py --workload 100k --mode fp16f8 --steps 5 --warmup 3 > gpurun_out/bench_100k_f8.json 2> gpurun_out/bench_100k_f8.err; echo "exit=$?"; cat gpurun_out/bench_100k_f8.json; tail -3 gpurun_out/bench_100k_f8.err
python bench.py --mode fp16f8 --steps 3 --warmup 3 > gpurun_out/bench_1m_f8.json 2> gpurun_out/bench_1m_f8.err; echo "exit=$?"; cat gpurun_out/bench_1m_f8.json; tail -3 gpurun_out/bench_1m_f8.err
python bench.py --mode fp16x3 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_1m_x3.json 2> gpurun_out/bench_1m_x3.err; echo "exit=$?"; cat gpurun_out/bench_1m_x3.json; tail -3 gpurun_out/bench_1m_x3.err
python scripts/bench_configs.py c1 c3 > gpurun_out/configs.jsonl 2> gpurun_out/configs.err; echo "exit=$?"; cat gpurun_out/configs.jsonl; tail -5 gpurun_out/configs.err
python -m pytest tests -m gpu -x -q -k "cluster_pairs or fp16f8 or lfw_size" 2>&1 | tail -5
