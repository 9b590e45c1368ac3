#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR bench.py --gpus $N --steps 5 --warmup 4 > gpurun_out/bench_1m_n$N.json 2> gpurun_out/bench_1m_n$N.err; echo "bench exit=$?"
grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"breakdown_ms": {[^}]*}\|"parallelism": "[^"]*"' gpurun_out/bench_1m_n$N.json | head -8
tail -2 gpurun_out/bench_1m_n$N.err | cut -c1-300
