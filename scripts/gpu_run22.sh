#!/bin/bash
# N-GPU: sharded bins == single-GPU bins, then the headline bench under torchrun ($1 = N)
N=${1:-8}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR scripts/check_multi_gpu.py > gpurun_out/check_n$N.log 2>&1; echo "check exit=$?" >> gpurun_out/check_n$N.log
grep -E "world=|exit=" gpurun_out/check_n$N.log
timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_1m_n$N.json 2> gpurun_out/bench_1m_n$N.err; echo "bench exit=$?"
tail -c 2500 gpurun_out/bench_1m_n$N.json; tail -3 gpurun_out/bench_1m_n$N.err
