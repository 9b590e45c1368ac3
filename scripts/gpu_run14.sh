#!/bin/bash
# two ranks over NCCL: bench at N=2 (1M workload) + integer-histogram equality against the 1-GPU result at 100k
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_1m_n2.json 2> gpurun_out/bench_1m_n2.err; echo "exit=$?"; cat gpurun_out/bench_1m_n2.json; tail -5 gpurun_out/bench_1m_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 scripts/check_multi_gpu.py > gpurun_out/check_n2.log 2>&1; echo "exit=$?"; tail -5 gpurun_out/check_n2.log
