#!/bin/bash
# GPU parity tests + smoke (logs under gpurun_out/)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -x -q "$@" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
