#!/bin/bash
# tests + smoke + bench + ncu launch list + one full ncu capture of the Gram kernel
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
python bench.py --workload 100k --steps 5 --warmup 3 > gpurun_out/bench_100k.json 2> gpurun_out/bench_100k.err; echo "bench100k exit=$?"; cat gpurun_out/bench_100k.json; tail -3 gpurun_out/bench_100k.err
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_1m.json 2> gpurun_out/bench_1m.err; echo "bench1m exit=$?"; cat gpurun_out/bench_1m.json; tail -3 gpurun_out/bench_1m.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>&1; cat gpurun_out/bench_ref.json
CMD="python bench.py --workload 20k --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit=$?"
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:gram_kernel -s 1 -c 2 -o gpurun_out/prof_gram $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit=$?"; tail -3 gpurun_out/ncu_full.log
