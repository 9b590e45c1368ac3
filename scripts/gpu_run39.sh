#!/bin/bash
# final 1-GPU pass: tests, smoke, bench lines (1M default, 100k), ncu of the default command
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
bash scripts/gpu_tests.sh > gpurun_out/tests39.log 2>&1; tail -5 gpurun_out/tests39.log
python bench.py > gpurun_out/bench_1m_final.json 2> gpurun_out/bench_1m_final.err; echo "bench exit=$?"; tail -c 1200 gpurun_out/bench_1m_final.json
python bench.py --workload 100k --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_100k_final.json 2> gpurun_out/bench_100k_final.err; echo "bench100k exit=$?"
grep -o '"value": [0-9.]*' gpurun_out/bench_100k_final.json | head -2
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_final.json 2>/dev/null; cut -c1-300 gpurun_out/bench_reference_final.json
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-parity"
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_1m.csv $CMD > gpurun_out/ncu_list_1m.log 2>&1
echo "ncu list exit=$?"
ncu --set full --clock-control none --import-source on -k regex:gram_kernel -s 1 -c 1 -o gpurun_out/prof_gram_1m -f $CMD > gpurun_out/ncu_full_1m.log 2>&1
echo "ncu full exit=$?"
