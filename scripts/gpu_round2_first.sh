#!/bin/bash
# PREPARED at the end of round 1 (GPU budget spent), NOT yet run: the measurements owed for the cluster-progress window.
#  1. tests + smoke;  2. bench lines (1M, 100k) with the window's default;  3. ncu launch list + --set full of the 1M default;
#  4. super-row height x window sweep with DRAM bytes (does a smaller super-row fit L2 better now that the clusters stay in step?)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest exit=$?" >> gpurun_out/r2_pytest_gpu.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke exit=$?" >> gpurun_out/r2_smoke.log
timeout 300 python bench.py > gpurun_out/r2_bench_1m.json 2> gpurun_out/r2_bench_1m.err
timeout 300 python bench.py --panel-window -1 --no-cpu-baseline > gpurun_out/r2_bench_1m_window_off.json 2> gpurun_out/r2_bench_1m_window_off.err
timeout 200 python bench.py --workload 100k --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_100k.json 2> gpurun_out/r2_bench_100k.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_1m.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-e2e > gpurun_out/r2_ncu_list.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gram_kernel -s 1 -c 1 -o gpurun_out/r2_prof_gram_1m -f \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-e2e > gpurun_out/r2_ncu_full.log 2>&1
timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none \
  -k regex:gram_kernel -c 13 python scripts/probe_rr_window.py 1000000 2 32768,32768,24576,16384,8192,49152 -1,2 2>&1 \
  | grep -E "dram__bytes|gpu__time_duration|hit_rate|rr=|rror" > gpurun_out/r2_rr_window.log
tail -3 gpurun_out/r2_pytest_gpu.log gpurun_out/r2_smoke.log; cat gpurun_out/r2_bench_1m.json | cut -c1-400; cut -c1-160 gpurun_out/r2_rr_window.log
