#!/bin/bash
# The ONE runner of GPU sessions:  gpurun --timeout S -- 'bash scripts/gpu_job.sh JOB [ARGS]'
# Every job writes into gpurun_out/ (merged back by gpurun); summaries worth judging are copied to profiles/ by hand.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
job=$1; shift

case "$job" in
tests)      # pytest -m gpu + smoke
    timeout 900 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest exit=$?" >> $O/pytest_gpu.log
    timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit=$?" >> $O/smoke.log
    tail -n 5 $O/pytest_gpu.log; tail -n 3 $O/smoke.log
    ;;
bias)       # signed error against the similarity, per arithmetic mode (probe_bias.py)
    timeout 600 python scripts/probe_bias.py "${1:-fp16x3,fp16f8,fp16}" ${2:-8192} ${3:-512} ${5:-gauss} ${6:-40} ${7:-corrected} > $O/bias_${4:-probe}.log 2>&1
    grep -v '^{' $O/bias_${4:-probe}.log | head -${LINES_MAX:-150}
    ;;
rrwindow)   # super-row height x cluster-progress window with DRAM bytes (one ncu metrics pass)
    timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none \
      -k regex:gram_kernel -c ${3:-13} python scripts/probe_rr_window.py 1000000 ${1:-2} ${2:-32768,32768,24576,16384,8192,49152} ${4:--1,2} 2>&1 \
      | grep -E "dram__bytes|gpu__time_duration|hit_rate|rr=|rror" > $O/rr_window.log
    cat $O/rr_window.log
    ;;
bench)      # bench.py with the given arguments; name of the output as first argument
    name=$1; shift
    timeout 900 python bench.py "$@" > $O/bench_$name.json 2> $O/bench_$name.err; echo "exit=$?" >> $O/bench_$name.err
    cut -c1-1500 $O/bench_$name.json; tail -3 $O/bench_$name.err
    ;;
ncu_list)   # launch list of one bench command
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$1.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-parity --no-e2e --no-strict "${@:2}" > $O/ncu_list_$1.log 2>&1
    tail -3 $O/ncu_list_$1.log
    ;;
ncu_full)   # --set full of one Gram launch of the bench command
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:gram_kernel -s ${NCU_SKIP:-2} -c 1 -o $O/prof_$1 -f \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --no-e2e --no-strict "${@:2}" > $O/ncu_full_$1.log 2>&1
    tail -3 $O/ncu_full_$1.log
    ;;
ncu_any)    # launch list (per-kernel durations) of any python command: name, then the command
    name=$1; shift
    timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c ${NCU_COUNT:-600} --csv --log-file $O/launches_$name.csv \
      python "$@" > $O/ncu_any_$name.log 2>&1
    python scripts/ncu_launch_summary.py $O/launches_$name.csv | tee $O/launches_$name.txt
    ;;
py)         # any python script with arguments, log name first
    name=$1; shift
    timeout 1200 python "$@" > $O/$name.log 2>&1; echo "exit=$?" >> $O/$name.log
    tail -${TAIL:-60} $O/$name.log
    ;;
torchrun)   # N ranks of any script: name, N, then the script and its arguments
    name=$1; n=$2; shift 2
    timeout ${TORCHRUN_TIMEOUT:-900} python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 "$@" > $O/$name.log 2>&1; echo "exit=$?" >> $O/$name.log
    grep -v "^\[W\|^W[0-9]\|OMP_NUM_THREADS\|^\*\*\*" $O/$name.log | tail -${TAIL:-80} | cut -c1-${CUT:-600}
    ;;
multi)      # several jobs in one session:  multi "job1 args" "job2 args" ...
    for j in "$@"; do echo "=== $j"; bash scripts/gpu_job.sh $j; done
    ;;
*)  echo "unknown job $job"; exit 2 ;;
esac
