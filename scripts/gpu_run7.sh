#!/bin/bash
# bottleneck decomposition with the profiling knob (FNB_DEBUG: 1 = no epilogue, 2 = no operand loads, 3 = both)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
L=gpurun_out/probe7.log
: > $L
run() {
  echo "=== FNB_DEBUG=$FNB_DEBUG $*" >> $L
  nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap --format=csv,noheader,nounits -lms 100 > gpurun_out/clk.tmp &
  SMI=$!
  timeout 300 python scripts/gpu_probe.py "$@" >> $L 2>&1; echo "exit=$?" >> $L
  kill $SMI
  sort -t, -k2 -n -r gpurun_out/clk.tmp | head -3 | tr '\n' ';' >> $L; echo >> $L
}
for dbg in 0 1 2 3; do
  export FNB_DEBUG=$dbg
  run bench fp16x3 2 100000 512 4
  run bench fp16f8 2 100000 512 4
  run bench bf16 2 100000 512 4
  run bench fp16x3 2 400000 512 4
  run bench fp16f8 2 400000 512 4
done
cat $L
