#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
L=gpurun_out/probe29.log
: > $L
for args in "1000000 8 0 0" "1000000 4 0 0"; do
  timeout 300 python scripts/check_world_emulated.py $args >> $L 2>&1; echo "exit=$?" >> $L
done
cat $L | cut -c1-700
