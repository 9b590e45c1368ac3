#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
L=gpurun_out/probe32.log
: > $L
for args in "900000 4 330000 28 1" "900000 4 400000 28 2" "900000 2 260000 16 1"; do
  echo "=== $args" >> $L
  timeout 200 python scripts/probe_concurrent.py $args >> $L 2>&1; echo "exit=$?" >> $L
done
cat $L | cut -c1-250
