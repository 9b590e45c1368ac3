#!/bin/bash
# targeted next-column-panel prefetch (debug bit 6) x region rows
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
L=gpurun_out/probe30.log
: > $L
run() { echo "=== FNB_DEBUG=$FNB_DEBUG $*" >> $L; timeout 200 python scripts/gpu_probe.py "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
for rr in 8192 16384 32768; do
  for dbg in 0 64; do
    export FNB_DEBUG=$dbg
    run bench fp16f8 2 1000000 512 3 2 $rr
  done
done
for rr in 4096 16384; do
  for dbg in 0 64; do
    export FNB_DEBUG=$dbg
    run bench fp16f8 2 100000 512 4 1 $rr
  done
done
export FNB_DEBUG=0
python - <<'PY'
import re
h=None; res={}
for l in open('gpurun_out/probe30.log'):
    if l.startswith('==='): h=l.strip()[4:]
    m=re.search(r'-> ([\d.]+) Gpairs',l)
    if m: res.setdefault(h,[]).append(float(m.group(1)))
for k,v in res.items(): print('%-60s'%k, v[1:])
PY
