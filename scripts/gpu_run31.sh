#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cluster_pairs or rank_sharding" > gpurun_out/tests31.log 2>&1; echo "pytest exit=$?" >> gpurun_out/tests31.log
tail -15 gpurun_out/tests31.log
L=gpurun_out/probe31.log
: > $L
run() { echo "=== FNB_DEBUG=$FNB_DEBUG $*" >> $L; timeout 200 python scripts/gpu_probe.py "$@" >> $L 2>&1; echo "exit=$?" >> $L; }
export FNB_DEBUG=0
for pairs in 2 4 2 4; do
  run bench fp16f8 2 1000000 512 3 $pairs 32768
done
run bench fp16f8 2 1000000 512 3 4 65536
run bench fp16x3 2 1000000 512 3 4 32768
run bench fp16f8 2 400000 512 4 4 32768
run bench fp16f8 2 100000 512 4 4 16384
python - <<'PY'
import re
h=None; res={}
for l in open('gpurun_out/probe31.log'):
    if l.startswith('==='): h=l.strip()[4:]
    m=re.search(r'grid=(\d+).*-> ([\d.]+) Gpairs',l)
    if m: res.setdefault(h,[]).append((int(m.group(1)), float(m.group(2))))
    if 'Error' in l or 'error' in l: print(h, l.strip()[:200])
for k,v in res.items(): print('%-60s'%k, v[1:])
PY
