#!/bin/bash
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
L=gpurun_out/probe33.log
: > $L
for rr in 131072 163840 196608 262144 327680; do
  timeout 300 python scripts/check_world_emulated.py 1000000 8 2 $rr >> $L 2>&1; echo "exit=$?" >> $L
done
python - <<'PY'
import re
for l in open('gpurun_out/probe33.log'):
    if l.startswith('N='): print(l.strip()[:110])
    if 'per rank' in l:
        ms=[float(x) for x in re.findall(r', ([\d.]+)\)', l)]
        print('   ms mean %.1f max %.1f min %.1f' % (sum(ms)/len(ms), max(ms), min(ms)))
PY
