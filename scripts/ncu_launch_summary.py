"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list:  python scripts/ncu_launch_summary.py file.csv"""
import csv
import re
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1], newline='') as f:
    lines = [ln for ln in f if not ln.startswith('==')]
for r in csv.DictReader(lines):
    if r.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(r['Metric Value'].replace(',', ''))
    unit = r.get('Metric Unit', 'ns')
    ns = v * {'ns': 1, 'us': 1e3, 'usecond': 1e3, 'ms': 1e6, 'msecond': 1e6, 'nsecond': 1, 's': 1e9, 'second': 1e9}.get(unit, 1)
    name = re.sub(r'\(.*', '', r['Kernel Name'])
    name = re.sub(r'<.*', '', name) if len(name) > 60 else name
    rows.append((name, ns))
tot = OrderedDict()
for name, ns in rows:
    c, t = tot.get(name, (0, 0.0))
    tot[name] = (c + 1, t + ns)
total = sum(t for _, t in tot.values())
print('%d launches, %.3f ms in kernels' % (len(rows), total / 1e6))
for name, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print('%-60s x%-5d total %10.3f ms  avg %9.2f us  share %5.1f %%' % (name[:60], c, t / 1e6, t / c / 1e3, 100 * t / total))
