"""bench.py --impl reference (the CPU arm the driver runs beside the GPU arm): one JSON line with the contract's keys."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, str(ROOT / 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1',
                          '--cpu-sample-rows', '2048'], capture_output=True, text=True, timeout=300, cwd=str(ROOT))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'G pair-distances/s' and d['higher_is_better'] is True
    assert d['metric'].startswith('G pair-distances/sec') and d['value'] > 0 and d['n_gpus'] == 1
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == {'value': d['value'], 'unit': d['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in d['config'] and 'sample' in d['config']


def test_reference_arm_other_ranks_stay_silent():
    import os
    env = dict(os.environ, RANK='1', WORLD_SIZE='2')
    out = subprocess.run([sys.executable, str(ROOT / 'bench.py'), '--impl', 'reference', '--gpus', '2'], capture_output=True, text=True,
                         timeout=120, cwd=str(ROOT), env=env)
    assert out.returncode == 0 and out.stdout.strip() == ''
