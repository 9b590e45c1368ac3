"""Two-rank form of the sharded entry point on a box with at least two GPUs (skipped on one): scripts/check_multi_gpu.py under
torchrun -- `fnb_comm_init` + `fnb_pair_histogram_sharded` (NCCL broadcasts chunk by chunk under the Gram launches, tile queues
shared over NVLink, ncclAllReduce of the bins) must equal the single-GPU bins on every rank for class-ordered and shuffled rows,
device / pageable / pinned shards, ragged shards, three modes, streaming on / off; un-normalised rows on one rank must raise on
all.  The 2- and 8-GPU logs of the round are profiles/r02n_check_n2.log and r02n_check_n8.log."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu


def test_sharded_histogram_two_ranks():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    root = Path(__file__).resolve().parent.parent
    env = dict(os.environ, MASTER_ADDR='127.0.0.1')
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2', '--master-addr', '127.0.0.1',
                        '--master-port', '29531', str(root / 'scripts' / 'check_multi_gpu.py')], cwd=root, env=env, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert 'all checks passed' in r.stdout
