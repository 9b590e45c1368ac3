"""One GPU plays every rank of an N-GPU job in turn: the ranks' integer bins must sum to the single-rank bins.
    python scripts/check_world_emulated.py N world [cluster_pairs] [region_rows] [mode]"""
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from facenet_b200 import _capi

n, world = int(sys.argv[1]), int(sys.argv[2])
pairs = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rr = int(sys.argv[4]) if len(sys.argv) > 4 else 0
mode = sys.argv[5] if len(sys.argv) > 5 else 'auto'
h = _capi.default_handle(0)
g = torch.Generator(device='cuda'); g.manual_seed(0)
ids = n // 50
labels = (torch.arange(n, device='cuda') % ids)[torch.randperm(n, generator=g, device='cuda')]
x = torch.randn((ids, 512), generator=g, device='cuda')[labels] + 1.1 * torch.randn((n, 512), generator=g, device='cuda')
x = (x / x.norm(dim=1, keepdim=True)).contiguous()
thr = np.linspace(0, 4, 100)
whole, st = h.pair_histogram_bins(x, labels, thr, 0, mode=mode, cluster_pairs=1, region_rows=32768)
total = np.zeros_like(whole)
per_rank = []
for r in range(world):
    b, s = h.pair_histogram_bins(x, labels, thr, 0, mode=mode, rank=r, world=world, cluster_pairs=pairs, region_rows=rr)
    total += b
    per_rank.append((int(b[0].sum()), s['tiles'], s['grid_ctas'], round(s['kernel_ms'], 2)))
want = n * (n - 1) // 2
print('N=%d world=%d pairs=%d rr=%d: whole %d (want %d) sum over ranks %d  equal bins: %s' %
      (n, world, pairs, rr, int(whole[0].sum()), want, int(total[0].sum()), bool((total == whole).all())))
print('   per rank (pairs, tiles, grid, ms):', per_rank)
