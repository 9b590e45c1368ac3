"""Two super-row heights back to back in ONE process (for a single-pass ncu metrics run): python scripts/probe_rr_pair.py N pairs rrA rrB"""
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from facenet_b200 import _capi

n, pairs, rra, rrb = (int(v) for v in sys.argv[1:5])
h = _capi.default_handle(0)
g = torch.Generator(device='cuda'); g.manual_seed(0)
ids = n // 50
labels = (torch.arange(n, device='cuda') % ids)[torch.randperm(n, generator=g, device='cuda')]
x = torch.randn((ids, 512), generator=g, device='cuda')[labels] + 1.1 * torch.randn((n, 512), generator=g, device='cuda')
x = (x / x.norm(dim=1, keepdim=True)).contiguous()
thr = np.linspace(0, 4, 100)
for rr in (rra, rrb, rra, rrb):
    _, st = h.pair_histogram_bins(x, labels, thr, 0, mode='fp16f8', cluster_pairs=pairs, region_rows=rr)
    print('rr=%d grid=%d kernel %.1f ms' % (rr, st['grid_ctas'], st['kernel_ms']))
