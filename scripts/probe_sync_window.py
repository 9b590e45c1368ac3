"""Cluster-progress window (GramParams::sync_window, FNB_DEBUG bits 2-4) against the free-running schedule, ONE process
(for a single-pass ncu metrics run): python scripts/probe_sync_window.py N pairs rr w0 w1 ...   (w = 0: off)
Prints the kernel time per launch and whether the bins equal those of the first launch (the window changes timing only)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from facenet_b200 import _capi

n, pairs, rr = (int(v) for v in sys.argv[1:4])
windows = [int(v) for v in sys.argv[4:]]
h = _capi.default_handle(0)
g = torch.Generator(device='cuda'); g.manual_seed(0)
ids = n // 50
labels = (torch.arange(n, device='cuda') % ids)[torch.randperm(n, generator=g, device='cuda')]
x = torch.randn((ids, 512), generator=g, device='cuda')[labels] + 1.1 * torch.randn((n, 512), generator=g, device='cuda')
x = (x / x.norm(dim=1, keepdim=True)).contiguous()
thr = np.linspace(0, 4, 100)
first = None
for w in windows:
    os.environ['FNB_DEBUG'] = str(w << 2)
    bins, st = h.pair_histogram_bins(x, labels, thr, 0, mode='fp16f8', cluster_pairs=pairs, region_rows=rr)
    if first is None:
        first = bins.copy()
    print('window=%d rr=%d grid=%d kernel %.1f ms bins_equal=%s' % (w, rr, st['grid_ctas'], st['kernel_ms'], bool((bins == first).all())), flush=True)
