"""Super-row height x cluster-progress window in ONE process (for a single-pass ncu metrics run):
python scripts/probe_rr_window.py N pairs rr0,rr1,... w0,w1,...     (w = -1: off; the first launch is cold: repeat it)
Prints the kernel time per launch and whether the bins equal those of the first launch."""
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from facenet_b200 import _capi

n, pairs = int(sys.argv[1]), int(sys.argv[2])
rrs = [int(v) for v in sys.argv[3].split(',')]
windows = [int(v) for v in sys.argv[4].split(',')]
h = _capi.default_handle(0)
g = torch.Generator(device='cuda'); g.manual_seed(0)
ids = n // 50
labels = (torch.arange(n, device='cuda') % ids)[torch.randperm(n, generator=g, device='cuda')]
x = torch.randn((ids, 512), generator=g, device='cuda')[labels] + 1.1 * torch.randn((n, 512), generator=g, device='cuda')
x = (x / x.norm(dim=1, keepdim=True)).contiguous()
thr = np.linspace(0, 4, 100)
first = None
for rr in rrs:
    for w in windows:
        bins, st = h.pair_histogram_bins(x, labels, thr, 0, mode='fp16f8', cluster_pairs=pairs, region_rows=rr, panel_window=w)
        if first is None:
            first = bins.copy()
        print('rr=%d window=%d (used %d) grid=%d kernel %.1f ms bins_equal=%s' %
              (rr, w, st['panel_window'], st['grid_ctas'], st['kernel_ms'], bool((bins == first).all())), flush=True)
