"""fnb_options.panel_window (cluster-progress window of the Gram kernel) on the GPU: the bins must not depend on it, for whole
sets and for row-block shards, and the auto rule must switch it on for the 1M launch.  python scripts/check_panel_window.py [N_big]"""
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from facenet_b200 import _capi

h = _capi.default_handle(0)
thr = np.linspace(0, 4, 100)


def make(n, seed):
    g = torch.Generator(device='cuda'); g.manual_seed(seed)
    ids = max(1, n // 50)
    labels = (torch.arange(n, device='cuda') % ids)[torch.randperm(n, generator=g, device='cuda')]
    x = torch.randn((ids, 512), generator=g, device='cuda')[labels] + 1.1 * torch.randn((n, 512), generator=g, device='cuda')
    return (x / x.norm(dim=1, keepdim=True)).contiguous(), labels


# small set: forced windows, whole set and two row-block shards
x, lab = make(6000, 1)
base, st = h.pair_histogram_bins(x, lab, thr, 0, panel_window=-1)
assert st['panel_window'] == 0
for w in (1, 2, 7):
    b, st = h.pair_histogram_bins(x, lab, thr, 0, panel_window=w)
    assert st['panel_window'] == w and (b == base).all(), ('whole set', w)
parts = [h.pair_histogram_bins(x, lab, thr, 0, panel_window=2, rank=r, world=2)[0] for r in range(2)]
assert ((parts[0] + parts[1]) == base).all(), 'shards'
b, st = h.pair_histogram_bins(x, lab, thr, 0)
assert st['panel_window'] == 0 and (b == base).all(), 'auto must stay off for a small set'
print('small set ok (windows 1/2/7, 2 shards, auto = off)', flush=True)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
x, lab = make(n, 0)
first = None
for w in (-1, 0, -1, 0):
    b, st = h.pair_histogram_bins(x, lab, thr, 0, mode='auto', panel_window=w)
    first = b.copy() if first is None else first
    print('N=%d option=%d window_used=%d grid=%d kernel %.1f ms bins_equal=%s' %
          (n, w, st['panel_window'], st['grid_ctas'], st['kernel_ms'], bool((b == first).all())), flush=True)
