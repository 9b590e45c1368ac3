"""Super-row height under the tile queue (1M x 512, auto -> fp16f8, queue + auxiliary pairs): kernel time per height in one
process.  With the queue a row block is read by whichever cluster asks (both dies), so the row panels compete for L2 twice."""
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from facenet_b200 import _capi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
heights = [int(v) for v in (sys.argv[2].split(',') if len(sys.argv) > 2 else '0,16896,25344,33792,50688'.split(','))]
mode = sys.argv[3] if len(sys.argv) > 3 else 'auto'
dev = torch.device('cuda', 0)
gen = torch.Generator(device=dev); gen.manual_seed(0)
ids = n // 50
centres = torch.randn((ids, 512), generator=gen, device=dev)
labels = (torch.arange(n, device=dev) % ids)[torch.randperm(n, generator=gen, device=dev)]
x = torch.empty((n, 512), device=dev)
for c0 in range(0, n, 1 << 17):
    c1 = min(n, c0 + (1 << 17))
    blk = centres[labels[c0:c1]] + 1.1 * torch.randn((c1 - c0, 512), generator=gen, device=dev)
    x[c0:c1] = blk / blk.norm(dim=1, keepdim=True)
del centres, blk
thr = np.linspace(0, 4, 100)
h = _capi.default_handle(0)
ref = None
for rep in range(2):
    for rr in heights:
        bins, _ = h.pair_histogram_bins(x, labels, thr, 0, mode=mode, region_rows=rr)
        ref = bins if ref is None else ref
        ms = []
        for _ in range(2):
            _, st = h.pair_histogram_bins(x, labels, thr, 0, mode=mode, region_rows=rr)
            ms.append(st['kernel_ms'])
        k = float(np.mean(ms))
        print('region_rows %6d grid %d: kernel %.2f ms (%.1f G pairs/s), bins identical %s' % (rr, st['grid_ctas'], k, n * (n - 1) / 2 / k / 1e6, bool((bins == ref).all())), flush=True)
