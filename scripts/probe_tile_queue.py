"""A/B of the tile queue in one process (same box, same clocks): static schedule (+ window) / queue / queue + second launch on the
free SMs, 1M (or argv[1]) rows, both headline modes; kernel time from the library's CUDA events, clocks and power sampled."""
import subprocess
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, '.')
from facenet_b200 import _capi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
modes = sys.argv[2].split(',') if len(sys.argv) > 2 else ['auto', 'fp16x3']
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
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
samples = []
stop = False


def sampler():
    while not stop:
        out = subprocess.run(['nvidia-smi', '--query-gpu=clocks.sm,power.draw', '--format=csv,noheader,nounits', '-i', '0'],
                             capture_output=True, text=True).stdout.strip().split(',')
        try:
            samples.append((float(out[0]), float(out[1])))
        except Exception:
            pass
        time.sleep(0.05)


ref = {}
for mode in modes:
    for q, pairs in ((-1, 0), (2, 0), (0, 0), (-1, 0), (0, 0), (2, 1), (-1, 1)):
        kw = dict(mode=mode, tile_queue=q, cluster_pairs=pairs)
        bins, _ = h.pair_histogram_bins(x, labels, thr, 0, **kw)      # warm
        same = (bins == ref.setdefault(mode, bins)).all()
        del samples[:]
        stop = False
        t = threading.Thread(target=sampler); t.start()
        ms = []
        for _ in range(reps):
            _, st = h.pair_histogram_bins(x, labels, thr, 0, **kw)
            ms.append(st['kernel_ms'])
        stop = True; t.join()
        clk = np.median([s[0] for s in samples]) if samples else 0
        pw = np.median([s[1] for s in samples]) if samples else 0
        k = float(np.mean(ms))
        print('mode %-6s tile_queue %2d cluster_pairs %d -> grid %3d: kernel %.2f ms (%.1f G pairs/s), SM clock %.0f MHz, power %.0f W, bins identical %s'
              % (mode, q, pairs, st['grid_ctas'], k, n * (n - 1) / 2 / k / 1e6, clk, pw, bool(same)), flush=True)
