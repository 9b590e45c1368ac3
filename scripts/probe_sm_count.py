"""Is the 1M launch bound by the number of SMs or by the board power limit?  The same launch on fewer CTAs (fnb_options.max_ctas)
and with one-pair clusters on all 148 SMs: kernel time (CUDA events inside the library), SM clock and power sampled meanwhile."""
import subprocess
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, '.')
from facenet_b200 import _capi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
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


for pairs, ctas in ((2, 0), (2, 116), (2, 100), (2, 64), (1, 0), (1, 132), (1, 116), (2, 0)):
    h.pair_histogram_bins(x, labels, thr, 0, mode='auto', cluster_pairs=pairs, max_ctas=ctas)      # warm
    del samples[:]
    stop = False
    t = threading.Thread(target=sampler); t.start()
    ms = []
    for _ in range(3):
        _, st = h.pair_histogram_bins(x, labels, thr, 0, mode='auto', cluster_pairs=pairs, max_ctas=ctas)
        ms.append(st['kernel_ms'])
    stop = True; t.join()
    clk = np.median([s[0] for s in samples]) if samples else 0
    pw = np.median([s[1] for s in samples]) if samples else 0
    k = float(np.mean(ms))
    print('cluster_pairs %d max_ctas %3d -> grid %3d: kernel %.1f ms (%.1f G pairs/s), per-CTA rate %.3f G pairs/s, SM clock %.0f MHz, power %.0f W'
          % (pairs, ctas, st['grid_ctas'], k, n * (n - 1) / 2 / k / 1e6, n * (n - 1) / 2 / k / 1e6 / st['grid_ctas'], clk, pw), flush=True)
