"""Throughput experiment: a 2 x 2-cluster (or two-pair) launch on most SMs and a small one-pair launch on the stranded SMs,
from two handles / two host threads.  Are the kernels co-resident, and what is the combined rate under the power limit?
    python scripts/probe_concurrent.py NA pairsA NB ctasB [pairsB]"""
import sys
import threading
import time

import numpy as np
import torch

sys.path.insert(0, '.')
from facenet_b200 import _capi

na, pa, nb, cb = (int(v) for v in sys.argv[1:5])
pb = int(sys.argv[5]) if len(sys.argv) > 5 else 1
thr = np.linspace(0, 4, 100)


def make(n, seed):
    g = torch.Generator(device='cuda'); g.manual_seed(seed)
    ids = n // 50
    labels = (torch.arange(n, device='cuda') % ids)[torch.randperm(n, generator=g, device='cuda')]
    x = torch.randn((ids, 512), generator=g, device='cuda')[labels] + 1.1 * torch.randn((n, 512), generator=g, device='cuda')
    return (x / x.norm(dim=1, keepdim=True)).contiguous(), labels


xa, la = make(na, 0)
xb, lb = make(nb, 1)
ha, hb = _capi.Handle(0), _capi.Handle(0)
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()
res = {}


def run(tag, h, s, x, l, pairs, ctas):
    with torch.cuda.stream(s):
        t0 = time.perf_counter()
        _, st = h.pair_histogram_bins(x, l, thr, 0, mode='fp16f8', cluster_pairs=pairs, max_ctas=ctas, region_rows=32768)
        res[tag] = (time.perf_counter() - t0, st['kernel_ms'], st['grid_ctas'], st['n_pairs'])


for rep in range(2):
    run('A', ha, sa, xa, la, pa, 0)
    run('B', hb, sb, xb, lb, pb, cb)
print('alone:  A %.1f ms (grid %d, %.1f Gpairs/s)   B %.1f ms (grid %d, %.1f Gpairs/s)' %
      (res['A'][1], res['A'][2], res['A'][3] / res['A'][1] / 1e6, res['B'][1], res['B'][2], res['B'][3] / res['B'][1] / 1e6))
alone = dict(res)
for rep in range(2):
    ta = threading.Thread(target=run, args=('A', ha, sa, xa, la, pa, 0))
    tb = threading.Thread(target=run, args=('B', hb, sb, xb, lb, pb, cb))
    t0 = time.perf_counter()
    ta.start(); time.sleep(0.02); tb.start()
    ta.join(); tb.join()
    wall = time.perf_counter() - t0
    tot = res['A'][3] + res['B'][3]
    print('together: wall %.1f ms  A kernel %.1f ms  B kernel %.1f ms  -> combined %.1f Gpairs/s  (A alone %.1f)' %
          (wall * 1e3, res['A'][1], res['B'][1], tot / wall / 1e9, alone['A'][3] / alone['A'][1] / 1e6))
