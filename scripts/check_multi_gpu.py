"""Run under torchrun on N GPUs: the sharded whole-set histogram (all-gather + tiles t % world == rank + all-reduce)
must equal, bin for bin, the single-GPU histogram of the same data, and the CPU-oracle counts within the eps window."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, '.')
from facenet_b200 import _capi, distributed as fd
from oracle import statistics_oracle as so

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
n_cls = 64 * world
x, labels = so.synthetic_embeddings([37] * n_cls + [1] * (3 * world), dim=512, sigma=1.1, seed=1)
n = x.shape[0] - x.shape[0] % world
x, labels = x[:n], labels[:n]
per = n // world
thr = so.default_thresholds(0)
xs = torch.from_numpy(x[rank * per:(rank + 1) * per]).cuda()
ls = torch.from_numpy(labels[rank * per:(rank + 1) * per]).cuda()
for mode in ('fp16x3', 'auto'):
    bins, st = fd.pair_histogram_sharded(xs, ls, thr, 0, mode=mode)
    whole, st1 = _capi.default_handle(local).pair_histogram_bins(torch.from_numpy(x).cuda(), torch.from_numpy(labels).cuda(), thr, 0, mode=mode)
    same = bool((bins.cpu().numpy().astype(np.uint64) == whole).all())
    if rank == 0:
        out = fd.counts_from_bins(bins, thr, 0)
        ref = so.pair_histogram(x, labels, thr, 0)
        l1 = int(np.abs(out['same'] - ref['same']).sum() + np.abs(out['diff'] - ref['diff']).sum())
        print('world=%d mode=%s N=%d: sharded bins == single-GPU bins: %s; tiles this rank %d of %d; L1 vs oracle %d (eps window %d)'
              % (world, mode, n, same, st['tiles'], st1['tiles'], l1, st1['eps_window']))
    assert same
dist.destroy_process_group()
