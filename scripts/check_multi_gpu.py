"""Run under torchrun on N GPUs: the sharded whole-set histogram computed INSIDE the library (fnb_comm_init +
fnb_pair_histogram_sharded: NCCL broadcasts of the rows chunk by chunk under the Gram launches, ncclAllReduce of the bins) must
equal, bin for bin, the single-GPU histogram of the same data -- for rows in class order and shuffled, device and host shards,
ragged shards, every mode, streaming on and off -- and the CPU-oracle counts within the eps window.  An un-normalised row on
one rank must raise on every rank.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/check_multi_gpu.py
"""
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
big = '--big' in sys.argv
n_cls = (400 if big else 64) * world
thr = so.default_thresholds(0)
handle = _capi.default_handle(local)
fails = 0


def say(*a):
    if rank == 0:
        print(*a, flush=True)


for order in ('class', 'shuffled'):
    x, labels = so.synthetic_embeddings([37] * n_cls + [1] * (3 * world), dim=512, sigma=1.1, seed=1, shuffle=(order == 'shuffled'))
    n = x.shape[0]
    # ragged shards: rank r holds rows [cut[r], cut[r + 1])
    cut = [0] + [int(n * (r + 1) / world) - (17 * (r + 1)) % 29 for r in range(world - 1)] + [n]
    xs_h, ls_h = np.ascontiguousarray(x[cut[rank]:cut[rank + 1]]), np.ascontiguousarray(labels[cut[rank]:cut[rank + 1]])
    xs_d, ls_d = torch.from_numpy(xs_h).cuda(), torch.from_numpy(ls_h).cuda()
    xs_p = torch.from_numpy(xs_h).pin_memory()            # pinned host rows: one DMA of the shard, class-order gather on the device
    ref = so.pair_histogram(x, labels, thr, 0) if rank == 0 else None
    for mode in ('fp16x3', 'auto', 'fp16f8'):
        whole, st1 = handle.pair_histogram_bins(torch.from_numpy(x).cuda(), torch.from_numpy(labels).cuda(), thr, 0, mode=mode)
        for where, (xa, la) in (('device', (xs_d, ls_d)), ('host', (xs_h, ls_h)), ('pinned', (xs_p.numpy(), ls_h))):
            for streamed, rr in ((None, 0), (1, 1024), (-1, 0)) + (((3, 0),) if where == 'pinned' else ()):
                bins, st = fd.pair_histogram_sharded(xa, la, thr, 0, mode=mode, streamed=streamed, region_rows=rr)
                same = bool((bins.cpu().numpy().astype(np.uint64) == whole).all())
                flag = torch.tensor([0 if same else 1], device='cuda')
                dist.all_reduce(flag)
                ok = int(flag.item()) == 0
                fails += 0 if ok else 1
                if rank == 0:
                    out = fd.counts_from_bins(bins, thr, 0)
                    l1 = int(np.abs(out['same'] - ref['same']).sum() + np.abs(out['diff'] - ref['diff']).sum())
                    print('world=%d order=%-8s mode=%-6s shards=%-6s streamed=%-4s N=%d: bins == single-GPU bins on every rank: %s; chunks %d; '
                          'tiles this rank %d of %d; mode_used %s; L1 vs oracle %d (eps window %d); gather %.2f ms kernel %.2f ms'
                          % (world, order, mode, where, streamed, n, ok, st['streamed_chunks'], st['tiles'], st1['tiles'],
                             _capi.MODE_NAMES[st['mode_used']], l1, st1['eps_window'], st['gather_ms'], st['kernel_ms']), flush=True)
    # an un-normalised pair on the LAST rank only: every rank raises the reference's ValueError
    bad = xs_d.clone()
    if rank == world - 1:
        bad[5] = bad[6] * 1.01
    try:
        fd.pair_histogram_sharded(bad, ls_d, thr, 0, mode='auto')
        raised = False
    except ValueError as e:
        raised = 'normalized' in str(e)
    flag = torch.tensor([0 if raised else 1], device='cuda')
    dist.all_reduce(flag)
    say('world=%d order=%s: un-normalised rows on rank %d raise on every rank: %s' % (world, order, world - 1, int(flag.item()) == 0))
    fails += int(flag.item())
    # and the communicator is still usable
    bins, _ = fd.pair_histogram_sharded(xs_d, ls_d, thr, 0, mode='fp16x3')
    whole, _ = handle.pair_histogram_bins(torch.from_numpy(x).cuda(), torch.from_numpy(labels).cuda(), thr, 0, mode='fp16x3')
    fails += 0 if bool((bins.cpu().numpy().astype(np.uint64) == whole).all()) else 1

say('comm', handle.comm_info(), 'FAILURES' if fails else 'all checks passed', fails)
dist.barrier()
dist.destroy_process_group()
sys.exit(1 if fails else 0)
