"""BASELINE config 5: VGGFace2-scale 3.3M x 512 embeddings (~9k identities, ragged), all-pairs verification in BF16
mode, reporting the eps-window disagreements against the fp32-equivalent result.  Any world size:

    python scripts/bench_c5.py [--n 3300000] [--ids 9000] [--ref-mode auto]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/bench_c5.py

Prints one JSON line (rank 0): seconds and G pair-distances/s of the BF16 pass and of the reference-precision pass,
the L1 difference of the per-threshold counts, and how many pairs each pass counted inside the eps window.
The BF16 inputs are the fp32 embeddings rounded to nearest-even (split_rows_kernel), as SURVEY.md section 8(d) states.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, '.')
from facenet_b200 import _capi, distributed as fd

ap = argparse.ArgumentParser()
ap.add_argument('--n', type=int, default=3_300_000)
ap.add_argument('--ids', type=int, default=9000)
ap.add_argument('--ref-mode', default='auto')
ap.add_argument('--reps', type=int, default=2)
args = ap.parse_args()

rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    dist.init_process_group('nccl', device_id=dev)
n = args.n - args.n % world
per = n // world

# ragged class sizes ~75..800 (models/20200724-231357/logs/report.txt:5-11 gives the flavour), same on every rank
rng = np.random.default_rng(0)
w = rng.uniform(75, 800, size=args.ids)
sizes = np.maximum(2, np.floor(w / w.sum() * n)).astype(np.int64)
sizes[0] += n - sizes.sum()
labels_np = np.repeat(np.arange(args.ids, dtype=np.int64), sizes)
labels_np = labels_np[rng.permutation(n)]
labels_full = torch.from_numpy(labels_np).to(dev)
gen = torch.Generator(device=dev)
gen.manual_seed(0)
centres = torch.randn((args.ids, 512), generator=gen, device=dev)
x_shard = torch.empty((per, 512), device=dev)
lo, hi = rank * per, (rank + 1) * per
for c0 in range(0, n, 1 << 17):
    c1 = min(n, c0 + (1 << 17))
    blk = centres[labels_full[c0:c1]] + 1.1 * torch.randn((c1 - c0, 512), generator=gen, device=dev)
    blk = blk / blk.norm(dim=1, keepdim=True)
    a, b = max(c0, lo), min(c1, hi)
    if a < b:
        x_shard[a - lo:b - lo] = blk[a - c0:b - c0]
labels_shard = labels_full[lo:hi].contiguous()
del blk, centres
thr = np.linspace(0, 4, 100)
pairs = n * (n - 1) // 2


def run(mode):
    best, out, st = None, None, None
    for _ in range(args.reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        bins, st = fd.pair_histogram_sharded(x_shard, labels_shard, thr, 0, mode=mode)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        eps = torch.tensor([st['eps_window']], device=dev, dtype=torch.int64)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            dist.all_reduce(eps)
        best = float(dt.item()) if best is None else min(best, float(dt.item()))
        out = bins
    return best, fd.counts_from_bins(out, thr, 0), int(eps.item()), st


t_bf, c_bf, eps_bf, st_bf = run('bf16')
t_ref, c_ref, eps_ref, st_ref = run(args.ref_mode)
sample = None
if rank == 0:
    # per-pair disagreement on a row sample: bin of the BF16 distance vs bin of the fp32-equivalent distance
    ns = min(16384, per)
    h = _capi.default_handle(local)
    xs = x_shard[:ns].contiguous()
    m = ns * (ns - 1) // 2
    d_bf = torch.empty(m, device=dev, dtype=torch.float32)
    d_rf = torch.empty(m, device=dev, dtype=torch.float32)
    h.pairwise(xs, None, 0, mode='bf16', out=d_bf)
    h.pairwise(xs, None, 0, mode='fp16x3', out=d_rf)
    torch.cuda.synchronize()
    thr_t = torch.from_numpy(thr).to(dev)
    b_bf = torch.searchsorted(thr_t, d_bf.double(), right=True)
    b_rf = torch.searchsorted(thr_t, d_rf.double(), right=True)
    differ = b_bf != b_rf
    # distance of the fp32-equivalent value to its nearest threshold
    gap = (d_rf.double()[:, None] - thr_t[torch.clamp(torch.stack([b_rf - 1, b_rf], 1), 0, thr.size - 1)]).abs().min(1).values
    err = (d_bf - d_rf).abs()
    sample = {'rows': ns, 'pairs': m, 'pairs_binned_differently': int(differ.sum().item()),
              'of_those_within_1e-5_of_a_threshold': int((differ & (gap <= 1e-5)).sum().item()),
              'of_those_within_bf16_error_bound_3e-3': int((differ & (gap <= 3e-3)).sum().item()),
              'max_abs_dd': float(err.max().item()), 'rms_dd': float(err.pow(2).mean().sqrt().item())}
    del d_bf, d_rf, b_bf, b_rf, differ, gap, err
if rank == 0:
    l1 = int(np.abs(c_bf['same'] - c_ref['same']).sum() + np.abs(c_bf['diff'] - c_ref['diff']).sum())
    assert c_bf['n_same'] == c_ref['n_same'] and c_bf['n_same'] + c_bf['n_diff'] == pairs
    print(json.dumps({
        'config': 'c5: synthetic %d x 512 embeddings, %d identities (ragged), all-pairs verification, %d GPU(s)' % (n, args.ids, world),
        'pairs': pairs,
        'bf16': {'seconds': t_bf, 'g_pair_distances_per_s': pairs / t_bf / 1e9, 'eps_window_pairs': eps_bf,
                 'kernel_ms_this_rank': st_bf['kernel_ms']},
        'reference_precision': {'mode': _capi.MODE_NAMES[st_ref['mode_used']], 'seconds': t_ref,
                                'g_pair_distances_per_s': pairs / t_ref / 1e9, 'eps_window_pairs': eps_ref,
                                'kernel_ms_this_rank': st_ref['kernel_ms']},
        'bf16_vs_reference_precision': {'l1_count_difference_over_100_thresholds': l1,
                                        'relative_to_pairs': l1 / pairs,
                                        'note': 'cumulative counts: one mis-binned pair contributes 1 per threshold it crosses'},
        'per_pair_sample': sample}))
if world > 1:
    dist.destroy_process_group()
