"""world_size-2 gloo test of the multi-GPU plumbing (shard all-gather, per-rank tile split, integer
all-reduce, bins -> counts) with the NumPy stand-in as the per-rank compute call."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import statistics_oracle as so


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _emulated_hist(emb, labels, thresholds, metric, rank, world, bins_out, **kw):
    """Per-rank stand-in: rows sorted by label, strict upper triangle in 64x64 tiles; a rank owns the row blocks rb with
    rb % mod in its residue list (mirrors the library's split by row block; default mod = world, the residue `rank`)."""
    from facenet_b200 import _capi
    x = emb.numpy()
    lab = labels.numpy()
    order = np.argsort(lab, kind='stable')
    x, lab = x[order], lab[order]
    cuts = np.sort(_capi.numpy_cuts(thresholds, metric))
    nt = cuts.size
    n = x.shape[0]
    tile = 64
    mod, mine = kw.get('shard') or (world, [rank])
    out = np.zeros((2, nt + 1), dtype=np.int64)
    for r0 in range(0, n, tile):
        if (r0 // tile) % mod not in mine:
            continue
        for c0 in range(r0, n, tile):
            s = np.clip(x[r0:r0 + tile] @ x[c0:c0 + tile].T, -1, 1)
            k = np.searchsorted(cuts, s.ravel(), side='right').reshape(s.shape)
            valid = np.arange(c0, c0 + s.shape[1])[None, :] > np.arange(r0, r0 + s.shape[0])[:, None]
            same = valid & (lab[r0:r0 + tile, None] == lab[None, c0:c0 + tile])
            out[0] += np.bincount(k[valid], minlength=nt + 1)
            out[1] += np.bincount(k[same], minlength=nt + 1)
    bins_out.copy_(torch.from_numpy(out))
    return {'emulated': True, 'kernel_ms': 1.0 + 0.5 * rank}      # rank 1 pretends to be the slower GPU


def _worker(rank, world, port, x, labels, result):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from facenet_b200 import distributed as fd
        per = x.shape[0] // world
        xs = torch.from_numpy(x[rank * per:(rank + 1) * per])
        ls = torch.from_numpy(labels[rank * per:(rank + 1) * per])
        thr = so.default_thresholds(0)
        bins, _ = fd.pair_histogram_sharded(xs, ls, thr, 0, hist_fn=_emulated_hist)
        # speed-weighted shares: same bins, and the slower rank's share shrinks step by step (identically on both ranks)
        bal = fd.ShardBalancer(world, slots_per_rank=8)
        for step in range(3):
            again, _ = fd.pair_histogram_sharded(xs, ls, thr, 0, hist_fn=_emulated_hist, balancer=bal)
            assert torch.equal(again, bins)
        assert sum(bal.widths) == bal.mod and bal.widths[0] > bal.widths[1] >= 1
        out = fd.counts_from_bins(bins, thr, 0)
        if rank == 0:
            result['same'] = out['same']
            result['diff'] = out['diff']
            result['n'] = (out['n_same'], out['n_diff'])
        # every rank holds the same reduced bins
        gathered = [torch.zeros_like(bins) for _ in range(world)]
        dist.all_gather(gathered, bins)
        assert all(torch.equal(g, bins) for g in gathered)

        # an error on ONE rank is raised on EVERY rank (no rank is left waiting in the all-reduce), and the next call works
        def failing(emb, lab, thresholds, metric, rank_, world_, bins_out, **kw):
            if rank_ == 1:
                raise ValueError('embeddings must be normalized to 1, range -1 1.5')
            return _emulated_hist(emb, lab, thresholds, metric, rank_, world_, bins_out, **kw)
        try:
            fd.pair_histogram_sharded(xs, ls, thr, 0, hist_fn=failing)
            raised = None
        except ValueError as e:
            raised = str(e)
        assert raised is not None and 'normalized' in raised
        result['raised_%d' % rank] = raised
        again, _ = fd.pair_histogram_sharded(xs, ls, thr, 0, hist_fn=_emulated_hist)
        assert torch.equal(again, bins)
    finally:
        dist.destroy_process_group()


def test_two_rank_histogram_equals_single_process():
    x, labels = so.synthetic_embeddings([7, 1, 30, 12, 2, 2, 18, 24], dim=64, sigma=1.0, seed=6)
    assert x.shape[0] % 2 == 0
    ref = so.pair_histogram(x, labels, so.default_thresholds(0), 0)
    mgr = mp.Manager()
    result = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, x, labels, result), nprocs=2, join=True)
    assert result['n'] == (ref['n_same'], ref['n_diff'])
    assert 'another rank' in result['raised_0'] and 'range' in result['raised_1']
    assert np.abs(result['same'] - ref['same']).sum() + np.abs(result['diff'] - ref['diff']).sum() <= 2


def test_shard_balancer_shares():
    from facenet_b200 import distributed as fd
    b = fd.ShardBalancer(8)
    assert b.spec(3) == (512, list(range(3, 512, 8)))              # equal shares: residue % world == rank
    ms = [110, 111, 105, 109, 104, 109, 101, 107]
    for _ in range(6):
        b.update(ms)
        # the ranges always partition [0, mod)
        assert sum(b.widths) == b.mod and min(b.widths) >= 1
        owned = [b.spec(r)[1] for r in range(8)]
        assert sorted(sum(owned, [])) == list(range(b.mod)) and [len(o) for o in owned] == b.widths
        # a rank's residues are spread evenly: the largest gap is close to mod / width
        assert all(max(np.diff(o)) <= 2 * b.mod / len(o) for o in owned)
        # pretend every GPU keeps its speed: its time follows its share
        speed = [64 / t for t in [110, 111, 105, 109, 104, 109, 101, 107]]
        ms = [w / s for w, s in zip(b.widths, speed)]
    assert max(ms) / min(ms) < 1.03                       # was 1.10 with equal shares
    b.update([0.0] * 8)                                   # unusable measurement: ignored
    assert sum(b.widths) == b.mod
