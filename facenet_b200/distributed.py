"""Multi-GPU form of the whole-set verification histogram: one process per GPU (torch.distributed,
NCCL over NVLink/NVSwitch for the plumbing).

The pair matrix shards naturally: every rank needs all N embeddings (columns) but computes only the
tiles of the row blocks ``rb % world == rank`` of every super-row of the tile order, so each rank does 1/world of the upper
triangle regardless of N (no triangle imbalance).  Exchange steps:

  1. all-gather of the embedding / label shards (N * 512 * 4 bytes in total; 2 GB at N = 1M);
  2. all-reduce (sum) of the [2, T+1] int64 histogram bins -- integer, so the result is identical for
     every world size.

Both steps run inside the library (``fnb_comm_init`` / ``fnb_pair_histogram_sharded``, NCCL behind the C ABI) -- a host
without torch can shard too; ``torch.distributed`` only carries the 128-byte communicator id here.  ``hist_fn`` replaces the
per-rank compute call: the CPU tests (gloo, world_size 2) inject a NumPy stand-in and ``torch.distributed`` then does the
all-gather and the all-reduce around it.
"""
import numpy as np


class ShardBalancer:
    """Shares of the pair matrix per rank, adapted to the measured speed of each GPU.

    Every rank computes the tiles of the row blocks ``rb`` (per super-row of the tile order) whose residue ``rb % mod`` it
    owns (``fnb_options.shard_*``); the ranks' residue sets partition ``[0, mod)`` and their sizes are the shares.  The all-reduce at the
    end of a step makes everybody wait for the slowest GPU, and under the board power limit the GPUs of one box differ
    by several per cent; after each step the ranks exchange their Gram-kernel times (one tiny all-reduce) and the widths
    move half-way towards ``speed / sum(speed)``.  The integer bins do not depend on the split, so adapting it never
    changes a result.  All ranks hold identical state (it is updated from all-reduced values only)."""

    def __init__(self, world, slots_per_rank=64):
        self.world = int(world)
        self.mod = self.world * int(slots_per_rank)
        self.widths = [int(slots_per_rank)] * self.world
        self.steps = 0
        self.gain = 0.5            # how far the shares move towards speed / sum(speed) per step
        self._table = None
        self._pending = None       # all-reduced kernel times of the previous step, read at the start of the next one

    def owners(self):
        """Owner of each residue in [0, mod): rank r's k-th residue sits at position (k + 1/2) / width_r of the unit interval
        and the residues are handed out in the order of those positions, so every rank's residues are spread evenly -- the
        long rows at the top and the short rows at the bottom of a triangular region are shared out in proportion, and
        equal widths give the plain ``residue % world``."""
        if self._table is None:
            pos = np.concatenate([(np.arange(w) + 0.5) / w for w in self.widths])
            who = np.concatenate([np.full(w, r, dtype=np.int64) for r, w in enumerate(self.widths)])
            self._table = who[np.lexsort((who, pos))]
        return self._table

    def spec(self, rank):
        """``(mod, residues)`` of ``rank`` -- the ``shard`` argument of ``Handle.pair_histogram_bins``."""
        return self.mod, np.flatnonzero(self.owners() == rank).tolist()

    def update(self, kernel_ms):
        ms = [float(v) for v in kernel_ms]
        if len(ms) != self.world or min(ms) <= 0.0:
            return
        speed = [w / t for w, t in zip(self.widths, ms)]
        total = sum(speed)
        g = self.gain
        target = [(1.0 - g) * w + g * self.mod * sp / total for w, sp in zip(self.widths, speed)]
        widths = [max(1, int(v)) for v in target]
        # largest remainders first until the widths fill [0, mod) again
        order = sorted(range(self.world), key=lambda r: (-(target[r] - int(target[r])), r))
        i = 0
        while sum(widths) < self.mod:
            widths[order[i % self.world]] += 1
            i += 1
        while sum(widths) > self.mod:
            r = max(range(self.world), key=lambda q: (widths[q], -q))
            widths[r] -= 1
        self.widths = widths
        self._table = None
        self.steps += 1


_balancers = {}


def default_balancer(world, group=None):
    """Process-wide balancer per (group, world)."""
    key = (id(group) if group is not None else None, int(world))
    if key not in _balancers:
        _balancers[key] = ShardBalancer(world)
    return _balancers[key]


def _default_hist_fn(device_index):
    from facenet_b200 import _capi
    handle = _capi.default_handle(device_index)

    def fn(emb, labels, thresholds, metric, rank, world, bins_out, **kw):
        import torch
        handle.set_stream(torch.cuda.current_stream().cuda_stream)
        _, stats = handle.pair_histogram_bins(emb, labels, thresholds, metric, rank=rank, world=world,
                                              bins_out=bins_out, **kw)
        return stats
    return fn


def gather_shards(emb_shard, labels_shard, group=None):
    """all-gather equally sized shards into the full [N, D] / [N] tensors (same on every rank)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return emb_shard, labels_shard
    emb = torch.empty((emb_shard.shape[0] * world, emb_shard.shape[1]), dtype=emb_shard.dtype, device=emb_shard.device)
    labels = torch.empty((labels_shard.shape[0] * world,), dtype=labels_shard.dtype, device=labels_shard.device)
    dist.all_gather_into_tensor(emb, emb_shard.contiguous(), group=group)
    dist.all_gather_into_tensor(labels, labels_shard.contiguous(), group=group)
    return emb, labels


_native = {}


def native_handle(device_index, group=None):
    """The library handle of this process's GPU with an NCCL communicator over ``group`` (made on first use: rank 0's
    ``ncclUniqueId`` travels through the group's own broadcast; the communicator itself lives behind the C ABI)."""
    import torch.distributed as dist
    from facenet_b200 import _capi
    key = (int(device_index), id(group) if group is not None else None)
    if key not in _native:
        h = _capi.default_handle(int(device_index))
        try:
            if getattr(h, 'comm', None) is None or h.comm[1] != dist.get_world_size(group):
                h.comm_init_from_torch(group)
        except Exception as exc:      # NCCL not loadable / communicator refused: the same on every rank (same process image)
            import warnings
            warnings.warn('facenet_b200: no NCCL communicator inside the library (%s); torch.distributed does the exchange' % exc)
            h = None
        _native[key] = h
    return _native[key]


def pair_histogram_sharded(emb_shard, labels_shard, thresholds, metric=0, group=None, hist_fn=None, balancer=None, **kw):
    """Every rank passes its shard (torch tensors on its device, or host arrays); returns on every rank
    ``(bins [2, T+1] int64 torch tensor summed over ranks, stats of this rank)``.

    Default (``hist_fn=None``, an initialised process group, a GPU): the exchange and the reduction run INSIDE the library
    (``fnb_pair_histogram_sharded``: labels first, rows by ncclBroadcast from their owner on a side stream -- chunk by chunk
    under the Gram launches when the shards are in class order -- and an ncclAllReduce of the integer bins).
    ``hist_fn`` given (the CPU tests inject a NumPy stand-in): ``torch.distributed`` does the all-gather and the all-reduce
    around it.  An error on one rank (embeddings not normalised, ...) is raised on EVERY rank -- never a hang in the all-reduce.

    ``balancer`` (a ``ShardBalancer``, e.g. ``default_balancer(world)``): split the work by measured GPU speed and keep
    adapting; ``None`` = equal shares."""
    import torch
    import torch.distributed as dist
    rank, world = (dist.get_rank(group), dist.get_world_size(group)) if dist.is_initialized() else (0, 1)
    thr = np.ascontiguousarray(np.atleast_1d(thresholds), dtype=np.float64)
    if balancer is not None and world > 1:
        if balancer._pending is not None:
            # times of the previous step: that all-reduce finished long ago, reading it now does not stall the GPU
            balancer.update(balancer._pending.tolist())
            balancer._pending = None
        kw = dict(kw, shard=balancer.spec(rank))
    if hist_fn is None and world > 1 and torch.cuda.is_available():
        dev = emb_shard.device if getattr(emb_shard, 'is_cuda', False) else torch.device('cuda', torch.cuda.current_device())
        handle = native_handle(dev.index or 0, group)
    else:
        handle = None
    if handle is not None:
        bins = torch.zeros((2, thr.size + 1), dtype=torch.int64, device=dev)
        from facenet_b200 import _capi
        try:
            _, stats = handle.pair_histogram_sharded(emb_shard, labels_shard, thr, metric, bins_out=bins, **kw)
        except _capi.FnbError as e:
            if e.code == _capi.FNB_ERR_NOT_NORMALIZED:       # the reference's exception (statistics.py:40-42), on every rank
                raise ValueError(str(e)) from None
            raise
        _exchange_times(balancer, stats, rank, world, dev, group)
        return bins, stats
    if isinstance(emb_shard, np.ndarray) and hist_fn is None and torch.cuda.is_available():      # (the library path takes host rows as they are)
        emb_shard, labels_shard = torch.from_numpy(emb_shard).cuda(), torch.from_numpy(np.asarray(labels_shard)).cuda()
    emb, labels = gather_shards(emb_shard, labels_shard, group)
    # one spare slot per row carries the ranks' error flags through the same all-reduce as the bins
    buf = torch.zeros((2, thr.size + 2), dtype=torch.int64, device=emb.device)
    bins = buf[:, :thr.size + 1]
    if hist_fn is None:
        hist_fn = _default_hist_fn(emb.device.index or 0)
    stats, err = None, None
    try:
        local = torch.zeros((2, thr.size + 1), dtype=torch.int64, device=emb.device)
        stats = hist_fn(emb, labels, thr, metric, rank, world, local, **kw)
        bins.copy_(local)
    except Exception as e:          # raised below, on every rank
        err = e
        buf[0 if isinstance(e, ValueError) else 1, thr.size + 1] = 1
    if world > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    flags = buf[:, thr.size + 1].tolist() if (world > 1 or err is not None) else (0, 0)
    if err is not None:
        raise err
    if flags[0]:
        raise ValueError('embeddings must be normalized to 1 (reported by another rank)')
    if flags[1]:
        raise RuntimeError('pair_histogram_sharded failed on another rank')
    _exchange_times(balancer, stats, rank, world, emb.device, group)
    return bins.contiguous(), stats


def _exchange_times(balancer, stats, rank, world, device, group):
    import torch
    import torch.distributed as dist
    if world > 1 and balancer is not None and isinstance(stats, dict) and 'kernel_ms' in stats:
        times = torch.zeros(world, dtype=torch.float64, device=device)
        times[rank] = float(stats['kernel_ms'])
        dist.all_reduce(times, op=dist.ReduceOp.SUM, group=group)
        balancer._pending = times


def counts_from_bins(bins, thresholds, metric=0):
    """Host conversion of the summed bins into per-threshold counts (same arithmetic as
    ``fnb_counts_from_bins``): dict with ``same``, ``diff`` (int64 [T]), ``n_same``, ``n_diff``."""
    from facenet_b200 import _capi
    from facenet_b200.statistics import _counts_lt
    b = np.asarray(bins.cpu() if hasattr(bins, 'cpu') else bins).astype(np.int64)
    cuts = _capi.numpy_cuts(thresholds, metric)
    lt = _counts_lt(b, cuts)
    return {'same': lt[1], 'diff': lt[0] - lt[1], 'n_same': int(b[1].sum()), 'n_diff': int(b[0].sum() - b[1].sum())}
