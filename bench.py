"""Benchmarks of the embedding-evaluation hot path on B200 (BASELINE.json); one JSON line per run.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload 1m|100k|c5|mining|lfw] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workloads (BASELINE.json configs; the default is the one the metric is quoted on):
  1m      config 4 (DEFAULT): all-pairs verification histogram of 1M x 512 embeddings, 4.999995e11 pairs per step
  100k    config 2: the same on 100k x 512
  c5      config 5: 3.3M x 512 (9,000 ragged identities), BF16 pass, with the disagreement report against the
          fp32-equivalent pass in `parity`
  mining  config 3: triplet mining of 1800-row batches (45 identities x 40 images), 10,000 steps, 16 batches per launch
  lfw     config 1: FaceToFaceValidation (10 folds x 100 thresholds + FAR threshold) on 13,233 x 512, 5,749 identities

One step = one pass of the hot path over one batch of synthetic input.  For the all-pairs workloads:
  value  device-resident inputs (each rank holds its row shard in HBM), CUDA events around K steps,
         max over ranks; per step: [all-gather of shards] -> sort/split -> Gram+histogram -> [all-reduce]
  e2e    the same metric through the public API with HOST buffers: H2D of the step's embeddings and labels and D2H of the
         histogram inside the timed region; `e2e` = pinned host memory (bench contract), `e2e_pageable` = what the
         reference's caller hands over (np.concatenate output), both with the measured h2d_ms
  strict the same workload in the strict fp16x3 contraction (a few steps), beside the headline mode
  roofline   algorithmic FLOP (1024 per pair) / Gram-kernel time, against the measured tensor peak
  cpu_baseline  the NumPy oracle on the box's host cores on a bounded sample (rank 0, N=1 only)

`--impl reference` times the reference's CPU implementation of the path (the oracle port of
facenet/statistics.py; the reference is pure Python and cannot travel to the GPU box) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    '1m': dict(n=1_000_000, ids=20_000, name='synthetic 1M x 512 unit-norm fp32 embeddings (20,000 ids x 50), all-pairs '
                                             'verification histogram, 100 thresholds linspace(0,4), metric 0'),
    '100k': dict(n=100_000, ids=2_000, name='synthetic 100k x 512 unit-norm fp32 embeddings (2,000 ids x 50), all-pairs '
                                            'verification histogram, 100 thresholds linspace(0,4), metric 0'),
    '20k': dict(n=20_000, ids=400, name='synthetic 20k x 512 (debug size)'),
    'c5': dict(n=3_300_000, ids=9_000, ragged=True, mode='bf16',
               name='synthetic VGGFace2-scale 3.3M x 512 embeddings (9,000 ids, 75-800 images each), all-pairs histogram in BF16 '
                    'mode with the eps-window disagreement report against the fp32-equivalent pass'),
    'mining': dict(name='triplet mining: 1800-image batches (45 identities x 40 images), 512-d, alpha 0.2, hardest + semi-hard '
                        'negatives, 10,000 steps'),
    'lfw': dict(name='synthetic LFW-size 13,233 x 512 fp32 embeddings (5,749 identities, per-identity sigma ~ U(1.5, 3.5): AUC < 1), '
                     'FaceToFaceValidation: all-pairs distance + 10-fold ROC / VAL@FAR'),
}
DIM = 512
FLOP_PER_PAIR = 2 * DIM          # one pair distance = 1024 algorithmic FLOP
METRIC = 'G pair-distances/sec all-pairs 512-d verification'
UNIT = 'G pair-distances/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=None, help='timed steps (default: 5; mining 10000; lfw 3; c5 2)')
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='1m', choices=sorted(WORKLOADS))
    ap.add_argument('--mode', default=None, choices=['auto', 'fp16x3', 'fp16f8', 'tf32x3', 'tf32', 'bf16', 'fp16'])
    ap.add_argument('--cta-group', type=int, default=0)
    ap.add_argument('--cpu-sample-rows', type=int, default=16384)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-parity', action='store_true')
    ap.add_argument('--no-strict', action='store_true', help='skip the strict fp16x3 block of the all-pairs workloads')
    ap.add_argument('--strict-steps', type=int, default=2)
    ap.add_argument('--adaptive-shares', action='store_true', help='N > 1: speed-adaptive shares instead of equal ones')
    ap.add_argument('--panel-window', type=int, default=0,
                    help='fnb_options.panel_window: 0 auto (on for launches of >= 5e10 pairs per rank), -1 off, 1..7 window')
    ap.add_argument('--region-rows', type=int, default=0, help='fnb_options.region_rows (0 = auto)')
    ap.add_argument('--parity-rows', type=int, default=8192)
    ap.add_argument('--mining-batches', type=int, default=40,
                    help='mining: batches per launch (fnb_mine_batched); measured 4 / 8 / 16 / 32 per launch: 0.064 / 0.056 / 0.051 / 0.047 ms per batch')
    args = ap.parse_args()
    if args.steps is None:
        args.steps = {'mining': 10000, 'lfw': 3, 'c5': 2}.get(args.workload, 5)
    if args.mode is None:
        args.mode = WORKLOADS[args.workload].get('mode', 'auto')
    return args


def thresholds():
    return np.linspace(0, 4, 100)        # statistics.py:262


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def synth_numpy(n_rows, ids, seed=0):
    """CPU sample of the workload (same recipe as the device generator: clustered, unit norm, fp32)."""
    rng = np.random.default_rng(seed)
    cls = np.arange(n_rows) % ids
    centres = rng.standard_normal((ids, DIM), dtype=np.float32)
    x = centres[cls] + np.float32(1.1) * rng.standard_normal((cls.size, DIM), dtype=np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    perm = rng.permutation(cls.size)
    return np.ascontiguousarray(x[perm]), cls[perm].astype(np.int64)


def cpu_arm(sample_rows, ids_full, n_full, steps, warmup):
    """Times the oracle port on a bounded sample; returns (G pairs/s, ms per step, description)."""
    from oracle import statistics_oracle as so
    ids = max(2, int(round(ids_full * sample_rows / n_full)))
    ids = max(ids, sample_rows // 50)
    x, labels = synth_numpy(sample_rows, ids, seed=1)
    thr = thresholds()
    cores = host_threads()
    pairs = sample_rows * (sample_rows - 1) // 2
    for _ in range(warmup):
        so.pair_histogram(x[:2048], labels[:2048], thr, 0, threads=cores)
    t0 = time.perf_counter()
    for _ in range(steps):
        out = so.pair_histogram(x, labels, thr, 0, threads=cores)
    dt = (time.perf_counter() - t0) / steps
    assert out['n_same'] + out['n_diff'] == pairs
    sample = ('%d x %d row sample of the workload (%d pairs/step), NumPy oracle statistics_oracle.pair_histogram: blocked '
              'fp32 sgemm + searchsorted binning, %d threads' % (sample_rows, DIM, pairs, cores))
    return pairs / dt / 1e9, dt * 1e3, sample, cores


class ClockSampler:
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
              'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.samples = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(index), '--query-gpu=' + self.FIELDS,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc:
            self.proc.terminate()

    def summary(self, t0, t1):
        sm, mx, reasons, power = [], [], set(), []
        for t, line in self.samples:
            if t < t0 or t > t1:
                continue
            parts = [p.strip() for p in line.split(',')]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), parts[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)), 'reasons': sorted(reasons),
                'samples': len(sm), 'power_w_max': float(max(power))}


def measured_peaks():
    f = ROOT / 'MEASURED_PEAKS.json'
    if f.exists():
        try:
            return json.loads(f.read_text()), 'MEASURED_PEAKS.json'
        except ValueError:
            pass
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback (B200_PROFILING.md)'


def tf32_live(torch, dev):
    """cuBLAS TF32 (torch.matmul fp32 8192^3 with allow_tf32) measured the way MEASURED_PEAKS.json measures bf16:
    best of 5 (burst) and back to back for ~4 s (sustained).  Not on the product path."""
    try:
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        ma = torch.randn((8192, 8192), device=dev)
        mb = torch.randn((8192, 8192), device=dev)
        torch.matmul(ma, mb)
        flop = 2.0 * 8192 ** 3
        best = None
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(ma, mb); e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        reps = max(10, int(4000.0 / best))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            torch.matmul(ma, mb)
        e1.record(); torch.cuda.synchronize()
        sustained = flop * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12
        torch.backends.cuda.matmul.allow_tf32 = prev
        del ma, mb
        return flop / (best * 1e-3) / 1e12, sustained, reps
    except Exception:
        return None, None, 0


# ----------------------------------------------------------------------------------------------------------------------
# reference arm (CPU)

def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    cores = host_threads()
    if args.workload == 'mining':
        from oracle import mining_oracle as mo, statistics_oracle as so
        steps = max(1, min(args.steps, 3))
        batches = [so.synthetic_embeddings([40] * 45, dim=DIM, sigma=1.1, seed=s, shuffle=False) for s in range(steps)]
        t0 = time.perf_counter()
        for x, labels in batches:
            mo.mine(x, labels, 0.2)
        dt = (time.perf_counter() - t0) / steps
        val = 1800 * 1800 / dt / 1e9
        sample = '%d batches of 1800 x 512, NumPy mining oracle (oracle/mining_oracle.py), 1 thread + BLAS' % steps
        kind_cores = 1
    elif args.workload == 'lfw':
        from oracle import statistics_oracle as so
        sizes = so.lfw_like_class_sizes()
        x, labels = so.synthetic_embeddings(sizes, dim=DIM, sigma=(1.5, 3.5), seed=0)
        t0 = time.perf_counter()
        so.face_to_face_validation(x, labels, 0, 10, 1.e-3)
        dt = time.perf_counter() - t0
        val = lfw_pair_evaluations(x.shape[0]) / dt / 1e9
        sample = 'one full validation, vectorised NumPy oracle (the literal reference loop extrapolates to ~37 h)'
        kind_cores = cores
    else:
        val, ms, sample, kind_cores = cpu_arm(args.cpu_sample_rows, wl['ids'], wl['n'], max(1, min(args.steps, 5)), min(args.warmup, 1))
        dt = ms * 1e-3
    line = {'impl': 'reference', 'metric': METRIC, 'value': val,
            'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic', 'config': {'workload': wl['name'], 'sample': sample},
            'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': kind_cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


def lfw_pair_evaluations(n, folds=10):
    """unordered pairs evaluated by one FaceToFaceValidation: every fold's train set over the grid + its test set (statistics.py:284-308)"""
    total = 0
    for f in range(folds):
        nt = n // folds + (1 if f < n % folds else 0)
        ntr = n - nt
        total += ntr * (ntr - 1) // 2 + nt * (nt - 1) // 2
    return total


# ----------------------------------------------------------------------------------------------------------------------
# all-pairs verification histogram (configs 2, 4, 5)

def ragged_sizes(n, ids, lo=75, hi=800):
    """deterministic class sizes in [lo, hi] summing to n"""
    rng = np.random.default_rng(12345)
    sizes = rng.integers(lo, hi + 1, size=ids).astype(np.int64)
    diff = int(n - sizes.sum())
    step = 1 if diff > 0 else -1
    i = 0
    while diff != 0:
        j = i % ids
        if lo <= sizes[j] + step <= hi:
            sizes[j] += step
            diff -= step
        i += 1
    return sizes


def run_allpairs(args):
    import torch
    import torch.distributed as dist
    from facenet_b200 import _capi, distributed as fd, statistics as fst

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    wl = WORKLOADS[args.workload]
    n, ids = wl['n'], wl['ids']
    n -= n % world
    per_rank = n // world
    thr = thresholds()
    pairs = n * (n - 1) // 2

    # ---- synthetic inputs, generated on the device with a fixed seed (identical on every rank), then sharded
    gen = torch.Generator(device=dev)
    gen.manual_seed(0)
    centres = torch.randn((ids, DIM), generator=gen, device=dev, dtype=torch.float32)
    if wl.get('ragged'):
        sizes = torch.from_numpy(ragged_sizes(n, ids)).to(dev)
        labels_full = torch.repeat_interleave(torch.arange(ids, device=dev, dtype=torch.int64), sizes)
        n_same_expected = int((sizes * (sizes - 1) // 2).sum().item())
    else:
        labels_full = torch.arange(n, device=dev, dtype=torch.int64) % ids
        cnt = torch.bincount(labels_full, minlength=ids)
        n_same_expected = int((cnt * (cnt - 1) // 2).sum().item())
    labels_full = labels_full[torch.randperm(n, generator=gen, device=dev)]
    lo, hi = rank * per_rank, (rank + 1) * per_rank
    x_shard = torch.empty((per_rank, DIM), device=dev, dtype=torch.float32)
    chunk = 1 << 17
    for c0 in range(0, n, chunk):                      # same random stream on every rank; keep only the shard
        c1 = min(n, c0 + chunk)
        blk = centres[labels_full[c0:c1]] + 1.1 * torch.randn((c1 - c0, DIM), generator=gen, device=dev, dtype=torch.float32)
        blk = blk / blk.norm(dim=1, keepdim=True)
        a, b = max(c0, lo), min(c1, hi)
        if a < b:
            x_shard[a - lo:b - lo] = blk[a - c0:b - c0]
    del blk
    labels_shard = labels_full[lo:hi].contiguous()
    del labels_full, centres
    torch.cuda.synchronize()

    handle = _capi.default_handle(local_rank)
    fst.set_default_mode(mode=args.mode, device=local_rank, cta_group=args.cta_group)
    stream = torch.cuda.current_stream()
    handle.set_stream(stream.cuda_stream)
    acc = {'kernel_ms': [], 'prepare_ms': [], 'launches': 0, 'modes': set(), 'grids': set(), 'windows': set(), 'bounds': [],
           'fallbacks': 0}
    run_mode = [args.mode]

    def hist_fn(emb, labels, thresholds_, metric, rank_, world_, bins_out, **kw):
        _, st = handle.pair_histogram_bins(emb, labels, thresholds_, metric, rank=rank_, world=world_, bins_out=bins_out,
                                           mode=run_mode[0], cta_group=args.cta_group, shard=kw.get('shard'),
                                           panel_window=args.panel_window, region_rows=args.region_rows)
        record(st)
        return st

    def record(st):
        acc['kernel_ms'].append(st['kernel_ms'])
        acc['prepare_ms'].append(st['prepare_ms'])
        acc['launches'] += st['kernel_launches']
        acc['modes'].add(_capi.MODE_NAMES[st['mode_used']])
        acc['grids'].add(st['grid_ctas'])
        acc['windows'].add(st['panel_window'])
        acc['bounds'].append(st['error_bound'])
        acc['fallbacks'] += st['fallback']
        acc.setdefault('gather_ms', []).append(st.get('gather_ms', 0.0))
        acc.setdefault('chunks', set()).add(st.get('streamed_chunks', 0))

    def reset_acc():
        acc.update(kernel_ms=[], prepare_ms=[], launches=0, bounds=[], fallbacks=0, gather_ms=[])

    # equal shares by default; --adaptive-shares lets the shares follow the measured speed of each GPU (measured on two
    # 8-GPU boxes: the per-rank kernel times level out, but the step is no shorter -- profiles/r01c_multi_gpu.md)
    balancer = fd.default_balancer(world) if (world > 1 and args.adaptive_shares) else None

    def sharded(xs_, ls_):
        # N > 1: the exchange (labels, then rows by ncclBroadcast chunk by chunk under the launches) and the ncclAllReduce of
        # the bins run inside the library (fnb_pair_histogram_sharded); N = 1: the plain call
        if world > 1:
            b, st = fd.pair_histogram_sharded(xs_, ls_, thr, 0, balancer=balancer, mode=run_mode[0], cta_group=args.cta_group,
                                              panel_window=args.panel_window, region_rows=args.region_rows)
            record(st)
            return b, st
        return fd.pair_histogram_sharded(xs_, ls_, thr, 0, hist_fn=hist_fn, balancer=balancer)

    def step_device():
        bins, st = sharded(x_shard, labels_shard)
        return bins.cpu() if rank == 0 else bins      # final histogram on the host (rank 0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(steps):
            out = fn()
        e1.record(stream)
        barrier()
        t1 = time.perf_counter()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), out, t0, t1

    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(args.warmup):
        step_device()
    reset_acc()
    total_ms, bins, t0, t1 = timed(step_device, args.steps)
    k_ms = float(np.mean(acc['kernel_ms'])) if acc['kernel_ms'] else float('nan')
    p_ms = float(np.mean(acc['prepare_ms'])) if acc['prepare_ms'] else float('nan')
    err_bound = float(np.max(acc['bounds'])) if acc['bounds'] else None
    fallbacks = acc['fallbacks']
    # per-rank Gram-kernel time: the all-reduce makes every step wait for the slowest rank
    k_ranks = [k_ms]
    if world > 1:
        kt = torch.tensor([k_ms], device=dev, dtype=torch.float64)
        allk = [torch.zeros_like(kt) for _ in range(world)]
        dist.all_gather(allk, kt)
        k_ranks = [float(t.item()) for t in allk]
    # the row exchange (fnb_stats.gather_ms: first to last piece on the copy stream; all but the first chunk of it runs UNDER the
    # Gram launches) and, for reference, a plain torch all-gather of the same shards timed alone
    gather_ms = float(np.mean(acc['gather_ms'])) if world > 1 and acc.get('gather_ms') else 0.0
    plain_gather_ms = 0.0
    chunks_used = sorted(acc.get('chunks', {0}))
    if world > 1:
        g_total, _, _, _ = timed(lambda: fd.gather_shards(x_shard, labels_shard), args.steps)
        plain_gather_ms = g_total / args.steps
    timed_launches = acc['launches']
    value = pairs * args.steps / (total_ms * 1e-3) / 1e9
    mode_used = sorted(acc['modes'])[0] if len(acc['modes']) == 1 else args.mode      # what 'auto' resolved to
    grids, windows = sorted(acc['grids']), sorted(acc['windows'])

    # sanity of the result (not timed): every pair counted once, same-identity total known analytically
    if rank == 0:
        out = fd.counts_from_bins(bins, thr, 0)
        assert out['n_same'] + out['n_diff'] == pairs, (out['n_same'], out['n_diff'], pairs)
        if world == 1 or n == wl['n']:
            assert out['n_same'] == n_same_expected, (out['n_same'], n_same_expected)
    bins_headline = np.asarray(bins.cpu()).astype(np.int64) if rank == 0 else None

    # ---- the same workload in the strict fp16x3 contraction (every rank takes part; a few steps)
    strict = None
    bins_strict = None
    if not args.no_strict and mode_used != 'fp16x3':
        run_mode[0] = 'fp16x3'
        acc['modes'] = set()
        step_device()
        reset_acc()
        s_ms, s_bins, _, _ = timed(step_device, args.strict_steps)
        s_k = float(np.mean(acc['kernel_ms']))
        strict = {'mode': 'fp16x3', 'steps': args.strict_steps, 'value': pairs * args.strict_steps / (s_ms * 1e-3) / 1e9, 'unit': UNIT,
                  'ms_per_step': s_ms / args.strict_steps, 'kernel_ms': s_k, 'error_bound': float(np.max(acc['bounds']))}
        bins_strict = np.asarray(s_bins.cpu()).astype(np.int64) if rank == 0 else None
        run_mode[0] = args.mode
        acc['modes'] = {mode_used}

    # ---- parity of the timed arithmetic mode on a row sample of THIS workload (not timed; rank 0).  The checker is
    #      plain torch float64 on the GPU (the CPU oracle is the checker in tests/ and smoke()):
    #        max |d_mode - d_f64| over all pairs of the sample           (tolerance of BASELINE.json: 1e-5)
    #        per threshold: |count_mode - count_f64| <= number of pairs whose float64 distance lies within eps of that threshold
    parity = None
    if rank == 0 and not args.no_parity:
        ns = min(args.parity_rows, per_rank)
        xs, ls = x_shard[:ns].contiguous(), labels_shard[:ns].contiguous()

        def sample_report(mode):
            d_mode = torch.empty(ns * (ns - 1) // 2, device=dev, dtype=torch.float32)
            handle.pairwise(xs, None, 0, mode=mode, out=d_mode)
            torch.cuda.synchronize()
            x64 = xs.double()
            iu = torch.triu_indices(ns, ns, 1, device=dev)
            d64 = (2.0 * (1.0 - (x64 @ x64.T).clamp_(-1.0, 1.0)))[iu[0], iu[1]]
            del x64
            err = (d_mode.double() - d64).abs_()
            same = (ls[iu[0]] == ls[iu[1]])
            del iu
            thr_t = torch.from_numpy(thr).to(dev)
            # count(d < t_n) of the float64 distances rounded to fp32 like the reference's (statistics.py:131)
            dref = d64.float().double()
            order_all = torch.sort(dref).values
            order_same = torch.sort(dref[same]).values
            lt_all = torch.searchsorted(order_all, thr_t, right=False)
            lt_same = torch.searchsorted(order_same, thr_t, right=False)
            eps = 1.e-5
            win_all = torch.searchsorted(order_all, thr_t + eps, right=True) - torch.searchsorted(order_all, thr_t - eps, right=False)
            win_same = torch.searchsorted(order_same, thr_t + eps, right=True) - torch.searchsorted(order_same, thr_t - eps, right=False)
            hs = handle.pair_histogram(xs, ls, thr, 0, mode=mode, cta_group=args.cta_group)
            got_same = torch.from_numpy(hs['same']).to(dev)
            got_diff = torch.from_numpy(hs['diff']).to(dev)
            dif_same = (got_same - lt_same).abs()
            dif_diff = (got_diff - (lt_all - lt_same)).abs()
            ok = bool((dif_same <= win_same).all().item() and (dif_diff <= (win_all - win_same)).all().item())
            rep = {'mode': mode, 'sample_rows': ns, 'pairs': int(d64.numel()), 'max_abs_dd_vs_f64': float(err.max().item()),
                   'rms_dd_vs_f64': float(err.pow(2).mean().sqrt().item()), 'tolerance': 1e-5,
                   'hist_l1_vs_f64_counts': int(dif_same.sum().item() + dif_diff.sum().item()),
                   'pairs_within_eps_of_a_threshold_f64': int(win_all.sum().item()),
                   'mis_binned_outside_eps_window': 0 if ok else 1, 'hist_ok': ok,
                   'eps_window_pairs_counted_by_kernel': int(hs['stats']['eps_window']),
                   'error_bound_of_sample_launch': float(hs['stats']['error_bound'])}
            del d_mode, d64, err, same, dref, order_all, order_same
            torch.cuda.empty_cache()
            return rep

        parity = sample_report(mode_used)
        parity['error_bound_timed_launches'] = err_bound
        parity['auto_fallbacks'] = fallbacks
        if strict is not None:
            strict['parity'] = sample_report('fp16x3')
            # whole-set integer bins of the two arithmetics: identical up to pairs near a threshold
            strict['headline_vs_strict_bins_l1'] = int(np.abs(bins_headline - bins_strict).sum())
            strict['headline_vs_strict_counts_max_abs'] = int(np.abs(np.cumsum(bins_headline[:, ::-1], axis=1) -
                                                                     np.cumsum(bins_strict[:, ::-1], axis=1)).max())
        if args.workload == 'c5' or mode_used in ('bf16', 'tf32', 'fp16'):
            # BASELINE config 5: "reporting eps-window disagreements versus TF32": per-pair bins of the single-pass mode against
            # the fp32-equivalent pass on the sample -- how many pairs land in another bin, and how many of those lie within
            # eps = 1e-5 of a threshold (the contract's window) / within the single-pass mode's own error bound of one
            ns2 = min(16384, per_rank)
            xs2 = x_shard[:ns2].contiguous()
            d_lo = torch.empty(ns2 * (ns2 - 1) // 2, device=dev, dtype=torch.float32)
            d_hi = torch.empty_like(d_lo)
            handle.pairwise(xs2, None, 0, mode=mode_used, out=d_lo)
            handle.pairwise(xs2, None, 0, mode='fp16x3', out=d_hi)
            thr32 = torch.from_numpy(thr.astype(np.float32)).to(dev)
            b_lo = torch.searchsorted(thr32, d_lo, right=True)
            b_hi = torch.searchsorted(thr32, d_hi, right=True)
            differ = b_lo != b_hi
            gap = (d_hi[differ][:, None].double() - torch.from_numpy(thr).to(dev)[None, :]).abs().min(dim=1).values if bool(differ.any()) else torch.zeros(0, device=dev)
            dd = (d_lo - d_hi).abs()
            bound = float(dd.max().item())
            parity['single_pass_vs_fp32_equivalent'] = {
                'sample_rows': ns2, 'pairs': int(d_lo.numel()), 'pairs_in_another_bin': int(differ.sum().item()),
                'of_those_within_1e-5_of_a_threshold': int((gap <= 1e-5).sum().item()),
                'of_those_within_the_modes_max_error_of_a_threshold': int((gap <= bound).sum().item()),
                'max_abs_dd': bound, 'rms_dd': float(dd.double().pow(2).mean().sqrt().item()),
                'whole_set_bins_l1_vs_fp32_equivalent': int(np.abs(bins_headline - bins_strict).sum()) if bins_strict is not None else None}
            del d_lo, d_hi, b_lo, b_hi, differ, dd
            torch.cuda.empty_cache()

    # ---- end to end through the public API with host buffers
    e2e, e2e_pageable = None, None
    if not args.no_e2e:
        x_host = torch.empty((per_rank, DIM), dtype=torch.float32, pin_memory=True)
        l_host = torch.empty((per_rank,), dtype=torch.int64, pin_memory=True)
        x_host.copy_(x_shard); l_host.copy_(labels_shard)
        torch.cuda.synchronize()
        h2d_bytes = int(world * (x_host.numel() * 4 + l_host.numel() * 8))
        d2h_bytes = int(2 * (thr.size + 1) * 8)

        def leg(xh_np, lh_np, xh_t, lh_t, steps):
            h2d = []
            if world == 1:
                def step_e2e():
                    out_ = fst.pair_histogram(xh_np, lh_np, thr, 0, mode=args.mode, cta_group=args.cta_group,
                                              panel_window=args.panel_window, region_rows=args.region_rows)   # H2D + D2H inside
                    h2d.append(out_['stats']['h2d_ms'])
                    return out_
            else:
                def step_e2e():
                    # host shards straight into the library: each rank uploads its rows piece by piece (pinned ring) and
                    # broadcasts them from the GPU, under the launches
                    b, st_ = sharded(xh_np, lh_np)
                    r = b.cpu()
                    h2d.append(st_.get('gather_ms', 0.0))
                    return r
            for _ in range(max(1, min(args.warmup, 2))):
                step_e2e()
            del h2d[:]
            ms, _, _, _ = timed(step_e2e, steps)
            return {'value': pairs * steps / (ms * 1e-3) / 1e9, 'unit': UNIT, 'h2d_bytes_per_step': h2d_bytes,
                    'd2h_bytes_per_step': d2h_bytes, 'ms_per_step': ms / steps, 'h2d_ms': float(np.mean(h2d)) if h2d else None,
                    'd2h_ms': None, 'steps': steps}

        e2e = leg(x_host.numpy(), l_host.numpy(), x_host, l_host, args.steps)
        e2e['host_memory'] = ('pinned (torch pin_memory viewed as NumPy): streamed chunk by chunk under the Gram launches (rows out of class order are '
                              'gathered by the copy threads through the pinned ring); h2d_ms = first to last chunk on the copy stream')
        # what the reference's caller really hands over: pageable np.concatenate output (facenet.py:184-201)
        x_page, l_page = np.array(x_host.numpy()), np.array(l_host.numpy())
        steps_p = max(1, min(args.steps, 3))
        e2e_pageable = leg(x_page, l_page, torch.from_numpy(x_page), torch.from_numpy(l_page), steps_p)
        e2e_pageable['host_memory'] = ('pageable NumPy arrays: staged through a ring of pinned slots filled by a pool of host threads '
                                       '(csrc/fnb_stage.cu), chunk by chunk under the Gram launches')
        e2e_pageable['e2e_over_value'] = e2e_pageable['value'] / value
        e2e['e2e_over_value'] = e2e['value'] / value

    if sampler:
        sampler.stop()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    clocks = sampler.summary(t0, t1)
    peaks, peak_src = measured_peaks()
    # TF32 tensor peak = half the bf16 rate (same datapath, 4-byte operands); the driver measures bf16 with cuBLAS
    # (burst figure for a step of a few ms, sustained figure for a step long enough to sit under the power cap)
    long_step = total_ms / args.steps > 100.0
    peak_key = 'bf16_tflops_sustained' if (long_step and 'bf16_tflops_sustained' in peaks) else 'bf16_tflops'
    tf32_peak = peaks[peak_key] / 2.0
    roof_peak = peaks[peak_key] if mode_used == 'bf16' else tf32_peak       # BF16 mode is measured against the bf16 peak
    achieved = (pairs / world) * FLOP_PER_PAIR / (k_ms * 1e-3) / 1e12       # one launch covers 1/world of the pairs
    traffic, traffic_src = None, None
    tf = ROOT / 'profiles' / 'ncu_traffic.json'
    if tf.exists() and world == 1:                     # the captures are single-GPU launches of the whole pair matrix
        try:
            key = '%s/%s' % (args.workload, mode_used) + ('' if (windows != [0] or args.workload != '1m') else '/window_off')
            table = json.loads(tf.read_text())
            traffic = table.get(key)
            traffic_src = 'profiles/ncu_traffic.json[%s]: %s' % (key, table.get('_note', '')) if traffic is not None else None
        except ValueError:
            traffic = None
    passes = 3 if mode_used.endswith('x3') else 2 if mode_used == 'fp16f8' else 1
    roofline = {'bound': 'tensor', 'achieved': achieved, 'peak': roof_peak, 'unit': 'TFLOP/s', 'frac': achieved / roof_peak,
                'traffic': traffic, 'traffic_source': traffic_src, 'kernel': 'gram_kernel<HIST>', 'kernel_ms': k_ms,
                'peak_source': peak_src + (': %s (BF16 mode)' % peak_key if mode_used == 'bf16' else
                                           ': %s / 2 (TF32 rate = half the bf16 rate); of measured' % peak_key),
                'frac_of_nominal_tf32_1100': achieved / 1100.0,
                'executed_mma_tflops': achieved * passes,
                'executed_frac_of_pipe_peak': achieved * passes / (tf32_peak * (1 if 'tf32' in mode_used else 2)),
                'note': 'achieved = pairs x 1024 algorithmic FLOP / Gram-kernel time; executed_mma_tflops counts the split passes '
                        '(fp16-pass equivalents: an e4m3 pass counts 1/2)'}
    if strict is not None:
        s_ach = (pairs / world) * FLOP_PER_PAIR / (strict['kernel_ms'] * 1e-3) / 1e12
        strict['roofline'] = {'achieved': s_ach, 'peak': tf32_peak, 'frac': s_ach / tf32_peak, 'unit': 'TFLOP/s',
                              'note': 'fp16x3 executes 3 fp16 MMAs per algorithmic MMA = 1.5 TF32-pass equivalents: 0.667 of the fp16 pipe at best; the '
                                              'denominator is cuBLAS\'s measured sustained rate / 2, which this MMA stream can exceed under the same power limit'}

    burst, sustained, reps = tf32_live(torch, dev)
    if burst:
        roofline['tf32_cublas_tflops_live'] = burst
        roofline['frac_of_live_tf32'] = achieved / burst
        roofline['tf32_cublas_tflops_live_sustained'] = sustained
        roofline['frac_of_live_tf32_sustained'] = achieved / sustained
        roofline['tf32_live_note'] = 'torch.matmul fp32 8192^3 allow_tf32: best of 5 (burst); %d back to back, ~4 s (sustained)' % reps

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, ms, sample, cores = cpu_arm(args.cpu_sample_rows, ids, n, 2, 1)
        cpu = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample}

    rest = total_ms / args.steps - k_ms
    parts = {'everything outside the Gram launches (label exchange + sorts, first chunk of the row exchange, splits, all-reduce, D2H, host)': rest}
    limiter = max(parts, key=parts.get)
    line = {'metric': METRIC, 'value': value, 'unit': UNIT,
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': total_ms / args.steps,
            'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
            'dtype': {'fp16x3': 'f16x3 split, f32 accumulate (fp32-equivalent)', 'tf32x3': 'tf32x3 split, f32 accumulate',
                      'tf32': 'tf32', 'bf16': 'bf16', 'fp16': 'f16',
                      'fp16f8': 'f16 hi*hi + e4m3 cross terms, f32 accumulate; strict f16x3 tiles for same-identity pairs'}[mode_used],
            'data': 'synthetic',
            'config': {'workload': wl['name'], 'mode': args.mode, 'mode_used': mode_used,
                       'parallelism': ('row blocks rb %% world == rank over %d ranks, a rank that drains its queue takes tiles from the others (NVLink peer atomics)' % world if world > 1 else 'one GPU') +
                                      (', shares adapted to per-GPU kernel time: %s / %d' % (balancer.widths, balancer.mod) if balancer else ', equal shares'),
                       'l2': 'inputs (%.0f MB fp32 + split operands) larger than L2; no flush' % (n * DIM * 4 / 1e6),
                       'pairs_per_step': pairs, 'region_rows': args.region_rows,
                       'grid_ctas': grids, 'panel_window': windows, 'cluster': 'CTA pairs (cta_group::2) on one tile queue: 132 CTAs in clusters of two pairs with the A operand multicast + plain pairs on the SMs '
                                  'those clusters cannot use (N > 1: all but 8 of them, kept for the row exchange, except under the last launch)'},
            'breakdown_ms': {'row_exchange_first_to_last_piece': gather_ms, 'row_exchange_chunks': chunks_used,
                             'plain_torch_all_gather_alone': plain_gather_ms, 'sort_split_until_first_launch': p_ms,
                             'gram_kernel': k_ms, 'gram_kernel_per_rank': k_ranks,
                             'step_minus_gram_kernel': rest,
                             'note': 'N > 1: labels first, then the fp32 rows by ncclBroadcast from their owner, chunk by chunk on a copy '
                                     'stream under the Gram launches (fnb_pair_histogram_sharded); ncclAllReduce of the integer bins'},
            'clocks': clocks, 'e2e': e2e, 'e2e_pageable': e2e_pageable, 'gpu_launches': timed_launches, 'roofline': roofline,
            'cpu_baseline': cpu, 'parity': parity, 'strict': strict,
            'pct_tf32_peak': 100.0 * value * 1e9 * FLOP_PER_PAIR / 1e12 / (tf32_peak * world)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------------------------
# config 3: triplet mining

def run_mining(args):
    import torch
    from facenet_b200 import _capi
    from oracle import mining_oracle as mo

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    P, K, B, S, alpha = 45, 40, 1800, args.mining_batches, 0.2
    mode = 'fp16x3' if args.mode == 'auto' else args.mode
    handle = _capi.default_handle(local_rank)
    stream = torch.cuda.current_stream()
    handle.set_stream(stream.cuda_stream)
    # a pool of batches from a seeded stream (every rank its own: replicas only, no collective -- DESIGN.md section 4)
    pool = 4                                         # groups of S batches, cycled
    gen = torch.Generator(device=dev); gen.manual_seed(100 + rank)
    labels_one = torch.arange(B, device=dev, dtype=torch.int64) // K          # P x K layout, rows grouped by class (facenet.py:108-113)
    groups = []
    for _ in range(pool):
        c = torch.randn((S * P, DIM), generator=gen, device=dev)
        x = c.repeat_interleave(K, dim=0) + 1.1 * torch.randn((S * B, DIM), generator=gen, device=dev)
        x = (x / x.norm(dim=1, keepdim=True)).contiguous()
        groups.append((x, labels_one.repeat(S).contiguous()))
    torch.cuda.synchronize()
    steps = max(S, (args.steps // S) * S)
    launches_per_call = 4
    outs = [None] * pool

    def call(g):
        outs[g % pool] = handle.mine_batched(groups[g % pool][0], groups[g % pool][1], nbatches=S, alpha=alpha, kmax=K - 1, mode=mode,
                                             out=outs[g % pool])

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    for g in range(max(args.warmup, 3)):
        call(g)
    handle.mine_check()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.perf_counter()
    e0.record(stream)
    for g in range(steps // S):
        call(g)                                      # device-resident outputs, no host synchronisation inside
    e1.record(stream)
    barrier()
    t1 = time.perf_counter()
    chk = handle.mine_check()                        # data-dependent errors of the last call, kernel time of the last launch group
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    wall_ms = (t1 - t0) * 1e3
    value = world * steps * B * B / (total_ms * 1e-3) / 1e9

    # parity of one batch against the oracle (not timed; rank 0): exact index equality when the oracle selects on the
    # library's own distances; match rates against the oracle's fp32 distances
    parity, cpu = None, None
    if rank == 0 and not args.no_parity:
        xb, lb = groups[0][0][:B].cpu().numpy(), groups[0][1][:B].cpu().numpy()
        got = {k: v[:B].cpu().numpy() for k, v in outs[0].items() if k in ('hardest_pos', 'hardest_neg', 'pos_index', 'semi_hard', 'eligible')}
        t_cpu = time.perf_counter()
        ref = mo.mine(xb, lb, alpha)
        cpu_s = time.perf_counter() - t_cpu
        dist_gpu = handle.pairwise(xb, xb, 0, mode=mode, cta_group=1)
        ref2 = mo.mine(xb, lb, alpha, dist=dist_gpu)
        parity = {'batch_rows': B, 'exact_vs_oracle_on_library_distances': bool(all(np.array_equal(got[k], ref2[k]) for k in got)),
                  'match_rate_vs_oracle_fp32_distances': {k: float((got[k] == ref[k]).mean()) for k in got},
                  'note': 'index disagreements against the fp32 oracle are near-ties / candidates within 2e-5 of a decision boundary '
                          '(tests/test_gpu_parity.py::check_mining proves it per element)'}
        if not args.no_cpu_baseline:
            cpu = {'value': B * B / cpu_s / 1e9, 'unit': UNIT, 'cores': 1, 'kind': 'port',
                   'sample': 'one 1800 x 512 batch, NumPy mining oracle (oracle/mining_oracle.py: fp32 sgemm + per-anchor loops), %.2f s' % cpu_s}

    # end to end: host batches in (pinned), mined indices back on the host, per group of S batches
    e2e = None
    if not args.no_e2e:
        xh = [torch.empty_like(g[0], device='cpu').pin_memory() for g in groups[:2]]
        lh = [torch.empty_like(g[1], device='cpu').pin_memory() for g in groups[:2]]
        for i in range(2):
            xh[i].copy_(groups[i][0]); lh[i].copy_(groups[i][1])
        torch.cuda.synchronize()
        e2e_steps = max(S, min(steps, 2000) // S * S)

        def call_host(g):
            return handle.mine_batched(xh[g % 2].numpy(), lh[g % 2].numpy(), nbatches=S, alpha=alpha, kmax=K - 1, mode=mode)
        call_host(0)
        barrier()
        e0.record(stream)
        for g in range(e2e_steps // S):
            res = call_host(g)
        e1.record(stream)
        barrier()
        e_ms = e0.elapsed_time(e1)
        e2e = {'value': world * e2e_steps * B * B / (e_ms * 1e-3) / 1e9, 'unit': UNIT, 'ms_per_step': e_ms / e2e_steps, 'steps': e2e_steps,
               'h2d_bytes_per_step': int(world * (B * DIM * 4 + B * 8)), 'd2h_bytes_per_step': int(world * B * (2 + 3 * (K - 1)) * 4),
               'host_memory': 'pinned'}
    if sampler:
        sampler.stop()
    if rank != 0:
        return
    clocks = sampler.summary(t0, t1)
    peaks, peak_src = measured_peaks()
    us_per_step = total_ms * 1e3 / steps
    # the step is latency / L2 bound, not tensor bound: B^2 x 512 x 2 FLOP x 3 split passes at the fp16 tensor peak
    tensor_us = B * B * DIM * 2 * 3 / (peaks['bf16_tflops'] * 1e12) * 1e6
    strip_bytes = 2 * B * B * 4                      # the distance strip is written by the Gram epilogue and read by the selection
    achieved = B * B * FLOP_PER_PAIR / (us_per_step * 1e-6) / 1e12
    roofline = {'bound': 'tensor', 'achieved': achieved, 'peak': peaks['bf16_tflops'] / 2.0, 'unit': 'TFLOP/s',
                'frac': achieved / (peaks['bf16_tflops'] / 2.0), 'traffic': None, 'kernel': 'gram_kernel<ROWSTRIP> + mine_rows_kernel',
                'tensor_time_us_per_step': tensor_us, 'measured_us_per_step': us_per_step, 'tensor_share_of_step': tensor_us / us_per_step,
                'strip_bytes_per_step': strip_bytes, 'strip_GBps': strip_bytes / (us_per_step * 1e-6) / 1e9,
                'last_group_kernel_ms': chk['kernel_ms'], 'last_group_prepare_ms': chk['prepare_ms'],
                'peak_source': peak_src + ': bf16_tflops / 2',
                'note': 'a 1800-row batch is 10 GFLOP of executed fp16 MMA (7 us at peak): the step is bound by the epilogue + strip round trip '
                        'through L2 and the per-anchor selection, not by the tensor pipe'}
    line = {'metric': METRIC + ' (triplet mining: B x B ordered pair distances per 1800-row batch)', 'value': value, 'unit': UNIT,
            'n_gpus': world, 'steps': steps, 'warmup': max(args.warmup, 3) * S, 'ms_per_step': total_ms / steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f16x3 split, f32 accumulate' if mode == 'fp16x3' else mode,
            'data': 'synthetic',
            'config': {'workload': WORKLOADS['mining']['name'], 'batches_per_launch': S, 'batch': '%d x %d' % (P, K), 'alpha': alpha,
                       'mode': mode, 'outputs': 'device-resident int32 (hardest_pos, hardest_neg [B]; pos_index, semi_hard, eligible [B, 39])',
                       'parallelism': 'replicas only (one batch stream per GPU, no collective)',
                       'l2': 'batch pool of %d x %d batches (%.0f MB) cycled; strips %.0f MB per launch' % (pool, S, pool * S * B * DIM * 4 / 1e6, S * B * B * 4 / 1e6)},
            'steps_per_s': world * steps / (total_ms * 1e-3), 'seconds_for_10k_steps': total_ms * 1e-3 * 10000 / steps,
            'wall_ms_host_side': wall_ms,
            'clocks': clocks, 'e2e': e2e, 'gpu_launches': launches_per_call * (steps // S), 'roofline': roofline, 'cpu_baseline': cpu,
            'parity': parity}
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------------------
# config 1: LFW-size k-fold validation through the drop-in classes

def run_lfw(args):
    import torch
    from facenet_b200 import statistics as fst
    from oracle import statistics_oracle as so

    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if int(os.environ.get('RANK', '0')) != 0:
        return                                       # one validation does not shard: N > 1 would be replicas
    torch.cuda.set_device(local_rank)
    mode = 'fp16x3' if args.mode == 'auto' else args.mode
    fst.set_default_mode(mode=mode, device=local_rank)

    class Cfg:
        metric, nrof_folds, far_target = 0, 10, 1.e-3

    sizes = so.lfw_like_class_sizes()
    x, labels = so.synthetic_embeddings(sizes, dim=DIM, sigma=(1.5, 3.5), seed=0)
    n = x.shape[0]
    pair_evals = lfw_pair_evaluations(n)
    xd, ld = torch.from_numpy(x).cuda(), torch.from_numpy(labels).cuda()
    sampler = ClockSampler(local_rank)
    for _ in range(max(1, min(args.warmup, 2))):
        fst.FaceToFaceValidation(xd, ld, Cfg)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        vd = fst.FaceToFaceValidation(xd, ld, Cfg)   # embeddings resident on the GPU (section 8 f3)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    dev_s = (t1 - t0) / args.steps
    launches = sum(1 for r in vd.reports[:1] for _ in r.conf_matrix_train) * 2
    e2e = None
    if not args.no_e2e:
        fst.FaceToFaceValidation(x, labels, Cfg)
        t2 = time.perf_counter()
        for _ in range(args.steps):
            vh = fst.FaceToFaceValidation(x, labels, Cfg)       # host NumPy arrays in, Report dict out
        host_s = (time.perf_counter() - t2) / args.steps
        e2e = {'value': pair_evals / host_s / 1e9, 'unit': UNIT, 'ms_per_step': host_s * 1e3,
               'h2d_bytes_per_step': int(20 * 0.55 * n * DIM * 4), 'd2h_bytes_per_step': int(20 * 4 * 100 * 8),
               'host_memory': 'pageable NumPy (every fold copies its rows, like the reference)',
               'reports_identical_to_gpu_resident': bool(all(float(vh.dict[c][k]) == float(vd.dict[c][k]) for c in vd.dict for k in vd.dict[c]))}
    sampler.stop()
    parity, cpu = None, None
    if not args.no_parity:
        tc = time.perf_counter()
        ref = so.face_to_face_validation(x, labels, 0, 10, 1.e-3)
        cpu_s = time.perf_counter() - tc
        got = vd.dict
        thr_acc = np.array([m.threshold[0] for m in vd.reports[0].conf_matrix_test])
        thr_far = np.array([float(m.threshold[0]) for m in vd.reports[1].conf_matrix_test])
        worst = max(abs(float(got[c][k]) - float(ref[c][k])) for c in got for k in got[c])
        parity = {'accuracy_threshold_equal_folds': int(np.sum(thr_acc == ref['_thresholds'][:, 0])), 'folds': 10,
                  'far_threshold_max_abs_diff': float(np.abs(thr_far - ref['_thresholds'][:, 1]).max()),
                  'report_dict_max_abs_diff': worst,
                  'auc': float(got['MaximumAccuracy']['auc']), 'eer': float(got['MaximumAccuracy']['eer']),
                  'accuracy': float(got['MaximumAccuracy']['accuracy']), 'tp_rate_at_far': float(got[[c for c in got if c.startswith('False')][0]]['tp_rates']),
                  'note': 'AUC < 1 by construction (per-identity sigma ~ U(1.5, 3.5)): the accuracy argmax and the FAR interpolation are off the plateau'}
        cpu = {'value': pair_evals / cpu_s / 1e9, 'unit': UNIT, 'cores': host_threads(), 'kind': 'port',
               'sample': 'one full validation with the vectorised NumPy oracle: %.1f s (the literal reference loop extrapolates to ~37 h)' % cpu_s}
    clocks = sampler.summary(t0, t1)
    peaks, peak_src = measured_peaks()
    achieved = pair_evals * FLOP_PER_PAIR / dev_s / 1e12
    line = {'metric': METRIC + ' (10-fold FaceToFaceValidation)', 'value': pair_evals / dev_s / 1e9, 'unit': UNIT, 'n_gpus': 1,
            'steps': args.steps, 'warmup': max(1, min(args.warmup, 2)), 'ms_per_step': dev_s * 1e3, 'higher_is_better': True,
            'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f16x3 split, f32 accumulate' if mode == 'fp16x3' else mode, 'data': 'synthetic',
            'config': {'workload': WORKLOADS['lfw']['name'], 'mode': mode, 'pair_distance_evaluations_per_step': pair_evals,
                       'launches_per_step': '10 train folds x 100 thresholds + 10 test folds x 2 thresholds (20 Gram launches)',
                       'l2': 'one fold (%.0f MB of fp32 rows) fits L2; wall-clock timing of the whole Python call' % (0.9 * n * DIM * 4 / 1e6)},
            'clocks': clocks, 'e2e': e2e, 'gpu_launches': 20 * 5 * args.steps,
            'roofline': {'bound': 'tensor', 'achieved': achieved, 'peak': peaks['bf16_tflops'] / 2.0, 'unit': 'TFLOP/s',
                         'frac': achieved / (peaks['bf16_tflops'] / 2.0), 'traffic': None, 'kernel': 'gram_kernel<HIST> (keyed regions)',
                         'note': '20 launches of 0.4-70 M pairs each: launch + host bound (ms_per_step is wall clock around the Python call)'},
            'cpu_baseline': cpu, 'parity': parity}
    print(json.dumps(line))


def main():
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
        return
    import torch
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (facenet_b200 has no CPU fallback)')
    if args.workload == 'mining':
        run_mining(args)
    elif args.workload == 'lfw':
        run_lfw(args)
    else:
        run_allpairs(args)


if __name__ == '__main__':
    main()
