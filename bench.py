"""Headline benchmark: all-pairs 512-d verification histogram, G pair-distances/s (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One step = one pass of the hot path over one batch of synthetic input: every unordered pair of the
N x 512 embedding set is evaluated once (Gram contraction on tcgen05 tensor cores) and binned against
the reference's 100 thresholds, split by same / different identity.  Default workload: BASELINE
config "synthetic 1M x 512 embeddings, all-pairs verification sharded over 1/2/4/8 B200" (it fits one
GPU; 4.999995e11 pairs per step); `--workload 100k` selects the 100k x 512 config.

  value  device-resident inputs (each rank holds its row shard in HBM), CUDA events around K steps,
         max over ranks; per step: [all-gather of shards] -> sort/split -> Gram+histogram -> [all-reduce]
  e2e    same metric through the public API with HOST (pinned) buffers: H2D of the step's embeddings
         and labels and D2H of the histogram inside the timed region
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
}
DIM = 512
FLOP_PER_PAIR = 2 * DIM          # one pair distance = 1024 algorithmic FLOP


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='1m', choices=sorted(WORKLOADS))
    ap.add_argument('--mode', default='auto', choices=['auto', 'fp16x3', 'fp16f8', 'tf32x3', 'tf32', 'bf16', 'fp16'])
    ap.add_argument('--cta-group', type=int, default=0)
    ap.add_argument('--cpu-sample-rows', type=int, default=16384)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-parity', action='store_true')
    ap.add_argument('--adaptive-shares', action='store_true', help='N > 1: speed-adaptive shares instead of equal ones')
    ap.add_argument('--panel-window', type=int, default=0,
                    help='fnb_options.panel_window: 0 auto (on for launches of >= 5e10 pairs per rank), -1 off, 1..7 window')
    ap.add_argument('--parity-rows', type=int, default=8192)
    return ap.parse_args()


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


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    val, ms, sample, cores = cpu_arm(args.cpu_sample_rows, wl['ids'], wl['n'], max(1, args.steps), min(args.warmup, 1))
    line = {'impl': 'reference', 'metric': 'G pair-distances/sec all-pairs 512-d verification', 'value': val,
            'unit': 'G pair-distances/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic', 'config': {'workload': wl['name'], 'sample': sample},
            'cpu_baseline': {'value': val, 'unit': 'G pair-distances/s', 'cores': cores, 'kind': 'port', 'sample': sample},
            'e2e': {'value': val, 'unit': 'G pair-distances/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line))


def main():
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from facenet_b200 import _capi, distributed as fd, statistics as fst

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (facenet_b200 has no CPU fallback)')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)
    wl = WORKLOADS[args.workload]
    n, ids = wl['n'], wl['ids']
    assert n % world == 0
    per_rank = n // world
    thr = thresholds()
    pairs = n * (n - 1) // 2

    # ---- synthetic inputs, generated on the device with a fixed seed (identical on every rank), then sharded
    gen = torch.Generator(device=dev)
    gen.manual_seed(0)
    centres = torch.randn((ids, DIM), generator=gen, device=dev, dtype=torch.float32)
    labels_full = torch.arange(n, device=dev, dtype=torch.int64) % ids
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
    kernel_ms, prepare_ms, launches, modes_used, grids, windows = [], [], [0], set(), set(), set()

    def hist_fn(emb, labels, thresholds_, metric, rank_, world_, bins_out, **kw):
        _, st = handle.pair_histogram_bins(emb, labels, thresholds_, metric, rank=rank_, world=world_, bins_out=bins_out,
                                           mode=args.mode, cta_group=args.cta_group, shard=kw.get('shard'),
                                           panel_window=args.panel_window)
        kernel_ms.append(st['kernel_ms'])
        prepare_ms.append(st['prepare_ms'])
        launches[0] += st['kernel_launches']
        modes_used.add(_capi.MODE_NAMES[st['mode_used']])
        grids.add(st['grid_ctas'])
        windows.add(st['panel_window'])
        return st

    # equal shares by default; --adaptive-shares lets the shares follow the measured speed of each GPU (measured on two
    # 8-GPU boxes: the per-rank kernel times level out, but the step is no shorter -- profiles/r01c_multi_gpu.md)
    balancer = fd.default_balancer(world) if (world > 1 and args.adaptive_shares) else None

    def step_device():
        bins, st = fd.pair_histogram_sharded(x_shard, labels_shard, thr, 0, hist_fn=hist_fn, balancer=balancer)
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
    kernel_ms.clear(); prepare_ms.clear(); launches[0] = 0
    total_ms, bins, t0, t1 = timed(step_device, args.steps)
    k_ms = float(np.mean(kernel_ms)) if kernel_ms else float('nan')
    p_ms = float(np.mean(prepare_ms)) if prepare_ms else float('nan')
    # per-rank Gram-kernel time: the all-reduce makes every step wait for the slowest rank
    k_ranks = [k_ms]
    if world > 1:
        kt = torch.tensor([k_ms], device=dev, dtype=torch.float64)
        allk = [torch.zeros_like(kt) for _ in range(world)]
        dist.all_gather(allk, kt)
        k_ranks = [float(t.item()) for t in allk]
    # where the rest of the step goes (not part of `value`): the all-gather alone, timed the same way
    gather_ms = 0.0
    if world > 1:
        g_total, _, _, _ = timed(lambda: fd.gather_shards(x_shard, labels_shard), args.steps)
        gather_ms = g_total / args.steps
    timed_launches = launches[0]
    value = pairs * args.steps / (total_ms * 1e-3) / 1e9
    mode_used = sorted(modes_used)[0] if len(modes_used) == 1 else args.mode      # what 'auto' resolved to

    # sanity of the result (not timed): every pair counted once, same-identity total known analytically
    if rank == 0:
        out = fd.counts_from_bins(bins, thr, 0)
        assert out['n_same'] + out['n_diff'] == pairs, (out['n_same'], out['n_diff'], pairs)
        assert out['n_same'] == ids * (n // ids) * (n // ids - 1) // 2

    # ---- parity of the timed arithmetic mode on a row sample of THIS workload (not timed; rank 0).  The checker is
    #      plain torch float64 on the GPU (the CPU oracle is the checker in tests/ and smoke()):
    #        max |d_mode - d_f64| over all pairs of the sample           (tolerance of BASELINE.json: 1e-5)
    #        histogram of the sample in this mode vs bins of the float64 distances: L1 difference, which must be
    #        covered by the pairs the kernel itself counted inside the eps window
    parity = None
    if rank == 0 and not args.no_parity:
        ns = min(args.parity_rows, per_rank)
        xs, ls = x_shard[:ns].contiguous(), labels_shard[:ns].contiguous()
        d_mode = torch.empty(ns * (ns - 1) // 2, device=dev, dtype=torch.float32)
        handle.pairwise(xs, None, 0, mode=mode_used, out=d_mode)
        torch.cuda.synchronize()
        x64 = xs.double()
        iu = torch.triu_indices(ns, ns, 1, device=dev)
        d64 = (2.0 * (1.0 - (x64 @ x64.T).clamp_(-1.0, 1.0)))[iu[0], iu[1]]
        err = (d_mode.double() - d64).abs_()
        same = (ls[iu[0]] == ls[iu[1]])
        thr_t = torch.from_numpy(thr).to(dev)
        # bin k of a distance = number of thresholds <= d  (d < t_n  <=>  k <= n), float64 compare like statistics.py:131
        b64 = torch.searchsorted(thr_t, d64.float().double(), right=True)
        ref_lt_all = torch.bincount(b64, minlength=thr.size + 1).cumsum(0)[:thr.size]
        ref_lt_same = torch.bincount(b64[same], minlength=thr.size + 1).cumsum(0)[:thr.size]
        hs = handle.pair_histogram(xs, ls, thr, 0, mode=mode_used, cta_group=args.cta_group)
        got_same = torch.from_numpy(hs['same']).to(dev)
        got_all = torch.from_numpy(hs['same'] + hs['diff']).to(dev)
        l1 = int((got_same - ref_lt_same).abs().sum().item() + ((got_all - got_same) - (ref_lt_all - ref_lt_same)).abs().sum().item())
        parity = {'sample_rows': ns, 'pairs': int(d64.numel()), 'max_abs_dd_vs_f64': float(err.max().item()),
                  'rms_dd_vs_f64': float(err.pow(2).mean().sqrt().item()), 'tolerance': 1e-5,
                  'hist_l1_vs_f64_bins': l1, 'eps_window_pairs_counted': int(hs['stats']['eps_window']),
                  'hist_ok': bool(l1 <= 2 * hs['stats']['eps_window'])}
        del d_mode, x64, d64, err, iu, same, b64
        torch.cuda.empty_cache()

    # ---- end to end through the public API with pinned host buffers
    e2e = None
    if not args.no_e2e:
        x_host = torch.empty((per_rank, DIM), dtype=torch.float32, pin_memory=True)
        l_host = torch.empty((per_rank,), dtype=torch.int64, pin_memory=True)
        x_host.copy_(x_shard); l_host.copy_(labels_shard)
        torch.cuda.synchronize()
        xh_np, lh_np = x_host.numpy(), l_host.numpy()

        if world == 1:
            def step_e2e():
                return fst.pair_histogram(xh_np, lh_np, thr, 0, mode=args.mode, cta_group=args.cta_group)   # H2D + D2H inside
        else:
            xd = torch.empty_like(x_shard); ld = torch.empty_like(labels_shard)

            def step_e2e():
                xd.copy_(x_host, non_blocking=True); ld.copy_(l_host, non_blocking=True)
                b, _ = fd.pair_histogram_sharded(xd, ld, thr, 0, hist_fn=hist_fn, balancer=balancer)
                return b.cpu()
        for _ in range(max(1, min(args.warmup, 2))):
            step_e2e()
        e2e_ms, _, _, _ = timed(step_e2e, args.steps)
        e2e = {'value': pairs * args.steps / (e2e_ms * 1e-3) / 1e9, 'unit': 'G pair-distances/s',
               'h2d_bytes_per_step': int(world * (x_host.numel() * 4 + l_host.numel() * 8)),
               'd2h_bytes_per_step': int(2 * (thr.size + 1) * 8), 'ms_per_step': e2e_ms / args.steps}

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
    achieved = (pairs / world) * FLOP_PER_PAIR / (k_ms * 1e-3) / 1e12       # one launch covers 1/world of the pairs
    traffic, traffic_src = None, None
    tf = ROOT / 'profiles' / 'ncu_traffic.json'
    if tf.exists() and world == 1:                     # the captures are single-GPU launches of the whole pair matrix
        try:
            # the DRAM bytes of the long launches depend on the cluster-progress window (profiles/r01d_panel_window.md)
            key = '%s/%s' % (args.workload, mode_used) + ('' if (windows != {0} or args.workload != '1m') else '/window_off')
            table = json.loads(tf.read_text())
            traffic = table.get(key)
            traffic_src = 'profiles/ncu_traffic.json[%s]: %s' % (key, table.get('_note', '')) if traffic is not None else None
        except ValueError:
            traffic = None
    passes = 3 if mode_used.endswith('x3') else 2 if mode_used == 'fp16f8' else 1
    roofline = {'bound': 'tensor', 'achieved': achieved, 'peak': tf32_peak, 'unit': 'TFLOP/s', 'frac': achieved / tf32_peak,
                'traffic': traffic, 'traffic_source': traffic_src, 'kernel': 'gram_kernel<HIST>', 'kernel_ms': k_ms,
                'peak_source': peak_src + ': %s / 2 (TF32 rate = half the bf16 rate); of measured' % peak_key,
                'frac_of_nominal_tf32_1100': achieved / 1100.0,
                'executed_mma_tflops': achieved * passes,
                'executed_frac_of_pipe_peak': achieved * passes / (tf32_peak * (1 if 'tf32' in mode_used else 2)),
                'note': 'achieved = pairs x 1024 algorithmic FLOP / Gram-kernel time; executed_mma_tflops counts the split passes '
                        '(fp16-pass equivalents: an e4m3 pass counts 1/2)'}

    # TF32 tensor peak measured live, the way MEASURED_PEAKS.json measures bf16 (SURVEY.md section 8 d): cuBLAS through
    # torch.matmul, fp32 8192^3 with allow_tf32, best of 5 (burst) -- reported beside the bf16/2 figure, not on the product path
    tf32_live = None
    try:
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        ma = torch.randn((8192, 8192), device=dev)
        mb = torch.randn((8192, 8192), device=dev)
        torch.matmul(ma, mb)
        best = None
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(ma, mb); e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        tf32_live = 2.0 * 8192 ** 3 / (best * 1e-3) / 1e12
        torch.backends.cuda.matmul.allow_tf32 = prev
        del ma, mb
    except Exception:
        tf32_live = None
    if tf32_live:
        roofline['tf32_cublas_tflops_live'] = tf32_live
        roofline['frac_of_live_tf32'] = achieved / tf32_live

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, ms, sample, cores = cpu_arm(args.cpu_sample_rows, ids, n, 2, 1)
        cpu = {'value': v, 'unit': 'G pair-distances/s', 'cores': cores, 'kind': 'port', 'sample': sample}

    line = {'metric': 'G pair-distances/sec all-pairs 512-d verification', 'value': value, 'unit': 'G pair-distances/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': total_ms / args.steps,
            'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
            'dtype': {'fp16x3': 'f16x3 split, f32 accumulate (fp32-equivalent)', 'tf32x3': 'tf32x3 split, f32 accumulate',
                      'tf32': 'tf32', 'bf16': 'bf16', 'fp16': 'f16',
                      'fp16f8': 'f16 hi*hi + e4m3 cross terms, f32 accumulate'}[mode_used],
            'data': 'synthetic',
            'config': {'workload': wl['name'], 'mode': args.mode, 'mode_used': mode_used,
                       'parallelism': ('row blocks split over %d ranks' % world) +
                                      (', shares adapted to per-GPU kernel time: %s / %d' % (balancer.widths, balancer.mod) if balancer else ', equal shares'),
                       'l2': 'inputs (%.0f MB fp32 + split operands) larger than L2; no flush' % (n * DIM * 4 / 1e6),
                       'pairs_per_step': pairs,
                       'grid_ctas': sorted(grids), 'panel_window': sorted(windows), 'cluster': 'CTA pairs (cta_group::2); 132-CTA grids are clusters of two pairs with the A operand multicast'},
            'breakdown_ms': {'all_gather': gather_ms, 'sort_split': p_ms, 'gram_kernel': k_ms, 'gram_kernel_per_rank': k_ranks,
                             'rest (all-reduce, D2H of the bins, host)': total_ms / args.steps - gather_ms - p_ms - k_ms},
            'clocks': clocks, 'e2e': e2e, 'gpu_launches': timed_launches, 'roofline': roofline, 'cpu_baseline': cpu,
            'parity': parity,
            'pct_tf32_peak': 100.0 * value * 1e9 * FLOP_PER_PAIR / 1e12 / (tf32_peak * world)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
