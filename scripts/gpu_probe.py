"""Bring-up probe (not a test): one case per process so that a device trap in one
configuration does not poison the others.  Usage: python scripts/gpu_probe.py <case> [args]"""
import sys
import time

import numpy as np

sys.path.insert(0, '.')
from facenet_b200 import _capi
from oracle import statistics_oracle as so


def unit(n, d, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d), dtype=np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x


def case_pairwise(mode, cg, na=200, nb=150, d=128):
    h = _capi.Handle(0)
    xa, xb = unit(na, d, 1), unit(nb, d, 2)
    ref = so.pairwise_similarities(xa.copy(), xb.copy(), 0)
    got = h.pairwise(xa, xb, 0, mode=mode, cta_group=cg)
    err = np.abs(got - ref)
    print('pairwise cross mode=%s cg=%d na=%d nb=%d d=%d: max|dd|=%.3e mean=%.3e range=%s' %
          (mode, cg, na, nb, d, err.max(), err.mean(), h.last_range))
    if err.max() > 1e-2:
        # help diagnose layout errors: where is it wrong?
        bad = np.argwhere(err > 1e-2)
        print('  bad count', len(bad), 'first', bad[:8].tolist())
        print('  got[0,:8]', got[0, :8], '\n  ref[0,:8]', ref[0, :8])
        rows_bad = np.unique(bad[:, 0]); cols_bad = np.unique(bad[:, 1])
        print('  bad rows', rows_bad[:16], '... cols', cols_bad[:16])
    ref = so.pairwise_similarities(xa.copy(), None, 0)
    got = h.pairwise(xa, None, 0, mode=mode, cta_group=cg)
    err = np.abs(got - ref)
    print('pairwise self  mode=%s cg=%d n=%d: max|dd|=%.3e' % (mode, cg, na, err.max()))
    ref = so.pairwise_similarities(xa.copy(), xb.copy(), 1)
    got = h.pairwise(xa, xb, 1, mode=mode, cta_group=cg)
    print('pairwise cross metric1: max|dd|=%.3e' % np.abs(got - ref).max())


def case_hist(mode, cg, n_classes=40, d=128, seed=3, pairs=1):
    h = _capi.Handle(0)
    rng = np.random.default_rng(seed)
    sizes = rng.integers(1, 40, size=n_classes)
    x, labels = so.synthetic_embeddings(sizes, dim=d, sigma=1.0, seed=seed)
    for metric in (0, 1):
        thr = so.default_thresholds(metric)
        ref = so.pair_histogram(x, labels, thr, metric)
        for force in ('fast', ):
            out = h.pair_histogram(x, labels, thr, metric, mode=mode, cta_group=cg, cluster_pairs=pairs)
            base = h.pair_histogram(x, labels, thr, metric, mode=mode, cta_group=cg)
            print('   pairs=%d bins identical to pairs=1: %s   grid %d' % (pairs, bool((out['bins'] == base['bins']).all()), out['stats']['grid_ctas']))
            ds = np.abs(out['same'] - ref['same']).sum()
            dd = np.abs(out['diff'] - ref['diff']).sum()
            print('hist mode=%s cg=%d N=%d metric=%d: n_same %d/%d n_diff %d/%d  L1(same)=%d L1(diff)=%d eps_window=%d tiles=%d kernel_ms=%.3f'
                  % (mode, cg, x.shape[0], metric, out['n_same'], ref['n_same'], out['n_diff'], ref['n_diff'], ds, dd,
                     out['stats']['eps_window'], out['stats']['tiles'], out['stats']['kernel_ms']))
            print('   range checked [%g, %g] max_abs %g' % (out['stats']['smin'], out['stats']['smax'], out['stats']['max_abs']))


def case_bench(mode, cg, n=20000, d=512, reps=3, pairs=1, region_rows=0):
    h = _capi.Handle(0)
    x, labels = so.synthetic_embeddings([50] * (n // 50), dim=d, sigma=1.1, seed=0)
    thr = so.default_thresholds(0)
    import torch
    xt = torch.from_numpy(x).cuda()
    lt = torch.from_numpy(labels).cuda()
    for r in range(reps):
        t0 = time.time()
        bins, st = h.pair_histogram_bins(xt, lt, thr, 0, mode=mode, cta_group=cg, cluster_pairs=pairs, region_rows=region_rows)
        dt = time.time() - t0
        npairs = n * (n - 1) / 2
        print('bench mode=%s cg=%d pairs=%d rr=%d grid=%d N=%d: kernel %.3f ms  prepare %.3f ms  wall %.1f ms  -> %.1f Gpairs/s (kernel)  %.1f TFLOP/s  pairs=%d'
              % (mode, cg, pairs, region_rows, st['grid_ctas'], n, st['kernel_ms'], st['prepare_ms'], dt * 1e3, npairs / st['kernel_ms'] / 1e6,
                 npairs * 1024 / st['kernel_ms'] / 1e9, st['n_pairs']))


def case_accuracy(mode, n=3000, d=512):
    """Error of a mode's distances against float64 on random, clustered and tight-cluster embeddings."""
    h = _capi.Handle(0)
    rng = np.random.default_rng(0)
    sets = {'random': unit(n, d, 1)}
    x, _ = so.synthetic_embeddings([50] * (n // 50), dim=d, sigma=1.1, seed=0)
    sets['clustered sigma=1.1'] = x
    x, _ = so.synthetic_embeddings([50] * (n // 50), dim=d, sigma=0.3, seed=0)
    sets['tight sigma=0.3'] = x
    for name, x in sets.items():
        x64 = x.astype(np.float64)
        iu = np.triu_indices(n, 1)
        exact = 2 * (1 - np.clip((x64 @ x64.T)[iu], -1, 1))
        ref32 = so.pairwise_similarities(x.copy(), None, 0)
        got = h.pairwise(x, None, 0, mode=mode)
        e = np.abs(got - exact)
        print('accuracy mode=%s %-20s n=%d: max|dd|=%.3e rms=%.3e  vs oracle(fp32 sgemm): max=%.3e   [oracle vs f64: max=%.3e]'
              % (mode, name, n, e.max(), np.sqrt((e ** 2).mean()), np.abs(got - ref32).max(), np.abs(ref32 - exact).max()))
        se = got - exact
        j = int(np.argmax(e))
        print('      bias=%.3e  p99.99=%.3e  worst pair (%d,%d): d=%.6f err=%.3e' %
              (se.mean(), np.quantile(e, 0.9999), iu[0][j], iu[1][j], exact[j], se[j]))
        near = exact < 0.5
        if near.any():
            print('      pairs with d<0.5: n=%d bias=%.3e rms=%.3e max=%.3e' % (near.sum(), se[near].mean(), np.sqrt((se[near]**2).mean()), e[near].max()))


if __name__ == '__main__':
    case = sys.argv[1]
    args = sys.argv[2:]
    conv = [int(a) if a.lstrip('-').isdigit() else a for a in args]
    globals()['case_' + case](*conv)
