"""Signed error of an arithmetic mode as a function of the similarity (not a test).

    python scripts/probe_bias.py MODE[,MODE...] [N] [D]

Classes of different tightness (same-identity similarities from ~0.3 to ~0.99, plus antipodal rows for s < 0); every pair of
the N x D set is evaluated by fnb_pairwise (metric 0) and compared with torch float64 on the GPU.  Per similarity bin: count,
mean / rms / max of the SIMILARITY error  e_s = s_mode - s_f64  (= -dd / 2).  The mean is what the tensor core's truncating
accumulation contributes (profiles/r01_accuracy.md); the table is the calibration of GramParams::acc_scale's bias factor."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from facenet_b200 import _capi


def make_set(n, d, seed=0, kind='gauss'):
    """kind: gauss (Gaussian coordinates), relu (all coordinates >= 0, like post-ReLU features), sparse (1/8 of the coordinates
    non-zero per class), t3 (heavy-tailed Student-t coordinates)"""
    g = torch.Generator(device='cuda'); g.manual_seed(seed)
    sig = torch.tensor([0.05, 0.1, 0.15, 0.2, 0.3, 0.4, 0.5, 0.65, 0.8, 1.0, 1.25, 1.6], device='cuda')
    k = 64
    ids = n // k
    cls = torch.arange(n, device='cuda') // k
    centres = torch.randn((ids, d), generator=g, device='cuda')
    noise = torch.randn((n, d), generator=g, device='cuda')
    if kind == 't3':
        chi = (torch.randn((3, ids, d), generator=g, device='cuda') ** 2).sum(0) / 3.0
        centres = centres / chi.sqrt()
    x = centres[cls] + sig[cls % sig.numel()][:, None] * noise
    if kind == 'relu':
        x = x.abs()
    if kind == 'sparse':
        mask = (torch.rand((ids, d), generator=g, device='cuda') < 0.125).float()
        x = x * mask[cls]
    flip = (torch.arange(n, device='cuda') % 4 == 3).float() * -2 + 1        # every 4th row antipodal: s < 0 pairs
    x = x * flip[:, None]
    x = (x / x.norm(dim=1, keepdim=True)).contiguous()
    return x


def main():
    modes = sys.argv[1].split(',')
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    d = int(sys.argv[3]) if len(sys.argv) > 3 else 512
    kind = sys.argv[4] if len(sys.argv) > 4 else 'gauss'
    nb = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    raw = len(sys.argv) > 6 and sys.argv[6] == 'raw'      # raw: the library's bias correction switched off (calibration runs)
    h = _capi.default_handle(0)
    x = make_set(n, d, kind=kind)
    print('data kind %s, bias correction %s' % (kind, 'OFF' if raw else 'on'))
    x64 = x.double()
    iu = torch.triu_indices(n, n, 1, device='cuda')
    s64 = (x64 @ x64.T)[iu[0], iu[1]]
    edges = torch.linspace(-1.0, 1.0, nb + 1, device='cuda', dtype=torch.float64)
    b = torch.bucketize(s64, edges).clamp_(1, nb) - 1
    out = {}
    for mode in modes:
        dm = torch.empty(n * (n - 1) // 2, device='cuda', dtype=torch.float32)
        h.pairwise(x, None, 0, mode=mode, out=dm, bias_correction=-1 if raw else 0)
        torch.cuda.synchronize()
        es = (1.0 - dm.double() / 2.0) - s64.clamp(-1.0, 1.0)          # similarity error (d = 2 (1 - s))
        rows = []
        print('mode %s  N=%d D=%d  pairs=%d   overall: mean %.3e rms %.3e max|e_s| %.3e (max|dd| %.3e)' %
              (mode, n, d, es.numel(), es.mean().item(), es.pow(2).mean().sqrt().item(), es.abs().max().item(), 2 * es.abs().max().item()))
        print('   s_lo    s_hi       count        mean         rms(centred)  max|e_s|    mean/s')
        for k in range(nb):
            m = b == k
            c = int(m.sum().item())
            if c < 50:
                continue
            e = es[m]
            mu = e.mean().item()
            sd = (e - mu).pow(2).mean().sqrt().item()
            mx = e.abs().max().item()
            sc = 0.5 * (edges[k] + edges[k + 1]).item()
            rows.append({'s_lo': edges[k].item(), 's_hi': edges[k + 1].item(), 'count': c, 'mean': mu, 'std': sd, 'max_abs': mx})
            print('  %6.2f  %6.2f  %10d  %12.4e  %12.4e  %10.3e  %10.3e' % (edges[k].item(), edges[k + 1].item(), c, mu, sd, mx, mu / sc if abs(sc) > 1e-9 else 0))
        out[mode] = rows
        # least-squares slope of the mean error against s over the bins with |s| > 0.2 (bias ~ -beta * s)
        sel = [(0.5 * (r['s_lo'] + r['s_hi']), r['mean'], r['count']) for r in rows if abs(0.5 * (r['s_lo'] + r['s_hi'])) > 0.2]
        if sel:
            sv = np.array([v[0] for v in sel]); mv = np.array([v[1] for v in sel]); wv = np.array([v[2] for v in sel], dtype=np.float64)
            beta = -(wv * sv * mv).sum() / (wv * sv * sv).sum()
            resid = mv + beta * sv
            print('   fit: e_s ~ -beta * s, beta = %.4e; largest residual of the bin means %.3e' % (beta, np.abs(resid).max()))
            out[mode + '_beta'] = beta
        del dm, es
    print(json.dumps(out))


if __name__ == '__main__':
    main()
