"""A/B two builds of the library on the same box: python scripts/ab_kernel.py libA.so libB.so"""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, '.')
from facenet_b200 import _capi

c = ctypes
P = c.POINTER


def handle_for(path):
    lib = c.CDLL(path)
    lib.fnb_create.argtypes = [c.c_int, P(c.c_void_p)]
    lib.fnb_destroy.argtypes = [c.c_void_p]
    lib.fnb_destroy.restype = None
    lib.fnb_last_error.argtypes = [c.c_void_p]
    lib.fnb_last_error.restype = c.c_char_p
    lib.fnb_default_options.argtypes = [P(_capi.Options)]
    lib.fnb_default_options.restype = None
    lib.fnb_set_stream.argtypes = [c.c_void_p, c.c_void_p]
    lib.fnb_pair_histogram_bins.argtypes = [c.c_void_p, P(_capi.DLTensor), P(_capi.DLTensor), P(c.c_double), c.c_int, P(_capi.Options),
                                            P(_capi.DLTensor), P(_capi.Stats)]
    h = c.c_void_p()
    assert lib.fnb_create(0, c.byref(h)) == 0
    obj = object.__new__(_capi.Handle)
    obj.lib, obj.h, obj.device, obj._stream = lib, h, 0, None
    obj._borrow = lambda o: _capi.Borrowed(o)          # own stream, host-synchronous calls
    return obj


def data(n):
    g = torch.Generator(device='cuda'); g.manual_seed(0)
    ids = n // 50
    labels = (torch.arange(n, device='cuda') % ids)[torch.randperm(n, generator=g, device='cuda')]
    x = torch.randn((ids, 512), generator=g, device='cuda')[labels] + 1.1 * torch.randn((n, 512), generator=g, device='cuda')
    return (x / x.norm(dim=1, keepdim=True)).contiguous(), labels


libs = sys.argv[1:3]
hs = [handle_for(p) for p in libs]
thr = np.linspace(0, 4, 100)
for n, pairs, rr, reps in ((100000, 1, 16384, 6), (1000000, 2, 32768, 3), (100000, 1, 16384, 6)):
    x, l = data(n)
    torch.cuda.synchronize()
    out = {0: [], 1: []}
    for rep in range(reps):
        for i, h in enumerate(hs):
            _, st = h.pair_histogram_bins(x, l, thr, 0, mode='fp16f8', cluster_pairs=pairs, region_rows=rr)
            out[i].append(round(st['kernel_ms'], 3))
    for i in (0, 1):
        print('N=%d pairs=%d rr=%d  %s: kernel ms %s' % (n, pairs, rr, libs[i].split('/')[-1], out[i]))
    del x, l
