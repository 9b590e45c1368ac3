"""NumPy stand-in for the CUDA library's histogram entry points, used ONLY by the CPU tests to
exercise the host-side logic of facenet_b200.statistics (row plan, rectangles, cuts, weights)
where no GPU exists.  It follows the C ABI contract of include/facenet_b200.h literally."""
import numpy as np


def region_histogram_bins(embeddings, perm, cls, regions, nkeys, thresholds, metric=0, cuts=None, **_):
    x = np.ascontiguousarray(embeddings, dtype=np.float32)[np.asarray(perm)]
    cls = np.asarray(cls)
    assert np.all(np.diff(cls) >= 0), 'cls must be non-decreasing'
    order = np.sort(np.asarray(cuts, dtype=np.float32))
    nt = order.size
    bins = np.zeros((nkeys, 2, nt + 1), dtype=np.uint64)
    for r in regions:
        r0, r1, c0, c1 = int(r['row_begin']), int(r['row_end']), int(r['col_begin']), int(r['col_end'])
        if r1 <= r0 or c1 <= c0:
            continue
        s = np.clip(x[r0:r1] @ x[c0:c1].T, -1, 1)
        k = np.searchsorted(order, s.ravel(), side='right').reshape(s.shape)
        valid = np.ones(s.shape, dtype=bool)
        if r['tri']:
            assert (r0, r1) == (c0, c1)
            valid = np.arange(c0, c1)[None, :] > np.arange(r0, r1)[:, None]
        same = (cls[r0:r1, None] == cls[None, c0:c1]) & valid
        bins[r['key'], 0] += np.bincount(k[valid], minlength=nt + 1).astype(np.uint64)
        bins[r['key'], 1] += np.bincount(k[same], minlength=nt + 1).astype(np.uint64)
    return bins, {'emulated': True}


class EmulatedHandle:
    def region_histogram_bins(self, *a, **kw):
        kw.pop('mode', None)
        kw.pop('cta_group', None)
        return region_histogram_bins(*a, **kw)
