"""NumPy stand-in for the CUDA library's histogram entry points, used ONLY by the CPU tests to
exercise the host-side logic of facenet_b200.statistics (row plan, rectangles, cuts, weights)
where no GPU exists.  It follows the C ABI contract of include/facenet_b200.h literally."""
import numpy as np


def region_histogram_bins(embeddings, perm, cls, regions, nkeys, thresholds, metric=0, cuts=None, raw_distance=False,
                          normalize=0, theta=0.0, **_):
    x = np.ascontiguousarray(embeddings, dtype=np.float32)[np.asarray(perm)]
    nrm = None
    if normalize == 1:                                   # fnb_options.normalize = 1 (+ theta): faceclass.py:57-71
        nrm = np.linalg.norm(x, axis=1, keepdims=True)
        x = x / nrm
    elif normalize == 2:                                 # tf.nn.l2_normalize(axis=1, epsilon=1e-10)
        x = (x * (1.0 / np.sqrt(np.maximum((x * x).sum(axis=1, keepdims=True), np.float32(1e-10))))).astype(np.float32)
    cls = np.asarray(cls)
    assert np.all(np.diff(cls) >= 0), 'cls must be non-decreasing'
    if cuts is None or isinstance(cuts, str):            # the library's default: NumPy-exact cuts
        from facenet_b200 import _capi
        cuts = _capi.numpy_cuts(thresholds, metric)
    order = np.sort(np.asarray(cuts, dtype=np.float32))
    nt = order.size
    bins = np.zeros((nkeys, 2, nt + 1), dtype=np.uint64)
    for r in regions:
        r0, r1, c0, c1 = int(r['row_begin']), int(r['row_end']), int(r['col_begin']), int(r['col_end'])
        if r1 <= r0 or c1 <= c0:
            continue
        s = x[r0:r1] @ x[c0:c1].T
        if nrm is not None and theta != 0.0:
            g = 2 * (nrm[r0:r1] - nrm[c0:c1].T) / (nrm[r0:r1] + nrm[c0:c1].T)
            s = np.float32(1) - np.float32(0.5) * (2 * (1 - s) + np.float32(theta) * g * g)
        elif not raw_distance:
            s = np.clip(s, -1, 1)
        k = np.searchsorted(order, s.ravel(), side='right').reshape(s.shape)
        valid = np.ones(s.shape, dtype=bool)
        if r['tri']:
            assert (r0, r1) == (c0, c1)
            valid = np.arange(c0, c1)[None, :] > np.arange(r0, r1)[:, None]
        same = (cls[r0:r1, None] == cls[None, c0:c1]) & valid
        bins[r['key'], 0] += np.bincount(k[valid], minlength=nt + 1).astype(np.uint64)
        bins[r['key'], 1] += np.bincount(k[same], minlength=nt + 1).astype(np.uint64)
        if r['tri'] == 2:                                # diagonal elements -> slot key + 1 (include/facenet_b200.h)
            bins[r['key'] + 1, 0] += np.bincount(np.diagonal(k), minlength=nt + 1).astype(np.uint64)
    return bins, {'emulated': True}


def confidence_from_bins(bins, w_same, w_diff, thresholds, cuts, far_target):
    """NumPy statement of fnb_confidence_from_last_bins (include/facenet_b200.h)."""
    thr = np.asarray(thresholds, dtype=np.float64)
    order = np.sort(np.asarray(cuts, dtype=np.float32))
    pos = np.searchsorted(order, cuts, side='right')
    b = bins.astype(np.int64)
    suffix = np.concatenate([np.cumsum(b[..., ::-1], axis=-1)[..., ::-1], np.zeros(b.shape[:-1] + (1,), dtype=np.int64)], axis=-1)
    lt = suffix[..., pos]                                            # [keys, 2, T]
    tot = suffix[..., 0]
    ws = np.asarray(w_same, dtype=np.float64)[:, None]
    wd = np.asarray(w_diff, dtype=np.float64)[:, None]
    same_lt, diff_lt = lt[:, 1], lt[:, 0] - lt[:, 1]
    same_tot, diff_tot = tot[:, 1][:, None], (tot[:, 0] - tot[:, 1])[:, None]
    tp = (same_lt * ws).sum(axis=0)
    fn = ((same_tot - same_lt) * ws).sum(axis=0)
    fp = (diff_lt * wd).sum(axis=0)
    tn = ((diff_tot - diff_lt) * wd).sum(axis=0)
    with np.errstate(invalid='ignore', divide='ignore'):
        acc = (tp + tn) / (tp + fp + tn + fn)
        tnr = np.where((tn + fp) > 0, tn / np.where((tn + fp) > 0, tn + fp, 1), 1.0)
    fpr = 1 - tnr
    far = 0.0
    if fpr.max() >= far_target:
        if thr.size < 2 or far_target < fpr[0] or far_target > fpr[-1]:
            far = float('nan')
        else:
            j = min(max(int(np.searchsorted(fpr, far_target, side='right')) - 1, 0), thr.size - 2)
            with np.errstate(divide='ignore', invalid='ignore'):
                w = np.float64(1.0) / (fpr[j + 1] - fpr[j])                     # the k = 1 B-spline weights (select_kernel)
                far = thr[j] * ((fpr[j + 1] - far_target) * w) + thr[j + 1] * ((far_target - fpr[j]) * w)
    return {'tp': tp, 'tn': tn, 'fp': fp, 'fn': fn, 'argmax_accuracy': int(np.argmax(acc)), 'far_threshold': float(far)}


class EmulatedHandle:
    def region_histogram_bins(self, *a, **kw):
        kw.pop('mode', None)
        kw.pop('cta_group', None)
        self.last_bins, st = region_histogram_bins(*a, **kw)
        return self.last_bins, st

    def confidence_from_last_bins(self, nkeys, w_same, w_diff, thresholds, metric=0, far_target=0.0, cuts=None, **_):
        assert self.last_bins.shape[0] == nkeys
        return confidence_from_bins(self.last_bins, w_same, w_diff, thresholds, cuts, far_target)
