"""Drop-in replacement for the evaluation-statistics path of sMedX/FaceNet
(``/root/reference/facenet/statistics.py``), computed on a B200.

Same names, positional order, defaults, attributes and exceptions as the reference:

    pairwise_similarities(xa, xb=None, metric=0, atol=1.e-5)          statistics.py:22-57
    split_embeddings(embeddings, labels)                              statistics.py:68-79
    SimilarityCalculator(embeddings, labels, metric=0)                statistics.py:82-108
    ConfidenceMatrix(calculator, threshold)                           statistics.py:111-175
    Report(criterion=None)                                            statistics.py:178-234
    FaceToFaceValidation(embeddings, labels, config)                  statistics.py:237-331

What runs where
  * every pair distance, threshold comparison and count runs in the CUDA library
    (``include/facenet_b200.h``) -- one call per ``ConfidenceMatrix``, never one per class pair;
  * this module only prepares the call (class ranks, row order, rectangles of the pair matrix,
    similarity cuts) and turns the integer histograms into the reference's float64
    class-balanced rates;
  * ``Report`` (AUC / EER / mean +- std over folds, statistics.py:187-234) is host arithmetic on
    ~100 numbers and stays in Python like the reference.

There is no CPU fallback: without the shared library or a CUDA device these functions raise.
"""
import datetime
import time
from pathlib import Path

import numpy as np

from facenet_b200 import _capi

__all__ = ['pairwise_similarities', 'split_embeddings', 'SimilarityCalculator', 'ConfidenceMatrix', 'Report',
           'FaceToFaceValidation', 'FalseExamples', 'pair_histogram', 'mean', 'std', 'set_default_mode', 'kfold_split']

_state = {'mode': 'fp16x3', 'device': 0, 'cta_group': 0}


def set_default_mode(mode=None, device=None, cta_group=None):
    """Select the Gram arithmetic ('fp16x3' default; 'tf32x3', 'fp16f8', 'auto' -- fp16f8 where its error model holds, else fp16x3
    --, and the single-pass 'tf32', 'bf16', 'fp16'), the CUDA device and the tile variant used by the functions of this module."""
    if mode is not None:
        if mode not in _capi.MODES:
            raise ValueError('unknown mode {}'.format(mode))
        _state['mode'] = mode
    if device is not None:
        _state['device'] = int(device)
    if cta_group is not None:
        _state['cta_group'] = int(cta_group)


def _handle():
    return _capi.default_handle(_state['device'])


def _raise_like_reference(err, metric=None):
    if err.code == _capi.FNB_ERR_NOT_NORMALIZED:
        # statistics.py:42
        msg = str(err)
        lo, hi = msg.split('range')[-1].split()
        raise ValueError('\nembeddings must be normalized to 1, range {} {}'.format(np.float32(lo), np.float32(hi))) from None
    if err.code == _capi.FNB_ERR_BAD_METRIC:
        # statistics.py:55
        raise ValueError('Undefined similarity metric {}'.format(metric)) from None
    raise err


def mean(x):
    return np.mean(np.array(x))


def std(x):
    return np.std(np.array(x))


def _pad64(x):
    """The tensor-core tiles take embedding dimensions that are multiples of 64 (64 .. 4096); the reference takes any D.  Zero
    columns change no dot product and no norm, so other dimensions are padded here (a copy; 512-d embeddings pass through)."""
    d = int(x.shape[1])
    dp = max(64, -(-d // 64) * 64)
    if dp == d or x.shape[0] == 0:
        return x
    if isinstance(x, np.ndarray):
        out = np.zeros((x.shape[0], dp), dtype=np.float32)
        out[:, :d] = x
        return out
    if isinstance(x, _capi.DLPackTensor):
        raise ValueError('a raw DLPack capsule must carry an embedding dimension that is a multiple of 64 (got {})'.format(d))
    import torch
    return torch.nn.functional.pad(x if hasattr(x, 'storage') else torch.from_dlpack(x), (0, dp - d)).contiguous()


def pairwise_similarities(xa, xb=None, metric=0, atol=1.e-5):
    """Evaluate pairwise distances between vectors xa and xb (statistics.py:22-57).

    ``xb is None``: 1-D float32 array of the strict upper triangle of the Gram matrix in row-major
    ``np.triu_indices(n, k=1)`` order; otherwise the 2-D ``[n_a, n_b]`` matrix.  metric 0 gives
    ``2 * (1 - s)``, metric 1 ``arccos(s)``.  Raises ``ValueError`` exactly where the reference does."""
    xa = np.ascontiguousarray(xa, dtype=np.float32)        # (float64 input is computed in float32 here, in float64 by NumPy)
    if xb is not None:
        xb = np.ascontiguousarray(xb, dtype=np.float32)
    n_out = xa.shape[0] * (xa.shape[0] - 1) // 2 if xb is None else xa.shape[0] * xb.shape[0]
    if n_out and xa.ndim == 2:
        xa = _pad64(xa)
        if xb is not None and xb.ndim == 2:
            xb = _pad64(xb)
    if n_out == 0:
        # statistics.py:38 -- empty in, empty out (no checks)
        return np.empty((0,) if xb is None else (xa.shape[0], xb.shape[0]), dtype=np.float32)
    if metric not in (0, 1):
        # the reference range-checks first (statistics.py:40-42) and only then rejects the metric (:55)
        try:
            _handle().pairwise(xa, xb, 0, atol, mode=_state['mode'], cta_group=_state['cta_group'])
        except _capi.FnbError as err:
            _raise_like_reference(err, metric)
        raise ValueError('Undefined similarity metric {}'.format(metric))
    try:
        return _handle().pairwise(xa, xb, metric, atol, mode=_state['mode'], cta_group=_state['cta_group'])
    except _capi.FnbError as err:
        _raise_like_reference(err, metric)


def split_embeddings(embeddings, labels):
    """split embeddings to structure [[], [], ...[]] in sorted ``np.unique(labels)`` order (statistics.py:68-79)."""
    embeddings = np.asarray(embeddings)
    labels = np.asarray(labels)
    order = np.argsort(labels, kind='stable')
    _, counts = np.unique(labels, return_counts=True)
    return np.split(embeddings[order], np.cumsum(counts)[:-1])


def _on_gpu(x):
    """True for a tensor that lives in CUDA memory and speaks DLPack (torch, TensorFlow via tf.experimental.dlpack, CuPy)."""
    if isinstance(x, np.ndarray) or not hasattr(x, '__dlpack_device__'):
        return False
    try:
        return int(x.__dlpack_device__()[0]) == 2          # kDLCUDA
    except Exception:
        return False


def _host_array(x):
    if isinstance(x, _capi.DLPackTensor):
        # labels handed over as a capsule: a small host copy through torch (any DLPack consumer would do)
        n = int(np.prod(x.shape))
        if not x.is_cuda:
            ct = {(0, 32): np.int32, (0, 64): np.int64, (2, 32): np.float32, (2, 64): np.float64}[(x.dtype_code, x.dtype_bits)]
            import ctypes
            t = x.ptr.contents
            buf = (ctypes.c_char * (n * x.dtype_bits // 8)).from_address(t.data + t.byte_offset)
            return np.frombuffer(buf, dtype=ct).reshape(x.shape).copy()
        raise TypeError('labels in a raw DLPack capsule must live on the host (hand GPU labels over as a tensor object)')
    return np.asarray(x.cpu() if hasattr(x, 'cpu') else x)


class SimilarityCalculator:
    """Class to evaluate similarities according to defined metric (statistics.py:82-108).

    Holds the whole subset; ``ConfidenceMatrix`` hands it to the GPU in one call.  ``.embeddings`` (the
    per-class list of the reference) is materialised lazily for callers that index it.

    Beyond the reference (the hand-off of ``facenet.evaluate_embeddings``, facenet.py:184-201):
    ``embeddings`` may be a float32 tensor that is already on the GPU (torch / TensorFlow through DLPack) -- it is used in
    place, never copied to the host; ``_rows`` selects the rows of it this calculator covers (one fold of a validation)
    and ``normalize=True`` applies ``tf.nn.l2_normalize(axis=1, epsilon=1e-10)`` (inception_resnet_v1.py:491-492) while
    the operands are prepared, so raw network outputs can be handed over."""

    def __init__(self, embeddings, labels, metric=0, _rows=None, normalize=False):
        self.metric = metric
        embeddings = _capi.from_dlpack(embeddings)   # a raw "dltensor" capsule (tf.experimental.dlpack.to_dlpack) is taken over
        labels = _capi.from_dlpack(labels)
        self._gpu = _on_gpu(embeddings)
        self._x = embeddings if (self._gpu or isinstance(embeddings, _capi.DLPackTensor)) else np.ascontiguousarray(embeddings, dtype=np.float32)
        self._dim = int(self._x.shape[1]) if len(self._x.shape) == 2 else None
        if len(self._x.shape) == 2:
            self._x = _pad64(self._x)
        self._rows = None if _rows is None else np.ascontiguousarray(_rows, dtype=np.int64)
        self._normalize = 2 if normalize else 0
        self._labels = _host_array(labels)
        self._n = int(self._x.shape[0]) if self._rows is None else int(self._rows.size)
        if self._n != len(self._labels):
            raise ValueError('embeddings and labels have different lengths')
        values, self._cls, self._sizes = np.unique(self._labels, return_inverse=True, return_counts=True)
        self._cls = np.asarray(self._cls).reshape(-1)
        self._split = None

    @property
    def embeddings(self):
        if self._split is None:
            if isinstance(self._x, _capi.DLPackTensor):
                raise TypeError('.embeddings needs a tensor object (torch / NumPy); a raw DLPack capsule is consumed by the GPU path only')
            x = _host_array(self._x) if self._gpu else self._x
            if self._rows is not None:
                x = x[self._rows]
            if self._normalize:
                x = (x * (1.0 / np.sqrt(np.maximum((x.astype(np.float32) ** 2).sum(axis=1, keepdims=True), np.float32(1e-10))))).astype(np.float32)
            order = np.argsort(self._cls, kind='stable')
            self._split = np.split(x[order][:, :self._dim], np.cumsum(self._sizes)[:-1])
        return self._split

    def evaluate(self, i, k):
        # statistics.py:90-101
        nrof_positive_class_pairs = self.nrof_classes
        nrof_negative_class_pairs = self.nrof_classes * (self.nrof_classes - 1) / 2
        if i == k:
            sims = pairwise_similarities(self.embeddings[i], metric=self.metric)
            weight = sims.size * nrof_positive_class_pairs
        else:
            sims = pairwise_similarities(self.embeddings[i], self.embeddings[k], metric=self.metric)
            weight = sims.size * nrof_negative_class_pairs
        return sims, weight

    @property
    def nrof_classes(self):
        return int(self._sizes.size)

    def nrof_images(self, i):
        return int(self._sizes[i])


def _size_group_plan(cls, sizes):
    """Row order and rectangles for the class-balanced confidence matrix.

    The reference weights every (class i, class k) block by ``1 / (block size * number of class
    pairs)`` (statistics.py:91-99,133-138); the weight depends on the classes only through their
    SIZES.  Rows are therefore ordered by (class size, class rank): classes of equal size become
    contiguous, each pair of size groups (a <= b) is one rectangle of the pair matrix with its own
    histogram slot, and the integer counts per slot are exact."""
    uniq_sizes, group_of_class = np.unique(sizes, return_inverse=True)
    group_of_class = np.asarray(group_of_class).reshape(-1)
    n_groups = uniq_sizes.size
    # new class rank: ordered by (size group, old class rank)
    class_order = np.lexsort((np.arange(sizes.size), group_of_class))
    new_rank = np.empty(sizes.size, dtype=np.int64)
    new_rank[class_order] = np.arange(sizes.size)
    row_rank = new_rank[cls]
    perm = np.argsort(row_rank, kind='stable')
    cls_sorted = row_rank[perm].astype(np.int32)
    rows_per_group = np.bincount(group_of_class, weights=sizes, minlength=n_groups).astype(np.int64)
    bounds = np.concatenate([[0], np.cumsum(rows_per_group)])
    classes_per_group = np.bincount(group_of_class, minlength=n_groups).astype(np.int64)

    ia, ib = np.triu_indices(n_groups)
    regions = np.zeros(ia.size, dtype=_capi.REGION_DTYPE)
    regions['row_begin'] = bounds[ia]
    regions['row_end'] = bounds[ia + 1]
    regions['col_begin'] = bounds[ib]
    regions['col_end'] = bounds[ib + 1]
    regions['tri'] = (ia == ib)
    regions['key'] = np.arange(ia.size)
    return perm, cls_sorted, regions, ia, ib, uniq_sizes.astype(np.int64), classes_per_group


def _counts_lt(bins, cuts):
    """bins [..., T+1] over ascending-cut bins -> counts of pairs with ``d < threshold_n`` [..., T]."""
    order = np.sort(cuts)
    pos = np.searchsorted(order, cuts, side='right')            # cuts <= cut_n
    suffix = np.cumsum(bins[..., ::-1].astype(np.int64), axis=-1)[..., ::-1]
    suffix = np.concatenate([suffix, np.zeros(suffix.shape[:-1] + (1,), dtype=np.int64)], axis=-1)
    return suffix[..., pos]


class ConfidenceMatrix:
    """Class to evaluate confidence matrix (tp, tn, fp, fn) and others metrics (statistics.py:111-175).

    ``tp[n] = (1/C) * sum_i count_ii(n) / (n_i (n_i - 1) / 2)`` and
    ``fp[n] = (2 / (C (C-1))) * sum_{i>k} count_ik(n) / (n_i n_k)`` with
    ``count(n) = #{pairs : d < threshold_n}`` (strict, float64 compare).  The integer counts come from
    one fused Gram + histogram launch; the float64 rates are formed here."""

    def __init__(self, calculator, threshold, _far_target=None):
        self.threshold = np.array(threshold, ndmin=1)
        nt = self.threshold.size
        self.tp = np.zeros(nt)
        self.tn = np.zeros(nt)
        self.fp = np.zeros(nt)
        self.fn = np.zeros(nt)
        self.stats = None
        # filled by the device selection kernel when the whole threshold grid went through one launch
        self._argmax_accuracy = None
        self._far_threshold = None
        if not hasattr(calculator, '_cls'):
            # a reference-style calculator (facenet/statistics.py:82-108: ``.embeddings`` = list of per-class arrays, ``.metric``,
            # ``.nrof_classes``): rebuild the row set from the list -- still ONE launch for the whole matrix
            parts = [np.asarray(e, dtype=np.float32).reshape(-1, np.asarray(e).shape[-1]) if np.asarray(e).size else
                     np.zeros((0, 1), dtype=np.float32) for e in calculator.embeddings]
            dim = max((p.shape[1] for p in parts if p.shape[0]), default=1)
            rows = np.concatenate([p if p.shape[0] else np.zeros((0, dim), dtype=np.float32) for p in parts]) if parts else np.zeros((0, dim), np.float32)
            lab = np.repeat(np.arange(len(parts)), [p.shape[0] for p in parts])
            calculator = SimilarityCalculator(rows, lab, metric=calculator.metric)
        if nt == 0 or calculator._n < 2:
            return
        thr = self.threshold.astype(np.float64).reshape(-1)
        metric = calculator.metric
        if metric not in (0, 1):
            raise ValueError('Undefined similarity metric {}'.format(metric))
        nc = calculator.nrof_classes
        perm, cls_sorted, regions, ia, ib, gsize, gcount = _size_group_plan(calculator._cls, calculator._sizes)
        # weight of one pair of the rectangle (size group a, size group b): statistics.py:91-99,133-138
        #   same identity (diagonal rectangles only): 1 / (n (n - 1) / 2 * C)
        #   different identity:                       1 / (n_a * n_b * C (C - 1) / 2)
        npairs = gsize[ia] * (gsize[ia] - 1) / 2
        w_same = np.zeros(ia.size)
        ok = (ia == ib) & (npairs > 0)
        w_same[ok] = 1.0 / (npairs[ok] * nc)
        w_diff = 1.0 / (gsize[ia] * gsize[ib] * (nc * (nc - 1) / 2)) if nc > 1 else np.zeros(ia.size)
        w_diff = np.asarray(w_diff, dtype=np.float64)
        h = _handle()
        for t0 in range(0, nt, _capi.MAX_THRESHOLDS):
            sl = slice(t0, min(nt, t0 + _capi.MAX_THRESHOLDS))
            cuts = _capi.numpy_cuts(thr[sl], metric)
            try:
                rows = calculator._rows
                bins, self.stats = h.region_histogram_bins(calculator._x, perm if rows is None else rows[perm], cls_sorted,
                                                           regions, regions.size, thr[sl], metric=metric, mode=_state['mode'],
                                                           cta_group=_state['cta_group'], cuts=cuts,
                                                           subset=rows is not None, normalize=calculator._normalize)
                sel = h.confidence_from_last_bins(regions.size, w_same, w_diff, thr[sl], metric=metric, cuts=cuts,
                                                  far_target=0.0 if _far_target is None else float(_far_target))
            except _capi.FnbError as err:
                _raise_like_reference(err, metric)
            self.tp[sl], self.tn[sl], self.fp[sl], self.fn[sl] = sel['tp'], sel['tn'], sel['fp'], sel['fn']
            if nt <= _capi.MAX_THRESHOLDS:
                self._argmax_accuracy = sel['argmax_accuracy']
                if _far_target is not None:
                    self._far_threshold = sel['far_threshold']

    @property
    def accuracy(self):
        return (self.tp + self.tn) / (self.tp + self.fp + self.tn + self.fn)

    @property
    def precision(self):
        i = (self.tp + self.fp) > 0
        precision = np.ones(self.threshold.size)
        precision[i] = self.tp[i] / (self.tp[i] + self.fp[i])
        return precision

    @property
    def tp_rates(self):
        # true positive rate, validation rate, sensitivity or recall
        i = (self.tp + self.fn) > 0
        tp_rates = np.ones(self.threshold.size)
        tp_rates[i] = self.tp[i] / (self.tp[i] + self.fn[i])
        return tp_rates

    @property
    def tn_rates(self):
        # true negative rate, 1 - false alarm rate, specificity
        i = (self.tn + self.fp) > 0
        tn_rates = np.ones(self.threshold.size)
        tn_rates[i] = self.tn[i] / (self.tn[i] + self.fp[i])
        return tn_rates

    @property
    def fp_rates(self):
        # false positive rate, false alarm rate
        return 1 - self.tn_rates

    @property
    def fn_rates(self):
        # false negative rate,
        return 1 - self.tp_rates


def _argsort_numpy_scalar(v):
    """The order ``np.argsort`` gives under the reference's pinned numpy 1.19.4 (requirements.txt:9): the scalar introsort --
    median-of-3 quicksort down to 16 elements, then insertion sort -- which is NOT stable, so runs of equal keys come out
    permuted (newer numpy sorts with SIMD kernels on AVX2 / AVX-512 hosts and orders ties differently: the result must not
    depend on the host)."""
    v = np.asarray(v, dtype=np.float64)
    num = v.size
    tosort = list(range(num))
    SMALL = 15
    def lt(a, b):      # npy DOUBLE_LT: a < b || (b != b && a == a)
        return a < b or (b != b and a == a)
    pl, pr = 0, num - 1
    stack = []
    depth_stack = []
    cdepth = (num.bit_length() - 1) * 2 if num > 0 else 0
    while True:
        heap = False
        if cdepth < 0:
            # heapsort fallback (never reached for the short arrays this is used on)
            sub = sorted(tosort[pl:pr + 1], key=lambda i: (v[i] != v[i], v[i]))
            tosort[pl:pr + 1] = sub
            heap = True
        if not heap:
            while (pr - pl) > SMALL:
                pm = pl + ((pr - pl) >> 1)
                if lt(v[tosort[pm]], v[tosort[pl]]): tosort[pm], tosort[pl] = tosort[pl], tosort[pm]
                if lt(v[tosort[pr]], v[tosort[pm]]): tosort[pr], tosort[pm] = tosort[pm], tosort[pr]
                if lt(v[tosort[pm]], v[tosort[pl]]): tosort[pm], tosort[pl] = tosort[pl], tosort[pm]
                vp = v[tosort[pm]]
                pi, pj = pl, pr - 1
                tosort[pm], tosort[pj] = tosort[pj], tosort[pm]
                while True:
                    pi += 1
                    while lt(v[tosort[pi]], vp): pi += 1
                    pj -= 1
                    while lt(vp, v[tosort[pj]]): pj -= 1
                    if pi >= pj: break
                    tosort[pi], tosort[pj] = tosort[pj], tosort[pi]
                pk = pr - 1
                tosort[pi], tosort[pk] = tosort[pk], tosort[pi]
                if pi - pl < pr - pi:
                    stack.append((pi + 1, pr)); pr = pi - 1
                else:
                    stack.append((pl, pi - 1)); pl = pi + 1
                cdepth -= 1
                depth_stack.append(cdepth)
            for pi in range(pl + 1, pr + 1):
                vi = tosort[pi]; vp = v[vi]; pj = pi
                while pj > pl and lt(vp, v[tosort[pj - 1]]):
                    tosort[pj] = tosort[pj - 1]; pj -= 1
                tosort[pj] = vi
        if not stack: break
        pl, pr = stack.pop()
        cdepth = depth_stack.pop()
    return np.asarray(tosort, dtype=np.int64)


def _slinear(x, y, xq):
    """``scipy.interpolate.interp1d(x, y, kind='slinear')(xq)`` as the reference's pinned scipy 1.4.1 evaluated it
    (statistics.py:301-302): samples sorted by ``np.argsort(x)`` (see ``_argsort_numpy_scalar``), a k = 1 B-spline on the knots
    ``r_[x[0], x, x[-1]]``, evaluated on the last span whose left end is <= xq with de Boor's weights.  (scipy >= 1.10 rejects
    the duplicate abscissae fp_rates always contains.)"""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    ind = _argsort_numpy_scalar(x)
    xs, ys = x[ind], y[ind]
    if xq < xs[0]:
        raise ValueError('A value in x_new is below the interpolation range.')
    if xq > xs[-1]:
        raise ValueError('A value in x_new is above the interpolation range.')
    j = int(np.searchsorted(xs, xq, side='right')) - 1
    j = min(max(j, 0), xs.size - 2)
    xa, xb = xs[j], xs[j + 1]
    with np.errstate(divide='ignore', invalid='ignore'):
        w = np.float64(1.0) / (xb - xa)
        return np.array(ys[j] * ((xb - xq) * w) + ys[j + 1] * ((xq - xa) * w))


class Report:
    """Class to generate statistical report (statistics.py:178-234)."""

    def __init__(self, criterion=None):
        self.criterion = criterion
        self.conf_matrix_train = []
        self.conf_matrix_test = []

    def __repr__(self):
        dct = self.dict
        info = self.criterion + '\n'
        info += ('Area under curve (AUC): {:1.5f}\n'.format(dct['auc']) +
                 'Equal error rate (EER): {:1.5f}\n'.format(dct['eer']) + '\n')
        info += ('Accuracy:  {:2.5f}+-{:2.5f}\n'.format(dct['accuracy'], dct['accuracy_std']) +
                 'Precision: {:2.5f}+-{:2.5f}\n'.format(dct['precision'], std(dct['precision_std'])) +
                 'Sensitivity (TPR, 1-a type 1 error): {:2.5f}+-{:2.5f}\n'.format(dct['tp_rates'], dct['tp_rates_std']) +
                 'Specificity (TNR, 1-b type 2 error): {:2.5f}+-{:2.5f}\n'.format(dct['tn_rates'], dct['tn_rates_std']) +
                 'Threshold: {:2.5f}+-{:2.5f}\n'.format(dct['threshold'], dct['threshold_std']) + '\n')
        return info

    def append_fold(self, name, conf_matrix):
        if name == 'train':
            self.conf_matrix_train.append(conf_matrix)
        else:
            self.conf_matrix_test.append(conf_matrix)

    @property
    def dict(self):
        import sklearn.metrics
        from scipy import interpolate
        from scipy.optimize import brentq

        tp_rates = np.mean(np.array([m.tp_rates for m in self.conf_matrix_train]), axis=0)
        tn_rates = np.mean(np.array([m.tn_rates for m in self.conf_matrix_train]), axis=0)

        dct = {'auc': -1, 'eer': -1}
        try:
            dct['auc'] = sklearn.metrics.auc(1 - tn_rates, tp_rates)
        except Exception:
            pass
        try:
            dct['eer'] = brentq(lambda x: 1. - x - interpolate.interp1d(1 - tn_rates, tp_rates)(x), 0., 1.)
        except Exception:
            pass

        for key in ('accuracy', 'precision', 'tp_rates', 'tn_rates', 'threshold'):
            x = [getattr(m, key) for m in self.conf_matrix_test]
            dct[key] = np.mean(x)
            dct[key + '_std'] = np.std(x)
        return dct


def kfold_split(n, n_splits, seed=0):
    """The index sets of ``sklearn.model_selection.KFold(n_splits, shuffle=True, random_state=seed)
    .split(np.arange(n))`` (statistics.py:278-287) without the sklearn dependency: a
    ``RandomState(seed)`` shuffle cut into folds, the first ``n % n_splits`` one longer.  Raises ``ValueError`` where
    sklearn's KFold does (fewer than 2 splits, more splits than samples)."""
    n_splits = int(n_splits)
    if n_splits < 2:
        raise ValueError('k-fold cross-validation requires at least one train/test split by setting n_splits=2 or more, '
                         'got n_splits={0}.'.format(n_splits))
    if n_splits > n:
        raise ValueError('Cannot have number of splits n_splits={0} greater than the number of samples: n_samples={1}.'.format(n_splits, n))
    perm = np.arange(n)
    np.random.RandomState(seed).shuffle(perm)
    fold_sizes = np.full(n_splits, n // n_splits, dtype=np.int64)
    fold_sizes[:n % n_splits] += 1
    start = 0
    for size in fold_sizes:
        test_mask = np.zeros(n, dtype=bool)
        test_mask[perm[start:start + size]] = True
        yield np.nonzero(~test_mask)[0], np.nonzero(test_mask)[0]
        start += size


class FaceToFaceValidation:
    """Class to perform face-to-face validation (statistics.py:237-331)."""

    def __init__(self, embeddings, labels, config):
        self.elapsed_time = time.monotonic()
        # raw "dltensor" capsules (tf.experimental.dlpack.to_dlpack / torch.utils.dlpack.to_dlpack) are taken over per protocol
        embeddings = _capi.from_dlpack(embeddings)
        labels = _capi.from_dlpack(labels)
        self.embeddings = embeddings
        self.labels = labels

        assert (embeddings.shape[0] == len(labels))

        self.config = config
        self.reports = None

        if self.config.metric == 0:
            upper_threshold = 4
        elif self.config.metric == 1:
            upper_threshold = np.pi
        else:
            raise ValueError('Undefined similarity metric {}'.format(self.config.metric))

        self.thresholds = np.linspace(0, upper_threshold, 100)
        self._evaluate()
        try:
            from loguru import logger
            logger.info(self)
        except ImportError:
            pass

    def __repr__(self):
        info = (f'{self.__class__.__name__}\n' +
                f'metric: {self.config.metric}\n\n')
        for r in self.reports:
            info += str(r)
        info += f'elapsed_time: {self.elapsed_time}\n'
        return info

    def _evaluate(self):
        # embeddings that already live on the GPU (torch / TensorFlow through DLPack) stay there: every fold is a row
        # subset of the one resident tensor (no D2H -> H2D round trip, facenet.py:184-201); ``config.normalize`` (not in
        # the reference's config) applies l2_normalize on load for raw network outputs
        gpu = _on_gpu(self.embeddings)
        if isinstance(self.embeddings, _capi.DLPackTensor) and not gpu:
            raise TypeError('host embeddings must be handed over as an array; the capsule path is for GPU tensors')
        embeddings = self.embeddings if gpu else np.asarray(self.embeddings)
        labels = _host_array(self.labels)
        normalize = bool(getattr(self.config, 'normalize', False))

        def subset(rows):
            if gpu or normalize:
                return SimilarityCalculator(embeddings, labels[rows], metric=self.config.metric, _rows=rows, normalize=normalize)
            return SimilarityCalculator(embeddings[rows], labels[rows], metric=self.config.metric)

        self.reports = (
            Report(criterion='MaximumAccuracy'),
            Report(criterion='FalseAlarmRate(FAR = {})'.format(self.config.far_target))
        )
        for train_set, test_set in kfold_split(len(labels), self.config.nrof_folds):
            # evaluations with train set and define the best threshold for the fold
            calculator = subset(train_set)
            matrix = ConfidenceMatrix(calculator, self.thresholds, _far_target=self.config.far_target)
            for report in self.reports:
                report.append_fold('train', matrix)

            if matrix._argmax_accuracy is not None:
                # selected on the device (fnb_confidence_from_last_bins): first accuracy maximum (statistics.py:296)
                # and the FAR threshold by linear interpolation over fp_rates (statistics.py:299-302)
                accuracy_threshold = self.thresholds[matrix._argmax_accuracy]
                far_threshold = matrix._far_threshold
                if far_threshold != far_threshold:
                    raise ValueError('A value in x_new is below the interpolation range.')
                if far_threshold != 0:
                    # The device interpolates on the bracketing samples in threshold order.  The reference sorts the samples
                    # with an unstable argsort first (scipy 1.4.1 / numpy 1.19): when the target sits right behind a RUN of
                    # equal fp_rates, the left sample is whichever tied one that sort puts last -- replayed on the host.
                    fpr = matrix.fp_rates
                    below = fpr[fpr <= self.config.far_target]
                    if below.size and np.count_nonzero(fpr == below.max()) > 1:
                        far_threshold = _slinear(fpr, self.thresholds, self.config.far_target)
                    else:
                        far_threshold = np.array(far_threshold)
                else:
                    far_threshold = 0
            else:
                # degenerate fold (fewer than two embeddings): nothing was launched
                accuracy_threshold = self.thresholds[np.argmax(matrix.accuracy)]
                far_threshold = 0
                if np.max(matrix.fp_rates) >= self.config.far_target:
                    far_threshold = _slinear(matrix.fp_rates, self.thresholds, self.config.far_target)

            # evaluations with test set: both thresholds in ONE launch, then split per report
            calculator = subset(test_set)
            both = ConfidenceMatrix(calculator, np.array([accuracy_threshold, float(far_threshold)]))
            for idx, thr in enumerate((accuracy_threshold, far_threshold)):
                one = ConfidenceMatrix.__new__(ConfidenceMatrix)
                one.threshold = np.array(thr, ndmin=1)
                one.tp, one.tn = both.tp[idx:idx + 1].copy(), both.tn[idx:idx + 1].copy()
                one.fp, one.fn = both.fp[idx:idx + 1].copy(), both.fn[idx:idx + 1].copy()
                one.stats = both.stats
                self.reports[idx].append_fold('test', one)

        self.elapsed_time = time.monotonic() - self.elapsed_time

    @property
    def dict(self):
        output = {r.criterion: r.dict for r in self.reports}
        return output

    def write_report(self, file):
        file = Path(file).expanduser()
        with file.open('at') as f:
            f.write(64 * '-' + '\n')
            f.write('{} {}\n'.format(self.__class__.__name__, datetime.datetime.now()))
            f.write('metric: {}\n\n'.format(self.config.metric))
            for r in self.reports:
                f.write(str(r))

    def write_h5file(self, h5file, tag=None):
        # statistics.py:330-331 -> h5utils.write_dict (h5utils.py:9-26): resizable gzip datasets, appended per call
        from facenet_b200 import h5utils
        h5utils.write_dict(h5file, self.dict, group=tag)


class FalseExamples:
    """The hardest false pairs at a threshold -- the class the reference keeps commented out (statistics.py:334-421).

    Reference algorithm (``write_false_pairs``, :341-387), per class ("folder") in ``np.unique(labels)`` order:
      * within the class: up to ``nrof_fpos_images`` times, take the pair with the LARGEST distance; if it exceeds the threshold
        it is a missed match: record it and retire both images (rows and columns of the class's matrix), else stop (:355-368);
      * against every later class: up to ``nrof_fneg_images`` times, take the pair with the SMALLEST distance; if it is below
        the threshold it is a false accept: record it and retire that row and that column, else stop (:371-386).
    (The reference's directory names are crossed: missed matches go to ``fneg_dir``, false accepts to ``fpos_dir``; kept.)

    Here the candidates come from ONE launch -- a filter epilogue of the Gram kernel appends every same-identity pair above and
    every different-identity pair below the threshold to a compact list (``fnb_false_pairs``) -- and the greedy selection runs
    on that short list.  ``embeddings [N, D]``, ``labels [N]``; ``files`` (optional, [N]) are the image paths the reference's
    ``dbase`` supplies.  ``subtract_mean`` subtracts the mean embedding first, like :346-349 (distances are then those of
    un-normalised vectors, exactly as the reference would compute them)."""

    def __init__(self, embeddings, labels, threshold, metric=0, subtract_mean=False, files=None):
        self.embeddings = np.ascontiguousarray(_host_array(embeddings), dtype=np.float32)
        self.labels = _host_array(labels)
        if self.embeddings.shape[0] != len(self.labels):
            raise ValueError('embeddings and labels have different lengths')
        if metric not in (0, 1):
            raise ValueError('Undefined similarity metric {}'.format(metric))
        self.threshold = threshold
        self.metric = metric
        self.subtract_mean = subtract_mean
        self.files = files
        self.stats = None

    def false_pairs(self, nrof_fpos_images=10, nrof_fneg_images=2):
        """``{'fneg': [(distance, a, b), ...], 'fpos': [...]}``: missed same-identity pairs and false accepts in the reference's
        visiting order (class by class; within a class the greedy order), ``a``, ``b`` = row indices into ``embeddings``."""
        x = self.embeddings
        if self.subtract_mean:
            x = x - np.mean(x, axis=0)
        x = _pad64(np.ascontiguousarray(x, dtype=np.float32))
        _, cls = np.unique(self.labels, return_inverse=True)
        cls = np.asarray(cls).reshape(-1)
        try:
            rows, cols, dist, self.stats = _handle().false_pairs(x, self.labels, float(self.threshold), metric=self.metric)
        except _capi.FnbError as err:
            _raise_like_reference(err, self.metric)
        # position of every row inside its class (files1[i] of dbase.extract_data: original order within the class)
        order = np.argsort(cls, kind='stable')
        start = np.concatenate([[0], np.cumsum(np.bincount(cls))])
        local = np.empty(cls.size, dtype=np.int64)
        local[order] = np.arange(cls.size) - start[cls[order]]
        # orient every pair: a belongs to the earlier class (within a class: the earlier image)
        swap = (cls[cols] < cls[rows]) | ((cls[cols] == cls[rows]) & (local[cols] < local[rows]))
        a = np.where(swap, cols, rows).astype(np.int64)
        b = np.where(swap, rows, cols).astype(np.int64)
        same = cls[a] == cls[b]
        out = {'fneg': [], 'fpos': []}
        # missed matches: np.argmax over the class's symmetric matrix = largest distance, first in row-major order
        idx = np.nonzero(same)[0]
        key = np.lexsort((local[b[idx]], local[a[idx]], -dist[idx].astype(np.float64), cls[a[idx]]))
        self._greedy(idx[key], cls[a], None, a, b, dist, nrof_fpos_images, out['fneg'], both=True)
        # false accepts: np.argmin over the [n_1, n_2] block = smallest distance, first in row-major order
        idx = np.nonzero(~same)[0]
        key = np.lexsort((local[b[idx]], local[a[idx]], dist[idx].astype(np.float64), cls[b[idx]], cls[a[idx]]))
        self._greedy(idx[key], cls[a], cls[b], a, b, dist, nrof_fneg_images, out['fpos'], both=False)
        return out

    @staticmethod
    def _greedy(sorted_idx, g1, g2, a, b, dist, limit, sink, both):
        """walk the candidates group by group (already sorted by group, then by the reference's pick order); a pick retires
        its row and its column (``both``: either image in either role)"""
        prev, used_a, used_b, taken = None, set(), set(), 0
        for i in sorted_idx:
            group = (g1[i],) if g2 is None else (g1[i], g2[i])
            if group != prev:
                prev, used_a, used_b, taken = group, set(), set(), 0
            if taken >= limit:
                continue
            ia, ib = int(a[i]), int(b[i])
            if both:
                if ia in used_a or ib in used_a:
                    continue
                used_a.update((ia, ib))
            else:
                if ia in used_a or ib in used_b:
                    continue
                used_a.add(ia)
                used_b.add(ib)
            sink.append((float(dist[i]), ia, ib))
            taken += 1

    def generate_filename(self, dirname, distance, file1, file2):
        # statistics.py:389-396
        import os
        dir1 = os.path.basename(os.path.dirname(file1))
        name1 = os.path.splitext(os.path.basename(file1))[0]
        dir2 = os.path.basename(os.path.dirname(file2))
        name2 = os.path.splitext(os.path.basename(file2))[0]
        return os.path.join(dirname, '{:2.3f} & {}|{} & {}|{}.png'.format(distance, dir1, name1, dir2, name2))

    def generate_text(self, distance, file1, file2):
        # statistics.py:398-403
        import os

        def text(file):
            return os.path.join(os.path.basename(os.path.dirname(file)), os.path.splitext(os.path.basename(file))[0])

        return '{} & {}\n{:2.3f}/{:2.3f}'.format(text(file1), text(file2), distance, self.threshold)

    def write_false_pairs(self, fpos_dir, fneg_dir, nrof_fpos_images=10, nrof_fneg_images=2):
        """The reference renders each pair side by side into a PNG (PIL, :405-421); here every directory receives
        ``false_pairs.txt`` with one line per pair -- the file name the reference would have written and its caption -- and the
        montage itself when PIL and the image files are available."""
        pairs = self.false_pairs(nrof_fpos_images, nrof_fneg_images)
        files = self.files if self.files is not None else ['{}/{}'.format(l, i) for i, l in enumerate(self.labels)]
        for dirname, key in ((Path(fneg_dir).expanduser(), 'fneg'), (Path(fpos_dir).expanduser(), 'fpos')):
            dirname.mkdir(parents=True, exist_ok=True)
            with (dirname / 'false_pairs.txt').open('wt') as f:
                for distance, ia, ib in pairs[key]:
                    name = self.generate_filename(str(dirname), distance, str(files[ia]), str(files[ib]))
                    f.write('{}\t{}\n'.format(name, self.generate_text(distance, str(files[ia]), str(files[ib])).replace('\n', ' ')))
                    self._write_image(name, str(files[ia]), str(files[ib]))
        return pairs

    @staticmethod
    def _write_image(fname, file1, file2):
        try:
            from PIL import Image
            if not (Path(file1).is_file() and Path(file2).is_file()):
                return
            img = Image.fromarray(np.concatenate([np.asarray(Image.open(file1)), np.asarray(Image.open(file2))], axis=1))
            img.save(fname)
        except Exception:
            pass


def pair_histogram(embeddings, labels, thresholds, metric=0, mode='auto', **kw):
    """Whole-set verification histogram (BASELINE configs 2/4/5): integer numbers of same-identity and
    different-identity pairs with ``d < thresholds[n]`` over all N(N-1)/2 unordered pairs -- the
    reference's inner statement ``count_nonzero(sims < threshold)`` (statistics.py:131) without the
    per-class weighting.  Returns a dict with ``same``, ``diff`` (int64 [T]), ``n_same``, ``n_diff``, ``stats``.

    ``mode='auto'`` (default here; the drop-in classes above default to the strict ``'fp16x3'``) lets the library use the
    faster ``fp16f8`` contraction when the embeddings are dense enough for its error model and ``fp16x3`` otherwise;
    ``stats['mode_used']`` reports the choice."""
    try:
        return _handle().pair_histogram(embeddings, labels, thresholds, metric, mode=mode or _state['mode'],
                                        cta_group=kw.pop('cta_group', _state['cta_group']), **kw)
    except _capi.FnbError as err:
        _raise_like_reference(err, metric)
