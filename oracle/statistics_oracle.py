"""NumPy restatement of the reference's evaluation-statistics path.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``): only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this.

Every function cites the lines of ``/root/reference/facenet/statistics.py`` it
follows.  Two forms of the same arithmetic are provided:

* the LITERAL form (``SimilarityCalculator`` / ``ConfidenceMatrix`` /
  ``FaceToFaceValidation``): the reference's per-class-pair, per-threshold loops,
  restated;
* the VECTORISED form (``pair_histogram``, ``confidence_matrix_exact_order``,
  ``confidence_matrix_weighted``, ``face_to_face_validation``): blocked fp32
  ``X @ X.T`` + ``searchsorted`` binning, used for sizes the literal loops cannot
  finish and as the fair CPU throughput baseline.

Pinned against outputs of the unmodified reference run in the build container
(``oracle/gen_golden.py`` -> ``tests/golden/*.npz``; ``tests/test_oracle.py``).
"""
import numpy as np

# --------------------------------------------------------------------------------------
# helpers


def thresholds_f32_up(thresholds):
    """float32 thresholds rounded UP so that, for every float32 ``d``,
    ``d < t32_up  <=>  float64(d) < t`` -- the comparison NumPy >= 2 performs for
    ``sims < threshold`` with float32 ``sims`` and a float64 scalar
    (statistics.py:131)."""
    t64 = np.atleast_1d(np.asarray(thresholds, dtype=np.float64))
    t32 = t64.astype(np.float32)
    low = t32.astype(np.float64) < t64
    t32[low] = np.nextafter(t32[low], np.float32(np.inf))
    return t32


def upper_threshold(metric):
    """statistics.py:255-260."""
    if metric == 0:
        return 4
    if metric == 1:
        return np.pi
    raise ValueError('Undefined similarity metric {}'.format(metric))


def default_thresholds(metric):
    """statistics.py:262."""
    return np.linspace(0, upper_threshold(metric), 100)


def similarity_to_distance(sims, metric, atol=1.e-5):
    """Range check, clamp and metric transform applied to raw Gram values
    (statistics.py:38-55).  ``sims`` is modified in place like the reference does."""
    if sims.size > 0:
        lim = 1 + atol
        if sims.min() < -lim or sims.max() > lim:
            raise ValueError('\nembeddings must be normalized to 1, range {} {}'.format(sims.min(), sims.max()))
        sims[sims < -1] = -1
        sims[sims > +1] = +1
        if metric == 0:
            sims = 2 * (1 - sims)
        elif metric == 1:
            sims = np.arccos(sims)
        else:
            raise ValueError('Undefined similarity metric {}'.format(metric))
    return sims


# --------------------------------------------------------------------------------------
# literal form


def pairwise_similarities(xa, xb=None, metric=0, atol=1.e-5):
    """statistics.py:22-57.  Self form: strict upper triangle of ``xa @ xa.T`` in
    row-major ``triu_indices(n, 1)`` order (1-D); cross form: ``xa @ xb.T`` (2-D)."""
    if xb is None:
        gram = xa @ xa.transpose()
        sims = gram[np.triu_indices(gram.shape[0], k=1)]
    else:
        sims = xa @ xb.transpose()
    return similarity_to_distance(sims, metric, atol)


def split_embeddings(embeddings, labels):
    """statistics.py:68-79: per-class row blocks in sorted ``np.unique(labels)`` order."""
    return [embeddings[labels == value] for value in np.unique(labels)]


class SimilarityCalculator:
    """statistics.py:82-108."""

    def __init__(self, embeddings, labels, metric=0):
        self.metric = metric
        self.embeddings = split_embeddings(embeddings, labels)

    @property
    def nrof_classes(self):
        return len(self.embeddings)

    def nrof_images(self, i):
        return self.embeddings[i].shape[0]

    def evaluate(self, i, k):
        # statistics.py:90-101 -- weights: C for same-class blocks, C(C-1)/2 (float) otherwise
        if i == k:
            sims = pairwise_similarities(self.embeddings[i], metric=self.metric)
            weight = sims.size * self.nrof_classes
        else:
            sims = pairwise_similarities(self.embeddings[i], self.embeddings[k], metric=self.metric)
            weight = sims.size * (self.nrof_classes * (self.nrof_classes - 1) / 2)
        return sims, weight


class _Rates:
    """Derived quantities of a confidence matrix (statistics.py:140-175)."""

    @property
    def accuracy(self):
        return (self.tp + self.tn) / (self.tp + self.fp + self.tn + self.fn)

    def _ratio(self, num, other):
        mask = (num + other) > 0
        out = np.ones(self.threshold.size)
        out[mask] = num[mask] / (num[mask] + other[mask])
        return out

    @property
    def precision(self):
        return self._ratio(self.tp, self.fp)

    @property
    def tp_rates(self):
        return self._ratio(self.tp, self.fn)

    @property
    def tn_rates(self):
        return self._ratio(self.tn, self.fp)

    @property
    def fp_rates(self):
        return 1 - self.tn_rates

    @property
    def fn_rates(self):
        return 1 - self.tp_rates


class ConfidenceMatrix(_Rates):
    """statistics.py:111-138: literal loops (i ascending, k = 0..i, thresholds inner)."""

    def __init__(self, calculator, threshold):
        self.threshold = np.array(threshold, ndmin=1)
        nt = self.threshold.size
        self.tp, self.tn, self.fp, self.fn = (np.zeros(nt) for _ in range(4))
        for i in range(calculator.nrof_classes):
            for k in range(i + 1):
                sims, weight = calculator.evaluate(i, k)
                if sims.size < 1:
                    continue
                for n, t in enumerate(self.threshold):
                    count = np.count_nonzero(sims < t)          # strict <, float64 compare
                    if i == k:
                        self.tp[n] += count / weight
                        self.fn[n] += (sims.size - count) / weight
                    else:
                        self.fp[n] += count / weight
                        self.tn[n] += (sims.size - count) / weight


class RatesFromArrays(_Rates):
    """Confidence matrix built from precomputed tp/tn/fp/fn arrays."""

    def __init__(self, threshold, tp, tn, fp, fn):
        self.threshold = np.array(threshold, ndmin=1)
        self.tp, self.tn, self.fp, self.fn = tp, tn, fp, fn


# --------------------------------------------------------------------------------------
# vectorised form


def _sorted_classes(labels):
    """class rank per row (rank of the label value, statistics.py:76) and sizes."""
    values, cls, sizes = np.unique(labels, return_inverse=True, return_counts=True)
    return values, cls.astype(np.int64), sizes.astype(np.int64)


def _blocked_lower_bins(x, thr32_up, metric, atol, block):
    """Yield ``(r0, r1, c0, c1, bins, valid)`` sub-blocks covering the strict lower
    triangle of ``x @ x.T``: ``bins[i, j] = #{n : thr_n <= d(r0+i, c0+j)}`` (so
    ``d < thr_n  <=>  bins <= n``); ``valid`` is None (whole block below the diagonal)
    or a boolean mask (diagonal block).  Also tracks the global min/max raw similarity
    for the normalisation check (statistics.py:40-42)."""
    n = x.shape[0]
    state = {'min': np.inf, 'max': -np.inf}
    lim = 1 + atol

    def transform(s, valid):
        vals = s if valid is None else s[valid]
        if vals.size:
            lo, hi = float(vals.min()), float(vals.max())
            state['min'] = min(state['min'], lo)
            state['max'] = max(state['max'], hi)
            if lo < -lim or hi > lim:
                raise ValueError('\nembeddings must be normalized to 1, range {} {}'.format(vals.min(), vals.max()))
        np.clip(s, -1, 1, out=s)                                       # statistics.py:45-46
        d = 2 * (1 - s) if metric == 0 else np.arccos(s)               # statistics.py:50,53
        return np.searchsorted(thr32_up, d.ravel(), side='right').reshape(d.shape)

    def gen():
        for r0 in range(0, n, block):
            r1 = min(n, r0 + block)
            s = x[r0:r1] @ x[:r1].T                                    # fp32 sgemm (statistics.py:33,36)
            if r0 > 0:
                yield r0, r1, 0, r0, transform(s[:, :r0], None), None
            tri = np.tri(r1 - r0, r1 - r0, -1, dtype=bool)
            yield r0, r1, r0, r1, transform(s[:, r0:r1], tri), tri
    return gen(), state


def _row_block_histogram(x, labels, thr32_up, metric, atol, r0, r1):
    """Histogram contribution of rows [r0, r1) against all earlier rows (strict lower triangle)."""
    nt = thr32_up.size
    lim = 1 + atol
    s = x[r0:r1] @ x[:r1].T                                            # fp32 sgemm (statistics.py:33,36)
    valid = np.ones(s.shape, dtype=bool)
    valid[:, r0:r1] = np.tri(r1 - r0, r1 - r0, -1, dtype=bool)
    vals = s[valid]
    lo, hi = (float(vals.min()), float(vals.max())) if vals.size else (np.inf, -np.inf)
    if vals.size and (lo < -lim or hi > lim):
        raise ValueError('\nembeddings must be normalized to 1, range {} {}'.format(vals.min(), vals.max()))
    np.clip(vals, -1, 1, out=vals)                                     # statistics.py:45-46
    d = 2 * (1 - vals) if metric == 0 else np.arccos(vals)             # statistics.py:50,53
    bins = np.searchsorted(thr32_up, d, side='right')
    same = (labels[r0:r1, None] == labels[None, :r1])[valid]
    return (np.bincount(bins, minlength=nt + 1), np.bincount(bins[same], minlength=nt + 1), lo, hi)


def pair_histogram(embeddings, labels, thresholds, metric=0, atol=1.e-5, block=1024, threads=1):
    """Whole-set verification histogram (SURVEY.md section 8 a, end): for every
    unordered pair {a,b}, a != b, evaluated once, the integer number of
    same-identity and different-identity pairs with ``d < thresholds[n]``
    (strict, statistics.py:131).  Returns dict with ``same``/``diff`` int64 [T]
    cumulative counts, ``n_same``/``n_diff`` totals and the raw similarity range.
    ``threads`` > 1 processes row blocks on a thread pool (NumPy releases the GIL in
    matmul / searchsorted / bincount), BLAS limited to one thread per worker."""
    if metric not in (0, 1):
        raise ValueError('Undefined similarity metric {}'.format(metric))
    x = np.ascontiguousarray(embeddings, dtype=np.float32)
    labels = np.asarray(labels)
    thr = thresholds_f32_up(thresholds)
    nt = thr.size
    blocks = [(r0, min(x.shape[0], r0 + block)) for r0 in range(0, x.shape[0], block)]

    def work(rr):
        return _row_block_histogram(x, labels, thr, metric, atol, rr[0], rr[1])

    if threads > 1 and len(blocks) > 1:
        from concurrent.futures import ThreadPoolExecutor
        from threadpoolctl import threadpool_limits
        with threadpool_limits(limits=1, user_api='blas'), ThreadPoolExecutor(max_workers=threads) as pool:
            parts = list(pool.map(work, blocks))
    else:
        parts = [work(rr) for rr in blocks]
    all_h = np.zeros(nt + 1, dtype=np.int64)
    same_h = np.zeros(nt + 1, dtype=np.int64)
    smin, smax = np.inf, -np.inf
    for a_h, s_h, lo, hi in parts:
        all_h += a_h
        same_h += s_h
        smin, smax = min(smin, lo), max(smax, hi)
    diff_h = all_h - same_h
    return {'same': np.cumsum(same_h)[:nt], 'diff': np.cumsum(diff_h)[:nt],
            'n_same': int(same_h.sum()), 'n_diff': int(diff_h.sum()), 'smin': smin, 'smax': smax}


def eps_window_pairs(embeddings, thresholds, metric=0, eps=1.e-5, block=1024):
    """Number of unordered pairs whose distance lies within ``eps`` of some threshold."""
    x = np.ascontiguousarray(embeddings, dtype=np.float32)
    t = np.sort(np.atleast_1d(np.asarray(thresholds, dtype=np.float64)))
    n = x.shape[0]
    total = 0
    for r0 in range(0, n, block):
        r1 = min(n, r0 + block)
        s = x[r0:r1] @ x[:r1].T
        mask = np.arange(r0, r1)[:, None] > np.arange(r1)[None, :]
        v = np.clip(s[mask], -1, 1)
        d = (2 * (1 - v) if metric == 0 else np.arccos(v)).astype(np.float64)
        j = np.clip(np.searchsorted(t, d), 1, t.size - 1)
        near = np.minimum(np.abs(d - t[j - 1]), np.abs(d - t[j]))
        total += int(np.count_nonzero(near <= eps))
    return total


def class_pair_counts(embeddings, labels, thresholds, metric=0, atol=1.e-5, block=1024):
    """Integer ``count_nonzero(sims < t_n)`` per (class i, class k <= i, threshold n)
    (statistics.py:131), as a dense int64 array [C, C, T] (lower triangle filled),
    plus class sizes.  For small C only (memory C*C*T)."""
    x = np.ascontiguousarray(embeddings, dtype=np.float32)
    _, cls, sizes = _sorted_classes(labels)
    order = np.argsort(cls, kind='stable')
    xs, cs = x[order], cls[order]
    thr = thresholds_f32_up(thresholds)
    nt, nc = thr.size, sizes.size
    hist = np.zeros(nc * nc * (nt + 1), dtype=np.int64)
    it, _ = _blocked_lower_bins(xs, thr, metric, atol, block)
    for r0, r1, c0, c1, bins, valid in it:
        key = (cs[r0:r1, None] * nc + cs[None, c0:c1]) * (nt + 1) + bins
        key = key.ravel() if valid is None else key[valid]
        hist += np.bincount(key, minlength=hist.size)
    hist = hist.reshape(nc, nc, nt + 1)
    return np.cumsum(hist, axis=2)[:, :, :nt], sizes


def confidence_matrix_exact_order(embeddings, labels, thresholds, metric=0):
    """tp/tn/fp/fn of statistics.py:115-138 from per-class-pair integer counts,
    replaying the reference's fp64 operation order (i ascending, k = 0..i; one
    correctly-rounded division per term; sequential accumulation) -- bit-identical
    to the literal loops whenever the integer counts are identical."""
    counts, sizes = class_pair_counts(embeddings, labels, thresholds, metric)
    nc = sizes.size
    nt = counts.shape[2]
    ii, kk = np.tril_indices(nc)                      # lexicographic (i asc, k asc) == loop order
    npairs = np.where(ii == kk, sizes[ii] * (sizes[ii] - 1) // 2, sizes[ii] * sizes[kk])
    keep = npairs > 0                                 # statistics.py:127-128
    ii, kk, npairs = ii[keep], kk[keep], npairs[keep]
    cnt = counts[ii, kk, :].astype(np.float64)
    diag = ii == kk
    w_same = (npairs[diag] * nc).astype(np.float64)[:, None]                 # int weight
    w_diff = (npairs[~diag] * (nc * (nc - 1) / 2))[:, None]                  # float weight
    size_same = npairs[diag].astype(np.float64)[:, None]
    size_diff = npairs[~diag].astype(np.float64)[:, None]

    def seq_sum(terms):
        return np.cumsum(terms, axis=0)[-1] if terms.shape[0] else np.zeros(nt)

    tp = seq_sum(cnt[diag] / w_same)
    fn = seq_sum((size_same - cnt[diag]) / w_same)
    fp = seq_sum(cnt[~diag] / w_diff)
    tn = seq_sum((size_diff - cnt[~diag]) / w_diff)
    return RatesFromArrays(thresholds, tp, tn, fp, fn)


def confidence_matrix_weighted(embeddings, labels, thresholds, metric=0, atol=1.e-5, block=1024):
    """Same quantities for any number of classes using the separable pair weights
    ``1/(C * n_i(n_i-1)/2)`` (same class) and ``2/(C(C-1) n_i n_k)`` (different);
    agrees with the literal loops to ~1e-14 (summation order differs)."""
    x = np.ascontiguousarray(embeddings, dtype=np.float32)
    _, cls, sizes = _sorted_classes(labels)
    thr = thresholds_f32_up(thresholds)
    nt, nc = thr.size, sizes.size
    same_w = np.zeros(nt + 1)
    diff_w = np.zeros(nt + 1)
    npairs_same = sizes * (sizes - 1) / 2
    w_same_cls = np.where(npairs_same > 0, 1.0 / np.maximum(npairs_same * nc, 1), 0.0)
    inv_n = 1.0 / sizes
    c_diff = 1.0 / (nc * (nc - 1) / 2) if nc > 1 else 0.0
    it, _ = _blocked_lower_bins(x, thr, metric, atol, block)
    for r0, r1, c0, c1, bins, valid in it:
        cr = np.broadcast_to(cls[r0:r1, None], bins.shape)
        cc = np.broadcast_to(cls[None, c0:c1], bins.shape)
        same = cr == cc
        diff = ~same
        if valid is not None:
            same &= valid
            diff &= valid
        same_w += np.bincount(bins[same], weights=w_same_cls[cr[same]], minlength=nt + 1)
        diff_w += np.bincount(bins[diff], weights=inv_n[cr[diff]] * inv_n[cc[diff]] * c_diff, minlength=nt + 1)
    tp = np.cumsum(same_w)[:nt]
    fp = np.cumsum(diff_w)[:nt]
    fn = same_w.sum() - tp
    tn = diff_w.sum() - fp
    return RatesFromArrays(thresholds, tp, tn, fp, fn)


# --------------------------------------------------------------------------------------
# k-fold validation


def kfold_split(n, n_splits, seed=0):
    """``KFold(n_splits, shuffle=True, random_state=seed).split(arange(n))``
    (statistics.py:278-287): a ``RandomState(seed)`` shuffle of ``arange(n)`` cut into
    chunks, the first ``n % n_splits`` one longer; train/test index sets ascending."""
    perm = np.arange(n)
    np.random.RandomState(seed).shuffle(perm)
    sizes = np.full(n_splits, n // n_splits, dtype=np.int64)
    sizes[:n % n_splits] += 1
    start = 0
    for size in sizes:
        mask = np.zeros(n, dtype=bool)
        mask[perm[start:start + size]] = True
        yield np.nonzero(~mask)[0], np.nonzero(mask)[0]
        start += size


def argsort_numpy_scalar(v):
    """``np.argsort(v)`` as NumPy's scalar introsort orders it (``aquicksort``, numpy/core/src/npysort/quicksort: median-of-3
    partition down to 16 elements, then insertion sort) -- NOT stable: runs of equal keys come out permuted.  The reference's
    pinned numpy 1.19.4 (requirements.txt:9) has only this path; numpy >= 1.25 dispatches to SIMD sorts on AVX2 / AVX-512
    machines, which order ties differently.  Checked against this container's numpy with the SIMD dispatch disabled
    (tests/test_oracle.py::test_argsort_restatement_matches_numpy_scalar_path)."""
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


def slinear_interp(x, y, xq):
    """``scipy.interpolate.interp1d(x, y, kind='slinear')(xq)`` as the reference's pinned scipy 1.4.1 evaluates it
    (statistics.py:300-302; scipy >= 1.10 rejects the duplicate abscissae ``fp_rates`` always contains):
      1. ``interp1d.__init__``: ``ind = np.argsort(x)`` (default kind: the unstable scalar introsort), ``x, y = x[ind], y[ind]``;
      2. ``make_interp_spline(x, y, k=1)``: knots ``t = r_[x[0], x, x[-1]]``, coefficients ``y`` (no sortedness check on this path);
      3. ``BSpline.__call__``: interval = the LAST knot span whose left end is <= xq, then de Boor for k = 1:
         ``w = 1 / (xb - xa); y[j] * ((xb - xq) * w) + y[j + 1] * ((xq - xa) * w)``.
    Where ``xq`` falls behind a run of tied abscissae the left end is whichever tied sample the unstable sort put last."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    xq = float(xq)
    ind = argsort_numpy_scalar(x)
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
        return np.float64(ys[j] * ((xb - xq) * w) + ys[j + 1] * ((xq - xa) * w))


def report_dict(train, test):
    """statistics.py:210-234 (``Report.dict``)."""
    import sklearn.metrics
    from scipy import interpolate
    from scipy.optimize import brentq
    tp_rates = np.mean(np.array([m.tp_rates for m in train]), axis=0)
    tn_rates = np.mean(np.array([m.tn_rates for m in train]), axis=0)
    dct = {'auc': -1, 'eer': -1}
    try:
        dct['auc'] = sklearn.metrics.auc(1 - tn_rates, tp_rates)
    except Exception:
        pass
    try:
        dct['eer'] = brentq(lambda v: 1. - v - interpolate.interp1d(1 - tn_rates, tp_rates)(v), 0., 1.)
    except Exception:
        pass
    for key in ('accuracy', 'precision', 'tp_rates', 'tn_rates', 'threshold'):
        values = [getattr(m, key) for m in test]
        dct[key] = np.mean(values)
        dct[key + '_std'] = np.std(values)
    return dct


def face_to_face_validation(embeddings, labels, metric=0, nrof_folds=10, far_target=1.e-3,
                            confidence=confidence_matrix_weighted):
    """statistics.py:241-313 with a vectorised confidence matrix.  Returns
    ``{'MaximumAccuracy': {...}, 'FalseAlarmRate(FAR = x)': {...}}`` plus the chosen
    per-fold thresholds under key ``'_thresholds'``."""
    embeddings = np.asarray(embeddings)
    labels = np.asarray(labels)
    assert embeddings.shape[0] == len(labels)
    thresholds = default_thresholds(metric)
    train_cms, test_acc, test_far, chosen = [], [], [], []
    for train_set, test_set in kfold_split(len(labels), nrof_folds):
        cm = confidence(embeddings[train_set], labels[train_set], thresholds, metric)
        train_cms.append(cm)
        acc_thr = thresholds[np.argmax(cm.accuracy)]                     # statistics.py:296
        far_thr = 0                                                      # statistics.py:299-302
        if np.max(cm.fp_rates) >= far_target:
            far_thr = slinear_interp(cm.fp_rates, thresholds, far_target)
        chosen.append((float(acc_thr), float(far_thr)))
        test_acc.append(confidence(embeddings[test_set], labels[test_set], acc_thr, metric))
        test_far.append(confidence(embeddings[test_set], labels[test_set], far_thr, metric))
    return {'MaximumAccuracy': report_dict(train_cms, test_acc),
            'FalseAlarmRate(FAR = {})'.format(far_target): report_dict(train_cms, test_far),
            '_thresholds': np.array(chosen)}


# --------------------------------------------------------------------------------------
# synthetic inputs (SURVEY.md section 8 d)


def false_examples(embeddings, labels, threshold, metric=0, subtract_mean=False, nrof_fpos_images=10, nrof_fneg_images=2):
    """Literal NumPy statement of the search in the reference's COMMENTED-OUT ``FalseExamples.write_false_pairs``
    (facenet/statistics.py:341-387; PARITY UNPINNED: the code is commented out upstream and needs its dbase / tfrecord objects,
    so it cannot be executed -- the loops below follow it line by line): ``folder`` = class in ``np.unique(labels)`` order,
    images of a class in original order.  Returns ``{'fneg': [(distance, a, b)], 'fpos': [...]}`` with row indices."""
    x = np.ascontiguousarray(embeddings, dtype=np.float32)
    labels = np.asarray(labels)
    mean = np.mean(x, axis=0) if subtract_mean else 0                                  # :346-349
    values = np.unique(labels)
    members = [np.nonzero(labels == v)[0] for v in values]
    out = {'fneg': [], 'fpos': []}
    for f1 in range(len(values)):                                                      # :351
        rows1 = members[f1]
        e1 = (x[rows1] - mean).astype(np.float32)
        n1 = rows1.size
        sims = np.zeros((n1, n1), dtype=np.float32)                                    # squareform of the triangle, :358-359
        if n1 > 1:
            iu = np.triu_indices(n1, 1)
            tri = pairwise_similarities(e1.copy(), None, metric)
            sims[iu] = tri
            sims[(iu[1], iu[0])] = tri
        for _ in range(nrof_fpos_images):                                              # :361-368
            if sims.size == 0:
                break
            i, k = np.unravel_index(np.argmax(sims), sims.shape)
            if sims[i, k] > threshold:
                out['fneg'].append((float(sims[i, k]), int(rows1[i]), int(rows1[k])))
                sims[[i, k], :] = -1
                sims[:, [i, k]] = -1
            else:
                break
        for f2 in range(f1 + 1, len(values)):                                          # :371
            rows2 = members[f2]
            e2 = (x[rows2] - mean).astype(np.float32)
            cross = pairwise_similarities(e1.copy(), e2.copy(), metric).astype(np.float32)
            for _ in range(nrof_fneg_images):                                          # :376-386
                i, k = np.unravel_index(np.argmin(cross), cross.shape)
                if cross[i, k] < threshold:
                    out['fpos'].append((float(cross[i, k]), int(rows1[i]), int(rows2[k])))
                    cross[i, :] = np.inf
                    cross[:, k] = np.inf
                else:
                    break
    return out


def pair_histogram_window(embeddings, labels, thresholds, metric=0, eps=1.e-5, threads=1):
    """Per threshold, the exact number of same-identity / different-identity pairs whose ORACLE distance lies within ``eps``
    of it (``t - eps <= d <= t + eps``): the only pairs the contract (BASELINE.json north_star) allows an implementation
    to count on the other side of that threshold.  Returns ``(window_same, window_diff)`` int64 [T]."""
    thr = np.atleast_1d(np.asarray(thresholds, dtype=np.float64))
    hi = pair_histogram(embeddings, labels, np.nextafter(thr + eps, np.inf), metric, threads=threads)     # d <= t + eps
    lo = pair_histogram(embeddings, labels, thr - eps, metric, threads=threads)                            # d <  t - eps
    return hi['same'] - lo['same'], hi['diff'] - lo['diff']


def synthetic_embeddings(class_sizes, dim=512, sigma=1.1, seed=0, shuffle=True, label_values=None):
    """Clustered unit-norm fp32 embeddings: centre ~ N(0,I), sample = centre +
    sigma*N(0,I), L2-normalised in fp32.  Returns (embeddings [N,dim] f32, labels [N] i64).
    ``sigma`` may be a pair (lo, hi): every class draws its own sigma ~ U(lo, hi) -- tight and loose identities side by
    side, which makes the same / different distance distributions overlap (AUC < 1; (1.5, 3.5) gives AUC ~ 0.97 at 512-d)."""
    rng = np.random.default_rng(seed)
    class_sizes = np.asarray(class_sizes, dtype=np.int64)
    nc = class_sizes.size
    centres = rng.standard_normal((nc, dim), dtype=np.float32)
    cls = np.repeat(np.arange(nc), class_sizes)
    if np.ndim(sigma) == 0:
        x = centres[cls] + np.float32(sigma) * rng.standard_normal((cls.size, dim), dtype=np.float32)
    else:
        per_class = rng.uniform(float(sigma[0]), float(sigma[1]), size=nc).astype(np.float32)
        x = centres[cls] + per_class[cls][:, None] * rng.standard_normal((cls.size, dim), dtype=np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    values = np.arange(nc, dtype=np.int64) if label_values is None else np.asarray(label_values, dtype=np.int64)
    labels = values[cls]
    if shuffle:
        perm = rng.permutation(cls.size)
        x, labels = x[perm], labels[perm]
    return np.ascontiguousarray(x, dtype=np.float32), labels


def lfw_like_class_sizes(n_images=13233, n_ids=5749, n_single=4069, max_size=530, seed=0):
    """Deterministic ragged class-size vector with LFW's gross statistics
    (13,233 images, 5,749 identities, 4,069 singletons, largest 530)."""
    rng = np.random.default_rng(seed)
    n_multi = n_ids - n_single
    budget = n_images - n_single
    raw = 2 + np.floor(rng.pareto(1.15, size=n_multi) * 1.2).astype(np.int64)
    raw = np.minimum(raw, max_size)
    raw[np.argmax(raw)] = max_size
    # adjust to hit the image budget exactly while keeping every size >= 2
    while raw.sum() > budget:
        j = rng.integers(n_multi)
        if 2 < raw[j] < max_size:
            raw[j] -= 1
    while raw.sum() < budget:
        j = rng.integers(n_multi)
        if raw[j] < max_size - 1:
            raw[j] += 1
    sizes = np.concatenate([np.ones(n_single, dtype=np.int64), raw])
    rng.shuffle(sizes)
    assert sizes.sum() == n_images and sizes.size == n_ids
    return sizes
