"""The oracle restatement against (a) committed outputs of the unmodified reference
(tests/golden, made by oracle/gen_golden.py) and (b) the reference itself when
/root/reference is mounted (build container)."""
import numpy as np
import pytest

from oracle import statistics_oracle as so
from oracle import mining_oracle as mo
from pathlib import Path


def test_pairwise_golden(golden_dir):
    g = np.load(golden_dir / 'pairwise.npz')
    xa, xb = g['xa'], g['xb']
    for metric in (0, 1):
        got = so.pairwise_similarities(xa.copy(), metric=metric)
        assert got.dtype == np.float32 and got.shape == (37 * 36 // 2,)
        np.testing.assert_array_equal(got, g['self_m%d' % metric])
        got = so.pairwise_similarities(xa.copy(), xb.copy(), metric=metric)
        assert got.shape == (37, 21)
        np.testing.assert_array_equal(got, g['cross_m%d' % metric])


def test_pairwise_errors_and_empty():
    x = np.eye(4, 8, dtype=np.float32)
    with pytest.raises(ValueError, match='Undefined similarity metric 2'):
        so.pairwise_similarities(x, metric=2)
    with pytest.raises(ValueError, match='embeddings must be normalized to 1'):
        so.pairwise_similarities(2 * x, x)
    e = so.pairwise_similarities(x[:1])
    assert e.size == 0


def test_confidence_literal_and_vectorised_golden(golden_dir):
    g = np.load(golden_dir / 'confidence.npz')
    x, labels = g['embeddings'], g['labels']
    for metric in (0, 1):
        thr = so.default_thresholds(metric)
        lit = so.ConfidenceMatrix(so.SimilarityCalculator(x, labels, metric), thr)
        for name in ('tp', 'tn', 'fp', 'fn', 'accuracy', 'precision', 'tp_rates', 'tn_rates'):
            np.testing.assert_array_equal(getattr(lit, name), g['%s_m%d' % (name, metric)], err_msg=name)
        counts, sizes = so.class_pair_counts(x, labels, thr, metric)
        ref_counts = g['counts_m%d' % metric]
        # whole-Gram vs per-block sgemm may differ in the last ulp for a pair sitting on a threshold
        assert np.abs(counts - ref_counts).sum() <= 2
        vec = so.confidence_matrix_exact_order(x, labels, thr, metric)
        wgt = so.confidence_matrix_weighted(x, labels, thr, metric)
        for name in ('tp', 'tn', 'fp', 'fn'):
            ref = g['%s_m%d' % (name, metric)]
            if np.array_equal(counts, ref_counts):
                np.testing.assert_array_equal(getattr(vec, name), ref, err_msg=name)
            np.testing.assert_allclose(getattr(vec, name), ref, rtol=0, atol=1e-3)
            np.testing.assert_allclose(getattr(wgt, name), getattr(vec, name), rtol=0, atol=1e-13)
        one = so.confidence_matrix_exact_order(x, labels, np.array(thr[31] + 0.0123), metric)
        np.testing.assert_allclose([one.tp[0], one.tn[0], one.fp[0], one.fn[0]], g['single_m%d' % metric],
                                   rtol=0, atol=1e-3)


def test_pair_histogram_consistency():
    x, labels = so.synthetic_embeddings([5, 1, 9, 30, 2, 2, 17], dim=64, sigma=1.0, seed=2)
    thr = so.default_thresholds(0)
    h = so.pair_histogram(x, labels, thr, 0, block=16)
    n = x.shape[0]
    assert h['n_same'] + h['n_diff'] == n * (n - 1) // 2
    d = so.pairwise_similarities(x.copy())
    iu = np.triu_indices(n, 1)
    same = labels[iu[0]] == labels[iu[1]]
    for t_i in (0, 17, 50, 99):
        assert abs(h['same'][t_i] - np.count_nonzero(d[same] < thr[t_i])) <= 1
        assert abs(h['diff'][t_i] - np.count_nonzero(d[~same] < thr[t_i])) <= 1
    assert np.all(np.diff(h['same']) >= 0) and np.all(np.diff(h['diff']) >= 0)
    assert h['same'][0] == 0 and h['diff'][0] == 0          # nothing is < 0


def test_thresholds_f32_up_equivalence():
    rng = np.random.default_rng(0)
    thr = so.default_thresholds(0)
    t32 = so.thresholds_f32_up(thr)
    near = np.concatenate([np.nextafter(t32, np.float32(-np.inf)), t32, np.nextafter(t32, np.float32(np.inf)),
                           rng.uniform(0, 4, 1000).astype(np.float32)])
    for t, tu in zip(thr, t32):
        np.testing.assert_array_equal(near < t, near < tu)


def test_kfold_matches_sklearn():
    from sklearn.model_selection import KFold
    for n, k in ((330, 10), (103, 10), (50, 3)):
        ref = list(KFold(n_splits=k, shuffle=True, random_state=0).split(np.arange(n)))
        got = list(so.kfold_split(n, k))
        assert len(ref) == len(got)
        for (a, b), (c, d) in zip(ref, got):
            np.testing.assert_array_equal(a, c)
            np.testing.assert_array_equal(b, d)


def test_validation_golden(golden_dir):
    g = np.load(golden_dir / 'validation.npz')
    x, labels = g['embeddings'], g['labels']
    for metric in (0, 1):
        for conf in (so.confidence_matrix_exact_order, so.confidence_matrix_weighted):
            out = so.face_to_face_validation(x, labels, metric, 10, 1.e-3, confidence=conf)
            for tag, key in (('acc', 'MaximumAccuracy'), ('far', 'FalseAlarmRate(FAR = 0.001)')):
                keys = [str(k) for k in g['%s_keys_m%d' % (tag, metric)]]
                vals = g['%s_vals_m%d' % (tag, metric)]
                assert sorted(out[key].keys()) == keys
                got = np.array([float(out[key][k]) for k in keys])
                np.testing.assert_allclose(got, vals, rtol=0, atol=2e-3, err_msg='%s m%d' % (key, metric))
            np.testing.assert_array_equal(out['_thresholds'][:, 0], g['acc_thr_m%d' % metric])
            np.testing.assert_allclose(out['_thresholds'][:, 1], g['far_thr_m%d' % metric], rtol=0, atol=2e-3)


@pytest.mark.reference
def test_restatement_equals_live_reference():
    """Build container only: the unmodified reference on a fresh ragged input."""
    from oracle.reference_loader import load_reference_statistics
    st = load_reference_statistics()
    x, labels = so.synthetic_embeddings([3, 1, 6, 2, 11, 1, 4], dim=32, sigma=1.5, seed=9)
    for metric in (0, 1):
        thr = so.default_thresholds(metric)
        ref = st.ConfidenceMatrix(st.SimilarityCalculator(x, labels, metric=metric), thr)
        lit = so.ConfidenceMatrix(so.SimilarityCalculator(x, labels, metric), thr)
        for name in ('tp', 'tn', 'fp', 'fn'):
            np.testing.assert_array_equal(getattr(lit, name), getattr(ref, name))
        np.testing.assert_array_equal(st.pairwise_similarities(x.copy(), metric=metric),
                                      so.pairwise_similarities(x.copy(), metric=metric))


def test_mining_oracle_small():
    x, labels = so.synthetic_embeddings([4, 4, 4], dim=32, sigma=1.0, seed=1, shuffle=False)
    out = mo.mine(x, labels, alpha=0.2)
    d = mo.distance_matrix(x)
    for a in range(12):
        p = out['hardest_pos'][a]
        n = out['hardest_neg'][a]
        assert labels[p] == labels[a] and p != a and labels[n] != labels[a]
        assert d[a, p] == max(d[a, q] for q in range(12) if labels[q] == labels[a] and q != a)
        assert d[a, n] == min(d[a, q] for q in range(12) if labels[q] != labels[a])
        for j in range(3):
            p = out['pos_index'][a, j]
            s = out['semi_hard'][a, j]
            if s >= 0:
                assert labels[s] != labels[a] and d[a, s] > d[a, p] and np.float32(d[a, s] - d[a, p]) < np.float32(0.2)
    assert out['pos_index'].shape == (12, 3)


def test_mining_oracle_kth_eligible_and_upstream_selection():
    """select_kth_eligible walks the ascending candidate list; the upstream restatement draws from exactly those lists."""
    sizes = [6] * 5 + [2, 1]
    x, labels = so.synthetic_embeddings(sizes, dim=32, sigma=1.0, seed=4, shuffle=False)
    alpha = 0.6
    d = mo.distance_matrix(x)
    out = mo.mine(x, labels, alpha=alpha)
    for a in range(x.shape[0]):
        for j, p in enumerate(out['pos_index'][a]):
            if p < 0:
                continue
            cand = [n for n in range(x.shape[0]) if labels[n] != labels[a] and np.float32(d[a, n] - d[a, p]) < np.float32(alpha)]
            assert len(cand) == out['eligible'][a, j]
            ks = np.arange(len(cand) + 1)
            got = mo.select_kth_eligible(x, labels, [a] * ks.size, [p] * ks.size, ks, alpha=alpha)
            np.testing.assert_array_equal(got, cand + [-1])
    trip, num_trips, m = mo.select_triplets_upstream(x, sizes, len(sizes), alpha, np.random.RandomState(0))
    assert num_trips == sum(k * (k - 1) // 2 for k in sizes) and m == trip.shape[0] > 0
    for a, p, n in trip:
        assert labels[a] == labels[p] and p > a and labels[n] != labels[a]
        assert np.float32(d[a, n] - d[a, p]) < np.float32(alpha)
    # the draw is reproducible from the RNG stream and the eligible counts alone (what the GPU replay relies on)
    rng = np.random.RandomState(0)
    replay = []
    for a in range(x.shape[0]):
        for j, p in enumerate(out['pos_index'][a]):
            if p > a and out['eligible'][a, j] > 0:
                k = rng.randint(int(out['eligible'][a, j]))
                replay.append((a, int(p), int(mo.select_kth_eligible(x, labels, [a], [p], [k], alpha=alpha)[0])))
    rng.shuffle(replay)
    np.testing.assert_array_equal(np.asarray(replay, dtype=np.int32), trip)


def test_pair_histogram_equals_reference_counts_summed(golden_dir):
    """The whole-set same / different histogram is the reference's per-class-pair counts (count_nonzero(sims < threshold),
    statistics.py:131, taken from the UNMODIFIED reference in tests/golden/confidence.npz) summed over the diagonal /
    off-diagonal class pairs."""
    g = np.load(golden_dir / 'confidence.npz')
    x, labels = g['embeddings'], g['labels']
    for metric in (0, 1):
        thr = so.default_thresholds(metric)
        counts = g['counts_m%d' % metric]                       # [C, C, T], lower triangle incl. diagonal
        same = np.einsum('iit->t', counts)
        diff = counts.sum(axis=(0, 1)) - same
        got = so.pair_histogram(x, labels, thr, metric)
        # whole-Gram vs per-block sgemm may differ in the last ulp for a pair sitting on a threshold
        assert np.abs(got['same'] - same).sum() + np.abs(got['diff'] - diff).sum() <= 2


def _edge_cases(golden_dir):
    g = np.load(golden_dir / 'validation_edge.npz')
    for name in (str(c) for c in g['cases']):
        metric, folds, far = g[name + '_cfg']
        yield g, name, int(metric), int(folds), float(far)


def test_validation_edge_cases_golden(golden_dir):
    """Edge cases of statistics.py:277-313 from the unmodified reference (oracle/gen_golden_edge.py): ragged folds, test
    folds without a same-identity pair (rates default to 1), two classes only."""
    for g, name, metric, folds, far in _edge_cases(golden_dir):
        x, labels = g[name + '_embeddings'], g[name + '_labels']
        for conf in (so.confidence_matrix_exact_order, so.confidence_matrix_weighted):
            out = so.face_to_face_validation(x, labels, metric, folds, far, confidence=conf)
            crit = [str(k) for k in g[name + '_criteria']]
            assert sorted(k for k in out if not k.startswith('_')) == crit
            for tag, key in (('acc', 'MaximumAccuracy'), ('far', [k for k in crit if k != 'MaximumAccuracy'][0])):
                keys = [str(k) for k in g['%s_%s_keys' % (name, tag)]]
                assert sorted(out[key].keys()) == keys
                got = np.array([float(out[key][k]) for k in keys])
                np.testing.assert_allclose(got, g['%s_%s_vals' % (name, tag)], rtol=0, atol=1e-9, err_msg='%s %s' % (name, key))
            np.testing.assert_array_equal(out['_thresholds'][:, 0], g[name + '_acc_thr'])
            np.testing.assert_allclose(out['_thresholds'][:, 1], g[name + '_far_thr'], rtol=0, atol=1e-12)


def test_false_examples_oracle_invariants():
    """The restated search of statistics.py:341-387: missed matches are same-identity pairs above the threshold, at most
    ``nrof_fpos_images`` per class, no image twice per class; false accepts are different-identity pairs below it, at most
    ``nrof_fneg_images`` per class pair, no row / column twice; each is the extreme of what was left."""
    x, labels = so.synthetic_embeddings([6, 1, 9, 3, 12, 1], dim=32, sigma=(1.0, 2.5), seed=1)
    thr = 1.7
    out = so.false_examples(x, labels, thr, nrof_fpos_images=3, nrof_fneg_images=2)
    d = mo.distance_matrix(x)
    assert out['fneg'] and out['fpos']
    seen = {}
    for dist, a, b in out['fneg']:
        assert labels[a] == labels[b] and dist > thr and abs(dist - d[a, b]) < 1e-6
        used = seen.setdefault(labels[a], set())
        assert a not in used and b not in used
        used.update((a, b))
    assert all(len(v) <= 6 for v in seen.values())
    first = {}
    for dist, a, b in out['fneg']:
        first.setdefault(labels[a], dist)
    for lab, dist in first.items():
        m = labels == lab
        assert abs(dist - d[np.ix_(m, m)].max()) < 1e-6
    per_pair = {}
    for dist, a, b in out['fpos']:
        assert labels[a] < labels[b] and dist < thr
        rows, cols = per_pair.setdefault((labels[a], labels[b]), (set(), set()))
        assert a not in rows and b not in cols
        rows.add(a); cols.add(b)
    assert all(len(r) <= 2 for r, _ in per_pair.values())


def test_argsort_restatement_matches_numpy_scalar_path():
    """``so.argsort_numpy_scalar`` (and the product's copy) against NumPy's own scalar introsort -- this container's numpy with
    every SIMD sort dispatch disabled, the only path numpy 1.19.4 had -- on arrays with long runs of ties (what fp_rates is)."""
    import os
    import subprocess
    import sys
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r)
from oracle import statistics_oracle as so
from facenet_b200 import statistics as fst
rng = np.random.default_rng(0)
for trial in range(200):
    n = int(rng.integers(1, 300))
    kind = trial %% 4
    if kind == 0: x = rng.integers(0, 5, n).astype(float)
    elif kind == 1: x = np.sort(rng.integers(0, 8, n).astype(float))
    elif kind == 2:
        z = int(rng.integers(0, n)); o = int(rng.integers(0, n - z + 1))
        x = np.concatenate([np.zeros(z), np.sort(rng.random(n - z - o)), np.ones(o)])
    else: x = rng.random(n)
    ref = np.argsort(x)
    assert np.array_equal(so.argsort_numpy_scalar(x), ref), (trial, n, kind)
    assert np.array_equal(fst._argsort_numpy_scalar(x), ref), (trial, n, kind)
x = np.concatenate([np.zeros(40), np.linspace(0.001, 0.9, 45), np.ones(15)])
assert not np.array_equal(np.argsort(x), np.arange(100)), 'the scalar path permutes ties: SIMD dispatch still active?'
print('ok')
''' % str(Path(__file__).resolve().parent.parent)
    env = dict(os.environ, NPY_DISABLE_CPU_FEATURES='AVX512F AVX512CD AVX512_SKX AVX512_CLX AVX512_CNL AVX512_ICL AVX512_SPR AVX2')
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, env=env, timeout=300)
    if 'SIMD dispatch still active' in out.stderr:
        pytest.skip('cannot disable the SIMD sort dispatch of this numpy build')
    assert out.returncode == 0 and out.stdout.strip().endswith('ok'), out.stderr[-2000:]


def test_far_threshold_behind_a_run_of_tied_fp_rates():
    """under scipy 1.4.1 / numpy 1.19 interp1d(kind='slinear') sorts (fp_rates, thresholds) with an UNSTABLE
    argsort; when far_target lies right behind a run of equal fp_rates the left sample is whichever tied threshold that sort
    put last -- not the largest one.  Oracle, product host code and the stable formula on a hand-made case."""
    from facenet_b200 import statistics as fst
    thr = np.linspace(0, 4, 100)
    fpr = np.concatenate([np.zeros(40), np.linspace(0.004, 0.9, 45), np.ones(15)])          # first non-zero rate above the target
    far = 1.e-3
    ind = so.argsort_numpy_scalar(fpr)
    assert sorted(ind[:40].tolist()) == list(range(40)) and ind[39] != 39                     # the zeros come out permuted
    j = int(ind[39])                                                                          # the tied sample the sort puts last
    w = 1.0 / (fpr[40] - 0.0)
    expect = thr[j] * ((fpr[40] - far) * w) + thr[40] * ((far - 0.0) * w)
    got = so.slinear_interp(fpr, thr, far)
    assert got == expect == float(fst._slinear(fpr, thr, far))
    stable = thr[39] + far / fpr[40] * (thr[40] - thr[39])
    assert abs(got - stable) > 0.1                                                            # the deviation the advisor flagged
    # no tie at the bracket: the sorted order is irrelevant and the result is ordinary linear interpolation
    far2 = 0.3
    k = int(np.searchsorted(fpr, far2, side='right')) - 1
    lin = thr[k] + (far2 - fpr[k]) / (fpr[k + 1] - fpr[k]) * (thr[k + 1] - thr[k])
    assert abs(so.slinear_interp(fpr, thr, far2) - lin) < 1e-14
    with pytest.raises(ValueError, match='above the interpolation range'):
        so.slinear_interp(fpr[:50], thr[:50], 0.99)
