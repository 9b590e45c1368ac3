"""GPU parity of the pairwise-classifier path through the C ABI: classifier distances,
predictions and ConfusionMatrix against the CPU oracle (oracle/faceclass_oracle.py) and the committed outputs of the
unmodified reference (tests/golden/faceclass.npz); normalise-on-load against pre-normalised input.

Tolerances: distances within 1e-5 absolute (BASELINE.json); predictions / block means exact except for pairs whose
distance lies within eps = 1e-5 of the threshold (counted by the kernel)."""
import numpy as np
import pytest

from oracle import faceclass_oracle as fo
from oracle import statistics_oracle as so

pytestmark = pytest.mark.gpu

DIST_TOL = 1.e-5


def _split(x, sizes):
    b = np.concatenate([[0], np.cumsum(sizes)])
    return [x[a:c] for a, c in zip(b[:-1], b[1:])]


def _unnormalised(sizes, dim, seed):
    x, labels = so.synthetic_embeddings(sizes, dim=dim, sigma=1.5, seed=seed, shuffle=False)
    scale = np.random.default_rng(seed + 1).uniform(0.5, 1.9, size=(x.shape[0], 1)).astype(np.float32)
    return x, (x * scale).astype(np.float32), labels


def test_distances_golden(golden_dir):
    from facenet_b200 import faceclass
    g = np.load(golden_dir / 'faceclass.npz')
    mn = faceclass.FaceToFaceNormalizedEmbeddingsClassifier()
    md = faceclass.FaceToFaceDistanceClassifier()
    md.variables['theta'] = g['theta']
    for got, ref in ((mn.distance(g['x'][:40], None), g['norm_self']), (mn.distance(g['x'][:23], g['x'][23:]), g['norm_cross']),
                     (md.distance(g['xu'][:40], None), g['dist_self']), (md.distance(g['xu'][:23], g['xu'][23:]), g['dist_cross'])):
        assert got.dtype == np.float32 and got.shape == ref.shape
        assert np.abs(got - ref).max() <= DIST_TOL
    # predictions: equal wherever the reference distance is not within the tolerance of the threshold
    for m, x, y, key, dkey in ((mn, g['x'][:40], None, 'predict_norm', 'norm_self'), (md, g['xu'][:23], g['xu'][23:], 'predict_dist', 'dist_cross')):
        got = m.predict(x, y)
        clear = np.abs(g[dkey] - 1.0) > DIST_TOL
        np.testing.assert_array_equal(got[clear], g[key][clear])
    # logits (faceclass.py:23-27)
    lg = md(g['xu'][:40])
    ref = fo.logits(g['dist_self'], 10, 1)
    assert lg.dtype == np.float32 and np.abs(lg - ref).max() <= 10 * DIST_TOL


@pytest.mark.parametrize('na,nb,d', [(1, 1, 64), (130, 70, 128), (300, 513, 512)])
def test_distances_ragged_shapes(na, nb, d):
    from facenet_b200 import faceclass
    _, xa, _ = _unnormalised([na], d, 5)
    _, xb, _ = _unnormalised([nb], d, 9)
    md = faceclass.FaceToFaceDistanceClassifier()
    md.variables['theta'] = np.float32(1.7)
    assert np.abs(md.distance(xa, xb) - fo.distance_unnormalized(xa, xb, 1.7)).max() <= DIST_TOL
    assert np.abs(md.distance(xa, None) - fo.distance_unnormalized(xa, None, 1.7)).max() <= DIST_TOL
    mn = faceclass.FaceToFaceNormalizedEmbeddingsClassifier()
    # no range check and no clamp (faceclass.py:106-116): un-normalised input goes straight through
    ref = fo.distance_normalized(xa, xb)
    assert np.abs(mn.distance(xa, xb) - ref).max() <= DIST_TOL * max(1.0, float(np.abs(ref).max()))


def test_confusion_matrix_golden(golden_dir):
    from facenet_b200 import faceclass
    from facenet_b200.apps import train_classifier as tc
    g = np.load(golden_dir / 'faceclass.npz')
    sizes = g['sizes']
    for tag, x, model in (('norm', g['x'], faceclass.FaceToFaceNormalizedEmbeddingsClassifier()),
                          ('dist', g['xu'], faceclass.FaceToFaceDistanceClassifier())):
        if tag == 'dist':
            model.variables['theta'] = g['theta']
        for t, ref in zip(g['thresholds'], g['confusion_' + tag]):
            model.variables['threshold'] = np.float32(t)
            cm = tc.ConfusionMatrix(_split(x, sizes), model)
            tol = 1e-12 + cm.stats['eps_window'] * 1.0           # a pair in the eps window may move a block mean
            np.testing.assert_allclose([cm.accuracy, cm.precision, cm.tp_rate, cm.tn_rate], ref, rtol=0, atol=tol)
            assert 'accuracy' in repr(cm) and model.__class__.__name__ in repr(cm)


@pytest.mark.parametrize('theta', [0.0, 1.3])
def test_confusion_matrix_ragged_vs_oracle(theta):
    from facenet_b200 import faceclass
    from facenet_b200.apps import train_classifier as tc
    sizes = [1, 1, 2, 2, 2, 3, 7, 7, 40, 41, 300, 5, 5, 5, 1, 90]
    x, xu, _ = _unnormalised(sizes, 128, 31)
    if theta == 0.0:
        model, data = faceclass.FaceToFaceNormalizedEmbeddingsClassifier(), x
        fn = fo.distance_normalized
    else:
        model, data = faceclass.FaceToFaceDistanceClassifier(), xu
        model.variables['theta'] = np.float32(theta)
        fn = lambda a, b: fo.distance_unnormalized(a, b, theta)
    model.variables['threshold'] = np.float32(1.05)
    cm = tc.ConfusionMatrix(_split(data, sizes), model)
    ref = fo.confusion_matrix(_split(data, sizes), fn, 1.05)
    # every pair outside the eps window is predicted like the oracle: the rates can differ by at most
    # (pairs in the window) x (largest weight of one pair, 1 / smallest block) -- and by float64 summation order
    tol = 1e-12 + cm.stats['eps_window'] * 1.0 / len(sizes)
    np.testing.assert_allclose([cm.accuracy, cm.precision, cm.tp_rate, cm.tn_rate],
                               [ref.accuracy, ref.precision, ref.tp_rate, ref.tn_rate], rtol=0, atol=tol)


def test_normalise_on_load_histogram():
    """fnb_options.normalize = 2 (tf.nn.l2_normalize, epsilon 1e-10) on raw rows == the histogram of rows normalised
    on the host with the same formula, up to pairs inside the eps window."""
    from facenet_b200 import _capi
    h = _capi.default_handle(0)
    sizes = [30] * 20 + [1] * 7
    x, xu, labels = _unnormalised(sizes, 256, 77)
    thr = so.default_thresholds(0)
    xn = (xu * (1.0 / np.sqrt(np.maximum((xu.astype(np.float32) ** 2).sum(axis=1, keepdims=True), np.float32(1e-10))))).astype(np.float32)
    ref = h.pair_histogram(xn, labels, thr, 0)
    got = h.pair_histogram(xu, labels, thr, 0, normalize=2)
    mism = int(np.abs(got['same'] - ref['same']).sum() + np.abs(got['diff'] - ref['diff']).sum())
    assert got['n_same'] == ref['n_same'] and got['n_diff'] == ref['n_diff']
    assert mism <= 2 * got['stats']['eps_window']
    with pytest.raises(_capi.FnbError):
        h.pair_histogram(xu, labels, thr, 0)               # not normalised and not asked to: statistics.py:40-42


# ------------------------------------------------------------------------------ f2 cross entropy

class _Opt:
    def __init__(self, p, k):
        self.nrof_classes_per_batch, self.nrof_examples_per_class = p, k


def test_cross_entropy_golden(golden_dir):
    from facenet_b200 import faceclass
    from facenet_b200.apps import train_classifier as tc
    g = np.load(golden_dir / 'faceclass.npz')
    P, K = (int(v) for v in g['PK'])
    mn = faceclass.FaceToFaceNormalizedEmbeddingsClassifier()
    out = tc.pair_cross_entropy(mn, g['batch'], _Opt(P, K))
    assert abs(out['loss'] - float(g['bce_norm'])) <= 2e-5 * max(1.0, abs(float(g['bce_norm'])))
    a, t, th = (float(v) for v in g['bce_dist_vars'])
    md = faceclass.FaceToFaceDistanceClassifier()
    md.variables.update(alpha=np.float32(a), threshold=np.float32(t), theta=np.float32(th))
    out = tc.pair_cross_entropy(md, g['batch_u'], _Opt(P, K))
    assert abs(out['loss'] - float(g['bce_dist'])) <= 2e-5 * max(1.0, abs(float(g['bce_dist'])))
    # the reference's own signature on a materialised logits matrix
    lg = md(g['batch_u'])
    assert abs(float(tc.binary_cross_entropy_loss(lg, _Opt(P, K))) - float(g['bce_dist'])) <= 2e-5 * max(1.0, abs(float(g['bce_dist'])))


@pytest.mark.parametrize('P,K,d', [(45, 40, 512), (7, 3, 64), (2, 130, 128)])
def test_cross_entropy_and_grads_vs_oracle(P, K, d):
    from facenet_b200 import faceclass
    from facenet_b200.apps import train_classifier as tc
    x, xu, _ = _unnormalised([K] * P, d, 100 + P)
    for model, data, dist in ((faceclass.FaceToFaceNormalizedEmbeddingsClassifier(), x, fo.distance_normalized(x)),
                              (faceclass.FaceToFaceDistanceClassifier(), xu, None)):
        a, t = 6.0, 1.2
        model.variables.update(alpha=np.float32(a), threshold=np.float32(t))
        dth = None
        if dist is None:
            model.variables['theta'] = np.float32(0.8)
            dist = fo.distance_unnormalized(xu, None, 0.8)
            nrm = np.linalg.norm(xu.astype(np.float64), axis=1)
            dth = (2 * (nrm[:, None] - nrm[None, :]) / (nrm[:, None] + nrm[None, :])) ** 2
        ref = fo.binary_cross_entropy_loss_and_grads(dist, a, t, P, K, dtheta_term=dth)
        out = tc.pair_cross_entropy(model, data, _Opt(P, K))
        assert abs(out['pos_weight'] - ref['pos_weight']) < 1e-12
        tol = 2e-5 * max(1.0, abs(ref['loss']))
        assert abs(out['loss'] - ref['loss']) <= tol
        assert abs(out['grads']['alpha'] - ref['dalpha']) <= 5e-5 * max(1.0, abs(ref['dalpha']))
        assert abs(out['grads']['threshold'] - ref['dthreshold']) <= 5e-5 * max(1.0, abs(ref['dthreshold']))
        if dth is not None:
            assert abs(out['grads']['theta'] - ref['dtheta']) <= 5e-5 * max(1.0, abs(ref['dtheta']))
        # float32 reference arithmetic (the golden form) agrees with the float64 statement to the same tolerance
        ref32 = float(fo.binary_cross_entropy_loss(fo.logits(dist, a, t), P, K))
        assert abs(out['loss'] - ref32) <= 5e-5 * max(1.0, abs(ref32))


def test_cross_entropy_rejects_bad_batches():
    from facenet_b200 import _capi
    h = _capi.default_handle(0)
    x = np.zeros((10, 64), dtype=np.float32)
    with pytest.raises(_capi.FnbError):
        h.pair_cross_entropy(x, 3, 10.0, 1.0)          # 10 rows are not P x 3
    with pytest.raises(_capi.FnbError):
        h.pair_cross_entropy(x, 1, 10.0, 1.0)          # K = 1: pos_weight undefined (division by zero in the reference)
