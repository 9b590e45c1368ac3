"""The classifier / ConfusionMatrix / cross-entropy restatement (oracle/faceclass_oracle.py) against committed outputs
of the unmodified reference (tests/golden/faceclass.npz, made by oracle/gen_golden_faceclass.py) and, when
/root/reference is mounted, against the reference itself on fresh inputs."""
import types

import numpy as np
import pytest

from oracle import faceclass_oracle as fo
from oracle import statistics_oracle as so


def _split(x, sizes):
    b = np.concatenate([[0], np.cumsum(sizes)])
    return [x[a:c] for a, c in zip(b[:-1], b[1:])]


def test_distances_golden(golden_dir):
    g = np.load(golden_dir / 'faceclass.npz')
    x, xu, theta = g['x'], g['xu'], float(g['theta'])
    np.testing.assert_array_equal(fo.distance_normalized(x[:40]), g['norm_self'])
    np.testing.assert_array_equal(fo.distance_normalized(x[:23], x[23:]), g['norm_cross'])
    np.testing.assert_array_equal(fo.distance_unnormalized(xu[:40], None, theta), g['dist_self'])
    np.testing.assert_array_equal(fo.distance_unnormalized(xu[:23], xu[23:], theta), g['dist_cross'])
    np.testing.assert_array_equal(fo.predict(fo.distance_normalized(x[:40]), 1.0), g['predict_norm'])
    np.testing.assert_array_equal(fo.predict(fo.distance_unnormalized(xu[:23], xu[23:], theta), 1.0), g['predict_dist'])


def test_confusion_matrix_golden(golden_dir):
    g = np.load(golden_dir / 'faceclass.npz')
    sizes, theta = g['sizes'], float(g['theta'])
    cls = np.repeat(np.arange(sizes.size), sizes)
    for tag, x, fn in (('norm', g['x'], fo.distance_normalized),
                       ('dist', g['xu'], lambda a, b: fo.distance_unnormalized(a, b, theta))):
        emb = _split(x, sizes)
        for t, ref in zip(g['thresholds'], g['confusion_' + tag]):
            lit = fo.confusion_matrix(emb, fn, t)
            np.testing.assert_array_equal([lit.accuracy, lit.precision, lit.tp_rate, lit.tn_rate], ref)
            vec = fo.confusion_matrix_vectorized(x, cls, fn, t, block=32)
            np.testing.assert_allclose([vec.accuracy, vec.precision, vec.tp_rate, vec.tn_rate], ref, rtol=0, atol=2e-3)
            # the integer form reproduces the literal rates exactly (same float64 operations per class pair)
            upper, diag, cross = fo.confusion_counts(emb, fn, t)
            tp = sum((2 * upper[i] + diag[i]) / float(sizes[i]) ** 2 for i in range(sizes.size)) / sizes.size
            assert abs(tp - lit.tp) < 1e-12


def test_cross_entropy_golden(golden_dir):
    g = np.load(golden_dir / 'faceclass.npz')
    P, K = (int(v) for v in g['PK'])
    lg = fo.logits(fo.distance_normalized(g['batch']), 10.0, 1.0)
    assert fo.binary_cross_entropy_loss(lg, P, K) == g['bce_norm']
    a, t, th = (float(v) for v in g['bce_dist_vars'])
    lg = fo.logits(fo.distance_unnormalized(g['batch_u'], None, th), a, t)
    assert fo.binary_cross_entropy_loss(lg, P, K) == g['bce_dist']
    # float64 loss + analytic gradients against the float32 value and finite differences
    d = fo.distance_unnormalized(g['batch_u'], None, th)
    out = fo.binary_cross_entropy_loss_and_grads(d, a, t, P, K)
    assert abs(out['loss'] - float(g['bce_dist'])) < 2e-6 * max(1.0, abs(out['loss']))
    h = 1e-6
    up = fo.binary_cross_entropy_loss_and_grads(d, a + h, t, P, K)['loss']
    dn = fo.binary_cross_entropy_loss_and_grads(d, a - h, t, P, K)['loss']
    assert abs((up - dn) / (2 * h) - out['dalpha']) < 1e-6
    up = fo.binary_cross_entropy_loss_and_grads(d, a, t + h, P, K)['loss']
    dn = fo.binary_cross_entropy_loss_and_grads(d, a, t - h, P, K)['loss']
    assert abs((up - dn) / (2 * h) - out['dthreshold']) < 1e-6


def test_pair_labels_rule():
    i, k, z = fo.pair_labels(3, 2)
    assert z.tolist() == [1, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1]
    assert len(z) / z.sum() - 1 == 4.0


@pytest.mark.reference
def test_against_live_reference():
    from oracle.reference_loader import load_reference_faceclass
    fc, tc = load_reference_faceclass()
    sizes = [3, 1, 7, 12, 2]
    x, _ = so.synthetic_embeddings(sizes, dim=32, sigma=1.4, seed=77, shuffle=False)
    xu = (x * np.random.default_rng(1).uniform(0.5, 2.0, size=(x.shape[0], 1))).astype(np.float32)
    md = fc.FaceToFaceDistanceClassifier()
    md.variables['theta'].value = np.float32(1.3)
    md.variables['threshold'].value = np.float32(0.9)
    np.testing.assert_array_equal(md.distance(xu, None), fo.distance_unnormalized(xu, None, 1.3))
    ref = tc.ConfusionMatrix(_split(xu, sizes), md)
    got = fo.confusion_matrix(_split(xu, sizes), lambda a, b: fo.distance_unnormalized(a, b, 1.3), 0.9)
    assert (ref.accuracy, ref.precision, ref.tp_rate, ref.tn_rate) == (got.accuracy, got.precision, got.tp_rate, got.tn_rate)
    mn = fc.FaceToFaceNormalizedEmbeddingsClassifier()
    opt = types.SimpleNamespace(nrof_classes_per_batch=5, nrof_examples_per_class=4)
    xb, _ = so.synthetic_embeddings([4] * 5, dim=32, sigma=1.2, seed=3, shuffle=False)
    assert tc.binary_cross_entropy_loss(mn(xb), opt) == fo.binary_cross_entropy_loss(fo.logits(fo.distance_normalized(xb), 10, 1), 5, 4)
