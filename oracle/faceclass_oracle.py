"""CPU restatement (NumPy) of the pairwise-classifier path of sMedX/FaceNet -- SURVEY.md section 8 (f1), (f2).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``): nothing under ``facenet_b200/`` imports this.

Restated from ``/root/reference``:

  facenet/faceclass.py:43-74     FaceToFaceDistanceClassifier.distance      -> distance_unnormalized
  facenet/faceclass.py:106-116   FaceToFaceNormalizedEmbeddingsClassifier.distance -> distance_normalized
  facenet/faceclass.py:23-27,76  __call__ (logits), predict                 -> logits, predict
  facenet/apps/train_classifier.py:17-49   ConfusionMatrix                  -> confusion_matrix (literal loops),
                                                                               confusion_matrix_vectorized
  facenet/apps/train_classifier.py:60-84   binary_cross_entropy_loss        -> pair_labels, binary_cross_entropy_loss
                                                                               (+ its gradients, which TensorFlow derives)

Pinning: ``tests/golden/faceclass.npz`` holds outputs of the unmodified reference files executed with a NumPy
stand-in for the handful of TensorFlow names they touch (``oracle/reference_loader.load_reference_faceclass``).
``tf.nn.weighted_cross_entropy_with_logits`` itself is absent (no TensorFlow in this image): its documented formula
is restated, so the LOSS VALUE is pinned to the reference's label / pos_weight / gather logic but "parity unpinned"
against TensorFlow's own kernel.  The gradients have no reference code (TensorFlow autodiff): analytic derivatives
of the same formula, checked against finite differences in the tests.
"""
import numpy as np

F32 = np.float32


# ---------------------------------------------------------------------------------------
# distances (float32 arithmetic like the reference's NumPy branch)

def distance_normalized(x, y=None):
    """faceclass.py:106-116: ``2 * (1 - x @ y.T)`` -- no range check, no clamp."""
    x = np.asarray(x, dtype=F32)
    y = x if y is None else np.asarray(y, dtype=F32)
    return 2 * (1 - x @ np.transpose(y))


def distance_unnormalized(x, y=None, theta=1.0):
    """faceclass.py:43-74: rows are normalised, and ``theta * (2 (|x| - |y|) / (|x| + |y|))^2`` is added."""
    x = np.asarray(x, dtype=F32)
    y = x if y is None else np.asarray(y, dtype=F32)
    theta = F32(theta)
    yt = np.transpose(y)
    norm_x = np.linalg.norm(x, axis=1, keepdims=True)
    norm_y = np.linalg.norm(yt, axis=0, keepdims=True)
    x1 = x / norm_x
    y1 = yt / norm_y
    return 2 * (1 - x1 @ y1) + theta * pow(2 * (norm_x - norm_y) / (norm_x + norm_y), 2)


def logits(dist, alpha, threshold):
    """faceclass.py:23-27: ``alpha * (threshold - distance)`` in float32."""
    return np.multiply(F32(alpha), np.subtract(F32(threshold), dist))


def predict(dist, threshold):
    """faceclass.py:76,118: ``distance < threshold`` (float32 compare)."""
    return dist < F32(threshold)


# ---------------------------------------------------------------------------------------
# ConfusionMatrix (train_classifier.py:17-49)

class Confusion:
    def __init__(self, tp, tn, fp, fn):
        self.tp, self.tn, self.fp, self.fn = tp, tn, fp, fn
        self.accuracy = (tp + tn) / (tp + fp + tn + fn)
        self.precision = tp / (tp + fp)
        self.tp_rate = tp / (tp + fn)
        self.tn_rate = tn / (tn + fp)


def confusion_matrix(embeddings, distance_fn, threshold):
    """Literal loop order of train_classifier.py:26-37: for i: for k < i: cross block mean; then the FULL
    n_i x n_i self block (diagonal and both triangles included) mean."""
    nc = len(embeddings)
    tp = tn = fp = fn = 0
    for i in range(nc):
        for k in range(i):
            m = np.mean(predict(distance_fn(embeddings[i], embeddings[k]), threshold))
            fp += m
            tn += 1 - m
        m = np.mean(predict(distance_fn(embeddings[i], None), threshold))
        tp += m
        fn += 1 - m
    npos = nc
    nneg = nc * (nc - 1) / 2
    return Confusion(tp / npos, tn / nneg, fp / nneg, fn / npos)


def confusion_counts(embeddings, distance_fn, threshold):
    """Integer form: per class, predictions true in the strict upper triangle / on the diagonal of the self block;
    per class pair (i > k), predictions true in the cross block.  Returns (upper[C], diag[C], cross[C, C])."""
    nc = len(embeddings)
    upper = np.zeros(nc, dtype=np.int64)
    diag = np.zeros(nc, dtype=np.int64)
    cross = np.zeros((nc, nc), dtype=np.int64)
    for i in range(nc):
        p = predict(distance_fn(embeddings[i], None), threshold)
        diag[i] = int(np.trace(p))
        upper[i] = int(np.triu(p, 1).sum())
        for k in range(i):
            cross[i, k] = int(predict(distance_fn(embeddings[i], embeddings[k]), threshold).sum())
    return upper, diag, cross


def confusion_matrix_vectorized(x, cls, distance_fn, threshold, block=2048):
    """Same rates from one blocked pass over the whole set (``cls`` = class rank per row); for sizes where the
    per-class-pair loop is too slow.  The self block mean counts every ordered pair and the diagonal."""
    x = np.asarray(x, dtype=F32)
    cls = np.asarray(cls)
    nc = int(cls.max()) + 1
    sizes = np.bincount(cls, minlength=nc).astype(np.float64)
    w_row = 1.0 / sizes[cls]
    tp_sum = 0.0
    fp_sum = 0.0
    for r0 in range(0, x.shape[0], block):
        r1 = min(x.shape[0], r0 + block)
        p = predict(distance_fn(x[r0:r1], x), threshold)
        same = cls[r0:r1, None] == cls[None, :]
        w = w_row[r0:r1, None] * w_row[None, :]
        tp_sum += float((w * (p & same)).sum())
        fp_sum += float((w * (p & ~same)).sum()) / 2.0       # every unordered cross pair appears twice
    npos = nc
    nneg = nc * (nc - 1) / 2
    tp = tp_sum / npos
    fp = fp_sum / nneg
    return Confusion(tp, 1.0 - fp, fp, 1.0 - tp)


# ---------------------------------------------------------------------------------------
# binary cross-entropy over the upper triangle of a P x K batch (train_classifier.py:60-84)

def pair_labels(nrof_classes_per_batch, nrof_examples_per_class):
    """train_classifier.py:62-73: label 1 iff ``i // K == k // K`` on ``np.triu_indices(B, k=1)``."""
    b = nrof_classes_per_batch * nrof_examples_per_class
    i, k = np.triu_indices(b, k=1)
    return i, k, ((i // nrof_examples_per_class) == (k // nrof_examples_per_class)).astype(np.int64)


def weighted_cross_entropy_with_logits(labels, x, pos_weight):
    """TensorFlow's documented formula: ``(1 - z) x + (1 + (q - 1) z) (log1p(exp(-|x|)) + max(-x, 0))``."""
    x = np.asarray(x)
    z = np.asarray(labels, dtype=x.dtype)
    one = x.dtype.type(1)
    lw = one + (x.dtype.type(pos_weight) - one) * z
    return (one - z) * x + lw * (np.log1p(np.exp(-np.abs(x))) + np.maximum(-x, x.dtype.type(0)))


def binary_cross_entropy_loss(logit_matrix, nrof_classes_per_batch, nrof_examples_per_class, dtype=None):
    """train_classifier.py:60-84: ``pos_weight = len(labels) / sum(labels) - 1``; mean of the weighted cross entropy
    of the ``B (B - 1) / 2`` upper-triangle logits."""
    i, k, z = pair_labels(nrof_classes_per_batch, nrof_examples_per_class)
    pos_weight = len(z) / z.sum() - 1
    lg = np.asarray(logit_matrix)[i, k]
    if dtype is not None:
        lg = lg.astype(dtype)
    ce = weighted_cross_entropy_with_logits(z, lg, pos_weight)
    return np.mean(ce, dtype=lg.dtype)


def binary_cross_entropy_loss_and_grads(dist_matrix, alpha, threshold, nrof_classes_per_batch, nrof_examples_per_class,
                                        dtheta_term=None):
    """float64 loss of ``logits = alpha (threshold - d)`` and its derivatives with respect to alpha, threshold (and
    theta when ``dtheta_term`` = d(distance)/d(theta) is given) -- what TensorFlow's autodiff hands the optimiser
    (train_classifier.py:127).  d loss / d logit = (1 - z) - (1 + (q - 1) z) sigmoid(-x)."""
    i, k, z = pair_labels(nrof_classes_per_batch, nrof_examples_per_class)
    q = len(z) / z.sum() - 1
    d = np.asarray(dist_matrix, dtype=np.float64)[i, k]
    zf = z.astype(np.float64)
    x = alpha * (threshold - d)
    lw = 1 + (q - 1) * zf
    loss = np.mean((1 - zf) * x + lw * (np.log1p(np.exp(-np.abs(x))) + np.maximum(-x, 0)))
    sig_neg = np.where(x >= 0, np.exp(-x) / (1 + np.exp(-x)), 1 / (1 + np.exp(x)))
    g = (1 - zf) - lw * sig_neg
    out = {'loss': float(loss), 'dalpha': float(np.mean(g * (threshold - d))), 'dthreshold': float(np.mean(g * alpha)),
           'pos_weight': float(q)}
    if dtheta_term is not None:
        out['dtheta'] = float(np.mean(g * (-alpha) * np.asarray(dtheta_term, dtype=np.float64)[i, k]))
    return out
