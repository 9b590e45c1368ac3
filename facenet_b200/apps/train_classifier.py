"""The evaluation pieces of ``/root/reference/facenet/apps/train_classifier.py`` on a B200:

    ConfusionMatrix(embeddings, classifier)           train_classifier.py:17-49
    binary_cross_entropy_loss(logits, options)        train_classifier.py:60-84  (see ``pair_cross_entropy``)

``ConfusionMatrix`` takes the per-class list of embedding arrays (``facenet.Embeddings.data()``) and a classifier of
``facenet_b200.faceclass``.  The reference loops over all class pairs and takes the mean of ``classifier.predict`` per
block (the self block is the FULL n_i x n_i matrix, diagonal and both triangles); here the whole set goes through ONE
fused Gram + histogram launch with a single threshold, keyed by the pair of class sizes (the block means depend on the
classes only through their sizes), and the float64 rates are formed from the exact integer counts.
"""
import numpy as np

from facenet_b200 import _capi
from facenet_b200 import statistics as _st
from facenet_b200.faceclass import FaceToFaceDistanceClassifier

__all__ = ['ConfusionMatrix', 'binary_cross_entropy_loss', 'pair_cross_entropy']


class ConfusionMatrix:
    def __init__(self, embeddings, classifier):
        nrof_classes = len(embeddings)
        nrof_positive_class_pairs = nrof_classes
        nrof_negative_class_pairs = nrof_classes * (nrof_classes - 1) / 2

        sizes = np.array([len(e) for e in embeddings], dtype=np.int64)
        x = np.ascontiguousarray(np.concatenate([np.asarray(e, dtype=np.float32).reshape(len(e), -1) for e in embeddings], axis=0))
        cls = np.repeat(np.arange(nrof_classes), sizes)
        threshold = classifier.variable('threshold', mode='numpy')
        perm, cls_sorted, regions, ia, ib, gsize, gcount = _st._size_group_plan(cls, sizes)
        regions['key'] = 2 * np.arange(regions.size)          # slot 2j: pairs of rectangle j, slot 2j + 1: its diagonal
        regions['tri'] = np.where(ia == ib, 2, 0)
        thr = np.array([threshold], dtype=np.float64)
        opts = classifier._gram_options() if isinstance(classifier, FaceToFaceDistanceClassifier) else {}
        bins, self.stats = _st._handle().region_histogram_bins(x, perm, cls_sorted, regions, 2 * regions.size, thr, metric=0,
                                                               mode=_st._state['mode'], cta_group=_st._state['cta_group'],
                                                               raw_distance=True, **opts)
        lt = _st._counts_lt(bins.astype(np.int64), _capi.numpy_cuts(thr, 0))[..., 0]     # [2 * regions, 2]: d < threshold
        pairs_all, pairs_same, diag = lt[0::2, 0], lt[0::2, 1], lt[1::2, 0]

        # train_classifier.py:26-37: mean over the full n x n self block = (2 upper + diagonal) / n^2, one term per class
        on_diag = ia == ib
        tp = float(np.sum((2.0 * pairs_same[on_diag] + diag[on_diag]) / (gsize[ia[on_diag]].astype(np.float64) ** 2)))
        fn = nrof_classes - tp
        fp = float(np.sum((pairs_all - pairs_same) / (gsize[ia] * gsize[ib]).astype(np.float64)))
        tn = nrof_negative_class_pairs - fp

        tp /= nrof_positive_class_pairs
        fn /= nrof_positive_class_pairs

        fp /= nrof_negative_class_pairs
        tn /= nrof_negative_class_pairs

        self.classifier = classifier
        self.accuracy = (tp + tn) / (tp + fp + tn + fn)
        self.precision = tp / (tp + fp)
        self.tp_rate = tp / (tp + fn)
        self.tn_rate = tn / (tn + fp)

    def __repr__(self):
        return (f'{self.__class__.__name__}\n' +
                f'{str(self.classifier)}\n' +
                f'accuracy  {self.accuracy}\n' +
                f'precision {self.precision}\n' +
                f'tp rate   {self.tp_rate}\n' +
                f'tn rate   {self.tn_rate}\n')


def binary_cross_entropy_loss(logits, options):
    """train_classifier.py:60-84 for a materialised ``[B, B]`` logits matrix (``B = nrof_classes_per_batch *
    nrof_examples_per_class``): labels 1 for pairs of the same class on the strict upper triangle,
    ``pos_weight = len(labels) / sum(labels) - 1``, mean weighted cross entropy.  Returns a float32 scalar like
    ``session.run(loss)``.  The training loop should call ``pair_cross_entropy`` instead, which never forms the logits."""
    batch_size = options.nrof_classes_per_batch * options.nrof_examples_per_class
    logits = np.ascontiguousarray(logits, dtype=np.float32) if isinstance(logits, np.ndarray) or not hasattr(logits, '__dlpack__') else logits
    if tuple(logits.shape) != (batch_size, batch_size):
        raise ValueError('logits must be [{0}, {0}]'.format(batch_size))
    return np.float32(_st._handle().logits_cross_entropy(logits, options.nrof_examples_per_class))


def pair_cross_entropy(model, embeddings_batch, options):
    """``binary_cross_entropy_loss(model(embeddings_batch), options)`` (train_classifier.py:109-110) and its derivatives with
    respect to the classifier's variables, from ONE fused Gram launch -- what one optimiser step needs
    (train_classifier.py:127).  Returns ``{'loss', 'grads': {'alpha', 'threshold'[, 'theta']}, 'pos_weight', 'stats'}``."""
    opts = model._gram_options()
    out = _st._handle().pair_cross_entropy(embeddings_batch, options.nrof_examples_per_class, float(model.variable('alpha')),
                                           float(model.variable('threshold')), **opts)
    grads = {'alpha': out['dalpha'], 'threshold': out['dthreshold']}
    if 'theta' in model.variables:
        grads['theta'] = out['dtheta']
    return {'loss': out['loss'], 'grads': grads, 'pos_weight': out['pos_weight'], 'stats': out['stats']}
