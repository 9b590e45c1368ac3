"""Drop-in for the pairwise classifiers of sMedX/FaceNet (``/root/reference/facenet/faceclass.py``), evaluated on a B200.

    FaceToFaceDistanceClassifier                 faceclass.py:8-76
    FaceToFaceNormalizedEmbeddingsClassifier     faceclass.py:79-118

Same names, methods and arithmetic as the reference's NumPy branch (``distance`` / ``predict`` on ``np.ndarray``
inputs, ``__call__`` -> logits ``alpha * (threshold - distance)``).  The reference keeps ``alpha``, ``threshold`` and
``theta`` in ``tf.Variable``s read through a session; here they are plain float32 values in ``.variables`` (settable),
``variable(name, mode)`` returns them either way.

Every distance comes from the fused Gram kernel of the CUDA library (``fnb_pairwise`` with
``fnb_options.raw_distance``: no range check and no clamp, exactly like faceclass.py:106-116; the un-normalised
classifier adds normalise-on-load and the norm term of faceclass.py:71 in the epilogue).  No CPU fallback.
"""
import numpy as np

from facenet_b200 import _capi
from facenet_b200 import statistics as _st

__all__ = ['FaceToFaceDistanceClassifier', 'FaceToFaceNormalizedEmbeddingsClassifier']


class _PairClassifier:
    _defaults = {}

    def __init__(self):
        self.variables = {k: np.float32(v) for k, v in self._defaults.items()}

    def __call__(self, x, y=None):
        # faceclass.py:23-27 / 86-90
        alpha = self.variable('alpha')
        threshold = self.variable('threshold')
        return np.multiply(alpha, np.subtract(threshold, self.distance(x, y)))

    def __repr__(self):
        variables = {name: self.variable(name, mode='numpy') for name in self.variables.keys()}
        return (f'{self.__class__.__name__}\n'
                f'variables {variables}\n')

    def variable(self, name, mode=None):
        return np.float32(self.variables[name])

    def _gram_options(self):
        return {}

    def distance(self, x, y):
        x = np.ascontiguousarray(x, dtype=np.float32)
        y = x if y is None else np.ascontiguousarray(y, dtype=np.float32)
        if x.shape[0] == 0 or y.shape[0] == 0:
            return np.empty((x.shape[0], y.shape[0]), dtype=np.float32)
        return _st._handle().pairwise(x, y, 0, mode=_st._state['mode'], cta_group=_st._state['cta_group'],
                                      raw_distance=True, **self._gram_options())

    def predict(self, x, y=None):
        # faceclass.py:76 / 118
        return self.distance(x, y) < self.variable('threshold', mode='numpy')


class FaceToFaceDistanceClassifier(_PairClassifier):
    """normalized distance between embeddings (faceclass.py:8-76):

    ``distance = 2 (1 - (x/|x|, y/|y|)) + theta * (2 (|x| - |y|) / (|x| + |y|))^2``"""
    _defaults = {'alpha': 10, 'threshold': 1, 'theta': 1}

    def _gram_options(self):
        return {'normalize': 1, 'theta': float(self.variable('theta'))}


class FaceToFaceNormalizedEmbeddingsClassifier(_PairClassifier):
    """``distance = 2 (1 - (x, y))`` on embeddings the caller normalised (faceclass.py:79-118)."""
    _defaults = {'alpha': 10.0, 'threshold': 1.0}
