"""Load the reference's ``facenet/statistics.py`` UNMODIFIED from the read-only mount.

Test infrastructure (see ``oracle/__init__.py``).  Only usable where
``/root/reference`` exists (the build container); never on the GPU box.

``import facenet.statistics`` fails without TensorFlow because
``facenet/__init__.py:10`` imports it and ``statistics.py:19`` imports
``utils, ioutils, h5utils`` (TF/h5py/PIL).  None of those are used by the
arithmetic, so a stub package is registered in ``sys.modules`` and the file is
executed as-is.

One shim: ``statistics.py:301`` calls
``scipy.interpolate.interp1d(fp_rates, thresholds, kind='slinear')`` which under
this container's scipy (1.18) raises ``ValueError`` for duplicate abscissae
(fp_rates always has repeated 0s/1s).  Under the pinned scipy 1.4.1 it sorted the
samples with ``np.argsort`` (numpy 1.19: unstable scalar introsort, ties come out
permuted) and built a k=1 B-spline.  The module attribute ``interpolate`` is replaced
by a proxy that special-cases ``kind='slinear'`` with a step-by-step restatement of
exactly that (see ``statistics_oracle.slinear_interp``); every other attribute is scipy's.
"""
import importlib.util
import sys
import types
from pathlib import Path

import numpy as np

REFERENCE_ROOT = Path('/root/reference')
REFERENCE_STATISTICS = REFERENCE_ROOT / 'facenet' / 'statistics.py'


def reference_available():
    return REFERENCE_STATISTICS.is_file()


def slinear_interp(x, y, xq):
    """scipy 1.4.1's ``interp1d(kind='slinear')`` restated step by step, including the unstable ``np.argsort`` of numpy 1.19 that
    permutes tied abscissae: ``oracle.statistics_oracle.slinear_interp``."""
    from oracle import statistics_oracle as so
    return so.slinear_interp(x, y, xq)


class _InterpolateProxy:
    def __init__(self, real):
        self._real = real

    def __getattr__(self, name):
        return getattr(self._real, name)

    def interp1d(self, x, y, kind='linear', **kw):
        if kind == 'slinear':
            xs = np.asarray(x, dtype=np.float64)
            ys = np.asarray(y, dtype=np.float64)
            return lambda xq: np.array(slinear_interp(xs, ys, xq))
        return self._real.interp1d(x, y, kind=kind, **kw)


_cached = None


def load_reference_statistics():
    """Return the reference ``statistics`` module (executed unmodified)."""
    global _cached
    if _cached is not None:
        return _cached
    if not reference_available():
        raise FileNotFoundError(str(REFERENCE_STATISTICS))

    saved = {k: sys.modules.get(k) for k in
             ('facenet', 'facenet.utils', 'facenet.ioutils', 'facenet.h5utils', 'facenet.statistics')}
    try:
        pkg = types.ModuleType('facenet')
        pkg.__path__ = []
        for sub in ('utils', 'ioutils', 'h5utils'):
            m = types.ModuleType('facenet.' + sub)
            setattr(pkg, sub, m)
            sys.modules['facenet.' + sub] = m
        sys.modules['facenet'] = pkg
        spec = importlib.util.spec_from_file_location('facenet.statistics', str(REFERENCE_STATISTICS))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v

    mod.interpolate = _InterpolateProxy(mod.interpolate)
    # silence the per-validation logger.info(self) (statistics.py:266)
    mod.logger = types.SimpleNamespace(info=lambda *a, **k: None)
    _cached = mod
    return mod


# ---------------------------------------------------------------------------------------
# faceclass.py / apps/train_classifier.py

REFERENCE_FACECLASS = REFERENCE_ROOT / 'facenet' / 'faceclass.py'
REFERENCE_TRAIN_CLASSIFIER = REFERENCE_ROOT / 'facenet' / 'apps' / 'train_classifier.py'


class _StubVariable:
    def __init__(self, initial_value=None, dtype=None, name=None, trainable=True):
        self.value = np.float32(initial_value)
        self.name = name


class _StubSession:
    def run(self, var):
        return var.value if isinstance(var, _StubVariable) else var


def weighted_cross_entropy_with_logits(labels, logits, pos_weight):
    """``tf.nn.weighted_cross_entropy_with_logits`` as its documentation states it (TensorFlow is absent here):
    ``(1 - z) * x + l * (log1p(exp(-|x|)) + max(-x, 0))`` with ``l = 1 + (q - 1) * z``, evaluated in the logits' dtype."""
    x = np.asarray(logits)
    z = np.asarray(labels, dtype=x.dtype)
    q = x.dtype.type(pos_weight)
    one = x.dtype.type(1)
    lw = one + (q - one) * z
    return (one - z) * x + lw * (np.log1p(np.exp(-np.abs(x))) + np.maximum(-x, x.dtype.type(0)))


def _numpy_tf():
    """A minimal NumPy stand-in for the TensorFlow names the two files touch."""
    tf = types.ModuleType('tensorflow')
    tf.float32 = np.float32
    tf.float64 = np.float64
    tf.Variable = _StubVariable
    tf.get_default_session = lambda: _StubSession()
    tf.multiply = lambda a, b: np.multiply(getattr(a, 'value', a), getattr(b, 'value', b))
    tf.subtract = lambda a, b: np.subtract(getattr(a, 'value', a), getattr(b, 'value', b))
    tf.transpose = np.transpose
    tf.gather_nd = lambda params, idx: np.asarray(params)[tuple(np.asarray(idx).T)]
    tf.constant = lambda v, dtype=None: np.asarray(v, dtype=dtype)
    tf.reduce_mean = lambda v: np.mean(v, dtype=np.asarray(v).dtype)
    tf.linalg = types.SimpleNamespace(norm=np.linalg.norm)
    tf.nn = types.SimpleNamespace(weighted_cross_entropy_with_logits=weighted_cross_entropy_with_logits)
    return tf


_cached_fc = None


def faceclass_available():
    return REFERENCE_FACECLASS.is_file() and REFERENCE_TRAIN_CLASSIFIER.is_file()


def load_reference_faceclass():
    """Return ``(faceclass, train_classifier)``: the reference modules executed unmodified with a NumPy stand-in
    for TensorFlow (only ``Variable``, ``get_default_session().run``, a few elementwise ops, ``gather_nd`` and
    ``nn.weighted_cross_entropy_with_logits`` -- the last restated from its documentation)."""
    global _cached_fc
    if _cached_fc is not None:
        return _cached_fc
    if not faceclass_available():
        raise FileNotFoundError(str(REFERENCE_FACECLASS))
    names = ('tensorflow', 'facenet', 'facenet.config', 'facenet.facenet', 'facenet.faceclass', 'facenet.ioutils',
             'facenet.apps', 'facenet.apps.train_classifier')
    saved = {k: sys.modules.get(k) for k in names}
    try:
        sys.modules['tensorflow'] = _numpy_tf()
        pkg = types.ModuleType('facenet')
        pkg.__path__ = []
        sys.modules['facenet'] = pkg
        for sub in ('config', 'facenet', 'ioutils'):
            m = types.ModuleType('facenet.' + sub)
            setattr(pkg, sub, m)
            sys.modules['facenet.' + sub] = m
        spec = importlib.util.spec_from_file_location('facenet.faceclass', str(REFERENCE_FACECLASS))
        fc = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(fc)
        pkg.faceclass = fc
        sys.modules['facenet.faceclass'] = fc
        spec = importlib.util.spec_from_file_location('facenet.apps.train_classifier', str(REFERENCE_TRAIN_CLASSIFIER))
        tc = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(tc)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _cached_fc = (fc, tc)
    return _cached_fc
