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
(fp_rates always has repeated 0s/1s).  Under the pinned scipy 1.4.1 it built a
k=1 B-spline == piecewise-linear interpolation on the bracketing interval.  The
module attribute ``interpolate`` is replaced by a proxy that special-cases
``kind='slinear'`` (see ``slinear_interp``); every other attribute is scipy's.
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
    """Piecewise-linear interpolation as scipy 1.4.1's ``interp1d(kind='slinear')``
    evaluated it for a non-decreasing ``x`` with duplicates: use the LAST interval
    [x[j], x[j+1]] with x[j] <= xq (j clipped so j+1 is valid), slope form.
    """
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    xq = float(xq)
    if xq < x[0] or xq > x[-1]:
        raise ValueError('A value in x_new is outside the interpolation range.')
    j = int(np.searchsorted(x, xq, side='right')) - 1
    j = min(max(j, 0), x.size - 2)
    if x[j + 1] == x[j]:
        return np.float64(y[j])
    return np.float64(y[j] + (xq - x[j]) / (x[j + 1] - x[j]) * (y[j + 1] - y[j]))


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
