"""Generate ``tests/golden/faceclass.npz`` by running the UNMODIFIED reference files
``facenet/faceclass.py`` and ``facenet/apps/train_classifier.py`` (NumPy stand-in for TensorFlow, see
``oracle.reference_loader.load_reference_faceclass``).  Build container only:

    python -m oracle.gen_golden_faceclass
"""
import json
import sys
import types
from pathlib import Path

import numpy as np

from oracle import statistics_oracle as so
from oracle.reference_loader import load_reference_faceclass

OUT = Path(__file__).resolve().parent.parent / 'tests' / 'golden'


def set_vars(model, **kw):
    for k, v in kw.items():
        model.variables[k].value = np.float32(v)


def main():
    fc, tc = load_reference_faceclass()
    sizes = [1, 2, 5, 9, 17, 3, 1, 30, 4]
    x, labels = so.synthetic_embeddings(sizes, dim=64, sigma=1.6, seed=21, shuffle=False)
    rng = np.random.default_rng(4)
    scale = rng.uniform(0.6, 1.7, size=(x.shape[0], 1)).astype(np.float32)
    xu = (x * scale).astype(np.float32)                      # un-normalised embeddings
    bounds = np.concatenate([[0], np.cumsum(sizes)])
    split_n = [x[a:b] for a, b in zip(bounds[:-1], bounds[1:])]
    split_u = [xu[a:b] for a, b in zip(bounds[:-1], bounds[1:])]
    out = {'x': x, 'xu': xu, 'labels': labels, 'sizes': np.array(sizes),
           'versions': json.dumps({'numpy': np.__version__, 'python': sys.version.split()[0],
                                   'shim': 'NumPy stand-in for tensorflow (oracle/reference_loader.py)'})}

    mn = fc.FaceToFaceNormalizedEmbeddingsClassifier()
    md = fc.FaceToFaceDistanceClassifier()
    set_vars(md, theta=0.7)
    out['theta'] = np.float32(0.7)
    out['norm_self'] = mn.distance(x[:40], None)
    out['norm_cross'] = mn.distance(x[:23], x[23:])
    out['dist_self'] = md.distance(xu[:40], None)
    out['dist_cross'] = md.distance(xu[:23], xu[23:])
    thresholds = np.array([0.85, 1.0, 1.3], dtype=np.float32)
    out['thresholds'] = thresholds
    cms_n, cms_d = [], []
    for t in thresholds:
        set_vars(mn, threshold=t)
        set_vars(md, threshold=t)
        c = tc.ConfusionMatrix(split_n, mn)
        cms_n.append([c.accuracy, c.precision, c.tp_rate, c.tn_rate])
        c = tc.ConfusionMatrix(split_u, md)
        cms_d.append([c.accuracy, c.precision, c.tp_rate, c.tn_rate])
    out['confusion_norm'] = np.array(cms_n, dtype=np.float64)
    out['confusion_dist'] = np.array(cms_d, dtype=np.float64)
    set_vars(mn, threshold=1.0)
    set_vars(md, threshold=1.0)
    out['predict_norm'] = mn.predict(x[:40])
    out['predict_dist'] = md.predict(xu[:23], xu[23:])

    # binary cross entropy over a P x K batch (rows grouped by class, facenet.py:108-113)
    P, K = 6, 5
    xb, _ = so.synthetic_embeddings([K] * P, dim=64, sigma=1.3, seed=8, shuffle=False)
    xbu = (xb * rng.uniform(0.7, 1.5, size=(P * K, 1))).astype(np.float32)
    opt = types.SimpleNamespace(nrof_classes_per_batch=P, nrof_examples_per_class=K)
    out['batch'] = xb
    out['batch_u'] = xbu
    out['PK'] = np.array([P, K])
    set_vars(mn, alpha=10.0, threshold=1.0)
    set_vars(md, alpha=7.5, threshold=1.1, theta=0.7)
    out['bce_norm'] = np.float32(tc.binary_cross_entropy_loss(mn(xb), opt))
    out['bce_dist'] = np.float32(tc.binary_cross_entropy_loss(md(xbu), opt))
    out['bce_dist_vars'] = np.array([7.5, 1.1, 0.7], dtype=np.float32)
    np.savez_compressed(OUT / 'faceclass.npz', **out)
    print('faceclass.npz', (OUT / 'faceclass.npz').stat().st_size, 'bytes')


if __name__ == '__main__':
    main()
