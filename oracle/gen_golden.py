"""Generate ``tests/golden/*.npz`` by running the UNMODIFIED reference
(``/root/reference/facenet/statistics.py`` via ``oracle.reference_loader``).

Run once in the build container (the only place ``/root/reference`` exists):

    python -m oracle.gen_golden

Inputs are seeded synthetic embeddings (``statistics_oracle.synthetic_embeddings``);
the fixtures store the inputs too, so the tests never depend on the RNG stream.
Library versions in effect are recorded in each file (``versions``).
"""
import json
import sys
from pathlib import Path

import numpy as np

from oracle import statistics_oracle as so
from oracle.reference_loader import load_reference_statistics

OUT = Path(__file__).resolve().parent.parent / 'tests' / 'golden'


def versions():
    import scipy
    import sklearn
    return json.dumps({'numpy': np.__version__, 'scipy': scipy.__version__, 'sklearn': sklearn.__version__,
                       'python': sys.version.split()[0],
                       'shim': "interp1d(kind='slinear') restated (oracle/reference_loader.py)"})


class Cfg:
    def __init__(self, metric, nrof_folds, far_target):
        self.metric, self.nrof_folds, self.far_target = metric, nrof_folds, far_target


def main():
    st = load_reference_statistics()
    OUT.mkdir(parents=True, exist_ok=True)

    # ---- A1: pairwise_similarities, self and cross form, both metrics; D = 128
    x, _ = so.synthetic_embeddings([37, 21], dim=128, sigma=1.1, seed=11, shuffle=False)
    xa, xb = x[:37], x[37:]
    out = {'xa': xa, 'xb': xb, 'versions': versions()}
    for metric in (0, 1):
        out['self_m%d' % metric] = st.pairwise_similarities(xa.copy(), metric=metric)
        out['cross_m%d' % metric] = st.pairwise_similarities(xa.copy(), xb.copy(), metric=metric)
    np.savez_compressed(OUT / 'pairwise.npz', **out)

    # ---- A2-A4: ConfidenceMatrix on a ragged set with singletons, a duplicate embedding
    #      and non-contiguous label values; D = 128, 100 thresholds, both metrics
    sizes = [1, 1, 1, 2, 2, 3, 5, 8, 13, 21, 1, 34, 4, 4, 1, 7]
    values = np.array([907, -3, 12, 5000, 77, 78, 79, 1 << 40, 4, 6, 8, 10, 11, 13, 15, 17])
    x, labels = so.synthetic_embeddings(sizes, dim=128, sigma=1.0, seed=5, shuffle=True, label_values=values)
    x[3] = x[60]                      # exact duplicate rows (distance exactly 0 if same row content)
    out = {'embeddings': x, 'labels': labels, 'versions': versions()}
    for metric in (0, 1):
        thr = np.linspace(0, 4 if metric == 0 else np.pi, 100)
        calc = st.SimilarityCalculator(x, labels, metric=metric)
        cm = st.ConfidenceMatrix(calc, thr)
        for name in ('tp', 'tn', 'fp', 'fn', 'accuracy', 'precision', 'tp_rates', 'tn_rates'):
            out['%s_m%d' % (name, metric)] = getattr(cm, name)
        # integer per-class-pair counts straight from the reference's calculator
        nc = calc.nrof_classes
        counts = np.zeros((nc, nc, thr.size), dtype=np.int64)
        for i in range(nc):
            for k in range(i + 1):
                sims, _ = calc.evaluate(i, k)
                if sims.size:
                    counts[i, k] = [np.count_nonzero(sims < t) for t in thr]
        out['counts_m%d' % metric] = counts
        # single (0-d) threshold as produced by interp1d (statistics.py:302,308)
        cm1 = st.ConfidenceMatrix(calc, np.array(thr[31] + 0.0123))
        out['single_m%d' % metric] = np.array([cm1.tp[0], cm1.tn[0], cm1.fp[0], cm1.fn[0]])
    np.savez_compressed(OUT / 'confidence.npz', **out)

    # ---- A5/A6: FaceToFaceValidation; D = 64, N = 330, 10 folds, both metrics
    sizes = [12] * 20 + [3] * 10 + [1] * 10 + [25, 25]
    x, labels = so.synthetic_embeddings(sizes, dim=64, sigma=2.2, seed=3, shuffle=True)
    out = {'embeddings': x, 'labels': labels, 'versions': versions()}
    for metric in (0, 1):
        v = st.FaceToFaceValidation(x, labels, Cfg(metric, 10, 1.e-3))
        for r, tag in zip(v.reports, ('acc', 'far')):
            dct = r.dict
            out['%s_keys_m%d' % (tag, metric)] = np.array(sorted(dct.keys()))
            out['%s_vals_m%d' % (tag, metric)] = np.array([float(dct[k]) for k in sorted(dct.keys())])
            out['%s_thr_m%d' % (tag, metric)] = np.array([float(m.threshold[0]) for m in r.conf_matrix_test])
            out['%s_test_m%d' % (tag, metric)] = np.array([[m.tp[0], m.tn[0], m.fp[0], m.fn[0]]
                                                           for m in r.conf_matrix_test])
        out['train_tp_m%d' % metric] = np.array([m.tp for m in v.reports[0].conf_matrix_train])
        out['train_fp_m%d' % metric] = np.array([m.fp for m in v.reports[0].conf_matrix_train])
        out['repr_m%d' % metric] = np.array(repr(v).split('elapsed_time')[0])
    np.savez_compressed(OUT / 'validation.npz', **out)

    for f in sorted(OUT.glob('*.npz')):
        print(f.name, f.stat().st_size, 'bytes')


if __name__ == '__main__':
    main()
