"""CPU tests of the product's host side: the C-ABI library loads and exports every symbol the header
declares, the similarity-cut tables reproduce NumPy's distance comparison, and the class-balanced
rate arithmetic of facenet_b200.statistics equals the reference's (golden fixtures) when the
histogram entry point is replaced by a NumPy stand-in (tests/emulator.py).  No GPU compute here."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

from facenet_b200 import _capi, statistics as fst
from oracle import statistics_oracle as so
from tests import emulator

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    lib = _capi.load_library()
    header = (ROOT / 'include' / 'facenet_b200.h').read_text()
    declared = set(re.findall(r'\b(fnb_[a-z_0-9]+)\s*\(', header))
    assert declared, 'no declarations parsed'
    assert declared == set(_capi.EXPORTS)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert lib.fnb_version() >= 100


def test_struct_layouts_match_header(tmp_path):
    """sizeof and the offset of EVERY field of the ctypes mirrors equal what a C compiler sees in the header (a field added in
    the middle of fnb_options / fnb_stats on one side only would shift every option behind it silently)."""
    import subprocess
    structs = (('fnb_options', _capi.Options), ('fnb_stats', _capi.Stats), ('fnb_region', _capi.Region), ('DLTensor', _capi.DLTensor))
    lines = []
    for cname, cls in structs:
        lines.append('printf("%%zu\\n", sizeof(%s));' % cname)
        for fname, _ in cls._fields_:
            lines.append('printf("%%zu\\n", offsetof(%s, %s));' % (cname, fname))
    src = tmp_path / 'sizes.c'
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "facenet_b200.h"\nint main(void) {\n' + '\n'.join(lines) + '\nreturn 0; }\n')
    exe = tmp_path / 'sizes'
    subprocess.run(['gcc', '-I', str(ROOT / 'include'), str(src), '-o', str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    want = []
    for _, cls in structs:
        want.append(ctypes.sizeof(cls))
        want += [getattr(cls, fname).offset for fname, _ in cls._fields_]
    assert got == want
    assert _capi.REGION_DTYPE.itemsize == ctypes.sizeof(_capi.Region)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(_capi.FnbError, match='no CUDA device'):
        _capi.Handle(0)
    with pytest.raises(_capi.FnbError):
        fst.pairwise_similarities(np.eye(4, 64, dtype=np.float32))


def _dist(s, metric):
    s = np.clip(s.astype(np.float32), np.float32(-1), np.float32(1))
    return 2 * (1 - s) if metric == 0 else np.arccos(s)


@pytest.mark.parametrize('metric', [0, 1])
def test_numpy_cuts_reproduce_distance_comparison(metric):
    thr = so.default_thresholds(metric)
    cuts = _capi.numpy_cuts(thr, metric)
    rng = np.random.default_rng(0)
    fin = np.isfinite(cuts)
    probe = np.concatenate([cuts[fin], np.nextafter(cuts[fin], np.float32(-2)), np.nextafter(cuts[fin], np.float32(2)),
                            rng.uniform(-1, 1, 20000).astype(np.float32), np.float32([-1, 1, 0])])
    d = _dist(probe, metric)
    for t, c in zip(thr, cuts):
        np.testing.assert_array_equal(d.astype(np.float64) < t, np.clip(probe, -1, 1) >= c)
    assert np.isinf(cuts[0])                   # nothing is < 0


@pytest.mark.parametrize('metric', [0, 1])
def test_counts_from_bins_c_abi(metric):
    """fnb_counts_from_bins (host-only C entry point) + cut tables against direct counting."""
    lib = _capi.load_library()
    thr = so.default_thresholds(metric)
    rng = np.random.default_rng(1)
    s = rng.uniform(-1.05, 1.05, 50000).astype(np.float32)
    same = rng.random(s.size) < 0.3
    cuts = _capi.numpy_cuts(thr, metric)
    order = np.sort(cuts)
    k = np.searchsorted(order, np.clip(s, -1, 1), side='right')
    bins = np.stack([np.bincount(k, minlength=thr.size + 1), np.bincount(k[same], minlength=thr.size + 1)]).astype(np.uint64)
    h = _capi.Handle.__new__(_capi.Handle)
    h.lib = lib
    h.h = None
    out = h.counts_from_bins(bins, thr, metric=metric, cuts=cuts)
    d = _dist(s, metric).astype(np.float64)
    np.testing.assert_array_equal(out['same'], [(d[same] < t).sum() for t in thr])
    np.testing.assert_array_equal(out['diff'], [(d[~same] < t).sum() for t in thr])
    assert out['n_same'] == same.sum() and out['n_diff'] == (~same).sum()
    if metric == 0:
        # library-computed cuts (libm) are identical for metric 0
        out2 = h.counts_from_bins(bins, thr, metric=metric, cuts=None)
        np.testing.assert_array_equal(out2['same'], out['same'])
    np.testing.assert_array_equal(fst._counts_lt(bins, cuts)[1], out['same'])


def test_size_group_plan_invariants():
    rng = np.random.default_rng(3)
    sizes = np.array([1, 1, 5, 2, 2, 7, 1, 5, 3])
    cls = rng.permutation(np.repeat(np.arange(sizes.size), sizes))
    perm, cls_sorted, regions, ia, ib, gsize, gcount = fst._size_group_plan(cls, sizes)
    assert sorted(perm.tolist()) == list(range(cls.size))
    assert np.all(np.diff(cls_sorted) >= 0)
    assert list(gsize) == [1, 2, 3, 5, 7] and list(gcount) == [3, 2, 1, 2, 1]
    # every unordered pair is covered exactly once
    cover = np.zeros((cls.size, cls.size), dtype=int)
    for r in regions:
        blk = np.ones((r['row_end'] - r['row_begin'], r['col_end'] - r['col_begin']), dtype=int)
        if r['tri']:
            blk = np.triu(blk, 1)
        cover[r['row_begin']:r['row_end'], r['col_begin']:r['col_end']] += blk
    assert np.array_equal(cover, np.triu(np.ones_like(cover), 1))
    # rows of one rectangle side all have the same class size
    size_of_row = sizes[cls][perm]
    for r in regions:
        assert len(set(size_of_row[r['row_begin']:r['row_end']])) == 1
        assert len(set(size_of_row[r['col_begin']:r['col_end']])) == 1


def test_kfold_split_matches_sklearn():
    from sklearn.model_selection import KFold
    for n, k in ((330, 10), (13, 5)):
        ref = list(KFold(n_splits=k, shuffle=True, random_state=0).split(np.arange(n)))
        for (a, b), (c, d) in zip(ref, fst.kfold_split(n, k)):
            np.testing.assert_array_equal(a, c)
            np.testing.assert_array_equal(b, d)


@pytest.fixture
def emulated(monkeypatch):
    handle = emulator.EmulatedHandle()
    monkeypatch.setattr(fst, '_handle', lambda: handle)


def test_confidence_matrix_host_math_golden(emulated, golden_dir):
    g = np.load(golden_dir / 'confidence.npz')
    x, labels = g['embeddings'], g['labels']
    for metric in (0, 1):
        thr = so.default_thresholds(metric)
        cm = fst.ConfidenceMatrix(fst.SimilarityCalculator(x, labels, metric), thr)
        for name in ('tp', 'tn', 'fp', 'fn', 'accuracy', 'precision', 'tp_rates', 'tn_rates'):
            np.testing.assert_allclose(getattr(cm, name), g['%s_m%d' % (name, metric)], rtol=0, atol=1e-3, err_msg=name)
        exact = so.confidence_matrix_exact_order(x, labels, thr, metric)
        for name in ('tp', 'tn', 'fp', 'fn'):
            np.testing.assert_allclose(getattr(cm, name), getattr(exact, name), rtol=0, atol=1e-13, err_msg=name)
        one = fst.ConfidenceMatrix(fst.SimilarityCalculator(x, labels, metric), np.array(thr[31] + 0.0123))
        assert one.threshold.shape == (1,)
        np.testing.assert_allclose([one.tp[0], one.tn[0], one.fp[0], one.fn[0]], g['single_m%d' % metric], rtol=0, atol=1e-3)


def test_validation_host_math_golden(emulated, golden_dir):
    g = np.load(golden_dir / 'validation.npz')
    x, labels = g['embeddings'], g['labels']

    class Cfg:
        nrof_folds, far_target = 10, 1.e-3

    for metric in (0, 1):
        Cfg.metric = metric
        v = fst.FaceToFaceValidation(x, labels, Cfg)
        for r, tag in zip(v.reports, ('acc', 'far')):
            dct = r.dict
            keys = [str(k) for k in g['%s_keys_m%d' % (tag, metric)]]
            assert sorted(dct.keys()) == keys
            np.testing.assert_allclose([float(dct[k]) for k in keys], g['%s_vals_m%d' % (tag, metric)], rtol=0, atol=2e-3)
            got_thr = np.array([float(m.threshold[0]) for m in r.conf_matrix_test])
            if tag == 'acc':
                np.testing.assert_array_equal(got_thr, g['acc_thr_m%d' % metric])
            else:
                np.testing.assert_allclose(got_thr, g['far_thr_m%d' % metric], rtol=0, atol=2e-3)
        assert repr(v).split('elapsed_time')[0].splitlines()[:3] == str(g['repr_m%d' % metric]).splitlines()[:3]
        assert set(v.dict.keys()) == {'MaximumAccuracy', 'FalseAlarmRate(FAR = 0.001)'}


def test_validation_rejects_bad_metric_and_length(emulated):
    class Cfg:
        metric, nrof_folds, far_target = 2, 10, 1.e-3
    x = np.eye(20, 64, dtype=np.float32)
    with pytest.raises(ValueError, match='Undefined similarity metric 2'):
        fst.FaceToFaceValidation(x, np.arange(20) // 2, Cfg)
    Cfg.metric = 0
    with pytest.raises(AssertionError):
        fst.FaceToFaceValidation(x, np.arange(19), Cfg)


def test_split_embeddings_matches_reference_semantics():
    x = np.arange(24, dtype=np.float32).reshape(8, 3)
    labels = np.array([5, -1, 5, 7, -1, -1, 7, 100])
    got = fst.split_embeddings(x, labels)
    ref = so.split_embeddings(x, labels)
    assert len(got) == len(ref)
    for a, b in zip(got, ref):
        np.testing.assert_array_equal(a, b)


def test_confusion_matrix_host_math_golden(emulated, golden_dir):
    """ConfusionMatrix plan (size-group rectangles, diagonal slots, block-mean weights) with the library's keyed histogram
    replaced by the NumPy stand-in: must reproduce the unmodified reference's rates (tests/golden/faceclass.npz)."""
    from facenet_b200 import faceclass
    from facenet_b200.apps import train_classifier as tc
    g = np.load(golden_dir / 'faceclass.npz')
    sizes = g['sizes']
    b = np.concatenate([[0], np.cumsum(sizes)])
    for tag, x, model in (('norm', g['x'], faceclass.FaceToFaceNormalizedEmbeddingsClassifier()),
                          ('dist', g['xu'], faceclass.FaceToFaceDistanceClassifier())):
        if tag == 'dist':
            model.variables['theta'] = g['theta']
        emb = [x[a:c] for a, c in zip(b[:-1], b[1:])]
        for t, ref in zip(g['thresholds'], g['confusion_' + tag]):
            model.variables['threshold'] = np.float32(t)
            cm = tc.ConfusionMatrix(emb, model)
            np.testing.assert_allclose([cm.accuracy, cm.precision, cm.tp_rate, cm.tn_rate], ref, rtol=0, atol=1e-12)


def test_validation_subset_rows_and_normalise_on_load(emulated, golden_dir):
    """The GPU-resident hand-off path (every fold = a row subset of ONE array, l2_normalize on load) through the NumPy
    stand-in: raw (scaled) embeddings with config.normalize give the golden reports of the normalised embeddings."""
    g = np.load(golden_dir / 'validation.npz')
    x, labels = g['embeddings'], g['labels']
    scale = np.random.default_rng(0).uniform(0.5, 3.0, size=(x.shape[0], 1)).astype(np.float32)

    class Cfg:
        metric, nrof_folds, far_target, normalize = 0, 10, 1.e-3, True

    v = fst.FaceToFaceValidation((x * scale).astype(np.float32), labels, Cfg)
    for r, tag in zip(v.reports, ('acc', 'far')):
        dct = r.dict
        keys = [str(k) for k in g['%s_keys_m0' % tag]]
        np.testing.assert_allclose([float(dct[k]) for k in keys], g['%s_vals_m0' % tag], rtol=0, atol=2e-3)
    calc = fst.SimilarityCalculator((x * scale).astype(np.float32), labels[10:50], _rows=np.arange(10, 50), normalize=True)
    assert calc.nrof_classes == np.unique(labels[10:50]).size
    assert abs(float(np.linalg.norm(calc.embeddings[0][0])) - 1.0) < 1e-6


def test_validation_edge_cases_host_math_golden(emulated, golden_dir):
    """The drop-in FaceToFaceValidation (host logic over the emulated library) on the reference's outputs for ragged folds
    (N % k != 0), test folds without a same-identity pair and a two-class set (tests/golden/validation_edge.npz)."""
    g = np.load(golden_dir / 'validation_edge.npz')
    for name in (str(c) for c in g['cases']):
        x, labels = g[name + '_embeddings'], g[name + '_labels']

        class Cfg:
            metric, nrof_folds, far_target = int(g[name + '_cfg'][0]), int(g[name + '_cfg'][1]), float(g[name + '_cfg'][2])

        v = fst.FaceToFaceValidation(x, labels, Cfg)
        assert sorted(v.dict.keys()) == [str(k) for k in g[name + '_criteria']]
        for r, tag in zip(v.reports, ('acc', 'far')):
            dct = r.dict
            keys = [str(k) for k in g['%s_%s_keys' % (name, tag)]]
            assert sorted(dct.keys()) == keys
            np.testing.assert_allclose([float(dct[k]) for k in keys], g['%s_%s_vals' % (name, tag)], rtol=0, atol=1e-6,
                                       err_msg='%s %s' % (name, tag))
            got_thr = np.array([float(m.threshold[0]) for m in r.conf_matrix_test])
            if tag == 'acc':
                np.testing.assert_array_equal(got_thr, g[name + '_acc_thr'])
            else:
                np.testing.assert_allclose(got_thr, g[name + '_far_thr'], rtol=0, atol=1e-9)
            got_test = np.array([[m.tp[0], m.tn[0], m.fp[0], m.fn[0]] for m in r.conf_matrix_test])
            np.testing.assert_allclose(got_test, g['%s_%s_test' % (name, tag)], rtol=0, atol=1e-12)
        np.testing.assert_allclose(np.array([m.tp for m in v.reports[0].conf_matrix_train]), g[name + '_train_tp'], rtol=0, atol=1e-12)
        np.testing.assert_allclose(np.array([m.fp for m in v.reports[0].conf_matrix_train]), g[name + '_train_fp'], rtol=0, atol=1e-12)
        assert repr(v).split('elapsed_time')[0].splitlines()[:3] == str(g[name + '_repr']).splitlines()[:3]


@pytest.mark.parametrize('seed', range(8))
def test_confidence_matrix_host_math_random_ragged_sets(emulated, seed):
    """Size-group rectangles + fp64 weights (statistics._size_group_plan) against the literal class-pair loop of the oracle
    (statistics.py:115-138) on random ragged sets: singletons, repeated sizes, one dominant class, shuffled rows,
    non-contiguous label values, both metrics, grids and single thresholds."""
    rng = np.random.default_rng(1000 + seed)
    nc = int(rng.integers(2, 40))
    sizes = rng.choice([1, 1, 2, 2, 3, 5, 8, 13], size=nc).tolist()
    if seed % 2:
        sizes[int(rng.integers(0, nc))] = int(rng.integers(40, 90))
    values = rng.choice(np.arange(-50, 5000), size=nc, replace=False)
    x, labels = so.synthetic_embeddings(sizes, dim=64, sigma=float(rng.uniform(0.8, 2.5)), seed=seed, shuffle=True, label_values=values)
    metric = seed % 2
    for thr in (so.default_thresholds(metric), np.array(float(rng.uniform(0.5, 2.5)))):
        cm = fst.ConfidenceMatrix(fst.SimilarityCalculator(x, labels, metric), thr)
        ref = so.confidence_matrix_exact_order(x, labels, thr, metric)
        for name in ('tp', 'tn', 'fp', 'fn'):
            np.testing.assert_allclose(getattr(cm, name), getattr(ref, name), rtol=0, atol=1e-13, err_msg='%s seed %d' % (name, seed))
        for name in ('accuracy', 'precision', 'tp_rates', 'tn_rates', 'fp_rates', 'fn_rates'):
            np.testing.assert_allclose(getattr(cm, name), getattr(ref, name), rtol=0, atol=1e-12, err_msg=name)


# ------------------------------------------------------------------------------ boundary hardening

def test_raw_dlpack_capsule_is_consumed_per_protocol():
    """A raw "dltensor" capsule (what tf.experimental.dlpack.to_dlpack / torch.utils.dlpack.to_dlpack return) is renamed
    "used_dltensor" and its deleter is called exactly when the taken-over tensor is released (SURVEY.md section 8 b)."""
    import gc
    import sys
    x = np.arange(12, dtype=np.float32).reshape(3, 4)
    base = sys.getrefcount(x)
    cap = x.__dlpack__()
    assert sys.getrefcount(x) > base                                 # the managed tensor holds the array
    t = _capi.from_dlpack(cap)
    assert isinstance(t, _capi.DLPackTensor) and t.shape == (3, 4) and not t.is_cuda and t.__dlpack_device__()[0] == 1
    assert _capi._pyapi.PyCapsule_IsValid(cap, b'used_dltensor') and not _capi._pyapi.PyCapsule_IsValid(cap, b'dltensor')
    b = _capi.Borrowed(t)                                            # the C ABI sees the DLTensor in place
    assert b.ptr.contents.data == x.ctypes.data and b.ptr.contents.ndim == 2
    with pytest.raises(ValueError):
        _capi.from_dlpack(cap)                                       # a consumed capsule cannot be taken twice
    del b
    t.release()
    del cap
    gc.collect()
    assert sys.getrefcount(x) == base                                # deleter ran once: the array is free again
    t.release()                                                      # idempotent
    assert _capi.from_dlpack(x) is x                                 # everything else passes through


def test_confidence_matrix_accepts_a_reference_style_calculator(emulated):
    """ConfidenceMatrix needs only the reference's surface of its calculator (statistics.py:82-108: .embeddings list, .metric,
    .nrof_classes): a foreign calculator gives the same matrix as the repo's own."""
    x, labels = so.synthetic_embeddings([5, 1, 9, 2, 14], dim=64, sigma=1.2, seed=3)
    thr = so.default_thresholds(0)

    class ForeignCalculator:                                         # e.g. the reference's own class, or a user's
        def __init__(self, embeddings, labels, metric=0):
            self.metric = metric
            self.embeddings = [embeddings[labels == v] for v in np.unique(labels)]

        @property
        def nrof_classes(self):
            return len(self.embeddings)

    own = fst.ConfidenceMatrix(fst.SimilarityCalculator(x, labels, 0), thr)
    foreign = fst.ConfidenceMatrix(ForeignCalculator(x, labels), thr)
    for name in ('tp', 'tn', 'fp', 'fn'):
        np.testing.assert_array_equal(getattr(own, name), getattr(foreign, name))
    ref = so.ConfidenceMatrix(so.SimilarityCalculator(x, labels, 0), thr)
    np.testing.assert_allclose(foreign.tp, ref.tp, atol=1e-12)
    np.testing.assert_allclose(foreign.fp, ref.fp, atol=1e-12)


def test_validate_callback_shape_of_the_reference(emulated):
    """facenet/callbacks.py:21-28 + facenet.py:184-201: model outputs per batch -> np.concatenate -> FaceToFaceValidation."""
    from facenet_b200 import callbacks
    x, labels = so.synthetic_embeddings([6] * 12 + [1] * 8, dim=64, sigma=1.5, seed=2)
    batches = [(x[i:i + 16], labels[i:i + 16]) for i in range(0, x.shape[0], 16)]

    class Validate:
        metric, nrof_folds, far_target = 0, 4, 1.e-2

    class Config:
        validate = Validate

    cb = callbacks.ValidateCallback(lambda images: images, batches, every_n_epochs=2, max_nrof_epochs=5, config=Config)
    cb.on_epoch_end(0)
    assert cb.validation is None                                     # epoch 1: not due
    cb.on_epoch_end(1)
    assert cb.validation is not None
    direct = fst.FaceToFaceValidation(x, labels, Validate)
    for a, b in zip(cb.validation.reports, direct.reports):
        assert a.dict == b.dict
    cb.validation = None
    cb.on_epoch_end(4)                                               # last epoch always validates
    assert cb.validation is not None
    e, l = callbacks.evaluate_embeddings(lambda images: images, batches)
    np.testing.assert_array_equal(e, x)
    np.testing.assert_array_equal(l, labels)


def test_kfold_split_rejects_what_sklearn_rejects():
    with pytest.raises(ValueError, match='at least one train/test split'):
        list(fst.kfold_split(10, 1))
    with pytest.raises(ValueError, match='greater than the number of samples'):
        list(fst.kfold_split(3, 4))
    from sklearn.model_selection import KFold
    for bad in ((10, 1), (3, 4)):
        with pytest.raises(ValueError):
            list(KFold(n_splits=bad[1], shuffle=True, random_state=0).split(np.arange(bad[0])))


def test_write_dict_appends_resizable_gzip_datasets(monkeypatch, tmp_path):
    """h5utils.write_dict (facenet/h5utils.py:9-26) against an in-memory stand-in for h5py (absent in this image): nested keys
    become group paths, every dataset is 1-D, resizable and gzip-compressed, and a second write appends."""
    import sys
    import types

    store = {}

    class Dataset:
        def __init__(self, data, maxshape, compression, dtype):
            self.data = np.array(data, dtype=dtype)
            self.maxshape, self.compression = maxshape, compression

        @property
        def shape(self):
            return self.data.shape

        def resize(self, size, axis=0):
            assert self.maxshape == (None,) and axis == 0
            self.data = np.concatenate([self.data, np.zeros(size - self.data.shape[0], dtype=self.data.dtype)])

        def __setitem__(self, key, value):
            self.data[key] = value

    class File:
        def __init__(self, name, mode='a'):
            assert mode == 'a'
            self.ds = store.setdefault(name, {})

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

        def __contains__(self, name):
            return name in self.ds

        def __getitem__(self, name):
            return self.ds[name]

        def create_dataset(self, name, data, maxshape, compression, dtype):
            self.ds[name] = Dataset(data, maxshape, compression, dtype)

    monkeypatch.setitem(sys.modules, 'h5py', types.SimpleNamespace(File=File))
    from facenet_b200 import h5utils
    f = tmp_path / 'report.h5'
    h5utils.write_dict(f, {'MaximumAccuracy': {'auc': 0.97, 'eer': np.float64(0.08)}, 'elapsed': 1.5}, group='epoch')
    h5utils.write_dict(f, {'MaximumAccuracy': {'auc': 0.98, 'eer': np.float64(0.07)}, 'elapsed': 1.25}, group='epoch')
    ds = store[str(f)]
    assert sorted(ds) == ['epoch/MaximumAccuracy/auc', 'epoch/MaximumAccuracy/eer', 'epoch/elapsed']
    np.testing.assert_array_equal(ds['epoch/MaximumAccuracy/auc'].data, [0.97, 0.98])
    np.testing.assert_array_equal(ds['epoch/elapsed'].data, [1.5, 1.25])
    assert all(d.compression == 'gzip' and d.maxshape == (None,) and d.data.ndim == 1 for d in ds.values())
    h5utils.write_dict(f, {'x': 1})                                  # no group: top level
    assert 'x' in ds


def test_chunk_plan_covers_every_pair_once():
    """The column-chunk plan of a streamed pass (host rows uploaded / shards broadcast chunk by chunk under the Gram launches):
    over all chunks, every pair (row < col) of the set lies in exactly one region, the regions of chunk k only touch rows and
    columns below chunk k's end (the rows fed so far), and the chunks grow with the work already queued.  Pure host logic of
    the library (fnb_debug_chunk_plan): runs without a GPU."""
    import ctypes
    from facenet_b200 import _capi
    lib = _capi.load_library()
    rng = np.random.default_rng(0)
    cases = [(1, 4, 2), (2, 1, 1), (37, 8, 8), (64, 16, 8), (100, 16, 24), (257, 32, 16), (300, 64, 32), (301, 50, 7)]
    cases += [(int(rng.integers(2, 400)), int(rng.integers(1, 90)), int(rng.integers(1, 90))) for _ in range(12)]
    for n, rr, g in cases:
        cap = 4096
        buf = (ctypes.c_int * (cap * 6))()
        nch = ctypes.c_int()
        cnt = lib.fnb_debug_chunk_plan(n, rr, g, cap, buf, ctypes.byref(nch))
        assert 0 <= cnt <= cap
        regs = np.frombuffer(buf, dtype=np.int32)[:cnt * 6].reshape(cnt, 6)
        cover = np.zeros((n, n), dtype=np.int32)
        ends = {}
        for k, r0, r1, c0, c1, tri in regs:
            assert 0 <= r0 < r1 <= n and 0 <= c0 < c1 <= n
            blk = np.ones((r1 - r0, c1 - c0), dtype=np.int32)
            if tri:
                assert (r0, r1) == (c0, c1)
                blk = np.triu(blk, 1)
            cover[r0:r1, c0:c1] += blk
            ends[k] = max(ends.get(k, 0), c1)
            assert r1 <= ends[k] or r1 <= c1                  # rows of a region lie at or above its chunk's last column
        want = np.triu(np.ones((n, n), dtype=np.int32), 1)
        np.testing.assert_array_equal(cover, want)
        assert sorted(ends) == list(range(nch.value)) or n == 1
        e = [ends[k] for k in sorted(ends)]
        assert e == sorted(e) and (not e or e[-1] == n)
        # chunk k needs only rows < its own end: no region of chunk k reaches past ends[k]
        for k, r0, r1, c0, c1, tri in regs:
            assert r1 <= ends[k] and c1 <= ends[k]
