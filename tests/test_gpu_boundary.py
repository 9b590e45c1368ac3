"""GPU tests of the drop-in boundary (SURVEY.md section 8 b / A7): raw DLPack capsules, the callback-shaped caller, a foreign
calculator, dimensions the tensor-core tiles do not divide, the staging path for pageable host arrays."""
import numpy as np
import pytest

from oracle import statistics_oracle as so

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def fst():
    from facenet_b200 import statistics
    return statistics


@pytest.fixture(scope='module')
def handle():
    from facenet_b200 import _capi
    return _capi.default_handle(0)


class Cfg:
    metric, nrof_folds, far_target = 0, 5, 1.e-2


def test_raw_capsule_from_a_gpu_framework(fst, handle):
    """torch.utils.dlpack.to_dlpack stands in for tf.experimental.dlpack.to_dlpack (no TensorFlow in the image): both return a
    bare "dltensor" capsule.  It is taken over per protocol (renamed, deleter called on release), used in place on the GPU, and
    gives the reports of the same embeddings handed over as a tensor / as a host array."""
    import torch
    from torch.utils import dlpack as tdl
    from facenet_b200 import _capi
    x, labels = so.synthetic_embeddings([12] * 20 + [1] * 15 + [4] * 10, dim=128, sigma=(1.0, 2.5), seed=8)
    xt = torch.from_numpy(x).cuda()
    cap = tdl.to_dlpack(xt)
    v_cap = fst.FaceToFaceValidation(cap, labels, Cfg)
    assert _capi._pyapi.PyCapsule_IsValid(cap, b'used_dltensor')
    v_t = fst.FaceToFaceValidation(xt, labels, Cfg)
    v_h = fst.FaceToFaceValidation(x, labels, Cfg)
    for a, b, c in zip(v_cap.reports, v_t.reports, v_h.reports):
        assert a.dict == b.dict == c.dict
    # the low-level entry points take capsules too
    thr = so.default_thresholds(0)
    out_c = handle.pair_histogram(tdl.to_dlpack(xt), labels, thr, 0)
    out_t = handle.pair_histogram(xt, labels, thr, 0)
    np.testing.assert_array_equal(out_c['bins'], out_t['bins'])
    # ownership: the producer's memory is released when the taken-over tensor goes away
    before = torch.cuda.memory_allocated()
    big = torch.empty((4096, 1024), device='cuda')
    t = _capi.from_dlpack(tdl.to_dlpack(big))
    del big
    assert torch.cuda.memory_allocated() >= before + 4096 * 1024 * 4          # still owned by the DLPackTensor
    t.release()
    assert torch.cuda.memory_allocated() <= before + 1024


def test_callback_with_gpu_resident_model_outputs(fst):
    """facenet/callbacks.py:21-28 with a model whose outputs live on the GPU: evaluate_embeddings concatenates on the device and
    the validation consumes the tensor in place; same reports as the reference's host route."""
    import torch
    from facenet_b200 import callbacks
    x, labels = so.synthetic_embeddings([9] * 25 + [1] * 12, dim=128, sigma=(1.0, 2.5), seed=4)
    dset = [(torch.from_numpy(x[i:i + 32]), labels[i:i + 32]) for i in range(0, x.shape[0], 32)]

    class Config:
        validate = Cfg

    cb = callbacks.ValidateCallback(lambda images: images.cuda(), dset, every_n_epochs=1, max_nrof_epochs=3, config=Config)
    cb.on_epoch_end(0)
    emb, lab = callbacks.evaluate_embeddings(lambda images: images.cuda(), dset)
    assert emb.is_cuda and emb.shape == x.shape and isinstance(lab, np.ndarray)
    host = fst.FaceToFaceValidation(x, labels, Cfg)
    for a, b in zip(cb.validation.reports, host.reports):
        assert a.dict == b.dict


def test_foreign_calculator_one_launch(fst):
    x, labels = so.synthetic_embeddings([5, 1, 9, 2, 14, 30], dim=128, sigma=1.2, seed=3)
    thr = so.default_thresholds(0)

    class ForeignCalculator:
        def __init__(self):
            self.metric = 0
            self.embeddings = [x[labels == v] for v in np.unique(labels)]
            self.nrof_classes = len(self.embeddings)

    own = fst.ConfidenceMatrix(fst.SimilarityCalculator(x, labels, 0), thr)
    foreign = fst.ConfidenceMatrix(ForeignCalculator(), thr)
    for name in ('tp', 'tn', 'fp', 'fn'):
        np.testing.assert_array_equal(getattr(own, name), getattr(foreign, name))
    assert foreign.stats['kernel_launches'] >= 1


@pytest.mark.parametrize('d', [1, 17, 100, 130, 500])
def test_any_embedding_dimension(fst, d):
    """The reference accepts any D (statistics.py:33); the tensor-core tiles want multiples of 64: the drop-in functions pad the
    columns with zeros (dot products unchanged)."""
    rng = np.random.default_rng(d)
    xa = rng.standard_normal((70, d)).astype(np.float32); xa /= np.linalg.norm(xa, axis=1, keepdims=True)
    xb = rng.standard_normal((33, d)).astype(np.float32); xb /= np.linalg.norm(xb, axis=1, keepdims=True)
    for metric in (0, 1):
        np.testing.assert_allclose(fst.pairwise_similarities(xa, xb, metric), so.pairwise_similarities(xa.copy(), xb.copy(), metric), atol=1e-5 if metric == 0 or d > 1 else 4e-3)
        np.testing.assert_allclose(fst.pairwise_similarities(xa, metric=metric), so.pairwise_similarities(xa.copy(), None, metric), atol=1e-5 if metric == 0 or d > 1 else 4e-3)
    labels = np.arange(70) % 9
    cm = fst.ConfidenceMatrix(fst.SimilarityCalculator(xa, labels, 0), so.default_thresholds(0))
    ref = so.confidence_matrix_weighted(xa, labels, so.default_thresholds(0), 0)
    np.testing.assert_allclose(cm.tp, ref.tp, atol=2e-3)
    np.testing.assert_allclose(cm.fp, ref.fp, atol=2e-3)


def test_pageable_and_pinned_host_inputs_agree(handle):
    """kDLCPU tensors: pageable memory goes through the pinned ring + copy threads (csrc/fnb_stage.cu), pinned memory is copied
    in place; both give the bins of the device-resident tensor, and the stats report the copy."""
    import torch
    n, d = 40000, 512                                   # 82 MB: several ring slots
    x, labels = so.synthetic_embeddings([40] * (n // 40), dim=d, sigma=1.1, seed=1)
    thr = so.default_thresholds(0)
    dev, st_d = handle.pair_histogram_bins(torch.from_numpy(x).cuda(), torch.from_numpy(labels).cuda(), thr, 0, mode='auto')
    page, st_p = handle.pair_histogram_bins(x, labels, thr, 0, mode='auto')
    xp = torch.from_numpy(x).pin_memory()
    pin, st_q = handle.pair_histogram_bins(xp.numpy(), labels, thr, 0, mode='auto')
    np.testing.assert_array_equal(page, dev)
    np.testing.assert_array_equal(pin, dev)
    assert st_d['h2d_bytes'] == 0 and st_p['h2d_bytes'] == st_q['h2d_bytes'] == n * d * 4 + n * 8
    assert st_p['h2d_ms'] > 0 and st_q['h2d_ms'] > 0
    # a second call re-uses the ring; an odd size exercises the tail chunk
    page2, _ = handle.pair_histogram_bins(x[:33333], labels[:33333], thr, 0, mode='auto')
    dev2, _ = handle.pair_histogram_bins(torch.from_numpy(x[:33333]).cuda(), torch.from_numpy(labels[:33333]).cuda(), thr, 0, mode='auto')
    np.testing.assert_array_equal(page2, dev2)


@pytest.mark.parametrize('metric,subtract_mean', [(0, False), (1, False), (0, True)])
def test_false_examples_vs_oracle(fst, metric, subtract_mean):
    """FalseExamples (the reference's commented-out class, statistics.py:334-387) through the filter epilogue: the same pairs in
    the same order as the literal NumPy statement of its loops, distances within 1e-5; candidates within 1e-5 of the threshold
    (or near-ties of the greedy pick) may differ."""
    x, labels = so.synthetic_embeddings([6, 1, 9, 3, 12, 1, 30, 2], dim=128, sigma=(1.0, 2.5), seed=7)
    thr = 1.7 if metric == 0 else 1.45
    ref = so.false_examples(x, labels, thr, metric=metric, subtract_mean=subtract_mean)
    got = fst.FalseExamples(x, labels, thr, metric=metric, subtract_mean=subtract_mean).false_pairs()
    for key in ('fneg', 'fpos'):
        assert len(ref[key]) > 0
        r_pairs = [(a, b) for _, a, b in ref[key]]
        g_pairs = [(a, b) for _, a, b in got[key]]
        if r_pairs != g_pairs:
            # only pairs at the edge of the threshold (or tied picks) may differ
            only = set(r_pairs) ^ set(g_pairs)
            d_of = {(a, b): d for d, a, b in ref[key] + got[key]}
            assert all(abs(d_of[p] - thr) <= 2e-5 for p in only), (key, only)
        for (dr, a, b), (dg, a2, b2) in zip(ref[key], got[key]):
            if (a, b) == (a2, b2):
                assert abs(dr - dg) <= 1e-5
    ex = fst.FalseExamples(x, labels, thr, metric=metric, subtract_mean=subtract_mean)
    assert ex.false_pairs(nrof_fpos_images=1, nrof_fneg_images=1)['fneg'] == [p for i, p in enumerate(got['fneg'])
                                                                               if i == 0 or labels[p[1]] != labels[got['fneg'][i - 1][1]]]


def test_false_examples_listing_and_capacity(fst, handle, tmp_path):
    x, labels = so.synthetic_embeddings([8] * 10 + [1] * 5, dim=64, sigma=(1.0, 2.5), seed=9)
    files = ['/data/id%03d/img%04d.png' % (l, i) for i, l in enumerate(labels)]
    ex = fst.FalseExamples(x, labels, 1.75, files=files)
    pairs = ex.write_false_pairs(tmp_path / 'fpos', tmp_path / 'fneg')
    lines = (tmp_path / 'fneg' / 'false_pairs.txt').read_text().splitlines()
    assert len(lines) == len(pairs['fneg']) > 0
    d, a, b = pairs['fneg'][0]
    assert lines[0].startswith(str(tmp_path / 'fneg' / ('%2.3f & id%03d|img%04d & id%03d|img%04d.png' % (d, labels[a], a, labels[b], b))))
    assert '{:2.3f}/{:2.3f}'.format(d, 1.75) in lines[0]
    # the candidate list outgrows a small buffer: the binding repeats the call with the reported size
    r1, c1, d1, _ = handle.false_pairs(x, labels, 1.75, capacity=4)
    r2, c2, d2, _ = handle.false_pairs(x, labels, 1.75)
    assert r1.size == r2.size > 4
    assert sorted(zip(r1.tolist(), c1.tolist())) == sorted(zip(r2.tolist(), c2.tolist()))
    with pytest.raises(ValueError, match='normalized'):
        fst.FalseExamples(x * 1.5, labels, 1.75).false_pairs()


def test_sharded_entry_point_on_one_rank():
    """fnb_comm_init / fnb_pair_histogram_sharded with a communicator of ONE rank (NCCL resolved at run time, behind the C ABI):
    the collective entry point equals the plain one, for host and device rows in any order, and a second handle on the same
    device is unaffected.  (The N > 1 forms are checked by scripts/check_multi_gpu.py under torchrun: profiles/r02n_check_n*.log.)"""
    import torch
    from facenet_b200 import _capi
    h = _capi.Handle(0)
    try:
        uid = h.comm_unique_id()
        assert len(uid) == 128
        h.comm_init(uid, 0, 1)
        info = h.comm_info()
        assert info['rank'] == 0 and info['world'] == 1 and info['nccl_version'] >= 22000
        x, labels = so.synthetic_embeddings([33] * 70 + [1] * 41, dim=512, sigma=0.8, seed=23)      # shuffled rows
        thr = so.default_thresholds(0)
        for mode in ('fp16x3', 'auto'):
            ref, st0 = h.pair_histogram_bins(x, labels, thr, 0, mode=mode)
            got, st = h.pair_histogram_sharded(x, labels, thr, 0, mode=mode)
            np.testing.assert_array_equal(got, ref)
            assert st['n_pairs'] == st0['n_pairs'] and st['mode_used'] == st0['mode_used']
            got, _ = h.pair_histogram_sharded(torch.from_numpy(x).cuda(), torch.from_numpy(labels).cuda(), thr, 0, mode=mode)
            np.testing.assert_array_equal(got, ref)
        bad = x.copy(); bad[7] = bad[8] * 1.01
        with pytest.raises(_capi.FnbError, match='normalized'):
            h.pair_histogram_sharded(bad, labels, thr, 0)
        h.comm_destroy()
        with pytest.raises(_capi.FnbError, match='communicator'):
            h.pair_histogram_sharded(x, labels, thr, 0)
    finally:
        h.close()
