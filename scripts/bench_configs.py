"""Timings + parity of the BASELINE.json configs that are not the bench.py headline (run on a B200):

  c1  synthetic LFW-size 13,233 x 512 (5,749 identities): FaceToFaceValidation (10 folds, 100 thresholds,
      FAR 1e-3) through the drop-in Python API, host arrays in, Report dict out; parity vs the vectorised
      oracle (chosen thresholds, fold rates) and the oracle's CPU time beside it
  c3  triplet mining, 1800-image batches (45 identities x 40 images), 512-d, alpha 0.2: steps/s with
      host batches (H2D inside) and with device-resident batches; parity of one batch vs the mining oracle

  f1  ConfusionMatrix of the pair classifiers on the 26,489-row set of the published runs (section 8 f1)
  f2  cross-entropy loss + gradients of one 1800-row P x K batch (section 8 f2)

    python scripts/bench_configs.py [c1] [c3] [f1] [f2] [--steps N]   -> one JSON line per config
"""
import json
import sys
import time

import numpy as np

sys.path.insert(0, '.')


class Cfg:
    def __init__(self, metric=0, nrof_folds=10, far_target=1.e-3):
        self.metric, self.nrof_folds, self.far_target = metric, nrof_folds, far_target


def run_c1(mode):
    from facenet_b200 import statistics as fst
    from oracle import statistics_oracle as so
    fst.set_default_mode(mode=mode)
    sizes = so.lfw_like_class_sizes()
    x, labels = so.synthetic_embeddings(sizes, dim=512, sigma=1.1, seed=0)
    cfg = Cfg()
    fst.FaceToFaceValidation(x[:2000], labels[:2000], cfg)          # warm-up (context, workspaces)
    times = []
    for _ in range(3):
        t0 = time.perf_counter()
        v = fst.FaceToFaceValidation(x, labels, cfg)
        times.append(time.perf_counter() - t0)
    # the same validation with the embeddings already resident on the GPU (section 8 f3: no D2H -> H2D round trip)
    import torch
    xd, ld = torch.from_numpy(x).cuda(), torch.from_numpy(labels).cuda()
    torch.cuda.synchronize()
    times_dev = []
    for _ in range(3):
        t0 = time.perf_counter()
        vd = fst.FaceToFaceValidation(xd, ld, cfg)
        times_dev.append(time.perf_counter() - t0)
    same_dict = all(float(vd.dict[c][k]) == float(v.dict[c][k]) for c in v.dict for k in v.dict[c])
    got = v.dict
    thr_acc = np.array([m.threshold[0] for m in v.reports[0].conf_matrix_test])
    thr_far = np.array([m.threshold[0] for m in v.reports[1].conf_matrix_test])
    t0 = time.perf_counter()
    ref = so.face_to_face_validation(x, labels, 0, 10, 1.e-3)
    cpu_s = time.perf_counter() - t0
    n = x.shape[0]
    ntr = n - n // 10
    pair_evals = 10 * (ntr * (ntr - 1) // 2) + 20 * ((n // 10) * (n // 10 - 1) // 2)
    worst = 0.0
    for crit in got:
        for k in got[crit]:
            worst = max(worst, abs(float(got[crit][k]) - float(ref[crit][k])))
    line = {'config': 'c1: synthetic LFW-size 13,233 x 512 fp32 (5,749 identities), FaceToFaceValidation 10 folds x 100 thresholds',
            'mode': mode, 'seconds': min(times), 'seconds_all': times, 'seconds_gpu_resident_embeddings': min(times_dev),
            'gpu_resident_reports_identical': bool(same_dict), 'pair_distance_evaluations': pair_evals,
            'g_pair_distances_per_s': pair_evals / min(times) / 1e9,
            'cpu_vectorised_oracle_seconds': cpu_s, 'cpu_literal_reference_seconds_extrapolated': 4.0e-3 * 5749 ** 2,
            'accuracy_threshold_equal_folds': int(np.sum(thr_acc == ref['_thresholds'][:, 0])),
            'far_threshold_max_abs_diff': float(np.abs(thr_far - ref['_thresholds'][:, 1]).max()),
            'report_dict_max_abs_diff': worst,
            'report': {k: {kk: float(vv) for kk, vv in d.items()} for k, d in got.items()}}
    print(json.dumps(line))


def run_c3(mode, steps):
    import torch
    from facenet_b200 import _capi
    from oracle import mining_oracle as mo
    from oracle import statistics_oracle as so
    h = _capi.default_handle(0)
    nb = 8
    batches = []
    for s in range(nb):
        x, labels = so.synthetic_embeddings([40] * 45, dim=512, sigma=1.1, seed=100 + s, shuffle=False)
        batches.append((x, labels))
    out = h.mine(batches[0][0], batches[0][1], alpha=0.2, mode=mode, kmax=39)
    ref = mo.mine(batches[0][0], batches[0][1], alpha=0.2)
    agree = {k: float(np.mean(out[k] == ref[k])) for k in ref}
    # host batches: H2D of the batch and D2H of the index arrays inside every step
    for i in range(5):
        h.mine(*batches[i % nb], alpha=0.2, mode=mode, kmax=39)
    t0 = time.perf_counter()
    for i in range(steps):
        h.mine(*batches[i % nb], alpha=0.2, mode=mode, kmax=39)
    host_s = (time.perf_counter() - t0) / steps
    dev = [(torch.from_numpy(x).cuda(), torch.from_numpy(l).cuda()) for x, l in batches]
    torch.cuda.synchronize()
    for i in range(5):
        h.mine(*dev[i % nb], alpha=0.2, mode=mode, kmax=39)
    t0 = time.perf_counter()
    for i in range(steps):
        h.mine(*dev[i % nb], alpha=0.2, mode=mode, kmax=39)
    dev_s = (time.perf_counter() - t0) / steps
    t0 = time.perf_counter()
    mo.mine(batches[1][0], batches[1][1], alpha=0.2)
    cpu_s = time.perf_counter() - t0
    b = 1800
    line = {'config': 'c3: triplet mining, 1800-image batches (45 identities x 40 images), 512-d, alpha 0.2',
            'mode': mode, 'steps': steps, 'ms_per_step_host_batches': host_s * 1e3, 'ms_per_step_device_batches': dev_s * 1e3,
            'steps_per_s_device_batches': 1.0 / dev_s, 'seconds_for_10k_steps_device_batches': 1e4 * dev_s,
            'g_ordered_pair_distances_per_s': b * b / dev_s / 1e9, 'cpu_oracle_ms_per_step': cpu_s * 1e3,
            'index_agreement_with_oracle': agree}
    print(json.dumps(line))


def run_f1():
    """ConfusionMatrix (train_classifier.py:17-49) at the size of the published validation runs: 26,489 x 512, 530 classes of
    45-50 images (models/20200724-231357/logs/report.txt:13-22); both classifiers; CPU oracle (vectorised) beside it and the
    literal per-class-pair loop on a class subsample."""
    from facenet_b200 import faceclass
    from facenet_b200.apps import train_classifier as tc
    from oracle import faceclass_oracle as fo
    from oracle import statistics_oracle as so
    rng = np.random.default_rng(0)
    sizes = rng.integers(45, 51, size=530).tolist()
    x, _ = so.synthetic_embeddings(sizes, dim=512, sigma=1.1, seed=2, shuffle=False)
    xu = (x * rng.uniform(0.6, 1.7, size=(x.shape[0], 1))).astype(np.float32)
    b = np.concatenate([[0], np.cumsum(sizes)])
    cls = np.repeat(np.arange(len(sizes)), sizes)
    for tag, data, model, fn in (('normalized', x, faceclass.FaceToFaceNormalizedEmbeddingsClassifier(), fo.distance_normalized),
                                 ('distance', xu, faceclass.FaceToFaceDistanceClassifier(), None)):
        if fn is None:
            model.variables['theta'] = np.float32(0.7)
            fn = lambda a, c: fo.distance_unnormalized(a, c, 0.7)
        model.variables['threshold'] = np.float32(1.1)
        emb = [data[a:c] for a, c in zip(b[:-1], b[1:])]
        tc.ConfusionMatrix(emb[:40], model)                 # warm-up
        times = []
        for _ in range(3):
            t0 = time.perf_counter()
            cm = tc.ConfusionMatrix(emb, model)
            times.append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        ref = fo.confusion_matrix_vectorized(data, cls, fn, 1.1)
        cpu_vec = time.perf_counter() - t0
        nsub = 60
        t0 = time.perf_counter()
        lit = fo.confusion_matrix(emb[:nsub], fn, 1.1)
        cpu_lit = time.perf_counter() - t0
        sub = tc.ConfusionMatrix(emb[:nsub], model)
        n = data.shape[0]
        line = {'config': 'f1: ConfusionMatrix, %s classifier, 26,489-row-class set (%d x 512, 530 classes of 45-50), threshold 1.1' % (tag, n),
                'seconds': min(times), 'seconds_all': times, 'unordered_pairs': n * (n - 1) // 2,
                'g_pair_distances_per_s': n * (n - 1) / 2 / min(times) / 1e9, 'kernel_ms': cm.stats['kernel_ms'],
                'cpu_vectorised_oracle_seconds': cpu_vec,
                'cpu_literal_loop_seconds_%d_classes' % nsub: cpu_lit,
                'cpu_literal_loop_seconds_extrapolated_530_classes': cpu_lit * (530.0 / nsub) ** 2,
                'rates': [cm.accuracy, cm.precision, cm.tp_rate, cm.tn_rate],
                'max_abs_diff_vs_vectorised_oracle': float(np.max(np.abs(np.array([cm.accuracy, cm.precision, cm.tp_rate, cm.tn_rate]) -
                                                                   np.array([ref.accuracy, ref.precision, ref.tp_rate, ref.tn_rate])))),
                'max_abs_diff_vs_literal_loop_%d_classes' % nsub: float(np.max(np.abs(
                    np.array([sub.accuracy, sub.precision, sub.tp_rate, sub.tn_rate]) - np.array([lit.accuracy, lit.precision, lit.tp_rate, lit.tn_rate])))),
                'eps_window_pairs': int(cm.stats['eps_window'])}
        print(json.dumps(line))


def run_f2(steps):
    """One optimiser step's loss + gradients of the pair classifier on an 1800-row P x K batch (45 x 40, 512-d)."""
    import torch
    from facenet_b200 import faceclass
    from facenet_b200.apps import train_classifier as tc
    from oracle import faceclass_oracle as fo
    from oracle import statistics_oracle as so

    class Opt:
        nrof_classes_per_batch, nrof_examples_per_class = 45, 40
    x, _ = so.synthetic_embeddings([40] * 45, dim=512, sigma=1.1, seed=9, shuffle=False)
    xu = (x * np.random.default_rng(3).uniform(0.6, 1.7, size=(x.shape[0], 1))).astype(np.float32)
    for tag, data, model in (('normalized', x, faceclass.FaceToFaceNormalizedEmbeddingsClassifier()),
                             ('distance', xu, faceclass.FaceToFaceDistanceClassifier())):
        out = tc.pair_cross_entropy(model, data, Opt)
        for _ in range(5):
            tc.pair_cross_entropy(model, data, Opt)
        t0 = time.perf_counter()
        for _ in range(steps):
            tc.pair_cross_entropy(model, data, Opt)
        host_s = (time.perf_counter() - t0) / steps
        dev = torch.from_numpy(data).cuda()
        torch.cuda.synchronize()
        for _ in range(5):
            tc.pair_cross_entropy(model, dev, Opt)
        t0 = time.perf_counter()
        for _ in range(steps):
            tc.pair_cross_entropy(model, dev, Opt)
        dev_s = (time.perf_counter() - t0) / steps
        t0 = time.perf_counter()
        if tag == 'normalized':
            dist = fo.distance_normalized(data)
        else:
            dist = fo.distance_unnormalized(data, None, 1.0)
        ref32 = float(fo.binary_cross_entropy_loss(fo.logits(dist, 10, 1), 45, 40))
        cpu_s = time.perf_counter() - t0
        ref = fo.binary_cross_entropy_loss_and_grads(dist, 10.0, 1.0, 45, 40)
        line = {'config': 'f2: pair-classifier weighted cross entropy + gradients, %s classifier, 1800-row batch (45 x 40), 512-d' % tag,
                'steps': steps, 'ms_per_step_host_batch': host_s * 1e3, 'ms_per_step_device_batch': dev_s * 1e3,
                'kernel_ms': out['stats']['kernel_ms'], 'cpu_oracle_ms_per_step_loss_only': cpu_s * 1e3,
                'loss': out['loss'], 'loss_oracle_f32': ref32, 'loss_oracle_f64': ref['loss'],
                'rel_diff_vs_f64': abs(out['loss'] - ref['loss']) / max(1.0, abs(ref['loss'])),
                'dalpha': [out['grads']['alpha'], ref['dalpha']], 'dthreshold': [out['grads']['threshold'], ref['dthreshold']]}
        print(json.dumps(line))


if __name__ == '__main__':
    args = [a for a in sys.argv[1:] if not a.startswith('--')]
    steps = 200
    mode = 'fp16x3'
    for i, a in enumerate(sys.argv):
        if a == '--steps':
            steps = int(sys.argv[i + 1])
        if a == '--mode':
            mode = sys.argv[i + 1]
    args = [a for a in args if a not in (str(steps), mode)]
    if not args or 'c1' in args:
        run_c1(mode)
    if not args or 'c3' in args:
        run_c3(mode, steps)
    if 'f1' in args:
        run_f1()
    if 'f2' in args:
        run_f2(steps)
