"""Triplet selection for P x K batches, computed on a B200 (BASELINE.json north_star: "per-anchor hard
and semi-hard negative selection used for triplet batches in facenet/facenet.py").

The sMedX fork contains NO triplet code; what it fixes is the batch layout
(``equal_batches_input_pipeline``, /root/reference/facenet/facenet.py:89-123: ``nrof_classes_per_batch``
x ``nrof_examples_per_class`` rows grouped by class), the same-identity rule
(facenet/apps/train_classifier.py:62-73) and the distance (facenet/statistics.py:33-50).  The functions
here therefore follow the semantics written down in ``oracle/mining_oracle.py``; ``select_triplets``
keeps the call shape of upstream davidsandberg/facenet's function of that name (not in the fork).

All distances, comparisons and arg-reductions run in the CUDA library (``fnb_mine``); there is no CPU
fallback.
"""
import numpy as np

from facenet_b200 import _capi
from facenet_b200.statistics import _handle, _raise_like_reference, _state

__all__ = ['mine', 'mine_batched', 'mine_check', 'select_kth_eligible', 'hardest_triplets', 'semi_hard_triplets', 'select_triplets']


def mine(embeddings, labels, alpha=0.2, mode=None):
    """Per-anchor mining on one batch.

    Returns a dict of int32 arrays (``-1`` = empty set; ties resolved to the lowest index):

    ``hardest_pos [B]``      argmax over p != a, label[p] == label[a] of d(a, p)
    ``hardest_neg [B]``      argmin over n, label[n] != label[a] of d(a, n)
    ``pos_index [B, K-1]``   the positives of every anchor in ascending index order
    ``semi_hard [B, K-1]``   per (a, p): argmin of d(a, n) over negatives with d(a,n) > d(a,p) and
                             fp32(d(a,n) - d(a,p)) < alpha
    ``eligible [B, K-1]``    per (a, p): #{negatives n : fp32(d(a,n) - d(a,p)) < alpha}
    with d = 2 * (1 - clamp(x_a . x_b)) (metric 0 of ``pairwise_similarities``)."""
    try:
        return _handle().mine(embeddings, labels, alpha=alpha, mode=mode or _state['mode'])
    except _capi.FnbError as err:
        _raise_like_reference(err, 0)


def mine_batched(embeddings, labels, nbatches=1, alpha=0.2, kmax=None, mode=None, out=None):
    """``nbatches`` P x K batches packed as ``[S * B, D]`` / ``[S * B]`` mined in ONE call (indices local to each batch).

    torch CUDA tensors in -> torch CUDA int32 tensors out and NO host synchronisation (a training loop keeps the mined
    indices on the GPU; pass ``kmax`` = K - 1 and re-use ``out``); data-dependent errors (un-normalised embeddings, ``kmax``
    too small) are raised by ``mine_check()``.  NumPy in -> NumPy out, synchronous.  ``kmax=0`` mines the hardest positive /
    negative only, fully fused in the Gram epilogue."""
    try:
        return _handle().mine_batched(embeddings, labels, nbatches=nbatches, alpha=alpha, kmax=kmax, mode=mode or _state['mode'], out=out)
    except _capi.FnbError as err:
        _raise_like_reference(err, 0)


def mine_check():
    """Synchronise with the last device-resident ``mine_batched`` call and raise its data-dependent errors."""
    try:
        return _handle().mine_check()
    except _capi.FnbError as err:
        _raise_like_reference(err, 0)


def select_kth_eligible(anchors, positives, kth, alpha=0.2):
    """For every query ``(a, p, k)``: the ``k``-th (0-based, ascending index) negative ``n`` of anchor ``a`` with
    ``fp32(d(a, n) - d(a, p)) < alpha`` in the batch mined last (``a`` = row of that call, ``p`` / result local to a's batch);
    -1 when fewer exist.  With ``k = randint(eligible[a, p])`` this is the draw of upstream ``select_triplets``."""
    return _handle().mine_select_kth(anchors, positives, kth, alpha=alpha)


def hardest_triplets(embeddings, labels, mode=None):
    """(anchor, hardest positive, hardest negative) for every anchor that has both; int32 [M, 3]."""
    out = mine(embeddings, labels, alpha=0.0, mode=mode)
    a = np.arange(out['hardest_pos'].size, dtype=np.int32)
    ok = (out['hardest_pos'] >= 0) & (out['hardest_neg'] >= 0)
    return np.stack([a[ok], out['hardest_pos'][ok], out['hardest_neg'][ok]], axis=1)


def semi_hard_triplets(embeddings, labels, alpha=0.2, mode=None, unique_pairs=True):
    """(anchor, positive, semi-hard negative) for every (a, p) that has one; ``unique_pairs`` keeps
    p > a only (upstream ``select_triplets`` visits each unordered positive pair once).  int32 [M, 3]."""
    out = mine(embeddings, labels, alpha=alpha, mode=mode)
    pos, neg = out['pos_index'], out['semi_hard']
    a = np.broadcast_to(np.arange(pos.shape[0], dtype=np.int32)[:, None], pos.shape)
    ok = (pos >= 0) & (neg >= 0)
    if unique_pairs:
        ok &= pos > a
    return np.stack([a[ok], pos[ok], neg[ok]], axis=1)


def select_triplets(embeddings, nrof_images_per_class, image_paths=None, people_per_batch=None, alpha=0.2, mode=None, rng=None):
    """Call shape of upstream ``select_triplets(embeddings, nrof_images_per_class, image_paths,
    people_per_batch, alpha)``: rows are grouped by class with the given class sizes.

    ``rng=None`` (deterministic variant): the negative of every (a, p), p after a, is the semi-hard argmin; returns
    ``(triplets, nrof_eligible)``.  ``rng`` = a ``np.random.RandomState`` (or the ``np.random`` module): upstream's selection
    is replayed -- for every (a, p) in upstream's loop order whose eligible set ``{n : d(a,n) - d(a,p) < alpha}`` is not
    empty, ``rng.randint(len(eligible))`` picks the entry of the ascending candidate list (``fnb_mine_select_kth``), then the
    triplets are shuffled with ``rng.shuffle``; returns upstream's ``(triplets, num_trips, len(triplets))``.

    triplets is int32 [M, 3] of row indices, or a list of ``(path_a, path_p, path_n)`` when ``image_paths`` is given."""
    sizes = np.asarray(nrof_images_per_class, dtype=np.int64)
    if people_per_batch is not None:
        sizes = sizes[:people_per_batch]
    labels = np.repeat(np.arange(sizes.size, dtype=np.int64), sizes)
    embeddings = np.asarray(embeddings)[:labels.size]
    out = mine(embeddings, labels, alpha=alpha, mode=mode)
    pos, neg = out['pos_index'], out['semi_hard']
    a = np.broadcast_to(np.arange(pos.shape[0], dtype=np.int32)[:, None], pos.shape)
    if rng is None:
        ok = (pos > a) & (neg >= 0)
        trip = np.stack([a[ok], pos[ok], neg[ok]], axis=1)
        elig = out['eligible'][pos > a]
        if image_paths is not None:
            trip = [(image_paths[i], image_paths[j], image_paths[k]) for i, j, k in trip]
        return trip, elig
    # upstream order: anchors ascending, for each the positives after it ascending (row-major over [B, K-1] is that order)
    visit = pos > a
    num_trips = int(np.count_nonzero(visit))
    take = visit & (out['eligible'] > 0)
    qa, qp, qe = a[take], pos[take], out['eligible'][take]
    kth = np.array([rng.randint(int(e)) for e in qe], dtype=np.int32)
    qn = select_kth_eligible(qa.astype(np.int32), qp.astype(np.int32), kth, alpha=alpha) if qa.size else np.zeros(0, dtype=np.int32)
    trip = [(int(i), int(j), int(k)) for i, j, k in zip(qa, qp, qn)]
    if image_paths is not None:
        trip = [(image_paths[i], image_paths[j], image_paths[k]) for i, j, k in trip]
    rng.shuffle(trip)
    if image_paths is None:
        trip = np.asarray(trip, dtype=np.int32).reshape(-1, 3)
    return trip, num_trips, len(trip)
