"""NumPy statement of the triplet-mining semantics DEFINED BY THIS REPO.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  PARITY UNPINNED: the
reference fork contains no triplet-mining code (SURVEY.md section 0 R1, section 8
row A8), so there is nothing to import, restate or take golden vectors from.
The nearest reference artefacts fix only the conventions used here:

* batch layout: P classes x K examples, rows grouped by class
  (``facenet/facenet.py:89-123``), same-identity rule ``label[a] == label[b]``
  (``facenet/apps/train_classifier.py:62-73`` uses ``i // K == k // K``);
* distance: metric 0 of ``pairwise_similarities`` -- ``2 * (1 - clamp(x @ y))`` in
  fp32 (``facenet/statistics.py:33-50``);
* hardest pairs: within-class ``argmax`` / cross-class ``argmin`` of the distance
  (commented-out search in ``facenet/statistics.py:357-387``).

Definitions (ties always resolved to the LOWEST index; ``-1`` = empty set):

``hardest_pos[a]``      argmax over p != a, label[p] == label[a] of d(a, p)
``hardest_neg[a]``      argmin over n, label[n] != label[a] of d(a, n)
``pos_index[a, j]``     j-th index p (ascending, p != a) with label[p] == label[a]
``semi_hard[a, j]``     for p = pos_index[a, j]: argmin of d(a, n) over n with
                        label[n] != label[a], d(a, n) > d(a, p) and
                        fp32(d(a, n) - d(a, p)) < fp32(alpha)
``eligible[a, j]``      #{n : label[n] != label[a], fp32(d(a,n) - d(a,p)) < fp32(alpha)} --
                        the candidate-set size of upstream davidsandberg/facenet's
                        random ``select_triplets`` (not in the reference fork)
"""
import numpy as np


def distance_matrix(embeddings):
    """fp32 squared-L2 distance of unit vectors, facenet/statistics.py:33,45-50."""
    x = np.ascontiguousarray(embeddings, dtype=np.float32)
    s = x @ x.T
    np.clip(s, -1, 1, out=s)
    return 2 * (1 - s)


def mine(embeddings, labels, alpha=0.2, dist=None):
    labels = np.asarray(labels)
    d = distance_matrix(embeddings) if dist is None else np.asarray(dist, dtype=np.float32)
    b = d.shape[0]
    alpha32 = np.float32(alpha)
    same = labels[:, None] == labels[None, :]
    not_self = ~np.eye(b, dtype=bool)
    pos_mask = same & not_self
    neg_mask = ~same

    kmax = int(pos_mask.sum(axis=1).max()) if b else 0
    hardest_pos = np.full(b, -1, dtype=np.int32)
    hardest_neg = np.full(b, -1, dtype=np.int32)
    pos_index = np.full((b, kmax), -1, dtype=np.int32)
    semi_hard = np.full((b, kmax), -1, dtype=np.int32)
    eligible = np.zeros((b, kmax), dtype=np.int32)

    for a in range(b):
        pos = np.nonzero(pos_mask[a])[0]
        neg = np.nonzero(neg_mask[a])[0]
        if pos.size:
            hardest_pos[a] = pos[np.argmax(d[a, pos])]          # first max == lowest index
            pos_index[a, :pos.size] = pos
        if neg.size:
            hardest_neg[a] = neg[np.argmin(d[a, neg])]
        if pos.size and neg.size:
            dn = d[a, neg]
            for j, p in enumerate(pos):
                dp = d[a, p]
                margin_ok = (dn - dp) < alpha32                 # fp32 subtraction
                eligible[a, j] = int(np.count_nonzero(margin_ok))
                cand = margin_ok & (dn > dp)
                if cand.any():
                    idx = np.nonzero(cand)[0]
                    semi_hard[a, j] = neg[idx[np.argmin(dn[idx])]]
    return {'hardest_pos': hardest_pos, 'hardest_neg': hardest_neg,
            'pos_index': pos_index, 'semi_hard': semi_hard, 'eligible': eligible}


def select_kth_eligible(embeddings, labels, anchors, positives, kth, alpha=0.2, dist=None):
    """``kth[i]``-th entry (0-based) of ``np.where(fp32(d[a] - d[a, p]) < alpha)[0]`` restricted to the negatives of ``a``; -1
    if the list is shorter -- the candidate list upstream ``select_triplets`` indexes with ``np.random.randint``."""
    labels = np.asarray(labels)
    d = distance_matrix(embeddings) if dist is None else np.asarray(dist, dtype=np.float32)
    alpha32 = np.float32(alpha)
    out = np.full(len(anchors), -1, dtype=np.int32)
    for i, (a, p, k) in enumerate(zip(anchors, positives, kth)):
        cand = np.nonzero((labels != labels[a]) & ((d[a] - d[a, p]) < alpha32))[0]
        if 0 <= k < cand.size:
            out[i] = cand[k]
    return out


def select_triplets_upstream(embeddings, nrof_images_per_class, people_per_batch, alpha, rng, dist=None):
    """Restatement of upstream davidsandberg/facenet ``src/train_tripletloss.py:select_triplets`` (NOT in the sMedX fork --
    SURVEY.md section 8 A8; written from its published algorithm, PARITY UNPINNED): for every anchor ``a`` and every positive
    ``p`` after it in the class, a uniformly random negative among ``{n : d(a,n) - d(a,p) < alpha}`` drawn with
    ``rng.randint``; the list is shuffled at the end.  Distances are this repo's metric 0 (== upstream's squared L2 for
    unit-norm embeddings).  Returns ``(triplets int32 [M, 3], num_trips, M)``."""
    sizes = np.asarray(nrof_images_per_class, dtype=np.int64)[:people_per_batch]
    labels = np.repeat(np.arange(sizes.size), sizes)
    d = distance_matrix(np.asarray(embeddings)[:labels.size]) if dist is None else np.asarray(dist, dtype=np.float32)
    alpha32 = np.float32(alpha)
    triplets, num_trips, start = [], 0, 0
    for n_img in sizes:
        n_img = int(n_img)
        for j in range(1, n_img):
            a = start + j - 1
            neg = d[a].copy()
            neg[start:start + n_img] = np.nan
            for pair in range(j, n_img):
                p = start + pair
                with np.errstate(invalid='ignore'):
                    all_neg = np.where((neg - d[a, p]) < alpha32)[0]
                if all_neg.size > 0:
                    triplets.append((a, p, int(all_neg[rng.randint(all_neg.size)])))
                num_trips += 1
        start += n_img
    rng.shuffle(triplets)
    return np.asarray(triplets, dtype=np.int32).reshape(-1, 3), num_trips, len(triplets)
