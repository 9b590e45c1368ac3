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
