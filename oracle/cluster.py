"""Oracle O5: NumPy restatement of ``AdvancedKeyframeExtractor.cluster_similar_frames``
(filter_research_update.py:113-134) -- sklearn cosine matrix, ``1 - sim`` distances, DBSCAN(eps,
min_samples, metric='precomputed') -- and of the DBSCAN labelling it delegates to.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

DBSCAN lives in scikit-learn (third party; the reference imports ``sklearn.cluster.DBSCAN``,
filter_research_update.py:11).  Its published algorithm as restated here (sklearn/cluster/_dbscan.py and
_dbscan_inner.pyx): neighbourhood of i = {j : dist[i, j] <= eps} (ascending j, i itself included); i is a core
sample iff its neighbourhood has >= min_samples members; points are scanned in index order, every unlabelled
core sample starts a new cluster which is grown depth-first (LIFO stack) through core samples, border points
take the label of the cluster that reaches them first; the rest is noise (-1).

PINNED: reproduces the outputs of the reference's own method (with real scikit-learn) on
tests/golden/research.* (tests/golden/make_golden_research.py) -- see tests/test_oracle_golden.py.
"""
from __future__ import annotations

import numpy as np

from .dedup import _normalize_rows


def cosine_matrix(embeddings) -> np.ndarray:
    """sklearn ``cosine_similarity(X)``: normalise rows (zero norm -> 1), then X @ X.T, in X's float dtype."""
    x = np.asarray(embeddings)
    if x.dtype not in (np.float32, np.float64):
        x = x.astype(np.float64)
    xn = _normalize_rows(x)
    return xn @ xn.T


def dbscan_labels(neighborhoods, min_samples: int) -> np.ndarray:
    """Labels of sklearn's ``dbscan_inner`` given the eps-neighbourhoods (lists of ascending indices)."""
    n = len(neighborhoods)
    labels = np.full(n, -1, dtype=np.int64)
    is_core = np.array([len(nb) >= min_samples for nb in neighborhoods], dtype=bool)
    label_num = 0
    stack = []
    for i in range(n):
        if labels[i] != -1 or not is_core[i]:
            continue
        while True:
            if labels[i] == -1:
                labels[i] = label_num
                if is_core[i]:
                    for v in neighborhoods[i]:
                        if labels[v] == -1:
                            stack.append(int(v))
            if not stack:
                break
            i = stack.pop()
        label_num += 1
    return labels


def groups_from_labels(labels) -> list:
    """filter_research_update.py:129-134: ``defaultdict(list)`` keyed by label in first-seen order -> list of
    index lists (the noise label -1 forms one group like any other)."""
    clusters = {}
    for idx, label in enumerate(labels):
        clusters.setdefault(int(label), []).append(idx)
    return list(clusters.values())


def cluster_similar_frames(embeddings, eps: float = 0.05, min_samples: int = 2) -> list:
    """filter_research_update.py:113-134."""
    if len(embeddings) < 2:
        return [[0]] if len(embeddings) else []
    sim = cosine_matrix(np.stack([np.asarray(e) for e in embeddings]))
    dist = 1 - sim
    radius = dist.dtype.type(eps)                       # NumPy weak-scalar promotion: compared in dist's dtype
    neighborhoods = [np.nonzero(row <= radius)[0] for row in dist]
    return groups_from_labels(dbscan_labels(neighborhoods, min_samples))


def select_representative_frame(cluster_indices, embeddings):
    """filter_research_update.py:136-155 -- the member closest (cosine) to the cluster's mean embedding; first
    maximum wins; a single-member cluster returns that member."""
    from .dedup import cos1
    if len(cluster_indices) == 1:
        return cluster_indices[0]
    centroid = np.mean([embeddings[i] for i in cluster_indices], axis=0)
    best_idx, best_sim = 0, -1
    for i, emb_idx in enumerate(cluster_indices):
        sim = cos1(embeddings[emb_idx], centroid)
        if sim > best_sim:
            best_sim, best_idx = sim, i
    return cluster_indices[best_idx]
