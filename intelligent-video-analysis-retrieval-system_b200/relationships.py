"""Per-folder similarity relationships on the B200 path ("next" row 2).

Mirrors ``MetadataManager._build_similarity_relationships`` (core.py:3493-3531): for every folder
(video) the reference computes the full cosine matrix of its keyframes with sklearn and keeps, for
each frame, the top-10 other frames whose cosine exceeds 0.7.

Two stages:
  1. candidates -- the batched search kernel with Q = X: the folder's rows are L2-normalised, added
     to a flat inner-product index and searched against themselves with k = top + 1 + ``margin``
     (the fused top-k epilogue never materialises the n x n matrix);
  2. decision   -- the reference takes its ``> 0.7`` cut and its ranking on float32 cosines, while
     the index scores fp16 rows (about 1e-4 absolute error): the few candidates of every frame, plus
     the frame itself BY ID, are therefore re-scored in float32 on the host (normalise, then dot --
     sklearn's order) and ranked the way the reference ranks them: ``np.argsort(sim)[::-1]``
     (ties: higher index first), position 0 dropped as "self" whatever it is, next ``top`` kept where
     the float32 cosine exceeds the threshold.  O(n * k * d) host work per folder; folders are small.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np

from . import faiss_compat as faiss

CANDIDATE_MARGIN = 6          # extra candidates per frame so that fp16 near-ties cannot push a true top-10 frame out


def rank_candidates(xn: np.ndarray, cand: np.ndarray, top: int, threshold: float) -> List[List[int]]:
    """Stage 2 on the host.  ``xn``: float32 [n, d] L2-normalised rows; ``cand``: int64 [n, kc] candidate
    row ids per frame (-1 = padding; may or may not contain the frame itself).  Returns, per frame, the
    row indices the reference keeps (core.py:3516-3524)."""
    n = xn.shape[0]
    cand = np.asarray(cand, np.int64)
    self_col = np.arange(n, dtype=np.int64)[:, None]
    cand = np.concatenate([self_col, np.where(cand == self_col, -1, cand)], axis=1)   # self exactly once, by id
    safe = np.where(cand < 0, 0, cand)
    sims = np.empty(cand.shape, np.float32)
    for s0 in range(0, n, 1024):                                    # bounded gather: [1024, kc, d] at a time
        s1 = min(n, s0 + 1024)
        sims[s0:s1] = np.einsum("id,ikd->ik", xn[s0:s1], xn[safe[s0:s1]], dtype=np.float32, optimize=False)
    out = []
    for i in range(n):
        ok = cand[i] >= 0
        ids, s = cand[i][ok], sims[i][ok]
        # the reference's order: ascending argsort reversed -> score descending, ties by HIGHER index first
        order = np.lexsort((ids, s))[::-1]
        order = order[1:top + 1]                                    # position 0 is dropped as "self" whatever it is
        out.append([int(ids[j]) for j in order if s[j] > threshold])
    return out


def build_similarity_relationships(all_metadata: Dict[str, List], top: int = 10, threshold: float = 0.7,
                                   device: int | None = None) -> Dict[str, List[str]]:
    """{frame key: [keys of its most similar frames in the same folder]} (core.py:3493-3531).

    ``all_metadata``: {folder_name: [KeyframeMetadata-like with .clip_features, .get_unique_key()]}.
    Frames without features are skipped; folders with fewer than 2 featured frames are skipped.
    """
    graph: Dict[str, List[str]] = {}
    index = None
    try:
        for _folder, metadata_list in all_metadata.items():
            feats, keys = [], []
            for m in metadata_list:
                if m.clip_features is not None:
                    feats.append(np.asarray(m.clip_features, dtype=np.float32).reshape(-1))
                    keys.append(m.get_unique_key())
            if len(feats) < 2:
                continue
            x = np.ascontiguousarray(np.stack(feats), dtype=np.float32)
            nrm = np.sqrt(np.einsum("ij,ij->i", x, x, dtype=np.float32))          # sklearn: normalise, zero norm -> 1
            nrm[nrm == 0] = 1
            xn = (x / nrm[:, None]).astype(np.float32)
            if index is None or index.d != xn.shape[1]:
                if index is not None:
                    index.close()
                index = faiss.IndexFlatIP(xn.shape[1], device=device)
            index.reset()
            index.add(xn)
            k = min(top + 1 + CANDIDATE_MARGIN, len(keys))
            _D, I = index.search(xn, k)
            for key, row in zip(keys, rank_candidates(xn, I, top, threshold)):
                graph[key] = [keys[j] for j in row]
    finally:
        if index is not None:
            index.close()
    return graph
