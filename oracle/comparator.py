"""Tie-aware top-k comparator (TEST INFRASTRUCTURE ONLY).

north_star: "identical top-k ids except ties within a stated 1e-3 score
tolerance, and scores within that tolerance".

Per query, with s_k = the oracle's k-th (smallest returned) score:
  (1) every returned id has oracle score >= s_k - tol;
  (2) every oracle id whose score > s_k + tol is returned;
  (3) |score_ours(id) - score_oracle(id)| <= tol for every returned id;
  (4) returned scores are non-increasing; ids are unique; -1 padding only where
      the oracle pads.
"""
from __future__ import annotations

import numpy as np

TOL = 1e-3


def compare_topk(D, I, D_ref, I_ref, score_of, tol: float = TOL):
    """Return a list of human-readable violations (empty list == parity).

    ``score_of(ids[nq,k]) -> float32[nq,k]`` gives the ORACLE's fp32 score of
    arbitrary ids for each query (e.g. ``IndexFlatIP.scores_of``).
    """
    D, I, D_ref, I_ref = map(np.asarray, (D, I, D_ref, I_ref))
    bad = []
    if D.shape != D_ref.shape or I.shape != I_ref.shape:
        return [f"shape mismatch {D.shape}/{I.shape} vs {D_ref.shape}/{I_ref.shape}"]
    if I.dtype != np.int64 or D.dtype != np.float32:
        bad.append(f"dtype mismatch D={D.dtype} I={I.dtype}")
    nq, k = I.shape
    if nq == 0 or k == 0:
        return bad
    oracle_scores = score_of(I)
    for q in range(nq):
        valid_ref = I_ref[q] >= 0
        valid = I[q] >= 0
        nv = int(valid_ref.sum())
        if int(valid.sum()) != nv or not valid[:nv].all():
            bad.append(f"q{q}: {int(valid.sum())} valid hits vs oracle {nv}")
            continue
        if nv == 0:
            continue
        ids = I[q, :nv]
        if len(np.unique(ids)) != nv:
            bad.append(f"q{q}: duplicate ids")
        d = D[q, :nv]
        if np.any(d[1:] > d[:-1]):
            bad.append(f"q{q}: scores not non-increasing")
        s_k = float(D_ref[q, nv - 1])
        os_ = oracle_scores[q, :nv]
        if np.any(os_ < s_k - tol):
            j = int(np.argmin(os_))
            bad.append(f"q{q}: id {ids[j]} has oracle score {os_[j]:.6f} < s_k-tol ({s_k:.6f})")
        must = I_ref[q, :nv][D_ref[q, :nv] > s_k + tol]
        missing = np.setdiff1d(must, ids)
        if missing.size:
            bad.append(f"q{q}: {missing.size} oracle ids above s_k+tol missing, e.g. {missing[:3]}")
        err = np.abs(d - os_)
        if np.any(err > tol):
            j = int(np.argmax(err))
            bad.append(f"q{q}: score err {err[j]:.2e} at id {ids[j]}")
        if nv < k and not np.all(I[q, nv:] == -1):
            bad.append(f"q{q}: padding ids are not -1")
        if len(bad) > 20:
            bad.append("... (truncated)")
            break
    return bad


def recall_at_k(I, I_ref) -> float:
    """Plain set overlap (diagnostic only; ties make <1.0 legitimate)."""
    I, I_ref = np.asarray(I), np.asarray(I_ref)
    hit = tot = 0
    for a, b in zip(I, I_ref):
        b = b[b >= 0]
        tot += len(b)
        hit += len(np.intersect1d(a, b))
    return hit / max(tot, 1)
