"""Seeded synthetic inputs for the BASELINE.json configs (TEST INFRASTRUCTURE ONLY).

SURVEY.md section 8(d): CLIP-like search data is a mixture of unit cluster
centres plus isotropic noise, L2-normalised (pure i.i.d. Gaussian rows make every
top-100 at 100 M rows a 1e-3 tie); dedup data is per-scene base vectors plus
per-frame noise whose level straddles the 0.95 cosine threshold, left
UN-normalised because the reference feeds raw DINO/CLIP outputs to its filter
(filter.py:50-58, 142-151).
"""
from __future__ import annotations

import numpy as np


def _unit(x):
    n = np.linalg.norm(x, axis=1, keepdims=True)
    n[n == 0] = 1
    return (x / n).astype(np.float32)


def clip_like(n: int, d: int, seed: int, n_centres: int = 4096, centre_seed: int = 1234,
              sigma_scale: float = 0.5) -> np.ndarray:
    """x = normalize(c[z] + sigma*g), sigma = sigma_scale/sqrt(d)."""
    crng = np.random.default_rng(centre_seed)
    c = _unit(crng.standard_normal((n_centres, d), dtype=np.float32))
    rng = np.random.default_rng(seed)
    z = rng.integers(0, n_centres, size=n)
    g = rng.standard_normal((n, d), dtype=np.float32)
    return _unit(c[z] + np.float32(sigma_scale / np.sqrt(d)) * g)


def gaussian_unit(n: int, d: int, seed: int) -> np.ndarray:
    """normalize(N(0,I)) -- the adversarial-ties variant."""
    rng = np.random.default_rng(seed)
    return _unit(rng.standard_normal((n, d), dtype=np.float32))


def dedup_frames(n: int, d: int, seed: int = 7, mean_scene: float = 20.0,
                 sig_lo: float = 0.10, sig_hi: float = 0.35, scale: float = 3.0):
    """Frames e_t = scale * (b_scene + sigma_t * g_t), un-normalised float32.

    Returns (x[n,d], scene_id[n]).
    """
    rng = np.random.default_rng(seed)
    lens = []
    tot = 0
    while tot < n:
        l = int(rng.geometric(1.0 / mean_scene))
        lens.append(l)
        tot += l
    scene_id = np.repeat(np.arange(len(lens)), lens)[:n]
    b = rng.standard_normal((len(lens), d), dtype=np.float32)
    sig = rng.uniform(sig_lo, sig_hi, size=n).astype(np.float32)
    g = rng.standard_normal((n, d), dtype=np.float32)
    x = (b[scene_id] + sig[:, None] * g) * np.float32(scale)
    return x.astype(np.float32), scene_id


def dedup_frames_guarded(n: int, d: int, window: int, thresholds, seed: int = 7, eps: float = 1e-4,
                         **kw):
    """dedup_frames with frames re-drawn until no banded cosine is within eps of a threshold."""
    from . import dedup as od
    x, sid = dedup_frames(n, d, seed, **kw)
    rng = np.random.default_rng(seed + 100003)
    for _ in range(50):
        c = od.banded_cosines(x, window)
        valid = np.arange(n)[:, None] >= np.arange(1, window + 1)[None, :]
        near = np.zeros_like(valid)
        for t in thresholds:
            near |= (np.abs(c - np.float32(t)) < eps) & valid
        rows = np.nonzero(near.any(axis=1))[0]
        if rows.size == 0:
            return x, sid
        x[rows] += (0.05 * rng.standard_normal((rows.size, d))).astype(np.float32)
    raise RuntimeError("could not build a guard-banded fixture")
