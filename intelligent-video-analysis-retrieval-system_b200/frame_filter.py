"""Near-duplicate keyframe pruning on the B200 path (boundary level B1 / B1').

Same function names, arguments and return values as the similarity stage of the
reference's ``filter.py`` (filter.py:142-315), plus the streaming rule of
``video_frame_filter.py`` (63-70), ``TemporalAnalyzer.detect_scene_boundaries``
(core.py:3584-3642) and the README-named ``FrameFilter.apply_filters`` facade.
All cosines are computed on the GPU by the fused normalise-and-compare kernels
behind the C ABI (``ivr_consecutive_cosine``, ``ivr_dedup_window``,
``ivr_dedup_chain``); the O(n) scene bookkeeping stays on the host, as in the
reference.  Inputs are processed in float32 (the reference feeds float32 DINO /
CLIP outputs to sklearn, which keeps the dtype).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Sequence, Tuple

import numpy as np

from . import _native as nat

__all__ = ["calculate_similarities", "detect_scene_transitions", "group_into_scenes",
           "filter_similar_frames_in_scene", "filter_similar_frames_advanced",
           "apply_similarity_filtering_to_scenes", "extract_unique_frames_rule",
           "detect_scene_boundaries", "detect_scene_changes", "temporal_window_filter",
           "FrameFilter", "create_config"]


def create_config(**overrides) -> Dict:
    """Defaults of the reference's similarity stage (filter.py:30-34, 624-646)."""
    cfg = {"enable_similarity_filtering": True, "similarity_threshold": 0.95,
           "min_frame_distance": 1, "similarity_window_size": 5,
           "use_advanced_similarity_filtering": False,
           "transition_threshold": 0.75, "min_scene_length": 2}
    cfg.update(overrides)
    return cfg


# ---------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------
def _as_matrix(embeddings) -> Tuple[np.ndarray, np.ndarray]:
    """list/array of rows (entries may be None) -> (float32 [n,d] C-contiguous, none_mask[n])."""
    if isinstance(embeddings, np.ndarray) and embeddings.ndim == 2:
        return np.ascontiguousarray(embeddings, dtype=np.float32), np.zeros(len(embeddings), bool)
    n = len(embeddings)
    none = np.fromiter((e is None for e in embeddings), bool, n)
    first = next((e for e in embeddings if e is not None), None)
    if first is None:
        return np.zeros((n, 1), np.float32), none
    d = int(np.asarray(first).reshape(-1).shape[0])
    x = np.zeros((n, d), np.float32)
    for i, e in enumerate(embeddings):
        if e is not None:
            x[i] = np.asarray(e, dtype=np.float32).reshape(-1)
    return x, none


def _scene_arrays(scenes: Sequence[Tuple[int, int]]):
    a = np.ascontiguousarray([s for s, _ in scenes], dtype=np.int64)
    b = np.ascontiguousarray([e for _, e in scenes], dtype=np.int64)
    return a, b


def _dedup_window(x: np.ndarray, scenes, window: int, thr: float, device=None):
    a, b = _scene_arrays(scenes)
    keep = np.zeros(x.shape[0], np.uint8)
    cosp = np.empty(x.shape[0], np.float32)
    nat.check(nat.lib.ivr_dedup_window(
        nat.default_device() if device is None else device, x.ctypes.data, x.shape[0], x.shape[1],
        a.ctypes.data if len(a) else None, b.ctypes.data if len(b) else None, len(a),
        int(window), float(thr), keep.ctypes.data, cosp.ctypes.data))
    return keep, cosp


def _dedup_chain(x: np.ndarray, scenes, min_distance: int, thr: float, force_last: bool, device=None):
    a, b = _scene_arrays(scenes)
    keep = np.zeros(x.shape[0], np.uint8)
    nat.check(nat.lib.ivr_dedup_chain(
        nat.default_device() if device is None else device, x.ctypes.data, x.shape[0], x.shape[1],
        a.ctypes.data if len(a) else None, b.ctypes.data if len(b) else None, len(a),
        int(min_distance), float(thr), int(bool(force_last)), keep.ctypes.data))
    return keep


# ---------------------------------------------------------------------------
# filter.py surface
# ---------------------------------------------------------------------------
def calculate_similarities(embeddings) -> List:
    """filter.py:142-151: cosine of consecutive frames (1.0 where either is None)."""
    n = len(embeddings)
    if n <= 1:
        return []
    x, none = _as_matrix(embeddings)
    out = np.empty(n - 1, np.float32)
    nat.check(nat.lib.ivr_consecutive_cosine(nat.default_device(), x.ctypes.data, n, x.shape[1],
                                             out.ctypes.data))
    sims: List = list(out)
    if none.any():
        for i in np.nonzero(none[1:] | none[:-1])[0]:
            sims[i] = 1.0
    return sims


def detect_scene_transitions(similarities, threshold) -> List[int]:
    """filter.py:153-159: cut before frame i+1 where s[i] < threshold (strict)."""
    s = np.asarray(similarities, dtype=np.float64)
    return [int(i) + 1 for i in np.nonzero(s < threshold)[0]]


def group_into_scenes(transition_points, total_frames, min_length) -> List[Tuple[int, int]]:
    """filter.py:161-176: inclusive scenes; scenes shorter than min_length are dropped."""
    scenes, start = [], 0
    for t in transition_points:
        if t - start >= min_length:
            scenes.append((start, t - 1))
        start = t
    if total_frames - start >= min_length:
        scenes.append((start, total_frames - 1))
    return scenes


def filter_similar_frames_in_scene(scene_embeddings, scene_indices, config):
    """filter.py:178-222: last-kept chain, last frame of the scene always kept."""
    if not config["enable_similarity_filtering"] or len(scene_embeddings) <= 1:
        return scene_indices
    x, _ = _as_matrix(scene_embeddings)
    keep = _dedup_chain(x, [(0, len(x) - 1)], config["min_frame_distance"],
                        config["similarity_threshold"], True)
    return [scene_indices[i] for i in np.nonzero(keep)[0]]


def filter_similar_frames_advanced(scene_embeddings, scene_indices, config):
    """filter.py:224-259: sliding-window rule against already-kept frames."""
    if not config["enable_similarity_filtering"] or len(scene_embeddings) <= 1:
        return scene_indices
    if config["similarity_window_size"] <= 0:        # reference: range(i - 0, i) is empty -> nothing is ever compared
        return scene_indices
    x, _ = _as_matrix(scene_embeddings)
    window = min(config["similarity_window_size"], len(x))
    keep, _ = _dedup_window(x, [(0, len(x) - 1)], min(window, len(x) - 1) or 1,
                            config["similarity_threshold"])
    return [scene_indices[i] for i in np.nonzero(keep)[0]]


def apply_similarity_filtering_to_scenes(embeddings, valid_rows, scenes, config, verbose: bool = False):
    """filter.py:261-315: all scenes in ONE launch; returns (embeddings, rows, stats)."""
    if not config["enable_similarity_filtering"]:
        idx = [i for s, e in scenes for i in range(s, e + 1)]
        return ([embeddings[i] for i in idx], [valid_rows[i] for i in idx],
                {"original": len(idx), "filtered": len(idx), "removed": 0})
    stats = {"original": int(sum(e - s + 1 for s, e in scenes)), "filtered": 0, "removed": 0}
    kept_all: List[int] = []
    if scenes:
        x, _ = _as_matrix(embeddings)
        if config.get("use_advanced_similarity_filtering", False) and config["similarity_window_size"] <= 0:
            keep = np.ones(len(x), np.uint8)         # filter.py:233,242: an empty window keeps every frame
        elif config.get("use_advanced_similarity_filtering", False):
            longest = max(e - s + 1 for s, e in scenes)
            window = max(1, min(config["similarity_window_size"], longest - 1 if longest > 1 else 1))
            keep, _ = _dedup_window(x, scenes, window, config["similarity_threshold"])
        else:
            keep = _dedup_chain(x, scenes, config["min_frame_distance"],
                                config["similarity_threshold"], True)
        for s, e in scenes:                      # scene order, as the reference concatenates
            kept_all.extend((np.nonzero(keep[s:e + 1])[0] + s).tolist())
    stats["filtered"] = len(kept_all)
    stats["removed"] = stats["original"] - stats["filtered"]
    if verbose:
        print(f"similarity filtering: {stats['original']} -> {stats['filtered']} frames")
    return ([embeddings[i] for i in kept_all], [valid_rows[i] for i in kept_all], stats)


def extract_unique_frames_rule(embeddings, threshold: float = 0.98) -> List[int]:
    """video_frame_filter.py:63-70: keep iff first or cos(e, e_prev_kept) < threshold."""
    n = len(embeddings)
    if n == 0:
        return []
    x, _ = _as_matrix(embeddings)
    keep = _dedup_chain(x, [(0, n - 1)], 1, threshold, False)
    return np.nonzero(keep)[0].tolist()


def detect_scene_boundaries(features: np.ndarray, threshold: float = 0.3,
                            min_scene_length: int = 5, validate_inputs: bool = True) -> List[Tuple[int, int]]:
    """TemporalAnalyzer.detect_scene_boundaries (core.py:3584-3642).  The type / shape checks AND the
    "fewer than 2 * min_scene_length frames -> one scene" shortcut apply only with ``validate_inputs``
    (core.py:3601-3608); without it a short clip gets real boundaries, as in the reference."""
    if validate_inputs:
        if not isinstance(features, np.ndarray):
            raise ValueError("Features must be numpy array")
        if features.ndim != 2:
            raise ValueError("Features must be 2D array")
        if len(features) < min_scene_length * 2:
            return [(0, len(features) - 1)]
    n = len(features)
    sims = calculate_similarities(features) if n > 1 else []
    bounds, start = [], 0
    for i, s in enumerate(sims):
        if s < threshold and i - start >= min_scene_length:
            bounds.append((start, i))
            start = i + 1
    if start < n:
        bounds.append((start, n - 1))
    return bounds


def detect_scene_changes(embeddings, threshold: float = 0.7) -> List[int]:
    """AdvancedKeyframeExtractor.detect_scene_changes (filter_research_update.py:101-111):
    [0] + every i with cos(e_i, e_{i-1}) < threshold + [len] (end marker)."""
    n = len(embeddings)
    sims = np.asarray(calculate_similarities(embeddings), dtype=np.float64)
    return [0] + [int(i) + 1 for i in np.nonzero(sims < threshold)[0]] + [n]


def temporal_window_filter(embeddings, threshold: float = 0.95, temporal_window: int = 10) -> List[int]:
    """Phase 4 of AdvancedKeyframeExtractor (filter_research_update.py:316-338): walk the frames in
    time order; a frame is kept iff its cosine with EVERY frame in a FIFO of the last
    ``temporal_window`` kept frames is < threshold.  Returns the kept indices."""
    n = len(embeddings)
    if n == 0:
        return []
    x, _ = _as_matrix(embeddings)
    a = np.zeros(1, np.int64)
    b = np.full(1, n - 1, np.int64)
    keep = np.zeros(n, np.uint8)
    nat.check(nat.lib.ivr_dedup_fifo(nat.default_device(), x.ctypes.data, n, x.shape[1], a.ctypes.data,
                                     b.ctypes.data, 1, int(temporal_window), float(threshold), keep.ctypes.data))
    return np.nonzero(keep)[0].tolist()


# ---------------------------------------------------------------------------
# README facade
# ---------------------------------------------------------------------------
def cluster_similar_frames(embeddings, frame_indices=None, eps: float = 0.05, min_samples: int = 2,
                           device=None) -> List[List[int]]:
    """filter_research_update.py:113-134 (Phase 2): DBSCAN on ``1 - cosine_similarity`` inside a scene.

    The n x n cosine matrix and the eps-neighbourhood bits come from the GPU (``ivr_cosine_neighbors``); the
    labelling is scikit-learn's published ``dbscan_inner`` rule on those bits (index-order scan, depth-first
    growth through core samples, border points to the first cluster that reaches them, noise = -1) and the
    groups are returned like the reference's ``defaultdict`` -- label first-seen order, the noise label
    forming one group like any other.  ``frame_indices`` is accepted and unused, as in the reference.

    Divergence kept on purpose: with recent scikit-learn the reference RAISES ("Negative values in data")
    whenever float rounding makes some 1 - cos(e_i, e_i) slightly negative; here such distances simply count
    as <= eps (what the rule means)."""
    n = len(embeddings)
    if n < 2:
        return [[0]] if n else []
    x, _ = _as_matrix(embeddings)
    words = (n + 31) // 32
    adj = np.empty((n, words), np.uint32)
    nat.check(nat.lib.ivr_cosine_neighbors(nat.default_device() if device is None else int(device),
                                           x.ctypes.data, n, x.shape[1], C.c_float(eps), adj.ctypes.data))
    bits = np.unpackbits(adj.view(np.uint8), axis=1, bitorder="little")[:, :n].astype(bool)
    neighborhoods = [np.nonzero(row)[0] for row in bits]
    is_core = [len(nb) >= min_samples for nb in neighborhoods]
    labels = [-1] * n
    label_num, stack = 0, []
    for i in range(n):                                     # sklearn/cluster/_dbscan_inner.pyx
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
    clusters: Dict[int, List[int]] = {}
    for idx, label in enumerate(labels):
        clusters.setdefault(label, []).append(idx)
    return list(clusters.values())


def select_representative_frame(cluster_indices, embeddings, frame_info=None) -> int:
    """filter_research_update.py:136-155 (Phase 3): the cluster member closest (cosine) to the cluster's mean
    embedding; first maximum wins.  O(members x d) host arithmetic in the input dtype, in sklearn's order
    (normalise, then dot) -- glue between the GPU clustering and the caller, like the scene bookkeeping."""
    if len(cluster_indices) == 1:
        return cluster_indices[0]
    members = np.stack([np.asarray(embeddings[i]) for i in cluster_indices])
    centroid = np.mean(members, axis=0)

    def _unit(v):
        v = np.array(v, copy=True)
        if v.dtype not in (np.float32, np.float64):
            v = v.astype(np.float64)
        nrm = np.sqrt(np.einsum("ij,ij->i", v, v))
        nrm[nrm == 0.0] = 1.0
        return v / nrm[:, np.newaxis]

    best_idx, best_sim = 0, -1
    for i in range(len(cluster_indices)):
        sim = (_unit(members[i:i + 1]) @ _unit(centroid[None, :]).T)[0][0]
        if sim > best_sim:
            best_sim, best_idx = sim, i
    return cluster_indices[best_idx]


class FrameFilter:
    """``FrameFilter().apply_filters(embeddings, window=8, threshold=0.95) -> kept indices``
    (README.md:192-196): scene split on the consecutive cosine, then the windowed
    near-duplicate rule inside every scene -- one pass over the frames on the GPU."""

    def __init__(self, window: int = 8, threshold: float = 0.95, transition_threshold: float = 0.75,
                 min_scene_length: int = 2, device: int | None = None):
        self.window, self.threshold = int(window), float(threshold)
        self.transition_threshold, self.min_scene_length = float(transition_threshold), int(min_scene_length)
        self.device = device
        self.last_stats: Dict = {}

    def apply_filters(self, embeddings, window: int | None = None, threshold: float | None = None) -> np.ndarray:
        """Kept frame indices (ascending int64).  ONE C call (``ivr_frame_filter``): the frames cross PCIe once, in
        chunks whose banded-cosine kernels are queued behind their copies; the scene split and the greedy rule run on
        the device too.  Pass a page-locked array (e.g. ``torch.empty(...).pin_memory().numpy()``) to skip the staging
        copy.  Windows above ``IVR_MAX_WINDOW`` fall back to the scene-list entry point (it clamps the window to the
        longest scene first, as the reference's ``min(window, len(scene))`` does)."""
        window = self.window if window is None else int(window)
        threshold = self.threshold if threshold is None else float(threshold)
        x, _ = _as_matrix(embeddings)
        n = x.shape[0]
        if n == 0:
            return np.zeros(0, np.int64)
        if window > nat.IVR_MAX_WINDOW:
            return self._apply_filters_scene_list(x, window, threshold)
        keep = np.empty(n, np.uint8)
        stats = (C.c_int64 * 2)()
        nat.check(nat.lib.ivr_frame_filter(
            nat.default_device() if self.device is None else self.device, x.ctypes.data, n, x.shape[1], window,
            C.c_float(threshold), C.c_float(self.transition_threshold), self.min_scene_length,
            keep.ctypes.data, None, stats))
        kept = np.flatnonzero(keep).astype(np.int64)
        self.last_stats = {"original": int(stats[1]), "filtered": int(kept.size), "removed": int(stats[1]) - int(kept.size),
                           "scenes": int(stats[0])}
        return kept

    def _apply_filters_scene_list(self, x: np.ndarray, window: int, threshold: float) -> np.ndarray:
        n = x.shape[0]
        sims = np.empty(max(n - 1, 0), np.float32)
        if n > 1:
            nat.check(nat.lib.ivr_consecutive_cosine(
                nat.default_device() if self.device is None else self.device, x.ctypes.data, n,
                x.shape[1], sims.ctypes.data))
        scenes = group_into_scenes(detect_scene_transitions(sims, self.transition_threshold), n,
                                   self.min_scene_length)
        if not scenes:
            self.last_stats = {"original": 0, "filtered": 0, "removed": 0, "scenes": 0}
            return np.zeros(0, np.int64)
        longest = max(e - s + 1 for s, e in scenes)
        keep, _ = _dedup_window(x, scenes, max(1, min(window, max(longest - 1, 1))), threshold, self.device)
        kept = np.nonzero(keep)[0].astype(np.int64)
        orig = int(sum(e - s + 1 for s, e in scenes))
        self.last_stats = {"original": orig, "filtered": int(kept.size), "removed": orig - int(kept.size),
                           "scenes": len(scenes)}
        return kept
