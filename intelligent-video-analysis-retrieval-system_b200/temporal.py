"""``TemporalAnalyzer`` on the B200 path ("next" row 2).

Mirrors the two array-in / list-out methods of the reference's ``core.TemporalAnalyzer``:

* ``detect_scene_boundaries`` (core.py:3584-3642) -- consecutive cosines on the GPU
  (``ivr_consecutive_cosine``), the run-length rule on the host (``frame_filter``);
* ``find_similar_sequences`` (core.py:3644-3702) -- every (target window, database window) pair's mean
  frame-wise cosine, thresholded, as ONE fp32 cosine-matrix kernel plus a diagonal-mean kernel
  (``ivr_sequence_similarity``) instead of O(nt * nd * L) sklearn calls.

Same argument names, defaults, validation messages and result shapes.  Unlike the reference, a
device failure is raised (``NativeError``), not swallowed into ``[]``: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Tuple

import numpy as np

from . import _native as nat
from . import frame_filter as ff


class TemporalAnalyzer:
    def __init__(self, config=None, logger=None, device: int | None = None):
        self.config = config
        self.logger = logger
        self.device = nat.default_device() if device is None else int(device)

    def detect_scene_boundaries(self, features: np.ndarray, threshold: float = 0.3, min_scene_length: int = 5,
                                validate_inputs: bool = True) -> List[Tuple[int, int]]:
        return ff.detect_scene_boundaries(features, threshold=threshold, min_scene_length=min_scene_length,
                                          validate_inputs=validate_inputs)

    def find_similar_sequences(self, target_features: np.ndarray, database_features: np.ndarray,
                               sequence_length: int = 5, similarity_threshold: float = 0.8,
                               validate_inputs: bool = True) -> List[Tuple[int, float]]:
        """List of ``(db_start_index, similarity)`` for every (target window, database window) pair whose
        mean frame-wise cosine is >= the threshold, sorted by similarity descending (stable: ties keep
        target-major, then database order -- the reference's append order)."""
        if validate_inputs:
            if not isinstance(target_features, np.ndarray) or not isinstance(database_features, np.ndarray):
                raise ValueError("Features must be numpy arrays")
            if target_features.ndim != 2 or database_features.ndim != 2:
                raise ValueError("Features must be 2D arrays")
            if len(target_features) < sequence_length or len(database_features) < sequence_length:
                if self.logger:
                    self.logger.warning("Insufficient features for sequence comparison")
                return []
        t = np.ascontiguousarray(target_features, dtype=np.float32)
        d = np.ascontiguousarray(database_features, dtype=np.float32)
        if t.ndim != 2 or d.ndim != 2 or t.shape[1] != d.shape[1]:
            raise ValueError(f"target {t.shape} and database {d.shape} must be 2-D with the same dimension")
        if len(t) < sequence_length or len(d) < sequence_length:
            return []
        cap = 1 << 16
        while True:
            ht = np.empty(cap, np.int32)
            hj = np.empty(cap, np.int64)
            hs = np.empty(cap, np.float32)
            n = C.c_int64(0)
            nat.check(nat.lib.ivr_sequence_similarity(
                self.device, t.ctypes.data, t.shape[0], d.ctypes.data, d.shape[0], t.shape[1], int(sequence_length),
                C.c_float(similarity_threshold), cap, ht.ctypes.data, hj.ctypes.data, hs.ctypes.data, C.byref(n)))
            if n.value <= cap:
                break
            cap = int(n.value)
        m = n.value
        ht, hj, hs = ht[:m], hj[:m], hs[:m]
        order = np.lexsort((hj, ht, -hs.astype(np.float64)))      # similarity desc; ties: target start, then db start
        return [(int(hj[i]), float(hs[i])) for i in order]
