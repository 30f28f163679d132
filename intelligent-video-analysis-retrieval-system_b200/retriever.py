"""``FAISSRetriever`` and its data classes on the B200 path (boundary level B1).

Same surface as core.py:83-173 (``KeyframeMetadata``, ``SearchResult`` -- the schema a drop-in must keep) and
core.py:687-958, 1176-1234 (``FAISSRetriever.build_index / search / search_by_id``), organised for the GPU:

* ``build_index`` keeps, next to the id maps, ONE float32 matrix of the stored ``clip_features`` and their norms;
* ``search`` runs the flat inner-product search on the device for the whole query batch, then re-scores ALL
  ``nq * k`` hits at once (one gather + one batched row-wise dot + one clamp) instead of the reference's per-hit
  Python loop of ``np.dot`` + two ``np.linalg.norm`` calls, which SURVEY.md section 3.2 measured at 96 % of the
  reference's wall time.

Result semantics reproduced on purpose: the returned score is NOT the index's inner
product but the cosine between the L2-normalised query and the stored, un-normalised ``metadata.clip_features``,
clamped to [0, 1] (0.0 when the frame has no features or a zero norm); ``rank`` is 1-based; multi-query results are
one flat list, query-major; hits whose metadata fails validation are skipped when ``validate_results`` is set.
The features are read at ``build_index`` time (the reference reads them per hit): replace a frame's
``clip_features`` after the build and this retriever keeps scoring against the old ones.
"""
from __future__ import annotations

import os
import threading
from dataclasses import asdict, dataclass
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from . import faiss_compat as faiss

_LIST_FIELDS = ("neighboring_frames", "scene_boundaries", "detected_objects", "scene_tags", "similar_frames",
                "transition_frames")
_TEXT_FIELDS_CHECKED_FIRST = ("folder_name", "image_name")


@dataclass
class KeyframeMetadata:
    """Field names, order and defaults of core.py:83-104 (positional construction must keep working)."""
    folder_name: str
    image_name: str
    frame_id: int
    file_path: str
    sequence_position: int = 0
    total_frames: int = 0
    neighboring_frames: List[int] = None
    scene_boundaries: List[Tuple[int, int]] = None
    clip_features: Optional[np.ndarray] = None
    llm_description: Optional[str] = None
    detected_objects: List[str] = None
    scene_tags: List[str] = None
    confidence_score: float = 0.0
    similar_frames: List[str] = None
    transition_frames: List[str] = None

    def __post_init__(self):
        for name in _LIST_FIELDS:
            if getattr(self, name) is None:
                setattr(self, name, [])
        self._validate()

    def _validate(self):
        """Raises ValueError with the reference's messages, in the reference's order (core.py:123-134)."""
        problem = self._first_problem()
        if problem:
            raise ValueError(problem)

    def _first_problem(self) -> Optional[str]:
        for name in _TEXT_FIELDS_CHECKED_FIRST:
            value = getattr(self, name)
            if not (isinstance(value, str) and value):
                return f"{name} must be a non-empty string"
        if not isinstance(self.frame_id, int):
            return "frame_id must be an integer"
        if not (isinstance(self.file_path, str) and self.file_path):
            return "file_path must be a non-empty string"
        return None

    def to_dict(self) -> Dict[str, Any]:
        data = asdict(self)
        if self.clip_features is not None:
            data["clip_features"] = self.clip_features.tolist()
        return data

    @classmethod
    def from_dict(cls, data: Dict[str, Any]) -> "KeyframeMetadata":
        if data.get("clip_features") is not None:
            data["clip_features"] = np.array(data["clip_features"])
        return cls(**data)

    def get_unique_key(self) -> str:
        return f"{self.folder_name}_{self.image_name}"

    def validate_file_exists(self) -> bool:
        return os.path.exists(self.file_path)


@dataclass
class SearchResult:
    """core.py:160-173."""
    metadata: KeyframeMetadata
    similarity_score: float
    rank: int
    query_relevance: float = 0.0
    temporal_context: List["SearchResult"] = None
    explanation: Optional[str] = None

    def __post_init__(self):
        if self.temporal_context is None:
            self.temporal_context = []


class _NullLogger:
    def __getattr__(self, _):
        return lambda *a, **k: None


class _DictConfig:
    _D = {"retrieval.faiss_index_type": "IndexFlatIP", "retrieval.enable_gpu": True}

    def get(self, key, default=None):
        return self._D.get(key, default)


def _unit_rows(features: np.ndarray) -> np.ndarray:
    """core.py:1176-1196: 1-D -> one row, finite check, x / ||x|| with zero norms left as they are (dtype kept)."""
    if not isinstance(features, np.ndarray):
        raise ValueError("Features must be numpy array")
    if features.size == 0:
        raise ValueError("Features array is empty")
    if features.ndim not in (1, 2):
        raise ValueError(f"Features must be 1D or 2D, got {features.ndim}D")
    rows = features.reshape(1, -1) if features.ndim == 1 else features
    if not np.isfinite(rows).all():
        raise ValueError("Features contain NaN or infinite values")
    length = np.linalg.norm(rows, axis=1, keepdims=True)
    return rows / np.where(length == 0, 1, length)


class FAISSRetriever:
    """Drop-in for core.FAISSRetriever's build/search surface."""

    def __init__(self, config=None, logger=None, cache=None, device: int | None = None):
        self.config = config or _DictConfig()
        self.logger = logger or _NullLogger()
        self.cache = cache
        self.index = None
        self.index_type = self.config.get("retrieval.faiss_index_type", "IndexIVFFlat")
        self.use_gpu = True
        self.device = device
        self.dimension = None
        self.is_trained = False
        self.id_to_metadata: Dict[int, KeyframeMetadata] = {}
        self.metadata_to_id: Dict[str, int] = {}
        self.next_id = 0
        self._lock = threading.RLock()
        self._stored = None            # float32 [N, d_feat]: metadata.clip_features (zero rows where absent)
        self._stored_norm = None       # float32 [N]: their norms (0 where absent -> score 0.0)

    # ---- helpers ---------------------------------------------------------------------------------
    _normalize_and_validate_features = staticmethod(_unit_rows)

    @staticmethod
    def _calculate_proper_similarity(query_vec, target_vec):
        """One hit's score (core.py:736-756); ``search`` uses the batched form below."""
        if target_vec is None:
            return 0.0
        s = FAISSRetriever._rescore(np.asarray(query_vec).reshape(1, -1), np.asarray(target_vec).reshape(1, 1, -1),
                                    np.array([[np.linalg.norm(target_vec)]]))
        return float(s[0, 0])

    @staticmethod
    def _rescore(queries: np.ndarray, targets: np.ndarray, target_norm: np.ndarray) -> np.ndarray:
        """clamp(<q, t> / (||q|| ||t||), 0, 1) for queries [nq, d] against targets [nq, k, d]; 0 where a norm is 0."""
        dots = np.einsum("qd,qkd->qk", queries, targets)
        scale = np.linalg.norm(queries, axis=1)[:, None] * target_norm
        with np.errstate(divide="ignore", invalid="ignore"):
            cos = np.where(scale == 0, 0.0, dots / scale)
        return np.clip(cos, 0.0, 1.0)

    def _create_index(self, index_type: str, features: np.ndarray):
        """core.py:1198-1234: IVF and unknown types are forced to exact FlatIP."""
        if index_type in ("IndexFlatL2", "IndexHNSW", "IndexLSH"):
            raise RuntimeError(f"Index creation failed: {index_type} is not provided by the B200 path "
                               "(exact inner-product search only)")
        if index_type not in ("IndexFlatIP", "IndexIVFFlat"):
            self.logger.warning(f"Unknown index type: {index_type}, using IndexFlatIP")
        return faiss.IndexFlatIP(features.shape[1], device=self.device)

    def _clear_index_data(self):
        if self.index is not None and hasattr(self.index, "close"):
            self.index.close()
        self.index = None
        self.id_to_metadata, self.metadata_to_id, self.next_id = {}, {}, 0
        self._stored = self._stored_norm = None
        self.is_trained = False

    def _gather_stored_features(self, metadata_list) -> None:
        with_feat = [(i, np.asarray(m.clip_features).reshape(-1)) for i, m in enumerate(metadata_list)
                     if m.clip_features is not None]
        width = with_feat[0][1].shape[0] if with_feat else 1
        self._stored = np.zeros((len(metadata_list), width), np.float32)
        for i, f in with_feat:
            if f.shape[0] != width:
                raise ValueError(f"clip_features of metadata {i} has {f.shape[0]} values, others have {width}")
            self._stored[i] = f
        self._stored_norm = np.linalg.norm(self._stored, axis=1).astype(np.float32)

    # ---- build (core.py:758-846) -----------------------------------------------------------------
    def build_index(self, features: np.ndarray, metadata_list: List[KeyframeMetadata],
                    index_type: Optional[str] = None, validate_consistency: bool = True) -> None:
        if len(features) != len(metadata_list):
            raise ValueError(f"Features count ({len(features)}) != metadata count ({len(metadata_list)})")
        if len(features) == 0:
            raise ValueError("Cannot build index from empty feature set")
        if validate_consistency:
            if not isinstance(features, np.ndarray):
                raise ValueError("Features must be numpy array")
            if features.ndim != 2:
                raise ValueError(f"Features must be 2D array, got {features.ndim}D")
            keys = set()
            for i, m in enumerate(metadata_list):
                if not hasattr(m, "get_unique_key"):
                    raise ValueError(f"Metadata at index {i} is not KeyframeMetadata instance")
                key = m.get_unique_key()
                if key in keys:
                    raise ValueError(f"Duplicate metadata key found: {key}")
                keys.add(key)
        with self._lock:
            self._clear_index_data()
            for i, m in enumerate(metadata_list):
                try:
                    m._validate()
                except Exception as e:
                    raise ValueError(f"Invalid metadata at index {i}: {e}")
            self.id_to_metadata = dict(enumerate(metadata_list))
            self.metadata_to_id = {m.get_unique_key(): i for i, m in enumerate(metadata_list)}
            self.next_id = len(metadata_list)
            self._gather_stored_features(metadata_list)
            rows = _unit_rows(features)
            self.dimension = rows.shape[1]
            self.index = self._create_index(index_type or self.index_type, rows)
            try:
                self.index.add(rows.astype(np.float32))
                self.is_trained = True
            except Exception as e:
                raise RuntimeError(f"Failed to add vectors to index: {e}")
            if validate_consistency and self.index.ntotal != len(self.id_to_metadata):
                raise RuntimeError("Index validation failed: index/metadata size mismatch")

    # ---- search (core.py:848-930) ----------------------------------------------------------------
    def search(self, query_features: np.ndarray, k: int = 50, search_params: Optional[Dict] = None,
               validate_results: bool = True) -> List[SearchResult]:
        if not self.is_trained or not self.index:
            raise RuntimeError("Index not trained. Call build_index first.")
        if not self.id_to_metadata:
            return []
        with self._lock:
            queries = _unit_rows(query_features)
            if queries.shape[1] != self.dimension:
                raise ValueError(f"Query dimension ({queries.shape[1]}) != index dimension ({self.dimension})")
            try:
                _ip, ids = self.index.search(queries.astype(np.float32), k)     # the index's own scores are not reported
            except Exception as e:
                raise RuntimeError(f"Search operation failed: {e}")
            usable = (ids >= 0) & (ids < len(self._stored_norm))
            if validate_results:                                                # one check per distinct frame, not per hit
                distinct = np.unique(ids[usable])
                bad = [i for i in distinct.tolist() if self.id_to_metadata[i]._first_problem()]
                if bad:
                    usable &= ~np.isin(ids, bad)
            safe = np.where(usable, ids, 0)
            if self._stored.shape[1] == queries.shape[1]:
                scores = self._rescore(queries, self._stored[safe], self._stored_norm[safe])
            else:                                                               # no frame carries features
                scores = np.zeros(ids.shape)
            meta = self.id_to_metadata
            out: List[SearchResult] = []
            for row_ids, row_scores, row_ok in zip(safe.tolist(), scores.tolist(), usable.tolist()):
                out.extend(SearchResult(metadata=meta[i], similarity_score=s, rank=r, query_relevance=s)
                           for r, (i, s, ok) in enumerate(zip(row_ids, row_scores, row_ok), 1) if ok)
            return out

    def search_by_id(self, metadata_key: str, k: int = 10) -> List[SearchResult]:
        """core.py:932-958: neighbours of a stored frame, searched with its own stored features."""
        meta = self.id_to_metadata.get(self.metadata_to_id.get(metadata_key, -1))
        if meta is None or meta.clip_features is None:
            return []
        return self.search(meta.clip_features, k)
