"""``FAISSRetriever`` and its data classes on the B200 path (boundary level B1).

Mirrors core.py:83-173 (``KeyframeMetadata``, ``SearchResult``) and
core.py:687-958, 1176-1234 (``FAISSRetriever`` build/search/search_by_id).
The flat inner-product search runs on the GPU; the result semantics the
reference adds on top are reproduced exactly:
every hit is RE-SCORED with a manual cosine against the stored
``metadata.clip_features`` clamped to [0, 1] (0.0 when absent), ``rank`` is
1-based, multi-query results are flattened query-major.
"""
from __future__ import annotations

import os
import threading
from dataclasses import asdict, dataclass
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

from . import faiss_compat as faiss


@dataclass
class KeyframeMetadata:
    """core.py:83-157 (same fields, defaults and validation)."""
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
        for f in ("neighboring_frames", "scene_boundaries", "detected_objects", "scene_tags",
                  "similar_frames", "transition_frames"):
            if getattr(self, f) is None:
                setattr(self, f, [])
        self._validate()

    def _validate(self):
        if not self.folder_name or not isinstance(self.folder_name, str):
            raise ValueError("folder_name must be a non-empty string")
        if not self.image_name or not isinstance(self.image_name, str):
            raise ValueError("image_name must be a non-empty string")
        if not isinstance(self.frame_id, int):
            raise ValueError("frame_id must be an integer")
        if not self.file_path or not isinstance(self.file_path, str):
            raise ValueError("file_path must be a non-empty string")

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

    # ---- helpers (core.py:736-756, 1176-1196) ---------------------------
    @staticmethod
    def _calculate_proper_similarity(query_vec, target_vec):
        if target_vec is None:
            return 0.0
        dot_product = np.dot(query_vec, target_vec)
        query_norm = np.linalg.norm(query_vec)
        target_norm = np.linalg.norm(target_vec)
        if query_norm == 0 or target_norm == 0:
            return 0.0
        cos = dot_product / (query_norm * target_norm)
        return max(0.0, min(1.0, cos))

    @staticmethod
    def _normalize_and_validate_features(features: np.ndarray) -> np.ndarray:
        if not isinstance(features, np.ndarray):
            raise ValueError("Features must be numpy array")
        if features.size == 0:
            raise ValueError("Features array is empty")
        if features.ndim == 1:
            features = features.reshape(1, -1)
        elif features.ndim != 2:
            raise ValueError(f"Features must be 1D or 2D, got {features.ndim}D")
        if not np.isfinite(features).all():
            raise ValueError("Features contain NaN or infinite values")
        norms = np.linalg.norm(features, axis=1, keepdims=True)
        norms[norms == 0] = 1
        return features / norms

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
        self.is_trained = False

    # ---- build (core.py:758-846) ----------------------------------------
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
            seen = set()
            for i, m in enumerate(metadata_list):
                if not isinstance(m, KeyframeMetadata) and not hasattr(m, "get_unique_key"):
                    raise ValueError(f"Metadata at index {i} is not KeyframeMetadata instance")
                key = m.get_unique_key()
                if key in seen:
                    raise ValueError(f"Duplicate metadata key found: {key}")
                seen.add(key)
        with self._lock:
            self._clear_index_data()
            for i, m in enumerate(metadata_list):
                try:
                    m._validate()
                except Exception as e:
                    raise ValueError(f"Invalid metadata at index {i}: {e}")
                self.id_to_metadata[i] = m
                self.metadata_to_id[m.get_unique_key()] = i
            self.next_id = len(metadata_list)
            features = self._normalize_and_validate_features(features)
            self.dimension = features.shape[1]
            self.index = self._create_index(index_type or self.index_type, features)
            try:
                self.index.add(features.astype(np.float32))
                self.is_trained = True
            except Exception as e:
                raise RuntimeError(f"Failed to add vectors to index: {e}")
            if validate_consistency and self.index.ntotal != len(self.id_to_metadata):
                raise RuntimeError("Index validation failed: index/metadata size mismatch")

    # ---- search (core.py:848-930) ---------------------------------------
    def search(self, query_features: np.ndarray, k: int = 50, search_params: Optional[Dict] = None,
               validate_results: bool = True) -> List[SearchResult]:
        if not self.is_trained or not self.index:
            raise RuntimeError("Index not trained. Call build_index first.")
        if len(self.id_to_metadata) == 0:
            return []
        with self._lock:
            query_features = self._normalize_and_validate_features(query_features)
            if query_features.ndim == 1:
                query_features = query_features.reshape(1, -1)
            if query_features.shape[1] != self.dimension:
                raise ValueError(f"Query dimension ({query_features.shape[1]}) != index dimension ({self.dimension})")
            try:
                similarities, indices = self.index.search(query_features.astype(np.float32), k)
            except Exception as e:
                raise RuntimeError(f"Search operation failed: {e}")
            results: List[SearchResult] = []
            for i, (sim_scores, idx_list) in enumerate(zip(similarities, indices)):
                for rank, (_sim, idx) in enumerate(zip(sim_scores, idx_list)):
                    if idx >= 0 and idx in self.id_to_metadata:
                        metadata = self.id_to_metadata[idx]
                        if validate_results:
                            try:
                                metadata._validate()
                            except Exception:
                                continue
                        score = self._calculate_proper_similarity(query_features[i], metadata.clip_features)
                        results.append(SearchResult(metadata=metadata, similarity_score=score,
                                                    rank=rank + 1, query_relevance=score))
            return results

    def search_by_id(self, metadata_key: str, k: int = 10) -> List[SearchResult]:
        """core.py:932-958."""
        if metadata_key not in self.metadata_to_id:
            return []
        vector_id = self.metadata_to_id[metadata_key]
        if vector_id not in self.id_to_metadata:
            return []
        metadata = self.id_to_metadata[vector_id]
        if metadata.clip_features is not None:
            return self.search(metadata.clip_features, k)
        return []
