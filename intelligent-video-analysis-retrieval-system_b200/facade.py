"""README-named facades (boundary level B1'): ``RAGRetriever`` / ``RAGBuilder``.

The reference README (README.md:124-136, 154-158, 175-185) names these classes
although its code calls them ``UnifiedBuilderIntegration`` / ``UnifiedIndex``
(SURVEY.md section 0, fact 2).  They are thin aliases over the B1 surface.
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np

from .unified_builder import UnifiedBuilderIntegration
from .unified_index import UnifiedIndex


class RAGBuilder:
    """``RAGBuilder().build_index(embeddings, metadata) -> RAGRetriever``."""

    def __init__(self, device: int | None = None):
        self.device = device

    def build_index(self, embeddings: np.ndarray, metadata: List[Dict], normalize: bool = True) -> "RAGRetriever":
        index = UnifiedIndex(device=self.device)
        index.build_from_embeddings(embeddings, metadata, normalize=normalize)
        return RAGRetriever(index)


class RAGRetriever:
    """``search`` returns the result tuple (ids, scores, metadata join)."""

    def __init__(self, index: UnifiedIndex):
        self.index = index
        self.builder = UnifiedBuilderIntegration(system=None, logger=None)
        self.builder.unified_index = index

    def search(self, query, top_k: int = 10):
        return self.index.search(query, top_k)

    def augmented_search(self, query, top_k: int = 10):
        return self.index.augmented_search(query, top_k)

    def search_unified_fast(self, query_vector, k: int = 50, similarity_threshold: float = 0.0):
        return self.builder.search_unified_fast(query_vector, k, similarity_threshold)

    def close(self):
        self.index.close()
