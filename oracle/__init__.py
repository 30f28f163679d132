"""CPU oracle for the retrieval hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import it, and only as the checker (or as the timed CPU
baseline), never as a fallback for the CUDA path.

Modules
-------
flat_ip     NumPy restatement of the ``faiss.IndexFlatIP`` contract the
            reference calls (unified_index.py:503, 1767-1779; core.py:891).
dedup       NumPy restatement of the reference's cosine dedup rules
            (filter.py:142-315, video_frame_filter.py:63-70, core.py:3612-3630).
temporal    ``TemporalAnalyzer.find_similar_sequences`` / ``_compute_sequence_similarity``
            (core.py:3644-3702, 3812-3832).
cluster     The DBSCAN phase and representative selection of
            filter_research_update.py (113-155), with scikit-learn's published
            ``dbscan_inner`` rule restated.
comparator  The tie-aware top-k comparator.
synth       Seeded synthetic generators for the BASELINE.json configs.
ref_shims   Import shims that let the *reference's own files* run in the
            authoring container (used by the tests/golden/make_golden*.py scripts only;
            /root/reference does not exist on the GPU box).

Pinning status
--------------
* dedup:   PINNED -- checked bit-for-bit against the reference's own
           ``filter.py`` functions (golden fixtures in tests/golden/ were
           produced by importing the reference here).
* temporal, cluster: PINNED -- bit-for-bit against the reference's own
           ``TemporalAnalyzer`` / ``AdvancedKeyframeExtractor`` methods (real
           scikit-learn), fixtures tests/golden/temporal.*, research.*.
           The FIFO rule of filter_research_update.py:316-338 is inline code of
           ``process_video`` and stays restated only.
* wrappers (``search_vectors``, ``search_unified_fast``,
           ``FAISSRetriever.search``): PINNED -- fixtures produced by running
           the reference's own classes on top of ``flat_ip`` through a
           ``faiss``-shaped shim.
* flat_ip arithmetic vs. real FAISS: PARITY UNPINNED -- FAISS (un-vendored,
           version un-pinned dependency of the reference: ``import faiss`` at
           unified_index.py:31, core.py:32) is not installed in this image and
           the reference holds no golden vectors for it.  ``flat_ip`` restates
           the published IndexFlatIP contract (exact fp32 inner products,
           k largest descending, int64 insertion-order ids, -1 / -FLT_MAX
           padding) and is cross-checked against ``torch.topk`` on CPU.
"""
