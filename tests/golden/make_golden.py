#!/usr/bin/env python
"""Generate the golden fixtures in this directory FROM THE UNMODIFIED REFERENCE.

Runs only in the authoring container (needs /root/reference); the fixtures it
writes are committed so the tests can run on the GPU box where the reference is
absent.  Usage:  python tests/golden/make_golden.py

Fixtures
--------
dedup_d64.npz, dedup_d512.npz
    Inputs (un-normalised float32 frames, guard-banded: no banded cosine within
    1e-4 of any threshold used) and the outputs of the reference's own
    filter.py functions: calculate_similarities, detect_scene_transitions,
    group_into_scenes, filter_similar_frames_advanced,
    filter_similar_frames_in_scene, apply_similarity_filtering_to_scenes.
search_wrappers.npz + search_wrappers.json
    A small normalised DB + queries and the outputs of the reference's own
    UnifiedIndex.search_vectors (unified_index.py:480-538),
    UnifiedBuilderIntegration.search_unified_fast (unified_builder.py:190-251)
    and FAISSRetriever.build_index/search/search_by_id (core.py:758-958), run on
    top of the oracle's IndexFlatIP through a ``faiss``-shaped shim (FAISS
    itself is not installed; see oracle/__init__.py).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shims, synth  # noqa: E402


def _cfg(**kw):
    c = dict(enable_similarity_filtering=True, similarity_threshold=0.95,
             min_frame_distance=1, similarity_window_size=5,
             use_advanced_similarity_filtering=False)
    c.update(kw)
    return c


def make_dedup(rf, name, n, d, seed):
    thresholds = (0.95, 0.75, 0.98, 0.9, 0.3)
    x, _ = synth.dedup_frames_guarded(n, d, window=8, thresholds=thresholds, seed=seed)
    emb = [r for r in x]
    out = {"x": x}
    sims = rf.calculate_similarities(emb)
    out["sims"] = np.asarray(sims, np.float32)
    for thr_t, tag in ((0.75, "t075"), (0.3, "t030")):
        tp = rf.detect_scene_transitions(sims, thr_t)
        out[f"transitions_{tag}"] = np.asarray(tp, np.int64)
        for ml in (1, 2, 5):
            sc = rf.group_into_scenes(tp, n, ml)
            out[f"scenes_{tag}_m{ml}"] = np.asarray(sc, np.int64).reshape(-1, 2)
    scenes = rf.group_into_scenes(rf.detect_scene_transitions(sims, 0.75), n, 2)
    rows = list(range(n))
    variants = {
        "adv_w8_t095": _cfg(use_advanced_similarity_filtering=True, similarity_window_size=8),
        "adv_w5_t095": _cfg(use_advanced_similarity_filtering=True, similarity_window_size=5),
        "adv_w3_t090": _cfg(use_advanced_similarity_filtering=True, similarity_window_size=3,
                            similarity_threshold=0.90),
        "adv_w1_t095": _cfg(use_advanced_similarity_filtering=True, similarity_window_size=1),
        "basic_m1_t095": _cfg(),
        "basic_m2_t095": _cfg(min_frame_distance=2),
        "basic_m1_t098": _cfg(similarity_threshold=0.98),
        "disabled": _cfg(enable_similarity_filtering=False),
    }
    import contextlib
    import io
    for tag, cfg in variants.items():
        with contextlib.redirect_stdout(io.StringIO()):
            _, kept_rows, stats = rf.apply_similarity_filtering_to_scenes(emb, rows, scenes, cfg)
        out[f"kept_{tag}"] = np.asarray(kept_rows, np.int64)
        out[f"stats_{tag}"] = np.asarray([stats["original"], stats["filtered"], stats["removed"]],
                                         np.int64)
    # whole sequence as ONE scene (long-scene case) through the per-scene functions
    whole = list(range(n))
    out["kept_whole_adv_w8"] = np.asarray(
        rf.filter_similar_frames_advanced(emb, whole, variants["adv_w8_t095"]), np.int64)
    out["kept_whole_basic"] = np.asarray(
        rf.filter_similar_frames_in_scene(emb, whole, variants["basic_m1_t095"]), np.int64)
    # None handling of calculate_similarities (filter.py:146-150)
    emb_none = list(emb[:40])
    for i in (0, 7, 8, 39):
        emb_none[i] = None
    out["sims_none40"] = np.asarray(rf.calculate_similarities(emb_none), np.float32)
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, {k: v.shape for k, v in out.items() if k.startswith("kept")})


def make_search(mods):
    core, ui, ub = mods["core"], mods["unified_index"], mods["unified_builder"]
    n, d, nq = 1500, 64, 6
    xb = synth.clip_like(n, d, seed=11, n_centres=32)
    xq = synth.clip_like(nq, d, seed=12, n_centres=32)
    meta = [{"file_path": f"keyframes/L{(i // 100):02d}_V001/{i % 100:04d}.jpg",
             "folder_name": f"L{(i // 100):02d}_V001", "image_name": f"{i % 100:04d}",
             "frame_id": i % 100, "file_hash": f"{i:016x}", "file_size": 1000 + i}
            for i in range(n)]
    res = {}

    # --- UnifiedIndex.search_vectors on the oracle faiss -----------------
    faiss = sys.modules["faiss"]
    index = faiss.IndexFlatIP(d)
    chunk = xb.copy()
    faiss.normalize_L2(chunk)
    index.add(chunk.astype("float32"))
    u = ui.UnifiedIndex()
    u.faiss_index, u.metadata_list, u.is_loaded, u.vectors = index, meta, True, xb
    u.memory_maps = {"thumbnails": {}, "temporal": {}}
    sv = {}
    for k in (10, 50, 2000):
        for qi in range(nq if k < 1000 else 1):      # k > ntotal: -1 padding path, one query is enough
            r = u.search_vectors(xq[qi], k=k)
            sv[f"k{k}_q{qi}"] = [[x["rank"], x["similarity_score"], x["index"],
                                  x["metadata"]["folder_name"], x["metadata"]["frame_id"]] for x in r]
    r = u.search_vectors(xq[0], k=20, filter_func=lambda m: m["frame_id"] % 2 == 0)
    sv["k20_q0_even"] = [[x["rank"], x["similarity_score"], x["index"],
                          x["metadata"]["folder_name"], x["metadata"]["frame_id"]] for x in r]
    res["search_vectors"] = sv

    # --- UnifiedBuilderIntegration.search_unified_fast -------------------
    class _Sys:
        logger = None
    b = ub.UnifiedBuilderIntegration(_Sys())
    b.unified_index = u
    suf = {}
    for thr in (0.0, 0.5, 0.9):
        r = b.search_unified_fast(xq[1], k=30, similarity_threshold=thr)
        suf[f"thr{thr}"] = [[x["rank"], x["similarity_score"], x["index"], x["temporal_context"],
                             type(x["metadata"]).__name__,
                             getattr(x["metadata"], "folder_name", None),
                             getattr(x["metadata"], "frame_id", None)] for x in r]
    res["search_unified_fast"] = suf

    # --- FAISSRetriever build/search/search_by_id ------------------------
    class _Log:
        def __getattr__(self, _):
            return lambda *a, **k: None

    class _Cfg:
        def get(self, key, default=None):
            return {"retrieval.faiss_index_type": "IndexFlatIP",
                    "retrieval.enable_gpu": False}.get(key, default)

    class _Timer:
        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

    class _Perf:
        def timer(self, *a, **k):
            return _Timer()

    fr = core.FAISSRetriever.__new__(core.FAISSRetriever)
    fr.config, fr.logger, fr.cache = _Cfg(), _Log(), None
    fr.perf_monitor = _Perf()
    fr.index, fr.index_type, fr.use_gpu = None, "IndexFlatIP", False
    fr.dimension, fr.is_trained = None, False
    fr.id_to_metadata, fr.metadata_to_id, fr.next_id = {}, {}, 0
    fr.validator = core.DataConsistencyValidator(_Log())
    import threading
    fr._lock = threading.RLock()
    raw = (xb * np.float32(2.5)).astype(np.float32)          # un-normalised on purpose
    kms = [core.KeyframeMetadata(folder_name=m["folder_name"], image_name=m["image_name"],
                                 frame_id=m["frame_id"], file_path=m["file_path"],
                                 clip_features=(raw[i] if i % 7 else None))
           for i, m in enumerate(meta)]
    fr.build_index(raw, kms, validate_consistency=False)  # validator stats real files (core.py:300-330)
    out = fr.search(xq[:3] * np.float32(1.7), k=12)
    res["faiss_retriever_search"] = [[r.metadata.folder_name, r.metadata.image_name,
                                      float(r.similarity_score), r.rank, float(r.query_relevance)]
                                     for r in out]
    out1 = fr.search(xq[4], k=5)
    res["faiss_retriever_search_1d"] = [[r.metadata.folder_name, r.metadata.image_name,
                                         float(r.similarity_score), r.rank] for r in out1]
    key = kms[10].get_unique_key()
    outid = fr.search_by_id(key, k=7)
    res["faiss_retriever_search_by_id"] = {"key": key,
                                           "hits": [[r.metadata.folder_name, r.metadata.image_name,
                                                     float(r.similarity_score), r.rank] for r in outid]}
    res["ntotal"] = int(fr.index.ntotal)

    np.savez_compressed(os.path.join(HERE, "search_wrappers.npz"), xb=xb, xq=xq)
    with open(os.path.join(HERE, "search_wrappers.json"), "w") as f:
        json.dump({"meta_rule": "see make_golden.py", "n": n, "d": d, "results": res}, f)
    print("search_wrappers", {k: (len(v) if hasattr(v, "__len__") else v) for k, v in res.items()})


def main():
    with ref_shims.reference_modules(names=("filter",)) as mods:
        rf = mods["filter"]
        make_dedup(rf, "dedup_d64.npz", n=900, d=64, seed=7)
        make_dedup(rf, "dedup_d512.npz", n=260, d=512, seed=8)
    with ref_shims.reference_modules(names=("core", "unified_index", "unified_builder")) as mods:
        make_search(mods)


if __name__ == "__main__":
    main()
