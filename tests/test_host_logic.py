"""CPU: the Python host wrappers (result semantics, metadata join, error behaviour) with the
oracle injected as the index backend, checked against the golden outputs of the REFERENCE'S
OWN wrappers.  (Injection is test-only; the product constructs the CUDA index.)"""
import json
import os

import numpy as np
import pytest

from ivr_b200 import frame_filter as ff
from ivr_b200.retriever import FAISSRetriever, KeyframeMetadata, SearchResult
from ivr_b200.unified_builder import UnifiedBuilderIntegration, add_unified_index_support
from ivr_b200.unified_index import UnifiedIndex, UnifiedIndexConfig
from oracle import flat_ip


def make_unified(sg):
    index = flat_ip.IndexFlatIP(sg["xb"].shape[1])
    chunk = sg["xb"].copy()
    flat_ip.normalize_L2(chunk)
    index.add(chunk)
    u = UnifiedIndex()
    u.faiss_index, u.metadata_list, u.is_loaded = index, sg["meta"], True
    u.memory_maps = {"thumbnails": {}, "temporal": {}}
    return u


def rows_of(res):
    return [[x["rank"], x["similarity_score"], x["index"], x["metadata"]["folder_name"],
             x["metadata"]["frame_id"]] for x in res]


def test_search_vectors_matches_reference(search_golden):
    sg = search_golden
    u = make_unified(sg)
    for key, want in sg["results"]["search_vectors"].items():
        if key == "k20_q0_even":
            got = rows_of(u.search_vectors(sg["xq"][0], k=20, filter_func=lambda m: m["frame_id"] % 2 == 0))
        else:
            k, q = key.split("_")
            got = rows_of(u.search_vectors(sg["xq"][int(q[1:])], k=int(k[1:])))
        assert got == want, key
    # 1 - ip: best hit has the LOWEST similarity_score; rank is 0-based
    r = u.search_vectors(sg["xq"][0], k=10)
    assert r[0]["rank"] == 0 and r[0]["similarity_score"] <= r[-1]["similarity_score"]


def test_search_vectors_requires_loaded():
    with pytest.raises(ValueError, match="Index not loaded"):
        UnifiedIndex().search_vectors(np.zeros(4, np.float32))


def test_search_unified_fast_matches_reference(search_golden):
    sg = search_golden
    b = UnifiedBuilderIntegration(system=None)
    with pytest.raises(ValueError, match="Unified index not loaded"):
        b.search_unified_fast(sg["xq"][1])
    b.unified_index = make_unified(sg)
    for key, want in sg["results"]["search_unified_fast"].items():
        thr = float(key[3:])
        res = b.search_unified_fast(sg["xq"][1], k=30, similarity_threshold=thr)
        got = [[x["rank"], x["similarity_score"], x["index"], x["temporal_context"],
                type(x["metadata"]).__name__, x["metadata"].folder_name, x["metadata"].frame_id] for x in res]
        assert got == want, key

    class Sys:
        logger = None
    s = Sys()
    assert add_unified_index_support(s) is s.unified_builder
    assert add_unified_index_support(s) is s.unified_builder


def test_batched_front_ends_equal_the_per_query_calls(search_golden):
    """search_vectors_batch / search_unified_fast_batch: one index call for all
    queries, element i identical to the single-query wrapper the reference golden vectors pin."""
    sg = search_golden
    u = make_unified(sg)
    even = lambda m: m["frame_id"] % 2 == 0      # noqa: E731
    for k, f in ((20, None), (7, even), (len(sg["xb"]) + 5, None)):
        batch = u.search_vectors_batch(sg["xq"], k=k, filter_func=f)
        assert len(batch) == len(sg["xq"])
        for i, got in enumerate(batch):
            want = u.search_vectors(sg["xq"][i], k=k, filter_func=f)
            assert [x["index"] for x in got] == [x["index"] for x in want]
            assert [x["rank"] for x in got] == [x["rank"] for x in want]
            np.testing.assert_allclose([x["similarity_score"] for x in got],
                                       [x["similarity_score"] for x in want], rtol=0, atol=1e-6)
    assert rows_of(u.search_vectors_batch(sg["xq"][0], k=20)[0])[0][:1] == [0]      # 1-D query -> one result list
    b = UnifiedBuilderIntegration(system=None)
    with pytest.raises(ValueError, match="Unified index not loaded"):
        b.search_unified_fast_batch(sg["xq"])
    b.unified_index = u
    for thr in (0.0, 0.3):
        batch = b.search_unified_fast_batch(sg["xq"], k=30, similarity_threshold=thr)
        for i, got in enumerate(batch):
            want = b.search_unified_fast(sg["xq"][i], k=30, similarity_threshold=thr)
            assert [(x["index"], x["rank"], x["temporal_context"]) for x in got] == \
                   [(x["index"], x["rank"], x["temporal_context"]) for x in want]
    with pytest.raises(ValueError, match="Index not loaded"):
        UnifiedIndex().search_vectors_batch(sg["xq"])


def test_facade_tuple(search_golden):
    sg = search_golden
    u = make_unified(sg)
    ids, scores, meta = u.search(sg["xq"], top_k=7)
    assert ids.shape == (len(sg["xq"]), 7) and ids.dtype == np.int64 and scores.dtype == np.float32
    assert np.all(np.diff(scores, axis=1) <= 0)
    assert meta[0][0] is sg["meta"][int(ids[0, 0])]
    ids2, scores2, meta2, ctx = u.augmented_search(sg["xq"][0], top_k=3)
    assert ids2.shape == (1, 3) and ctx == [[[], [], []]]


def test_faiss_retriever_matches_reference(search_golden, monkeypatch):
    sg = search_golden
    import ivr_b200.retriever as R
    monkeypatch.setattr(R.faiss, "IndexFlatIP", lambda d, device=None: flat_ip.IndexFlatIP(d))
    raw = (sg["xb"] * np.float32(2.5)).astype(np.float32)
    kms = [KeyframeMetadata(folder_name=m["folder_name"], image_name=m["image_name"], frame_id=m["frame_id"],
                            file_path=m["file_path"], clip_features=(raw[i] if i % 7 else None))
           for i, m in enumerate(sg["meta"])]
    fr = FAISSRetriever()
    with pytest.raises(RuntimeError, match="Index not trained"):
        fr.search(sg["xq"][0])
    fr.build_index(raw, kms, validate_consistency=False)
    assert fr.index.ntotal == sg["results"]["ntotal"]
    out = fr.search(sg["xq"][:3] * np.float32(1.7), k=12)
    got = [[r.metadata.folder_name, r.metadata.image_name, float(r.similarity_score), r.rank,
            float(r.query_relevance)] for r in out]
    want = sg["results"]["faiss_retriever_search"]
    assert [g[:2] + g[3:4] for g in got] == [w[:2] + w[3:4] for w in want]         # ids + 1-based ranks, flattened
    assert np.allclose([g[2] for g in got], [w[2] for w in want], atol=1e-6)
    assert all(isinstance(r, SearchResult) for r in out)
    out1 = fr.search(sg["xq"][4], k=5)
    assert [[r.metadata.folder_name, r.metadata.image_name, r.rank] for r in out1] == \
        [[w[0], w[1], w[3]] for w in sg["results"]["faiss_retriever_search_1d"]]
    byid = sg["results"]["faiss_retriever_search_by_id"]
    hits = fr.search_by_id(byid["key"], k=7)
    assert [[r.metadata.folder_name, r.metadata.image_name, r.rank] for r in hits] == \
        [[w[0], w[1], w[3]] for w in byid["hits"]]
    assert fr.search_by_id("nope") == []
    with pytest.raises(ValueError, match="Query dimension"):
        fr.search(np.ones(5, np.float32))
    with pytest.raises(ValueError, match="NaN"):
        fr.search(np.full(sg["xb"].shape[1], np.nan, np.float32))
    with pytest.raises(ValueError, match="Features count"):
        fr.build_index(raw[:3], kms[:2])


def test_scene_bookkeeping_matches_reference(dedup_golden):
    g = dedup_golden["dedup_d64"]
    sims = g["sims"]
    n = len(g["x"])
    for thr, tag in ((0.75, "t075"), (0.3, "t030")):
        tp = ff.detect_scene_transitions(list(sims), thr)
        assert tp == g[f"transitions_{tag}"].tolist()
        for ml in (1, 2, 5):
            assert np.array_equal(np.asarray(ff.group_into_scenes(tp, n, ml), np.int64).reshape(-1, 2),
                                  g[f"scenes_{tag}_m{ml}"])
    cfg = ff.create_config()
    assert cfg["similarity_threshold"] == 0.95 and cfg["similarity_window_size"] == 5
    assert UnifiedIndexConfig().thumbnail_size == (224, 224)


def test_select_representative_frame_matches_reference(research_golden):
    """Phase 3 of filter_research_update.py (136-155) on every cluster the reference produced."""
    n = 0
    for name in ("s384", "s64", "s512"):
        case, x = research_golden["cases"][name], research_golden["arrays"][name]
        for span, groups in case["clusters"].items():
            if groups == "ValueError":
                continue
            a, b = (int(v) for v in span.split(":"))
            emb = list(x[a:b])
            assert [ff.select_representative_frame(g, emb, None) for g in groups] == case["representatives"][span]
            n += sum(len(g) > 1 for g in groups)
    assert n >= 20


def test_temporal_and_cluster_wrappers_validate_before_touching_the_device():
    """Argument handling that needs no GPU: same messages / early returns as the reference (core.py:3663-3670;
    filter_research_update.py:115-116)."""
    from ivr_b200.temporal import TemporalAnalyzer

    class _Log:
        def __init__(self):
            self.warnings = []

        def warning(self, msg):
            self.warnings.append(msg)

    log = _Log()
    ta = TemporalAnalyzer(logger=log, device=0)
    x = np.zeros((20, 8), np.float32)
    with pytest.raises(ValueError, match="Features must be numpy arrays"):
        ta.find_similar_sequences(x.tolist(), x)
    with pytest.raises(ValueError, match="Features must be 2D arrays"):
        ta.find_similar_sequences(x[0], x)
    assert ta.find_similar_sequences(x[:3], x, sequence_length=5) == []
    assert log.warnings == ["Insufficient features for sequence comparison"]
    with pytest.raises(ValueError, match="Features must be numpy array"):
        ta.detect_scene_boundaries(x.tolist())
    assert ff.cluster_similar_frames([]) == [] and ff.cluster_similar_frames([x[0]]) == [[0]]
    assert ff.select_representative_frame([7], [None] * 8) == 7


def test_relationship_ranking_on_float32_cosines_reproduces_the_reference_graph(relationships_golden):
    """relationships.rank_candidates (host stage of build_similarity_relationships): given the candidates an fp16-row
    index would return (emulated here in NumPy), the float32 re-score + reference ordering reproduces the graph of
    the unmodified MetadataManager._build_similarity_relationships (core.py:3493-3531) exactly."""
    from ivr_b200.relationships import CANDIDATE_MARGIN, rank_candidates
    rg = relationships_golden
    n_keys = 0
    for folder, x in rg["features"].items():
        keep = [i for i in range(len(x)) if (i % 11) != 5]
        if len(keep) < 2:
            continue
        keys = [f"{folder}_{i:04d}" for i in keep]
        xk = x[keep].astype(np.float32)
        nrm = np.sqrt(np.einsum("ij,ij->i", xk, xk, dtype=np.float32))
        nrm[nrm == 0] = 1
        xn = (xk / nrm[:, None]).astype(np.float32)
        h = xn.astype(np.float16).astype(np.float32)                 # what the index stores
        cand = np.argsort(-(h @ h.T), axis=1, kind="stable")[:, :min(11 + CANDIDATE_MARGIN, len(keys))]
        for key, row in zip(keys, rank_candidates(xn, cand, 10, 0.7)):
            assert [keys[j] for j in row] == rg["graph"][key], key
            n_keys += 1
        # the frame itself may be missing from the candidates (an fp16 near-duplicate outranked it): same result
        cand_noself = np.where(cand == np.arange(len(keys))[:, None], -1, cand)
        assert rank_candidates(xn, cand_noself, 10, 0.7) == rank_candidates(xn, cand, 10, 0.7)
    assert n_keys == len(rg["graph"])


def test_window_size_zero_keeps_every_frame_like_the_reference():
    """filter.py:233,242: window_size = min(0, len) = 0 -> range(i, i) is empty -> nothing is compared, all frames kept.
    Needs no device (nothing is computed)."""
    from oracle import dedup as od
    x = [np.full(8, 1.0, np.float32)] * 12                     # identical frames: any real window would drop 11 of them
    for w in (0, -3):
        cfg = ff.create_config(use_advanced_similarity_filtering=True, similarity_window_size=w)
        assert ff.filter_similar_frames_advanced(x, list(range(100, 112)), cfg) == list(range(100, 112))
        assert od.filter_similar_frames_advanced(x, list(range(100, 112)), cfg) == list(range(100, 112))
        emb, rows, stats = ff.apply_similarity_filtering_to_scenes(x, list(range(12)), [(0, 4), (6, 11)], cfg)
        want = od.apply_similarity_filtering_to_scenes(x, list(range(12)), [(0, 4), (6, 11)], cfg)
        assert rows == want[1] == [0, 1, 2, 3, 4, 6, 7, 8, 9, 10, 11] and stats == want[2]


def test_faiss_flat_index_header_layout():
    """The 45-byte IndexFlatIP header faiss_compat writes and parses (faiss/impl/index_write.cpp: fourcc,
    write_index_header, WRITEXBVECTOR), pinned field by field on a hand-built buffer.  No device needed."""
    import struct
    from ivr_b200 import faiss_compat as fc
    h = fc._flat_header(768, 851_284)
    assert len(h) == 45 and h[:4] == b"IxFI"
    assert struct.unpack("<i", h[4:8]) == (768,) and struct.unpack("<q", h[8:16]) == (851_284,)
    assert struct.unpack("<qq", h[16:32]) == (1 << 20, 1 << 20)
    assert h[32:33] == b"\x01" and struct.unpack("<i", h[33:37]) == (0,)
    assert struct.unpack("<Q", h[37:45]) == (851_284 * 768,)
    assert fc._parse_flat_header(h) == (768, 851_284)
    with pytest.raises(NotImplementedError, match="IndexFlatL2"):
        fc._parse_flat_header(b"IxF2" + h[4:])
    with pytest.raises(NotImplementedError, match="unsupported FAISS index type"):
        fc._parse_flat_header(b"IwFl" + h[4:])
    with pytest.raises(ValueError, match="corrupt"):
        fc._parse_flat_header(h[:37] + struct.pack("<Q", 5))
    with pytest.raises(ValueError, match="shorter"):
        fc._parse_flat_header(h[:20])


class _OracleBackedRetriever:
    """ivr_b200.FAISSRetriever with the device index swapped for the NumPy oracle: exercises the HOST logic
    (validation, id maps, batched re-score, rank / flattening) without a GPU."""

    def __new__(cls):
        import ivr_b200
        from oracle import flat_ip

        class R(ivr_b200.FAISSRetriever):
            def _create_index(self, index_type, features):
                return flat_ip.IndexFlatIP(features.shape[1])
        return R()


def test_retriever_host_logic_reproduces_the_reference_outputs(search_golden):
    """The vectorised re-score (one gather + one batched dot for all nq * k hits) against the golden outputs of the
    reference's per-hit loop (core.py:899-924): same hits, same order, same 1-based ranks, cosines within 1e-6."""
    import ivr_b200
    sg = search_golden
    raw = (sg["xb"] * np.float32(2.5)).astype(np.float32)
    kms = [ivr_b200.KeyframeMetadata(folder_name=m["folder_name"], image_name=m["image_name"], frame_id=m["frame_id"],
                                     file_path=m["file_path"], clip_features=(raw[i] if i % 7 else None))
           for i, m in enumerate(sg["meta"])]
    fr = _OracleBackedRetriever()
    fr.build_index(raw, kms, validate_consistency=False)
    out = fr.search(sg["xq"][:3] * np.float32(1.7), k=12)
    want = sg["results"]["faiss_retriever_search"]
    assert [(r.metadata.folder_name, r.metadata.image_name, r.rank) for r in out] == [(w[0], w[1], w[3]) for w in want]
    np.testing.assert_allclose([r.similarity_score for r in out], [w[2] for w in want], rtol=0, atol=1e-6)
    assert all(r.query_relevance == r.similarity_score and r.temporal_context == [] for r in out)
    assert any(r.similarity_score == 0.0 for r in out) or all(w[2] > 0 for w in want)      # frames without features score 0.0
    one = fr.search(sg["xq"][4], k=5)                                                       # 1-D query
    assert [(r.metadata.folder_name, r.metadata.image_name, r.rank) for r in one] == \
        [(w[0], w[1], w[3]) for w in sg["results"]["faiss_retriever_search_1d"]]
    by_id = sg["results"]["faiss_retriever_search_by_id"]
    hits = fr.search_by_id(by_id["key"], k=7)
    assert [(r.metadata.folder_name, r.metadata.image_name, r.rank) for r in hits] == [(h[0], h[1], h[3]) for h in by_id["hits"]]
    np.testing.assert_allclose([r.similarity_score for r in hits], [h[2] for h in by_id["hits"]], rtol=0, atol=1e-6)
    assert fr.search_by_id("no_such_key") == []
    # validate_results: a frame whose metadata went bad after the build is skipped, the ranks of the others keep their gaps
    victim = out[1].metadata
    victim.file_path = ""
    again = fr.search(sg["xq"][:1] * np.float32(1.7), k=12)
    assert victim not in [r.metadata for r in again] and [r.rank for r in again] == [1] + list(range(3, 13))
    assert len(fr.search(sg["xq"][:1] * np.float32(1.7), k=12, validate_results=False)) == 12
    victim.file_path = "restored.jpg"
    # argument errors carry the reference's messages
    with pytest.raises(ValueError, match="Query dimension"):
        fr.search(np.zeros((1, 3), np.float32) + 1, k=3)
    with pytest.raises(ValueError, match="NaN or infinite"):
        fr.search(np.full((1, raw.shape[1]), np.nan, np.float32), k=3)
    with pytest.raises(ValueError, match="1D or 2D"):
        fr.search(np.zeros((1, 1, raw.shape[1]), np.float32), k=3)
    with pytest.raises(RuntimeError, match="Index not trained"):
        ivr_b200.FAISSRetriever().search(raw[:1], k=3)
    with pytest.raises(ValueError, match="Features count"):
        fr.build_index(raw[:5], kms[:4])
    with pytest.raises(ValueError, match="Duplicate metadata key"):
        fr.build_index(raw[:2], [kms[0], kms[0]])
    for field, msg in (("folder_name", "folder_name must be a non-empty string"), ("frame_id", "frame_id must be an integer"),
                       ("file_path", "file_path must be a non-empty string")):
        kw = dict(folder_name="f", image_name="i", frame_id=1, file_path="p")
        kw[field] = None
        with pytest.raises(ValueError, match=msg):
            ivr_b200.KeyframeMetadata(**kw)
    assert ivr_b200.FAISSRetriever._calculate_proper_similarity(raw[1], None) == 0.0
    assert abs(ivr_b200.FAISSRetriever._calculate_proper_similarity(raw[1], 3 * raw[1]) - 1.0) < 1e-6
    assert ivr_b200.FAISSRetriever._calculate_proper_similarity(raw[1], -raw[1]) == 0.0            # clamped


# ---------------------------------------------------------------------------
# .rvdb container pieces (host only: the decoders are plain C inside the .so)
# ---------------------------------------------------------------------------
def test_lz4_frame_decoder_against_a_real_lz4_library():
    """rvdb_reader.lz4_frame_decompress vs frames written by pyarrow's LZ4 codec (the real liblz4 frame writer):
    empty, tiny, multi-block, compressible and incompressible payloads, concatenated frames."""
    pa = pytest.importorskip("pyarrow")
    from ivr_b200 import rvdb_reader as rr
    rng = np.random.default_rng(0)
    payloads = [b"", b"x", bytes(range(256)) * 3, rng.integers(0, 4, 70_000, dtype=np.uint8).tobytes(),
                rng.integers(0, 256, 300_000, dtype=np.uint8).tobytes(),           # incompressible: stored blocks
                rng.integers(0, 8, 5_000_000, dtype=np.uint8).tobytes()]           # > one 4 MB block
    for p in payloads:
        frame = pa.compress(p, codec="lz4", asbytes=True)
        assert frame[:4] == b"\x04\x22\x4d\x18"
        assert rr.lz4_frame_decompress(frame) == p
        assert rr.lz4_frame_decompress(np.frombuffer(frame, np.uint8)) == p
    two = pa.compress(payloads[2], codec="lz4", asbytes=True) + pa.compress(payloads[3], codec="lz4", asbytes=True)
    assert rr.lz4_frame_decompress(two) == payloads[2] + payloads[3]
    with pytest.raises(rr.RvdbFormatError):
        rr.lz4_frame_decompress(b"\x00\x01\x02\x03\x04\x05\x06\x07")
    bad = bytearray(pa.compress(payloads[3], codec="lz4", asbytes=True))
    bad[40] ^= 0xFF
    try:                                                          # a corrupted block must fail or differ, never crash
        assert rr.lz4_frame_decompress(bytes(bad)) != payloads[3]
    except rr.RvdbFormatError:
        pass


def _rvdb_fixture(path, emb, meta, with_faiss_bytes=None, old_layout=False):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import hdf5_fixture as hf
    import pyarrow as pa
    w = hf.Writer()
    frame = pa.compress(json.dumps(meta).encode(), codec="lz4", asbytes=True)
    links = {"vectors": w.group({"embeddings": w.dataset(emb, chunks=(256, emb.shape[1]), lzf=True, do_shuffle=True,
                                                         skip_lzf_on=(1,))}),
             "metadata": w.group({"data": w.dataset(np.frombuffer(frame, np.uint8))}),
             "system": w.group({"build_info": w.dataset(np.frombuffer(b'{"v": 1}', np.uint8), compact=True)})}
    if with_faiss_bytes is not None and not old_layout:
        links["faiss_index"] = w.dataset(np.frombuffer(with_faiss_bytes, np.uint8), chunks=(1 << 16,), lzf=True)
    if with_faiss_bytes is not None and old_layout:
        links["index"] = w.group({"faiss": w.dataset(np.frombuffer(
            pa.compress(with_faiss_bytes, codec="lz4", asbytes=True), np.uint8))})
    with open(path, "wb") as f:
        f.write(w.finish(w.group(links)))


def test_rvdb_reader_on_a_spec_built_container(tmp_path):
    """HDF5 subset reader (superblock 0, v1 object headers + continuation, symbol-table groups, compact / contiguous /
    chunked layouts with a two-level chunk B-tree, shuffle + LZF with a skipped-filter chunk) on a file assembled by
    tests/hdf5_fixture.py from the published format -- self-consistency, NOT a file written by h5py (absent here)."""
    pytest.importorskip("pyarrow")
    from ivr_b200 import rvdb_reader as rr
    rng = np.random.default_rng(1)
    emb = rng.standard_normal((1500, 96)).astype(np.float32)
    emb[100:400] = np.round(emb[100:400], 1)                     # compressible region
    meta = [{"file_path": f"keyframes/L01_V001/{i:04d}.jpg", "folder_name": "L01_V001", "image_name": f"{i:04d}.jpg",
             "frame_id": i, "file_hash": f"{i:016x}", "file_size": 1000 + i} for i in range(len(emb))]
    path = str(tmp_path / "index.rvdb")
    _rvdb_fixture(path, emb, meta)
    with rr.Hdf5File(path) as f:
        assert f.keys("/") == ["metadata", "system", "vectors"] and f.keys("vectors") == ["embeddings"]
        assert f.is_group("vectors") and not f.is_group("vectors/embeddings") and "nope/x" not in f
        ds = f.dataset("vectors/embeddings")
        assert ds.shape == emb.shape and ds.dtype == np.float32 and ds.layout["chunk"] == (256, 96)
        assert [fid for fid, _ in ds.filters] == [2, 32000]
        assert np.array_equal(ds.read(), emb)
        firsts, blocks = zip(*ds.iter_row_blocks())
        assert list(firsts) == list(range(0, 1500, 256)) and np.array_equal(np.concatenate(blocks), emb)
        assert bytes(f.dataset("system/build_info").read()) == b'{"v": 1}'
        del ds, blocks
    r = rr.read_rvdb(path)
    try:
        assert r["metadata"] == meta and r["faiss_index"] is None and r["embeddings"].shape == emb.shape
    finally:
        r["file"].close()
    with open(path, "r+b") as fh:                                 # libver='latest' files are refused, not mis-read
        fh.seek(8)
        fh.write(b"\x02")
    with pytest.raises(rr.RvdbFormatError, match="superblock version 2"):
        rr.Hdf5File(path)


def test_bench_reference_arm_runs_on_the_host_and_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver launches beside ours): one JSON line with the contract keys,
    the reference's own FAISSRetriever path when oracle/_ref (or /root/reference) is present, every host core in use
    even when the launcher exported OMP_NUM_THREADS=1, and nothing printed by non-zero ranks."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
           "--rows", "500000", "--nq", "64", "--cpu-sample-rows", "20000", "--cpu-sample-queries", "8"]
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2")
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "queries/s" and line["higher_is_better"] is True
    assert line["n_gpus"] == 2 and line["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["config"]["rows"] == 500000 and line["config"]["nq"] == 64 and "workload" in line["config"]
    cb = line["cpu_baseline"]
    assert cb["value"] == line["value"] and cb["kind"] in ("reference", "port") and "sample" in cb
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count()
    assert cb["cores"] == cores
    from oracle import ref_shims
    if ref_shims.reference_available():
        assert cb["kind"] == "reference" and cb["raw_flat_search"]["kind"] == "port"
    other = subprocess.run(cmd, env=dict(env, RANK="1"), capture_output=True, text=True, timeout=600)
    assert other.returncode == 0 and other.stdout.strip() == ""
