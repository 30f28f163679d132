"""GPU: drop-in proof -- the REFERENCE'S OWN, UNMODIFIED classes running on the B200 kernels.

``core.FAISSRetriever``, ``unified_index.UnifiedIndex.search_vectors`` and
``unified_builder.UnifiedBuilderIntegration.search_unified_fast`` are imported from ``oracle/_ref`` (the byte-for-byte
copy tools/make_ref.py makes at build time; ``/root/reference`` does not exist on the GPU box) with
``ivr_b200.faiss_compat`` installed as ``sys.modules['faiss']`` -- INTEGRATION.md, path 1.  Their outputs are compared
with the golden outputs the same classes produced over the CPU oracle (tests/golden/search_wrappers.*): same ranks,
same 1 - ip / clamped-cosine scores, same metadata join, ids identical except ties inside the 1e-3 band.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import comparator, flat_ip, ref_runner, ref_shims  # noqa: E402

TOL = 1e-3


@pytest.fixture(scope="module")
def ref_on_gpu():
    if not ref_shims.reference_available():
        pytest.skip("oracle/_ref missing: run __graft_entry__.build() where /root/reference exists")
    import ivr_b200
    with ref_shims.reference_modules(faiss_module=ivr_b200.faiss_compat, names=("core", "unified_index", "unified_builder"),
                                     stub_transformers=True) as mods:
        yield mods


def test_reference_unified_index_search_vectors_on_b200(ref_on_gpu, search_golden):
    import ivr_b200
    sg = search_golden
    u = ref_runner.make_unified_index(ref_on_gpu["unified_index"], ivr_b200.faiss_compat, sg["xb"], sg["meta"])
    assert type(u.faiss_index) is ivr_b200.faiss_compat.IndexFlatIP and u.faiss_index.ntotal == len(sg["xb"])
    ref = flat_ip.IndexFlatIP(sg["xb"].shape[1])
    ref.add(sg["xb"])
    for key, want in sg["results"]["search_vectors"].items():
        if key == "k20_q0_even":
            got = u.search_vectors(sg["xq"][0], k=20, filter_func=lambda m: m["frame_id"] % 2 == 0)
            assert all(g["metadata"]["frame_id"] % 2 == 0 for g in got) and abs(len(got) - len(want)) <= 1
            continue
        k, q = key.split("_")
        k, qi = int(k[1:]), int(q[1:])
        got = u.search_vectors(sg["xq"][qi], k=k)                    # the reference's own loop: unified_index.py:480-538
        assert len(got) == len(want)
        assert [g["rank"] for g in got] == [w[0] for w in want]
        D = np.array([[1.0 - g["similarity_score"] for g in got]], np.float32)
        I = np.array([[g["index"] for g in got]], np.int64)
        Dr = np.array([[1.0 - w[1] for w in want]], np.float32)
        Ir = np.array([[w[2] for w in want]], np.int64)
        bad = comparator.compare_topk(D, I, Dr, Ir, lambda ids: ref.scores_of(sg["xq"][qi:qi + 1], ids), TOL)
        assert not bad, (key, bad)
        assert all(g["metadata"] is sg["meta"][g["index"]] for g in got)
    assert u.faiss_index.last_timing()["path"] in ("stream", "mma")   # it really ran on the device

    class _Sys:
        logger = None
    b = ref_on_gpu["unified_builder"].UnifiedBuilderIntegration(_Sys())
    b.unified_index = u
    for key, want in sg["results"]["search_unified_fast"].items():
        got = b.search_unified_fast(sg["xq"][1], k=30, similarity_threshold=float(key[3:]))
        assert abs(len(got) - len(want)) <= 1
        for g, w in zip(got, want):
            assert g["similarity_score"] >= float(key[3:]) and g["temporal_context"] == []
            assert type(g["metadata"]).__name__ == "KeyframeMetadata"
            assert abs(g["similarity_score"] - w[1]) < 2 * TOL


def test_reference_faiss_retriever_on_b200(ref_on_gpu, search_golden):
    import ivr_b200
    core, sg = ref_on_gpu["core"], search_golden
    raw = (sg["xb"] * np.float32(2.5)).astype(np.float32)
    kms = [core.KeyframeMetadata(folder_name=m["folder_name"], image_name=m["image_name"], frame_id=m["frame_id"],
                                 file_path=m["file_path"], clip_features=(raw[i] if i % 7 else None))
           for i, m in enumerate(sg["meta"])]
    fr = ref_runner.make_faiss_retriever(core)
    fr.build_index(raw, kms, validate_consistency=False)           # core.py:758-846, unmodified
    assert type(fr.index) is ivr_b200.faiss_compat.IndexFlatIP and fr.index.ntotal == sg["results"]["ntotal"]
    out = fr.search(sg["xq"][:3] * np.float32(1.7), k=12)           # core.py:848-930, unmodified
    want = sg["results"]["faiss_retriever_search"]
    assert len(out) == len(want) == 36
    assert [r.rank for r in out] == [w[3] for w in want]
    for q in range(3):
        g = {(r.metadata.folder_name, r.metadata.image_name): r.similarity_score for r in out[q * 12:(q + 1) * 12]}
        w = {(x[0], x[1]): x[2] for x in want[q * 12:(q + 1) * 12]}
        common = set(g) & set(w)
        assert len(common) >= 11                                    # a tie at the 12th place may swap one hit
        for key in common:
            assert abs(g[key] - w[key]) < 1e-6                      # the reference re-scores in float64/32 on the host
    one = fr.search(sg["xq"][4], k=5)
    assert [(r.metadata.folder_name, r.metadata.image_name, r.rank) for r in one] == \
        [(w[0], w[1], w[3]) for w in sg["results"]["faiss_retriever_search_1d"]]
    by_id = sg["results"]["faiss_retriever_search_by_id"]
    hits = fr.search_by_id(by_id["key"], k=7)
    assert [(r.metadata.folder_name, r.metadata.image_name) for r in hits][:1] == [tuple(by_id["hits"][0][:2])]
    assert len(hits) == len(by_id["hits"])
