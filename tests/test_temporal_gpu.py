"""GPU: TemporalAnalyzer.find_similar_sequences on the CUDA path vs. the reference's golden outputs and the oracle.

Tolerance: hit SETS identical on guard-banded inputs (no window within 1e-4 of the threshold); similarities within
2e-6 (fp32 everywhere; only the summation order inside the dot products differs from BLAS)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import synth, temporal as ot  # noqa: E402


def check_against(hits, sims, thr):
    """hits: our [(db_start, sim)] list; sims: oracle float32 [nt_windows, nd_windows]."""
    t_idx, j_idx = np.nonzero(sims >= np.float32(thr))
    want = sorted((int(j), float(sims[t, j])) for t, j in zip(t_idx, j_idx))
    got = sorted(hits)
    assert [j for j, _ in got] == [j for j, _ in want]
    np.testing.assert_allclose([s for _, s in got], [s for _, s in want], rtol=0, atol=2e-6)
    s = [x[1] for x in hits]
    assert all(a >= b for a, b in zip(s, s[1:])), "not sorted by similarity descending"


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_find_similar_sequences_matches_reference_golden(temporal_golden, name):
    import ivr_b200
    tg = temporal_golden
    case, db, target = tg["cases"][name], tg["arrays"][f"{name}_db"], tg["arrays"][f"{name}_target"]
    ta = ivr_b200.TemporalAnalyzer()
    hits = ta.find_similar_sequences(target, db, sequence_length=case["sequence_length"],
                                     similarity_threshold=case["threshold"])
    want = case["hits"]
    assert len(hits) == len(want)
    assert sorted(j for j, _ in hits) == sorted(j for j, _ in want)
    np.testing.assert_allclose(sorted(s for _, s in hits), sorted(s for _, s in want), rtol=0, atol=2e-6)
    assert all(isinstance(j, int) and isinstance(s, float) for j, s in hits)
    # the best windows come out in the reference's order (their similarities are well separated)
    assert [j for j, _ in hits[:3]] == [j for j, _ in want[:3]]
    assert [list(b) for b in ta.detect_scene_boundaries(db, threshold=0.3, min_scene_length=5)] == case["scene_boundaries"]


def test_scene_boundaries_short_clip_with_and_without_validation(temporal_golden):
    """TemporalAnalyzer(validate_inputs=False) on a clip shorter than 2 * min_scene_length computes real boundaries,
    like the reference (core.py:3601-3608) -- golden outputs of the unmodified reference, both ways."""
    import ivr_b200
    ta = ivr_b200.TemporalAnalyzer()
    sc, db = temporal_golden["cases"]["short_clip"], temporal_golden["arrays"]["a_db"]
    for c in sc["slices"]:
        sl = np.ascontiguousarray(db[c["start"]:c["start"] + sc["length"]])
        for flag, key in ((True, "validated"), (False, "unvalidated")):
            got = ta.detect_scene_boundaries(sl, threshold=sc["threshold"], min_scene_length=sc["min_scene_length"],
                                             validate_inputs=flag)
            assert [list(b) for b in got] == c[key], (c["start"], key)


@pytest.mark.parametrize("L,d,nt,nd", [(8, 128, 40, 20000), (1, 64, 5, 3000), (17, 100, 30, 5000), (128, 32, 130, 1500)])
def test_find_similar_sequences_vs_oracle(L, d, nt, nd):
    import ivr_b200
    db, _ = synth.dedup_frames_guarded(nd, d, window=1, thresholds=(0.5,), seed=21 + L)
    rng = np.random.default_rng(22 + L)
    t0 = 100
    target = (db[t0:t0 + nt] + np.float32(0.3) * rng.standard_normal((nt, d)).astype(np.float32)).astype(np.float32)
    sims = ot.window_similarities(target, db, L)
    thr = float(np.quantile(sims, 0.999))
    while (np.abs(sims - np.float32(thr)) < 1e-4).any():          # guard band around the threshold
        thr += 2.5e-4
    hits = ivr_b200.TemporalAnalyzer().find_similar_sequences(target, db, sequence_length=L, similarity_threshold=thr)
    assert len(hits) > 0
    check_against(hits, sims, thr)


def test_find_similar_sequences_chunked_database(monkeypatch):
    """The database is swept in column blocks with an (L-1)-frame overlap: same hits whatever the block size."""
    import ivr_b200
    db, _ = synth.dedup_frames_guarded(9000, 64, window=1, thresholds=(0.5,), seed=31)
    target = db[500:530].copy()
    ta = ivr_b200.TemporalAnalyzer()
    whole = ta.find_similar_sequences(target, db, sequence_length=6, similarity_threshold=0.7)
    for cols in (12, 100, 1000, 4097):
        monkeypatch.setenv("IVR_SEQ_BLOCK_COLS", str(cols))
        assert ta.find_similar_sequences(target, db, sequence_length=6, similarity_threshold=0.7) == whole
    assert any(j == 500 and s > 0.9999 for j, s in whole)        # the sequence finds itself


def test_find_similar_sequences_many_hits_regrow_the_buffer():
    import ivr_b200
    x = np.ones((400, 8), np.float32)                             # every window pair has similarity 1.0
    hits = ivr_b200.TemporalAnalyzer().find_similar_sequences(x, x, sequence_length=2, similarity_threshold=0.5)
    assert len(hits) == 399 * 399 and abs(hits[0][1] - 1.0) < 1e-6   # > 65536: the wrapper retried with a larger buffer
    assert hits[:3] == [(0, hits[0][1]), (1, hits[0][1]), (2, hits[0][1])]   # ties keep target-major, then db order


def test_find_similar_sequences_validation_and_edges():
    import ivr_b200
    from ivr_b200 import _native as nat
    ta = ivr_b200.TemporalAnalyzer()
    x = synth.gaussian_unit(20, 16, seed=1)
    with pytest.raises(ValueError, match="numpy arrays"):
        ta.find_similar_sequences(x.tolist(), x)
    with pytest.raises(ValueError, match="2D arrays"):
        ta.find_similar_sequences(x[0], x)
    assert ta.find_similar_sequences(x[:3], x, sequence_length=5) == []       # reference: warning + []
    assert ta.find_similar_sequences(x, x[:4], sequence_length=5) == []
    with pytest.raises(ValueError, match="same dimension"):
        ta.find_similar_sequences(x, synth.gaussian_unit(20, 8, seed=2))
    big = synth.gaussian_unit(200, 8, seed=3)
    with pytest.raises(nat.NativeError):
        ta.find_similar_sequences(big, big, sequence_length=129)
    z = np.zeros((10, 16), np.float32)                            # zero rows: norm -> 1, cosine 0 (sklearn)
    assert ta.find_similar_sequences(z, x, sequence_length=3, similarity_threshold=0.1) == []
    assert len(ta.find_similar_sequences(z, x, sequence_length=3, similarity_threshold=0.0)) == 8 * 18
