"""Test configuration.

``-m "not gpu"`` : oracle vs. the golden vectors made from the reference, host
logic, C-ABI symbol check, 2-rank gloo test of the sharded path -- all CPU.
``-m gpu``       : parity tests proper; they call the CUDA path through the C ABI
and use ``oracle/`` only as the checker.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (sm_100a); run on the B200 box")


def _has_gpu() -> bool:
    try:
        import ivr_b200
        return ivr_b200._native.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def dedup_golden():
    return {name: dict(np.load(os.path.join(GOLDEN, name + ".npz")))
            for name in ("dedup_d64", "dedup_d512")}


@pytest.fixture(scope="session")
def search_golden():
    import json
    arrs = dict(np.load(os.path.join(GOLDEN, "search_wrappers.npz")))
    with open(os.path.join(GOLDEN, "search_wrappers.json")) as f:
        js = json.load(f)
    n = js["n"]
    meta = [{"file_path": f"keyframes/L{(i // 100):02d}_V001/{i % 100:04d}.jpg",
             "folder_name": f"L{(i // 100):02d}_V001", "image_name": f"{i % 100:04d}",
             "frame_id": i % 100, "file_hash": f"{i:016x}", "file_size": 1000 + i}
            for i in range(n)]
    return {"xb": arrs["xb"], "xq": arrs["xq"], "meta": meta, "results": js["results"]}


@pytest.fixture(scope="session")
def temporal_golden():
    """Outputs of the reference's TemporalAnalyzer (tests/golden/make_golden_temporal.py)."""
    import json
    arrs = dict(np.load(os.path.join(GOLDEN, "temporal.npz")))
    with open(os.path.join(GOLDEN, "temporal.json")) as f:
        cases = json.load(f)
    return {"arrays": arrs, "cases": cases}


@pytest.fixture(scope="session")
def research_golden():
    """Outputs of the reference's AdvancedKeyframeExtractor (tests/golden/make_golden_research.py)."""
    import json
    arrs = dict(np.load(os.path.join(GOLDEN, "research.npz")))
    with open(os.path.join(GOLDEN, "research.json")) as f:
        cases = json.load(f)
    return {"arrays": arrs, "cases": cases}


@pytest.fixture(scope="session")
def relationships_golden():
    """Output of the reference's MetadataManager._build_similarity_relationships (make_golden_relationships.py)."""
    import json
    feats = dict(np.load(os.path.join(GOLDEN, "relationships.npz")))
    with open(os.path.join(GOLDEN, "relationships.json")) as f:
        graph = json.load(f)["graph"]
    return {"features": feats, "graph": graph}


@pytest.fixture(scope="session")
def chain_golden():
    """Outputs of the reference's own inline keep-chain rules (tests/golden/make_golden_chain.py)."""
    import json
    arrs = dict(np.load(os.path.join(GOLDEN, "chain_rules.npz")))
    with open(os.path.join(GOLDEN, "chain_rules.json")) as f:
        cases = json.load(f)
    return {"arrays": arrs, "cases": cases}
