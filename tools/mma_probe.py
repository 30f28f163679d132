"""Bring-up probe for the tcgen05 path: prints diagnostics instead of asserting."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ivr_b200  # noqa: E402
from oracle import comparator, flat_ip, synth  # noqa: E402


def probe(n, d, nq, k, cg):
    os.environ["IVR_MMA_CTA_GROUP"] = str(cg)
    xb = synth.clip_like(n, d, seed=5, n_centres=64)
    xq = synth.clip_like(nq, d, seed=6, n_centres=64)
    ref = flat_ip.IndexFlatIP(d)
    ref.add(xb)
    idx = ivr_b200.IndexFlatIP(d)
    idx.add(xb)
    idx.search_path = 2
    idx.set_timing(True)
    t0 = time.time()
    D, I = idx.search(xq, k)
    dt = time.time() - t0
    Dr, Ir = ref.search(xq, k)
    bad = comparator.compare_topk(D, I, Dr, Ir, lambda ids: ref.scores_of(xq, ids), 1e-3)
    rec = comparator.recall_at_k(I, Ir)
    so = ref.scores_of(xq, np.where(I >= 0, I, 0))
    err = np.abs(so - D)[I >= 0].max() if (I >= 0).any() else -1
    print(f"n={n} d={d} nq={nq} k={k} cg={cg}: recall={rec:.4f} max|score-exact|={err:.2e} "
          f"violations={len(bad)} wall={dt*1e3:.1f}ms timing={idx.last_timing()}", flush=True)
    if bad:
        print("   first:", bad[:3], flush=True)
        print("   D[0,:5]", D[0, :5], "I[0,:5]", I[0, :5], "ref D", Dr[0, :5], "ref I", Ir[0, :5], flush=True)


if __name__ == "__main__":
    cgs = [int(a) for a in sys.argv[1:]] or [1]
    for cg in cgs:
        print(f"--- cta_group={cg} IVR_MMA_MODE={os.environ.get('IVR_MMA_MODE', '0')}", flush=True)
        probe(1000, 64, 8, 10, cg)
        probe(5000, 128, 130, 100, cg)
        probe(100_000, 512, 1000, 100, cg)
        probe(1_000_000, 512, 4096, 100, cg)
