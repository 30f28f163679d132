"""BASELINE config E on one GPU: rows x queries sweep with the per-phase breakdown.

    python tools/sweep_e.py [--rows 1,3,10,30,100] [--nq 1,16,256,4096,16384] [--dim 512] [--out gpurun_out/sweep_e]

For every cell: median of `--reps` host-timed calls of IndexFlatIP.search_tensor (CUDA events on torch's
stream, inputs resident), plus the handle's own score / merge / prep timing of the last call.  Writes one
JSON line per cell and a markdown table.  Synthetic CLIP-like rows.
"""
import argparse
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ivr_b200  # noqa: E402

HBM_PEAK, TC_PEAK = 6552.6, 1654.2
try:
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as f:
        _p = json.load(f)
        HBM_PEAK = float(_p.get("hbm_gbs", HBM_PEAK))
        TC_PEAK = float(_p.get("bf16_tflops", TC_PEAK))
except Exception:
    pass


def clip_like(n, d, cen, g):
    z = torch.randint(0, cen.shape[0], (n,), generator=g, device="cuda")
    return torch.nn.functional.normalize(
        cen[z] + (0.5 / d ** 0.5) * torch.randn(n, d, generator=g, device="cuda"), dim=1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", default="1,3,10,30,100")
    ap.add_argument("--nq", default="1,16,256,4096,16384")
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default="gpurun_out/sweep_e")
    a = ap.parse_args()
    rows = sorted(int(float(r) * 1e6) for r in a.rows.split(","))
    nqs = [int(v) for v in a.nq.split(",")]
    d, k = a.dim, a.k
    g = torch.Generator(device="cuda").manual_seed(1234)
    cen = torch.nn.functional.normalize(torch.randn(4096, d, generator=g, device="cuda"), dim=1)
    gq = torch.Generator(device="cuda").manual_seed(4321)
    queries = clip_like(max(nqs), d, cen, gq)

    idx = ivr_b200.IndexFlatIP(d)
    idx.reserve(rows[-1])
    idx.set_timing(True)
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    cells = []
    with open(a.out + ".jsonl", "w") as fj:
        for n in rows:                                   # grow the same index: rows are a prefix of the next size
            while idx.ntotal < n:
                m = min(1_000_000, n - idx.ntotal)
                idx.add(clip_like(m, d, cen, g))
            for nq in nqs:
                q = queries[:nq].contiguous()
                for _ in range(2):
                    idx.search_tensor(q, k)
                torch.cuda.synchronize()
                ts = []
                for _ in range(a.reps):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    idx.search_tensor(q, k)
                    e1.record()
                    e1.synchronize()
                    ts.append(e0.elapsed_time(e1))
                ms = statistics.median(ts)
                t = idx.last_timing()
                c = {"rows": n, "dim": d, "nq": nq, "k": k, "ms": round(ms, 4), "qps": round(nq / ms * 1e3, 1),
                     "kernel": t["kernel"], "score_ms": round(t["score_ms"], 4), "merge_ms": round(t["merge_ms"], 4),
                     "prep_ms": round(t["prep_ms"], 4),
                     "tflops": round(2.0 * n * d * nq / ms / 1e9, 1),
                     "gbps_rows_once": round(n * d * 2 / ms / 1e6, 1)}
                c["frac_hbm"] = round(c["gbps_rows_once"] / HBM_PEAK, 3)
                c["frac_tensor"] = round(c["tflops"] / TC_PEAK, 3)
                cells.append(c)
                fj.write(json.dumps(c) + "\n")
                fj.flush()
                print(json.dumps(c), flush=True)
    with open(a.out + ".md", "w") as fm:
        fm.write(f"| rows | queries | ms (median of {a.reps}) | q/s | kernel | score / merge / prep ms | TFLOP/s (% burst peak) "
                 f"| GB/s rows-once (% HBM peak) |\n|---|---|---|---|---|---|---|---|\n")
        for c in cells:
            fm.write(f"| {c['rows'] / 1e6:g} M | {c['nq']} | {c['ms']:.3f} | {c['qps']:.0f} | `{c['kernel']}` | "
                     f"{c['score_ms']:.3f} / {c['merge_ms']:.3f} / {c['prep_ms']:.3f} | {c['tflops']:.0f} ({100 * c['frac_tensor']:.0f} %) | "
                     f"{c['gbps_rows_once']:.0f} ({100 * c['frac_hbm']:.0f} %) |\n")


if __name__ == "__main__":
    main()
