"""Golden vectors for the two keep-chain rules that the reference only has INLINE, produced by executing the
reference's own, unmodified statements in the authoring container:

* ``video_frame_filter.extract_unique_frames`` (video_frame_filter.py:35-90) is called as it is, with the video decoder,
  the DINO model and the JPEG writer replaced by stand-ins (a fake ``cv2.VideoCapture`` that yields frame numbers, an
  ``extract_embedding`` that looks the frame's synthetic embedding up): the keep rule -- cos(e, e_prev_kept) >= 0.98
  drops -- runs exactly as written, and the CSV it writes lists the kept frame indices.
* Phase 4 of ``AdvancedKeyframeExtractor.extract_advanced_keyframes_optimized`` (filter_research_update.py:316-338) is inline code in a method
  that needs a video file.  Its statements are cut out of the parsed source by position (the assignment to
  ``final_frames`` up to the end of the ``for frame_data in selected_frames`` loop) and executed unmodified against the
  module's own globals (``cosine_similarity``, ``SIM_THRESHOLD = 0.95``, ``TEMPORAL_WINDOW = 10``).

Every cosine either rule evaluated is recorded: the script refuses data with a cosine within 1e-4 of the threshold
(guard band), so the GPU test can ask for bit-identical keep lists.

    python tests/golden/make_golden_chain.py        (needs /root/reference; writes chain_rules.npz / .json)

TEST INFRASTRUCTURE ONLY.
"""
import ast
import csv
import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_shims, synth  # noqa: E402


def frames(n, d, seed):
    """Scenes of near-duplicates with small and large noise, so both rules drop a good share of the frames."""
    x, _ = synth.dedup_frames(n, d, seed=seed, sig_lo=0.03, sig_hi=0.30)
    return x.astype(np.float32)


class Recorder:
    """Wraps sklearn's cosine_similarity: same result, every value remembered."""

    def __init__(self, fn):
        self.fn, self.values = fn, []

    def __call__(self, a, b):
        out = self.fn(a, b)
        self.values.append(float(out[0][0]))
        return out


def run_video_frame_filter(x):
    with ref_shims.reference_modules(names=("filter",)) as _:          # installs the transformers stub
        sys.path.insert(0, ref_shims.REFERENCE_DIR)
        sys.modules.pop("video_frame_filter", None)
        import video_frame_filter as vff                                  # module-level model load hits the stub
        sys.path.pop(0)
        rec = Recorder(vff.cosine_similarity)
        vff.cosine_similarity = rec

        class FakeCapture:
            def __init__(self, path):
                self.i = 0

            def isOpened(self):
                return True

            def read(self):
                if self.i >= len(x):
                    return False, None
                self.i += 1
                return True, np.full((2, 2, 3), self.i - 1, np.int32)     # the "frame" is its own index

            def get(self, prop):
                return 25.0

            def release(self):
                pass

        fake_cv2 = types.SimpleNamespace(VideoCapture=FakeCapture, CAP_PROP_FPS=5, CAP_PROP_POS_MSEC=0,
                                         COLOR_BGR2RGB=4, cvtColor=lambda f, c: f, imwrite=lambda p, f: True)
        vff.cv2 = fake_cv2

        class FakeImage:
            @staticmethod
            def fromarray(a):
                return types.SimpleNamespace(resize=lambda size: int(a[0, 0, 0]))
        vff.Image = FakeImage
        vff.extract_embedding = lambda idx: x[idx]
        tmp = tempfile.mkdtemp(prefix="ivr_vff_")
        os.makedirs(os.path.join(tmp, "map"))                           # process_videos() creates it (video_frame_filter.py:101-102)
        saved = vff.extract_unique_frames(os.path.join(tmp, "clip.mp4"), os.path.join(tmp, "kf"), os.path.join(tmp, "map"))
        with open(os.path.join(tmp, "map", "clip.csv")) as f:
            rows = list(csv.reader(f))[1:]
        kept = [int(r[3]) for r in rows]
        assert saved == len(kept)
        sys.modules.pop("video_frame_filter", None)
        return kept, rec.values, float(vff.SIM_THRESHOLD)


def run_phase4(x):
    with ref_shims.reference_modules(names=("filter_research_update",)) as mods:
        fr = mods["filter_research_update"]
        src = open(os.path.join(ref_shims.REFERENCE_DIR, "filter_research_update.py"), encoding="utf-8").read()
        tree = ast.parse(src)
        body = None
        for node in ast.walk(tree):
            if isinstance(node, ast.FunctionDef) and node.name == "extract_advanced_keyframes_optimized":
                stmts = node.body
                # the try-block wrappers differ between versions: search every statement list of the function
                lists = [n.body for n in ast.walk(node) if hasattr(n, "body") and isinstance(n.body, list)]
                for lst in lists:
                    for i, st in enumerate(lst):
                        if (isinstance(st, ast.Assign) and getattr(st.targets[0], "id", "") == "final_frames"):
                            j = i
                            while not (isinstance(lst[j], ast.For) and getattr(lst[j].iter, "id", "") == "selected_frames"):
                                j += 1
                            body = lst[i:j + 1]
                            break
                    if body:
                        break
        assert body, "Phase 4 not found in extract_advanced_keyframes_optimized"
        first, last = body[0].lineno, body[-1].end_lineno
        rec = Recorder(fr.cosine_similarity)
        g = dict(fr.__dict__)
        g["cosine_similarity"] = rec
        # frames arrive already time-ordered; the reference sorts them by pts_time first (stable)
        loc = {"selected_frames": [{"embedding": x[i], "pts_time": float(i), "idx": i} for i in range(len(x))]}
        exec(compile(ast.Module(body=body, type_ignores=[]), "filter_research_update.py", "exec"), g, loc)
        kept = [f["idx"] for f in loc["final_frames"]]
        return kept, rec.values, float(fr.SIM_THRESHOLD), int(fr.TEMPORAL_WINDOW), (first, last)


def main():
    arrays, cases = {}, {}
    for name, n, d in (("vff_d384", 1200, 384), ("vff_d64", 900, 64)):
        for seed in range(40, 400):
            x = frames(n, d, seed)
            kept, sims, thr = run_video_frame_filter(x)
            if all(abs(s - thr) > 1e-4 for s in sims):
                break
        else:
            raise SystemExit("no guard-banded data found for " + name)
        arrays[name] = x
        cases[name] = {"rule": "video_frame_filter.extract_unique_frames", "threshold": thr, "kept": kept,
                       "cosines_evaluated": len(sims), "seed": seed}
        print(name, "kept", len(kept), "of", n, "seed", seed)
    for name, n, d in (("phase4_d384", 900, 384), ("phase4_d64", 700, 64)):
        for seed in range(400, 900):
            x = frames(n, d, seed)
            kept, sims, thr, win, lines = run_phase4(x)
            if all(abs(s - thr) > 1e-4 for s in sims):
                break
        else:
            raise SystemExit("no guard-banded data found for " + name)
        arrays[name] = x
        cases[name] = {"rule": f"filter_research_update.py:{lines[0]}-{lines[1]} (Phase 4, executed from the parsed source)",
                       "threshold": thr, "temporal_window": win, "kept": kept, "cosines_evaluated": len(sims), "seed": seed}
        print(name, "kept", len(kept), "of", n, "seed", seed, "lines", lines)
    np.savez_compressed(os.path.join(HERE, "chain_rules.npz"), **arrays)
    with open(os.path.join(HERE, "chain_rules.json"), "w") as f:
        json.dump(cases, f)


if __name__ == "__main__":
    main()
