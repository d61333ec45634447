"""Prints the engine-vs-reference-golden score deviations per case / kind (how much of the 1e-2 tolerance is used)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blim_b200.engine import TVG, TVG_PRIOR, VTG, VTG_PRIOR, Engine  # noqa: E402
from oracle.make_golden import CASES, build_case  # noqa: E402

KIND = {("vtg", False): VTG, ("vtg", True): VTG_PRIOR, ("tvg", False): TVG, ("tvg", True): TVG_PRIOR}
for name in sorted(CASES):
    cfg, weights, corpus = build_case(CASES[name])
    gold = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", f"{name}.npz"))
    for rope in (torch.float32, torch.bfloat16):
        eng = Engine(cfg, max_run_tokens=4096, max_prefix_tokens=4096)
        eng.load_state_dict(weights, rope_table_dtype=rope)
        eng.set_videos(corpus.video)
        eng.set_texts(0, corpus.vtg_ids, corpus.vtg_labels)
        eng.set_texts(1, corpus.tvg_ids, corpus.tvg_labels)
        eng.set_video_vocab(corpus.video_vocab, corpus.tvg_video_labels.numpy())
        eng.set_tvg_prefix_length(corpus.tvg_prefix_length)
        out = {}
        for direction in ("v2t", "t2v"):
            for ft, cpn in KIND:
                ref = gold[f"{direction}_{ft}_{'cpn' if cpn else 'lik'}"]
                rows, cols = np.nonzero(ref != -100.0)
                pv, pt = (rows, cols) if direction == "v2t" else (cols, rows)
                got = eng.score_pairs(KIND[(ft, cpn)], pv, pt).cpu().numpy()
                d = np.abs(got - ref[rows, cols])
                out[f"{direction}_{ft}{'_cpn' if cpn else ''}"] = (round(float(d.max()), 5), round(float(d.mean()), 5))
        print(name, "rope", rope, out)
        eng.close()
