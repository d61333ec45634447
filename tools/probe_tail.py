#!/usr/bin/env python
"""Where do the milliseconds around the scoring kernels go at N = 8?  (dev tool, one GPU)
Times the host + small-kernel pieces of one evaluation step that do not shrink with the number of ranks -- PairPlan,
ShardPlan, unpacking the all-gathered buffer, compact_terms, fused rerank -- on an MSRVTT-1k-shaped problem with a tiny
model (none of them depends on the model size)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from blim_b200 import evalloop, retrieval, synth  # noqa: E402
from blim_b200.engine import TVG, TVG_PRIOR, VTG, VTG_PRIOR, ModelConfig  # noqa: E402
from blim_b200.model import BlimModel  # noqa: E402


def main():
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    cfg = ModelConfig.tiny()
    dev = torch.device("cuda", 0)
    model = BlimModel(cfg, state_dict=synth.init_weights(cfg, seed=0), device=0)
    eng = model.engine
    corpus = synth.make_corpus(cfg, "msrvtt", n=1000, seed=1)
    eng.set_videos(corpus.video)
    eng.set_texts(0, corpus.vtg_ids, corpus.vtg_labels)
    eng.set_texts(1, corpus.tvg_ids, corpus.tvg_labels)
    t2v, v2t = corpus.t2v_iv2.to(dev), corpus.v2t_iv2.to(dev)

    def timed(name, fn, reps=5):
        torch.cuda.synchronize()
        out = None
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            out = fn()
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
        print(f"{name:34s} {np.median(ts):8.2f} ms (min {min(ts):.2f})")
        return out

    plan = timed("PairPlan (topk x2, unique, D2H)", lambda: retrieval.PairPlan(v2t, t2v, 16, dev, engine=eng))
    jobs = [("vtg", VTG) + tuple(plan.union_np), ("vtg_prior", VTG_PRIOR) + tuple(plan.v2t_np), ("tvg", TVG) + tuple(plan.union_np),
            ("tvg_prior", TVG_PRIOR) + tuple(plan.t2v_np)]
    sp = timed(f"ShardPlan (world {world})", lambda: retrieval.ShardPlan(eng, jobs, world, plan.n_videos, plan.n_texts))
    gathered = torch.randn(world * sp.width, device=dev)

    def unpack():
        out = {}
        for name, kind, pv, pt in jobs:
            src, dst = sp.unpack_indices(name)
            res = torch.empty(len(pv), dtype=torch.float32, device=dev)
            res[torch.from_numpy(dst).to(dev, non_blocking=True)] = gathered[torch.from_numpy(src).to(dev, non_blocking=True)]
            out[name] = res
        return out
    s = timed("unpack all-gathered scores", unpack)
    comp = timed("compact_terms", lambda: retrieval.compact_terms(plan, s, cpn=True, full=True))
    timed("fused_rerank (2 kernels + recall)", lambda: evalloop.fused_rerank(eng, comp[0], comp[1], t2v, v2t, (0.0, 0.8), (1.0, 0.6, 0.8, 0.4)))
    eng.close()


if __name__ == "__main__":
    main()
