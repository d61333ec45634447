#!/usr/bin/env python
"""How even is the multi-GPU sharding?  (dev tool, ONE GPU)  Builds the C2 job, derives the shards of a W-rank run
(retrieval.ShardPlan: prefix owners balanced on their estimated decoder tokens summed over the score kinds) and scores every
rank's shard in turn on this GPU, timing each with CUDA events.  max / mean of those times is what the ranks of a real
W-GPU run wait for each other at the single all-gather."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from blim_b200 import retrieval, synth  # noqa: E402
from blim_b200.engine import TVG, TVG_PRIOR, VTG, VTG_PRIOR, ModelConfig  # noqa: E402
from blim_b200.model import BlimModel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--worlds", default="2,4,8")
    ap.add_argument("--n", type=int, default=1000)
    ap.add_argument("--dataset", default="msrvtt")
    ap.add_argument("--topk", type=int, default=16)
    a = ap.parse_args()
    cfg = ModelConfig.qwen2_7b()
    dev = torch.device("cuda", 0)
    model = BlimModel(cfg, device=0)
    eng = model.engine
    for idx, name in enumerate(synth.param_shapes(cfg)):
        eng.load_weight(name, synth.init_weight(cfg, name, idx, seed=0, device=dev, std=0.02))
    eng.set_rope(torch.float32)
    corpus = synth.make_corpus(cfg, a.dataset, n=a.n, seed=1, feat_device=dev)
    eng.set_videos(corpus.video)
    eng.set_texts(0, corpus.vtg_ids, corpus.vtg_labels)
    eng.set_texts(1, corpus.tvg_ids, corpus.tvg_labels)
    eng.set_video_vocab(corpus.video_vocab, corpus.tvg_video_labels.numpy())
    model.set_tvg_prefix_length(corpus.tvg_prefix_length)
    plan = retrieval.PairPlan(corpus.v2t_iv2.to(dev), corpus.t2v_iv2.to(dev), a.topk, dev, engine=eng)
    jobs = [("vtg", VTG) + tuple(plan.union_np), ("vtg_prior", VTG_PRIOR) + tuple(plan.v2t_np), ("tvg", TVG) + tuple(plan.union_np),
            ("tvg_prior", TVG_PRIOR) + tuple(plan.t2v_np)]
    for name, kind, pv, pt in jobs:      # warm-up: one full pass
        eng.score_pairs(kind, pv, pt)
    for world in [int(w) for w in a.worlds.split(",")]:
        sp = retrieval.ShardPlan(eng, jobs, world, plan.n_videos, plan.n_texts)
        ms = []
        for r in range(world):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for name, kind, pv, pt in jobs:
                mine = sp.shards[name][r]
                if len(mine):
                    eng.score_pairs(kind, pv[mine], pt[mine])
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        ms = np.array(ms)
        print(f"world {world}: per-rank ms {np.round(ms, 1).tolist()}  max/mean {ms.max() / ms.mean():.4f}  sum {ms.sum():.0f} ms  ideal efficiency bound {ms.mean() / ms.max():.3f}")
    eng.close()


if __name__ == "__main__":
    main()
