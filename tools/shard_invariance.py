#!/usr/bin/env python
"""How much do per-pair scores depend on how the pair set is sharded?  One GPU emulates the W ranks of a multi-GPU run
(retrieval.ShardPlan) and compares every score kind with the unsharded run.  Prints max |d| and how many scores
differ at all.  (Different shards batch different sequences into one attention tile, so own-key chunk boundaries and the
lazy-rescale history differ: last-bit differences before a bf16 rounding, never more than rounding noise.)

  python tools/shard_invariance.py [--n 200 --world 8 --model qwen2_7b|tiny]"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=200)
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--model", default="qwen2_7b")
    args = ap.parse_args()
    from blim_b200 import retrieval, synth
    from blim_b200.engine import TVG, TVG_PRIOR, VTG, VTG_PRIOR, ModelConfig
    from blim_b200.model import BlimModel
    cfg = ModelConfig.qwen2_7b() if args.model == "qwen2_7b" else ModelConfig.tiny()
    dev = torch.device("cuda", 0)
    model = BlimModel(cfg, device=0)
    eng = model.engine
    for idx, name in enumerate(synth.param_shapes(cfg)):
        eng.load_weight(name, synth.init_weight(cfg, name, idx, seed=0, device=dev, std=0.02))
    eng.set_rope(torch.float32)
    corpus = synth.make_corpus(cfg, "msrvtt", n=args.n, seed=1, feat_device=dev)
    eng.set_videos(corpus.video)
    eng.set_texts(0, corpus.vtg_ids, corpus.vtg_labels)
    eng.set_texts(1, corpus.tvg_ids, corpus.tvg_labels)
    eng.set_video_vocab(corpus.video_vocab, corpus.tvg_video_labels.numpy())
    model.set_tvg_prefix_length(corpus.tvg_prefix_length)
    plan = retrieval.PairPlan(corpus.v2t_iv2.to(dev), corpus.t2v_iv2.to(dev), 16, dev, engine=eng)
    jobs = [("vtg", VTG) + tuple(plan.union_np), ("vtg_prior", VTG_PRIOR) + tuple(plan.v2t_np), ("tvg", TVG) + tuple(plan.union_np),
            ("tvg_prior", TVG_PRIOR) + tuple(plan.t2v_np)]
    sp = retrieval.ShardPlan(eng, jobs, args.world, plan.n_videos, plan.n_texts)
    for label, kind, pv, pt in jobs:
        full = eng.score_pairs(kind, pv, pt).cpu().numpy()
        got = np.empty_like(full)
        for r in range(args.world):
            mine = sp.shards[label][r]
            if len(mine):
                got[mine] = eng.score_pairs(kind, pv[mine], pt[mine]).cpu().numpy()
        d = np.abs(got - full)
        print(f"{label:10s} pairs {len(full):6d}  differing {int((d > 0).sum()):6d}  max |d| {d.max():.2e}  mean |d| {d.mean():.2e}")


if __name__ == "__main__":
    main()
