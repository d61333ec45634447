#!/usr/bin/env python
"""Which decoder runs does a scoring job turn into?  (dev tool, CPU only -- no GPU, no weights)

Builds the pair set of a BASELINE.json config on a synthetic corpus, derives every rank's shard of a W-rank run
(retrieval.ShardPlan), restates the scheduler's units per score kind (csrc/engine.cu: score_vtg / score_tvg) and asks the
engine's own batch planner (blim_debug_plan_batches, host-only entry of the shared object) how it cuts them into batches.
One line per (rank, kind): number of batches, suffix tokens and prefix rows of each.  Used to see tail batches and the
run sizes a rank of an 8-GPU job works with, without a GPU box visit."""
import argparse
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from blim_b200 import _lib, retrieval, synth  # noqa: E402
from blim_b200.engine import TVG, TVG_PRIOR, VTG, VTG_PRIOR, ModelConfig  # noqa: E402


class _Lens:
    """The attributes of an Engine that the shard cost model reads (retrieval._shard_costs)."""

    def __init__(self, corpus):
        self.n_clips = corpus.n_clips
        self.tvg_prefix_length = corpus.tvg_prefix_length
        self.text_lens = {}
        for which, ids, labels in ((retrieval.TEXTS_VTG, corpus.vtg_ids, corpus.vtg_labels), (retrieval.TEXTS_TVG, corpus.tvg_ids, corpus.tvg_labels)):
            self.text_lens[which] = {"total": np.array([len(x) for x in ids], np.float64),
                                     "scored": np.array([int((l != -100).sum()) for l in labels], np.float64)}


def plan(run_tokens, prefix_tokens, max_items, reserve, prefix_len, item_lens):
    lib = _lib.load()
    counts = np.asarray([len(x) for x in item_lens], np.int32)
    flat = np.asarray([v for x in item_lens for v in x] or [0], np.int32)
    pre = np.asarray(prefix_len, np.int32)
    out = np.full(max(1, int(counts.sum())), -1, np.int32)
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    nb = lib.blim_debug_plan_batches(prefix_tokens, run_tokens, min(prefix_tokens, 8192), max_items, reserve, p(pre), p(counts), len(item_lens), p(flat), p(out))
    if nb < 0:
        return None
    unit_of = np.repeat(np.arange(len(item_lens)), counts)
    out = out[:len(flat)] if counts.sum() else out[:0]
    suf = np.bincount(out, weights=flat[:len(out)], minlength=nb).astype(int)
    rows = [int(sum(pre[u] for u in np.unique(unit_of[out == b]))) for b in range(nb)]
    return suf.tolist(), rows


def units_of(kind, pv, pt, corpus, lens):
    """(prefix_len per unit, suffix lengths per unit, max_items divisor, reserved root rows) as score_vtg / score_tvg build them."""
    nc = corpus.n_clips
    n_vis = nc * 64
    scored = lens.text_lens[retrieval.TEXTS_VTG]["scored"].astype(int)
    vtg_prompt = lens.text_lens[retrieval.TEXTS_VTG]["total"].astype(int) - scored   # header + <image> + tail
    t0 = lens.text_lens[retrieval.TEXTS_TVG]["total"].astype(int) - 3                   # position of the image token
    if kind == VTG:
        vids = np.unique(pv)
        return [int(vtg_prompt[0]) - 1 + n_vis] * len(vids), [(scored[pt[pv == v]] - 1).tolist() for v in vids], 2, 14
    if kind == VTG_PRIOR:       # one unit: the shared 26-token prompt; one suffix per distinct text
        texts = np.unique(pt)
        return [int(vtg_prompt[0]) - 1], [(scored[texts] - 1).tolist()], 2, 0
    if kind == TVG:             # unit = text, one (n_clips - 1)-row suffix per candidate video
        texts = np.unique(pt)
        return [int(t0[t]) for t in texts], [[nc - 1] * int((pt == t).sum()) for t in texts], nc, corpus.tvg_prefix_length
    keys = set(zip(t0[pt].tolist(), pv.tolist()))   # TVG prior: one unit (the visible header), one suffix per distinct (T0, video)
    return [corpus.tvg_prefix_length], [[nc] * len(keys)], nc, 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dataset", default="msrvtt")
    ap.add_argument("--n", type=int, default=None)
    ap.add_argument("--n-clips", type=int, default=None)
    ap.add_argument("--topk", type=int, default=16)
    ap.add_argument("--worlds", default="1,8")
    ap.add_argument("--ranks", type=int, default=2, help="ranks shown per world size")
    ap.add_argument("--run-tokens", type=int, default=49152)
    ap.add_argument("--prefix-tokens", type=int, default=49152)
    a = ap.parse_args()
    cfg = ModelConfig.qwen2_7b()
    cfg.mm_hidden_size = 8      # the features themselves are not needed here
    corpus = synth.make_corpus(cfg, a.dataset, n=a.n, n_clips=a.n_clips, seed=1)
    lens = _Lens(corpus)
    pp = retrieval.PairPlan(corpus.v2t_iv2, corpus.t2v_iv2, a.topk, "cpu")
    jobs = [("vtg", VTG) + tuple(pp.union_np), ("vtg_prior", VTG_PRIOR) + tuple(pp.v2t_np),
            ("tvg", TVG) + tuple(pp.union_np), ("tvg_prior", TVG_PRIOR) + tuple(pp.t2v_np)]
    for world in [int(w) for w in a.worlds.split(",")]:
        sp = retrieval.ShardPlan(lens, jobs, world, pp.n_videos, pp.n_texts) if world > 1 else None
        for r in range(min(world, a.ranks)):
            total = 0
            for name, kind, pv, pt in jobs:
                sel = sp.shards[name][r] if sp else np.arange(len(pv))
                pre, item_lens, div, reserve = units_of(kind, pv[sel], pt[sel], corpus, lens)
                res = plan(a.run_tokens, a.prefix_tokens, a.run_tokens // div, reserve, pre, item_lens)
                if res is None:
                    print(f"world {world} rank {r} {name}: does not fit")
                    continue
                suf, rows = res
                total += sum(suf) + sum(rows)
                print(f"world {world} rank {r} {name:9s}: {len(suf):3d} batches  suffix tokens {suf}  prefix rows {rows}")
            print(f"world {world} rank {r} decoder tokens in total: {total}")


if __name__ == "__main__":
    main()
