#!/usr/bin/env python
"""7B parity study on one B200 (dev tool; imports oracle/ = test infrastructure; needs baseline/_ref on the GPU box).

On a slice of the C2 workload (MSRVTT-1k shape, `--rows` rows x top-16 of every one of the six score matrices =
rows*16 pairs per matrix kind) it runs, on IDENTICAL random-init weights and synthetic features:

    ref_bf16   the unmodified reference, model in bf16 under torch.autocast(bf16), sdpa attention, batch 16 -- the
               north-star comparator, timed (= the "reference on one B200" throughput)
    ref_fp32   the unmodified reference in plain fp32 (no TF32, math attention) -- the exact value
    engine     blim_score_pairs through the C ABI

and reports per matrix max / mean |d score| for the three pairings, then the rerank agreement (candidate order,
ground-truth ranks, R@K on the slice, top-1 margins) of the fused BLiM scores.  JSON to --out, raw scores to --out.npz.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def stats(a, b):
    d = np.abs(a - b).reshape(-1)
    return {"max": float(d.max()), "mean": float(d.mean()), "p99": float(np.quantile(d, 0.99)), "frac_gt_1e-2": float((d > 1e-2).mean())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=16)
    ap.add_argument("--row0", type=int, default=0)
    ap.add_argument("--topk", type=int, default=16)
    ap.add_argument("--n", type=int, default=1000)
    ap.add_argument("--dataset", default="msrvtt")
    ap.add_argument("--model", default="qwen2_7b")
    ap.add_argument("--rich", type=int, default=0, help="1: random biases / norm weights too (tests); 0: the reference's _init_weights")
    ap.add_argument("--fp32", type=int, default=1)
    ap.add_argument("--budget-rows", type=int, default=0, help="also run tools/error_budget.py's rounding study on this many v2t rows")
    ap.add_argument("--out", default="gpurun_out/parity_7b.json")
    a = ap.parse_args()

    from blim_b200 import synth
    from blim_b200.engine import ModelConfig
    from blim_b200.model import BlimModel
    from oracle import ref_gpu, ref_harness

    if not ref_harness.reference_available():
        raise SystemExit("the reference sources are not available (baseline/_ref missing: run __graft_entry__.build() in the build container)")
    cfg = ModelConfig.qwen2_7b() if a.model == "qwen2_7b" else ModelConfig.tiny()
    dev = torch.device("cuda", 0)
    t_start = time.time()
    weights = synth.init_weights(cfg, seed=0, device=dev, std=0.02, rich=bool(a.rich))
    corpus = synth.make_corpus(cfg, a.dataset, n=a.n, seed=1)
    alpha, c = (0.0, 0.8), (1.0, 0.6, 0.8, 0.4)   # README.md:143 (MSRVTT fine-tuned recipe = bench.py C2)
    report = {"config": {"model": a.model, "dataset": a.dataset, "n": corpus.n, "rows": a.rows, "row0": a.row0, "topk": a.topk, "rich": a.rich,
                         "pairs_per_matrix": a.rows * min(a.topk, corpus.n), "alpha": alpha, "c": c}}

    # ---- engine
    model = BlimModel(cfg, state_dict=weights, device=0)
    eng = model.engine
    eng.set_videos(corpus.video)
    eng.set_texts(0, corpus.vtg_ids, corpus.vtg_labels)
    eng.set_texts(1, corpus.tvg_ids, corpus.tvg_labels)
    eng.set_video_vocab(corpus.video_vocab, corpus.tvg_video_labels.numpy())
    model.set_tvg_prefix_length(corpus.tvg_prefix_length)
    m_eng = ref_gpu.engine_matrices(eng, corpus, a.row0, a.rows, a.topk)
    torch.cuda.synchronize()
    eng.close()
    del model, eng
    torch.cuda.empty_cache()
    print(f"[{time.time() - t_start:.0f}s] engine done", flush=True)

    # ---- reference bf16 on the GPU (timed)
    rr = ref_gpu.ReferenceRunner(cfg, weights, corpus, dev, dtype=torch.bfloat16)
    rr.all_matrices((a.row0 + a.rows) % corpus.n, 1, a.topk)                       # warm-up: one row of every matrix
    m_b16, secs, pairs = rr.all_matrices(a.row0, a.rows, a.topk)
    report["reference_gpu"] = {"value": pairs / secs, "unit": "pairs/s", "pairs": pairs, "seconds": secs,
                               "what": "unmodified reference (retrieval_utils.compute_*_scores_x, sdpa, batch 16, bf16 autocast) on one B200, "
                                       "six matrices, host wall clock with synchronize on both sides, weights resident"}
    rr.close()
    del rr
    print(f"[{time.time() - t_start:.0f}s] reference bf16 done: {pairs / secs:.2f} pairs/s", flush=True)

    # ---- reference fp32 (exact)
    m_f32 = None
    if a.fp32:
        rr = ref_gpu.ReferenceRunner(cfg, weights, corpus, dev, dtype=torch.float32)
        m_f32, secs32, _ = rr.all_matrices(a.row0, a.rows, a.topk)
        report["reference_gpu_fp32_seconds"] = secs32
        if a.budget_rows:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import error_budget
            p = {k: v for k, v in rr.model.module.state_dict().items()}
            rep, raw, _ = error_budget.budget(p, cfg, corpus, list(range(a.row0, a.row0 + a.budget_rows)), a.topk)
            k = min(a.topk, corpus.n)
            e = m_eng["v2t_candidate_likelihood"][1][:a.budget_rows].reshape(-1)
            rep["ENGINE"] = stats(e, raw["fp32"])
            rep["ENGINE vs emulation"] = stats(e, raw["all (engine emulation)"])
            rep["oracle fp32 vs reference fp32"] = stats(raw["fp32"], m_f32["v2t_candidate_likelihood"][1][:a.budget_rows].reshape(-1))
            report["error_budget_vtg"] = {"pairs": a.budget_rows * k, "report": rep}
            del p
        rr.close()
        del rr
        print(f"[{time.time() - t_start:.0f}s] reference fp32 done", flush=True)

    # ---- score parity
    per = {}
    for name in ref_gpu.MATRICES:
        assert (m_eng[name][0] == m_b16[name][0]).all()
        row = {"engine_vs_ref_bf16": stats(m_eng[name][1], m_b16[name][1])}
        if m_f32 is not None:
            row["engine_vs_ref_fp32"] = stats(m_eng[name][1], m_f32[name][1])
            row["ref_bf16_vs_ref_fp32"] = stats(m_b16[name][1], m_f32[name][1])
        row["score_mean"], row["score_std_within_row"] = float(m_b16[name][1].mean()), float(m_b16[name][1].std(1).mean())
        per[name] = row
    report["score_parity"] = per

    # ---- rerank parity on the fused BLiM scores
    f_eng = ref_gpu.fused_rows(m_eng, corpus, a.row0, a.rows, alpha, c)
    f_b16 = ref_gpu.fused_rows(m_b16, corpus, a.row0, a.rows, alpha, c)
    report["rank_parity"] = {"engine_vs_ref_bf16": ref_gpu.rank_parity(f_b16, f_eng)}
    if m_f32 is not None:
        f_f32 = ref_gpu.fused_rows(m_f32, corpus, a.row0, a.rows, alpha, c)
        report["rank_parity"]["engine_vs_ref_fp32"] = ref_gpu.rank_parity(f_f32, f_eng)
        report["rank_parity"]["ref_bf16_vs_ref_fp32"] = ref_gpu.rank_parity(f_f32, f_b16)
    report["seconds_total"] = time.time() - t_start
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    json.dump(report, open(a.out, "w"), indent=1)
    raw = {}
    for tag, m in (("engine", m_eng), ("ref_bf16", m_b16), ("ref_fp32", m_f32)):
        if m is not None:
            for name, (idx, sc) in m.items():
                raw[f"{tag}.{name}"] = sc
                raw[f"idx.{name}"] = idx
    np.savez_compressed(a.out.replace(".json", ".npz"), **raw)
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
