"""TEST INFRASTRUCTURE -- generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference):   python oracle/make_golden.py
The fixtures pin oracle/blim_oracle.py (tests/test_oracle_golden.py) and, through it, the CUDA engine.  Inputs are not
stored: they are regenerated from seeds by blim_b200.synth (CPU generators, deterministic for the pinned torch).
"""
import argparse
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from blim_b200.engine import ModelConfig  # noqa: E402
from blim_b200 import synth  # noqa: E402
from oracle import ref_harness  # noqa: E402

# name -> fixture definition.  alpha / c follow the README recipes' shape (README.md:116-171).
CASES = {
    "tiny_a": dict(n=12, n_clips=4, cap_mean=6, cap_std=2, topk=4, bs=3, wseed=0, dseed=1, std=0.06,
                   alpha=(0.0, 0.9), c=(0.9, 0.2, 0.9, 0.9)),
    "tiny_b": dict(n=10, n_clips=2, cap_mean=9, cap_std=3, topk=5, bs=5, wseed=3, dseed=7, std=0.05,
                   alpha=(0.2, 0.8), c=(1.0, 0.4, 0.8, 0.6)),
}


def pad_left(seqs, fill):
    L = max(len(s) for s in seqs)
    out = torch.full((len(seqs), L), fill, dtype=torch.long)
    for i, s in enumerate(seqs):
        out[i, L - len(s):] = s
    return out


def build_case(case):
    cfg = ModelConfig.tiny()
    weights = synth.init_weights(cfg, seed=case["wseed"], std=case["std"], rich=True)
    corpus = synth.make_corpus(cfg, "msrvtt", n=case["n"], n_clips=case["n_clips"], cap_mean=case["cap_mean"], cap_std=case["cap_std"],
                               seed=case["dseed"])
    return cfg, weights, corpus


def run_reference(case):
    cfg, weights, corpus = build_case(case)
    sd = {k: v.float() for k, v in weights.items()}
    model, (ru, tu, mvf) = ref_harness.build_reference_model(cfg, sd, dtype=torch.float32, image_token_id=cfg.image_token_id)
    model.module.set_tvg_prefix_length(corpus.tvg_prefix_length)
    args = types.SimpleNamespace(topk=case["topk"], batch_size_eval=case["bs"], num_clips=corpus.n_clips)
    dev = torch.device("cpu")
    video = [v.float() for v in corpus.video]
    vocab = corpus.video_vocab.float()
    n = corpus.n
    out = {}
    with torch.no_grad():
        for direction, fn, sims in (("v2t", ru.compute_v2t_scores_x, corpus.v2t_iv2), ("t2v", ru.compute_t2v_scores_x, corpus.t2v_iv2)):
            for ft, ids_l, lab_l in (("vtg", corpus.vtg_ids, corpus.vtg_labels), ("tvg", corpus.tvg_ids, corpus.tvg_labels)):
                ids, labels = pad_left(ids_l, corpus.pad_token_id), pad_left(lab_l, -100)
                masks = pad_left([torch.ones_like(x) for x in ids_l], 0)
                for cpn in (False, True):
                    m = torch.full((n, n), -100.0)
                    m = fn(m, sims, 0, ids, masks, labels, video, vocab, corpus.tvg_video_labels, model, dev, args, forward_type=ft, cpn=cpn)
                    out[f"{direction}_{ft}_{'cpn' if cpn else 'lik'}"] = m.numpy().astype(np.float32)
    # dictionaries exactly as evaluation() packs them (retrieval_utils.py:264-276), full six-matrix branch
    t2v = {"candidate_likelihood": out["t2v_tvg_lik"], "query_likelihood": out["t2v_vtg_lik"], "internvideo2": corpus.t2v_iv2.numpy(),
           "candidate_prior": out["t2v_tvg_cpn"]}
    v2t = {"candidate_likelihood": out["v2t_vtg_lik"], "query_likelihood": out["v2t_tvg_lik"], "internvideo2": corpus.v2t_iv2.numpy(),
           "candidate_prior": out["v2t_vtg_cpn"]}
    # val_one_epoch arithmetic + get_recall of the reference (training_utils.py:154-165, 173-221)
    a, c = case["alpha"], case["c"]
    ids_map = {i: i for i in range(n)}
    cpn_t2v = t2v["candidate_likelihood"] - a[0] * t2v["candidate_prior"]
    cpn_v2t = v2t["candidate_likelihood"] - a[1] * v2t["candidate_prior"]
    blim_t2v = c[0] * t2v["query_likelihood"] + (1 - c[0]) * cpn_t2v
    blim_v2t = c[1] * v2t["query_likelihood"] + (1 - c[1]) * cpn_v2t
    blim_t2v = c[2] * blim_t2v + (1 - c[2]) * t2v["internvideo2"]
    blim_v2t = c[3] * blim_v2t + (1 - c[3]) * v2t["internvideo2"]
    res_blim = tu.get_recall(blim_t2v, blim_v2t, ids_map, ids_map)
    res_cpn = tu.get_recall(cpn_t2v, cpn_v2t, ids_map, ids_map)
    res_cand = tu.get_recall(t2v["candidate_likelihood"], v2t["candidate_likelihood"], ids_map, ids_map)
    # zero-shot arithmetic (float64 zeros branch, training_utils.py:154,161-162)
    zs_t2v = c[0] * t2v["query_likelihood"] + (1 - c[0]) * np.zeros((n, n))
    zs_t2v = c[2] * zs_t2v + (1 - c[2]) * t2v["internvideo2"]
    zs_v2t = c[3] * cpn_v2t + (1 - c[3]) * v2t["internvideo2"]
    res_zs = tu.get_recall(zs_t2v, zs_v2t, ids_map, ids_map)
    out.update(blim_t2v=blim_t2v, blim_v2t=blim_v2t, cpn_t2v=cpn_t2v, cpn_v2t=cpn_v2t, zs_t2v=zs_t2v, zs_v2t=zs_v2t)
    keys = sorted(res_blim)
    out["recall_keys"] = np.array(keys)
    out["recall_blim"] = np.array([res_blim[k] for k in keys])
    out["recall_cpn"] = np.array([res_cpn[k] for k in keys])
    out["recall_cand"] = np.array([res_cand[k] for k in keys])
    out["recall_zs"] = np.array([res_zs[k] for k in keys])
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", nargs="*", default=list(CASES))
    a = ap.parse_args()
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    for name in a.cases:
        out = run_reference(CASES[name])
        path = os.path.join(ROOT, "tests", "golden", f"{name}.npz")
        np.savez_compressed(path, **out)
        print(name, "->", path, {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items() if k.startswith(("v2t", "t2v"))})


if __name__ == "__main__":
    main()
