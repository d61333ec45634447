#!/usr/bin/env python
"""Where does the engine's distance from the fp32 reference come from?  (dev tool, GPU; imports oracle/ = test infrastructure)

Runs the oracle's fp32 per-pair forward with bf16 roundings inserted at exactly the places where the CUDA engine rounds
(bf16 tensor-core operands, fp32 accumulation, fp32 residual stream):

    proj   projector: GELU output and the projected visual rows are stored as bf16
    norm   RMSNorm output: the two roundings of modeling_qwen2_flash.py:93-98 (x*rstd -> bf16, weight*that -> bf16)
    norm1  RMSNorm output rounded once (fp32 product, one rounding) -- variant
    qkv    Q / K (after RoPE) / V stored as bf16
    p      softmax numerators P rounded to bf16 for P.V (row sum from the unrounded values)
    attn   attention output stored as bf16 (o_proj A operand)
    act    silu(gate)*up stored as bf16 (down_proj A operand)
    final  final-norm hidden state as bf16 (LM-head A operand)

and reports max / mean |d score| against the unrounded fp32 run for: all roundings (= an emulation of the engine),
each rounding alone, and all-but-one.  With --engine the real engine is scored on the same pairs too.
"""
import argparse
import json
import math
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import blim_oracle as O  # noqa: E402

ALL = ("proj", "norm", "qkv", "p", "attn", "act", "final")


ROUND_DTYPE = torch.bfloat16   # --fmt fp16: what the same pipeline gives with fp16 activations (weights stay as they are)


def rb(x):
    return x.to(ROUND_DTYPE).to(torch.float32)


def rms(x, w, eps, mode):
    xf = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps)
    if mode == "norm":
        return rb(w * rb(xf))
    if mode == "norm1":
        return rb(w * xf)
    return w * xf


def forward(p, cfg, embeds, mask, R):
    """decoder_forward of the oracle (fp32) with the roundings named in the set R."""
    B, L, _ = embeds.shape
    dev = embeds.device
    nh, nkv, dh = cfg.num_heads, cfg.num_kv_heads, cfg.head_dim
    mask4d = O.causal_key_mask(mask, L, torch.float32, dev)
    cos, sin = O.rope_cos_sin(dh, cfg.rope_theta, L, torch.float32, dev)
    c, s = cos[None, None], sin[None, None]
    nmode = "norm" if "norm" in R else ("norm1" if "norm1" in R else "")
    h = embeds
    for i in range(cfg.num_layers):
        pre = f"model.layers.{i}."
        a = pre + "self_attn."
        x = rms(h, p[pre + "input_layernorm.weight"], cfg.rms_norm_eps, nmode)
        q = F.linear(x, p[a + "q_proj.weight"], p[a + "q_proj.bias"]).view(B, L, nh, dh).transpose(1, 2)
        k = F.linear(x, p[a + "k_proj.weight"], p[a + "k_proj.bias"]).view(B, L, nkv, dh).transpose(1, 2)
        v = F.linear(x, p[a + "v_proj.weight"], p[a + "v_proj.bias"]).view(B, L, nkv, dh).transpose(1, 2)
        q = q * c + O._rot_half(q) * s
        k = k * c + O._rot_half(k) * s
        if "qkv" in R:
            q, k, v = rb(q), rb(k), rb(v)
        rep = nh // nkv
        k = k[:, :, None].expand(B, nkv, rep, L, dh).reshape(B, nh, L, dh)
        v = v[:, :, None].expand(B, nkv, rep, L, dh).reshape(B, nh, L, dh)
        w = torch.matmul(q, k.transpose(2, 3)) / math.sqrt(dh) + mask4d
        w = torch.exp(w - w.amax(-1, keepdim=True))
        l = w.sum(-1, keepdim=True)
        o = torch.matmul(rb(w) if "p" in R else w, v) / l
        o = o.transpose(1, 2).reshape(B, L, nh * dh)
        if "attn" in R:
            o = rb(o)
        h = h + F.linear(o, p[a + "o_proj.weight"])
        x = rms(h, p[pre + "post_attention_layernorm.weight"], cfg.rms_norm_eps, nmode)
        act = F.silu(F.linear(x, p[pre + "mlp.gate_proj.weight"])) * F.linear(x, p[pre + "mlp.up_proj.weight"])
        if "act" in R:
            act = rb(act)
        h = h + F.linear(act, p[pre + "mlp.down_proj.weight"])
    h = rms(h, p["model.norm.weight"], cfg.rms_norm_eps, "norm" if "final" in R else "")
    return F.linear(h, p["lm_head.weight"]), h


def project(p, feats, tvg, R):
    name = "model.mm_projector.tvg_mlp." if tvg else "model.mm_projector.mlp."
    x = F.gelu(F.linear(feats, p[name + "0.weight"], p[name + "0.bias"]))
    if "proj" in R:
        x = rb(x)
    x = F.linear(x, p[name + "2.weight"], p[name + "2.bias"])
    return rb(x) if "proj" in R else x


@torch.no_grad()
def vtg_scores(p, cfg, corpus, video_row, text_ids, R, cpn=False):
    """VTG scores of one video against the given texts (one padded batch, like retrieval_utils.py:62-97)."""
    dev = p["lm_head.weight"].device
    ids = O._pad_left([corpus.vtg_ids[t] for t in text_ids], corpus.pad_token_id).to(dev)
    lab = O._pad_left([corpus.vtg_labels[t] for t in text_ids], -100).to(dev)
    msk = O._pad_left([torch.ones_like(corpus.vtg_ids[t]) for t in text_ids], 0).to(dev)
    orig = O.project_video
    O.project_video = lambda pp, f, tvg: project(pp, f, tvg, R)
    try:
        embeds, lab2, mask, cpn_mask = O.prepare_inputs(p, cfg, ids, msk, lab, [corpus.video[video_row].to(dev)] * len(text_ids), False,
                                                        corpus.tvg_prefix_length)
    finally:
        O.project_video = orig
    logits, _ = forward(p, cfg, embeds.float(), cpn_mask if cpn else mask, R)
    return O.vtg_criterion(logits, lab2).cpu().numpy()


def budget(p, cfg, corpus, rows, topk, configs=None):
    """-> {config name: {max, mean}} of |score(config) - score(no rounding)| over rows x topk VTG pairs."""
    if configs is None:
        configs = {"all (engine emulation)": set(ALL), "all, norm rounded once": (set(ALL) - {"norm"}) | {"norm1"}}
        for r in ALL:
            configs[f"only {r}"] = {r}
        for r in ALL:
            configs[f"all but {r}"] = set(ALL) - {r}
    sel = [corpus.v2t_iv2[r].topk(k=topk).indices.tolist() for r in rows]
    base = np.concatenate([vtg_scores(p, cfg, corpus, r, s, set()) for r, s in zip(rows, sel)])
    out, raw = {}, {"fp32": base}
    for name, R in configs.items():
        got = np.concatenate([vtg_scores(p, cfg, corpus, r, s, R) for r, s in zip(rows, sel)])
        d = np.abs(got - base)
        out[name] = {"max": float(d.max()), "mean": float(d.mean())}
        raw[name] = got
    return out, raw, sel


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=2)
    ap.add_argument("--topk", type=int, default=16)
    ap.add_argument("--n", type=int, default=1000)
    ap.add_argument("--model", default="qwen2_7b")
    ap.add_argument("--rich", type=int, default=0)
    ap.add_argument("--engine", type=int, default=1)
    ap.add_argument("--out", default="")
    ap.add_argument("--fmt", default="bf16", choices=["bf16", "fp16"])
    ap.add_argument("--only-all", type=int, default=0)
    a = ap.parse_args()
    global ROUND_DTYPE
    ROUND_DTYPE = torch.float16 if a.fmt == "fp16" else torch.bfloat16
    from blim_b200 import synth
    from blim_b200.engine import VTG, ModelConfig
    cfg = ModelConfig.qwen2_7b() if a.model == "qwen2_7b" else ModelConfig.tiny()
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    w = synth.init_weights(cfg, seed=0, device=dev, std=0.02, rich=bool(a.rich))
    corpus = synth.make_corpus(cfg, "msrvtt", n=a.n, seed=1)
    rows = list(range(a.rows))
    eng_scores = None
    if a.engine and dev == "cuda":
        from blim_b200.model import BlimModel
        model = BlimModel(cfg, state_dict=w, device=0)
        e = model.engine
        e.set_videos(corpus.video)
        e.set_texts(0, corpus.vtg_ids, corpus.vtg_labels)
        sel = [corpus.v2t_iv2[r].topk(k=a.topk).indices.tolist() for r in rows]
        pv = np.repeat(rows, a.topk)
        pt = np.concatenate(sel)
        eng_scores = e.score_pairs(VTG, pv, pt).cpu().numpy()
        e.close()
    p = {k: v.float() for k, v in w.items()}
    del w
    rep, raw, sel = budget(p, cfg, corpus, rows, a.topk, {"all (engine emulation)": set(ALL)} if a.only_all else None)
    if eng_scores is not None:
        d = np.abs(eng_scores - raw["fp32"])
        rep["ENGINE"] = {"max": float(d.max()), "mean": float(d.mean())}
        d = np.abs(eng_scores - raw["all (engine emulation)"])
        rep["ENGINE vs emulation"] = {"max": float(d.max()), "mean": float(d.mean())}
    for k, v in rep.items():
        print(f"{k:32s} max {v['max']:.5f}  mean {v['mean']:.5f}")
    if a.out:
        json.dump({"pairs": a.rows * a.topk, "rich": a.rich, "fmt": a.fmt, "report": rep}, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
