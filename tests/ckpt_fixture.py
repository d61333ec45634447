"""A checkpoint with the key layout the reference writes (test helper).

main.py:99-111 wraps mm_projector.mlp in a PeftModel (targets "0", "2"), deep-copies the WRAPPED module into tvg_mlp,
then wraps the whole model (targets q/k/v/o_proj, lm_head) and makes visual_head trainable in fp32; util/misc.py:276-297
saves only the requires_grad tensors of that object under 'model', next to optimizer / scaler state and the argparse
namespace.  So the tensor names are PEFT's: `base_model.model.` in front of the outer model's names, the inner wrapper's
`base_model.model.` inside the projector, `lora_A.default.weight` / `lora_B.default.weight` per adapted Linear.  peft is
not installed offline, hence this hand-built layout.
"""
import argparse

import torch

from blim_b200 import synth


def adapted_linears(cfg):
    """PEFT stem (without .lora_X) -> (plain weight name, out_features, in_features)."""
    H, MM, V = cfg.hidden_size, cfg.mm_hidden_size, cfg.vocab_size
    NQ, NKV = cfg.num_heads * cfg.head_dim, cfg.num_kv_heads * cfg.head_dim
    out = {}
    for mlp in ("mlp", "tvg_mlp"):
        out[f"base_model.model.model.mm_projector.{mlp}.base_model.model.0"] = (f"model.mm_projector.{mlp}.0.weight", H, MM)
        out[f"base_model.model.model.mm_projector.{mlp}.base_model.model.2"] = (f"model.mm_projector.{mlp}.2.weight", H, H)
    for i in range(cfg.num_layers):
        for proj, o, n_in in (("q_proj", NQ, H), ("k_proj", NKV, H), ("v_proj", NKV, H), ("o_proj", H, NQ)):
            out[f"base_model.model.model.layers.{i}.self_attn.{proj}"] = (f"model.layers.{i}.self_attn.{proj}.weight", o, n_in)
    out["base_model.model.lm_head"] = ("lm_head.weight", V, H)
    return out


def make_reference_checkpoint(cfg, lora_r=8, lora_alpha=32, seed=11, scale=0.05):
    """-> (checkpoint dict as torch.save'd by the reference, {plain weight name: fp32 delta = alpha/r * B @ A})."""
    g = torch.Generator().manual_seed(seed)
    model, deltas = {}, {}
    for stem, (plain, o, n_in) in adapted_linears(cfg).items():
        A = torch.randn(lora_r, n_in, generator=g) * scale
        B = torch.randn(o, lora_r, generator=g) * scale
        model[stem + ".lora_A.default.weight"] = A.half()        # the model is .half() (main.py:97)
        model[stem + ".lora_B.default.weight"] = B.half()
        deltas[plain] = (lora_alpha / lora_r) * (B.half().float() @ A.half().float())
    model["base_model.model.visual_head.weight"] = torch.randn(cfg.mm_hidden_size, cfg.hidden_size, generator=g) * 0.05   # fp32 (main.py:108-111)
    ckpt = {"model": model, "optimizer": {"state": {}, "param_groups": []}, "epoch": 3, "scaler": {"scale": 65536.0},
            "args": argparse.Namespace(lora_r=lora_r, lora_alpha=lora_alpha, resume="", eval=False)}
    return ckpt, deltas


def base_state_dict(cfg, seed=0, std=0.05, with_random_tvg_mlp=True):
    """What `from_pretrained(...).state_dict()` gives: trained mlp, a tvg_mlp that is either absent (HF checkpoint keys) or
    the constructor's unrelated random init (mm_projector_builder.py:91-93)."""
    sd = synth.init_weights(cfg, seed=seed, std=std, rich=True)
    if not with_random_tvg_mlp:
        return {k: v for k, v in sd.items() if ".tvg_mlp." not in k}
    return sd   # synth draws tvg_mlp from its own (seed, index): unrelated to mlp, like the constructor's init


def expected_merged(cfg, base, ckpt, deltas):
    """fp32 parameters of the fine-tuned model, written out by hand (NOT through blim_b200.checkpoint)."""
    p = {k: v.float() for k, v in base.items()}
    for k in list(p):
        if k.startswith("model.mm_projector.mlp."):
            p["model.mm_projector.tvg_mlp." + k[len("model.mm_projector.mlp."):]] = p[k].clone()   # deepcopy of the wrapped mlp
    for plain, d in deltas.items():
        p[plain] = p[plain] + d
    p["visual_head.weight"] = ckpt["model"]["base_model.model.visual_head.weight"].float()
    return p
