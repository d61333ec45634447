"""Fine-tuned BLiM checkpoints -> plain reference parameter names.

The reference fine-tunes LoRA adapters (r = args.lora_r, alpha = args.lora_alpha) on q/k/v/o_proj and lm_head of the LLM
and on the two Linear layers ("0", "2") of mm_projector.mlp / tvg_mlp, plus the full visual_head (reference
main.py:99-111), and its checkpoints hold ONLY the trainable tensors (util/misc.py:276-297).  At evaluation time the
adapters are a constant: W' = W + (alpha / r) * B @ A.  `merge_lora` folds them into the base weights so the engine
loads an ordinary state dict (blim_load_weight knows nothing about adapters).  Pure load-time weight algebra.
"""
import re

import torch

_LORA_A = re.compile(r"^(.*)\.lora_A(?:\.[^.]+)?\.weight$")


def _plain_name(name):
    """PEFT wrapper prefixes -> the reference's own state_dict key."""
    prev = None
    while prev != name:
        prev = name
        if name.startswith("base_model.model."):
            name = name[len("base_model.model."):]
        name = name.replace(".base_model.model.", ".")
    name = name.replace(".base_layer.", ".")
    name = re.sub(r"\.modules_to_save\.[^.]+\.", ".", name)
    return name


_MLP = "model.mm_projector.mlp."
_TVG_MLP = "model.mm_projector.tvg_mlp."


def seed_tvg_mlp(state_dict):
    """The reference builds tvg_mlp as `copy.deepcopy(mm_projector.mlp)` AFTER the PEFT wrap (main.py:100-101): its frozen
    base weights and biases are the base mlp's, whatever the constructor's random tvg_mlp was
    (mm_projector_builder.py:91-93), and they are never saved (util/misc.py:282-285 stores requires_grad tensors only).
    Overwrites model.mm_projector.tvg_mlp.{0,2}.{weight,bias} with the mlp tensors, in place; returns the dict."""
    for key in [k for k in state_dict if k.startswith(_MLP)]:
        state_dict[_TVG_MLP + key[len(_MLP):]] = state_dict[key]
    return state_dict


def merge_lora(base_state_dict, trainable_state_dict, lora_r, lora_alpha, dtype=torch.float16, tvg_from_mlp=True):
    """Returns a new state dict: base weights with every LoRA pair merged (W + alpha/r * B @ A) and every other tensor of
    the checkpoint (visual_head.weight, ...) overriding the base tensor of the same plain name.  `tvg_from_mlp`: the
    tvg_mlp adapters sit on a copy of the base mlp (see seed_tvg_mlp), not on whatever tvg_mlp the base dict carries.
    `dtype` of the merged tensors: fp16 by default -- the engine's operand format (csrc/act_type.cuh) and the precision the
    reference itself runs the adapters in (main.py:97); bf16 would round away adapter deltas below 2^-8 of the weight."""
    scale = float(lora_alpha) / float(lora_r)
    out = {_plain_name(k): v for k, v in base_state_dict.items()}
    if tvg_from_mlp:
        seed_tvg_mlp(out)
    pending = dict(trainable_state_dict)
    merged = []
    for key in list(pending):
        m = _LORA_A.match(key)
        if not m:
            continue
        stem = m.group(1)
        b_key = key.replace(".lora_A", ".lora_B")
        if b_key not in pending:
            raise KeyError(f"LoRA A without its B: {key}")
        A, B = pending.pop(key), pending.pop(b_key)
        if A.shape[0] != lora_r or B.shape[1] != lora_r:
            raise ValueError(f"{key}: adapter rank {A.shape[0]} does not match lora_r={lora_r}")
        target = _plain_name(stem) + ".weight"
        if target not in out:
            raise KeyError(f"LoRA adapter for a weight the base model does not have: {target}")
        W = out[target]
        delta = (B.to(W.device, torch.float32) @ A.to(W.device, torch.float32)) * scale
        if delta.shape != W.shape:
            raise ValueError(f"{target}: adapter shape {tuple(delta.shape)} vs weight {tuple(W.shape)}")
        out[target] = (W.to(torch.float32) + delta).to(dtype)
        merged.append(target)
    for key, v in pending.items():
        out[_plain_name(key)] = v
    return out, merged


def load_finetuned(model, base_state_dict, checkpoint, lora_r, lora_alpha):
    """model: blim_b200.model.BlimModel.  checkpoint: the reference's `{'model': trainable tensors, ...}` dict or a path."""
    if isinstance(checkpoint, str):
        # reference checkpoints also hold the optimizer / scaler state and an argparse.Namespace (util/misc.py:288-294):
        # a trusted local file, so the full unpickler is what the reference itself uses (main.py:126)
        checkpoint = torch.load(checkpoint, map_location="cpu", weights_only=False)
    trainable = checkpoint["model"] if "model" in checkpoint else checkpoint
    sd, merged = merge_lora(base_state_dict, trainable, lora_r, lora_alpha)
    ignored = model.load_state_dict(sd)
    return merged, ignored
