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


def merge_lora(base_state_dict, trainable_state_dict, lora_r, lora_alpha, dtype=torch.bfloat16):
    """Returns a new state dict: base weights with every LoRA pair merged (W + alpha/r * B @ A) and every other tensor of
    the checkpoint (visual_head.weight, ...) overriding the base tensor of the same plain name."""
    scale = float(lora_alpha) / float(lora_r)
    out = {_plain_name(k): v for k, v in base_state_dict.items()}
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
        checkpoint = torch.load(checkpoint, map_location="cpu")
    trainable = checkpoint["model"] if "model" in checkpoint else checkpoint
    sd, merged = merge_lora(base_state_dict, trainable, lora_r, lora_alpha)
    ignored = model.load_state_dict(sd)
    return merged, ignored
