"""Seeded synthetic weights and corpora shaped like the reference's inputs: no tokenizer, dataset or
checkpoint is needed.  The tensors follow the output contract of BaseDataset (reference dataloader/base_dataset.py:60-114):
token ids with one -200 image sentinel, labels with -100 on the prompt part, [n_clips, 64, mm_hidden] features,
video_vocab = features.mean(1), tvg_video_labels = index of the video, InternVideo2 score matrices."""
from dataclasses import dataclass, field
from typing import List

import torch

from .engine import ModelConfig

IGNORE_INDEX = -100
IMAGE_TOKEN_INDEX = -200

# caption-length statistics (tokens) and clip counts per dataset shape [external, approximate: SURVEY.md 8(d)]
DATASET_SHAPES = {
    "msrvtt": dict(n=1000, cap_mean=12, cap_std=4, n_clips=4),
    "lsmdc": dict(n=1000, cap_mean=15, cap_std=6, n_clips=4),
    "didemo": dict(n=1004, cap_mean=45, cap_std=15, n_clips=4),
    "activitynet": dict(n=4917, cap_mean=80, cap_std=30, n_clips=16),
}

# ids of the Qwen2 chat template pieces (listed in the reference at modeling_videochat_flash.py:408)
_HEADER = [151644, 8948, 198, 2610, 525, 264, 10950, 17847, 13, 151645, 198, 151644, 872, 198]          # system + user header (14)
_TVG_INSTR = [31115, 264, 2766, 2661, 279, 17256, 624]                                                  # "Generate a video given the caption." (7)
_VTG_TAIL = [198, 74785, 419, 2766, 26753, 13, 151645, 198, 151644, 77091, 198, 198]                    # "\n{instruction}<|im_end|>\n<|im_start|>assistant\n" (12)
_TVG_CAPTION = [45, 25]                                                                                 # "Caption", ":"
_IM_START, _IM_END, _ASSISTANT, _NL = 151644, 151645, 77091, 198


def _remap(ids, cfg: ModelConfig):
    """Small-vocabulary configs: fold template ids into the vocabulary, keeping <|im_end|> == cfg.image_token_id."""
    if cfg.vocab_size > 151645:
        return list(ids)
    out = []
    for t in ids:
        if t == _IM_END:
            out.append(cfg.image_token_id)
        elif t == _IM_START:
            out.append(cfg.image_token_id - 1)
        else:
            out.append(t % (cfg.image_token_id - 1000))
    return out


def param_shapes(cfg: ModelConfig):
    """Reference state_dict names -> shapes for the parameters on the scoring path."""
    H, I, V, MM = cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size, cfg.mm_hidden_size
    NQ, NKV = cfg.num_heads * cfg.head_dim, cfg.num_kv_heads * cfg.head_dim
    shapes = {"model.embed_tokens.weight": (V, H)}
    for i in range(cfg.num_layers):
        p = f"model.layers.{i}."
        shapes.update({
            p + "self_attn.q_proj.weight": (NQ, H), p + "self_attn.q_proj.bias": (NQ,),
            p + "self_attn.k_proj.weight": (NKV, H), p + "self_attn.k_proj.bias": (NKV,),
            p + "self_attn.v_proj.weight": (NKV, H), p + "self_attn.v_proj.bias": (NKV,),
            p + "self_attn.o_proj.weight": (H, NQ),
            p + "mlp.gate_proj.weight": (I, H), p + "mlp.up_proj.weight": (I, H), p + "mlp.down_proj.weight": (H, I),
            p + "input_layernorm.weight": (H,), p + "post_attention_layernorm.weight": (H,),
        })
    shapes["model.norm.weight"] = (H,)
    for mlp in ("mlp", "tvg_mlp"):
        p = f"model.mm_projector.{mlp}."
        shapes.update({p + "0.weight": (H, MM), p + "0.bias": (H,), p + "2.weight": (H, H), p + "2.bias": (H,)})
    shapes["lm_head.weight"] = (V, H)
    shapes["visual_head.weight"] = (MM, H)
    return shapes


def init_weight(cfg: ModelConfig, name, idx, seed=0, device="cpu", std=0.02, rich=False, dtype=torch.bfloat16):
    """One parameter of init_weights (generated from (seed, idx): tensors can be streamed one at a time)."""
    shape = param_shapes(cfg)[name]
    dev = torch.device(device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed * 100003 + idx)
    if name.endswith("norm.weight") or name.endswith("layernorm.weight"):
        t = torch.ones(shape, device=dev)
        if rich:
            t = t + 0.1 * torch.randn(shape, generator=g, device=dev)
    elif name.endswith(".bias"):
        t = torch.zeros(shape, device=dev)
        if rich:
            t = 0.5 * std * torch.randn(shape, generator=g, device=dev)
    else:
        t = torch.randn(shape, generator=g, device=dev, dtype=torch.float32 if dev.type == "cpu" else dtype) * std
    return t.to(dtype)


def init_weights(cfg: ModelConfig, seed=0, device="cpu", std=0.02, rich=False, dtype=torch.bfloat16):
    """Random parameters.  rich=False follows the reference's _init_weights (modeling_qwen2_flash.py:834-843): N(0, std)
    Linear / Embedding weights, zero biases, unit RMSNorm weights.  rich=True also randomises biases and norm weights so
    that every term of the path is exercised by parity tests.  Values are generated per tensor from (seed, index) in
    fp32 and rounded to `dtype`, so the engine and the oracle see identical, bf16-representable numbers."""
    return {name: init_weight(cfg, name, idx, seed, device, std, rich, dtype) for idx, name in enumerate(param_shapes(cfg))}


@dataclass
class SynthCorpus:
    cfg: ModelConfig
    name: str
    n: int
    n_clips: int
    video: torch.Tensor                 # [n, n_clips, 64, MM] bf16
    video_vocab: torch.Tensor           # [n, n_clips, MM] bf16
    tvg_video_labels: torch.Tensor      # [n] int64
    vtg_ids: List[torch.Tensor] = field(default_factory=list)
    vtg_labels: List[torch.Tensor] = field(default_factory=list)
    tvg_ids: List[torch.Tensor] = field(default_factory=list)
    tvg_labels: List[torch.Tensor] = field(default_factory=list)
    t2v_iv2: torch.Tensor = None        # [n_texts, n_videos] fp32
    v2t_iv2: torch.Tensor = None        # [n_videos, n_texts] fp32
    tvg_prefix_length: int = 21
    pad_token_id: int = 0

    def masks(self, ids_list):
        return [torch.ones_like(x) for x in ids_list]


def make_corpus(cfg: ModelConfig, name="msrvtt", n=None, n_clips=None, cap_mean=None, cap_std=None, seed=1, feat_device="cpu") -> SynthCorpus:
    shp = dict(DATASET_SHAPES[name])
    n = n or shp["n"]
    n_clips = n_clips or shp["n_clips"]
    cap_mean = cap_mean or shp["cap_mean"]
    cap_std = shp["cap_std"] if cap_std is None else cap_std
    dev = torch.device(feat_device)
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    video = (torch.randn(n, n_clips, cfg.tokens_per_clip, cfg.mm_hidden_size, generator=g, device=dev) * 0.5).to(torch.bfloat16)
    video_vocab = video.float().mean(2).to(torch.bfloat16)

    gt = torch.Generator()
    gt.manual_seed(seed + 1)
    lens = (torch.randn(n, generator=gt) * cap_std + cap_mean).round().clamp(min=2, max=max(4, cap_mean + 4 * max(cap_std, 1))).long()
    lo, hi = (1000, cfg.vocab_size - 2000) if cfg.vocab_size > 151645 else (100, cfg.image_token_id - 1000)
    header, instr, tail, capt = _remap(_HEADER, cfg), _remap(_TVG_INSTR, cfg), _remap(_VTG_TAIL, cfg), _remap(_TVG_CAPTION, cfg)
    im_start, im_end, assistant, nl = _remap([_IM_START, _IM_END, _ASSISTANT, _NL], cfg)
    corpus = SynthCorpus(cfg=cfg, name=name, n=n, n_clips=n_clips, video=video, video_vocab=video_vocab,
                         tvg_video_labels=torch.arange(n), tvg_prefix_length=len(header) + len(instr))
    for i in range(n):
        cap = torch.randint(lo, hi, (int(lens[i]),), generator=gt).tolist()
        # VTG: header <image> tail | caption <|im_end|> \n        labels: prompt part = -100 (base_dataset.py:80-81)
        prompt = header + [IMAGE_TOKEN_INDEX] + tail
        ids = prompt + cap + [im_end, nl]
        lab = [IGNORE_INDEX] * len(prompt) + cap + [im_end, nl]
        corpus.vtg_ids.append(torch.tensor(ids, dtype=torch.long))
        corpus.vtg_labels.append(torch.tensor(lab, dtype=torch.long))
        # TVG: header instr "Caption:" caption <|im_end|>\n<|im_start|>assistant\n | <image> <|im_end|> \n  (base_dataset.py:86-105)
        prompt = header + instr + capt + cap + [im_end, nl, im_start, assistant, nl]
        ids = prompt + [IMAGE_TOKEN_INDEX, im_end, nl]
        lab = [IGNORE_INDEX] * len(prompt) + [IMAGE_TOKEN_INDEX, im_end, nl]
        corpus.tvg_ids.append(torch.tensor(ids, dtype=torch.long))
        corpus.tvg_labels.append(torch.tensor(lab, dtype=torch.long))

    gs = torch.Generator()
    gs.manual_seed(seed + 2)
    t2v = torch.randn(n, n, generator=gs) + 3.0 * torch.eye(n)
    v2t = t2v.t().contiguous() + 0.1 * torch.randn(n, n, generator=gs)
    corpus.t2v_iv2, corpus.v2t_iv2 = t2v.float(), v2t.float()
    return corpus
