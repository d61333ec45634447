"""Video feature extraction on the B200 extractor -- the host-side mirror of what the reference's
extract.py does with `model.encode_video_image(video, idx, return_video_feature=True)`:

    frames [T, 3, S, S] (T = 16, S = 448)  ->  clips of 4 frames  ->  UMT ViT-L encoder (UMTVisionTower.forward,
    vision_tower_builder.py:564-577)  ->  ToMe merge to 16 tokens per frame (ToMe16_mlp_hd64.forward with
    return_video_feature=True, mm_projector_builder.py:134-154)  ->  [T/4, 64, 1024] features, saved as fp16 `.pth`
    (extract.py:104-106) and read back by the dataloader (base_dataset.py:26-31).

All compute is in libblim_b200.so (include/blim_vision.h); this file holds the configuration, the position table (a
non-persistent buffer of the reference module, vision_tower_builder.py:191-268, built once on the host) and the ctypes
calls.  There is no PyTorch fallback.
"""
import ctypes
import math
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib

_DTYPE_CODE = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}
PREFIX = "model.vision_tower.vision_tower."
PROFILE_KINDS = ("gemm", "attention", "layernorm", "patchify_pos", "token_merging")


class VisionError(RuntimeError):
    pass


@dataclass
class VisionConfig:
    """build_vit() (vision_tower_builder.py:506-523) + UMTVisionConfig (:480-503) + the ToMe target."""
    image_size: int = 448
    patch_size: int = 16
    frames_per_clip: int = 4          # mm_local_num_frames
    hidden_size: int = 1024
    encoder_depth: int = 24
    select_layer: int = -2            # mm_vision_select_layer -> return_index: blocks run = depth + select_layer + 1
    num_heads: int = 16
    mlp_ratio: float = 4.0
    tome_tokens_per_frame: int = 16
    ckpt_num_frame: int = 4
    ln_eps: float = 1e-6
    final_ln_eps: float = 1e-12

    @property
    def num_layers(self):
        return self.encoder_depth + self.select_layer + 1

    @property
    def patches_per_frame(self):
        return (self.image_size // self.patch_size) ** 2

    @property
    def tokens_per_clip(self):
        return self.frames_per_clip * self.patches_per_frame

    @property
    def mlp_hidden_size(self):
        return int(self.hidden_size * self.mlp_ratio)

    @staticmethod
    def umt_l(image_size=448):
        return VisionConfig(image_size=image_size)

    @staticmethod
    def tiny():
        """2 heads of 64, 3 blocks run (depth 4, select -2), 96x96 frames -> 36 patches / frame, 144 tokens / clip."""
        return VisionConfig(image_size=96, hidden_size=128, encoder_depth=4, num_heads=2)


def _sinusoid(n_position, d_hid):
    pos = torch.arange(n_position, dtype=torch.float64)[:, None]
    j = torch.arange(d_hid, dtype=torch.float64)[None, :]
    angle = pos / torch.pow(torch.tensor(10000.0, dtype=torch.float64), 2 * torch.div(j, 2, rounding_mode="floor") / d_hid)
    table = angle.clone()
    table[:, 0::2] = torch.sin(angle[:, 0::2])
    table[:, 1::2] = torch.cos(angle[:, 1::2])
    return table.to(torch.float32).unsqueeze(0)   # float64 math then one cast, like the numpy reference


def position_table(cfg: VisionConfig):
    """`encoder.pos_embed` of the reference for (image_size, frames_per_clip): [tokens_per_clip, C] fp32.
    image_size == 224: get_sinusoid_encoding_table (vision_tower_builder.py:191-222, temporal interpolation only when the
    frame count differs from the checkpoint's); otherwise get_sinusoid_encoding_table2 (:225-268): the 4 x 14 x 14 table
    of the checkpoint, bicubically resized to the new patch grid, then linearly along time."""
    C, T, ck = cfg.hidden_size, cfg.frames_per_clip, cfg.ckpt_num_frame
    n_position = cfg.tokens_per_clip
    interp = torch.nn.functional.interpolate
    if cfg.image_size == 224:
        if ck != -1 and ck != T:
            n_ck = n_position // T * ck
            table = _sinusoid(n_ck, C)
            P = int((n_ck // ck) ** 0.5)
            table = table.reshape(-1, ck, P, P, C).permute(0, 2, 3, 4, 1).reshape(-1, C, ck)
            table = interp(table, size=T, mode="linear")
            table = table.reshape(1, P, P, C, T).permute(0, 4, 1, 2, 3).flatten(1, 3)
        else:
            table = _sinusoid(n_position, C)
        return table[0].contiguous()
    pre_n_position = 784
    table = _sinusoid(pre_n_position, C)
    if n_position != pre_n_position:
        P = 14
        new_P = int((n_position // T) ** 0.5)
        table = table.reshape(-1, ck, P, P, C).reshape(-1, P, P, C).permute(0, 3, 1, 2)
        table = interp(table, size=(new_P, new_P), mode="bicubic", align_corners=False)
        table = table.permute(0, 2, 3, 1).reshape(-1, ck, new_P, new_P, C).flatten(1, 3)
    if T != ck:
        P = int((n_position // T) ** 0.5)
        table = table.reshape(-1, ck, P, P, C).permute(0, 2, 3, 4, 1).reshape(-1, C, ck)
        table = interp(table, size=T, mode="linear")
        table = table.reshape(1, P, P, C, T).permute(0, 4, 1, 2, 3).flatten(1, 3)
    return table[0].contiguous()


def param_shapes(cfg: VisionConfig):
    """Reference parameter names (below model.vision_tower.vision_tower.) and shapes of the blocks that are run."""
    C, F, P = cfg.hidden_size, cfg.mlp_hidden_size, cfg.patch_size
    shapes = {"encoder.patch_embed.proj.weight": (C, 3, 1, P, P), "encoder.patch_embed.proj.bias": (C,)}
    for i in range(cfg.num_layers):
        b = f"encoder.blocks.{i}."
        shapes.update({b + "norm1.weight": (C,), b + "norm1.bias": (C,), b + "attn.q_bias": (C,), b + "attn.v_bias": (C,),
                       b + "attn.qkv.weight": (3 * C, C), b + "attn.proj.weight": (C, C), b + "attn.proj.bias": (C,),
                       b + "norm2.weight": (C,), b + "norm2.bias": (C,), b + "mlp.fc1.weight": (F, C), b + "mlp.fc1.bias": (F,),
                       b + "mlp.fc2.weight": (C, F), b + "mlp.fc2.bias": (C,)})
    shapes["encoder.vision_layernorm.weight"] = (C,)
    shapes["encoder.vision_layernorm.bias"] = (C,)
    return shapes


def init_weights(cfg: VisionConfig, seed=0, device="cpu", dtype=torch.bfloat16, rich=True):
    """Synthetic parameters: xavier-uniform Linear weights like the reference's _init_weights (vision_tower_builder.py:
    414-422); rich=True also randomises biases and LayerNorm parameters so every term is exercised.  Generated per tensor
    from (seed, index) in fp32 and rounded to `dtype`."""
    out = {}
    dev = torch.device(device)
    for idx, (name, shape) in enumerate(param_shapes(cfg).items()):
        g = torch.Generator(device=dev)
        g.manual_seed(seed * 100003 + idx)
        if name.endswith("weight") and len(shape) >= 2:
            fan_out, fan_in = shape[0], int(np.prod(shape[1:]))
            bound = math.sqrt(6.0 / (fan_in + fan_out))
            t = (torch.rand(shape, generator=g, device=dev) * 2 - 1) * bound
        elif "norm" in name and name.endswith("weight"):
            t = torch.ones(shape, device=dev) + (0.1 * torch.randn(shape, generator=g, device=dev) if rich else 0)
        else:
            t = 0.02 * torch.randn(shape, generator=g, device=dev) if rich else torch.zeros(shape, device=dev)
        out[name] = t.to(dtype)
    return out


class VisionEncoder:
    """One extractor per process / device (blim_vision_create ... blim_vision_destroy)."""

    def __init__(self, cfg: VisionConfig, state_dict=None, device=0, max_clips=16):
        self.cfg = cfg
        self.lib = _lib.load()
        self.h = None
        if not torch.cuda.is_available():
            raise VisionError("blim_b200.vision needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", device) if not isinstance(device, torch.device) else device
        c = _lib.VisionCfg(image_size=cfg.image_size, patch_size=cfg.patch_size, frames_per_clip=cfg.frames_per_clip, hidden_size=cfg.hidden_size,
                           num_layers=cfg.num_layers, num_heads=cfg.num_heads, mlp_hidden_size=cfg.mlp_hidden_size,
                           tome_tokens_per_frame=cfg.tome_tokens_per_frame, max_clips=max_clips, ln_eps=cfg.ln_eps, final_ln_eps=cfg.final_ln_eps)
        h = ctypes.c_void_p()
        if self.lib.blim_vision_create(ctypes.byref(c), self.device.index or 0, ctypes.byref(h)) != 0:
            raise VisionError(self.lib.blim_vision_last_error(None).decode())
        self.h = h
        self.max_clips = max_clips
        with torch.cuda.device(self.device):
            pos = position_table(cfg).to(self.device)
            self._check(self.lib.blim_vision_set_pos_embed(self.h, ctypes.c_void_p(pos.data_ptr()), pos.shape[0], self._stream()))
            torch.cuda.synchronize(self.device)
        if state_dict is not None:
            self.load_state_dict(state_dict)

    def _check(self, rc):
        if rc != 0:
            raise VisionError(self.lib.blim_vision_last_error(self.h).decode())

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def close(self):
        if getattr(self, "h", None):
            self.lib.blim_vision_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_state_dict(self, state_dict):
        """Reference names, with or without the `model.vision_tower.vision_tower.` prefix; other keys are skipped."""
        with torch.cuda.device(self.device):
            for name, t in state_dict.items():
                short = name[len(PREFIX):] if name.startswith(PREFIX) else name
                if not short.startswith("encoder."):
                    continue
                if t.dtype not in _DTYPE_CODE:
                    t = t.float()
                src = t.to(self.device).contiguous()
                shape = (ctypes.c_int64 * src.dim())(*src.shape)
                self._check(self.lib.blim_vision_load_weight(self.h, short.encode(), ctypes.c_void_p(src.data_ptr()), _DTYPE_CODE[src.dtype], shape,
                                                             src.dim(), self._stream()))
            torch.cuda.synchronize(self.device)

    def _frames(self, frames):
        assert frames.dim() == 4 and frames.shape[1] == 3 and frames.shape[2] == frames.shape[3] == self.cfg.image_size, frames.shape
        if frames.dtype not in _DTYPE_CODE:
            frames = frames.float()
        return frames.to(self.device).contiguous()

    def encode(self, frames):
        """UMTVisionTower.forward: [n_frames, 3, S, S] -> fp32 [n_clips, tokens_per_clip, C] (final-LayerNorm states)."""
        with torch.cuda.device(self.device):
            src = self._frames(frames)
            n_clips = src.shape[0] // self.cfg.frames_per_clip
            out = torch.empty((n_clips, self.cfg.tokens_per_clip, self.cfg.hidden_size), dtype=torch.float32, device=self.device)
            self._check(self.lib.blim_vision_encode(self.h, ctypes.c_void_p(src.data_ptr()), _DTYPE_CODE[src.dtype], src.shape[0],
                                                    ctypes.c_void_p(out.data_ptr()), self._stream()))
        return out

    def merge_tokens(self, x, target, debug=False):
        """ToMe16_mlp_hd64.merge_tokens: fp32 [b, p, C] -> [b, target, C] (+ first-round edge_idx / node_idx with debug)."""
        with torch.cuda.device(self.device):
            x = x.to(self.device, torch.float32).contiguous()
            b, p, C = x.shape
            assert C == self.cfg.hidden_size
            out = torch.empty((b, target, C), dtype=torch.float32, device=self.device)
            na = (p + 1) // 2
            edge = torch.empty((b, na), dtype=torch.int32, device=self.device) if debug else None
            nidx = torch.empty((b, na), dtype=torch.int32, device=self.device) if debug else None
            self._check(self.lib.blim_vision_merge_tokens(self.h, ctypes.c_void_p(x.data_ptr()), b, p, target, ctypes.c_void_p(out.data_ptr()),
                                                          ctypes.c_void_p(edge.data_ptr()) if debug else None,
                                                          ctypes.c_void_p(nidx.data_ptr()) if debug else None, self._stream()))
        return (out, edge, nidx) if debug else out

    def extract(self, frames, out_dtype=torch.float16):
        """encode_video_image(..., return_video_feature=True) for one batch of frames (n_frames <= 4 * max_clips):
        -> [n_clips, 16 * frames_per_clip, C] in out_dtype (the reference stores fp16, extract.py:104)."""
        with torch.cuda.device(self.device):
            src = self._frames(frames)
            n_clips = src.shape[0] // self.cfg.frames_per_clip
            out = torch.empty((n_clips, self.cfg.tome_tokens_per_frame * self.cfg.frames_per_clip, self.cfg.hidden_size), dtype=out_dtype,
                              device=self.device)
            self._check(self.lib.blim_vision_extract(self.h, ctypes.c_void_p(src.data_ptr()), _DTYPE_CODE[src.dtype], src.shape[0],
                                                     ctypes.c_void_p(out.data_ptr()), _DTYPE_CODE[out_dtype], self._stream()))
        return out

    def extract_videos(self, videos, out_dtype=torch.float16):
        """extract.py:98-106 for a list of videos ([T_i, 3, S, S] pixel tensors): clips of several videos share one launch
        sequence (up to max_clips clips); returns one [T_i / frames_per_clip, 64, C] tensor per video."""
        fpc = self.cfg.frames_per_clip
        clips = [v.shape[0] // fpc for v in videos]
        outs, i = [], 0
        while i < len(videos):
            j, n = i, 0
            while j < len(videos) and n + clips[j] <= self.max_clips:
                n += clips[j]
                j += 1
            if j == i:
                raise VisionError(f"video {i} has {clips[i]} clips, more than max_clips={self.max_clips}")
            feats = self.extract(torch.cat([v[:c * fpc].to(self.device) for v, c in zip(videos[i:j], clips[i:j])], 0), out_dtype)
            outs += list(torch.split(feats, clips[i:j], 0))
            i = j
        return outs

    # -- measurement
    def kernel_launches(self):
        return int(self.lib.blim_vision_kernel_launches(self.h))

    def gemm_flops(self):
        return float(self.lib.blim_vision_gemm_flops(self.h))

    def profile(self, enable=True):
        self._check(self.lib.blim_vision_profile(self.h, int(enable)))

    def profile_read(self):
        n = len(PROFILE_KINDS)
        ms, ln = (ctypes.c_double * n)(), (ctypes.c_int64 * n)()
        self._check(self.lib.blim_vision_profile_read(self.h, n, ms, ln))
        return {k: dict(ms=ms[i], launches=ln[i]) for i, k in enumerate(PROFILE_KINDS)}
