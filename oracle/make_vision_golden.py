"""TEST INFRASTRUCTURE -- tests/golden/vision_tiny.npz: outputs of the UNMODIFIED reference vision classes on CPU (fp32).

Run in the build container only (needs /root/reference):   python oracle/make_vision_golden.py

What runs: UMTVisionTower.forward (vision_tower_builder.py:564-577) on a PretrainVisionTransformer of the tiny test
geometry -- build_vit() hard-codes ViT-L, so the module-level factory is pointed at the SAME class with small sizes; the
attention modules are switched to their own 'origin' branch (attn_type is a constructor argument; flash-attn has no CPU
kernels) -- then the three lines of encode_video_image that reshape clips (modeling_videochat_flash.py:152-154) and
ToMe16_mlp_hd64.forward(compress=True, local_num_frames=4, return_video_feature=True).  Inputs are regenerated from
seeds by blim_b200.vision.init_weights / the frame generator below.
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from blim_b200 import vision as V  # noqa: E402
from oracle import ref_harness  # noqa: E402

CASES = {"tiny_96": dict(cfg=V.VisionConfig.tiny(), wseed=0, fseed=1, n_frames=8),
         "tiny_224": dict(cfg=V.VisionConfig(image_size=224, hidden_size=128, encoder_depth=3, num_heads=2), wseed=2, fseed=3, n_frames=4)}


def make_frames(case):
    g = torch.Generator().manual_seed(case["fseed"])
    s = case["cfg"].image_size
    return torch.randn(case["n_frames"], 3, s, s, generator=g).to(torch.bfloat16).float()   # bf16-representable pixels


def make_weights(case):
    return {k: v.float() for k, v in V.init_weights(case["cfg"], seed=case["wseed"]).items()}


def run_reference(case):
    ref_harness._install_stubs()
    if ref_harness.REF_ROOT not in sys.path:
        sys.path.insert(0, ref_harness.REF_ROOT)
    from videochat_flash import vision_tower_builder as vtb
    from videochat_flash import mm_projector_builder as mpb
    cfg = case["cfg"]

    def small_vit(config, pt_type="origin"):
        return vtb.PretrainVisionTransformer(img_size=config.image_size, patch_size=cfg.patch_size, encoder_embed_dim=cfg.hidden_size,
                                             encoder_depth=cfg.encoder_depth, encoder_num_heads=cfg.num_heads, drop_path_rate=0.,
                                             num_frames=config.num_frames, tubelet_size=1, use_checkpoint=False, checkpoint_num=0,
                                             return_index=config.return_idx, with_ln=True)
    vtb.build_vit = small_vit
    tcfg = types.SimpleNamespace(mm_local_num_frames=cfg.frames_per_clip, mm_vision_select_layer=cfg.select_layer)
    tower = vtb.UMTVisionTower("umt-hd-tiny", tcfg, delay_load=False, image_size=cfg.image_size)
    for blk in tower.vision_tower.encoder.blocks:
        blk.attn.attn_type = "origin"
        blk.attn.attn_drop = torch.nn.Dropout(0.0)
    sd = {"vision_tower." + k: v for k, v in make_weights(case).items()}
    missing, unexpected = tower.load_state_dict(sd, strict=False)
    assert not unexpected and not missing, (missing, unexpected)
    tower = tower.float().eval()
    frames = make_frames(case)
    fpc = cfg.frames_per_clip
    pcfg = types.SimpleNamespace(mm_hidden_size=cfg.hidden_size, hidden_size=64, mm_pos_num_frames=8)
    vcfg = types.SimpleNamespace(image_size=cfg.image_size, patch_size=cfg.patch_size, num_attention_heads=cfg.num_heads)
    proj = mpb.ToMe16_mlp_hd64(pcfg, vcfg).float().eval()
    out = {}
    with torch.no_grad():
        video = frames                                                                              # one "video" of n_frames frames
        clips = video.reshape(video.shape[0] // fpc, fpc, video.shape[1], video.shape[2], video.shape[3])   # mvf:152
        feats = tower(clips)                                                                        # mvf:153  [n_clips, fpc * L, C]
        out["pos_embed"] = tower.vision_tower.encoder.pos_embed[0].numpy()
        out["encoded"] = feats.numpy()
        per_frame = feats.reshape(-1, feats.shape[-2] // fpc, feats.shape[-1])                      # mvf:154
        merged = proj(per_frame, compress=True, local_num_frames=fpc, return_video_feature=True)    # mvf:168
        out["merged"] = merged.numpy()
        # first-round matching decisions, read from the closure bipartite_soft_matching returns
        metric = feats.reshape(feats.shape[0], feats.shape[1], cfg.num_heads, -1).mean(2)
        r = feats.shape[1] // 2 if feats.shape[1] - merged.shape[1] > feats.shape[1] // 2 else feats.shape[1] - merged.shape[1]
        merge, _ = mpb.bipartite_soft_matching(metric, r)
        cells = dict(zip(merge.__code__.co_freevars, (c.cell_contents for c in merge.__closure__)))
        out["round1_r"] = np.array(r)
        out["round1_edge"] = torch.cat([cells["src_idx"], cells["unm_idx"]], dim=-2)[..., 0].numpy().astype(np.int32)
        out["round1_dst"] = cells["dst_idx"][..., 0].numpy().astype(np.int32)
    return out


if __name__ == "__main__":
    res = {}
    for name, case in CASES.items():
        for k, v in run_reference(case).items():
            if name == "tiny_224" and k == "pos_embed":
                v = v[::8]              # every 8th row keeps the fixture small (the 96-pixel case stores all rows)
            if name == "tiny_224" and k == "encoded":
                v = v[:, ::8]
            res[f"{name}/{k}"] = v
            print(name, k, getattr(v, "shape", v))
    path = os.path.join(ROOT, "tests", "golden", "vision_tiny.npz")
    np.savez_compressed(path, **res)
    print(path, os.path.getsize(path), "bytes")
