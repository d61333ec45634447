"""TEST INFRASTRUCTURE -- CPU restatement (plain PyTorch fp32) of the reference's video feature extraction path.

Only tests/, __graft_entry__.smoke() and benchmark baselines may import this; the product path (blim_b200.vision ->
libblim_b200.so) never does.  Pinned: tests/golden/vision_tiny.npz is written by oracle/make_vision_golden.py from the
UNMODIFIED reference classes (UMTVisionTower / PretrainVisionTransformer / ToMe16_mlp_hd64) and checked in
tests/test_vision_cpu.py.

Follows, function by function:
    position_table      get_sinusoid_encoding_table / get_sinusoid_encoding_table2   vision_tower_builder.py:191-268
    vit_encode          UMTVisionTower.forward :564-577, PretrainVisionTransformer.forward :427-433,
                        PretrainVisionTransformerEncoder.forward_features :329-348, PatchEmbed :181-188, Block :152-159,
                        Attention.forward ('origin' branch == flash_v2 numerically) :99-126, Mlp :57-63
    bipartite_matching  bipartite_soft_matching   mm_projector_builder.py:6-58
    merge_tokens        ToMe16_mlp_hd64.merge_tokens + merge_wavg   mm_projector_builder.py:61-77, 101-130
    extract             encode_video_image(return_video_feature=True)   modeling_videochat_flash.py:152-154, 168 and
                        ToMe16_mlp_hd64.forward :134-154
"""
import numpy as np
import torch
import torch.nn.functional as F


def _sinusoid_np(n_position, d_hid):
    def vec(position):
        return [position / np.power(10000, 2 * (j // 2) / d_hid) for j in range(d_hid)]
    table = np.array([vec(p) for p in range(n_position)])
    table[:, 0::2] = np.sin(table[:, 0::2])
    table[:, 1::2] = np.cos(table[:, 1::2])
    return torch.tensor(table, dtype=torch.float).unsqueeze(0)


def position_table(image_size, patch_size, frames, hidden, ckpt_num_frame=4):
    n_position = frames * (image_size // patch_size) ** 2
    if image_size == 224:
        if ckpt_num_frame != -1 and ckpt_num_frame != frames:
            T, new_T = ckpt_num_frame, frames
            n_ck = n_position // new_T * T
            table = _sinusoid_np(n_ck, hidden)
            P = int((n_ck // T) ** 0.5)
            table = table.reshape(-1, T, P, P, hidden).permute(0, 2, 3, 4, 1).reshape(-1, hidden, T)
            table = F.interpolate(table, size=new_T, mode="linear")
            return table.reshape(1, P, P, hidden, new_T).permute(0, 4, 1, 2, 3).flatten(1, 3)[0]
        return _sinusoid_np(n_position, hidden)[0]
    table = _sinusoid_np(784, hidden)
    if n_position != 784:
        T, P = ckpt_num_frame, 14
        new_P = int((n_position // frames) ** 0.5)
        table = table.reshape(-1, T, P, P, hidden).reshape(-1, P, P, hidden).permute(0, 3, 1, 2)
        table = F.interpolate(table, size=(new_P, new_P), mode="bicubic", align_corners=False)
        table = table.permute(0, 2, 3, 1).reshape(-1, T, new_P, new_P, hidden).flatten(1, 3)
    if frames != ckpt_num_frame:
        T, new_T = ckpt_num_frame, frames
        P = int((n_position // frames) ** 0.5)
        table = table.reshape(-1, T, P, P, hidden).permute(0, 2, 3, 4, 1).reshape(-1, hidden, T)
        table = F.interpolate(table, size=new_T, mode="linear")
        table = table.reshape(1, P, P, hidden, new_T).permute(0, 4, 1, 2, 3).flatten(1, 3)
    return table[0]


def vit_encode(w, cfg, frames, n_layers=None):
    """frames [n_frames, 3, S, S] -> [n_clips, frames_per_clip * L, C] fp32; `w` = reference-named parameters (no prefix)."""
    w = {k: v.float() for k, v in w.items()}
    fpc, C, H = cfg.frames_per_clip, cfg.hidden_size, cfg.num_heads
    x = frames.float().reshape(-1, fpc, 3, cfg.image_size, cfg.image_size).permute(0, 2, 1, 3, 4)       # B C T H W  (:571)
    x = F.conv3d(x, w["encoder.patch_embed.proj.weight"], w["encoder.patch_embed.proj.bias"], stride=(1, cfg.patch_size, cfg.patch_size))
    x = x.flatten(2).transpose(1, 2)                                                                     # B, T*L, C   (:187)
    x = x + position_table(cfg.image_size, cfg.patch_size, fpc, C, cfg.ckpt_num_frame).to(x)
    B, N, _ = x.shape
    scale = (C // H) ** -0.5
    for i in range(cfg.num_layers if n_layers is None else n_layers):
        p = f"encoder.blocks.{i}."
        h = F.layer_norm(x, (C,), w[p + "norm1.weight"], w[p + "norm1.bias"], cfg.ln_eps)
        bias = torch.cat([w[p + "attn.q_bias"], torch.zeros_like(w[p + "attn.v_bias"]), w[p + "attn.v_bias"]])
        qkv = F.linear(h, w[p + "attn.qkv.weight"], bias).reshape(B, N, 3, H, -1).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0] * scale, qkv[1], qkv[2]
        a = (q @ k.transpose(-2, -1)).softmax(dim=-1)
        h = (a @ v).transpose(1, 2).reshape(B, N, -1)
        x = x + F.linear(h, w[p + "attn.proj.weight"], w[p + "attn.proj.bias"])
        h = F.layer_norm(x, (C,), w[p + "norm2.weight"], w[p + "norm2.bias"], cfg.ln_eps)
        h = F.linear(F.gelu(F.linear(h, w[p + "mlp.fc1.weight"], w[p + "mlp.fc1.bias"])), w[p + "mlp.fc2.weight"], w[p + "mlp.fc2.bias"])
        x = x + h
    return F.layer_norm(x, (C,), w["encoder.vision_layernorm.weight"], w["encoder.vision_layernorm.bias"], cfg.final_ln_eps)


def bipartite_matching(metric, r):
    """-> (unm_idx, src_idx, dst_idx, node_idx, edge_idx) for metric [b, t, c] (mm_projector_builder.py:19-34)."""
    t = metric.shape[1]
    r = min(r, t // 2)
    metric = metric / metric.norm(dim=-1, keepdim=True)
    a, b = metric[..., ::2, :], metric[..., 1::2, :]
    scores = a @ b.transpose(-1, -2)
    node_max, node_idx = scores.max(dim=-1)
    edge_idx = node_max.argsort(dim=-1, descending=True, stable=True)[..., None]
    unm_idx, src_idx = edge_idx[..., r:, :], edge_idx[..., :r, :]
    dst_idx = node_idx[..., None].gather(dim=-2, index=src_idx)
    return unm_idx, src_idx, dst_idx, node_idx, edge_idx[..., 0], r, node_max


def merge_tokens(x, target, heads, debug=False):
    """x [b, p, c] fp32 -> [b, target, c]; debug=True also returns the first round's (edge_idx, node_idx, node_max).
    Equal node_max values are ordered by index (stable sort); the reference leaves that order unspecified."""
    b, p, c = x.shape
    rs, tmp = [], p
    assert tmp > target
    while tmp != target:
        if tmp - target <= tmp // 2:
            rs.append(tmp - target)
            break
        rs.append(tmp // 2)
        tmp -= tmp // 2
    size, first = None, None
    for r in rs:
        metric = x.reshape(b, p, heads, c // heads).mean(2)
        unm_idx, src_idx, dst_idx, node_idx, edge_idx, r, node_max = bipartite_matching(metric, r)
        if first is None:
            first = (edge_idx.clone(), node_idx.clone(), node_max.clone())

        def merge(t):
            src, dst = t[..., ::2, :], t[..., 1::2, :]
            n, t1, cc = src.shape
            unm = src.gather(dim=-2, index=unm_idx.expand(n, t1 - r, cc))
            s = src.gather(dim=-2, index=src_idx.expand(n, r, cc))
            dst = dst.scatter_add(-2, dst_idx.expand(n, r, cc), s)
            return torch.cat([unm, dst], dim=1)

        if size is None:
            size = torch.ones_like(x[..., 0, None])
        x = merge(x * size)
        size = merge(size)
        x = x / size
        p = x.shape[1]
    return ((x,) + first) if debug else x


def extract(w, cfg, frames):
    """[n_frames, 3, S, S] -> [n_clips, tome_tokens_per_frame * frames_per_clip, C] fp32."""
    feats = vit_encode(w, cfg, frames)
    return merge_tokens(feats, cfg.tome_tokens_per_frame * cfg.frames_per_clip, cfg.num_heads)
