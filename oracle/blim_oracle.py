"""TEST INFRASTRUCTURE -- CPU/torch restatement of the reference's scoring path.  NOT part of the product.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module; the
product (blim_b200/) never does and fails loudly without its CUDA extension.

This is a plain-PyTorch (fp32 by default) restatement of the algorithm the reference executes for the path, pair by
pair and padded batch by padded batch exactly like the reference does -- no prefix reuse, no dedupe, full-vocabulary
logits -- so it is slow and simple.  Every function cites the reference lines it follows.  Parity pin: the fixtures in
tests/golden/ were produced by the UNMODIFIED reference (oracle/make_golden.py, run in the build container where
/root/reference exists) and tests/test_oracle_golden.py checks this restatement against them.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

IGNORE_INDEX = -100        # videochat_flash/conversation.py:10
IMAGE_TOKEN_INDEX = -200   # videochat_flash/conversation.py:11


# ------------------------------------------------------------------------------------------------ model pieces
def rms_norm(x, weight, eps):
    """Qwen2RMSNorm.forward, modeling_qwen2_flash.py:93-98."""
    dt = x.dtype
    xf = x.to(torch.float32)
    var = xf.pow(2).mean(-1, keepdim=True)
    xf = xf * torch.rsqrt(var + eps)
    return weight * xf.to(dt)


def rope_cos_sin(head_dim, theta, seq_len, dtype, device):
    """Qwen2RotaryEmbedding, modeling_qwen2_flash.py:109,119-135 (tables cached in fp32, cast to the activation dtype)."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))
    t = torch.arange(seq_len, dtype=torch.int64).type_as(inv_freq)
    freqs = torch.outer(t, inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos().to(device=device, dtype=dtype), emb.sin().to(device=device, dtype=dtype)


def _rot_half(x):
    """rotate_half, modeling_qwen2_flash.py:139-143."""
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def attention(p, prefix, cfg, h, mask4d, cos, sin):
    """Qwen2SdpaAttention.forward, modeling_qwen2_flash.py:640-716 (eager maths, identical semantics: q2:291-310)."""
    B, L, _ = h.shape
    nh, nkv, dh = cfg.num_heads, cfg.num_kv_heads, cfg.head_dim
    q = F.linear(h, p[prefix + "q_proj.weight"], p[prefix + "q_proj.bias"]).view(B, L, nh, dh).transpose(1, 2)
    k = F.linear(h, p[prefix + "k_proj.weight"], p[prefix + "k_proj.bias"]).view(B, L, nkv, dh).transpose(1, 2)
    v = F.linear(h, p[prefix + "v_proj.weight"], p[prefix + "v_proj.bias"]).view(B, L, nkv, dh).transpose(1, 2)
    c, s = cos[:L][None, None], sin[:L][None, None]          # position_ids = arange(L): q2:998-1003, apply_rotary_pos_emb q2:147-172
    q = q * c + _rot_half(q) * s
    k = k * c + _rot_half(k) * s
    rep = nh // nkv                                            # repeat_kv, q2:192-201
    k = k[:, :, None].expand(B, nkv, rep, L, dh).reshape(B, nh, L, dh)
    v = v[:, :, None].expand(B, nkv, rep, L, dh).reshape(B, nh, L, dh)
    w = torch.matmul(q, k.transpose(2, 3)) / math.sqrt(dh) + mask4d
    w = torch.softmax(w, dim=-1, dtype=torch.float32).to(q.dtype)
    o = torch.matmul(w, v).transpose(1, 2).reshape(B, L, nh * dh)
    return F.linear(o, p[prefix + "o_proj.weight"])


def causal_key_mask(attention_mask, L, dtype, device):
    """_prepare_4d_causal_attention_mask (q2:1019-1040): additive mask = causal AND key-valid."""
    neg = torch.finfo(dtype).min
    causal = torch.triu(torch.ones(L, L, dtype=torch.bool, device=device), diagonal=1)
    m = torch.zeros(attention_mask.shape[0], 1, L, L, dtype=dtype, device=device)
    m = m.masked_fill(causal[None, None], neg)
    m = m.masked_fill((attention_mask == 0)[:, None, None, :], neg)
    return m


def decoder_forward(p, cfg, inputs_embeds, attention_mask):
    """Qwen2Model_Flash.forward + Qwen2DecoderLayer.forward (q2:952-1156, 742-800) and the LM head (q2:1452-1453).
    Returns (logits fp32 [B,L,V], final-norm hidden states [B,L,H]) like the reference output object (q2:1472-1478)."""
    B, L, _ = inputs_embeds.shape
    dt, dev = inputs_embeds.dtype, inputs_embeds.device
    mask4d = causal_key_mask(attention_mask, L, dt, dev)
    cos, sin = rope_cos_sin(cfg.head_dim, cfg.rope_theta, L, dt, dev)
    h = inputs_embeds
    for i in range(cfg.num_layers):
        pre = f"model.layers.{i}."
        r = h
        h = rms_norm(h, p[pre + "input_layernorm.weight"], cfg.rms_norm_eps)
        h = r + attention(p, pre + "self_attn.", cfg, h, mask4d, cos, sin)
        r = h
        h = rms_norm(h, p[pre + "post_attention_layernorm.weight"], cfg.rms_norm_eps)
        gate = F.linear(h, p[pre + "mlp.gate_proj.weight"])
        up = F.linear(h, p[pre + "mlp.up_proj.weight"])
        h = r + F.linear(F.silu(gate) * up, p[pre + "mlp.down_proj.weight"])          # Qwen2MLP, q2:188
    h = rms_norm(h, p["model.norm.weight"], cfg.rms_norm_eps)
    logits = F.linear(h, p["lm_head.weight"]).float()
    return logits, h


def project_video(p, feats, tvg):
    """ToMe16_mlp_hd64.forward with video_feature=True (mm_projector_builder.py:156-159): Linear -> GELU -> Linear."""
    name = "model.mm_projector.tvg_mlp." if tvg else "model.mm_projector.mlp."
    x = F.linear(feats, p[name + "0.weight"], p[name + "0.bias"])
    x = F.gelu(x)
    return F.linear(x, p[name + "2.weight"], p[name + "2.bias"])


# ------------------------------------------------------------------------------------------------ multimodal glue
def prepare_inputs(p, cfg, input_ids, attention_mask, labels, videos, tvg, tvg_prefix_length):
    """prepare_inputs_labels_for_multimodal, video_feature=True branch (modeling_videochat_flash.py:185-515):
    strip padding by the mask (333-334), split at the image sentinel and embed the text (395-404), splice the projected
    visual rows (410-433; TVG rows are the per-clip mean, 243), build labels (-100 on visual rows, 429) and the CPN
    mask (VTG: zeros on the visual rows, 433; TVG: only the first tvg_prefix_length text tokens stay visible, 414-417),
    then right-pad to the batch maximum (470-485).  Returns (embeds [B,L,H], labels [B,L], mask [B,L], cpn_mask [B,L])."""
    embed_w = p["model.embed_tokens.weight"]
    seqs, labs, cpns = [], [], []
    for b in range(input_ids.shape[0]):
        keep = attention_mask[b].bool()
        ids, lab = input_ids[b][keep], labels[b][keep]
        pos = (ids == IMAGE_TOKEN_INDEX).nonzero().flatten().tolist()
        assert len(pos) == 1, "exactly one video per sequence on the scoring path"
        feat = project_video(p, videos[b].to(embed_w.dtype), tvg)          # [n_clips, 64, H]
        vis = feat.mean(1) if tvg else feat.flatten(0, 1)
        i = pos[0]
        left, right = ids[:i], ids[i + 1:]
        emb = torch.cat([F.embedding(left, embed_w), vis, F.embedding(right, embed_w)], dim=0)
        lab_new = torch.cat([lab[:i], torch.full((vis.shape[0],), IGNORE_INDEX, dtype=lab.dtype, device=lab.device), lab[i + 1:]])
        if tvg:
            m_left = torch.zeros(i, dtype=torch.long, device=ids.device)
            m_left[:tvg_prefix_length] = 1
            m_vis = torch.ones(vis.shape[0], dtype=torch.long, device=ids.device)
        else:
            m_left = torch.ones(i, dtype=torch.long, device=ids.device)
            m_vis = torch.zeros(vis.shape[0], dtype=torch.long, device=ids.device)
        cpn = torch.cat([m_left, m_vis, torch.ones(right.shape[0], dtype=torch.long, device=ids.device)])
        seqs.append(emb), labs.append(lab_new), cpns.append(cpn)
    L = max(s.shape[0] for s in seqs)
    B = len(seqs)
    embeds = torch.zeros(B, L, embed_w.shape[1], dtype=embed_w.dtype, device=embed_w.device)
    out_lab = torch.full((B, L), IGNORE_INDEX, dtype=torch.long, device=embed_w.device)
    mask = torch.zeros(B, L, dtype=torch.long, device=embed_w.device)
    cpn_mask = torch.zeros(B, L, dtype=torch.long, device=embed_w.device)
    for b in range(B):
        n = seqs[b].shape[0]
        embeds[b, :n], out_lab[b, :n], mask[b, :n], cpn_mask[b, :n] = seqs[b], labs[b], 1, cpns[b]
    return embeds, out_lab, mask, cpn_mask


# ------------------------------------------------------------------------------------------------ criteria
def vtg_criterion(logits, labels):
    """VTGCriterion.forward, retrieval_utils.py:23-33: shifted CE, sum / count of non-zero losses, negated."""
    sl = logits[..., :-1, :].contiguous()
    tl = labels[..., 1:].contiguous()
    loss = F.cross_entropy(sl.view(-1, sl.shape[-1]), tl.view(-1), reduction="none").reshape(logits.shape[0], -1)
    return -(loss.sum(1) / loss.bool().sum(1))


def tvg_criterion(logits, labels):
    """TVGCriterion.forward, retrieval_utils.py:40-43."""
    loss = F.cross_entropy(logits.reshape(-1, logits.shape[-1]), labels.reshape(-1), reduction="none").reshape(logits.shape[0], -1)
    return -loss.mean(1)


def score_batch(p, cfg, forward_type, cpn, ids, masks, labels, videos, video_vocab, vocab_labels, tvg_prefix_length, num_clips):
    """One batched forward + criterion of compute_*_scores_x (retrieval_utils.py:78-108 / 124-149)."""
    tvg = forward_type == "tvg"
    embeds, lab, mask, cpn_mask = prepare_inputs(p, cfg, ids, masks, labels, videos, tvg, tvg_prefix_length)
    logits, hidden = decoder_forward(p, cfg, embeds, cpn_mask if cpn else mask)
    if not tvg:
        return vtg_criterion(logits, lab)
    pcol = (lab == cfg.image_token_id).nonzero()[:, 1]
    idx = pcol[:, None] + (torch.arange(num_clips, device=pcol.device) - (num_clips + 1))[None]           # ru:99
    vis = torch.gather(hidden, 1, idx[..., None].expand(-1, -1, hidden.shape[-1]))                        # ru:104
    vis = F.linear(vis, p["visual_head.weight"])                                                          # forward_visual, mvf:598-599
    tl = torch.bmm(vis.permute(1, 0, 2), video_vocab.to(vis.dtype).permute(1, 2, 0)).transpose(0, 1) / math.sqrt(video_vocab.shape[-1])  # ru:106
    return tvg_criterion(tl.float(), vocab_labels)


def _pad_left(seqs, fill):
    """padding_ids, retrieval_utils.py:155-167 (LEFT padding to the global maximum)."""
    L = max(len(s) for s in seqs)
    out = torch.full((len(seqs), L), fill, dtype=torch.long)
    for i, s in enumerate(seqs):
        out[i, L - len(s):] = s
    return out


def compute_scores_x(p, cfg, corpus, direction, forward_type, cpn, topk, batch_size, rows=None, device="cpu"):
    """compute_v2t_scores_x / compute_t2v_scores_x (retrieval_utils.py:48-153): per row top-k candidates from the
    InternVideo2 scores, batched forwards, scatter into a -100-filled matrix (ru:219)."""
    ids_l, lab_l = (corpus.tvg_ids, corpus.tvg_labels) if forward_type == "tvg" else (corpus.vtg_ids, corpus.vtg_labels)
    ids = _pad_left(ids_l, corpus.pad_token_id)
    labels = _pad_left(lab_l, IGNORE_INDEX)
    masks = _pad_left([torch.ones_like(x) for x in ids_l], 0)
    sims = corpus.v2t_iv2 if direction == "v2t" else corpus.t2v_iv2
    n_rows, n_cols = sims.shape
    out = torch.full((n_rows, n_cols), -100.0)
    vocab = corpus.video_vocab.to(device)
    rows = range(n_rows) if rows is None else rows
    for r in rows:
        k = min(n_cols, topk)
        idx = sims[r].topk(k=k, dim=0).indices
        scores = []
        for j in range(0, k, batch_size):
            sel = idx[j:j + batch_size]
            n = len(sel)
            if direction == "v2t":           # the same video against n candidate texts (ru:55-89)
                b_ids, b_mask, b_lab = ids[sel], masks[sel], labels[sel]
                vids = [corpus.video[r].to(device)] * n
                vlab = corpus.tvg_video_labels[r].repeat(n, corpus.n_clips)
            else:                            # the same text against n candidate videos (ru:121-149)
                b_ids, b_mask, b_lab = ids[r].repeat(n, 1), masks[r].repeat(n, 1), labels[r].repeat(n, 1)
                vids = [corpus.video[int(v)].to(device) for v in sel]
                vlab = corpus.tvg_video_labels[sel][:, None].repeat(1, corpus.n_clips)
            s = score_batch(p, cfg, forward_type, cpn, b_ids.to(device), b_mask.to(device), b_lab.to(device), vids, vocab,
                            vlab.to(device), corpus.tvg_prefix_length, corpus.n_clips)
            scores.append(s.float().cpu())
        out[r, idx] = torch.cat(scores)
    return out


def evaluation(p, cfg, corpus, topk, batch_size, cpn=True, zero_shot=False, device="cpu"):
    """evaluation(), retrieval_utils.py:169-281 at world size 1: returns (t2v_dict, v2t_dict) of numpy fp32 matrices."""
    full = not zero_shot
    run = lambda d, ft, c: compute_scores_x(p, cfg, corpus, d, ft, c, topk, batch_size, device=device).numpy()
    t2v, v2t = {}, {}
    v2t["candidate_likelihood"] = run("v2t", "vtg", False)
    if cpn:
        v2t["candidate_prior"] = run("v2t", "vtg", True)
    if full:
        v2t["query_likelihood"] = run("v2t", "tvg", False)
    t2v["query_likelihood"] = run("t2v", "vtg", False)
    if full:
        t2v["candidate_likelihood"] = run("t2v", "tvg", False)
        if cpn:
            t2v["candidate_prior"] = run("t2v", "tvg", True)
    t2v["internvideo2"] = corpus.t2v_iv2.numpy()
    v2t["internvideo2"] = corpus.v2t_iv2.numpy()
    return t2v, v2t


# ------------------------------------------------------------------------------------------------ fuse + recall (numpy)
def get_recall(t2v, v2t):
    """get_recall, training_utils.py:173-221 with the diagonal ground truth of val_one_epoch (tu:146-147)."""
    def side(m):
        if np.count_nonzero(m == 0) != 0:          # "matrix absent" guard, tu:174,195
            return 0.0, 0.0, 0.0, None
        ranks = np.zeros(m.shape[0])
        for i, row in enumerate(m):
            order = np.argsort(row)[::-1]
            ranks[i] = np.where(order == i)[0][0]
        return (100.0 * np.sum(ranks < 1) / len(ranks), 100.0 * np.sum(ranks < 5) / len(ranks), 100.0 * np.sum(ranks < 10) / len(ranks), ranks)
    v1, v5, v10, vr = side(v2t)
    t1, t5, t10, tr = side(t2v)
    vm, tm = (v1 + v5 + v10) / 3, (t1 + t5 + t10) / 3
    res = {"t2v_r1": t1, "t2v_r5": t5, "t2v_r10": t10, "t2v_r_mean": tm, "v2t_r1": v1, "v2t_r5": v5, "v2t_r10": v10, "v2t_r_mean": vm,
           "r_mean": (vm + tm) / 2}
    return {k: round(float(v), 2) for k, v in res.items()}, tr, vr


def fuse(t2v_dict, v2t_dict, alpha, c, cpn=True, zero_shot=False):
    """The "blim" branch of val_one_epoch, training_utils.py:154-165 (numpy semantics preserved: Python-float
    coefficients, float32 matrices, float64 zeros in the zero-shot text->video branch)."""
    n_t, n_v = t2v_dict["internvideo2"].shape
    full = not zero_shot
    if cpn:
        cpn_t2v = t2v_dict["candidate_likelihood"] - alpha[0] * t2v_dict["candidate_prior"] if full else np.zeros((n_t, n_v))
        cpn_v2t = v2t_dict["candidate_likelihood"] - alpha[1] * v2t_dict["candidate_prior"]
    else:
        cpn_t2v = t2v_dict["candidate_likelihood"] if full else np.zeros((n_t, n_v))
        cpn_v2t = v2t_dict["candidate_likelihood"]
    blim_t2v = c[0] * t2v_dict["query_likelihood"] + (1 - c[0]) * cpn_t2v
    blim_v2t = c[1] * v2t_dict["query_likelihood"] + (1 - c[1]) * cpn_v2t if full else cpn_v2t
    blim_t2v = c[2] * blim_t2v + (1 - c[2]) * t2v_dict["internvideo2"]
    blim_v2t = c[3] * blim_v2t + (1 - c[3]) * v2t_dict["internvideo2"]
    return blim_t2v, blim_v2t, cpn_t2v, cpn_v2t


def val_results(t2v_dict, v2t_dict, alpha, c, cpn=True, zero_shot=False):
    """val_one_epoch's result table (tu:149-167) for the five named rows."""
    n = t2v_dict["internvideo2"].shape[0]
    z = np.zeros((n, n))
    blim_t2v, blim_v2t, cpn_t2v, cpn_v2t = fuse(t2v_dict, v2t_dict, alpha, c, cpn, zero_shot)
    res = {}
    for name in ("internvideo2", "candidate_likelihood", "query_likelihood"):
        res[name] = get_recall(t2v_dict.get(name, z), v2t_dict.get(name, z))[0]
    if cpn:
        res["cpn_candidate_likelihood"] = get_recall(cpn_t2v, cpn_v2t)[0]
    res["blim"] = get_recall(blim_t2v, blim_v2t)[0]
    return res
