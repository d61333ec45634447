"""Host-side stand-in for the reference model object, backed by the CUDA engine.

`BlimModel` offers the attributes retrieval_utils.py touches on `model` / `model.module` (reference
retrieval_utils.py:66,93,105,210): prepare_inputs_labels_for_multimodal, forward_visual, set_tvg_prefix_length and
__call__(inputs_embeds=, attention_mask=) returning an object with .logits / .hidden_states.  That literal path
materialises logits and exists for compatibility and testing; the fast path is blim_b200.retrieval, which hands whole
pair lists to Engine.score_pairs.
"""
from types import SimpleNamespace

import torch

from .engine import Engine, ModelConfig

IGNORE_INDEX = -100
IMAGE_TOKEN_INDEX = -200


def upload_videos(video, device, shard=None):
    """Host feature list / tensor -> ONE device tensor [n, n_clips, tokens, mm].  No host-side torch.stack: a list that is a
    run of consecutive views of a single host buffer (what a loader over a pre-stacked / pinned feature tensor yields) goes
    in one copy, anything else in one async copy per video.  With shard = (rank, world), world > 1, only videos
    [rank * per, (rank + 1) * per), per = ceil(n / world), are copied from the host; torch.distributed all-gathers the
    slices into the full tensor on the devices."""
    if torch.is_tensor(video):
        n, v0 = video.shape[0], video[0]
        get = lambda lo, hi: video[lo:hi]
        whole = True
    else:
        video = list(video)
        n, v0 = len(video), video[0]
        step = v0.numel() * v0.element_size()
        whole = (v0.device.type == "cpu" and v0.is_contiguous()
                 and all(v.dtype == v0.dtype and v.shape == v0.shape and v.is_contiguous() and v.data_ptr() == v0.data_ptr() + i * step
                         and v.untyped_storage().data_ptr() == v0.untyped_storage().data_ptr() for i, v in enumerate(video)))
        flat = torch.as_strided(v0, (n,) + tuple(v0.shape), (v0.numel(),) + tuple(v0.stride())) if whole else None
        get = (lambda lo, hi: flat[lo:hi]) if whole else None

    def copy_into(dst, lo, hi):           # dst[:hi - lo] <- videos lo..hi-1
        if hi <= lo:
            return
        if whole:
            dst[:hi - lo].copy_(get(lo, hi), non_blocking=True)
        else:
            for i in range(lo, hi):
                dst[i - lo].copy_(video[i], non_blocking=True)

    shape = tuple(v0.shape)
    rank, world = shard if shard else (0, 1)
    if world <= 1:
        feats = torch.empty((n,) + shape, dtype=v0.dtype, device=device)
        copy_into(feats, 0, n)
        return feats
    import torch.distributed as dist
    per = (n + world - 1) // world
    mine = torch.zeros((per,) + shape, dtype=v0.dtype, device=device)
    copy_into(mine, min(n, rank * per), min(n, (rank + 1) * per))
    full = torch.empty((per * world,) + shape, dtype=v0.dtype, device=device)
    dist.all_gather_into_tensor(full, mine)
    return full[:n]


class BlimModel:
    def __init__(self, cfg: ModelConfig, state_dict=None, device=0, **engine_kw):
        self.config = cfg
        self.engine = Engine(cfg, device=device, **engine_kw)
        self.device = self.engine.device
        self.module = self  # DDP-style access used by the reference (model.module....)
        self.tvg_prefix_length = 21
        self._corpus_keys = {}
        if state_dict is not None:
            self.load_state_dict(state_dict)

    # ------------------------------------------------------------------ nn.Module-ish surface
    def load_state_dict(self, state_dict, strict=False):
        return self.engine.load_state_dict(state_dict)

    def eval(self):
        return self

    def to(self, *a, **k):
        return self

    def half(self):
        return self

    def set_tvg_prefix_length(self, n):
        """modeling_videochat_flash.py:592"""
        self.tvg_prefix_length = int(n)
        self.engine.set_tvg_prefix_length(int(n))

    def set_video_vocab(self, video_vocab):
        """modeling_videochat_flash.py:589"""
        self.video_vocab = video_vocab

    def forward_visual(self, visual_token_embeds):
        """modeling_videochat_flash.py:598-599"""
        shp = visual_token_embeds.shape
        out = self.engine.forward_visual(visual_token_embeds.reshape(-1, shp[-1]))
        return out.reshape(*shp[:-1], -1).to(visual_token_embeds.dtype)

    def __call__(self, input_ids=None, attention_mask=None, position_ids=None, past_key_values=None, inputs_embeds=None, labels=None,
                 **unused):
        """VideoChatFlashQwenForCausalLM.forward as called on the scoring path (modeling_videochat_flash.py:601-629):
        model(inputs_embeds=..., attention_mask=...)."""
        if inputs_embeds is None:
            raise NotImplementedError("the scoring path always passes inputs_embeds (retrieval_utils.py:93-103)")
        logits, hidden = self.engine.forward_logits(inputs_embeds, attention_mask)
        return SimpleNamespace(logits=logits, hidden_states=hidden.to(inputs_embeds.dtype), loss=None, past_key_values=None, attentions=None)

    forward = __call__

    # ------------------------------------------------------------------ multimodal glue (compat path)
    def prepare_inputs_labels_for_multimodal(self, input_ids, position_ids, attention_mask, past_key_values, labels, images,
                                             modalities=["image"], image_sizes=None, video_feature=False, tvg=False, cpn=False):
        """video_feature=True branch of modeling_videochat_flash.py:185-515 with the projector and the embedding lookup
        running in the engine.  Returns the reference's 6-tuple."""
        assert video_feature, "only pre-extracted video features are on the scoring path"
        dev = self.device
        cfg = self.config
        B = input_ids.shape[0]
        feats = torch.stack([x.to(dev) for x in images], 0)                    # [B, n_clips, 64, MM]
        n_clips, tpc = feats.shape[1], feats.shape[2]
        proj = self.engine.project_video(feats.reshape(-1, feats.shape[-1]), tvg=tvg).reshape(B, n_clips, tpc, cfg.hidden_size)
        vis = proj.float().mean(2).to(torch.bfloat16) if tvg else proj.reshape(B, n_clips * tpc, cfg.hidden_size)
        seqs, labs, cpns = [], [], []
        for b in range(B):
            keep = attention_mask[b].bool()
            ids, lab = input_ids[b][keep], labels[b][keep]
            pos = (ids == IMAGE_TOKEN_INDEX).nonzero().flatten().tolist()
            assert len(pos) == 1
            i = pos[0]
            left, right = ids[:i], ids[i + 1:]
            emb = torch.cat([self.engine.embed_tokens(left), vis[b], self.engine.embed_tokens(right)], 0)
            nv = vis[b].shape[0]
            labs.append(torch.cat([lab[:i], torch.full((nv,), IGNORE_INDEX, dtype=lab.dtype, device=lab.device), lab[i + 1:]]))
            if tvg:
                m_left = torch.zeros(i, dtype=attention_mask.dtype, device=dev)
                m_left[:self.tvg_prefix_length] = 1
                m_vis = torch.ones(nv, dtype=attention_mask.dtype, device=dev)
            else:
                m_left = torch.ones(i, dtype=attention_mask.dtype, device=dev)
                m_vis = torch.zeros(nv, dtype=attention_mask.dtype, device=dev)
            cpns.append(torch.cat([m_left, m_vis, torch.ones(right.shape[0], dtype=attention_mask.dtype, device=dev)]))
            seqs.append(emb)
        L = max(s.shape[0] for s in seqs)
        embeds = torch.zeros(B, L, cfg.hidden_size, dtype=torch.bfloat16, device=dev)
        new_labels = torch.full((B, L), IGNORE_INDEX, dtype=labels.dtype, device=dev)
        mask = torch.zeros(B, L, dtype=attention_mask.dtype, device=dev)
        cpn_mask = torch.zeros(B, L, dtype=attention_mask.dtype, device=dev)
        for b in range(B):
            n = seqs[b].shape[0]
            embeds[b, :n], new_labels[b, :n], mask[b, :n], cpn_mask[b, :n] = seqs[b], labs[b].to(dev), 1, cpns[b]
        if cpn:
            return None, None, (mask, cpn_mask), past_key_values, embeds, new_labels
        return None, None, mask, past_key_values, embeds, new_labels

    # ------------------------------------------------------------------ corpus registration for the fast path
    # The caches below avoid re-uploading a corpus for each of the six compute_*_scores_x calls of one evaluation.  They
    # are keyed on the CONTENT location (storage pointers of the first / last tensor, shapes, in-place version counters),
    # not on id(): a new list of the same length can reuse the id of a garbage-collected one.
    @staticmethod
    def _tensors_key(x):
        if torch.is_tensor(x):
            return ("t", x.data_ptr(), tuple(x.shape), str(x.dtype), x._version)
        x = list(x)
        ends = [t for t in (x[0], x[-1])] if x else []
        return ("l", len(x)) + tuple((t.data_ptr(), tuple(t.shape), str(t.dtype), t._version) if torch.is_tensor(t) else id(t) for t in ends)

    def ensure_videos(self, video, shard=None):
        """Feature list / tensor -> engine.  shard = (rank, world) of an initialised process group: every rank holds the same
        host copy of the corpus (like every rank of the reference loads the whole loader, retrieval_utils.py:182-193), so each
        uploads only its 1/world slice over PCIe and the slices are all-gathered device to device (NVLink) instead of every
        rank pulling all 0.5 GB through its host link."""
        key = ("video", self._tensors_key(video))
        if self._corpus_keys.get("video") != key:
            feats = video if torch.is_tensor(video) and video.device.type == "cuda" else upload_videos(video, self.device, shard)
            self.engine.set_videos(feats)
            self._corpus_keys["video"] = key
            self._corpus_keys.pop("vocab", None)

    def ensure_texts(self, which, input_ids, attention_masks, labels):
        key = (which, self._tensors_key(input_ids), self._tensors_key(labels), self._tensors_key(attention_masks) if attention_masks is not None else None)
        if self._corpus_keys.get(("texts", which)) != key:
            if torch.is_tensor(input_ids):   # padded [N, Lmax] + masks (padding_ids output, retrieval_utils.py:155-167)
                m = attention_masks.bool().cpu()
                ids_l = [input_ids[i].cpu()[m[i]] for i in range(input_ids.shape[0])]
                lab_l = [labels[i].cpu()[m[i]] for i in range(labels.shape[0])]
            else:
                ids_l, lab_l = list(input_ids), list(labels)
            self.engine.set_texts(which, ids_l, lab_l)
            self._corpus_keys[("texts", which)] = key

    def ensure_vocab(self, video_vocab, tvg_video_labels):
        key = ("vocab", self._tensors_key(video_vocab), self._tensors_key(torch.as_tensor(tvg_video_labels)))
        if self._corpus_keys.get("vocab") != key:
            self.engine.set_video_vocab(video_vocab, torch.as_tensor(tvg_video_labels).cpu().numpy())
            self._corpus_keys["vocab"] = key
