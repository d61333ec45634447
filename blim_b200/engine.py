"""Python handle on the C-ABI engine.  PyTorch is used only for device memory, streams and dtype bookkeeping."""
import ctypes
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib

VTG, VTG_PRIOR, TVG, TVG_PRIOR = 0, 1, 2, 3
TEXTS_VTG, TEXTS_TVG = 0, 1
_DTYPE_CODE = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}


@dataclass
class ModelConfig:
    """Architecture of VideoChatFlashQwenForCausalLM (reference modeling_videochat_flash.py:572-590 + Qwen2Config)."""
    hidden_size: int = 3584
    num_layers: int = 28
    num_heads: int = 28
    num_kv_heads: int = 4
    intermediate_size: int = 18944
    vocab_size: int = 152064
    mm_hidden_size: int = 1024
    tokens_per_clip: int = 64
    rope_theta: float = 1000000.0
    rms_norm_eps: float = 1e-6
    max_positions: int = 4096
    image_token_id: int = 151645  # conversation.py:13 (remapped for small-vocabulary test configs)

    @property
    def head_dim(self):
        return self.hidden_size // self.num_heads

    @staticmethod
    def qwen2_7b():
        return ModelConfig()

    @staticmethod
    def tiny():
        """BASELINE.json configs[0]: 2 layers, hidden 256 (4 heads / 2 KV heads of 64, I=1024, V=4096)."""
        return ModelConfig(hidden_size=256, num_layers=2, num_heads=4, num_kv_heads=2, intermediate_size=1024,
                           vocab_size=4096, mm_hidden_size=1024, rope_theta=1000000.0, max_positions=2048, image_token_id=4000)


class EngineError(RuntimeError):
    pass


def rope_tables(head_dim, theta, n_positions, table_dtype=torch.float32):
    """cos/sin [n_positions, head_dim/2] fp32, built like Qwen2RotaryEmbedding._set_cos_sin_cache
    (reference modeling_qwen2_flash.py:109,119-127).  The reference casts the cached tables to the activation dtype
    (modeling_qwen2_flash.py:133-134); table_dtype=torch.bfloat16 reproduces that rounding, the default keeps the fp32
    tables (what the fp32 reference / oracle uses, and strictly closer to the exact rotation)."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))
    t = torch.arange(n_positions, dtype=torch.int64).type_as(inv_freq)
    freqs = torch.outer(t, inv_freq)
    cos, sin = freqs.cos(), freqs.sin()
    if table_dtype is not None and table_dtype != torch.float32:
        cos, sin = cos.to(table_dtype).float(), sin.to(table_dtype).float()
    return cos.contiguous(), sin.contiguous()


class Engine:
    """One engine per process / device (blim_create ... blim_destroy)."""

    def __init__(self, cfg: ModelConfig, device=0, max_run_tokens=0, max_prefix_tokens=0, gemm_cta_group=0):
        if not torch.cuda.is_available():
            raise EngineError("blim_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.cfg = cfg
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        c = _lib.ModelCfg(cfg.hidden_size, cfg.num_layers, cfg.num_heads, cfg.num_kv_heads, cfg.head_dim, cfg.intermediate_size,
                          cfg.vocab_size, cfg.mm_hidden_size, cfg.tokens_per_clip, cfg.max_positions, cfg.rms_norm_eps,
                          max_run_tokens, max_prefix_tokens, gemm_cta_group)
        h = ctypes.c_void_p()
        rc = self.lib.blim_create(ctypes.byref(c), self.device.index, ctypes.byref(h))
        if rc != 0:
            raise EngineError(self.lib.blim_last_error(None).decode())
        self.h = h
        self._keep = {}
        self.n_clips = 0
        # 16-bit format of the engine's tensor-core operands (fp16 by default, csrc/act_type.cuh)
        self.act_dtype = {1: torch.bfloat16, 2: torch.float16}[int(self.lib.blim_act_dtype())]

    # ---------------------------------------------------------------- helpers
    def _check(self, rc):
        if rc != 0:
            raise EngineError(self.lib.blim_last_error(self.h).decode())

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _dev(self, t, dtype=None):
        t = t.to(self.device)
        if dtype is not None:
            t = t.to(dtype)
        return t.contiguous()

    def close(self):
        if getattr(self, "h", None):
            torch.cuda.synchronize(self.device)
            self.lib.blim_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---------------------------------------------------------------- weights
    def load_weight(self, name, t):
        """One parameter by its reference state_dict name; returns False if the name is not on the scoring path."""
        if not torch.is_tensor(t) or t.dtype not in _DTYPE_CODE:
            return False
        with torch.cuda.device(self.device):
            src = self._dev(t)
            shape = (ctypes.c_int64 * src.dim())(*src.shape)
            rc = self.lib.blim_load_weight(self.h, name.encode(), ctypes.c_void_p(src.data_ptr()), _DTYPE_CODE[src.dtype], shape, src.dim(),
                                           self._stream())
            if rc == 2:
                return False
            self._check(rc)
            torch.cuda.current_stream(self.device).synchronize()  # src may be a temporary
        return True

    def set_rope(self, table_dtype=torch.float32):
        with torch.cuda.device(self.device):
            cos, sin = rope_tables(self.cfg.head_dim, self.cfg.rope_theta, self.cfg.max_positions, table_dtype)
            cos, sin = self._dev(cos), self._dev(sin)
            self._check(self.lib.blim_set_rope(self.h, ctypes.c_void_p(cos.data_ptr()), ctypes.c_void_p(sin.data_ptr()),
                                               self.cfg.max_positions, self._stream()))
            torch.cuda.synchronize(self.device)

    def load_state_dict(self, state_dict, rope_table_dtype=torch.float32):
        """Load parameters by their reference state_dict names; returns the list of ignored keys."""
        ignored = [name for name, t in state_dict.items() if not self.load_weight(name, t)]
        self.set_rope(rope_table_dtype)
        return ignored

    # ---------------------------------------------------------------- corpus
    def set_videos(self, feats):
        """feats: [n_videos, n_clips, tokens_per_clip, mm_hidden] tensor (any float dtype; converted to bf16 on device)."""
        if isinstance(feats, (list, tuple)):
            feats = torch.stack([f for f in feats], 0)
        assert feats.dim() == 4 and feats.shape[2] == self.cfg.tokens_per_clip and feats.shape[3] == self.cfg.mm_hidden_size, feats.shape
        if feats.dtype not in _DTYPE_CODE:
            feats = feats.float()
        with torch.cuda.device(self.device):
            src = self._dev(feats)
            self._check(self.lib.blim_set_videos(self.h, ctypes.c_void_p(src.data_ptr()), _DTYPE_CODE[src.dtype], feats.shape[0],
                                                 feats.shape[1], self._stream()))
            torch.cuda.synchronize(self.device)
        self.n_videos, self.n_clips = feats.shape[0], feats.shape[1]

    def set_texts(self, which, ids_list, labels_list, masks_list=None):
        """Ragged token ids (with the -200 image sentinel) and labels (-100 = ignore); with `masks_list` the entries whose
        mask is 0 (padding) are stripped first (what prepare_inputs_labels_for_multimodal does, mvf:333-334)."""
        lens = np.fromiter((len(x) for x in ids_list), dtype=np.int64, count=len(ids_list))
        if all(torch.is_tensor(x) for x in ids_list):
            ids = torch.cat(list(ids_list)).numpy().astype(np.int32)
            labels = torch.cat(list(labels_list)).numpy().astype(np.int32)
        else:
            ids = np.concatenate([np.asarray(x, dtype=np.int64) for x in ids_list]).astype(np.int32)
            labels = np.concatenate([np.asarray(x, dtype=np.int64) for x in labels_list]).astype(np.int32)
        if masks_list is not None:
            mask = (torch.cat(list(masks_list)).numpy() if torch.is_tensor(masks_list[0])
                    else np.concatenate([np.asarray(x) for x in masks_list])).astype(bool)
            if not mask.all():
                seg = np.repeat(np.arange(len(lens)), lens)
                lens = np.bincount(seg[mask], minlength=len(lens)).astype(np.int64)
                ids, labels = ids[mask], labels[mask]
        off = np.zeros(len(ids_list) + 1, dtype=np.int64)
        np.cumsum(lens, out=off[1:])
        # token counts per text for the multi-GPU load balancer (retrieval.balanced_owner_ranks)
        seg = np.repeat(np.arange(len(lens)), lens)
        if not hasattr(self, "text_lens"):
            self.text_lens = {}
        self.text_lens[which] = {"total": lens.astype(np.float64),
                                 "scored": np.bincount(seg, weights=(labels != -100), minlength=len(lens)).astype(np.float64)}
        ids, labels = np.ascontiguousarray(ids), np.ascontiguousarray(labels)
        assert len(ids) == len(labels) == off[-1]
        self._check(self.lib.blim_set_texts(self.h, which, ids.ctypes.data_as(ctypes.c_void_p), labels.ctypes.data_as(ctypes.c_void_p),
                                            off.ctypes.data_as(ctypes.c_void_p), len(ids_list)))

    def set_video_vocab(self, vocab, video_labels):
        """vocab [n_vocab, n_clips, mm_hidden]; video_labels[n_videos] = vocab row of each video (base_dataset.py:33-37,114)."""
        if vocab.dtype not in _DTYPE_CODE:
            vocab = vocab.float()
        labels = np.ascontiguousarray(np.asarray(video_labels, dtype=np.int32))
        with torch.cuda.device(self.device):
            src = self._dev(vocab)
            self._check(self.lib.blim_set_video_vocab(self.h, ctypes.c_void_p(src.data_ptr()), _DTYPE_CODE[src.dtype], vocab.shape[0],
                                                      labels.ctypes.data_as(ctypes.c_void_p), len(labels), self._stream()))
            torch.cuda.synchronize(self.device)

    def build_video_vocab(self, video_labels, n_vocab=None):
        """video_vocab on the device from the features of set_videos (base_dataset.py:33-37)."""
        labels = np.ascontiguousarray(np.asarray(video_labels, dtype=np.int32))
        n_vocab = int(n_vocab if n_vocab is not None else labels.max() + 1)
        with torch.cuda.device(self.device):
            self._check(self.lib.blim_build_video_vocab(self.h, labels.ctypes.data_as(ctypes.c_void_p), len(labels), n_vocab, self._stream()))

    def set_tvg_prefix_length(self, n):
        self._check(self.lib.blim_set_tvg_prefix_length(self.h, int(n)))
        self.tvg_prefix_length = int(n)

    # ---------------------------------------------------------------- scoring
    def score_pairs(self, kind, pair_v, pair_t, out=None):
        """Scores of (video, text) pairs -> fp32 device tensor [n_pairs]."""
        pv = np.ascontiguousarray(np.asarray(pair_v, dtype=np.int32))
        pt = np.ascontiguousarray(np.asarray(pair_t, dtype=np.int32))
        assert pv.shape == pt.shape and pv.ndim == 1
        with torch.cuda.device(self.device):
            if out is None:
                out = torch.empty(len(pv), dtype=torch.float32, device=self.device)
            self._check(self.lib.blim_score_pairs(self.h, kind, pv.ctypes.data_as(ctypes.c_void_p), pt.ctypes.data_as(ctypes.c_void_p),
                                                  len(pv), ctypes.c_void_p(out.data_ptr()), self._stream()))
        return out

    def forward_logits(self, inputs_embeds, attention_mask=None, want_logits=True, want_hidden=True):
        """Compat path: model(inputs_embeds=..., attention_mask=...) -> (logits fp32 [B,L,V], hidden bf16 [B,L,H])."""
        B, L, H = inputs_embeds.shape
        with torch.cuda.device(self.device):
            emb = self._dev(inputs_embeds, torch.bfloat16)
            mask = self._dev(attention_mask, torch.int32) if attention_mask is not None else None
            logits = torch.empty(B, L, self.cfg.vocab_size, dtype=torch.float32, device=self.device) if want_logits else None
            hidden = torch.empty(B, L, H, dtype=torch.bfloat16, device=self.device) if want_hidden else None
            self._check(self.lib.blim_forward_logits(self.h, ctypes.c_void_p(emb.data_ptr()),
                                                     ctypes.c_void_p(mask.data_ptr()) if mask is not None else None, B, L,
                                                     ctypes.c_void_p(logits.data_ptr()) if want_logits else None,
                                                     ctypes.c_void_p(hidden.data_ptr()) if want_hidden else None, self._stream()))
        return logits, hidden

    def project_video(self, feats_rows, tvg=False):
        """mm_projector.mlp / tvg_mlp on [n_rows, mm_hidden] -> [n_rows, hidden] bf16."""
        with torch.cuda.device(self.device):
            x = self._dev(feats_rows, torch.bfloat16)
            out = torch.empty(x.shape[0], self.cfg.hidden_size, dtype=torch.bfloat16, device=self.device)
            self._check(self.lib.blim_project_video(self.h, ctypes.c_void_p(x.data_ptr()), x.shape[0], int(bool(tvg)),
                                                    ctypes.c_void_p(out.data_ptr()), self._stream()))
        return out

    def forward_visual(self, hidden_rows):
        with torch.cuda.device(self.device):
            x = self._dev(hidden_rows, torch.bfloat16)
            out = torch.empty(x.shape[0], self.cfg.mm_hidden_size, dtype=torch.bfloat16, device=self.device)
            self._check(self.lib.blim_forward_visual(self.h, ctypes.c_void_p(x.data_ptr()), x.shape[0], ctypes.c_void_p(out.data_ptr()),
                                                     self._stream()))
        return out

    def embed_tokens(self, ids):
        with torch.cuda.device(self.device):
            i = self._dev(ids.reshape(-1), torch.int32)
            out = torch.empty(i.numel(), self.cfg.hidden_size, dtype=torch.bfloat16, device=self.device)
            self._check(self.lib.blim_embed_tokens(self.h, ctypes.c_void_p(i.data_ptr()), i.numel(), ctypes.c_void_p(out.data_ptr()),
                                                   self._stream()))
        return out.reshape(*ids.shape, self.cfg.hidden_size)

    # ---------------------------------------------------------------- multi-GPU exchange
    comm_world = 1

    def comm_init(self, group=None):
        """Join this engine's NCCL communicator (one engine per rank): rank 0 creates the unique id, torch.distributed only
        carries its 128 bytes to the other ranks; the all-gather itself runs in the engine on the compute stream."""
        import torch.distributed as dist
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        if self.comm_world == world and world > 1:
            return
        box = [None]
        if rank == 0:
            buf = ctypes.create_string_buffer(128)
            if self.lib.blim_comm_unique_id(buf) != 0:
                raise EngineError(self.lib.blim_last_error(None).decode())
            box[0] = bytes(buf.raw)
        dist.broadcast_object_list(box, src=0, group=group)
        with torch.cuda.device(self.device):
            self._check(self.lib.blim_comm_init(self.h, ctypes.c_char_p(box[0]), rank, world))
        self.comm_world = world

    def allgather_scores(self, send, recv=None):
        """recv[r * n : (r + 1) * n] = rank r's send (fp32 device tensors), on the current stream."""
        n = send.numel()
        with torch.cuda.device(self.device):
            if recv is None:
                recv = torch.empty(self.comm_world * n, dtype=torch.float32, device=self.device)
            self._check(self.lib.blim_allgather_scores(self.h, None, ctypes.c_void_p(send.data_ptr()), ctypes.c_void_p(recv.data_ptr()), n,
                                                       self._stream()))
        return recv

    # ---------------------------------------------------------------- fuse / rerank
    def fuse_rerank(self, cand_idx, cand, prior, query, iv2, alpha, c_query, c_ens, use_prior=True, use_query=True,
                    cpn_zero_f64=False, row0=0):
        """CPN + ensemble + rerank of one direction on compact [rows, k] arrays (see include/blim_b200.h)."""
        with torch.cuda.device(self.device):
            n_rows, k = cand_idx.shape
            n_cols = iv2.shape[1]
            ci = self._dev(cand_idx, torch.int32)
            ca = self._dev(cand, torch.float32) if cand is not None else None
            pr = self._dev(prior, torch.float32) if prior is not None else None
            qu = self._dev(query, torch.float32) if query is not None else None
            iv = self._dev(iv2, torch.float32)
            fused = torch.empty(n_rows, k, dtype=torch.float64, device=self.device)
            order = torch.empty(n_rows, k, dtype=torch.int32, device=self.device)
            rank = torch.empty(n_rows, dtype=torch.int32, device=self.device)
            zero = torch.zeros(1, dtype=torch.int32, device=self.device)
            cfg = _lib.FuseCfg(float(alpha), float(c_query), float(c_ens), int(use_prior and prior is not None), int(use_query),
                               int(cpn_zero_f64))
            p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
            self._check(self.lib.blim_fuse_rerank(self.h, ctypes.byref(cfg), p(ci), p(ca), p(pr), p(qu), p(iv), n_rows, n_cols, k, row0,
                                                  p(fused), p(order), p(rank), p(zero), self._stream()))
        return fused, order, rank, zero

    def rank_dense(self, mat, row0=0):
        with torch.cuda.device(self.device):
            m = self._dev(mat, torch.float32)
            rank = torch.empty(m.shape[0], dtype=torch.int32, device=self.device)
            zero = torch.zeros(1, dtype=torch.int32, device=self.device)
            self._check(self.lib.blim_rank_dense(self.h, ctypes.c_void_p(m.data_ptr()), m.shape[0], m.shape[1], row0,
                                                 ctypes.c_void_p(rank.data_ptr()), ctypes.c_void_p(zero.data_ptr()), self._stream()))
        return rank, zero

    def topk_rows(self, mat, k):
        """Per-row top-k of a dense fp32 matrix on the device -> (idx int64 [rows, k], val fp32 [rows, k]), descending."""
        with torch.cuda.device(self.device):
            m = self._dev(mat, torch.float32)
            k = min(int(k), m.shape[1])
            idx = torch.empty(m.shape[0], k, dtype=torch.int32, device=self.device)
            val = torch.empty(m.shape[0], k, dtype=torch.float32, device=self.device)
            self._check(self.lib.blim_topk_rows(self.h, ctypes.c_void_p(m.data_ptr()), m.shape[0], m.shape[1], k,
                                                ctypes.c_void_p(idx.data_ptr()), ctypes.c_void_p(val.data_ptr()), self._stream()))
        return idx.long(), val

    def scatter_scores(self, n_rows, n_cols, row, col, val, fill=-100.0, dense=None):
        """torch.full((n_rows, n_cols), fill) then dense[row, col] = val  (reference retrieval_utils.py:219,110)."""
        with torch.cuda.device(self.device):
            do_fill = dense is None
            if dense is None:
                dense = torch.empty(n_rows, n_cols, dtype=torch.float32, device=self.device)
            r, c, v = self._dev(row, torch.int32), self._dev(col, torch.int32), self._dev(val, torch.float32)
            self._check(self.lib.blim_scatter_scores(self.h, ctypes.c_void_p(dense.data_ptr()), n_rows, n_cols, int(do_fill), float(fill),
                                                     ctypes.c_void_p(r.data_ptr()), ctypes.c_void_p(c.data_ptr()),
                                                     ctypes.c_void_p(v.data_ptr()), r.numel(), self._stream()))
        return dense

    # ---------------------------------------------------------------- counters / debug
    def kernel_launches(self):
        return int(self.lib.blim_kernel_launches(self.h))

    def gemm_flops(self):
        return float(self.lib.blim_gemm_flops(self.h))

    def profile(self, enable=True):
        self._check(self.lib.blim_profile(self.h, int(enable)))

    def profile_read(self):
        """-> dict(gemm_ms, attn_ms, gemm_launches, attn_launches) since the last read (synchronises the device)."""
        g, a = ctypes.c_double(), ctypes.c_double()
        ng, na = ctypes.c_int64(), ctypes.c_int64()
        self._check(self.lib.blim_profile_read(self.h, ctypes.byref(g), ctypes.byref(a), ctypes.byref(ng), ctypes.byref(na)))
        return dict(gemm_ms=g.value, attn_ms=a.value, gemm_launches=ng.value, attn_launches=na.value)

    PROFILE_KINDS = ("qkv_rope", "o_proj", "gate_up_swiglu", "down_proj", "head_lse", "other_gemm", "attention", "rmsnorm")

    def profile_read_detail(self):
        """-> {kind: dict(ms, flops, launches)} of the intervals recorded so far; does not reset (call before profile_read)."""
        n = len(self.PROFILE_KINDS)
        ms, fl, ln = (ctypes.c_double * n)(), (ctypes.c_double * n)(), (ctypes.c_int64 * n)()
        self._check(self.lib.blim_profile_read_detail(self.h, n, ms, fl, ln))
        return {k: dict(ms=ms[i], flops=fl[i], launches=ln[i]) for i, k in enumerate(self.PROFILE_KINDS)}

    def debug_umma(self, A, B, b_mn_major, lbo=0, sbo=0, kstep=0):
        """Single-CTA tcgen05 probe (see blim_debug_umma): A [128, K], B [N, K] or (b_mn_major) [K, N] -> fp32 [128, N].
        Each operand is passed in its own 16-bit format (torch.float16 stays fp16, anything else becomes bf16)."""
        with torch.cuda.device(self.device):
            fa = torch.float16 if A.dtype == torch.float16 else torch.bfloat16
            fb = torch.float16 if B.dtype == torch.float16 else torch.bfloat16
            A, B = self._dev(A, fa), self._dev(B, fb)
            K = A.shape[1]
            N = B.shape[1] if b_mn_major else B.shape[0]
            C = torch.zeros(128, N, dtype=torch.float32, device=self.device)
            flags = int(bool(b_mn_major)) | (2 if fa == torch.float16 else 0) | (4 if fb == torch.float16 else 0)
            self._check(self.lib.blim_debug_umma(self.h, ctypes.c_void_p(A.data_ptr()), ctypes.c_void_p(B.data_ptr()),
                                                 ctypes.c_void_p(C.data_ptr()), K, N, flags, lbo, sbo, kstep, self._stream()))
        return C

    def debug_gemm(self, epilogue, A, W, bias=None, target=None, scale=1.0, cta_group=0, C=None):
        """Unit-test entry for the tcgen05 GEMM core (see blim_debug_gemm): A and W are converted to the engine's operand
        format; 16-bit results come back in that format."""
        with torch.cuda.device(self.device):
            A = self._dev(A, self.act_dtype)
            W = self._dev(W, self.act_dtype)
            M, K = A.shape
            N = W.shape[0]
            if C is None:
                if epilogue in (0, 1, 2):
                    C = torch.empty(M, N, dtype=self.act_dtype, device=self.device)
                elif epilogue == 3:
                    C = torch.empty(M, N, dtype=torch.float32, device=self.device)
                elif epilogue == 5:
                    C = torch.empty(M, N // 2, dtype=self.act_dtype, device=self.device)
                elif epilogue == 6:
                    C = torch.empty(M, dtype=torch.float32, device=self.device)
                else:
                    raise ValueError("epilogue 4 (residual add) needs an explicit fp32 C")
            b = self._dev(bias, torch.float32) if bias is not None else None
            t = self._dev(target, torch.int32) if target is not None else None
            p = lambda x: ctypes.c_void_p(x.data_ptr()) if x is not None else None
            self._check(self.lib.blim_debug_gemm(self.h, epilogue, p(A), p(W), p(C), M, N, K, p(b), p(t), float(scale), cta_group,
                                                 self._stream()))
        return C
