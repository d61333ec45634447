"""TEST INFRASTRUCTURE -- drive the UNMODIFIED reference (oracle/ref_harness.py) on a slice of a synthetic corpus.

Used as the comparator of the 7B parity tests, by tools/parity_7b.py and by bench.py (reference arm: the reference's
own PyTorch path timed on the B200; own arm: rank-parity check against it after the timed region).  Nothing here is on
the product path.

What is run is the reference's stock code: retrieval_utils.compute_v2t_scores_x / compute_t2v_scores_x
(retrieval_utils.py:48-153) on the reference model object (modeling_videochat_flash.py:572-629) with sdpa attention,
under torch.autocast like training_utils.py:142 does (bf16 here: BASELINE.json north_star), batch_size_eval = 16.
"""
import contextlib
import time
import types

import numpy as np
import torch

# the six score matrices of evaluation() (retrieval_utils.py:219-250): name -> (direction, forward_type, cpn)
MATRICES = {
    "v2t_candidate_likelihood": ("v2t", "vtg", False),
    "v2t_candidate_prior": ("v2t", "vtg", True),
    "v2t_query_likelihood": ("v2t", "tvg", False),
    "t2v_query_likelihood": ("t2v", "vtg", False),
    "t2v_candidate_likelihood": ("t2v", "tvg", False),
    "t2v_candidate_prior": ("t2v", "tvg", True),
}


def pad_left(seqs, fill):
    """padding_ids, retrieval_utils.py:155-167."""
    L = max(len(s) for s in seqs)
    out = torch.full((len(seqs), L), fill, dtype=torch.long)
    for i, s in enumerate(seqs):
        out[i, L - len(s):] = s
    return out


class ReferenceRunner:
    """The reference model for `cfg` with the given parameters, on `device`, in `dtype`."""

    def __init__(self, cfg, state_dict, corpus, device, dtype=torch.bfloat16, autocast=True):
        from oracle import ref_harness
        self.cfg, self.corpus, self.device, self.dtype = cfg, corpus, torch.device(device), dtype
        self.autocast = autocast and dtype != torch.float32
        sd = {k: v.to(dtype) for k, v in state_dict.items()} if dtype != torch.bfloat16 else state_dict
        self.model, (self.ru, self.tu, self.mvf) = ref_harness.build_reference_model(
            cfg, sd, dtype=dtype, image_token_id=cfg.image_token_id, device=self.device)
        del sd
        self.model.module.set_tvg_prefix_length(corpus.tvg_prefix_length)          # retrieval_utils.py:210
        self.video = [v.to(dtype) for v in corpus.video.cpu()]                     # the loader hands out CPU tensors (ru:55)
        self.vocab = corpus.video_vocab.to(self.device, dtype)                     # dataset.video_vocab.cuda(), ru:209
        self.text = {}
        for ft, ids_l, lab_l in (("vtg", corpus.vtg_ids, corpus.vtg_labels), ("tvg", corpus.tvg_ids, corpus.tvg_labels)):
            self.text[ft] = (pad_left(ids_l, corpus.pad_token_id), pad_left([torch.ones_like(x) for x in ids_l], 0), pad_left(lab_l, -100))

    def _ctx(self):
        if self.device.type != "cuda":
            return contextlib.nullcontext()
        if self.autocast:
            return torch.autocast("cuda", dtype=self.dtype)                        # training_utils.py:142 (fp16 there)
        # fp32 comparator: plain fp32 maths everywhere (no TF32, eager-equivalent attention)
        from torch.nn.attention import SDPBackend, sdpa_kernel
        return sdpa_kernel(SDPBackend.MATH)

    @torch.no_grad()
    def matrix_rows(self, name, row0, n_rows, topk, batch_size=16):
        """Rows [row0, row0 + n_rows) of one score matrix through the reference's own loop.
        Returns (idx [n_rows, k] candidate ids in top-k order, scores [n_rows, k] fp32)."""
        direction, ft, cpn = MATRICES[name]
        c = self.corpus
        sims = c.v2t_iv2 if direction == "v2t" else c.t2v_iv2
        fn = self.ru.compute_v2t_scores_x if direction == "v2t" else self.ru.compute_t2v_scores_x
        ids, masks, labels = self.text[ft]
        args = types.SimpleNamespace(topk=topk, batch_size_eval=batch_size, num_clips=c.n_clips)
        out = torch.full(tuple(sims.shape), -100.0).to(self.device)                # retrieval_utils.py:219
        with self._ctx():
            out = fn(out, sims[row0:row0 + n_rows], row0, ids, masks, labels, self.video, self.vocab, c.tvg_video_labels, self.model,
                     self.device, args, forward_type=ft, cpn=cpn)
        k = min(sims.shape[1], topk)
        idx = sims[row0:row0 + n_rows].topk(k=k, dim=1).indices
        return idx.numpy(), torch.gather(out[row0:row0 + n_rows].cpu(), 1, idx).float().numpy()

    def all_matrices(self, row0, n_rows, topk, batch_size=16, names=None):
        """-> ({name: (idx, scores)}, seconds spent, pairs scored in the BASELINE.json sense = 2 * n_rows * k)."""
        names = list(MATRICES) if names is None else names
        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)
        t0 = time.time()
        out = {n: self.matrix_rows(n, row0, n_rows, topk, batch_size) for n in names}
        if self.device.type == "cuda":
            torch.cuda.synchronize(self.device)
        dt = time.time() - t0
        k = next(iter(out.values()))[0].shape[1]
        return out, dt, 2 * n_rows * k

    def close(self):
        del self.model
        if self.device.type == "cuda":
            torch.cuda.empty_cache()


def engine_matrices(engine, corpus, row0, n_rows, topk, names=None):
    """The same rows of the same matrices from the CUDA engine (blim_score_pairs)."""
    from blim_b200.engine import TVG, TVG_PRIOR, VTG, VTG_PRIOR
    kind = {("vtg", False): VTG, ("vtg", True): VTG_PRIOR, ("tvg", False): TVG, ("tvg", True): TVG_PRIOR}
    out = {}
    for name in (list(MATRICES) if names is None else names):
        direction, ft, cpn = MATRICES[name]
        sims = corpus.v2t_iv2 if direction == "v2t" else corpus.t2v_iv2
        k = min(sims.shape[1], topk)
        idx = sims[row0:row0 + n_rows].topk(k=k, dim=1).indices.numpy()
        rows = np.repeat(np.arange(row0, row0 + n_rows), k)
        cols = idx.reshape(-1)
        pv, pt = (rows, cols) if direction == "v2t" else (cols, rows)
        out[name] = (idx, engine.score_pairs(kind[(ft, cpn)], pv, pt).cpu().numpy().reshape(n_rows, k))
    return out


def fused_rows(mats, corpus, row0, n_rows, alpha, c):
    """val_one_epoch's "blim" arithmetic (training_utils.py:154-165) on row slices -> {direction: (fused [n_rows, N] fp32,
    candidate order [n_rows, k] = ids of the k best columns best first, ground-truth rank [n_rows])}."""
    from oracle import blim_oracle as O
    n = corpus.n
    dense = {}
    for name, (idx, sc) in mats.items():
        m = np.full((n_rows, n), -100.0, dtype=np.float32)
        np.put_along_axis(m, idx, sc.astype(np.float32), axis=1)
        dense[name] = m
    t2v = {"candidate_likelihood": dense["t2v_candidate_likelihood"], "query_likelihood": dense["t2v_query_likelihood"],
           "candidate_prior": dense["t2v_candidate_prior"], "internvideo2": corpus.t2v_iv2[row0:row0 + n_rows].numpy()}
    v2t = {"candidate_likelihood": dense["v2t_candidate_likelihood"], "query_likelihood": dense["v2t_query_likelihood"],
           "candidate_prior": dense["v2t_candidate_prior"], "internvideo2": corpus.v2t_iv2[row0:row0 + n_rows].numpy()}
    blim_t2v, blim_v2t, _, _ = O.fuse(t2v, v2t, alpha, c, cpn=True, zero_shot=False)
    out = {}
    k = next(iter(mats.values()))[0].shape[1]
    for d, m in (("t2v", blim_t2v), ("v2t", blim_v2t)):
        order = np.argsort(m, axis=1)[:, ::-1]                                    # get_recall, training_utils.py:173-221
        gt = np.array([int(np.where(order[i] == row0 + i)[0][0]) for i in range(n_rows)])
        out[d] = (m, order[:, :k].copy(), gt)
    return out


def rank_parity(fused_a, fused_b):
    """Agreement of two fused_rows() results (a = comparator, b = candidate): reranked candidate ids, ground-truth ranks,
    R@1/5/10 on the slice, and -- for the rows whose order differs -- how close the swapped scores were in the comparator."""
    rep = {}
    for d in ("t2v", "v2t"):
        ma, oa, ga = fused_a[d]
        mb, ob, gb = fused_b[d]
        n_rows, k = oa.shape
        same_order = (oa == ob).all(1)
        top1 = oa[:, 0] == ob[:, 0]
        sa = np.take_along_axis(ma, oa, 1)
        margins = sa[:, 0] - sa[:, 1]
        gaps = []    # comparator score gap of every adjacent pair the candidate orders differently
        for i in np.nonzero(~same_order)[0]:
            pos_b = {int(v): j for j, v in enumerate(ob[i])}
            for j in range(k - 1):
                x, y = int(oa[i, j]), int(oa[i, j + 1])
                if x in pos_b and y in pos_b and pos_b[x] > pos_b[y]:
                    gaps.append(float(ma[i, x] - ma[i, y]))
        rec = lambda g: [float(100.0 * (g < t).sum() / n_rows) for t in (1, 5, 10)]
        rep[d] = {"rows": int(n_rows), "k": int(k), "rows_same_order": int(same_order.sum()), "rows_same_top1": int(top1.sum()),
                  "rows_same_gt_rank": int((ga == gb).sum()), "recall_a_r1_r5_r10": rec(ga), "recall_b_r1_r5_r10": rec(gb),
                  "recall_equal": rec(ga) == rec(gb), "min_top1_margin_a": float(margins.min()), "median_top1_margin_a": float(np.median(margins)),
                  "swapped_adjacent_pairs": len(gaps), "max_gap_of_swapped_pairs_a": max(gaps) if gaps else 0.0}
    return rep
