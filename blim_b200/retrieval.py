"""Drop-in mirror of the reference's retrieval_utils.py (same names, argument lists, return values) on top of the
CUDA engine.

    evaluation(model, data_loader, device, tokenizer, args)                      reference retrieval_utils.py:170
    compute_v2t_scores_x / compute_t2v_scores_x (same 14 parameters)             reference retrieval_utils.py:48 / 113
    padding_ids(input_ids, labels, masks, tokenizer)                             reference retrieval_utils.py:155

Differences in HOW (not in WHAT): a row's top-k candidates are not pushed through per-row padded batches; all rows'
(video, text) pairs go to Engine.score_pairs in one call, which shares video / text prefixes and never materialises the
logits.  evaluation() additionally scores each distinct (video, text) pair once for both directions and shards the
pairs by prefix owner over the ranks, combining them with ONE all-gather of compact scores (NCCL, enqueued by the engine
on the compute stream) instead of dense all-reduces (reference retrieval_utils.py:252-262).
"""
import datetime
import time

import numpy as np
import torch
import torch.distributed as dist

from .engine import TEXTS_TVG, TEXTS_VTG, TVG, TVG_PRIOR, VTG, VTG_PRIOR

IGNORE_INDEX = -100


def padding_ids(input_ids, labels, masks, tokenizer=None):
    """LEFT-pad ragged id / label / mask lists to the global maximum (reference retrieval_utils.py:155-167)."""
    n = len(input_ids)
    width = max(len(x) for x in input_ids)
    pad_id = tokenizer.pad_token_id if tokenizer is not None else 0
    ids_p = torch.full((n, width), pad_id, dtype=torch.long)
    lab_p = torch.full((n, width), IGNORE_INDEX, dtype=torch.long)
    msk_p = torch.zeros((n, width), dtype=torch.long)
    for i in range(n):
        m = len(input_ids[i])
        ids_p[i, width - m:] = input_ids[i]
        lab_p[i, width - m:] = labels[i]
        msk_p[i, width - m:] = masks[i]
    return ids_p, lab_p, msk_p


def _engine_model(model):
    m = getattr(model, "module", model)
    if not hasattr(m, "engine"):
        raise TypeError("blim_b200.retrieval needs a blim_b200.model.BlimModel (CUDA engine); there is no PyTorch fallback")
    return m


def _rows_topk(iterator, topk, device, engine):
    rows = [r if torch.is_tensor(r) else torch.as_tensor(r) for r in iterator]
    if not rows:
        return None
    sims = torch.stack(rows, 0).to(device)
    return engine.topk_rows(sims.float(), topk)[0]  # [rows, k], same per-row result as sims.topk(k, dim=0) (ru:52,117)


def _score_rows(scores_x, iterator, start, input_ids, attention_masks, labels, video, video_vocab, tvg_video_labels, model, device, args,
                forward_type, cpn, rows_are_videos):
    m = _engine_model(model)
    eng = m.engine
    idx = _rows_topk(iterator, args.topk, eng.device, eng)
    if idx is None:
        return scores_x
    n_rows, k = idx.shape
    which = TEXTS_TVG if forward_type == "tvg" else TEXTS_VTG
    m.ensure_videos(video)
    m.ensure_texts(which, input_ids, attention_masks, labels)
    if forward_type == "tvg":
        m.ensure_vocab(video_vocab, tvg_video_labels)
    kind = {("vtg", False): VTG, ("vtg", True): VTG_PRIOR, ("tvg", False): TVG, ("tvg", True): TVG_PRIOR}[(forward_type, bool(cpn))]
    rows = torch.arange(start, start + n_rows, device=idx.device)[:, None].expand(n_rows, k).reshape(-1)
    cols = idx.reshape(-1)
    pv, pt = (rows, cols) if rows_are_videos else (cols, rows)
    scores = eng.score_pairs(kind, pv.cpu().numpy(), pt.cpu().numpy())
    scores_x[rows.to(scores_x.device), cols.to(scores_x.device)] = scores.to(scores_x.device, scores_x.dtype)
    return scores_x


def compute_v2t_scores_x(v2t_scores_x, iterator, start, input_ids, attention_masks, labels, video, video_vocab, tvg_video_labels, model,
                         device, args, forward_type=None, cpn=False):
    """Video -> text rows: row `start + i` is a video, its top-k texts are scored (reference retrieval_utils.py:48-111)."""
    return _score_rows(v2t_scores_x, iterator, start, input_ids, attention_masks, labels, video, video_vocab, tvg_video_labels, model,
                       device, args, forward_type, cpn, rows_are_videos=True)


def compute_t2v_scores_x(t2v_scores_x, iterator, start, input_ids, attention_masks, labels, video, video_vocab, tvg_video_labels, model,
                         device, args, forward_type=None, cpn=False):
    """Text -> video rows: row `start + i` is a text, its top-k videos are scored (reference retrieval_utils.py:113-153)."""
    return _score_rows(t2v_scores_x, iterator, start, input_ids, attention_masks, labels, video, video_vocab, tvg_video_labels, model,
                       device, args, forward_type, cpn, rows_are_videos=False)


# ------------------------------------------------------------------------------------------------ evaluation
def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


class PairPlan:
    """Candidate pairs of one evaluation: per direction the top-k index arrays, plus the deduplicated union.
    Device tensors for the kernels and numpy copies for the host-side scheduler (one D2H at construction, so nothing
    later has to synchronise the stream to look at indices)."""

    def __init__(self, v2t_iv2, t2v_iv2, topk, device, engine=None):
        if engine is not None:   # stage-1 candidates on the device with the engine's warp-level top-k kernel
            self.v2t_idx = engine.topk_rows(v2t_iv2, topk)[0]                                     # [Nv, k] text ids
            self.t2v_idx = engine.topk_rows(t2v_iv2, topk)[0]                                     # [Nt, k] video ids
        else:                    # host-logic tests without a GPU
            self.v2t_idx = v2t_iv2.to(device).topk(k=min(v2t_iv2.shape[1], topk), dim=1).indices
            self.t2v_idx = t2v_iv2.to(device).topk(k=min(t2v_iv2.shape[1], topk), dim=1).indices
        nv, k = self.v2t_idx.shape
        nt = self.t2v_idx.shape[0]
        self.n_videos, self.n_texts, self.k = nv, nt, k
        v_a = torch.arange(nv, device=device)[:, None].expand(nv, k).reshape(-1)
        t_a = self.v2t_idx.reshape(-1)
        t_b = torch.arange(nt, device=device)[:, None].expand(nt, self.t2v_idx.shape[1]).reshape(-1)
        v_b = self.t2v_idx.reshape(-1)
        self.v2t_pairs = (v_a, t_a)   # entries of the v2t matrices  [v, t]
        self.t2v_pairs = (v_b, t_b)   # entries of the t2v matrices  [t, v]
        key = torch.cat([v_a * nt + t_a, v_b * nt + t_b])
        self.union_key, inverse = torch.unique(key, return_inverse=True)  # sorted by (video, text)
        self.v2t_in_union = inverse[:v_a.numel()]
        self.t2v_in_union = inverse[v_a.numel():]
        self.union_v = torch.div(self.union_key, nt, rounding_mode="floor")
        self.union_t = self.union_key - self.union_v * nt
        packed = torch.cat([self.union_v, self.union_t, v_a, t_a, v_b, t_b]).cpu().numpy().astype(np.int32)
        nu, na, nb = self.union_v.numel(), v_a.numel(), v_b.numel()
        self.union_np = (packed[:nu], packed[nu:2 * nu])
        self.v2t_np = (packed[2 * nu:2 * nu + na], packed[2 * nu + na:2 * nu + 2 * na])
        self.t2v_np = (packed[2 * nu + 2 * na:2 * nu + 2 * na + nb], packed[2 * nu + 2 * na + nb:])


LM_ROW_COST = 0.084   # one LM-head row in units of one decoder token (2*H*V / 13.05 GFLOP at 7B; a few % either way elsewhere)


def _shard_costs(eng, kind, pv, pt):
    """(owner type, owner ids, per-pair cost, per-owner cost) in decoder-token equivalents for the scheduler's four run
    shapes (csrc/engine.cu), i.e. what the rank that owns the prefix will actually execute after the engine's own
    deduplication.  Owner = the id whose shared prefix the pairs hang off: "v" (video) or "t" (text)."""
    lens = getattr(eng, "text_lens", None)
    if not lens or TEXTS_VTG not in lens or TEXTS_TVG not in lens:
        lens = None
    n_vis = getattr(eng, "n_clips", 4) * 64
    n_clips = getattr(eng, "n_clips", 4)
    root = getattr(eng, "tvg_prefix_length", 21)
    if lens:
        scored = lens[TEXTS_VTG]["scored"]                     # caption + <|im_end|> + "\n": LM rows of a pair
        suffix = (scored - 1.0) + LM_ROW_COST * scored          # decoder tokens of the suffix sequence + its LM rows
        t0 = lens[TEXTS_TVG]["total"] - 3.0                    # text part of the TVG prompt
    if kind == VTG:          # owner = video: [visual rows + prompt tail] once (the header is a shared root), one caption suffix per pair
        return "v", pv, (suffix[pt] if lens else np.full(len(pv), 15.0)), n_vis + 12.0
    if kind == VTG_PRIOR:    # owner = text: ONE suffix per distinct text however many videos list it; the 26-token prefix is shared
        per_text = np.zeros(int(pt.max()) + 1 if len(pt) else 0)
        if len(pt):
            per_text[np.unique(pt)] = suffix[np.unique(pt)] if lens else 15.0
        return "t", pt, np.zeros(len(pt)), per_text
    if kind == TVG:          # owner = text: its text prefix behind the shared root once, n_clips - 1 visual rows per pair
        return "t", pt, np.full(len(pt), n_clips - 1.0), (np.maximum(t0 - root, 1.0) if lens else 25.0)
    # TVG_PRIOR, owner = video: one n_clips-token suffix per distinct text length among the video's pairs (the engine dedupes
    # on (header, T0, last token, video))
    per_video = np.zeros(int(pv.max()) + 1 if len(pv) else 0)
    if len(pv):
        key = np.unique(pv.astype(np.int64) * 65536 + (t0[pt].astype(np.int64) if lens else 0))
        per_video += np.bincount(key // 65536, minlength=len(per_video)) * float(n_clips)
    return "v", pv, np.zeros(len(pv)), per_video


def balanced_owner_ranks(costs, world):
    """Which rank scores which prefix owner.  `costs` = {"v": per-video decoder tokens, "t": per-text decoder tokens}
    summed over ALL score kinds of the evaluation: videos and texts go into one list, largest first, each to the currently
    lightest rank (LPT).  The ranks meet once, at the single all-gather, so only the total per rank matters.
    Deterministic: every rank derives the same assignment from the plan.  Returns {"v": rank_of_video, "t": rank_of_text}."""
    import heapq
    items = [(float(c), typ, i) for typ in ("v", "t") for i, c in enumerate(costs.get(typ, ())) if c > 0]
    items.sort(key=lambda x: (-x[0], x[1], x[2]))
    heap = [(0.0, r) for r in range(world)]          # (load, rank): ties go to the lowest rank, like argmin
    rank_of = {typ: np.zeros(len(costs.get(typ, ())), dtype=np.int64) for typ in ("v", "t")}
    for c, typ, i in items:
        load, r = heapq.heappop(heap)
        rank_of[typ][i] = r
        heapq.heappush(heap, (load + c, r))
    return rank_of


class ShardPlan:
    """Who scores what, and where it lands: per score kind the pair indices of every rank, the offset of that shard in the
    rank's send buffer (all kinds concatenated), and the gather indices that unpack the all-gathered buffer."""

    def __init__(self, eng, jobs, world, n_videos, n_texts):
        costs = {"v": np.zeros(n_videos), "t": np.zeros(n_texts)}
        owners = {}
        for name, kind, pv, pt in jobs:
            typ, owner, pair_cost, base = _shard_costs(eng, kind, pv, pt)
            n_owner = len(costs[typ])
            c = np.bincount(owner, weights=pair_cost, minlength=n_owner).astype(np.float64)
            used = np.bincount(owner, minlength=n_owner) > 0
            if isinstance(base, np.ndarray):
                base = np.pad(base, (0, max(0, n_owner - len(base))))[:n_owner]
            costs[typ] += c + np.where(used, base, 0.0)
            owners[name] = (typ, owner)
        rank_of = balanced_owner_ranks(costs, world)
        self.shards = {}                      # name -> [pair indices of rank r]
        fill = np.zeros(world, dtype=np.int64)
        self.offsets = {}                     # name -> [offset of that shard in rank r's send buffer]
        for name, kind, pv, pt in jobs:
            typ, owner = owners[name]
            pair_rank = rank_of[typ][owner]
            self.shards[name] = [np.nonzero(pair_rank == r)[0] for r in range(world)]
            self.offsets[name] = fill.copy()
            fill += np.array([len(x) for x in self.shards[name]])
        self.width = int(max(1, fill.max()))  # floats every rank contributes (padded to the longest)
        self.world = world

    def unpack_indices(self, name):
        """(src positions in the gathered [world * width] buffer, dst positions in the kind's per-pair array)."""
        src = np.concatenate([r * self.width + self.offsets[name][r] + np.arange(len(x)) for r, x in enumerate(self.shards[name])])
        dst = np.concatenate(self.shards[name])
        return src, dst


def score_all(model, plan: PairPlan, cpn=True, full=True, distributed=False):
    """Scores every term evaluation() needs on the deduplicated pair set.  Multi-GPU: the pairs of all score kinds are
    sharded by the id that owns the shared prefix (each video / text prefix is prefilled on exactly one rank), balanced
    on the SUM of the kinds' decoder tokens; every rank derives all ranks' shards from the same plan, scores its own into
    one send buffer, and the ranks meet ONCE: a single all-gather of padded fp32 scores (blim_allgather_scores, NCCL on
    the compute stream; no index or size exchange, no host synchronisation) replaces the reference's barrier + 2-6 dense
    all-reduces (retrieval_utils.py:252-262).
    Returns a dict of compact device tensors aligned with plan.union_* (vtg, tvg) / plan.v2t_pairs (vtg_prior) /
    plan.t2v_pairs (tvg_prior)."""
    m = _engine_model(model)
    eng = m.engine
    rank, world = _world() if distributed else (0, 1)
    jobs = [("vtg", VTG) + tuple(plan.union_np)]
    if cpn:
        jobs.append(("vtg_prior", VTG_PRIOR) + tuple(plan.v2t_np))
    if full:
        jobs.append(("tvg", TVG) + tuple(plan.union_np))
        if cpn:
            jobs.append(("tvg_prior", TVG_PRIOR) + tuple(plan.t2v_np))
    if world == 1:
        return {name: eng.score_pairs(kind, pv, pt) for name, kind, pv, pt in jobs}
    sp = ShardPlan(eng, jobs, world, plan.n_videos, plan.n_texts)
    send = torch.zeros(sp.width, dtype=torch.float32, device=eng.device)
    timing = getattr(m, "shard_timing", None)      # bench.py: device time of this rank's own scoring, before the ranks meet
    if timing is not None:
        timing["t0"] = torch.cuda.Event(enable_timing=True)
        timing["t0"].record()
    for name, kind, pv, pt in jobs:
        mine = sp.shards[name][rank]
        if len(mine):
            off = int(sp.offsets[name][rank])
            eng.score_pairs(kind, pv[mine], pt[mine], out=send[off:off + len(mine)])
    if timing is not None:
        timing["t1"] = torch.cuda.Event(enable_timing=True)
        timing["t1"].record()
    if hasattr(eng, "allgather_scores"):
        eng.comm_init()                        # first call only: joins the engine's NCCL communicator
        gathered = eng.allgather_scores(send)
    else:                                      # host-logic tests on CPU (gloo) with a stand-in engine
        gathered = torch.empty(world * sp.width, dtype=torch.float32, device=eng.device)
        dist.all_gather_into_tensor(gathered, send)
    out = {}
    for name, kind, pv, pt in jobs:
        src, dst = sp.unpack_indices(name)
        res = torch.empty(len(pv), dtype=torch.float32, device=eng.device)
        res[torch.from_numpy(dst).to(eng.device, non_blocking=True)] = gathered[torch.from_numpy(src).to(eng.device, non_blocking=True)]
        out[name] = res
    return out


def compact_terms(plan: PairPlan, s, cpn=True, full=True):
    """[rows, k] arrays per direction in the reference's vocabulary (retrieval_utils.py:264-276)."""
    nv, nt = plan.n_videos, plan.n_texts
    v2t = {"idx": plan.v2t_idx, "candidate_likelihood": s["vtg"][plan.v2t_in_union].view(nv, -1)}
    t2v = {"idx": plan.t2v_idx, "query_likelihood": s["vtg"][plan.t2v_in_union].view(nt, -1)}
    if cpn:
        v2t["candidate_prior"] = s["vtg_prior"].view(nv, -1)
    if full:
        v2t["query_likelihood"] = s["tvg"][plan.v2t_in_union].view(nv, -1)
        t2v["candidate_likelihood"] = s["tvg"][plan.t2v_in_union].view(nt, -1)
        if cpn:
            t2v["candidate_prior"] = s["tvg_prior"].view(nt, -1)
    return t2v, v2t


def dense_matrices(eng, compact, n_rows, n_cols):
    """-100-filled dense matrices with the candidate entries scattered in (reference retrieval_utils.py:219,110,152)."""
    idx = compact["idx"]
    rows = torch.arange(n_rows, device=idx.device)[:, None].expand_as(idx).reshape(-1)
    out = {}
    for name, val in compact.items():
        if name == "idx":
            continue
        out[name] = eng.scatter_scores(n_rows, n_cols, rows, idx.reshape(-1), val.reshape(-1))
    return out


@torch.no_grad()
def evaluation(model, data_loader, device, tokenizer, args):
    """Same contract as the reference's evaluation() (retrieval_utils.py:169-281): returns (t2v_dict, v2t_dict) of numpy
    fp32 matrices under the keys candidate_likelihood / query_likelihood / internvideo2 / candidate_prior.
    The InternVideo2 scores come from `./scores/{dataset}[_zeroshot].pth` like in the reference, or from
    `args.iv2_scores = {"v2t": ..., "t2v": ...}` when present (synthetic runs)."""
    m = _engine_model(model)
    eng = m.engine
    m._corpus_keys.clear()   # one upload per evaluation() call: this call's loader output is what gets scored
    start_time = time.time()
    video, tvg_video_labels = [], []
    vtg_ids, vtg_labels, vtg_masks = [], [], []
    tvg_ids, tvg_labels, tvg_masks = [], [], []
    staged = getattr(getattr(data_loader, "dataset", data_loader), "staged_corpus", None)   # blim_b200.dataset.stage_corpus
    if staged is None:
        for data in data_loader:                                       # retrieval_utils.py:182-193
            video += [v for v in data["video"]]
            vtg_ids += data["vtg_ids"]; vtg_labels += data["vtg_labels"]; vtg_masks += data["vtg_masks"]
            tvg_ids += data["tvg_ids"]; tvg_labels += data["tvg_labels"]; tvg_masks += data["tvg_masks"]
            tvg_video_labels.append(data["tvg_video_labels"])
        tvg_video_labels = torch.cat(tvg_video_labels, dim=0)
    zero_shot = args.resume == "" and args.eval                         # retrieval_utils.py:199
    scores = getattr(args, "iv2_scores", None)
    if scores is None:
        name = f"./scores/{args.dataset.lower()}_zeroshot.pth" if zero_shot else f"./scores/{args.dataset.lower()}.pth"
        scores = torch.load(name, weights_only=True)
    v2t_iv2, t2v_iv2 = scores["v2t"], scores["t2v"]
    num_texts, num_videos = t2v_iv2.shape
    full = not zero_shot

    # pad-stripped ragged texts go straight to the engine (padding_ids + the mask strip of mvf:333-334 cancel out)
    if staged is None:
        m.ensure_videos(video, shard=_world() if getattr(args, "distributed", False) else None)
        eng.set_texts(TEXTS_VTG, vtg_ids, vtg_labels, vtg_masks)
        eng.set_texts(TEXTS_TVG, tvg_ids, tvg_labels, tvg_masks)
        if full:
            m.ensure_vocab(data_loader.dataset.video_vocab, tvg_video_labels)
    # else: features, token tables and the device-built video vocabulary are already in the engine
    m.set_tvg_prefix_length(data_loader.dataset.tvg_prefix_length)      # retrieval_utils.py:210

    plan = PairPlan(v2t_iv2.float(), t2v_iv2.float(), args.topk, eng.device, engine=eng)
    s = score_all(model, plan, cpn=bool(args.cpn), full=full, distributed=bool(getattr(args, "distributed", False)))
    t2v_c, v2t_c = compact_terms(plan, s, cpn=bool(args.cpn), full=full)
    m.last_plan, m.last_compact = plan, (t2v_c, v2t_c)                  # kept on the device for the fused rerank
    t2v_dev, v2t_dev = dense_matrices(eng, t2v_c, num_texts, num_videos), dense_matrices(eng, v2t_c, num_videos, num_texts)
    t2v_dev["internvideo2"], v2t_dev["internvideo2"] = t2v_iv2.float().to(eng.device), v2t_iv2.float().to(eng.device)
    m.last_dense = (t2v_dev, v2t_dev)                                   # device copies: val_one_epoch ranks them in place

    t2v_dict = {k: v.cpu().numpy() for k, v in t2v_dev.items() if k != "internvideo2"}
    v2t_dict = {k: v.cpu().numpy() for k, v in v2t_dev.items() if k != "internvideo2"}
    t2v_dict["internvideo2"] = t2v_iv2.cpu().numpy()
    v2t_dict["internvideo2"] = v2t_iv2.cpu().numpy()
    print(f"Evaluation time {str(datetime.timedelta(seconds=int(time.time() - start_time)))}")
    return t2v_dict, v2t_dict
