"""Drop-in mirror of the evaluation half of the reference's training_utils.py on top of the CUDA engine:

    val_one_epoch(model, data_loader, optimizer, device, epoch, loss_scaler, tokenizer=None, args=None)   training_utils.py:140
    get_recall(t2v, v2t, t2v_ids, v2t_ids)                                                                training_utils.py:173

The CPN subtraction, the BLiM / InternVideo2 ensemble and the rank search run in the engine's fuse_rerank / rank_dense
kernels; their arithmetic reproduces numpy's float32 (and, in the zero-shot text->video branch, float64) semantics
operation by operation, so R@K and the candidate order are bit-identical to the reference for identical scores.
"""
import numpy as np
import torch

from .retrieval import _engine_model, evaluation


def _recall_from_ranks(ranks, n):
    ranks = np.asarray(ranks)
    return (100.0 * int((ranks < 1).sum()) / n, 100.0 * int((ranks < 5).sum()) / n, 100.0 * int((ranks < 10).sum()) / n)


def _pack(t, v):
    t1, t5, t10 = t
    v1, v5, v10 = v
    vm, tm = (v1 + v5 + v10) / 3, (t1 + t5 + t10) / 3
    res = {"t2v_r1": t1, "t2v_r5": t5, "t2v_r10": t10, "t2v_r_mean": tm, "v2t_r1": v1, "v2t_r5": v5, "v2t_r10": v10, "v2t_r_mean": vm,
           "r_mean": (vm + tm) / 2}
    return {k: round(vv, 2) for k, vv in res.items()}          # training_utils.py:219-220


def get_recall(t2v, v2t, t2v_ids=None, v2t_ids=None, engine=None):
    """R@1/5/10 of dense score matrices with the diagonal ground truth of val_one_epoch (training_utils.py:146-147,
    173-221), the rank search running on the device.  A matrix containing an exact 0 counts as absent (tu:174,195)."""
    if engine is None:
        raise TypeError("get_recall needs the CUDA engine (engine=...)")
    out = []
    for m in (t2v, v2t):
        m = np.asarray(m)
        if m.dtype != np.float32:
            # float64 matrices (zero-shot t2v branch): ranks on the host copy, exactly numpy
            if np.count_nonzero(m == 0) != 0:
                out.append((0.0, 0.0, 0.0))
                continue
            ranks = np.array([np.where(np.argsort(r)[::-1] == i)[0][0] for i, r in enumerate(m)])
            out.append(_recall_from_ranks(ranks, m.shape[0]))
            continue
        rank, zero = engine.rank_dense(torch.from_numpy(np.ascontiguousarray(m)))
        if int(zero.item()) != 0:
            out.append((0.0, 0.0, 0.0))
        else:
            out.append(_recall_from_ranks(rank.cpu().numpy(), m.shape[0]))
    return _pack(out[0], out[1])


def fused_rerank(engine, t2v_c, v2t_c, t2v_iv2, v2t_iv2, alpha, c, cpn=True, zero_shot=False, use_query=True):
    """The "blim" row of val_one_epoch on compact candidate arrays (training_utils.py:154-165 + 173-221).
    Returns (result dict, per-direction dict with fused scores / candidate order / ground-truth ranks, all on device)."""
    full = not zero_shot
    detail, recalls = {}, []
    for d, comp, iv2, a, cq, ce in (("t2v", t2v_c, t2v_iv2, alpha[0], c[0], c[2]), ("v2t", v2t_c, v2t_iv2, alpha[1], c[1], c[3])):
        zero_f64 = zero_shot and d == "t2v"
        fused, order, rank, zero = engine.fuse_rerank(
            comp["idx"], comp.get("candidate_likelihood"), comp.get("candidate_prior") if cpn else None, comp.get("query_likelihood"),
            iv2, a, cq, ce, use_prior=cpn and "candidate_prior" in comp, use_query=use_query and (full or d == "t2v"),
            cpn_zero_f64=zero_f64)
        detail[d] = {"fused": fused, "order": order, "rank": rank, "zero": zero}
    for d in ("t2v", "v2t"):
        if int(detail[d]["zero"].item()) != 0:
            recalls.append((0.0, 0.0, 0.0))
        else:
            r = detail[d]["rank"].cpu().numpy()
            recalls.append(_recall_from_ranks(r, len(r)))
    return _pack(recalls[0], recalls[1]), detail


def _recall_device(engine, t2v_mat, v2t_mat):
    """get_recall on device-resident dense fp32 matrices (None = the reference's np.zeros placeholder -> recalls 0)."""
    out = []
    for m in (t2v_mat, v2t_mat):
        if m is None:
            out.append((0.0, 0.0, 0.0))
            continue
        rank, zero = engine.rank_dense(m)
        out.append((0.0, 0.0, 0.0) if int(zero.item()) != 0 else _recall_from_ranks(rank.cpu().numpy(), m.shape[0]))
    return _pack(out[0], out[1])


def val_one_epoch(model, data_loader, optimizer, device, epoch, loss_scaler, tokenizer=None, args=None):
    """Same contract as the reference's val_one_epoch (training_utils.py:140-169): the result table for the rows
    internvideo2 / candidate_likelihood / query_likelihood / cpn_candidate_likelihood / blim.  Everything is ranked on
    the device from the matrices evaluation() left there; nothing is uploaded again."""
    m = _engine_model(model)
    eng = m.engine
    t2v_dict, v2t_dict = evaluation(model, data_loader, device, tokenizer, args)
    zero_shot = args.resume == "" and args.eval
    full = not zero_shot
    t2v_dev, v2t_dev = m.last_dense
    t2v_c, v2t_c = m.last_compact
    results = {}
    for name in ("internvideo2", "candidate_likelihood", "query_likelihood"):
        results[name] = _recall_device(eng, t2v_dev.get(name), v2t_dev.get(name))
    if args.cpn:
        # cand - alpha * prior through the fuse kernel with the ensemble switched off (c = 1: 1*x + 0*iv2 is exact)
        res, _ = fused_rerank(eng, t2v_c, v2t_c, t2v_dev["internvideo2"], v2t_dev["internvideo2"], args.alpha, (0.0, 0.0, 1.0, 1.0),
                              cpn=True, zero_shot=False, use_query=False)
        if not full:   # reference: cpn_t2v = np.zeros -> "matrix absent" -> text->video recalls 0 (tu:154,195)
            res = _pack((0.0, 0.0, 0.0), (res["v2t_r1"], res["v2t_r5"], res["v2t_r10"]))
        results["cpn_candidate_likelihood"] = res
    results["blim"], m.last_rerank = fused_rerank(eng, t2v_c, v2t_c, t2v_dev["internvideo2"], v2t_dev["internvideo2"], args.alpha, args.c,
                                                  cpn=bool(args.cpn), zero_shot=zero_shot)
    return results
