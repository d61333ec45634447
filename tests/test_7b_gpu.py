"""Full-size (VideoChat-Flash-Qwen2-7B architecture, random init) parity and property tests.  GPU only, ~3 minutes.

 * engine vs the unmodified reference run on the GPU in fp32 and in bf16, 256 pairs of every score matrix:
   |d log-likelihood| <= 1e-2 against the fp32 run (BASELINE.json north_star tolerance), fixed bounds against the bf16 run;
 * size-independent properties on MSRVTT-shaped data: re-batching invariance (a pair's score does not depend on
   which other pairs share its decoder run) and direction symmetry of the deduplicated pair set."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from blim_b200 import synth
from blim_b200.engine import TVG, TVG_PRIOR, VTG, VTG_PRIOR, ModelConfig
from blim_b200.model import BlimModel
from oracle import blim_oracle as O

KIND = {("vtg", False): VTG, ("vtg", True): VTG_PRIOR, ("tvg", False): TVG, ("tvg", True): TVG_PRIOR}


@pytest.fixture(scope="module")
def seven_b():
    cfg = ModelConfig.qwen2_7b()
    dev = torch.device("cuda", 0)
    model = BlimModel(cfg, device=0)
    shapes = synth.param_shapes(cfg)
    weights = {}
    for idx, name in enumerate(shapes):
        t = synth.init_weight(cfg, name, idx, seed=0, device=dev, std=0.02, rich=True)
        model.engine.load_weight(name, t)
        weights[name] = t
    model.engine.set_rope(torch.float32)
    yield cfg, model, weights
    model.engine.close()


def _set_corpus(model, corpus):
    eng = model.engine
    eng.set_videos(corpus.video)
    eng.set_texts(0, corpus.vtg_ids, corpus.vtg_labels)
    eng.set_texts(1, corpus.tvg_ids, corpus.tvg_labels)
    eng.set_video_vocab(corpus.video_vocab, corpus.tvg_video_labels.numpy())
    model.set_tvg_prefix_length(corpus.tvg_prefix_length)


def test_7b_scores_match_reference(seven_b):
    """256 pairs of every one of the six score matrices (16 rows x top-16 of an MSRVTT-1k-shaped corpus) against the
    UNMODIFIED reference run on this GPU (oracle/ref_gpu.py; sources shipped in baseline/_ref by build()):

      * engine vs the reference in fp32 (the exact value): |d log-likelihood| <= 1e-2 for every pair -- the north-star
        tolerance, asserted as a fixed bound;
      * engine vs the reference in bf16 (autocast, sdpa, batch 16 -- the reference's own reduced-precision path): that
        run itself sits up to 5e-2 from the reference's fp32 run (profiles/r02_parity_7b_*.json), so the bound here is
        what two correct 16-bit pipelines can differ by: 8e-2 (VTG kinds) / 2.5e-2 (TVG kinds);
      * the engine is closer to the exact value than the reference's bf16 run is, per kind, in the mean;
      * reranked candidate order of the fused BLiM scores vs the fp32 reference: only candidates whose exact fused scores
        are closer than 2e-2 may swap, ground-truth ranks / top-1 agree wherever the exact top-1 margin exceeds 2e-2."""
    from oracle import ref_gpu, ref_harness
    cfg, model, weights = seven_b
    rows, topk = 16, 16
    corpus = synth.make_corpus(cfg, "msrvtt", n=1000, seed=5)
    _set_corpus(model, corpus)
    m_eng = ref_gpu.engine_matrices(model.engine, corpus, 0, rows, topk)
    if ref_harness.reference_available():
        rr = ref_gpu.ReferenceRunner(cfg, weights, corpus, "cuda:0", dtype=torch.bfloat16)
        m_b16, secs, pairs = rr.all_matrices(0, rows, topk)
        rr.close()
        del rr
        rr = ref_gpu.ReferenceRunner(cfg, weights, corpus, "cuda:0", dtype=torch.float32)
        m_f32, _, _ = rr.all_matrices(0, rows, topk)
        rr.close()
        del rr
        print(f"reference (bf16, sdpa, batch 16) on this GPU: {pairs / secs:.1f} pairs/s")
    else:   # no shipped reference: the oracle restatement (pinned to the reference by tests/golden) is the fp32 comparator
        m_b16 = None
        p = {k: v.float() for k, v in weights.items()}
        m_f32 = {}
        for name, (direction, ft, cpn) in ref_gpu.MATRICES.items():
            with torch.no_grad():
                dense = O.compute_scores_x(p, cfg, corpus, direction, ft, cpn, topk=topk, batch_size=16, rows=range(rows), device="cuda").numpy()
            idx = m_eng[name][0]
            m_f32[name] = (idx, np.take_along_axis(dense[:rows], idx, 1))
        del p
    torch.cuda.empty_cache()
    worst = {}
    for name in ref_gpu.MATRICES:
        e, f = m_eng[name][1], m_f32[name][1]
        assert np.isfinite(e).all()
        row = {"engine_vs_fp32": float(np.abs(e - f).max()), "engine_vs_fp32_mean": float(np.abs(e - f).mean())}
        if m_b16 is not None:
            b = m_b16[name][1]
            row.update(engine_vs_bf16=float(np.abs(e - b).max()), bf16_vs_fp32=float(np.abs(b - f).max()), bf16_vs_fp32_mean=float(np.abs(b - f).mean()))
        worst[name] = {k: round(v, 5) for k, v in row.items()}
    for name, row in worst.items():
        print("7B", name, row)
    for name, row in worst.items():
        tvg = ref_gpu.MATRICES[name][1] == "tvg"
        assert row["engine_vs_fp32"] <= 1e-2, f"{name}: engine is {row['engine_vs_fp32']} from the fp32 reference"
        if m_b16 is not None:
            assert row["engine_vs_bf16"] <= (2.5e-2 if tvg else 8e-2), f"{name}: engine is {row['engine_vs_bf16']} from the bf16 reference"
            assert row["engine_vs_fp32_mean"] <= row["bf16_vs_fp32_mean"], f"{name}: the reference's bf16 run is closer to fp32 than the engine"
    alpha, c = (0.0, 0.8), (1.0, 0.6, 0.8, 0.4)
    rp = ref_gpu.rank_parity(ref_gpu.fused_rows(m_f32, corpus, 0, rows, alpha, c), ref_gpu.fused_rows(m_eng, corpus, 0, rows, alpha, c))
    print("7B rank parity vs the fp32 reference:", rp)
    for d, r in rp.items():
        assert r["max_gap_of_swapped_pairs_a"] <= 2e-2, (d, r)
        if r["min_top1_margin_a"] > 2e-2:
            assert r["rows_same_top1"] == r["rows"] and r["recall_equal"], (d, r)


def test_7b_rebatching_invariance_and_symmetry(seven_b):
    cfg, model, weights = seven_b
    corpus = synth.make_corpus(cfg, "msrvtt", n=96, seed=9, feat_device="cuda")
    _set_corpus(model, corpus)
    from blim_b200.retrieval import PairPlan, score_all, compact_terms
    plan = PairPlan(corpus.v2t_iv2, corpus.t2v_iv2, 8, model.device)
    s = score_all(model, plan, cpn=True, full=True)
    uv, ut = plan.union_v.cpu().numpy(), plan.union_t.cpu().numpy()
    rng = np.random.default_rng(0)
    sel = rng.permutation(len(uv))[:40]
    for kind, key in ((VTG, "vtg"), (TVG, "tvg")):
        again = model.engine.score_pairs(kind, uv[sel], ut[sel]).cpu().numpy()     # different run composition, reversed order
        full = s[key].cpu().numpy()[sel]
        # the GEMMs are row-independent (bit-identical under re-batching); the tcgen05 attention aligns its 64-key chunks
        # to the first sequence of each 128-row block, so the bf16 rounding of P can differ with the block composition
        assert np.abs(again - full).max() <= 5e-3, f"{key}: scores depend on batch composition, max |d| {np.abs(again - full).max()}"
    # direction symmetry through the dedupe: v2t[v,t] and t2v[t,v] are the same number wherever both exist
    t2v_c, v2t_c = compact_terms(plan, s)
    a = {(int(v), int(t)): float(x) for v, row, xs in zip(range(corpus.n), plan.v2t_idx.cpu().numpy(), v2t_c["candidate_likelihood"].cpu().numpy()) for t, x in zip(row, xs)}
    n_both = 0
    for t, row, xs in zip(range(corpus.n), plan.t2v_idx.cpu().numpy(), t2v_c["query_likelihood"].cpu().numpy()):
        for v, x in zip(row, xs):
            if (int(v), t) in a:
                n_both += 1
                assert a[(int(v), t)] == float(x)
    assert n_both > 0
    # the VTG prior is a function of the text only
    pr = v2t_c["candidate_prior"].cpu().numpy()
    idx = plan.v2t_idx.cpu().numpy()
    by_text = {}
    for r in range(corpus.n):
        for t, x in zip(idx[r], pr[r]):
            by_text.setdefault(int(t), set()).add(float(x))
    assert all(len(v) == 1 for v in by_text.values())
