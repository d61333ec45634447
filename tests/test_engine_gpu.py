"""CUDA engine (through the C ABI) vs the oracle and the reference-generated golden fixtures.  GPU only.

Tolerance: BASELINE.json north_star -- per-pair log-likelihoods within 1e-2 absolute (bf16 pipeline vs fp32 reference);
integer / index outputs (ranks, orders) bit-exact."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from blim_b200 import synth
from blim_b200.engine import TVG, TVG_PRIOR, VTG, VTG_PRIOR, Engine, ModelConfig
from oracle import blim_oracle as O
from oracle.make_golden import CASES, build_case

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-2
KIND = {("vtg", False): VTG, ("vtg", True): VTG_PRIOR, ("tvg", False): TVG, ("tvg", True): TVG_PRIOR}


def make_engine(cfg, weights, corpus, **kw):
    eng = Engine(cfg, **kw)
    eng.load_state_dict(weights)
    eng.set_videos(corpus.video)
    eng.set_texts(0, corpus.vtg_ids, corpus.vtg_labels)
    eng.set_texts(1, corpus.tvg_ids, corpus.tvg_labels)
    eng.set_video_vocab(corpus.video_vocab, corpus.tvg_video_labels.numpy())
    eng.set_tvg_prefix_length(corpus.tvg_prefix_length)
    return eng


@pytest.fixture(scope="module", params=[(n, g) for n in sorted(CASES) for g in (1, 2)], ids=lambda p: f"{p[0]}-cg{p[1]}")
def golden_case(request):
    name, cg = request.param
    cfg, weights, corpus = build_case(CASES[name])
    eng = make_engine(cfg, weights, corpus, max_run_tokens=4096, max_prefix_tokens=4096, gemm_cta_group=cg)
    yield name, CASES[name], cfg, weights, corpus, eng, np.load(os.path.join(GOLDEN, f"{name}.npz"))
    eng.close()


@pytest.mark.parametrize("direction", ["v2t", "t2v"])
@pytest.mark.parametrize("ft,cpn", [("vtg", False), ("vtg", True), ("tvg", False), ("tvg", True)])
def test_scores_match_reference_golden(golden_case, direction, ft, cpn):
    name, spec, cfg, weights, corpus, eng, gold = golden_case
    ref = gold[f"{direction}_{ft}_{'cpn' if cpn else 'lik'}"]
    rows, cols = np.nonzero(ref != -100.0)
    pv, pt = (rows, cols) if direction == "v2t" else (cols, rows)
    got = eng.score_pairs(KIND[(ft, cpn)], pv, pt).cpu().numpy()
    want = ref[rows, cols]
    err = np.abs(got - want)
    assert np.isfinite(got).all(), got
    assert err.max() <= TOL, f"{name} {direction} {ft} cpn={cpn}: max |d|={err.max():.4g} at pair {int(err.argmax())}: got {got[:6]} want {want[:6]}"


def test_small_workspace_batches_agree(golden_case):
    """Tiny workspaces force many prefix/suffix batches and split units: results must not depend on the batching."""
    name, spec, cfg, weights, corpus, eng, gold = golden_case
    small = make_engine(cfg, weights, corpus, max_run_tokens=320, max_prefix_tokens=600)
    try:
        ref = gold["v2t_vtg_lik"]
        rows, cols = np.nonzero(ref != -100.0)
        for kind, pv, pt in ((VTG, rows, cols), (VTG_PRIOR, rows, cols), (TVG, rows, cols), (TVG_PRIOR, rows, cols)):
            a = eng.score_pairs(kind, pv, pt).cpu().numpy()
            b = small.score_pairs(kind, pv, pt).cpu().numpy()
            assert np.abs(a - b).max() < 2e-3, (kind, np.abs(a - b).max())
    finally:
        small.close()


def test_sharded_scores_are_bit_identical(golden_case):
    """What the ranks of a multi-GPU run score (retrieval.score_all: prefix owners assigned to ranks, strided or balanced):
    every score of every kind must be BIT-identical to the unsharded run -- attention tiles never stack sequences of
    different scheduling units, so a unit's numbers do not depend on what else shares its run."""
    from blim_b200 import retrieval
    name, spec, cfg, weights, corpus, eng, gold = golden_case
    ref = gold["v2t_vtg_lik"]
    rows, cols = np.nonzero(ref != -100.0)
    jobs = [("vtg", VTG, rows, cols), ("vtg_prior", VTG_PRIOR, rows, cols), ("tvg", TVG, rows, cols), ("tvg_prior", TVG_PRIOR, rows, cols)]
    full = {name: eng.score_pairs(kind, pv, pt).cpu().numpy() for name, kind, pv, pt in jobs}
    for world in (2, 3, 8):
        sp = retrieval.ShardPlan(eng, jobs, world, corpus.n, corpus.n)     # owners balanced on the summed cost of the four kinds
        strided = {name: [np.nonzero((pv if kind in (VTG, TVG_PRIOR) else pt) % world == r)[0] for r in range(world)] for name, kind, pv, pt in jobs}
        for shards in (sp.shards, strided):
            for name, kind, pv, pt in jobs:
                got = np.empty_like(full[name])
                seen = np.zeros(len(pv), dtype=int)
                for r in range(world):
                    mine = shards[name][r]
                    seen[mine] += 1
                    if len(mine):
                        got[mine] = eng.score_pairs(kind, pv[mine], pt[mine]).cpu().numpy()
                assert (seen == 1).all()
                assert np.array_equal(got, full[name]), (name, world, np.abs(got - full[name]).max())


def _mid_cfg():
    return ModelConfig(hidden_size=512, num_layers=3, num_heads=4, num_kv_heads=2, intermediate_size=1536, vocab_size=8192,
                       mm_hidden_size=1024, max_positions=2048, image_token_id=8000)


@pytest.mark.parametrize("cfg_name,cg", [("tiny", 1), ("tiny", 2), ("mid", 1), ("mid", 2)])
def test_forward_logits_matches_oracle(cfg_name, cg):
    cfg = ModelConfig.tiny() if cfg_name == "tiny" else _mid_cfg()
    weights = synth.init_weights(cfg, seed=5, std=0.05, rich=True)
    eng = Engine(cfg, max_run_tokens=2048, max_prefix_tokens=256, gemm_cta_group=cg)
    try:
        eng.load_state_dict(weights, rope_table_dtype=torch.float32)
        g = torch.Generator().manual_seed(11)
        B, L = 3, 150
        emb = (torch.randn(B, L, cfg.hidden_size, generator=g) * 0.05).bfloat16()
        mask = torch.ones(B, L, dtype=torch.long)
        mask[0, 20:90] = 0          # CPN-style hole
        mask[1, 120:] = 0           # right padding
        mask[2, 5] = 0
        logits, hidden = eng.forward_logits(emb.cuda(), mask.cuda())
        torch.cuda.synchronize()
        p = {k: v.float().cuda() for k, v in weights.items()}
        with torch.no_grad():
            ref_logits, ref_hidden = O.decoder_forward(p, cfg, emb.float().cuda(), mask.cuda())
        eh = (hidden.float() - ref_hidden).abs().max().item()
        el = (logits - ref_logits).abs().max().item()
        scale_h, scale_l = ref_hidden.abs().max().item(), ref_logits.abs().max().item()
        assert torch.isfinite(logits).all()
        assert eh <= 3e-2 * scale_h and el <= 3e-2 * scale_l, f"hidden err {eh} (scale {scale_h}), logits err {el} (scale {scale_l})"
        # log-softmax of the logits is what the scores are made of
        lp = (torch.log_softmax(logits, -1) - torch.log_softmax(ref_logits, -1)).abs().max().item()
        assert lp < 5e-2 * max(1.0, scale_l), (lp, scale_l)
    finally:
        eng.close()


def test_scores_match_oracle_mid_headdim128():
    cfg = _mid_cfg()
    weights = synth.init_weights(cfg, seed=9, std=0.04, rich=True)
    corpus = synth.make_corpus(cfg, "didemo", n=8, n_clips=4, cap_mean=20, cap_std=8, seed=21)
    eng = make_engine(cfg, weights, corpus, max_run_tokens=4096, max_prefix_tokens=4096)
    try:
        p = {k: v.float().cuda() for k, v in weights.items()}
        for direction in ("v2t", "t2v"):
            for ft, cpn in (("vtg", False), ("vtg", True), ("tvg", False), ("tvg", True)):
                with torch.no_grad():
                    ref = O.compute_scores_x(p, cfg, corpus, direction, ft, cpn, topk=3, batch_size=2, device="cuda").numpy()
                rows, cols = np.nonzero(ref != -100.0)
                pv, pt = (rows, cols) if direction == "v2t" else (cols, rows)
                got = eng.score_pairs(KIND[(ft, cpn)], pv, pt).cpu().numpy()
                err = np.abs(got - ref[rows, cols]).max()
                assert err <= TOL, f"{direction} {ft} cpn={cpn}: {err}"
    finally:
        eng.close()


def test_projector_and_heads_match_oracle():
    cfg = ModelConfig.tiny()
    weights = synth.init_weights(cfg, seed=2, std=0.05, rich=True)
    eng = Engine(cfg, max_run_tokens=2048, max_prefix_tokens=256)
    try:
        eng.load_state_dict(weights)
        p = {k: v.float().cuda() for k, v in weights.items()}
        feats = (torch.randn(300, cfg.mm_hidden_size, device="cuda") * 0.5).bfloat16()
        for tvg in (False, True):
            got = eng.project_video(feats, tvg=tvg).float()
            ref = O.project_video(p, feats.float(), tvg)
            assert (got - ref).abs().max() <= 2e-2 * ref.abs().max()
        hid = torch.randn(77, cfg.hidden_size, device="cuda").bfloat16()
        got = eng.forward_visual(hid).float()
        ref = hid.float() @ p["visual_head.weight"].t()
        assert (got - ref).abs().max() <= 2e-2 * ref.abs().max()
        ids = torch.randint(0, cfg.vocab_size, (5, 7), device="cuda")
        assert torch.equal(eng.embed_tokens(ids).float(), p["model.embed_tokens.weight"][ids])
    finally:
        eng.close()


# ---------------------------------------------------------------------------------------------- fuse + rerank: bit exact
def _compact(mat, idx):
    return np.take_along_axis(mat, idx, axis=1)


@pytest.mark.parametrize("zero_shot", [False, True])
def test_fuse_rerank_bit_exact_on_reference_scores(golden_case, zero_shot):
    name, spec, cfg, weights, corpus, eng, gold = golden_case
    n, k = corpus.n, spec["topk"]
    t2v = {"candidate_likelihood": gold["t2v_tvg_lik"], "query_likelihood": gold["t2v_vtg_lik"], "internvideo2": corpus.t2v_iv2.numpy(),
           "candidate_prior": gold["t2v_tvg_cpn"]}
    v2t = {"candidate_likelihood": gold["v2t_vtg_lik"], "query_likelihood": gold["v2t_tvg_lik"], "internvideo2": corpus.v2t_iv2.numpy(),
           "candidate_prior": gold["v2t_vtg_cpn"]}
    bt, bv, _, _ = O.fuse(t2v, v2t, spec["alpha"], spec["c"], cpn=True, zero_shot=zero_shot)
    _, tr, vr = O.get_recall(bt, bv)
    a, c = spec["alpha"], spec["c"]
    for d, mats, blim, ranks, alpha, cq, ce in (("t2v", t2v, bt, tr, a[0], c[0], c[2]), ("v2t", v2t, bv, vr, a[1], c[1], c[3])):
        iv2 = mats["internvideo2"]
        idx = np.ascontiguousarray(torch.from_numpy(iv2).topk(k, dim=1).indices.numpy())
        fused, order, rank, zero = eng.fuse_rerank(
            torch.from_numpy(idx), torch.from_numpy(_compact(mats["candidate_likelihood"], idx)),
            torch.from_numpy(_compact(mats["candidate_prior"], idx)), torch.from_numpy(_compact(mats["query_likelihood"], idx)),
            torch.from_numpy(iv2), alpha, cq, ce, use_prior=True,
            use_query=not (zero_shot and d == "v2t"), cpn_zero_f64=(zero_shot and d == "t2v"))
        fused, order, rank = fused.cpu().numpy(), order.cpu().numpy(), rank.cpu().numpy()
        want = _compact(blim, idx).astype(np.float64)
        assert np.array_equal(fused, want), f"{d}: fused scores differ"
        assert np.array_equal(rank, ranks.astype(np.int64)), f"{d}: ranks differ"
        want_order = np.take_along_axis(idx, np.argsort(-want, axis=1, kind="stable"), axis=1)
        assert np.array_equal(order, want_order), f"{d}: candidate order differs"
        assert int(zero.item()) == 0
        dense_rank, _ = eng.rank_dense(torch.from_numpy(blim.astype(np.float32)))
        if blim.dtype == np.float32:
            assert np.array_equal(dense_rank.cpu().numpy(), ranks.astype(np.int64))


def test_fuse_rerank_random_large():
    eng = Engine(ModelConfig.tiny(), max_run_tokens=512, max_prefix_tokens=256)
    try:
        rng = np.random.default_rng(0)
        n, k = 1000, 16
        iv2 = (rng.standard_normal((n, n)) + 3 * np.eye(n)).astype(np.float32)
        idx = np.ascontiguousarray(torch.from_numpy(iv2).topk(k, dim=1).indices.numpy())
        mats = {}
        for key in ("candidate_likelihood", "candidate_prior", "query_likelihood"):
            m = np.full((n, n), -100.0, dtype=np.float32)
            np.put_along_axis(m, idx, (rng.standard_normal((n, k)) * 0.3 - 3).astype(np.float32), axis=1)
            mats[key] = m
        alpha, cq, ce = 0.9, 0.3, 0.8
        cpn = mats["candidate_likelihood"] - alpha * mats["candidate_prior"]
        blim = ce * (cq * mats["query_likelihood"] + (1 - cq) * cpn) + (1 - ce) * iv2
        ranks = np.array([np.where(np.argsort(r)[::-1] == i)[0][0] for i, r in enumerate(blim)])
        fused, order, rank, zero = eng.fuse_rerank(torch.from_numpy(idx), torch.from_numpy(_compact(mats["candidate_likelihood"], idx)),
                                                   torch.from_numpy(_compact(mats["candidate_prior"], idx)),
                                                   torch.from_numpy(_compact(mats["query_likelihood"], idx)), torch.from_numpy(iv2), alpha, cq, ce)
        assert np.array_equal(fused.cpu().numpy(), _compact(blim, idx).astype(np.float64))
        assert np.array_equal(rank.cpu().numpy(), ranks)
        dense = eng.scatter_scores(n, n, torch.from_numpy(np.repeat(np.arange(n), k)), torch.from_numpy(idx.reshape(-1)),
                                   torch.from_numpy(_compact(mats["query_likelihood"], idx).reshape(-1)))
        assert np.array_equal(dense.cpu().numpy(), mats["query_likelihood"])
    finally:
        eng.close()


def test_topk_rows_matches_torch():
    eng = Engine(ModelConfig.tiny(), max_run_tokens=512, max_prefix_tokens=256)
    try:
        for n_rows, n_cols, k in ((7, 5, 5), (100, 1000, 16), (33, 4917, 64), (4, 77, 1)):
            g = torch.Generator(device="cuda").manual_seed(n_cols)
            m = torch.randn(n_rows, n_cols, generator=g, device="cuda")
            idx, val = eng.topk_rows(m, k)
            ref = m.topk(k, dim=1)
            assert torch.equal(val, ref.values) and torch.equal(idx, ref.indices), (n_rows, n_cols, k)
        m = torch.tensor([[1.0, 3.0, 3.0, 2.0, 3.0]], device="cuda")      # ties: lowest column first
        assert eng.topk_rows(m, 4)[0].tolist() == [[1, 2, 4, 3]]
    finally:
        eng.close()


def test_device_built_video_vocab_matches_uploaded(golden_case):
    """blim_build_video_vocab (mean over the 64 tokens on the device) gives the same TVG scores as the uploaded vocab, up to
    the rounding of the vocabulary itself: the uploaded one was rounded to bf16 on the host (synth.make_corpus), the
    device-built one goes straight to the engine's operand format."""
    name, spec, cfg, weights, corpus, eng, gold = golden_case
    ref = gold["t2v_tvg_lik"]
    rows, cols = np.nonzero(ref != -100.0)
    a = eng.score_pairs(TVG, cols, rows).cpu().numpy()
    eng.build_video_vocab(corpus.tvg_video_labels.numpy())
    b = eng.score_pairs(TVG, cols, rows).cpu().numpy()
    assert np.abs(a - b).max() <= 1e-3, np.abs(a - b).max()
    eng.set_video_vocab(corpus.video_vocab, corpus.tvg_video_labels.numpy())
