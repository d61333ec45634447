"""Pins oracle/blim_oracle.py (the CPU restatement) against fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py -> tests/golden/*.npz).  CPU only."""
import os

import numpy as np
import pytest
import torch

from oracle import blim_oracle as O
from oracle.make_golden import CASES, build_case

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
PAIRS = [("v2t", "vtg", False), ("v2t", "vtg", True), ("v2t", "tvg", False), ("v2t", "tvg", True),
         ("t2v", "vtg", False), ("t2v", "vtg", True), ("t2v", "tvg", False), ("t2v", "tvg", True)]


@pytest.fixture(scope="module", params=sorted(CASES))
def case(request):
    name = request.param
    cfg, weights, corpus = build_case(CASES[name])
    gold = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    p = {k: v.float() for k, v in weights.items()}
    return name, CASES[name], cfg, p, corpus, gold


def test_oracle_matrices_match_reference(case):
    name, spec, cfg, p, corpus, gold = case
    with torch.no_grad():
        for direction, ft, cpn in PAIRS:
            got = O.compute_scores_x(p, cfg, corpus, direction, ft, cpn, spec["topk"], spec["bs"]).numpy()
            ref = gold[f"{direction}_{ft}_{'cpn' if cpn else 'lik'}"]
            assert (got == -100.0).sum() == (ref == -100.0).sum()
            np.testing.assert_allclose(got, ref, atol=2e-5, rtol=0, err_msg=f"{name} {direction} {ft} cpn={cpn}")


def test_oracle_fuse_and_recall_bit_exact(case):
    name, spec, cfg, p, corpus, gold = case
    t2v = {"candidate_likelihood": gold["t2v_tvg_lik"], "query_likelihood": gold["t2v_vtg_lik"], "internvideo2": corpus.t2v_iv2.numpy(),
           "candidate_prior": gold["t2v_tvg_cpn"]}
    v2t = {"candidate_likelihood": gold["v2t_vtg_lik"], "query_likelihood": gold["v2t_tvg_lik"], "internvideo2": corpus.v2t_iv2.numpy(),
           "candidate_prior": gold["v2t_vtg_cpn"]}
    bt, bv, ct, cv = O.fuse(t2v, v2t, spec["alpha"], spec["c"], cpn=True, zero_shot=False)
    for got, key in ((bt, "blim_t2v"), (bv, "blim_v2t"), (ct, "cpn_t2v"), (cv, "cpn_v2t")):
        assert got.dtype == gold[key].dtype
        assert np.array_equal(got, gold[key]), key
    zt, zv, _, _ = O.fuse(t2v, v2t, spec["alpha"], spec["c"], cpn=True, zero_shot=True)
    assert zt.dtype == np.float64 and np.array_equal(zt, gold["zs_t2v"])
    assert np.array_equal(zv, gold["zs_v2t"])
    keys = list(gold["recall_keys"])
    for mats, key in (((bt, bv), "recall_blim"), ((ct, cv), "recall_cpn"), ((zt, zv), "recall_zs"),
                      ((t2v["candidate_likelihood"], v2t["candidate_likelihood"]), "recall_cand")):
        res = O.get_recall(*mats)[0]
        assert [res[k] for k in keys] == list(gold[key]), key


def test_reference_invariants(case):
    """Invariants the engine's work sharing relies on, checked on the reference's own outputs."""
    name, spec, cfg, p, corpus, gold = case
    # (1) the VTG prior does not depend on the video
    m = gold["v2t_vtg_cpn"]
    for t in range(m.shape[1]):
        col = m[:, t][m[:, t] != -100.0]
        if len(col) > 1:
            assert np.ptp(col) < 1e-5
    # (2) v2t[v, t] and t2v[t, v] hold the same quantity wherever both were computed
    for a, b in (("v2t_vtg_lik", "t2v_vtg_lik"), ("v2t_tvg_lik", "t2v_tvg_lik"), ("v2t_tvg_cpn", "t2v_tvg_cpn")):
        x, y = gold[a], gold[b].T
        both = (x != -100.0) & (y != -100.0)
        assert both.sum() > 0
        assert np.abs(x[both] - y[both]).max() < 2e-5


def test_reference_runner_reproduces_goldens_and_rank_parity_helpers():
    """oracle/ref_gpu.py (the driver of the UNMODIFIED reference used as comparator at 7B and by bench.py) on the CPU: its
    matrices are the goldens bit for bit, the fused rows equal the goldens' BLiM matrices, and rank_parity reports a perfect
    match of a result with itself and a swap when two close candidates are exchanged."""
    from oracle import ref_gpu, ref_harness
    if not ref_harness.reference_available():
        pytest.skip("reference tree not present")
    from oracle.make_golden import CASES, build_case
    case = CASES["tiny_a"]
    cfg, w, corpus = build_case(case)
    g = np.load(os.path.join(GOLDEN, "tiny_a.npz"))
    rr = ref_gpu.ReferenceRunner(cfg, {k: v.float() for k, v in w.items()}, corpus, "cpu", dtype=torch.float32)
    mats, secs, pairs = rr.all_matrices(0, corpus.n, case["topk"], case["bs"])
    assert pairs == 2 * corpus.n * case["topk"]
    key = {"v2t_candidate_likelihood": "v2t_vtg_lik", "v2t_candidate_prior": "v2t_vtg_cpn", "v2t_query_likelihood": "v2t_tvg_lik",
           "t2v_query_likelihood": "t2v_vtg_lik", "t2v_candidate_likelihood": "t2v_tvg_lik", "t2v_candidate_prior": "t2v_tvg_cpn"}
    for name, (idx, sc) in mats.items():
        assert np.array_equal(np.take_along_axis(g[key[name]], idx, 1), sc), name
    f = ref_gpu.fused_rows(mats, corpus, 0, corpus.n, case["alpha"], case["c"])
    assert np.array_equal(f["t2v"][0], g["blim_t2v"]) and np.array_equal(f["v2t"][0], g["blim_v2t"])
    same = ref_gpu.rank_parity(f, f)
    for d in ("t2v", "v2t"):
        assert same[d]["rows_same_order"] == corpus.n and same[d]["recall_equal"] and same[d]["swapped_adjacent_pairs"] == 0
    # exchange the scores of the two best candidates of row 0 in one matrix: exactly that row's order changes
    swapped = {k: (i.copy(), s.copy()) for k, (i, s) in mats.items()}
    o = f["t2v"][1][0]                                         # candidate ids of row 0, best first
    idx, sc = swapped["t2v_query_likelihood"]
    a, b = int(np.where(idx[0] == o[0])[0][0]), int(np.where(idx[0] == o[1])[0][0])
    sc[0, a], sc[0, b] = sc[0, b] - 1.0, sc[0, a] + 1.0        # push them past each other
    f2 = ref_gpu.fused_rows(swapped, corpus, 0, corpus.n, case["alpha"], case["c"])
    rp = ref_gpu.rank_parity(f, f2)
    assert rp["t2v"]["rows_same_order"] == corpus.n - 1 and rp["t2v"]["swapped_adjacent_pairs"] >= 1 and rp["v2t"]["rows_same_order"] == corpus.n
