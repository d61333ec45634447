"""evaluation() / val_one_epoch() drop-ins end to end (host inputs -> result table) against the oracle's numpy
val_one_epoch arithmetic applied to the SAME score matrices: recalls must be identical.  GPU only."""
import argparse
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import bench
from blim_b200 import evalloop, retrieval
from blim_b200.model import BlimModel
from oracle import blim_oracle as O
from oracle.make_golden import CASES, build_case

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("zero_shot,cpn", [(False, True), (False, False), (True, True)])
def test_val_one_epoch_table(name, zero_shot, cpn):
    spec = CASES[name]
    cfg, weights, corpus = build_case(spec)
    gold = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    model = BlimModel(cfg, state_dict=weights, device=0, max_run_tokens=4096, max_prefix_tokens=4096)
    try:
        loader = bench.Loader(corpus, batch=5)
        args = argparse.Namespace(topk=spec["topk"], batch_size_eval=spec["bs"], num_clips=corpus.n_clips, cpn=cpn, eval=True,
                                  resume="" if zero_shot else "ckpt", dataset="msrvtt", distributed=False, alpha=list(spec["alpha"]),
                                  c=list(spec["c"]), iv2_scores={"v2t": corpus.v2t_iv2, "t2v": corpus.t2v_iv2})
        t2v_dict, v2t_dict = retrieval.evaluation(model, loader, model.device, None, args)
        # dictionaries have the reference's keys / shapes / fill and agree with the reference's matrices
        want_keys = {"query_likelihood", "internvideo2"} | ({"candidate_likelihood"} if not zero_shot else set()) | \
            ({"candidate_prior"} if (cpn and not zero_shot) else set())
        assert set(t2v_dict) == want_keys
        assert set(v2t_dict) == {"candidate_likelihood", "internvideo2"} | ({"query_likelihood"} if not zero_shot else set()) | \
            ({"candidate_prior"} if cpn else set())
        for got, ref_key in ((v2t_dict["candidate_likelihood"], "v2t_vtg_lik"), (t2v_dict["query_likelihood"], "t2v_vtg_lik")):
            ref = gold[ref_key]
            assert got.dtype == np.float32 and ((got == -100.0) == (ref == -100.0)).all()
            assert np.abs(got - ref).max() <= 1e-2
        res = evalloop.val_one_epoch(model, loader, None, model.device, 0, None, tokenizer=None, args=args)
        t2v_dict, v2t_dict = retrieval.evaluation(model, loader, model.device, None, args)   # same matrices again (deterministic)
        want = O.val_results(t2v_dict, v2t_dict, spec["alpha"], spec["c"], cpn=cpn, zero_shot=zero_shot)
        assert res == want, (res, want)
    finally:
        model.engine.close()
