"""Device-side input pipeline: blim_b200.dataset.stage_corpus against the loader loop of
evaluation() on the same miniature data tree -- the staged path must reproduce the looped path (VTG bit-exactly; TVG
within the score tolerance, because its video vocabulary is reduced on the device from bf16 features instead of
being averaged in fp16 on the host), and both must agree with the oracle.  GPU only."""
import argparse

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from blim_b200 import dataset as D
from blim_b200 import retrieval, synth
from blim_b200.engine import ModelConfig
from blim_b200.model import BlimModel
from oracle import dataset_fixture
from oracle.stub_tokenizer import StubTokenizer


@pytest.mark.parametrize("name", ["MSRVTT", "DiDeMo"])
def test_staged_corpus_matches_loader_loop(tmp_path, name):
    root = str(tmp_path / "data")
    dataset_fixture.write(root)
    cfg = ModelConfig.tiny()
    # Qwen2 specials remapped into the tiny vocabulary the way blim_b200.synth does (image_token_id = <|im_end|> = 4000)
    tok = StubTokenizer(special={"<|im_start|>": cfg.image_token_id - 1, "<|im_end|>": cfg.image_token_id, "\n": 198}, lo=300, span=3000, pad=0)
    args = argparse.Namespace(dataset=name, topk=3, batch_size_eval=4, num_clips=4, cpn=True, eval=True, resume="ckpt", distributed=False)
    loader = D.load_data(args, tokenizer=tok, split="test", root=root)
    ds = loader.dataset
    n = len(ds)
    g = torch.Generator().manual_seed(5)
    t2v = torch.randn(n, n, generator=g) + 3.0 * torch.eye(n)
    args.iv2_scores = {"t2v": t2v, "v2t": t2v.t().contiguous() + 0.1 * torch.randn(n, n, generator=g)}
    weights = synth.init_weights(cfg, seed=2, std=0.05, rich=True)
    model = BlimModel(cfg, state_dict=weights, device=0, max_run_tokens=4096, max_prefix_tokens=4096)
    try:
        t2v_a, v2t_a = retrieval.evaluation(model, loader, model.device, tok, args)
        staged = D.stage_corpus(model, ds)
        assert staged.n == n and staged.n_clips == 4 and staged.tvg_video_labels.tolist() == [ds._vid_index[d["vid"]] for d in ds.data]
        assert ds.staged_corpus is staged
        t2v_b, v2t_b = retrieval.evaluation(model, loader, model.device, tok, args)
        for a, b in ((t2v_a, t2v_b), (v2t_a, v2t_b)):
            assert set(a) == set(b)
            for key in a:
                assert ((a[key] == -100.0) == (b[key] == -100.0)).all()
        # VTG terms (video -> text candidate likelihood and prior, text -> video query likelihood): same inputs, same bits
        np.testing.assert_array_equal(v2t_a["candidate_likelihood"], v2t_b["candidate_likelihood"])
        np.testing.assert_array_equal(v2t_a["candidate_prior"], v2t_b["candidate_prior"])
        np.testing.assert_array_equal(t2v_a["query_likelihood"], t2v_b["query_likelihood"])
        # TVG terms: vocabulary built on the device
        for key, d_a, d_b in (("candidate_likelihood", t2v_a, t2v_b), ("candidate_prior", t2v_a, t2v_b), ("query_likelihood", v2t_a, v2t_b)):
            assert np.abs(d_a[key] - d_b[key]).max() <= 1e-2, key
        # captions that share a video (video0 owns two) see the same vocabulary row
        assert len(ds.vids) == n - 1
    finally:
        model.engine.close()
