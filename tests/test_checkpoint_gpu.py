"""SURVEY.md 8(f) rank 1 through the engine: a checkpoint with the reference's key layout (tests/ckpt_fixture.py) is merged,
loaded with load_finetuned and scored; the comparator is the oracle run on the fine-tuned model's parameters written out
by hand in fp32 (base + alpha/r * B @ A per adapted Linear, tvg_mlp = copy of the base mlp + its own adapters, fp32
visual_head) -- i.e. PEFT's adapter forward.  GPU only."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from blim_b200 import synth
from blim_b200.checkpoint import load_finetuned
from blim_b200.engine import TVG, TVG_PRIOR, VTG, VTG_PRIOR, ModelConfig
from blim_b200.model import BlimModel
from oracle import blim_oracle as O
from tests import ckpt_fixture as F

KIND = {("vtg", False): VTG, ("vtg", True): VTG_PRIOR, ("tvg", False): TVG, ("tvg", True): TVG_PRIOR}


@pytest.mark.parametrize("with_random_tvg_mlp", [False, True])
def test_finetuned_checkpoint_scores_match_adapter_forward(tmp_path, with_random_tvg_mlp):
    cfg = ModelConfig.tiny()
    base = F.base_state_dict(cfg, with_random_tvg_mlp=with_random_tvg_mlp)
    ckpt, deltas = F.make_reference_checkpoint(cfg)
    path = tmp_path / "checkpoint_best.pth"
    torch.save(ckpt, path)
    corpus = synth.make_corpus(cfg, "msrvtt", n=10, n_clips=4, cap_mean=7, cap_std=2, seed=4)
    model = BlimModel(cfg, device=0, max_run_tokens=4096, max_prefix_tokens=4096)
    try:
        merged, ignored = load_finetuned(model, base, str(path), ckpt["args"].lora_r, ckpt["args"].lora_alpha)
        assert len(merged) == len(F.adapted_linears(cfg)) and not ignored
        eng = model.engine
        eng.set_videos(corpus.video)
        eng.set_texts(0, corpus.vtg_ids, corpus.vtg_labels)
        eng.set_texts(1, corpus.tvg_ids, corpus.tvg_labels)
        eng.set_video_vocab(corpus.video_vocab, corpus.tvg_video_labels.numpy())
        model.set_tvg_prefix_length(corpus.tvg_prefix_length)
        p = F.expected_merged(cfg, base, ckpt, deltas)
        p_base = {k: v.float() for k, v in base.items()}
        moved = 0.0
        for (ft, cpn), kind in KIND.items():
            with torch.no_grad():
                want = O.compute_scores_x(p, cfg, corpus, "v2t", ft, cpn, topk=4, batch_size=4).numpy()
                if "model.mm_projector.tvg_mlp.0.weight" in p_base:
                    plain = O.compute_scores_x(p_base, cfg, corpus, "v2t", ft, cpn, topk=4, batch_size=4).numpy()
                    moved = max(moved, float(np.abs(plain - want)[want != -100.0].max()))
            rows, cols = np.nonzero(want != -100.0)
            got = eng.score_pairs(kind, rows, cols).cpu().numpy()
            err = float(np.abs(got - want[rows, cols]).max())
            assert err <= 1e-2, f"{ft} cpn={cpn}: engine(merged) vs oracle(adapter forward) |d|={err}"
        if with_random_tvg_mlp:
            assert moved > 5e-2, "the adapters must change the scores, or the test proves nothing"
    finally:
        model.engine.close()
