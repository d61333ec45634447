"""Parity at the shapes of BASELINE.json configs[2..4] on a mid-size model (head_dim 128):
DiDeMo-like long captions, ActivityNet-like videos with 16 clips (1024 visual tokens per prefix, prefix cache streamed
in small groups) and an LSMDC-like top-64 candidate list.  All four score kinds, both directions, against the fp32
oracle (the CPU restatement pinned to the reference by tests/golden).  GPU only."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from blim_b200 import synth
from blim_b200.engine import TVG, TVG_PRIOR, VTG, VTG_PRIOR, ModelConfig
from blim_b200.model import BlimModel
from oracle import blim_oracle as O

KIND = {("vtg", False): VTG, ("vtg", True): VTG_PRIOR, ("tvg", False): TVG, ("tvg", True): TVG_PRIOR}
SHAPES = {
    # name: (dataset shape, n, n_clips, caption mean / std, topk, oracle rows, engine workspace (run tokens, prefix rows))
    "c3_didemo": ("didemo", 10, 4, 45, 15, 6, [0, 7], (4096, 4096)),
    "c4_activitynet_16clips": ("activitynet", 6, 16, 60, 20, 4, [1, 4], (4096, 2304)),     # 2 x 1050-token prefixes per prefix run
    "c5_lsmdc_top64": ("lsmdc", 72, 4, 15, 6, 64, [3], (8192, 8192)),
}


def _mid_cfg():
    return ModelConfig(hidden_size=512, num_layers=3, num_heads=4, num_kv_heads=2, intermediate_size=1536, vocab_size=8192,
                       mm_hidden_size=1024, max_positions=2048, image_token_id=8000)


@pytest.mark.parametrize("name", sorted(SHAPES))
def test_config_shapes_match_oracle(name):
    dataset, n, n_clips, cap_mean, cap_std, topk, rows, (run_tokens, prefix_rows) = SHAPES[name]
    cfg = _mid_cfg()
    weights = synth.init_weights(cfg, seed=7, std=0.05, rich=True)
    corpus = synth.make_corpus(cfg, dataset, n=n, n_clips=n_clips, cap_mean=cap_mean, cap_std=cap_std, seed=3)
    model = BlimModel(cfg, state_dict=weights, device=0, max_run_tokens=run_tokens, max_prefix_tokens=prefix_rows)
    try:
        eng = model.engine
        eng.set_videos(corpus.video)
        eng.set_texts(0, corpus.vtg_ids, corpus.vtg_labels)
        eng.set_texts(1, corpus.tvg_ids, corpus.tvg_labels)
        eng.set_video_vocab(corpus.video_vocab, corpus.tvg_video_labels.numpy())
        model.set_tvg_prefix_length(corpus.tvg_prefix_length)
        p = {k: v.float() for k, v in weights.items()}
        worst = {}
        for direction in ("v2t", "t2v"):
            for (ft, cpn), kind in KIND.items():
                with torch.no_grad():
                    want = O.compute_scores_x(p, cfg, corpus, direction, ft, cpn, topk=topk, batch_size=8, rows=rows).numpy()
                r, c = np.nonzero(want != -100.0)
                assert len(r) == len(rows) * min(topk, n)
                pv, pt = (r, c) if direction == "v2t" else (c, r)
                got = eng.score_pairs(kind, pv, pt).cpu().numpy()
                worst[(direction, ft, cpn)] = float(np.abs(got - want[r, c]).max())
        print(name, {k: round(v, 5) for k, v in worst.items()})
        assert max(worst.values()) <= 1e-2, worst
    finally:
        model.engine.close()
