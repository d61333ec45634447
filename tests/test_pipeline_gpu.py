"""The §8(f) rows composed: frames -> feature extractor -> `.pth` files (extract.py flow) -> dataset classes -> staged
corpus -> evaluation() / val_one_epoch.  The features that reach the scoring engine through the files must be the ones
the extractor produced, and scoring them must give the same numbers as handing the tensors over directly.  GPU only."""
import argparse
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from blim_b200 import dataset as D
from blim_b200 import evalloop, extract, retrieval, synth
from blim_b200 import vision as V
from blim_b200.engine import ModelConfig
from blim_b200.model import BlimModel
from oracle.stub_tokenizer import StubTokenizer


def test_frames_to_recall_table(tmp_path):
    root = tmp_path / "data"
    n = 6
    vcfg = V.VisionConfig(image_size=96, hidden_size=1024, encoder_depth=3, num_heads=16)      # ViT-L width, 2 blocks run
    enc = V.VisionEncoder(vcfg, state_dict=V.init_weights(vcfg, seed=4), device=0, max_clips=6)
    g = torch.Generator().manual_seed(8)
    frames = {f"video{i}": torch.randn(8, 3, 96, 96, generator=g).to(torch.bfloat16) for i in range(n)}
    paths = [f"/videos/{v}.mp4" for v in frames] + ["/videos/broken.mp4"]

    def load_frames(path):
        vid = os.path.basename(path).split(".")[0]
        if vid not in frames:
            raise IOError("cannot decode")
        return frames[vid]

    feat_dir = root / "MSRVTT" / "features"
    try:
        a = extract.extract_dataset(enc, paths, load_frames, str(feat_dir), dataset="MSRVTT", num_chunk=2, chunk_idx=0, batch_videos=2)
        b = extract.extract_dataset(enc, paths, load_frames, str(feat_dir), dataset="MSRVTT", num_chunk=2, chunk_idx=1, batch_videos=2)
        assert sorted(a + b) == sorted(frames)                                   # the undecodable file is skipped
        direct = {v: enc.extract(f, out_dtype=torch.float16).cpu() for v, f in frames.items()}
    finally:
        enc.close()
    for v in frames:
        saved = torch.load(feat_dir / f"{v}.pth", weights_only=True)
        assert saved.dtype == torch.float16 and saved.shape == (2, 64, 1024)
        assert torch.equal(saved, direct[v])                                      # batching / chunking never changes a video

    caps = ["a man is talking about a car", "two dogs run across the field", "someone slices an onion", "the band starts to play",
            "a girl opens the door", "kids are playing football"]
    json.dump([{"video": f"video{i}.mp4", "caption": caps[i]} for i in range(n)], open(root / "MSRVTT" / "msrvtt_ret_test.json", "w"))
    cfg = ModelConfig.tiny()
    tok = StubTokenizer(special={"<|im_start|>": cfg.image_token_id - 1, "<|im_end|>": cfg.image_token_id, "\n": 198}, lo=300, span=3000, pad=0)
    gs = torch.Generator().manual_seed(5)
    t2v = torch.randn(n, n, generator=gs) + 3.0 * torch.eye(n)
    args = argparse.Namespace(dataset="MSRVTT", topk=3, batch_size_eval=4, num_clips=2, cpn=True, eval=True, resume="ckpt", distributed=False,
                              alpha=[0.2, 0.9], c=[0.9, 0.3, 0.9, 0.8],
                              iv2_scores={"t2v": t2v, "v2t": t2v.t().contiguous() + 0.1 * torch.randn(n, n, generator=gs)})
    loader = D.load_data(args, tokenizer=tok, split="test", root=str(root))
    model = BlimModel(cfg, state_dict=synth.init_weights(cfg, seed=2, std=0.05, rich=True), device=0, max_run_tokens=4096, max_prefix_tokens=4096)
    try:
        D.stage_corpus(model, loader.dataset)
        res = evalloop.val_one_epoch(model, loader, None, model.device, 0, None, tokenizer=tok, args=args)
        assert set(res) == {"internvideo2", "candidate_likelihood", "query_likelihood", "cpn_candidate_likelihood", "blim"}
        t2v_a, v2t_a = retrieval.evaluation(model, loader, model.device, tok, args)
        # the same scoring with the extractor's tensors handed over directly (no files, no dataset classes)
        ds = loader.dataset
        eng = model.engine
        eng.set_videos(torch.stack([direct[d["vid"]] for d in ds.data], 0))
        eng.build_video_vocab(np.arange(n), n_vocab=n)
        t2v_b, v2t_b = retrieval.evaluation(model, loader, model.device, tok, args)
        for x, y in ((t2v_a, t2v_b), (v2t_a, v2t_b)):
            for key in x:
                np.testing.assert_array_equal(x[key], y[key])
        assert np.isfinite(v2t_a["candidate_likelihood"][v2t_a["candidate_likelihood"] != -100.0]).all()
    finally:
        model.engine.close()


def test_chunk_bounds_cover_the_list():
    for n, k in ((10, 3), (7, 7), (5, 1), (1000, 8)):
        spans = [extract.chunk_bounds(n, k, i) for i in range(k)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(k - 1))
