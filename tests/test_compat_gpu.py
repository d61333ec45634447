"""The literal reference call sequence (retrieval_utils.py:62-110) against BlimModel's model-object surface:
prepare_inputs_labels_for_multimodal -> model(inputs_embeds=, attention_mask=) -> criterion.  This is the slow
compatibility path (logits materialised by blim_forward_logits); it must agree with the reference goldens and with
the engine's fast path.  GPU only."""
import math
import os
import types

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from blim_b200.engine import TVG, TVG_PRIOR, VTG, VTG_PRIOR
from blim_b200.model import BlimModel
from blim_b200 import retrieval
from oracle import blim_oracle as O
from oracle.make_golden import CASES, build_case, pad_left

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def reference_style_rows(model, corpus, cfg, sims, ids, masks, labels, forward_type, cpn, topk, bs, direction, rows):
    """Per-row loop with the reference's call sequence (same video replicated / same text replicated, ragged tail)."""
    dev = model.device
    n = sims.shape[0]
    out = torch.full((n, n), -100.0)
    vocab = corpus.video_vocab.to(dev)
    for r in rows:
        idx = sims[r].topk(k=min(n, topk), dim=0).indices
        scores = []
        for j in range(0, len(idx), bs):
            sel = idx[j:j + bs]
            m = len(sel)
            if direction == "v2t":
                b_ids, b_msk, b_lab = ids[sel], masks[sel], labels[sel]
                vids = [corpus.video[r].to(dev)] * m
                vlab = corpus.tvg_video_labels[r].repeat(m, corpus.n_clips)
            else:
                b_ids, b_msk, b_lab = ids[r].repeat(m, 1), masks[r].repeat(m, 1), labels[r].repeat(m, 1)
                vids = [corpus.video[int(v)].to(dev) for v in sel]
                vlab = corpus.tvg_video_labels[sel][:, None].repeat(1, corpus.n_clips)
            (_, _, (msk, cpn_msk), _, embeds, lab) = model.module.prepare_inputs_labels_for_multimodal(
                b_ids.to(dev), None, b_msk.to(dev), None, b_lab.to(dev), vids, ["video"] * m, image_sizes=[(448, 448)] * m,
                video_feature=True, tvg=(forward_type == "tvg"), cpn=True)
            outp = model(inputs_embeds=embeds, attention_mask=cpn_msk if cpn else msk)
            if forward_type == "vtg":
                s = O.vtg_criterion(outp.logits, lab)
            else:
                pcol = (lab == cfg.image_token_id).nonzero()[:, 1]
                vi = pcol[:, None] + (torch.arange(corpus.n_clips, device=dev) - (corpus.n_clips + 1))[None]
                emb = torch.gather(outp.hidden_states, 1, vi[..., None].expand(-1, -1, outp.hidden_states.shape[-1]))
                emb = model.module.forward_visual(emb)
                tl = torch.bmm(emb.float().permute(1, 0, 2), vocab.float().permute(1, 2, 0)).transpose(0, 1) / math.sqrt(vocab.shape[-1])
                s = O.tvg_criterion(tl, vlab.to(dev))
            scores.append(s.float().cpu())
        out[r, idx] = torch.cat(scores)
    return out.numpy()


@pytest.mark.parametrize("name", sorted(CASES))
def test_reference_call_sequence_on_engine_model(name):
    spec = CASES[name]
    cfg, weights, corpus = build_case(spec)
    gold = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    model = BlimModel(cfg, state_dict=weights, device=0, max_run_tokens=4096, max_prefix_tokens=4096)
    try:
        model.set_tvg_prefix_length(corpus.tvg_prefix_length)
        rows = [0, 3]
        for direction, sims in (("v2t", corpus.v2t_iv2), ("t2v", corpus.t2v_iv2)):
            for ft, ids_l, lab_l in (("vtg", corpus.vtg_ids, corpus.vtg_labels), ("tvg", corpus.tvg_ids, corpus.tvg_labels)):
                ids, labels = pad_left(ids_l, corpus.pad_token_id), pad_left(lab_l, -100)
                masks = pad_left([torch.ones_like(x) for x in ids_l], 0)
                for cpn in (False, True):
                    got = reference_style_rows(model, corpus, cfg, sims, ids, masks, labels, ft, cpn, spec["topk"], spec["bs"], direction, rows)
                    ref = gold[f"{direction}_{ft}_{'cpn' if cpn else 'lik'}"]
                    for r in rows:
                        sel = ref[r] != -100.0
                        err = np.abs(got[r][sel] - ref[r][sel]).max()
                        assert err <= 1e-2, f"{name} {direction} {ft} cpn={cpn} row {r}: {err}"
        # the drop-in compute_*_scores_x (fast path) fills the same matrix entries
        args = types.SimpleNamespace(topk=spec["topk"], batch_size_eval=spec["bs"], num_clips=corpus.n_clips)
        ids, labels = pad_left(corpus.vtg_ids, 0), pad_left(corpus.vtg_labels, -100)
        masks = pad_left([torch.ones_like(x) for x in corpus.vtg_ids], 0)
        video = [v for v in corpus.video]
        m = torch.full((corpus.n, corpus.n), -100.0)
        m = retrieval.compute_v2t_scores_x(m, corpus.v2t_iv2, 0, ids, masks, labels, video, corpus.video_vocab, corpus.tvg_video_labels,
                                           model, model.device, args, forward_type="vtg", cpn=False).numpy()
        ref = gold["v2t_vtg_lik"]
        assert ((m == -100.0) == (ref == -100.0)).all()
        assert np.abs(m - ref).max() <= 1e-2
        ids, labels = pad_left(corpus.tvg_ids, 0), pad_left(corpus.tvg_labels, -100)
        masks = pad_left([torch.ones_like(x) for x in corpus.tvg_ids], 0)
        m = torch.full((corpus.n, corpus.n), -100.0)
        m = retrieval.compute_t2v_scores_x(m, corpus.t2v_iv2, 0, ids, masks, labels, video, corpus.video_vocab, corpus.tvg_video_labels,
                                           model, model.device, args, forward_type="tvg", cpn=True).numpy()
        ref = gold["t2v_tvg_cpn"]
        assert ((m == -100.0) == (ref == -100.0)).all()
        assert np.abs(m - ref).max() <= 1e-2
    finally:
        model.engine.close()


@pytest.mark.parametrize("name", sorted(CASES))
def test_unmodified_reference_loops_run_on_engine_model(name):
    """The reference's OWN retrieval_utils.compute_v2t_scores_x / compute_t2v_scores_x (imported unmodified from the shipped
    baseline/_ref, not restated) driving BlimModel: the model-object surface they touch -- .module.prepare_inputs_labels_for_
    multimodal, __call__(inputs_embeds=, attention_mask=) -> .logits / .hidden_states, .module.forward_visual -- is served by
    the engine, and the matrices equal the goldens the same functions produced with the reference model."""
    from oracle import ref_harness
    if not ref_harness.reference_available():
        pytest.skip("reference sources not on this box (baseline/_ref is written by build() in the build container)")
    ru, tu, mvf = ref_harness.import_reference()
    spec = CASES[name]
    cfg, weights, corpus = build_case(spec)
    ru.IMAGE_TOKEN_ID = cfg.image_token_id                       # small vocabulary: same remap the golden generator uses
    gold = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    model = BlimModel(cfg, state_dict=weights, device=0, max_run_tokens=4096, max_prefix_tokens=4096)
    try:
        model.set_tvg_prefix_length(corpus.tvg_prefix_length)
        dev = model.device
        args = types.SimpleNamespace(topk=spec["topk"], batch_size_eval=spec["bs"], num_clips=corpus.n_clips)
        video = [v for v in corpus.video]
        vocab = corpus.video_vocab.to(dev)
        rows = 3
        for direction, fn, sims in (("v2t", ru.compute_v2t_scores_x, corpus.v2t_iv2), ("t2v", ru.compute_t2v_scores_x, corpus.t2v_iv2)):
            for ft, ids_l, lab_l in (("vtg", corpus.vtg_ids, corpus.vtg_labels), ("tvg", corpus.tvg_ids, corpus.tvg_labels)):
                ids, labels = pad_left(ids_l, corpus.pad_token_id), pad_left(lab_l, -100)
                masks = pad_left([torch.ones_like(x) for x in ids_l], 0)
                for cpn in (False, True):
                    m = torch.full((corpus.n, corpus.n), -100.0).to(dev)
                    # like val_one_epoch (training_utils.py:142) the loops run under autocast: the criteria are computed in fp32
                    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                        m = fn(m, sims[:rows], 0, ids, masks, labels, video, vocab, corpus.tvg_video_labels, model, dev, args, forward_type=ft, cpn=cpn)
                    got = m.float().cpu().numpy()[:rows]
                    ref = gold[f"{direction}_{ft}_{'cpn' if cpn else 'lik'}"][:rows]
                    assert ((got == -100.0) == (ref == -100.0)).all()
                    # TVG: the loop's own torch.bmm (retrieval_utils.py:106) returns bf16 logits under autocast -- its rounding, not the engine's
                    tol = 1e-2 if ft == "vtg" else 2e-2
                    assert np.abs(got - ref).max() <= tol, f"{name} {direction} {ft} cpn={cpn}: {np.abs(got - ref).max()}"
    finally:
        model.engine.close()
