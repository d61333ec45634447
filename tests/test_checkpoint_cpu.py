"""LoRA merge loader: W' = W + alpha/r * B @ A under the PEFT key names the reference produces."""
import pytest
import torch

from blim_b200.checkpoint import _plain_name, merge_lora


def test_plain_names():
    assert _plain_name("base_model.model.model.layers.3.self_attn.q_proj.base_layer.weight") == "model.layers.3.self_attn.q_proj.weight"
    assert _plain_name("base_model.model.model.mm_projector.mlp.base_model.model.0.base_layer.bias") == "model.mm_projector.mlp.0.bias"
    assert _plain_name("base_model.model.visual_head.weight") == "visual_head.weight"
    assert _plain_name("model.norm.weight") == "model.norm.weight"


def test_merge_matches_adapter_forward():
    g = torch.Generator().manual_seed(0)
    r, alpha = 8, 32
    base = {"model.layers.0.self_attn.q_proj.weight": torch.randn(64, 48, generator=g).bfloat16(),
            "model.mm_projector.tvg_mlp.2.weight": torch.randn(48, 48, generator=g).bfloat16(),
            "lm_head.weight": torch.randn(100, 48, generator=g).bfloat16(),
            "visual_head.weight": torch.zeros(16, 48).bfloat16(),
            "model.norm.weight": torch.ones(48).bfloat16()}
    ckpt = {}
    adapters = {"base_model.model.model.layers.0.self_attn.q_proj": (64, 48),
                "base_model.model.model.mm_projector.tvg_mlp.base_model.model.2": (48, 48),
                "base_model.model.lm_head": (100, 48)}
    for stem, (o, i) in adapters.items():
        ckpt[stem + ".lora_A.default.weight"] = torch.randn(r, i, generator=g) * 0.1
        ckpt[stem + ".lora_B.default.weight"] = torch.randn(o, r, generator=g) * 0.1
    ckpt["base_model.model.visual_head.weight"] = torch.randn(16, 48, generator=g)
    merged, names = merge_lora(base, ckpt, r, alpha, dtype=torch.float32)
    assert sorted(names) == sorted(["model.layers.0.self_attn.q_proj.weight", "model.mm_projector.tvg_mlp.2.weight", "lm_head.weight"])
    x = torch.randn(5, 48, generator=g)
    for stem, _ in adapters.items():
        from blim_b200.checkpoint import _plain_name as pn
        W = base[pn(stem) + ".weight"].float()
        A, B = ckpt[stem + ".lora_A.default.weight"], ckpt[stem + ".lora_B.default.weight"]
        want = x @ W.t() + (alpha / r) * (x @ A.t()) @ B.t()          # PEFT's adapter forward
        got = x @ merged[pn(stem) + ".weight"].t()
        assert torch.allclose(got, want, atol=1e-4)
    assert torch.equal(merged["visual_head.weight"], ckpt["base_model.model.visual_head.weight"])
    assert torch.equal(merged["model.norm.weight"], base["model.norm.weight"])
    with pytest.raises(ValueError):
        merge_lora(base, ckpt, 4, alpha)


@pytest.mark.parametrize("with_random_tvg_mlp", [False, True])
def test_reference_checkpoint_layout(tmp_path, with_random_tvg_mlp):
    """The reference's real key layout (main.py:99-111, util/misc.py:276-297): tvg_mlp adapters sit on a COPY OF THE BASE
    mlp, whether the base dict has no tvg_mlp at all (HF checkpoint) or an unrelated random one (constructor init)."""
    from blim_b200.checkpoint import load_finetuned
    from blim_b200.engine import ModelConfig
    from tests import ckpt_fixture as F
    cfg = ModelConfig.tiny()
    base = F.base_state_dict(cfg, with_random_tvg_mlp=with_random_tvg_mlp)
    ckpt, deltas = F.make_reference_checkpoint(cfg)
    path = tmp_path / "checkpoint_best.pth"
    torch.save(ckpt, path)                                  # holds an argparse.Namespace + optimizer state like the reference's

    class Sink:                                             # stands in for BlimModel (no GPU here)
        def load_state_dict(self, sd):
            self.sd = sd
            return []
    sink = Sink()
    merged_names, _ = load_finetuned(sink, base, str(path), 8, 32)
    want = F.expected_merged(cfg, base, ckpt, deltas)
    assert len(merged_names) == len(F.adapted_linears(cfg))
    for k, w in want.items():
        got = sink.sd[k].float()
        tol = 0.0 if k == "visual_head.weight" else 2.0 ** -8 * float(w.abs().max())       # one bf16 rounding of the merged weight
        assert (got - w).abs().max() <= tol, k
    # the tvg_mlp base really is the mlp base, not the base dict's own tvg_mlp
    k = "model.mm_projector.tvg_mlp.0.bias"
    assert torch.equal(sink.sd[k].float(), base["model.mm_projector.mlp.0.bias"].float())
