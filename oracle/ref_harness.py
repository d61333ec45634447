"""TEST INFRASTRUCTURE -- imports the UNMODIFIED reference (mlvlab/BLiM) to pin the oracle and to serve as comparator.

Where the reference comes from, in this order: $BLIM_REF, /root/reference (build container only), baseline/_ref/ (a
git-ignored verbatim copy of the reference's .py files that __graft_entry__.build() makes in the build container; it
travels to the GPU box with the repo snapshot, never enters the history).  Nothing in the product imports this module.
Users: oracle/make_golden.py (fixtures under tests/golden/), the tests that compare oracle/blim_oracle.py and the CUDA
engine with the real reference, tools/parity_7b.py and bench.py's reference arm / rank-parity check.

Shims: stub modules for packages the vision tower / video IO import but the scoring path never
calls (timm.layers, av, imageio, decord), and three transformers-5 workarounds on the config object.
"""
import os
import sys
import types

import torch

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIPPED = os.path.join(_REPO, "baseline", "_ref")


def _find_root():
    for cand in (os.environ.get("BLIM_REF"), "/root/reference", SHIPPED):
        if cand and os.path.isfile(os.path.join(cand, "retrieval_utils.py")) and os.path.isdir(os.path.join(cand, "videochat_flash")):
            return cand
    return os.environ.get("BLIM_REF", "/root/reference")


REF_ROOT = _find_root()


def reference_available():
    return os.path.isfile(os.path.join(REF_ROOT, "retrieval_utils.py")) and os.path.isdir(os.path.join(REF_ROOT, "videochat_flash"))


def ship_reference(src="/root/reference", dst=SHIPPED):
    """Verbatim copy of the reference's Python sources into the git-ignored baseline/_ref/ so that they reach the GPU box
    (which has no /root/reference).  Called by __graft_entry__.build() in the build container; returns the file count."""
    import shutil
    if not os.path.isdir(os.path.join(src, "videochat_flash")):
        return 0
    n = 0
    for dirpath, dirnames, filenames in os.walk(src):
        dirnames[:] = [d for d in dirnames if d not in (".git", "asset", "__pycache__")]
        for f in filenames:
            if not f.endswith((".py", ".md", ".sh")) and f != "LICENSE":
                continue
            rel = os.path.relpath(os.path.join(dirpath, f), src)
            out = os.path.join(dst, rel)
            os.makedirs(os.path.dirname(out), exist_ok=True)
            shutil.copyfile(os.path.join(dirpath, f), out)
            n += 1
    return n


def _install_stubs():
    import transformers  # noqa: F401  (must be imported before the stubs shadow anything it probes)

    def stub(name, **attrs):
        if name in sys.modules:
            return sys.modules[name]
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    try:
        import timm.layers  # noqa: F401
    except Exception:
        timm = stub("timm")
        layers = stub("timm.layers", drop_path=lambda x, p=0.0, training=False: x, to_2tuple=lambda x: (x, x),
                      trunc_normal_=lambda t, std=0.02: torch.nn.init.trunc_normal_(t, std=std))
        timm.layers = layers
    for name in ("av", "imageio"):
        try:
            __import__(name)
        except Exception:
            stub(name)
    try:
        import decord  # noqa: F401
    except Exception:
        stub("decord", VideoReader=object, cpu=lambda *a, **k: None)


def import_reference():
    """Returns (retrieval_utils, training_utils, modeling_videochat_flash) modules of the reference."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    _install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import retrieval_utils  # noqa: E402
    import training_utils  # noqa: E402
    from videochat_flash import modeling_videochat_flash as mvf  # noqa: E402
    return retrieval_utils, training_utils, mvf


import contextlib


@contextlib.contextmanager
def _skip_random_init():
    """Parameters are overwritten by load_state_dict right after construction: skip the constructors' random fills (minutes
    for 7.6 G parameters on host cores)."""
    names = ("kaiming_uniform_", "uniform_", "normal_", "trunc_normal_", "xavier_uniform_", "zeros_", "ones_", "constant_")
    saved = {n: getattr(torch.nn.init, n) for n in names}
    try:
        for n in names:
            setattr(torch.nn.init, n, lambda t, *a, **k: t)
        yield
    finally:
        for n, f in saved.items():
            setattr(torch.nn.init, n, f)


class _Wrap(torch.nn.Module):
    """evaluation()/compute_*_scores_x expect a DDP-like object with `.module` (retrieval_utils.py:66)."""

    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, *a, **k):
        return self.module(*a, **k)


def build_reference_model(cfg, state_dict, dtype=torch.float32, image_token_id=None, device=None):
    """Instantiate the reference VideoChatFlashQwenForCausalLM for `cfg` (blim_b200.engine.ModelConfig) and load
    `state_dict` (reference parameter names).  Returns the wrapped model (with .module) and the reference modules.
    With `device` the module is constructed directly there in `dtype` (7B: no 30 GB host-side fp32 init)."""
    ru, tu, mvf = import_reference()
    c = mvf.VideoChatFlashQwenConfig(
        hidden_size=cfg.hidden_size, num_hidden_layers=cfg.num_layers, num_attention_heads=cfg.num_heads,
        num_key_value_heads=cfg.num_kv_heads, intermediate_size=cfg.intermediate_size, vocab_size=cfg.vocab_size,
        rms_norm_eps=cfg.rms_norm_eps, max_position_embeddings=cfg.max_positions, hidden_act="silu",
        use_sliding_window=False, attention_dropout=0.0, tie_word_embeddings=False)
    c.mm_vision_tower = "umt-hd-stub"
    c.delay_load = True
    c.mm_projector_type = "tome16_mlp_hd64"
    c.vision_encode_type = "video_image"
    c.mm_patch_merge_type = "spatial_nopad"
    c.mm_newline_position = "nothing"
    c.mm_local_num_frames = 4
    c.mm_vision_select_layer = -2
    c.mm_hidden_size = cfg.mm_hidden_size
    c.mm_pos_num_frames = 8
    c.tokenizer_padding_side = "right"
    c.use_cache = False
    c._attn_implementation = "sdpa"
    object.__setattr__(c, "rope_theta", cfg.rope_theta)
    if device is None:
        model = mvf.VideoChatFlashQwenForCausalLM(c)
    else:
        old = torch.get_default_dtype()
        torch.set_default_dtype(dtype)
        saved_init = mvf.VideoChatFlashQwenForCausalLM._init_weights
        mvf.VideoChatFlashQwenForCausalLM._init_weights = lambda self, module: None
        try:
            with torch.device(device), _skip_random_init():
                model = mvf.VideoChatFlashQwenForCausalLM(c)
        finally:
            torch.set_default_dtype(old)
            mvf.VideoChatFlashQwenForCausalLM._init_weights = saved_init
    missing, unexpected = model.load_state_dict(state_dict, strict=False)
    need = [k for k in missing if not k.startswith("model.vision_tower") and "rotary_emb" not in k]
    assert not need, f"reference parameters not provided: {need[:5]}"
    if device is not None:   # nothing may be left at its (skipped) initial value
        loaded = set(state_dict)
        left = [k for k, _ in model.named_parameters() if k not in loaded and not k.startswith("model.vision_tower")]
        assert not left, f"reference parameters left uninitialised: {left[:5]}"
    model = model.to(dtype).eval()
    if image_token_id is not None:
        ru.IMAGE_TOKEN_ID = image_token_id  # small vocabularies: remap the module attribute (retrieval_utils.py:14,99)
    return _Wrap(model), (ru, tu, mvf)
