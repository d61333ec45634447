"""tcgen05 GEMM core vs plain PyTorch fp32 (same bf16 inputs), through the C ABI debug entry.  GPU only."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from blim_b200.engine import Engine, ModelConfig

SHAPES = [(128, 256, 64), (128, 256, 256), (200, 512, 128), (77, 256, 1024), (1000, 768, 3584), (2500, 4608, 3584), (333, 1000, 1024),
          (4096, 512, 18944)]


@pytest.fixture(scope="module")
def eng():
    e = Engine(ModelConfig.tiny(), max_run_tokens=8192, max_prefix_tokens=4096)
    yield e
    e.close()


def _inputs(M, N, K, seed):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    A = (torch.randn(M, K, generator=g, device="cuda") * 0.5).bfloat16()
    W = (torch.randn(N, K, generator=g, device="cuda") / math.sqrt(K)).bfloat16()
    return A, W


def _report(name, got, ref, tol):
    err = (got.float() - ref).abs()
    bad = err > tol
    if bad.any():
        idx = bad.nonzero()[0].tolist()
        msg = (f"{name}: {int(bad.sum())}/{bad.numel()} elements off (max err {err.max().item():.4g}, tol {tol:.3g}), first at {idx}; "
               f"rows with errors: {bad.any(1).nonzero().flatten()[:16].tolist()} cols with errors: {bad.any(0).nonzero().flatten()[:16].tolist()}\n"
               f"got[:2,:8]={got[:2, :8].float().tolist()}\nref[:2,:8]={ref[:2, :8].tolist()}")
        pytest.fail(msg)


@pytest.mark.parametrize("cta_group", [1, 2])
@pytest.mark.parametrize("shape", SHAPES)
def test_gemm_store(eng, shape, cta_group):
    M, N, K = shape
    A, W = _inputs(M, N, K, 1)
    ref = A.float() @ W.float().t()
    tol = 2e-2 * ref.abs().max().item() + 1e-3
    got = eng.debug_gemm(0, A, W, cta_group=cta_group)
    torch.cuda.synchronize()
    _report(f"bf16 store {shape} cg{cta_group}", got, ref, tol)
    got32 = eng.debug_gemm(3, A, W, cta_group=cta_group)
    torch.cuda.synchronize()
    _report(f"fp32 store {shape} cg{cta_group}", got32, ref, 2e-3 * ref.abs().max().item() + 1e-4)


@pytest.mark.parametrize("cta_group", [1, 2])
def test_gemm_bias_gelu_resid(eng, cta_group):
    M, N, K = 700, 1024, 512
    A, W = _inputs(M, N, K, 2)
    bias = torch.randn(N, device="cuda")
    ref = A.float() @ W.float().t() + bias
    got = eng.debug_gemm(1, A, W, bias=bias, cta_group=cta_group)
    _report("bias", got, ref, 2e-2 * ref.abs().max().item())
    got = eng.debug_gemm(2, A, W, bias=bias, cta_group=cta_group)
    _report("bias+gelu", got, torch.nn.functional.gelu(ref), 2e-2 * ref.abs().max().item())
    C = torch.randn(M, N, device="cuda")
    ref2 = C + A.float() @ W.float().t()
    eng.debug_gemm(4, A, W, cta_group=cta_group, C=C)
    torch.cuda.synchronize()
    _report("residual", C, ref2, 2e-3 * ref2.abs().max().item())


@pytest.mark.parametrize("cta_group", [1, 2])
def test_gemm_swiglu(eng, cta_group):
    M, I, K = 500, 1024, 256
    g = torch.Generator(device="cuda")
    g.manual_seed(3)
    A = torch.randn(M, K, generator=g, device="cuda").bfloat16()
    Wg = (torch.randn(I, K, generator=g, device="cuda") / math.sqrt(K)).bfloat16()
    Wu = (torch.randn(I, K, generator=g, device="cuda") / math.sqrt(K)).bfloat16()
    # 128-row interleave: [gate 0..127 | up 0..127 | gate 128..255 | ...]
    W = torch.stack([Wg.view(I // 128, 128, K), Wu.view(I // 128, 128, K)], dim=1).reshape(2 * I, K).contiguous()
    ref = torch.nn.functional.silu(A.float() @ Wg.float().t()) * (A.float() @ Wu.float().t())
    got = eng.debug_gemm(5, A, W, cta_group=cta_group)
    torch.cuda.synchronize()
    _report("swiglu", got, ref, 2e-2 * ref.abs().max().item())


@pytest.mark.parametrize("cta_group", [1, 2])
@pytest.mark.parametrize("shape", [(300, 4096, 256), (1500, 1000, 1024), (640, 152064, 256)])
def test_gemm_lse(eng, shape, cta_group):
    M, N, K = shape
    A, W = _inputs(M, N, K, 4)
    A = A * 4
    scale = 0.5
    tgt = torch.randint(0, N, (M,), device="cuda")
    logits = (A.float() @ W.float().t()) * scale
    ref = torch.log_softmax(logits, -1).gather(1, tgt[:, None]).squeeze(1)
    got = eng.debug_gemm(6, A, W, target=tgt, scale=scale, cta_group=cta_group)
    torch.cuda.synchronize()
    err = (got - ref).abs().max().item()
    assert err < 2e-3, f"lse {shape} cg{cta_group}: max err {err}, got {got[:4].tolist()} ref {ref[:4].tolist()}"
