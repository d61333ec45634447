"""Pins the tcgen05 shared-memory descriptor semantics the attention kernel relies on: operands staged by ordinary
threads with a manual 128-byte swizzle, and an MN-major B operand (V of attention: [key][head_dim]).  GPU only."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from blim_b200.engine import Engine, ModelConfig


@pytest.fixture(scope="module")
def eng():
    e = Engine(ModelConfig.tiny(), max_run_tokens=256, max_prefix_tokens=256)
    yield e
    e.close()


def _rand(shape, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda") * 0.5).bfloat16()


@pytest.mark.parametrize("K,N", [(64, 64), (128, 64), (64, 128), (128, 128)])
def test_k_major_manual_swizzle(eng, K, N):
    A, B = _rand((128, K), 1), _rand((N, K), 2)
    ref = A.float() @ B.float().t()
    got = eng.debug_umma(A, B, b_mn_major=False)
    torch.cuda.synchronize()
    assert (got - ref).abs().max() < 1e-3 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("K,N", [(64, 64), (128, 64), (64, 128), (128, 128), (16, 128), (48, 64)])
def test_mn_major_b_descriptor(eng, K, N):
    """B = [K][N] (row = key): sub-tiles of [K x 64] per 64 columns of N; LBO = byte distance between those sub-tiles,
    SBO = 1024 (8 K-rows), 2048 bytes per 16-row MMA step.  K < 64: rows beyond K are never touched by the K/16 MMAs."""
    Kp = 64 if K < 64 else K
    A, B = _rand((128, Kp), 3), _rand((Kp, N), 4)
    if K < Kp:
        A[:, K:] = 0
    ref = A.float() @ B.float()
    report = {}
    for name, (lbo, sbo, kstep) in {"lbo=tile,sbo=1024,k=2048": (Kp * 128, 1024, 2048), "lbo=1024,sbo=tile,k=2048": (1024, Kp * 128, 2048),
                                    "lbo=tile,sbo=1024,k=32": (Kp * 128, 1024, 32), "lbo=0,sbo=1024,k=2048": (0, 1024, 2048)}.items():
        got = eng.debug_umma(A, B, b_mn_major=True, lbo=lbo, sbo=sbo, kstep=kstep)
        torch.cuda.synchronize()
        report[name] = (got - ref).abs().max().item()
    print(f"K={K} N={N}:", {k: round(v, 4) for k, v in report.items()})
    assert report["lbo=tile,sbo=1024,k=2048"] < 1e-3 * max(1.0, ref.abs().max().item()), report


def test_fp16_operands(eng):
    """kind::f16 with both operands in fp16 (the scoring path's operand format): values use the full 11-bit significand,
    so reading the fp16 bits as bf16 cannot pass.  (Mixed fp16 x bf16 operands are NOT supported by the hardware: the
    MMA raises an illegal-instruction error, measured on B200 -- which is why weights are converted to fp16 at load.)"""
    g = torch.Generator(device="cuda").manual_seed(11)
    A = (torch.randn(128, 128, generator=g, device="cuda") * 0.5).half()
    B = (torch.randn(64, 128, generator=g, device="cuda") * 0.5).half()
    ref = A.float() @ B.float().t()
    got = eng.debug_umma(A, B, b_mn_major=False)
    torch.cuda.synchronize()
    assert (got - ref).abs().max() < 1e-4 * max(1.0, ref.abs().max().item()), (got - ref).abs().max().item()
