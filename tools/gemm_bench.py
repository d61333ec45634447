"""Micro-benchmark of the tcgen05 GEMM core through the C ABI (CUDA events, L2 flushed between timed launches)."""
import json
import math
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blim_b200.engine import Engine, ModelConfig  # noqa: E402


def main():
    eng = Engine(ModelConfig.tiny(), max_run_tokens=16384, max_prefix_tokens=256)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    shapes = [(8192, 4608, 3584, 0), (8192, 3584, 3584, 0), (8192, 37888, 3584, 5), (8192, 3584, 18944, 0), (16384, 37888, 3584, 5),
              (4096, 152064, 3584, 6), (8192, 8192, 8192, 0), (27072, 3584, 18944, 4), (27072, 3584, 18944, 7), (27072, 3584, 3584, 4),
              (27072, 3584, 3584, 7), (27072, 3584, 18944, 4), (27072, 3584, 18944, 7), (27072, 3584, 3584, 0), (27072, 3584, 3584, 3),
              (27072, 4608, 3584, 0), (27072, 3584, 3584, 4), (27136, 3584, 3584, 0), (18944, 3584, 3584, 0)]
    out = []
    for M, N, K, epi in shapes:
        A = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
        W = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
        tgt = torch.randint(0, N, (M,), device="cuda", dtype=torch.int32)
        for cg in ((2,) if epi in (4, 7) else (1, 2)):
            C = torch.zeros(M, N, device="cuda") if epi in (4, 7) else None
            ts = []
            for it in range(6):
                flush.zero_()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                C = eng.debug_gemm(epi, A, W, target=tgt, cta_group=cg, C=C)
                e.record()
                torch.cuda.synchronize()
                ts.append(s.elapsed_time(e))
            t = min(ts[2:])
            rec = dict(M=M, N=N, K=K, epilogue=epi, cta_group=cg, ms=round(t, 4), tflops=round(2.0 * M * N * K / t / 1e9, 1))
            print(json.dumps(rec), flush=True)
            out.append(rec)
        # torch (cuBLAS) for comparison, plain GEMM only
        ts = []
        for it in range(6):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            torch.matmul(A, W.t())
            e.record()
            torch.cuda.synchronize()
            ts.append(s.elapsed_time(e))
        t = min(ts[2:])
        rec = dict(M=M, N=N, K=K, impl="torch.matmul", ms=round(t, 4), tflops=round(2.0 * M * N * K / t / 1e9, 1))
        print(json.dumps(rec), flush=True)
        out.append(rec)
        del A, W
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/gemm_bench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
