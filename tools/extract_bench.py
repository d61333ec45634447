#!/usr/bin/env python
"""Feature-extraction bench (not the BASELINE.json headline, which bench.py measures).

A "step" = one batch of videos (default 4 videos x 16 frames @ 448 px = 16 clips of 3136 tokens) through the ViT-L/16
encoder (23 blocks) and the ToMe merge to [4, 64, 1024] features per video -- what the reference's extract.py:96-110 does
per video.  Prints ONE JSON line: videos/s with frames resident on the device (`value`), end to end from pinned host
frames to host fp16 features (`e2e`), the tcgen05 GEMM roofline, per-family device times, and the oracle on the host
cores beside it (`cpu_baseline`, one clip).

  python tools/extract_bench.py [--videos 4 --frames 16 --steps 5 --warmup 3 --image-size 448]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--videos", type=int, default=4)
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--image-size", type=int, default=448)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit("extract_bench.py: no CUDA device -- the extractor has no CPU fallback")
    import bench
    from blim_b200 import vision as V
    cfg = V.VisionConfig.umt_l(args.image_size)
    dev = torch.device("cuda", 0)
    w = V.init_weights(cfg, seed=0, device=dev)
    n_clips = args.videos * args.frames // cfg.frames_per_clip
    enc = V.VisionEncoder(cfg, state_dict=w, device=0, max_clips=n_clips)
    g = torch.Generator(device=dev).manual_seed(1)
    frames = torch.randn(args.videos * args.frames, 3, args.image_size, args.image_size, generator=g, device=dev).to(torch.float16)
    host = frames.cpu().pin_memory()
    out_host = torch.empty((n_clips, 64, cfg.hidden_size), dtype=torch.float16).pin_memory()

    def step_device():
        return enc.extract(frames, out_dtype=torch.float16)

    def step_e2e():
        out_host.copy_(enc.extract(host.to(dev, non_blocking=True), out_dtype=torch.float16), non_blocking=True)
        return out_host

    def timed(fn, steps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    for _ in range(args.warmup):
        step_device()
    enc.profile(True)
    enc.profile_read()
    l0, f0 = enc.kernel_launches(), enc.gemm_flops()
    sampler = bench.ClockSampler(0)
    sampler.start()
    ms = timed(step_device, args.steps)
    clocks = sampler.stop()
    prof = enc.profile_read()
    enc.profile(False)
    launches = (enc.kernel_launches() - l0) // args.steps
    gemm_flops = (enc.gemm_flops() - f0) / args.steps
    step_e2e()
    e2e_ms = timed(step_e2e, args.steps)

    C, F, L, T = cfg.hidden_size, cfg.mlp_hidden_size, cfg.num_layers, cfg.tokens_per_clip
    M = n_clips * T
    alg_gemm = 2.0 * M * (3 * cfg.patch_size ** 2 * C + L * (3 * C * C + C * C + 2 * C * F))
    alg_attn = 4.0 * L * n_clips * T * T * C
    peaks = bench.measured_peaks()
    gemm_s = prof["gemm"]["ms"] / 1000.0 / args.steps
    achieved = alg_gemm / gemm_s / 1e12
    line = {"metric": "videos/s, feature extraction (ViT-L/16 encoder + ToMe merge to [clips, 64, 1024])", "value": args.videos / (ms / args.steps / 1000.0),
            "unit": "videos/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{args.videos} videos x {args.frames} frames @ {args.image_size}px = {n_clips} clips x {T} tokens, 23 blocks, random-init UMT ViT-L",
                       "tokens_per_step": M, "l2": "activations 0.6 GB per layer >> 126 MB L2"},
            "clocks": clocks,
            "e2e": {"value": args.videos / (e2e_ms / args.steps / 1000.0), "unit": "videos/s", "ms_per_step": e2e_ms / args.steps,
                    "h2d_bytes_per_step": host.numel() * 2, "d2h_bytes_per_step": out_host.numel() * 2,
                    "api": "blim_b200.vision.VisionEncoder.extract, pinned host frames -> host fp16 features"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "kernel": "gemm_tcgen05_kernel (patch embed, qkv, proj, fc1, fc2)", "achieved": achieved,
                         "peak": peaks["bf16_sustained"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16_sustained"],
                         "peak_source": f"{peaks['source']} (sustained cuBLAS bf16; burst {peaks['bf16_burst']})",
                         "algorithmic_gemm_flops_per_step": alg_gemm, "executed_gemm_flops_per_step": gemm_flops,
                         "algorithmic_attention_flops_per_step": alg_attn,
                         "attention_tflops": alg_attn / (prof["attention"]["ms"] / 1000.0 / args.steps) / 1e12,
                         "whole_step_tflops": (alg_gemm + alg_attn) / (ms / args.steps / 1000.0) / 1e12,
                         "by_kernel": {k: {"ms_per_step": d["ms"] / args.steps, "share_of_step": d["ms"] / ms, "launches_per_step": d["launches"] // args.steps}
                                       for k, d in prof.items()}}}
    if not args.no_cpu_baseline:
        from oracle import vision_oracle as VO
        torch.set_num_threads(os.cpu_count())
        wc = {k: v.float().cpu() for k, v in w.items()}
        clip = host[:cfg.frames_per_clip].float()
        t0 = time.time()
        with torch.no_grad():
            VO.extract(wc, cfg, clip)
        dt = time.time() - t0
        line["cpu_baseline"] = {"value": 1.0 / (dt * args.frames / cfg.frames_per_clip), "unit": "videos/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": f"one clip ({cfg.frames_per_clip} frames) through the fp32 oracle in {dt:.1f} s, scaled to {args.frames} frames per video"}
    print(json.dumps(line), flush=True)
    enc.close()


if __name__ == "__main__":
    main()
