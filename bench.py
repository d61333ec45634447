#!/usr/bin/env python
"""bench.py -- BLiM bidirectional likelihood scoring on B200 (BASELINE.json metric).

A "step" is ONE pass of the whole hot path over the workload: for every text query and every video the top-k
InternVideo2 candidates are scored in both directions (VTG, TVG) with their CPN priors (six score matrices), combined by
the CPN + ensemble arithmetic and reranked into R@1/5/10.  Default workload = BASELINE.json configs[1]:
VideoChat-Flash-Qwen2-7B (random init), MSRVTT-1k shape (1000 queries x top-16), bf16, synthetic inputs.

  python bench.py [--gpus N --steps K --warmup W]            own arm (CUDA engine through the C ABI)
  python bench.py --impl reference [...]                     reference arm: the UNMODIFIED reference (baseline/_ref, shipped by
                                                             build(); else its CPU restatement in oracle/) on the host
                                                             cores, each step a bounded sample of the same workload; the
                                                             same reference timed on one B200 through PyTorch is reported
                                                             beside it as `reference_gpu`
N > 1: launched by torchrun, one rank per GPU; pairs are sharded by prefix owner (strong scaling of the fixed job),
ONE NCCL all-gather of compact scores per step, issued by the engine on its compute stream.
"""
import argparse
import contextlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "candidate pairs scored/sec (both directions+CPN)"
UNIT = "pairs/s"
DTYPE = "fp16"
DTYPE_NOTE = ("tcgen05 kind::f16 with fp16 operands (the bf16 checkpoint values are exactly representable; activations fp16), fp32 accumulation, "
              "fp32 residual stream -- the same tensor-core rate as bf16 operands with an 11-bit significand, the reference's own dtype (main.py:97); "
              "a -DBLIM_ACT_BF16 build runs all-bf16 operands at the same speed (csrc/act_type.cuh)")

WORKLOADS = {
    # name: (model config factory, dataset shape, n (None = dataset size), topk, alpha, c, n_clips)
    "c2": dict(model="qwen2_7b", dataset="msrvtt", n=None, topk=16, alpha=(0.0, 0.8), c=(1.0, 0.6, 0.8, 0.4), n_clips=None,
               desc="C2: VideoChat-Flash-Qwen2-7B random-init, MSRVTT-1k shape, 1000 queries x top-16, both directions + CPN (6 matrices)"),
    "c3": dict(model="qwen2_7b", dataset="didemo", n=None, topk=16, alpha=(0.0, 0.9), c=(0.9, 0.2, 0.9, 0.9), n_clips=None,
               desc="C3: 7B, DiDeMo shape (1004 queries, long captions), top-16, CPN + ensemble"),
    "c4": dict(model="qwen2_7b", dataset="activitynet", n=None, topk=16, alpha=(0.2, 0.9), c=(1.0, 0.4, 0.9, 0.8), n_clips=16,
               desc="C4: 7B, ActivityNet val_1 shape (4917 videos x 16 clips = 1024 visual tokens, long captions), top-16 (sized for 8 GPUs; use --n to subsample)"),
    "c5": dict(model="qwen2_7b", dataset="lsmdc", n=None, topk=64, alpha=(0.2, 1.0), c=(1.0, 0.6, 0.9, 0.6), n_clips=None,
               desc="C5: 7B, LSMDC test shape (1000 queries), top-64 (use --topk 16/32/64 for the sweep)"),
    "tiny": dict(model="tiny", dataset="msrvtt", n=64, topk=8, alpha=(0.0, 0.8), c=(1.0, 0.6, 0.8, 0.4), n_clips=None,
                 desc="tiny: 2-layer hidden-256 model, 64 queries x top-8 (debug)"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2",
                    help="one of %s, 'extract' (feature-extractor bench, tools/extract_bench.py), or a comma list with optional top-k "
                         "overrides sharing one engine / one weight load, e.g. c3,c5@16,c5@32,c5@64,c4 (one JSON line each)" % sorted(WORKLOADS))
    ap.add_argument("--n", type=int, default=0, help="override the number of queries/videos (debug)")
    ap.add_argument("--topk", type=int, default=0, help="override the workload's top-k")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the rank-parity check against the reference run on this GPU")
    ap.add_argument("--cta-group", type=int, default=0)
    ap.add_argument("--run-tokens", type=int, default=0, help="engine workspace: tokens per decoder run / prefix-cache rows (0 = engine default)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ helpers
def shared_config(wl, n, topk, world):
    """The `config` object of BOTH arms (the driver compares them): what is being measured, nothing arm-specific."""
    return {"workload": wl["desc"], "n_queries": int(n), "topk": int(topk), "pairs_per_step": int(2 * n * topk), "matrices": 6,
            "alpha": list(wl["alpha"]), "c": list(wl["c"]),
            "l2": "no flush needed: every step streams 15 GB of weights per decoder run (>> 126 MB L2)",
            "parallelism": f"pairs sharded by prefix owner over {world} GPU(s), weights replicated"}



class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 200 ms during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        # samples under load = upper half (the sampler also sees the idle edges)
        sm_sorted = sorted(sm)
        med = sm_sorted[len(sm_sorted) // 2] if sm_sorted else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"bf16_sustained": p.get("bf16_tflops_sustained"), "bf16_burst": p.get("bf16_tflops"), "hbm": p.get("hbm_gbs"), "source": "measured"}
    return {"bf16_sustained": 1400.0, "bf16_burst": 1590.0, "hbm": 6650.0, "source": "fallback"}


def ncu_traffic():
    """DRAM bytes per launch of the dominant kernel (gate|up + SwiGLU tcgen05 GEMM) from the committed `ncu --set full`
    capture (profiles/r01_traffic_*.json, written by tools/ncu_summary.py + the curation step in profiles/README.md);
    None when no capture is committed.  Returns (bytes, detail dict)."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic_*.json")))
    if not files:
        return None, None
    rows = json.load(open(files[-1]))["launches"]
    top = max((r for r in rows if r["kind"] == "gate_up_swiglu"), key=lambda r: r["ms"], default=None)
    if top is None:
        return None, None
    return top["traffic_bytes"], {"file": os.path.relpath(files[-1], ROOT), "kernel": top["kernel"], "M": top["M"],
                                  "algorithmic_bytes": top["algorithmic_bytes"], "traffic_over_algorithmic": top["traffic_over_algorithmic"],
                                  "per_kind_traffic_over_algorithmic": {r["kind"]: r["traffic_over_algorithmic"] for r in rows}}


def algorithmic_flops(cfg, corpus, plan, cpn=True, full=True):
    """FLOPs the algorithm needs -- non-padding tokens, every shared prefix once (including the chat-template
    header shared by all video prefixes / all TVG text prefixes), logits only at scored positions, and in the last layer
    of a prefix only the QKV projection (KV cache) for all tokens plus the rest of the layer for the one token whose state
    is read.  Returns (gemm_flops, attention_flops)."""
    H, I, V, MM, L = cfg.hidden_size, cfg.intermediate_size, cfg.vocab_size, cfg.mm_hidden_size, cfg.num_layers
    nkv = cfg.num_kv_heads * cfg.head_dim
    p_qkv = 2.0 * H * (H + 2 * nkv)                      # per token, one layer
    p_layer = p_qkv + 2.0 * (H * H + 3 * H * I)          # per token, one layer
    p_dec = L * p_layer
    p_lm = 2.0 * H * V
    prefix = lambda n_tok, n_last: n_tok * ((L - 1) * p_layer + p_qkv) + n_last * (p_layer - p_qkv)
    att = lambda q, c: 4.0 * L * cfg.num_heads * cfg.head_dim * q * c
    n_vis = corpus.n_clips * cfg.tokens_per_clip
    nc = corpus.n_clips
    uv, ut = plan.union_np
    cap = np.array([int((lab != -100).sum()) for lab in corpus.vtg_labels])             # scored tokens per text (caption + 2)
    pre = np.array([int((lab == -100).sum()) - 1 for lab in corpus.vtg_labels])          # prompt tokens without the image sentinel
    t0 = np.array([int((ids == -200).nonzero()[0]) for ids in corpus.tvg_ids])           # TVG text length
    last_tok = np.array([int(ids[int((ids == -200).nonzero()[0]) - 1]) for ids in corpus.tvg_ids])
    g = a = 0.0
    vids = np.unique(uv)
    s_v = pre[0] + n_vis
    r_v = int((corpus.vtg_ids[0] == -200).nonzero()[0])                                   # shared header before the video (root)
    g += len(vids) * (prefix(s_v - r_v, 1) + n_vis * 2.0 * (MM * H + H * H)) + prefix(r_v, 0)   # video prefixes (header once) + projector
    a += len(vids) * att(s_v, (s_v + 1) / 2)
    suf = cap[ut] - 1                                                                     # decoder tokens per pair
    g += float(suf.sum()) * (p_dec + p_lm) + len(vids) * p_lm
    a += float(sum(att(q, s_v + (q + 1) / 2) for q in suf))
    if cpn:
        texts = np.unique(plan.v2t_np[1])
        q = cap[texts] - 1
        g += prefix(pre[0], 1) + float(q.sum()) * (p_dec + p_lm) + p_lm
        a += float(sum(att(x, pre[0] + (x + 1) / 2) for x in q))
    if full:
        texts = np.unique(ut)
        head = nc * (2.0 * H * MM + 2.0 * MM * corpus.n)
        r_t = corpus.tvg_prefix_length                                                    # shared header + instruction (root)
        g += float(sum(prefix(x - r_t, 1) for x in t0[texts])) + prefix(r_t, 0) + len(uv) * ((nc - 1) * p_dec + head)
        g += len(vids) * n_vis * 2.0 * (MM * H + H * H)                                   # tvg_mlp projector
        a += float(sum(att(x, (x + 1) / 2) for x in t0[texts])) + float(sum(att(nc - 1, t0[t] + nc / 2) for t in ut))
        if cpn:
            pv, pt = plan.t2v_np
            uniq = len(set(zip(pv.tolist(), t0[pt].tolist(), last_tok[pt].tolist())))
            uniq_lt = len(set(zip(t0[pt].tolist(), last_tok[pt].tolist())))
            g += prefix(corpus.tvg_prefix_length, 0) + (uniq * (nc - 1) + uniq_lt) * p_dec + uniq * head
            a += uniq * att(nc - 1, corpus.tvg_prefix_length + nc / 2) + uniq_lt * att(1, corpus.tvg_prefix_length)
    return g, a


class Loader:
    """Minimal data_loader / dataset pair carrying the reference's collate_fn output (base_dataset.py:152-163)."""

    def __init__(self, corpus, batch=64, pin=False):
        self.c = corpus
        self.batch = batch
        self.dataset = self
        self.video_vocab = corpus.video_vocab
        self.tvg_prefix_length = corpus.tvg_prefix_length
        self.video = corpus.video.cpu()
        if pin:
            self.video = self.video.pin_memory()
            self.video_vocab = self.video_vocab.cpu().pin_memory()

    def __len__(self):
        return self.c.n

    def __iter__(self):
        c = self.c
        for i in range(0, c.n, self.batch):
            j = min(c.n, i + self.batch)
            yield {"video": [self.video[x] for x in range(i, j)],
                   "vtg_ids": c.vtg_ids[i:j], "vtg_labels": c.vtg_labels[i:j], "vtg_masks": [torch.ones_like(x) for x in c.vtg_ids[i:j]],
                   "tvg_ids": c.tvg_ids[i:j], "tvg_labels": c.tvg_labels[i:j], "tvg_masks": [torch.ones_like(x) for x in c.tvg_ids[i:j]],
                   "tvg_video_labels": c.tvg_video_labels[i:j]}


# ------------------------------------------------------------------------------------------------ CPU (reference) arm
CPU_SAMPLE_PAIRS = 4   # candidate pairs per score-matrix kind in one CPU sample (one batched forward of 4 per kind)


def cpu_sample_reference(runner):
    """One bounded sample of the workload through the UNMODIFIED reference on the host cores (oracle/ref_gpu.py): for each
    of the six score matrices one row with its top-4 candidates as ONE batched forward + criterion
    (retrieval_utils.compute_*_scores_x, MSRVTT-1k-sized corpus so the TVG vocabulary has 1000 videos).
    Returns (pairs/s, description)."""
    _, secs, pairs = runner.all_matrices(0, 1, CPU_SAMPLE_PAIRS, batch_size=CPU_SAMPLE_PAIRS)
    desc = (f"unmodified reference (retrieval_utils.compute_v2t/t2v_scores_x, eager PyTorch, bf16 weights on the host): one row x top-{CPU_SAMPLE_PAIRS} "
            f"of each of the six score matrices = 6 batched forwards of {CPU_SAMPLE_PAIRS} pairs + criteria, {runner.corpus.n}-video corpus; "
            f"{pairs} pairs in {secs:.1f} s, extrapolates linearly to the full workload")
    return pairs / secs, desc


def cpu_sample_port(cfg, weights_cpu, corpus):
    """Fallback when the reference sources are not on the box: the same sample through oracle/blim_oracle.py (the CPU
    restatement pinned to the reference by tests/golden)."""
    from oracle import blim_oracle as O
    from oracle import ref_gpu
    torch.set_num_threads(os.cpu_count())
    t0 = time.time()
    for name, (direction, ft, cpn) in ref_gpu.MATRICES.items():
        with torch.no_grad():
            O.compute_scores_x(weights_cpu, cfg, corpus, direction, ft, cpn, topk=CPU_SAMPLE_PAIRS, batch_size=CPU_SAMPLE_PAIRS, rows=[0])
    secs = time.time() - t0
    pairs = 2 * CPU_SAMPLE_PAIRS
    return pairs / secs, (f"oracle port of the reference's per-pair algorithm: one row x top-{CPU_SAMPLE_PAIRS} of each of the six score matrices "
                          f"(6 batched forwards of {CPU_SAMPLE_PAIRS} pairs + criteria), {corpus.n}-video corpus; {pairs} pairs in {secs:.1f} s")


def make_cpu_sampler(cfg, weights_cpu, wl):
    """-> (callable returning (pairs/s, description), kind)."""
    from blim_b200 import synth
    from oracle import ref_harness
    torch.set_num_threads(os.cpu_count())
    corpus = synth.make_corpus(cfg, wl["dataset"], n=wl["n"], n_clips=wl["n_clips"], seed=1)
    if ref_harness.reference_available():
        from oracle import ref_gpu
        runner = ref_gpu.ReferenceRunner(cfg, weights_cpu, corpus, "cpu", dtype=torch.bfloat16)
        return (lambda: cpu_sample_reference(runner)), "reference"
    return (lambda: cpu_sample_port(cfg, weights_cpu, corpus)), "port"


REF_GPU_ROWS = 16   # rows of every score matrix scored by the reference on the GPU (x top-k = 256 pairs per matrix kind at k = 16)


def reference_on_gpu(cfg, weights, corpus, topk, device, rows=REF_GPU_ROWS, dtype=torch.bfloat16):
    """The unmodified reference through PyTorch on one B200 (sdpa attention, batch 16, bf16 autocast): one warm-up row, then
    `rows` rows of all six matrices timed with a device synchronize on both sides.  -> (matrices, dict for the JSON line)."""
    from oracle import ref_gpu
    rr = ref_gpu.ReferenceRunner(cfg, weights, corpus, device, dtype=dtype)
    try:
        rr.all_matrices(rows % corpus.n, 1, topk)
        mats, secs, pairs = rr.all_matrices(0, rows, topk)
    finally:
        rr.close()
    return mats, {"value": pairs / secs, "unit": UNIT, "pairs": pairs, "seconds": secs, "dtype": str(dtype).replace("torch.", ""),
                  "what": f"unmodified reference (retrieval_utils.compute_*_scores_x, sdpa, batch_size_eval 16, {'bf16 autocast' if dtype != torch.float32 else 'fp32, no TF32'}) "
                          f"on one B200: rows 0..{rows - 1} x top-{topk} of all six score matrices, weights resident, wall clock with synchronize on both sides"}


def run_reference_arm(args, wl, rank, world):
    if rank != 0:
        return
    from blim_b200 import synth
    from blim_b200.engine import ModelConfig
    from oracle import ref_harness
    cfg = ModelConfig.qwen2_7b() if wl["model"] == "qwen2_7b" else ModelConfig.tiny()
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    weights_dev = synth.init_weights(cfg, seed=0, device=dev, std=0.02)
    weights = {k: v.cpu() for k, v in weights_dev.items()}
    reference_gpu = None
    if dev == "cuda" and ref_harness.reference_available():
        try:
            corpus = synth.make_corpus(cfg, wl["dataset"], n=args.n or wl["n"], n_clips=wl["n_clips"], seed=1)
            _, reference_gpu = reference_on_gpu(cfg, weights_dev, corpus, args.topk or wl["topk"], torch.device("cuda", 0))
        except Exception as ex:   # a reported number, never a reason to lose the line
            reference_gpu = {"value": None, "unit": UNIT, "what": f"failed: {ex!r}"}
    del weights_dev
    if dev == "cuda":
        torch.cuda.empty_cache()
    sampler, kind = make_cpu_sampler(cfg, weights, wl)
    vals = []
    desc = ""
    for it in range(args.warmup + args.steps):
        v, desc = sampler()
        if it >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    pairs = 2 * CPU_SAMPLE_PAIRS
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * pairs / value, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic", "config": shared_config(wl, args.n or wl["n"] or synth.DATASET_SHAPES[wl["dataset"]]["n"], args.topk or wl["topk"], args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": os.cpu_count(), "kind": kind, "sample": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "reference_gpu": reference_gpu}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ own arm
def parity_with_reference(cfg, eng, corpus, plan, topk, alpha, c, dev, t2v_iv2, v2t_iv2):
    """north_star: "per-pair log-likelihoods within 1e-2, reranked candidate indices and R@K identical", checked on the scores
    the timed steps left in the engine's plan: rows 0..15 of all six matrices against the unmodified reference run on this
    GPU in bf16 (its own 16-bit path, also timed: `reference_gpu`) and rows 0..7 in fp32 (the exact value)."""
    from blim_b200 import retrieval, synth
    from oracle import ref_gpu, ref_harness
    if not ref_harness.reference_available():
        return {"unavailable": "reference sources not on this box (baseline/_ref is written by build() in the build container)"}, None
    model_scores = plan.last_scores
    t2v_c, v2t_c = retrieval.compact_terms(plan, model_scores, cpn=True, full=True)
    names = {"v2t_candidate_likelihood": (v2t_c, "candidate_likelihood"), "v2t_candidate_prior": (v2t_c, "candidate_prior"),
             "v2t_query_likelihood": (v2t_c, "query_likelihood"), "t2v_query_likelihood": (t2v_c, "query_likelihood"),
             "t2v_candidate_likelihood": (t2v_c, "candidate_likelihood"), "t2v_candidate_prior": (t2v_c, "candidate_prior")}
    rows = REF_GPU_ROWS
    m_eng = {n: (comp["idx"][:rows].cpu().numpy(), comp[key][:rows].float().cpu().numpy()) for n, (comp, key) in names.items()}
    weights = synth.init_weights(cfg, seed=0, device=dev, std=0.02)   # the same (seed, index) streams the engine was loaded from
    host = corpus                                                     # ReferenceRunner takes CPU copies of what it needs
    m_b16, reference_gpu = reference_on_gpu(cfg, weights, host, topk, dev, rows=rows)
    rows32 = rows // 2
    m_f32, _ = reference_on_gpu(cfg, weights, host, topk, dev, rows=rows32, dtype=torch.float32)
    del weights
    torch.cuda.empty_cache()
    stat = lambda a, b: {"max": float(np.abs(a - b).max()), "mean": float(np.abs(a - b).mean())}
    scores = {}
    for n in names:
        assert (m_eng[n][0] == m_b16[n][0]).all(), "candidate sets differ"
        scores[n] = {"engine_vs_ref_fp32": stat(m_eng[n][1][:rows32], m_f32[n][1]), "engine_vs_ref_bf16": stat(m_eng[n][1], m_b16[n][1]),
                     "ref_bf16_vs_ref_fp32": stat(m_b16[n][1][:rows32], m_f32[n][1])}
    cut = lambda m, r: {n: (i[:r], x[:r]) for n, (i, x) in m.items()}
    f_eng, f_b16 = ref_gpu.fused_rows(m_eng, host, 0, rows, alpha, c), ref_gpu.fused_rows(m_b16, host, 0, rows, alpha, c)
    f_eng32, f_b1632, f_f32 = (ref_gpu.fused_rows(cut(m, rows32), host, 0, rows32, alpha, c) for m in (m_eng, m_b16, m_f32))
    out = {"pairs_per_matrix": {"vs_ref_bf16": rows * topk, "vs_ref_fp32": rows32 * topk},
           "max_abs_dscore_engine_vs_ref_fp32": max(v["engine_vs_ref_fp32"]["max"] for v in scores.values()),
           "max_abs_dscore_engine_vs_ref_bf16": max(v["engine_vs_ref_bf16"]["max"] for v in scores.values()),
           "max_abs_dscore_ref_bf16_vs_ref_fp32": max(v["ref_bf16_vs_ref_fp32"]["max"] for v in scores.values()),
           "scores": scores,
           "rerank_engine_vs_ref_fp32": ref_gpu.rank_parity(f_f32, f_eng32), "rerank_engine_vs_ref_bf16": ref_gpu.rank_parity(f_b16, f_eng),
           "rerank_ref_bf16_vs_ref_fp32": ref_gpu.rank_parity(f_f32, f_b1632),
           "note": "a = comparator, b = candidate; swapped pairs are candidates whose comparator scores are closer than the listed gap"}
    return out, reference_gpu


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload == "extract":     # the step before the path: its own bench, same JSON contract
        if rank == 0:
            sys.argv = [sys.argv[0], "--steps", str(max(args.steps, 1)), "--warmup", str(args.warmup)] + (["--no-cpu-baseline"] if args.no_cpu_baseline else [])
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import extract_bench
            extract_bench.main()
        return
    specs = []
    for item in args.workload.split(","):
        name, _, k = item.partition("@")
        if name not in WORKLOADS:
            raise SystemExit(f"bench.py: unknown workload {name!r} (choose from {sorted(WORKLOADS)} or 'extract')")
        specs.append((WORKLOADS[name], int(k) if k else 0))
    if args.impl == "reference":
        run_reference_arm(args, specs[0][0], rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the BLiM B200 engine has no CPU fallback")
    if len({wl["model"] for wl, _ in specs}) != 1:
        raise SystemExit("bench.py: the workloads of one run must share the model")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from blim_b200 import synth
    from blim_b200.engine import ModelConfig
    from blim_b200.model import BlimModel

    wl0 = specs[0][0]
    cfg = ModelConfig.qwen2_7b() if wl0["model"] == "qwen2_7b" else ModelConfig.tiny()
    dev = torch.device("cuda", local_rank)
    model = BlimModel(cfg, device=local_rank, gemm_cta_group=args.cta_group, max_run_tokens=args.run_tokens, max_prefix_tokens=args.run_tokens)
    eng = model.engine
    want_cpu = (not args.no_cpu_baseline) and rank == 0 and world == 1
    weights_cpu = {}
    shapes = synth.param_shapes(cfg)
    # stream the random-init parameters through the engine one tensor at a time (same seeds on every rank)
    for idx, name in enumerate(shapes):
        t = synth.init_weight(cfg, name, idx, seed=0, device=dev, std=0.02)
        eng.load_weight(name, t)
        if want_cpu:
            weights_cpu[name] = t.cpu()
        del t
    eng.set_rope(torch.float32)
    for i, (wl, k) in enumerate(specs):
        last = i == len(specs) - 1
        run_workload(args, wl, k or args.topk, cfg, model, dev, rank, world, local_rank, weights_cpu if (want_cpu and last) else None)
        torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


def run_workload(args, wl, topk_override, cfg, model, dev, rank, world, local_rank, weights_cpu):
    """One workload on an engine whose weights are already loaded: device-resident steps, e2e steps, checks, one JSON line."""
    from blim_b200 import evalloop, retrieval, synth
    eng = model.engine
    want_cpu = weights_cpu is not None
    want_parity = (not args.no_parity) and rank == 0 and world == 1 and wl["model"] == "qwen2_7b"
    corpus = synth.make_corpus(cfg, wl["dataset"], n=args.n or wl["n"], n_clips=wl["n_clips"], seed=1, feat_device=dev)
    n, topk = corpus.n, (topk_override or wl["topk"])
    alpha, c = wl["alpha"], wl["c"]

    # device-resident inputs for `value`
    eng.set_videos(corpus.video)
    eng.set_texts(0, corpus.vtg_ids, corpus.vtg_labels)
    eng.set_texts(1, corpus.tvg_ids, corpus.tvg_labels)
    eng.set_video_vocab(corpus.video_vocab, corpus.tvg_video_labels.numpy())
    model.set_tvg_prefix_length(corpus.tvg_prefix_length)
    t2v_iv2, v2t_iv2 = corpus.t2v_iv2.to(dev), corpus.v2t_iv2.to(dev)
    distributed = world > 1

    def step_device():
        plan = retrieval.PairPlan(v2t_iv2, t2v_iv2, topk, dev, engine=eng)
        s = retrieval.score_all(model, plan, cpn=True, full=True, distributed=distributed)
        plan.last_scores = s
        t2v_c, v2t_c = retrieval.compact_terms(plan, s, cpn=True, full=True)
        res, detail = evalloop.fused_rerank(eng, t2v_c, v2t_c, t2v_iv2, v2t_iv2, alpha, c, cpn=True, zero_shot=False)
        return res, plan

    def timed(fn, steps):
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for _ in range(steps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        if distributed:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if distributed:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), out

    for _ in range(args.warmup):
        step_device()
    eng.profile(True)
    eng.profile_read()
    launches0 = eng.kernel_launches()
    flops0 = eng.gemm_flops()
    sampler = ClockSampler(local_rank)
    sampler.start()
    if distributed:
        model.shard_timing = {}
    total_ms, (res, plan) = timed(step_device, args.steps)
    clocks = sampler.stop()
    rank_busy = None
    if distributed and model.shard_timing.get("t1") is not None:   # last step: this rank's own scoring, before the single all-gather
        own = torch.tensor([model.shard_timing["t0"].elapsed_time(model.shard_timing["t1"])], device=dev, dtype=torch.float64)
        allr = [torch.zeros_like(own) for _ in range(world)]
        dist.all_gather(allr, own)
        rank_busy = [round(float(x.item()), 1) for x in allr]
        model.shard_timing = None
    detail = eng.profile_read_detail()
    prof = eng.profile_read()
    eng.profile(False)
    launches = (eng.kernel_launches() - launches0) // max(1, args.steps)
    executed_flops = (eng.gemm_flops() - flops0) / args.steps
    pairs = 2 * n * topk
    ms_per_step = total_ms / args.steps
    value = pairs / (ms_per_step / 1000.0)

    # roofline of the dominant kernel (tcgen05 GEMM): algorithmic GEMM FLOPs / summed GEMM device time
    peaks = measured_peaks()
    f_gemm, f_attn = algorithmic_flops(cfg, corpus, plan, cpn=True, full=True)
    share = 1.0 / world
    gemm_s = prof["gemm_ms"] / 1000.0 / args.steps
    achieved = f_gemm * share / gemm_s / 1e12 if gemm_s > 0 else None
    traffic, traffic_detail = ncu_traffic()
    roofline = {"bound": "tensor", "kernel": "gemm_tcgen05_kernel (all epilogues)", "achieved": achieved, "peak": peaks["bf16_sustained"],
                "unit": "TFLOP/s", "frac": (achieved / peaks["bf16_sustained"]) if achieved else None, "traffic": traffic,
                "traffic_detail": traffic_detail,
                "peak_source": f"{peaks['source']} (sustained cuBLAS bf16; burst {peaks['bf16_burst']})",
                "algorithmic_gemm_flops_per_step": f_gemm, "algorithmic_attention_flops_per_step": f_attn,
                "executed_gemm_flops_per_step": executed_flops, "gemm_launches_per_step": prof["gemm_launches"] // max(1, args.steps),
                "gemm_ms_per_step": prof["gemm_ms"] / args.steps, "attention_ms_per_step": prof["attn_ms"] / args.steps,
                "gemm_share_of_step": prof["gemm_ms"] / total_ms, "attention_share_of_step": prof["attn_ms"] / total_ms,
                "whole_step_tflops": (f_gemm + f_attn) * share / (ms_per_step / 1000.0) / 1e12,
                # per-kind device time (CUDA events on the launching stream, this rank); TFLOP/s = EXECUTED 2*M*N*K / time
                "by_kernel": {k: {"ms_per_step": d["ms"] / args.steps, "launches_per_step": d["launches"] // max(1, args.steps),
                                  "share_of_step": d["ms"] / total_ms,
                                  "tflops": (d["flops"] / (d["ms"] / 1000.0) / 1e12) if d["flops"] > 0 and d["ms"] > 0 else None}
                              for k, d in detail.items()}}

    # end to end through the reference-facing API with HOST inputs (pinned), copies inside the timed region
    e2e = None
    if not args.no_e2e and corpus.video.numel() * 2 > (4 << 30):
        e2e = {"value": None, "unit": UNIT, "skipped": f"{corpus.video.numel() * 2 / 2**30:.1f} GB of features per rank: the pinned host copy of every rank "
                                                        "does not fit this bench's host budget; measured device-resident only"}
    elif not args.no_e2e:
        loader = Loader(corpus, pin=True)
        host_scores = {"v2t": corpus.v2t_iv2.cpu().pin_memory(), "t2v": corpus.t2v_iv2.cpu().pin_memory()}
        eargs = argparse.Namespace(topk=topk, batch_size_eval=16, num_clips=corpus.n_clips, cpn=True, eval=True, resume="synthetic",
                                   dataset=wl["dataset"], distributed=distributed, alpha=list(alpha), c=list(c), iv2_scores=host_scores)
        d2h = [0]

        def step_e2e():
            model._corpus_keys.clear()
            with contextlib.redirect_stdout(sys.stderr):   # the JSON line is the only thing on stdout
                r = evalloop.val_one_epoch(model, loader, None, dev, 0, None, tokenizer=None, args=eargs)
            return r

        step_e2e()
        e2e_ms, res_e2e = timed(step_e2e, args.steps)
        h2d = corpus.video.numel() * 2 + corpus.video_vocab.numel() * 2 + 2 * n * n * 4 + sum(len(x) for x in corpus.vtg_ids + corpus.tvg_ids) * 8
        d2h_bytes = 6 * n * n * 4 + 2 * n * 4
        e2e = {"value": pairs / (e2e_ms / args.steps / 1000.0), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h_bytes),
               "ms_per_step": e2e_ms / args.steps, "api": "blim_b200.evalloop.val_one_epoch (evaluation + CPN/ensemble/rerank), host inputs",
               "recall_blim": res_e2e["blim"]}

    # parity of the timed path's own results with the unmodified reference run on this GPU (checker, after the timed region)
    rank_parity = reference_gpu = None
    if want_parity:
        try:
            rank_parity, reference_gpu = parity_with_reference(cfg, eng, corpus, plan, topk, alpha, c, dev, t2v_iv2, v2t_iv2)
        except Exception as ex:
            rank_parity = {"error": repr(ex)}

    cpu_baseline = None
    if want_cpu:
        try:
            eng.close()   # free the engine's HBM and stop its profiling events before the host-side baseline
            sampler, kind = make_cpu_sampler(cfg, weights_cpu, wl)
            v, desc = sampler()
            cpu_baseline = {"value": v, "unit": UNIT, "cores": os.cpu_count(), "kind": kind, "sample": desc}
        except Exception as ex:  # the baseline is a reported number, never a reason to lose the bench line
            cpu_baseline = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {ex!r}"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": DTYPE,
                "data": "synthetic",
                "config": shared_config(wl, n, topk, world),
                "engine": {"precision": DTYPE_NOTE, "unique_vtg_pairs": int(plan.union_key.numel()), "rank_scoring_ms_last_step": rank_busy},
                "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu_baseline,
                "rank_parity": rank_parity, "reference_gpu": reference_gpu, "recall_blim": res}
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
