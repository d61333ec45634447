/*
 * blim_b200 -- C ABI of the B200-native BLiM scoring engine (libblim_b200.so).
 *
 * Scope: the bidirectional likelihood scoring path of mlvlab/BLiM and nothing else:
 *   p(t|v) "VTG" and p(v|t) "TVG" prefill-only scoring, their CPN priors, the InternVideo2 ensemble and the rerank.
 * Every entry point cites the reference interface (file:line under the reference repo) it replaces.
 *
 * Conventions
 *   - One engine per process / device, not re-entrant, driven by one host thread (the reference is one process per
 *     GPU, single host thread: README.md:117, util/misc.py:199-229).
 *   - All work is enqueued on the cudaStream_t passed in (PyTorch's current stream); nothing synchronises the device
 *     unless stated.  `void* stream` is a cudaStream_t.
 *   - "dev" pointers are device memory owned by the caller; "host" pointers are ordinary host memory.
 *   - Return value: 0 = ok, non-zero = error; blim_last_error() returns the message.  No exception crosses the ABI.
 *   - dtype codes: 0 = float32, 1 = bfloat16, 2 = float16.
 */
#ifndef BLIM_B200_H_
#define BLIM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct blim_engine blim_engine;

/* Architecture of VideoChatFlashQwenForCausalLM (reference: videochat_flash/modeling_videochat_flash.py:572-590,
 * Qwen2Config fields read at modeling_qwen2_flash.py:219-246, 920-950). */
typedef struct blim_model_cfg {
  int32_t hidden_size;
  int32_t num_layers;
  int32_t num_heads;
  int32_t num_kv_heads;
  int32_t head_dim;          /* 64 or 128 */
  int32_t intermediate_size; /* multiple of 128 */
  int32_t vocab_size;
  int32_t mm_hidden_size;    /* video feature width (1024), multiple of 64 */
  int32_t tokens_per_clip;   /* 64: rows per clip of the pre-extracted features (dataloader/base_dataset.py:28) */
  int32_t max_positions;     /* rows of the rotary table */
  float rms_norm_eps;
  int32_t max_run_tokens;    /* workspace: tokens per decoder run (0 = default 32768) */
  int32_t max_prefix_tokens; /* workspace: rows of the shared-prefix KV cache (0 = default 32768) */
  int32_t gemm_cta_group;    /* 1 = one CTA per 128x256 tile, 2 = CTA pairs (cta_group::2, 256x256 tiles); 0 = default (2) */
} blim_model_cfg;

/* Score kinds of blim_score_pairs.
 *   VTG        score_VTG(v,t)   -> v2t "candidate_likelihood" / t2v "query_likelihood"   retrieval_utils.py:91-97, 124-134
 *   VTG_PRIOR  prior_VTG(t)     -> v2t "candidate_prior" (cpn mask)                      retrieval_utils.py:92-93
 *   TVG        score_TVG(t,v)   -> v2t "query_likelihood" / t2v "candidate_likelihood"   retrieval_utils.py:98-108, 135-150
 *   TVG_PRIOR  prior_TVG(t,v)   -> t2v "candidate_prior" (cpn mask)                      retrieval_utils.py:142-143 */
enum { BLIM_VTG = 0, BLIM_VTG_PRIOR = 1, BLIM_TVG = 2, BLIM_TVG_PRIOR = 3 };

/* Which tokenised prompt family a text table holds (dataloader/base_dataset.py:60-84 vs 86-105). */
enum { BLIM_TEXTS_VTG = 0, BLIM_TEXTS_TVG = 1 };

/* Create / destroy.  Replaces model construction + .to(device) (main.py:97). */
int blim_create(const blim_model_cfg* cfg, int device, blim_engine** out);
void blim_destroy(blim_engine* e);
const char* blim_last_error(const blim_engine* e); /* e may be NULL: last create error */

/* Load one parameter by its reference state_dict key (e.g. "model.layers.3.self_attn.q_proj.weight",
 * "model.mm_projector.tvg_mlp.0.bias", "lm_head.weight", "visual_head.weight").  The engine converts and repacks into
 * its own bf16 layouts (fused QKV, 128-row interleaved gate|up); the caller keeps ownership of `dev_ptr`.
 * Replaces from_pretrained / load_state_dict (main.py:97, util/misc.py:303-311). */
int blim_load_weight(blim_engine* e, const char* name, const void* dev_ptr, int dtype, const int64_t* shape, int ndim, void* stream);

/* Rotary tables cos/sin [max_positions, head_dim/2] fp32 (device), built by the host exactly like
 * Qwen2RotaryEmbedding._set_cos_sin_cache (modeling_qwen2_flash.py:119-127), including its cast to the model dtype. */
int blim_set_rope(blim_engine* e, const float* cos_dev, const float* sin_dev, int n_positions, void* stream);

/* Corpus.  Replaces the per-row H2D copies and list handling of evaluation() (retrieval_utils.py:179-197, 209-210).
 *   videos : [n_videos, n_clips, tokens_per_clip, mm_hidden] features (device; copied and converted to bf16)
 *   texts  : pad-stripped token ids with the -200 image sentinel and labels with -100 = ignore (host, ragged,
 *            offsets[n_texts + 1]) -- the output contract of BaseDataset.get_vtg_id / get_tvg_id
 *   vocab  : video_vocab [n_vocab, n_clips, mm_hidden] (device) and tvg_video_labels[n_videos] (host)
 *            (dataloader/base_dataset.py:33-37, 114) */
int blim_set_videos(blim_engine* e, const void* feats_dev, int dtype, int n_videos, int n_clips, void* stream);
int blim_set_texts(blim_engine* e, int which, const int32_t* ids_host, const int32_t* labels_host, const int64_t* offsets_host, int n_texts);
int blim_set_video_vocab(blim_engine* e, const void* vocab_dev, int dtype, int n_vocab, const int32_t* video_labels_host, int n_videos,
                         void* stream);
/* video_vocab built on the device from the features of blim_set_videos: vocab[label[v]] = features[v].mean(tokens)
 * (get_video_vocab, dataloader/base_dataset.py:33-37) -- no host-side mean, no vocab upload. */
int blim_build_video_vocab(blim_engine* e, const int32_t* video_labels_host, int n_videos, int n_vocab, void* stream);
int blim_set_tvg_prefix_length(blim_engine* e, int n); /* set_tvg_prefix_length, modeling_videochat_flash.py:592 */

/* Score n_pairs (video, text) pairs of one kind; pair_v / pair_t are host arrays, out_scores_dev[n_pairs] is device fp32.
 * Duplicate work is shared: a video prefix is prefilled once for all its captions, a text prefix once for all its
 * candidate videos, identical (v,t) keys are scored once, priors once per distinct conditioning.
 * Replaces the forward passes of compute_v2t_scores_x / compute_t2v_scores_x (retrieval_utils.py:48-153). */
int blim_score_pairs(blim_engine* e, int kind, const int32_t* pair_v_host, const int32_t* pair_t_host, int64_t n_pairs,
                     float* out_scores_dev, void* stream);

/* Compatibility path with the reference model's forward signature (modeling_videochat_flash.py:601-629 called as
 * model(inputs_embeds=..., attention_mask=...)): embeds [B, L, H] bf16, key mask [B, L] int32 (1 = visible key);
 * writes logits [B, L, V] fp32 (if not NULL) and final-norm hidden states [B, L, H] bf16 (if not NULL). */
int blim_forward_logits(blim_engine* e, const void* embeds_dev, const int32_t* mask_dev_or_null, int B, int L, float* logits_dev,
                        void* hidden_dev, void* stream);

/* Projector: features [n_rows, mm_hidden] bf16 -> mlp / tvg_mlp output [n_rows, hidden] bf16
 * (mm_projector_builder.py:156-159). */
int blim_project_video(blim_engine* e, const void* feats_dev, int n_rows, int tvg, void* out_dev, void* stream);
/* visual_head: [n_rows, hidden] bf16 -> [n_rows, mm_hidden] bf16 (forward_visual, modeling_videochat_flash.py:598-599). */
int blim_forward_visual(blim_engine* e, const void* hidden_dev, int n_rows, void* out_dev, void* stream);
/* embed_tokens lookup: ids (device int32) -> [n, hidden] bf16 (modeling_videochat_flash.py:388,403). */
int blim_embed_tokens(blim_engine* e, const int32_t* ids_dev, int n, void* out_dev, void* stream);

/* CPN + ensemble + rerank (val_one_epoch training_utils.py:154-165, get_recall training_utils.py:173-221) on compact
 * candidate arrays.  One direction per call; rows [row0, row0 + n_rows) of the full problem (ground truth of row r is
 * column row0 + r).  All arrays device: cand_idx int32 [n_rows, k]; cand / prior / query fp32 [n_rows, k] (prior, query
 * may be NULL); iv2 fp32 [n_rows, n_cols].  Outputs: fused fp64 [n_rows, k], order int32 [n_rows, k] (candidate columns
 * by descending fused score), gt_rank int32 [n_rows], zero_count int32 [1] (accumulated). */
typedef struct blim_fuse_cfg {
  double alpha;      /* args.alpha[0] (t2v) or [1] (v2t) */
  double c_query;    /* args.c[0] (t2v) or args.c[1] (v2t) */
  double c_ens;      /* args.c[2] (t2v) or args.c[3] (v2t) */
  int32_t use_prior; /* args.cpn and a prior matrix exists for this direction */
  int32_t use_query; /* 0: blim = cpn (zero-shot v2t, training_utils.py:162) */
  int32_t cpn_zero_f64; /* zero-shot t2v: cpn term is np.zeros float64 (training_utils.py:154,161) */
} blim_fuse_cfg;
int blim_fuse_rerank(blim_engine* e, const blim_fuse_cfg* cfg, const int32_t* cand_idx, const float* cand, const float* prior,
                     const float* query, const float* iv2, int n_rows, int n_cols, int k, int row0, double* fused_out,
                     int32_t* order_out, int32_t* gt_rank_out, int32_t* zero_count, void* stream);
/* get_recall's rank search on a dense score matrix [n_rows, n_cols] fp32 (device). */
int blim_rank_dense(blim_engine* e, const float* mat, int n_rows, int n_cols, int row0, int32_t* gt_rank_out, int32_t* zero_count,
                    void* stream);
/* Stage-1 candidates: per-row top-k (descending, ties lowest column first) of a dense fp32 similarity matrix -- replaces
 * sims.topk(k) on the InternVideo2 rows (retrieval_utils.py:52,117).  All device arrays; idx int32 / val fp32 [n_rows, k]. */
int blim_topk_rows(blim_engine* e, const float* mat, int n_rows, int n_cols, int k, int32_t* idx_out, float* val_out, void* stream);
/* dense[row[i], col[i]] = val[i] after an optional fill (retrieval_utils.py:219, 110, 152); all device arrays. */
int blim_scatter_scores(blim_engine* e, float* dense, int n_rows, int n_cols, int do_fill, float fill, const int32_t* row,
                        const int32_t* col, const float* val, int64_t n, void* stream);

/* Counters for bench.py: kernels launched by this engine since creation / FLOPs of its tensor-core GEMMs. */
int64_t blim_kernel_launches(const blim_engine* e);
double blim_gemm_flops(const blim_engine* e);

/* Measurement support for bench.py: with profiling enabled every tcgen05 GEMM / attention launch is bracketed by CUDA
 * events on the launching stream; blim_profile_read synchronises the device, returns the summed device time and launch
 * count per kernel family since the last read, and resets. */
int blim_profile(blim_engine* e, int enable);
int blim_profile_read(blim_engine* e, double* gemm_ms, double* attn_ms, int64_t* gemm_launches, int64_t* attn_launches);
/* Per-kind breakdown of the intervals recorded so far, without resetting them (call it before blim_profile_read).
 * n >= 8 entries per array: 0 QKV+RoPE GEMM, 1 o_proj, 2 gate|up+SwiGLU, 3 down_proj, 4 LM / TVG head + log-sum-exp,
 * 5 other GEMMs (projectors, visual head), 6 attention, 7 RMSNorm.  ms = summed device time, flops = executed 2*M*N*K
 * of the GEMM launches (0 for attention / RMSNorm). */
int blim_profile_read_detail(blim_engine* e, int n, double* ms, double* flops, int64_t* launches);

/* ---- Multi-GPU exchange (one process / engine per GPU).  Replaces the reference's dist.barrier() + 2-6 dense N x N
 * all_reduce(SUM) of -100-filled matrices (retrieval_utils.py:252-262) by ONE all-gather of compact per-pair scores.
 * NCCL is resolved at run time (dlopen: the copy already in the process, $BLIM_NCCL_LIB, or the linker path).
 *   blim_comm_unique_id : rank 0 creates the 128-byte ncclUniqueId; the caller distributes it (any channel).
 *   blim_comm_init      : every rank joins (ncclCommInitRank on the engine's device); collective, call on all ranks.
 *   blim_allgather_scores: recv[r*count .. (r+1)*count) = rank r's send[0 .. count), fp32 device buffers, enqueued on
 *                         `stream` (the compute stream: no host synchronisation between scoring and rerank).
 *                         nccl_comm = an ncclComm_t of the caller, or NULL for the engine's own communicator. */
int blim_comm_unique_id(void* id_out128);
int blim_comm_init(blim_engine* e, const void* id128, int rank, int world);
int blim_comm_destroy(blim_engine* e);
int blim_allgather_scores(blim_engine* e, void* nccl_comm, const float* send_dev, float* recv_dev, int64_t count_per_rank, void* stream);

/* 16-bit format of the engine's tensor-core operands (weights as stored after blim_load_weight, activations), in
 * blim_load_weight's dtype codes: 2 = fp16 (default build), 1 = bf16 (-DBLIM_ACT_BF16); see csrc/act_type.cuh.  Only the
 * debug entries below expose operand-format buffers; every other entry point speaks the dtypes documented with it. */
int blim_act_dtype(void);

/* Debug / unit-test entry: C = epilogue(A[M,K] · W[N,K]^T) with the engine's tcgen05 GEMM.  A and W are in the operand
 * format (blim_act_dtype); "act" outputs are in that format too.
 * epilogue: 0 = act store, 1 = act store + bias, 2 = act store + bias + GELU, 3 = fp32 store,
 *           4 = fp32 residual add (C += A·W^T), 5 = SwiGLU (W = 128-row interleaved gate|up, C act [M, N/2]),
 *           6 = log-sum-exp (C fp32 [M]: logp of target[M] at `scale`). */
int blim_debug_gemm(blim_engine* e, int epilogue, const void* A, const void* W, void* C, int M, int N, int K, const float* bias,
                    const int32_t* target, float scale, int cta_group, void* stream);

/* Debug / unit-test entry: single-CTA tcgen05 probe C[128,N] = A[128,K] · B with thread-staged (manually swizzled)
 * operands; B is [N][K] (K-major) or, with bit 0 of b_mn_major, [K][N] with explicit descriptor byte offsets.  Bit 1 of
 * b_mn_major: A holds fp16 instead of bf16; bit 2: B holds fp16.  Pins the shared-memory descriptor semantics the
 * attention kernel relies on, and that both operands of an MMA must share one format
 * (tests/test_umma_probe_gpu.py). */
int blim_debug_umma(blim_engine* e, const void* A, const void* B, float* C, int K, int N, int b_mn_major, uint32_t lbo, uint32_t sbo,
                    uint32_t kstep_bytes, void* stream);

/* Debug / unit-test entry, HOST ONLY (needs neither an engine nor a device): the batch planner blim_score_pairs uses to
 * cut a scoring job into decoder runs, on a described workload.  Unit u (one shared prefix: a video with its captions, a
 * text with its candidate videos) has unit_prefix_len[u] prefix tokens and unit_item_count[u] suffix sequences whose
 * lengths follow each other in item_suf_len, unit after unit.  Capacities as in the model configuration struct: max_prefix_tokens,
 * max_run_tokens; max_units prefixes and max_items suffix sequences per batch, reserve_rows cache rows kept for the
 * shared prompt header.  Writes the batch index of every suffix sequence and returns the number of batches, -1 when the
 * workload does not fit.  Contract (tests/test_host_cpu.py): capacities hold, order is preserved, and a unit that fits
 * a batch on its own is never split -- the precondition for scores that do not depend on the sharding. */
int blim_debug_plan_batches(int max_prefix_tokens, int max_run_tokens, int max_units, int max_items, int reserve_rows,
                            const int32_t* unit_prefix_len, const int32_t* unit_item_count, int n_units,
                            const int32_t* item_suf_len, int32_t* batch_of_item_out);

#ifdef __cplusplus
}
#endif
#endif /* BLIM_B200_H_ */
