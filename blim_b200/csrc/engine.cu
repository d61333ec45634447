// Host side of the B200 BLiM scoring engine + its C ABI (include/blim_b200.h).
//
// The host code here is the "scheduler" of SURVEY.md 7(6): it turns a list of (video, text) pairs into a few large
// decoder runs so that every GEMM sees thousands of rows:
//   * VTG      : per video one PREFIX sequence [system/user header | 64*n_clips projected visual rows | prompt tail]
//                prefilled once into the prefix KV cache; every caption of that video is a SUFFIX sequence that attends
//                to the cached prefix (cascade attention) -- replaces the per-pair 300-token prefill of
//                compute_v2t_scores_x / compute_t2v_scores_x (reference retrieval_utils.py:62-97, 121-134).
//   * VTG prior: the CPN mask hides the visual rows as keys (modeling_videochat_flash.py:433), so the prior only depends
//                on the text: one shared prefix (header + prompt tail at gapped rotary positions), one suffix per text.
//   * TVG      : prefix = text part of the TVG prompt (per text), suffix = the first n_clips-1 pooled visual rows.
//   * TVG prior: prefix = the tvg_prefix_length visible header tokens (modeling_videochat_flash.py:414-417), suffix =
//                [last text token, invisible as a key] + pooled visual rows; shared by all texts of equal length.
// Scoring never materialises logits: LM head / TVG head run through the fused log-sum-exp GEMM epilogue.
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/blim_b200.h"
#include "attention.cuh"
#include "attention_tc.cuh"
#include "attention_ws.cuh"
#include "comm.cuh"
#include "gemm_sm100.cuh"
#include "kernels_misc.cuh"
#include "umma_probe.cuh"

using namespace blim;
typedef __nv_bfloat16 bf16;   // the embedding table and the bf16 tensors of the compat entry points; GEMM operands are act_t (act_type.cuh)

static std::string g_create_error;

namespace {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

// sub-tensors of a layer that blim_load_weight has filled (the fused buffers are allocated by the first of them)
enum : unsigned {
  kLdQ = 1u << 0, kLdK = 1u << 1, kLdV = 1u << 2, kLdQb = 1u << 3, kLdKb = 1u << 4, kLdVb = 1u << 5, kLdO = 1u << 6, kLdGate = 1u << 7,
  kLdUp = 1u << 8, kLdDown = 1u << 9, kLdLn1 = 1u << 10, kLdLn2 = 1u << 11, kLdAll = (1u << 12) - 1
};
static const char* const kLdNames[12] = {"self_attn.q_proj.weight", "self_attn.k_proj.weight", "self_attn.v_proj.weight", "self_attn.q_proj.bias",
                                         "self_attn.k_proj.bias", "self_attn.v_proj.bias", "self_attn.o_proj.weight", "mlp.gate_proj.weight",
                                         "mlp.up_proj.weight", "mlp.down_proj.weight", "input_layernorm.weight", "post_attention_layernorm.weight"};
struct LayerW {
  DevBuf w_qkv, w_o, w_gu, w_down, b_qkv, ln1, ln2;
  unsigned loaded = 0;  // kLd* bits
  bool folded = false;  // input_layernorm / post_attention_layernorm weights folded into w_qkv / w_gu (fused RMSNorm)
};
struct ProjW {
  DevBuf w0, b0, w2, b2;
};

struct PromptGroup {
  std::vector<int32_t> pre, post;  // prompt ids before / after the image sentinel (VTG), or the visible header (TVG prior)
};
struct TextTable {
  int n = 0;
  std::vector<int32_t> ids, labels;
  std::vector<int64_t> off;
  std::vector<int> img_pos;   // index of the -200 sentinel
  std::vector<int> first_lab; // first index with label != -100 (VTG)
  std::vector<int> group;     // prompt group of each text
  std::vector<PromptGroup> groups;
};

// One flat decoder run: T tokens, S sequences.
struct Run {
  std::vector<int> tok_src, tok_pos, tok_slot;
  std::vector<uint8_t> key_valid;
  std::vector<AttnSeq> seqs;
  bool any_invalid = false;
  bool any_slot = false;  // some token's K/V row differs from its run index (prefix runs behind replicated root rows)
  int unit = 0;           // tag for the sequences begun from now on (AttnSeq::unit)
  int T() const { return static_cast<int>(tok_src.size()); }
  void clear() {
    tok_src.clear(); tok_pos.clear(); tok_slot.clear(); key_valid.clear(); seqs.clear(); any_invalid = false; any_slot = false; unit = 0;
  }
  // returns index of the first token; b_start = K/V row of the sequence's first token (-1: its run index)
  int begin_seq(int a_start, int a_len, int b_start = -1) {
    AttnSeq s;
    s.q_start = T(); s.q_len = 0; s.a_start = a_start; s.a_len = a_len; s.b_start = b_start < 0 ? T() : b_start;
    static const bool per_seq = getenv("BLIM_ATTN_PERSEQ") != nullptr;   // A/B: never stack two sequences in one tile
    s.unit = per_seq ? static_cast<int>(seqs.size()) : unit;
    if (s.b_start != s.q_start) any_slot = true;
    seqs.push_back(s);
    return s.q_start;
  }
  void push(int src, int pos, bool valid = true) {
    tok_slot.push_back(seqs.back().b_start + seqs.back().q_len);
    tok_src.push_back(src); tok_pos.push_back(pos); key_valid.push_back(valid ? 1 : 0);
    if (!valid) any_invalid = true;
    seqs.back().q_len++;
  }
  void end_seq() {
    if (seqs.back().q_len == 0) seqs.pop_back();
  }
};

}  // namespace

struct blim_engine {
  blim_model_cfg cfg;
  int device = 0;
  std::string err;
  GemmLaunchCtx gemm;
  long long launches = 0;
  double flops = 0.0;

  int H, NL, NH, NKV, DH, I, V, MM, TPC, NQ, NKVD, NQKV, G;
  int Tmax, Pmax, Umax;

  // weights
  DevBuf embed, lm_head, visual_head, norm;
  std::vector<LayerW> layers;
  ProjW proj[2];  // 0 = mlp (VTG), 1 = tvg_mlp
  DevBuf rope_cos, rope_sin;
  int rope_n = 0;
  long long weights_loaded = 0;

  // corpus
  DevBuf feats;  // [n_videos, n_clips*TPC, MM] bf16
  int n_videos = 0, n_clips = 0;
  TextTable texts[2];
  DevBuf vocab;  // [n_clips, n_vocab, MM] bf16
  int n_vocab = 0;
  std::vector<int32_t> video_labels;
  int tvg_prefix_len = 21;
  DevBuf tvg_vis;  // [n_videos * n_clips, H] bf16 pooled tvg_mlp rows
  bool tvg_vis_ready = false;

  // workspaces
  DevBuf x, xn, q, attn, k_own, v_own, act, kp, vp, prefix_last, vis, proj_tmp, lm_a, pred, partial, tgt_logit, logp, uniq_scores;
  DevBuf io_in, io_out;      // bf16 <-> activation-format staging of the compat entry points (allocated on first use)
  DevBuf vis_in, d_vis_idx;  // projector input staging for non-contiguous video sets (gathered feature rows + their indices)
  DevBuf d_tok_slot, d_tok_src, d_tok_pos, d_key_valid, d_seqs, d_works, d_idx, d_targets, d_row_off, d_map, d_seq_start;
  int ksplit = 1, nsplit = 0;             // K slices of the long-K residual GEMM (1 = off) / N slices of gate|up (0 = by weight size)
  size_t ksplit_min_bytes = 100u << 20;   // weights larger than this are K-sliced
  int resid_tma = 0;                      // BLIM_RESID_TMA=1|2: residual add of o_proj (| and down_proj) through a TMA reduce (EpiResidTma)
  int nsplit_min_rows = 20000;            // gate|up is only N-sliced for runs of at least this many rows: below, A fits the L2 next
                                          // to the streaming weights and the slices only add wave quantisation (BLIM_GEMM_NSPLIT_MIN_ROWS)
  int attn_version = kAttnWarpSpecialized;  // BLIM_ATTN=tc2p|tc2: the round-1 kernels (A/B against the warp-specialised default)
  CUtensorMap tm_q;                          // e->q as a (head_dim, head, token) tensor: Q tiles by TMA (attention_ws.cuh)
  uint8_t* arena = nullptr;  // pinned staging arena for scheduler metadata (see upload())
  size_t arena_cap = 0, arena_off = 0;
  bool arena_disabled = false;
  bool root_share = true;   // shared prompt-header root for the prefixes (BLIM_ROOT=0 disables)
  NcclComm comm = nullptr;   // blim_comm_init: this engine's NCCL communicator (one rank per engine / process / GPU)
  int comm_rank = 0, comm_world = 1;
  bool attr_fuse = false, attr_topk = false;   // cudaFuncSetAttribute done on this engine's device
  bool fuse_norm = false;  // BLIM_FUSE_NORM=1: RMSNorm fused into the GEMMs around it (measured slower than the standalone kernel)
  DevBuf ssq, rstd;
  CUtensorMap tm_kp, tm_vp, tm_kown, tm_vown;  // K / V buffers as TMA tensors (tcgen05 attention)
  size_t partial_tiles = 0;

  // optional per-launch device timing (bench.py roofline): CUDA events around every GEMM / attention launch
  bool profiling = false;
  struct Timed { cudaEvent_t a, b; int cat; int sub; double flops; };
  std::vector<Timed> timed;
  std::vector<cudaEvent_t> event_pool;
  cudaEvent_t get_event() {
    if (!event_pool.empty()) { cudaEvent_t ev = event_pool.back(); event_pool.pop_back(); return ev; }
    cudaEvent_t ev; cudaEventCreate(&ev); return ev;
  }
  void tic(int cat, cudaStream_t st, int sub = 0, double flops = 0.0) {
    if (!profiling) return;
    Timed t; t.a = get_event(); t.b = get_event(); t.cat = cat; t.sub = sub; t.flops = flops;
    cudaEventRecord(t.a, st);
    timed.push_back(t);
  }
  void toc(cudaStream_t st) {
    if (!profiling) return;
    cudaEventRecord(timed.back().b, st);
  }

  int fail(const std::string& m) {
    err = m;
    return 1;
  }
  int fail_cuda(const char* what, cudaError_t e) {
    err = std::string(what) + ": " + cudaGetErrorString(e);
    return 1;
  }
};

#define CKE(expr)                                                   \
  do {                                                              \
    cudaError_t _e = (expr);                                        \
    if (_e != cudaSuccess) return e->fail_cuda(#expr, _e);          \
  } while (0)
#define CKL()                                                       \
  do {                                                              \
    e->launches++;                                                  \
    cudaError_t _e = cudaGetLastError();                            \
    if (_e != cudaSuccess) return e->fail_cuda("kernel launch", _e); \
  } while (0)
#define CKR(expr)                \
  do {                           \
    int _r = (expr);             \
    if (_r != 0) return _r;      \
  } while (0)

static inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// ------------------------------------------------------------------------------------------------ GEMM wrappers
// profile sub-categories (blim_profile_read_detail): which contraction a GEMM launch is
enum { kProfQkv = 0, kProfOProj = 1, kProfGateUp = 2, kProfDown = 3, kProfLse = 4, kProfOtherGemm = 5, kProfAttn = 6, kProfNorm = 7, kProfCats = 8 };
template <class Epi> struct EpiProf { static int sub(const blim_engine*, int) { return kProfOtherGemm; } };
template <int D> struct EpiProf<EpiQkvRope<D>> { static int sub(const blim_engine*, int) { return kProfQkv; } };
template <> struct EpiProf<EpiSwiglu> { static int sub(const blim_engine*, int) { return kProfGateUp; } };
template <> struct EpiProf<EpiLse> { static int sub(const blim_engine*, int) { return kProfLse; } };
template <> struct EpiProf<EpiResidTma> { static int sub(const blim_engine* e, int K) { return K == e->NQ ? kProfOProj : kProfDown; } };
template <int G> struct EpiProf<EpiResidT<G>> { static int sub(const blim_engine* e, int K) { return K == e->NQ ? kProfOProj : kProfDown; } };   // down_proj may be K-sliced
template <> struct EpiProf<EpiResidNorm> { static int sub(const blim_engine* e, int K) { return K == e->NQ ? kProfOProj : kProfDown; } };

template <class Epi>
// A = activation operand, W = weight operand, both in the operand format act_t.
static int gemm(blim_engine* e, const void* A, int lda, const void* W, int ldw, int M, int N, int K, const typename Epi::Params& p,
                cudaStream_t st, int b_fmt = kActFmt, int a_fmt = kActFmt) {
  if (M <= 0) return 0;
  e->tic(0, st, EpiProf<Epi>::sub(e, K), 2.0 * M * static_cast<double>(N) * K);
  cudaError_t r = launch_gemm<Epi>(e->gemm, A, lda, W, ldw, M, N, K, p, st, a_fmt, b_fmt);
  e->toc(st);
  if (r != cudaSuccess) return e->fail_cuda("tcgen05 gemm launch", r);
  e->flops += 2.0 * M * static_cast<double>(N) * K;
  return 0;
}

static int gemm_qkv(blim_engine* e, const act_t* A, const LayerW& w, int M, act_t* q_out, act_t* k_out, act_t* v_out, const int* pos,
                    const int* kv_slot, const float* rstd, cudaStream_t st) {
  if (e->DH == 128) {
    EpiQkvRope<128>::Params p{q_out, k_out, v_out, w.b_qkv.as<float>(), pos, kv_slot, e->rope_cos.as<float>(), e->rope_sin.as<float>(),
                              e->NQ, e->NKVD, e->rope_n, rstd};
    return gemm<EpiQkvRope<128>>(e, A, e->H, w.w_qkv.as<act_t>(), e->H, M, e->NQKV, e->H, p, st);
  }
  EpiQkvRope<64>::Params p{q_out, k_out, v_out, w.b_qkv.as<float>(), pos, kv_slot, e->rope_cos.as<float>(), e->rope_sin.as<float>(),
                           e->NQ, e->NKVD, e->rope_n, rstd};
  return gemm<EpiQkvRope<64>>(e, A, e->H, w.w_qkv.as<act_t>(), e->H, M, e->NQKV, e->H, p, st);
}

// ------------------------------------------------------------------------------------------------ create / destroy
extern "C" const char* blim_last_error(const blim_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

extern "C" void blim_destroy(blim_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  DevBuf* bufs[] = {&e->embed, &e->lm_head, &e->visual_head, &e->norm, &e->rope_cos, &e->rope_sin, &e->feats, &e->vocab, &e->tvg_vis,
                    &e->x, &e->xn, &e->q, &e->attn, &e->k_own, &e->v_own, &e->act, &e->kp, &e->vp, &e->prefix_last, &e->vis, &e->vis_in, &e->d_vis_idx,
                    &e->proj_tmp, &e->lm_a, &e->pred, &e->partial, &e->tgt_logit, &e->logp, &e->uniq_scores, &e->d_tok_src,
                    &e->d_tok_pos, &e->d_key_valid, &e->d_seqs, &e->d_works, &e->d_idx, &e->d_targets, &e->d_row_off, &e->d_map, &e->d_seq_start, &e->ssq, &e->rstd, &e->d_tok_slot, &e->io_in, &e->io_out};
  for (DevBuf* b : bufs) b->release();
  for (LayerW& l : e->layers) {
    l.w_qkv.release(); l.w_o.release(); l.w_gu.release(); l.w_down.release(); l.b_qkv.release(); l.ln1.release(); l.ln2.release();
  }
  for (ProjW& p : e->proj) { p.w0.release(); p.b0.release(); p.w2.release(); p.b2.release(); }
  if (e->comm && nccl_api().CommDestroy) nccl_api().CommDestroy(e->comm);
  if (e->arena) cudaFreeHost(e->arena);
  for (auto& t : e->timed) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
  for (cudaEvent_t ev : e->event_pool) cudaEventDestroy(ev);
  delete e;
}

extern "C" int blim_create(const blim_model_cfg* cfg, int device, blim_engine** out) {
  if (!cfg || !out) { g_create_error = "null argument"; return 1; }
  *out = nullptr;
  int n_dev = 0;
  cudaError_t ce = cudaGetDeviceCount(&n_dev);
  if (ce != cudaSuccess || n_dev <= 0) {
    g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(ce) + " (the engine has no CPU fallback)";
    return 1;
  }
  if ((ce = cudaSetDevice(device)) != cudaSuccess) { g_create_error = cudaGetErrorString(ce); return 1; }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major != 10) {
    g_create_error = "blim_b200 needs an sm_100a device (got sm_" + std::to_string(prop.major) + std::to_string(prop.minor) + ")";
    return 1;
  }
  blim_engine* e = new blim_engine();
  e->cfg = *cfg;
  e->device = device;
  e->H = cfg->hidden_size; e->NL = cfg->num_layers; e->NH = cfg->num_heads; e->NKV = cfg->num_kv_heads; e->DH = cfg->head_dim;
  e->I = cfg->intermediate_size; e->V = cfg->vocab_size; e->MM = cfg->mm_hidden_size; e->TPC = cfg->tokens_per_clip;
  e->NQ = e->NH * e->DH; e->NKVD = e->NKV * e->DH; e->NQKV = e->NQ + 2 * e->NKVD;
  e->G = e->NKV > 0 ? e->NH / e->NKV : 0;
  // 49 152-token runs: measured against 32 768 on the same box, +0.9 % on the full C2 job and -1.9 % time on one rank's
  // share of an 8-GPU job (its 125 video prefixes then fit ONE prefix run: fewer, larger launches, less wave quantisation);
  // 65 536 gives nothing more.  ~10 GB of workspace at 7B.
  e->Tmax = cfg->max_run_tokens > 0 ? cfg->max_run_tokens : 49152;
  e->Pmax = cfg->max_prefix_tokens > 0 ? cfg->max_prefix_tokens : 49152;
  e->Umax = std::min(e->Pmax, 8192);
  e->gemm.num_sms = prop.multiProcessorCount;
  e->gemm.device = device;
  {
    const char* a = getenv("BLIM_ATTN");
    e->attn_version = (a && std::string(a) == "tc2") ? kAttnPerItem : (a && std::string(a) == "tc2p") ? kAttnPersistent : kAttnWarpSpecialized;
    const char* rs = getenv("BLIM_ROOT");
    e->root_share = !(rs && std::string(rs) == "0");
    const char* f = getenv("BLIM_FUSE_NORM");
    e->fuse_norm = f && std::string(f) == "1";
  }
  e->gemm.cta_group = cfg->gemm_cta_group == 1 ? 1 : 2;  // default: CTA pairs (cta_group::2)
  if (const char* v = getenv("BLIM_GEMM_KSPLIT")) e->ksplit = std::max(1, std::min(8, atoi(v)));
  if (const char* v = getenv("BLIM_GEMM_NSPLIT")) e->nsplit = std::max(1, std::min(16, atoi(v)));
  if (const char* v = getenv("BLIM_RESID_TMA")) e->resid_tma = std::max(0, std::min(2, atoi(v)));
  if (const char* v = getenv("BLIM_GEMM_NSPLIT_MIN_ROWS")) e->nsplit_min_rows = std::max(1, atoi(v));
  if (const char* h = getenv("BLIM_GEMM_HINTS")) e->gemm.l2_hints = atoi(h) != 0;
  if (const char* m = getenv("BLIM_GEMM_SB_MB")) { e->gemm.sb_mb = std::max(4, std::min(96, atoi(m))); e->gemm.sb_auto = false; }
  if (const char* m = getenv("BLIM_GEMM_SB_MIN")) { e->gemm.sb_min = std::max(1, std::min(8, atoi(m))); e->gemm.sb_auto = false; }
  auto bad = [&](const char* m) {
    g_create_error = m;
    delete e;
    return 1;
  };
  if (e->DH != 64 && e->DH != 128) return bad("head_dim must be 64 or 128");
  if (e->NKV <= 0 || e->NH % e->NKV) return bad("num_heads must be a multiple of num_kv_heads");
  if (e->G > 128) return bad("more than 128 query heads per KV head");
  if (e->H % 64 || e->NQ % 64 || e->I % 128 || e->MM % 64) return bad("hidden/intermediate/mm sizes must be multiples of 64/128/64");
  if (e->H % 8 || e->NQ != e->H) return bad("num_heads*head_dim must equal hidden_size");
  if (kBN % e->DH) return bad("head_dim must divide 256");
  if (cfg->max_positions <= 0) return bad("max_positions must be > 0");

  const size_t T = e->Tmax, P = e->Pmax;
  e->layers.resize(e->NL);
  struct { DevBuf* b; size_t bytes; } allocs[] = {
      {&e->x, T * e->H * 4}, {&e->xn, T * e->H * 2}, {&e->q, T * e->NQ * 2}, {&e->attn, T * e->NQ * 2},
      {&e->k_own, T * e->NKVD * 2}, {&e->v_own, T * e->NKVD * 2}, {&e->act, T * static_cast<size_t>(e->I) * 2},
      {&e->kp, static_cast<size_t>(e->NL) * P * e->NKVD * 2}, {&e->vp, static_cast<size_t>(e->NL) * P * e->NKVD * 2},
      {&e->prefix_last, static_cast<size_t>(e->Umax) * e->H * 4}, {&e->vis, P * e->H * 2}, {&e->vis_in, P * e->MM * 2}, {&e->d_vis_idx, P * 4}, {&e->proj_tmp, T * e->H * 2},
      {&e->lm_a, T * e->H * 2}, {&e->pred, T * e->MM * 2}, {&e->tgt_logit, T * 4}, {&e->logp, T * 4},
      {&e->d_tok_src, T * 4}, {&e->d_tok_pos, T * 4}, {&e->d_key_valid, T}, {&e->d_seqs, T * sizeof(AttnSeq)},
      {&e->d_works, (T * e->G / 64 + T + 1) * sizeof(AttnWorkTc)}, {&e->d_idx, T * 4}, {&e->d_targets, T * 4},
      {&e->d_row_off, (T + 1) * 4}, {&e->d_seq_start, T * 4}, {&e->ssq, T * 2 * static_cast<size_t>((e->H + kBN - 1) / kBN) * 4}, {&e->rstd, T * 4}, {&e->d_tok_slot, T * 4}};
  for (auto& a : allocs) {
    cudaError_t r = a.b->reserve(a.bytes);
    if (r != cudaSuccess) {
      g_create_error = std::string("workspace allocation failed: ") + cudaGetErrorString(r);
      blim_destroy(e);
      return 1;
    }
  }
  // K / V buffers are read in whole 64-row TMA boxes: rows outside a chunk must hold finite values
  cudaMemset(e->kp.p, 0, e->kp.cap); cudaMemset(e->vp.p, 0, e->vp.cap);
  cudaMemset(e->k_own.p, 0, e->k_own.cap); cudaMemset(e->v_own.p, 0, e->v_own.cap);
  if (!make_kv_tmap(&e->tm_kp, e->kp.p, static_cast<uint64_t>(e->NL) * P, e->NKVD) || !make_kv_tmap(&e->tm_vp, e->vp.p, static_cast<uint64_t>(e->NL) * P, e->NKVD) ||
      !make_kv_tmap(&e->tm_kown, e->k_own.p, T, e->NKVD) || !make_kv_tmap(&e->tm_vown, e->v_own.p, T, e->NKVD) ||
      !make_q_tmap(&e->tm_q, e->q.p, T, e->NH, e->DH, e->G, 128 / std::max(1, std::min(e->G, 128)))) {
    g_create_error = "cuTensorMapEncodeTiled failed for the K/V buffers";
    blim_destroy(e);
    return 1;
  }
  *out = e;
  return 0;
}

// ------------------------------------------------------------------------------------------------ weights
// GEMM weights go to the tensor-core operand format (act_t, see act_type.cuh: both operands of a kind::f16 MMA must share
// one format); T = bf16 keeps the checkpoint's format (the embedding table, which is gathered, never multiplied).
template <typename T>
static int repack_w(blim_engine* e, DevBuf& dst, size_t dst_rows, const void* src, int dtype, int rows, int cols, int dst_row0,
                    int interleave, int half, cudaStream_t st) {
  CKE(dst.reserve(dst_rows * cols * 2));
  const size_t n = static_cast<size_t>(rows) * cols;
  const int blocks = static_cast<int>(std::min<size_t>((n + 255) / 256, 4096));
  repack_rows_16_kernel<T><<<blocks, 256, 0, st>>>(dst.as<T>(), src, dtype, rows, cols, dst_row0, interleave, half);
  CKL();
  return 0;
}
// same for activation-format destinations (video features: the A operand of the projector)
static int repack_act(blim_engine* e, DevBuf& dst, size_t dst_rows, const void* src, int dtype, int rows, int cols, cudaStream_t st) {
  CKE(dst.reserve(dst_rows * cols * 2));
  const size_t n = static_cast<size_t>(rows) * cols;
  const int blocks = static_cast<int>(std::min<size_t>((n + 255) / 256, 4096));
  repack_rows_16_kernel<act_t><<<blocks, 256, 0, st>>>(dst.as<act_t>(), src, dtype, rows, cols, 0, 0, 0);
  CKL();
  return 0;
}
static int repack_f32(blim_engine* e, DevBuf& dst, size_t dst_elems, size_t dst_off, const void* src, int dtype, size_t n, int round_bf16,
                      cudaStream_t st) {
  CKE(dst.reserve(dst_elems * 4));
  const int blocks = static_cast<int>(std::min<size_t>((n + 255) / 256, 4096));
  repack_f32_kernel<<<blocks, 256, 0, st>>>(dst.as<float>() + dst_off, src, dtype, n, round_bf16);
  CKL();
  return 0;
}

extern "C" int blim_load_weight(blim_engine* e, const char* name_c, const void* src, int dtype, const int64_t* shape, int ndim, void* stream) {
  if (!e || !name_c || !src) return e ? e->fail("null argument") : 1;
  if (dtype < 0 || dtype > 2) return e->fail("bad dtype");
  CKE(cudaSetDevice(e->device));
  cudaStream_t st = S(stream);
  const std::string name(name_c);
  auto is2 = [&](int64_t r, int64_t c) { return ndim == 2 && shape[0] == r && shape[1] == c; };
  auto is1 = [&](int64_t n) { return ndim == 1 && shape[0] == n; };
  auto shape_err = [&]() { return e->fail("unexpected shape for " + name); };
  const int H = e->H, I = e->I, V = e->V, MM = e->MM, NQ = e->NQ, NKVD = e->NKVD, NQKV = e->NQKV;
  e->tvg_vis_ready = false;
  e->weights_loaded++;
  if (name == "model.embed_tokens.weight") {
    if (!is2(V, H)) return shape_err();
    return repack_w<bf16>(e, e->embed, V, src, dtype, V, H, 0, 0, 0, st);
  }
  if (name == "lm_head.weight") {
    if (!is2(V, H)) return shape_err();
    return repack_w<act_t>(e, e->lm_head, V, src, dtype, V, H, 0, 0, 0, st);
  }
  if (name == "visual_head.weight") {
    if (!is2(MM, H)) return shape_err();
    return repack_w<act_t>(e, e->visual_head, MM, src, dtype, MM, H, 0, 0, 0, st);
  }
  if (name == "model.norm.weight") {
    if (!is1(H)) return shape_err();
    return repack_f32(e, e->norm, H, 0, src, dtype, H, 1, st);
  }
  const std::string pj = "model.mm_projector.";
  if (name.compare(0, pj.size(), pj) == 0) {
    std::string rest = name.substr(pj.size());
    int which;
    if (rest.compare(0, 4, "mlp.") == 0) { which = 0; rest = rest.substr(4); }
    else if (rest.compare(0, 8, "tvg_mlp.") == 0) { which = 1; rest = rest.substr(8); }
    else { e->weights_loaded--; return 2; }
    ProjW& p = e->proj[which];
    if (rest == "0.weight") { if (!is2(H, MM)) return shape_err(); return repack_w<act_t>(e, p.w0, H, src, dtype, H, MM, 0, 0, 0, st); }
    if (rest == "0.bias") { if (!is1(H)) return shape_err(); return repack_f32(e, p.b0, H, 0, src, dtype, H, 1, st); }
    if (rest == "2.weight") { if (!is2(H, H)) return shape_err(); return repack_w<act_t>(e, p.w2, H, src, dtype, H, H, 0, 0, 0, st); }
    if (rest == "2.bias") { if (!is1(H)) return shape_err(); return repack_f32(e, p.b2, H, 0, src, dtype, H, 1, st); }
    e->weights_loaded--;
    return 2;
  }
  const std::string ly = "model.layers.";
  if (name.compare(0, ly.size(), ly) == 0) {
    size_t dot = name.find('.', ly.size());
    if (dot == std::string::npos) { e->weights_loaded--; return 2; }
    const int li = atoi(name.substr(ly.size(), dot - ly.size()).c_str());
    if (li < 0 || li >= e->NL) return e->fail("layer index out of range: " + name);
    LayerW& w = e->layers[li];
    const std::string rest = name.substr(dot + 1);
    if (w.folded && (rest.find("q_proj.weight") != std::string::npos || rest.find("k_proj.weight") != std::string::npos ||
                     rest.find("v_proj.weight") != std::string::npos || rest.find("gate_proj") != std::string::npos ||
                     rest.find("up_proj") != std::string::npos || rest.find("layernorm") != std::string::npos))
      return e->fail("layer " + std::to_string(li) + " is already finalised (norm weights folded into its GEMM weights): create a new engine to reload " + name);
    if (rest == "self_attn.q_proj.weight") { if (!is2(NQ, H)) return shape_err(); w.loaded |= kLdQ; return repack_w<act_t>(e, w.w_qkv, NQKV, src, dtype, NQ, H, 0, 0, 0, st); }
    if (rest == "self_attn.k_proj.weight") { if (!is2(NKVD, H)) return shape_err(); w.loaded |= kLdK; return repack_w<act_t>(e, w.w_qkv, NQKV, src, dtype, NKVD, H, NQ, 0, 0, st); }
    if (rest == "self_attn.v_proj.weight") { if (!is2(NKVD, H)) return shape_err(); w.loaded |= kLdV; return repack_w<act_t>(e, w.w_qkv, NQKV, src, dtype, NKVD, H, NQ + NKVD, 0, 0, st); }
    if (rest == "self_attn.q_proj.bias") { if (!is1(NQ)) return shape_err(); w.loaded |= kLdQb; return repack_f32(e, w.b_qkv, NQKV, 0, src, dtype, NQ, 1, st); }
    if (rest == "self_attn.k_proj.bias") { if (!is1(NKVD)) return shape_err(); w.loaded |= kLdKb; return repack_f32(e, w.b_qkv, NQKV, NQ, src, dtype, NKVD, 1, st); }
    if (rest == "self_attn.v_proj.bias") { if (!is1(NKVD)) return shape_err(); w.loaded |= kLdVb; return repack_f32(e, w.b_qkv, NQKV, NQ + NKVD, src, dtype, NKVD, 1, st); }
    if (rest == "self_attn.o_proj.weight") { if (!is2(H, NQ)) return shape_err(); w.loaded |= kLdO; return repack_w<act_t>(e, w.w_o, H, src, dtype, H, NQ, 0, 0, 0, st); }
    if (rest == "mlp.gate_proj.weight") { if (!is2(I, H)) return shape_err(); w.loaded |= kLdGate; return repack_w<act_t>(e, w.w_gu, 2 * static_cast<size_t>(I), src, dtype, I, H, 0, 128, 0, st); }
    if (rest == "mlp.up_proj.weight") { if (!is2(I, H)) return shape_err(); w.loaded |= kLdUp; return repack_w<act_t>(e, w.w_gu, 2 * static_cast<size_t>(I), src, dtype, I, H, 0, 128, 1, st); }
    if (rest == "mlp.down_proj.weight") { if (!is2(H, I)) return shape_err(); w.loaded |= kLdDown; return repack_w<act_t>(e, w.w_down, H, src, dtype, H, I, 0, 0, 0, st); }
    if (rest == "input_layernorm.weight") { if (!is1(H)) return shape_err(); w.loaded |= kLdLn1; return repack_f32(e, w.ln1, H, 0, src, dtype, H, 1, st); }
    if (rest == "post_attention_layernorm.weight") { if (!is1(H)) return shape_err(); w.loaded |= kLdLn2; return repack_f32(e, w.ln2, H, 0, src, dtype, H, 1, st); }
  }
  e->weights_loaded--;
  return 2;  // not a parameter of the scoring path (vision tower, rotary buffers, ...): ignored
}

extern "C" int blim_set_rope(blim_engine* e, const float* cos_dev, const float* sin_dev, int n_positions, void* stream) {
  if (!e || !cos_dev || !sin_dev || n_positions <= 0) return e ? e->fail("bad rope arguments") : 1;
  CKE(cudaSetDevice(e->device));
  const size_t bytes = static_cast<size_t>(n_positions) * (e->DH / 2) * 4;
  CKE(e->rope_cos.reserve(bytes));
  CKE(e->rope_sin.reserve(bytes));
  // stored transposed: [head_dim/2][n_positions] (see EpiQkvRope)
  const int n_el = n_positions * (e->DH / 2);
  transpose_f32_kernel<<<(n_el + 255) / 256, 256, 0, S(stream)>>>(e->rope_cos.as<float>(), cos_dev, n_positions, e->DH / 2);
  CKL();
  transpose_f32_kernel<<<(n_el + 255) / 256, 256, 0, S(stream)>>>(e->rope_sin.as<float>(), sin_dev, n_positions, e->DH / 2);
  CKL();
  e->rope_n = n_positions;
  return 0;
}

// ------------------------------------------------------------------------------------------------ corpus
static int upload(blim_engine* e, DevBuf& dst, const void* src, size_t bytes, cudaStream_t st);
extern "C" int blim_set_videos(blim_engine* e, const void* feats_dev, int dtype, int n_videos, int n_clips, void* stream) {
  if (!e || !feats_dev || n_videos <= 0 || n_clips <= 0) return e ? e->fail("bad video arguments") : 1;
  CKE(cudaSetDevice(e->device));
  const int rows = n_videos * n_clips * e->TPC;
  CKR(repack_act(e, e->feats, rows, feats_dev, dtype, rows, e->MM, S(stream)));
  e->n_videos = n_videos;
  e->n_clips = n_clips;
  e->tvg_vis_ready = false;
  return 0;
}

extern "C" int blim_set_texts(blim_engine* e, int which, const int32_t* ids, const int32_t* labels, const int64_t* off, int n_texts) {
  if (!e || which < 0 || which > 1 || !ids || !labels || !off || n_texts <= 0) return e ? e->fail("bad text arguments") : 1;
  TextTable& t = e->texts[which];
  t = TextTable();
  t.n = n_texts;
  t.off.assign(off, off + n_texts + 1);
  t.ids.assign(ids, ids + off[n_texts]);
  t.labels.assign(labels, labels + off[n_texts]);
  t.img_pos.resize(n_texts);
  t.first_lab.resize(n_texts);
  t.group.resize(n_texts);
  std::map<std::vector<int32_t>, int> groups;
  for (int i = 0; i < n_texts; ++i) {
    const int64_t b = off[i], n = off[i + 1] - off[i];
    int img = -1, n_img = 0, fl = static_cast<int>(n);
    for (int64_t j = 0; j < n; ++j) {
      if (ids[b + j] == -200) { if (img < 0) img = static_cast<int>(j); n_img++; }
      else if (ids[b + j] < 0 || ids[b + j] >= e->V) return e->fail("token id out of range in text " + std::to_string(i));
    }
    for (int64_t j = 0; j < n; ++j)
      if (labels[b + j] != -100) { fl = static_cast<int>(j); break; }
    if (n_img != 1) return e->fail("text " + std::to_string(i) + " must contain exactly one image sentinel (-200)");
    t.img_pos[i] = img;
    t.first_lab[i] = fl;
    if (which == BLIM_TEXTS_VTG) {
      // scored tail must be contiguous, lie after the image and hold valid targets (base_dataset.py:80-81)
      if (fl <= img) return e->fail("VTG text " + std::to_string(i) + ": labels start before the image token");
      if (fl >= n) return e->fail("VTG text " + std::to_string(i) + ": nothing to score");
      for (int64_t j = fl; j < n; ++j)
        if (labels[b + j] < 0 || labels[b + j] >= e->V) return e->fail("VTG text " + std::to_string(i) + ": non-contiguous labels");
    } else {
      if (img < 1) return e->fail("TVG text " + std::to_string(i) + ": no text before the image token");
    }
    // prompt group: VTG = everything before the first scored token; TVG = the CPN-visible header
    std::vector<int32_t> key;
    if (which == BLIM_TEXTS_VTG) key.assign(ids + b, ids + b + fl);
    auto it = groups.find(key);
    if (which == BLIM_TEXTS_VTG) {
      if (it == groups.end()) {
        PromptGroup g;
        g.pre.assign(ids + b, ids + b + img);
        g.post.assign(ids + b + img + 1, ids + b + fl);
        t.groups.push_back(g);
        it = groups.emplace(key, static_cast<int>(t.groups.size()) - 1).first;
      }
      t.group[i] = it->second;
    } else {
      t.group[i] = 0;  // TVG prior groups depend on tvg_prefix_len: resolved at scoring time
    }
  }
  return 0;
}

extern "C" int blim_set_video_vocab(blim_engine* e, const void* vocab_dev, int dtype, int n_vocab, const int32_t* labels, int n_videos,
                                    void* stream) {
  if (!e || !vocab_dev || n_vocab <= 0 || !labels || n_videos <= 0) return e ? e->fail("bad vocab arguments") : 1;
  if (e->n_clips <= 0) return e->fail("call blim_set_videos before blim_set_video_vocab");
  CKE(cudaSetDevice(e->device));
  const size_t n = static_cast<size_t>(n_vocab) * e->n_clips * e->MM;
  CKE(e->vocab.reserve(n * 2));
  const int blocks = static_cast<int>(std::min<size_t>((n + 255) / 256, 4096));
  repack_vocab_kernel<<<blocks, 256, 0, S(stream)>>>(e->vocab.as<act_t>(), vocab_dev, dtype, n_vocab, e->n_clips, e->MM);
  CKL();
  e->n_vocab = n_vocab;
  e->video_labels.assign(labels, labels + n_videos);
  for (int v = 0; v < n_videos; ++v)
    if (labels[v] < 0 || labels[v] >= n_vocab) return e->fail("video label out of range");
  return 0;
}

// video_vocab computed on the device from the features given to blim_set_videos (no host-side mean / upload).
extern "C" int blim_build_video_vocab(blim_engine* e, const int32_t* labels, int n_videos, int n_vocab, void* stream) {
  if (!e || !labels || n_vocab <= 0) return e ? e->fail("bad vocab arguments") : 1;
  if (e->n_videos <= 0 || n_videos != e->n_videos) return e->fail("blim_build_video_vocab: call blim_set_videos first (same number of videos)");
  for (int v = 0; v < n_videos; ++v)
    if (labels[v] < 0 || labels[v] >= n_vocab) return e->fail("video label out of range");
  CKE(cudaSetDevice(e->device));
  cudaStream_t st = S(stream);
  const size_t n = static_cast<size_t>(n_vocab) * e->n_clips * e->MM;
  CKE(e->vocab.reserve(n * 2));
  CKE(cudaMemsetAsync(e->vocab.p, 0, n * 2, st));
  CKR(upload(e, e->d_map, labels, static_cast<size_t>(n_videos) * sizeof(int), st));
  vocab_from_feats_kernel<<<n_videos * e->n_clips, 128, 0, st>>>(e->vocab.as<act_t>(), e->feats.as<act_t>(), e->d_map.as<int>(), e->n_clips, e->TPC,
                                                                 e->MM, n_vocab);
  CKL();
  e->n_vocab = n_vocab;
  e->video_labels.assign(labels, labels + n_videos);
  return 0;
}

extern "C" int blim_set_tvg_prefix_length(blim_engine* e, int n) {
  if (!e || n < 0) return e ? e->fail("bad tvg prefix length") : 1;
  e->tvg_prefix_len = n;
  return 0;
}

// ------------------------------------------------------------------------------------------------ building blocks
static int check_ready(blim_engine* e, cudaStream_t st) {
  if (!e->embed.p || !e->lm_head.p || !e->norm.p) return e->fail("weights not loaded (embed_tokens / lm_head / norm)");
  for (int l = 0; l < e->NL; ++l) {
    const LayerW& w = e->layers[l];
    if (w.loaded != kLdAll) {   // the fused QKV / gate|up buffers exist as soon as ONE of their parts is loaded: check every part
      std::string missing;
      for (int b = 0; b < 12; ++b)
        if (!(w.loaded & (1u << b))) missing += std::string(missing.empty() ? "" : ", ") + kLdNames[b];
      return e->fail("weights of layer " + std::to_string(l) + " not loaded: " + missing);
    }
  }
  if (!e->rope_cos.p) return e->fail("rotary table not set (blim_set_rope)");
  if (e->fuse_norm) {
    for (int l = 0; l < e->NL; ++l) {
      LayerW& w = e->layers[l];
      if (w.folded) continue;
      fold_norm_weight_kernel<<<2048, 256, 0, st>>>(w.w_qkv.as<act_t>(), w.ln1.as<float>(), static_cast<size_t>(e->NQKV), e->H);
      fold_norm_weight_kernel<<<2048, 256, 0, st>>>(w.w_gu.as<act_t>(), w.ln2.as<float>(), 2 * static_cast<size_t>(e->I), e->H);
      w.folded = true;
    }
  }
  return 0;
}

// Host -> device copy of scheduler metadata.  A cudaMemcpyAsync from pageable memory synchronises the stream before it
// starts, which would serialise the host-side planning of run k+1 with the GPU execution of run k; the data is therefore
// staged through a pinned bump arena (truly asynchronous copies).  The arena wraps with one stream synchronisation
// every ~64 MB of metadata (a run uploads well under 1 MB).
static int upload(blim_engine* e, DevBuf& dst, const void* src, size_t bytes, cudaStream_t st) {
  if (bytes == 0) return 0;
  CKE(dst.reserve(bytes));
  if (!e->arena && !e->arena_disabled) {
    const char* pe = getenv("BLIM_PINNED");
    if (pe && std::string(pe) == "0") e->arena_disabled = true;
  }
  if (!e->arena && !e->arena_disabled) {
    e->arena_cap = 64u << 20;
    if (cudaMallocHost(reinterpret_cast<void**>(&e->arena), e->arena_cap) != cudaSuccess) {
      e->arena = nullptr;
      e->arena_cap = 0;
      cudaGetLastError();
    }
  }
  const size_t need = (bytes + 255) & ~static_cast<size_t>(255);
  if (need > e->arena_cap) {
    CKE(cudaMemcpyAsync(dst.p, src, bytes, cudaMemcpyHostToDevice, st));  // pageable source: staged before return
    return 0;
  }
  if (e->arena_off + need > e->arena_cap) {
    CKE(cudaStreamSynchronize(st));  // every earlier staged copy has been consumed
    e->arena_off = 0;
  }
  memcpy(e->arena + e->arena_off, src, bytes);
  CKE(cudaMemcpyAsync(dst.p, e->arena + e->arena_off, bytes, cudaMemcpyHostToDevice, st));
  e->arena_off += need;
  return 0;
}

static int rmsnorm(blim_engine* e, act_t* out, const float* x0, const float* x1, const int* idx, const float* w, int R, cudaStream_t st) {
  if (R <= 0) return 0;
  e->tic(2, st);
  rmsnorm_kernel<<<R, 256, 0, st>>>(out, x0, x1, idx, w, R, e->H, e->cfg.rms_norm_eps);
  e->toc(st);
  CKL();
  return 0;
}

// Projector MLP over `rows` feature rows: Linear -> GELU -> Linear (mm_projector_builder.py:156-159).
static int project(blim_engine* e, const act_t* feats, int rows, int which, act_t* out, cudaStream_t st) {
  const ProjW& p = e->proj[which];
  if (!p.w0.p || !p.b0.p || !p.w2.p || !p.b2.p) return e->fail(which ? "tvg_mlp weights not loaded" : "mm_projector.mlp weights not loaded");
  for (int r0 = 0; r0 < rows; r0 += e->Tmax) {
    const int n = std::min(e->Tmax, rows - r0);
    EpiStore<act_t, true, true>::Params p1{e->proj_tmp.as<act_t>(), e->H, p.b0.as<float>()};
    CKR((gemm<EpiStore<act_t, true, true>>(e, feats + static_cast<size_t>(r0) * e->MM, e->MM, p.w0.as<act_t>(), e->MM, n, e->H, e->MM, p1, st)));
    EpiStore<act_t, true, false>::Params p2{out + static_cast<size_t>(r0) * e->H, e->H, p.b2.as<float>()};
    CKR((gemm<EpiStore<act_t, true, false>>(e, e->proj_tmp.as<act_t>(), e->H, p.w2.as<act_t>(), e->H, n, e->H, e->H, p2, st)));
  }
  return 0;
}

// Run the decoder over a flat token list.  x must already hold the input embeddings when `assembled` is true.
// `last_rows` (prefix runs): the only tokens whose final state is needed.  In the last layer everything after the
// attention is then computed for those rows only and their final residual-stream rows land in e->prefix_last[0..U);
// an empty list means the run only has to fill the KV cache (the last layer stops after its QKV projection).
static int run_decoder(blim_engine* e, Run& run, bool to_prefix_cache, bool assembled, cudaStream_t st,
                       const std::vector<int>* last_rows = nullptr) {
  const int T = run.T();
  if (T == 0) return 0;
  if (T > e->Tmax) return e->fail("internal: run exceeds max_run_tokens");
  if (to_prefix_cache && T > e->Pmax) return e->fail("internal: prefix run exceeds max_prefix_tokens");
  for (int p : run.tok_pos)
    if (p < 0 || p >= e->rope_n) return e->fail("sequence longer than the rotary table (max_positions)");
  std::vector<AttnWorkTc> works_tc;
  std::vector<int> seq_start;
  build_attn_works_tc(run.seqs.data(), static_cast<int>(run.seqs.size()), e->G, works_tc, seq_start, T);
  const int n_works = static_cast<int>(works_tc.size());
  CKR(upload(e, e->d_works, works_tc.data(), works_tc.size() * sizeof(AttnWorkTc), st));
  CKR(upload(e, e->d_seq_start, seq_start.data(), T * sizeof(int), st));
  CKR(upload(e, e->d_tok_pos, run.tok_pos.data(), T * sizeof(int), st));
  if (run.any_slot) {
    if (!to_prefix_cache) return e->fail("internal: K/V slots are only remapped in prefix runs");
    for (int sl : run.tok_slot)
      if (sl < 0 || sl >= e->Pmax) return e->fail("internal: K/V slot outside the prefix cache");
    CKR(upload(e, e->d_tok_slot, run.tok_slot.data(), T * sizeof(int), st));
  }
  const int* kv_slot = run.any_slot ? e->d_tok_slot.as<int>() : nullptr;
  if (run.any_invalid) CKR(upload(e, e->d_key_valid, run.key_valid.data(), T, st));
  if (!assembled) {
    CKR(upload(e, e->d_tok_src, run.tok_src.data(), T * sizeof(int), st));
    assemble_tokens_kernel<<<T, 128, 0, st>>>(e->x.as<float>(), e->embed.as<bf16>(), e->vis.as<act_t>(), e->d_tok_src.as<int>(), T, e->H);
    CKL();
  }
  const size_t kv_layer = static_cast<size_t>(e->Pmax) * e->NKVD;
  const bool fuse = e->fuse_norm;
  const int n_parts = 2 * ((e->H + kBN - 1) / kBN);
  const float* rstd = fuse ? e->rstd.as<float>() : nullptr;
  auto rowprep = [&](const float* x, int R) -> int {   // xn = bf16(x), rstd = 1/rms(x)
    rowprep_kernel<<<R, 256, 0, st>>>(e->xn.as<act_t>(), e->rstd.as<float>(), x, R, e->H, e->cfg.rms_norm_eps);
    CKL();
    return 0;
  };
  auto resid_gemm = [&](const act_t* A, int lda, const act_t* W, int K, float* x, int R, bool want_norm) -> int {
    if (want_norm) {   // x += A W^T, xn = bf16(x), rstd = 1/rms(x)  (first half of the next RMSNorm)
      EpiResidNorm::Params pn{x, e->xn.as<act_t>(), e->ssq.as<float>(), e->H};
      CKR(gemm<EpiResidNorm>(e, A, lda, W, lda, R, e->H, K, pn, st));
      rstd_rows_kernel<<<(R + 255) / 256, 256, 0, st>>>(e->rstd.as<float>(), e->ssq.as<float>(), R, n_parts, e->H, e->cfg.rms_norm_eps);
      CKL();
      return 0;
    }
    if (e->resid_tma && (e->H % 32) == 0 && (e->resid_tma == 2 || K == e->NQ)) {   // 1: o_proj only (short K), 2: down_proj too
      EpiResidTma::Params pt;
      if (!make_tmap_f32_32x32(&pt.tm_x, x, static_cast<uint64_t>(R), static_cast<uint64_t>(e->H), static_cast<uint64_t>(e->H)))
        return e->fail("cuTensorMapEncodeTiled failed for the residual stream");
      return gemm<EpiResidTma>(e, A, lda, W, lda, R, e->H, K, pt, st);
    }
    EpiResid::Params pr{x, e->H};
    // A/B knob (BLIM_GEMM_KSPLIT, default 1 = off): cut a long-K contraction (down_proj, 136 MB of weights) into K slices,
    // one launch each -- the epilogue accumulates into x anyway.  Measured: DRAM reads per launch at
    // M = 25 728 drop from 8.7 GB to 6.4 GB (2 slices), but the 68 MB slices still do not stay L2-resident and every slice
    // pays the fp32 read-modify-write of x again: -1 % pairs/s on the same box with 2 or 3 slices, so it stays off.
    int splits = 1;
    if (R >= 2048 && static_cast<size_t>(e->H) * K * 2 > e->ksplit_min_bytes) splits = e->ksplit;
    const int kb = K / kBK;
    for (int i = 0; i < splits; ++i) {
      const int k0 = (kb * i / splits) * kBK, k1 = (kb * (i + 1) / splits) * kBK;
      if (k1 > k0) CKR(gemm<EpiResid>(e, A + k0, lda, W + k0, lda, R, e->H, k1 - k0, pr, st));
    }
    return 0;
  };
  // gate|up: 271 MB of weights against a 126 MB L2 whose two halves each cache what their own SMs touch.  One launch
  // per N slice of <= 46 MB (6 at 7B): the slice stays L2-resident while the A panels stream through once per slice (m-tile
  // by m-tile, n fastest), instead of the whole weight streaming once per 40 MB super-block of A.  ncu at M = 25 728: DRAM
  // reads 5.8 GB -> 2.3 GB per layer, writes unchanged; same box: +1.1 % pairs/s (the saved DRAM power goes to the SM
  // clock: 1 275 -> 1 305 MHz under the cap).  BLIM_GEMM_NSPLIT overrides (1 = one launch).
  auto swiglu_gemm = [&](const act_t* A, const act_t* W, act_t* act, int R) -> int {
    const int n_tiles = (2 * e->I + kBN - 1) / kBN;
    const int want = e->nsplit > 0 ? e->nsplit : static_cast<int>((2 * static_cast<size_t>(e->I) * e->H * 2 + (46u << 20) - 1) / (46u << 20));
    const int splits = (R >= e->nsplit_min_rows) ? std::max(1, std::min(want, n_tiles)) : 1;
    for (int i = 0; i < splits; ++i) {
      const int t0 = n_tiles * i / splits, t1 = n_tiles * (i + 1) / splits;
      if (t1 <= t0) continue;
      const int n0 = t0 * kBN, n1 = std::min(t1 * kBN, 2 * e->I);
      EpiSwiglu::Params ps{act + static_cast<size_t>(t0) * (kBN / 2), e->I, rstd};
      CKR(gemm<EpiSwiglu>(e, A, e->H, W + static_cast<size_t>(n0) * e->H, e->H, R, n1 - n0, e->H, ps, st));
    }
    return 0;
  };
  if (fuse) CKR(rowprep(e->x.as<float>(), T));
  for (int l = 0; l < e->NL; ++l) {
    const LayerW& w = e->layers[l];
    act_t* kpl = e->kp.as<act_t>() + l * kv_layer;
    act_t* vpl = e->vp.as<act_t>() + l * kv_layer;
    act_t* k_out = to_prefix_cache ? kpl : e->k_own.as<act_t>();
    act_t* v_out = to_prefix_cache ? vpl : e->v_own.as<act_t>();
    if (!fuse) CKR(rmsnorm(e, e->xn.as<act_t>(), e->x.as<float>(), nullptr, nullptr, w.ln1.as<float>(), T, st));
    CKR(gemm_qkv(e, e->xn.as<act_t>(), w, T, e->q.as<act_t>(), k_out, v_out, e->d_tok_pos.as<int>(), kv_slot, rstd, st));
    const bool prune = last_rows != nullptr && l == e->NL - 1;
    if (prune && last_rows->empty()) break;
    const float scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(e->DH));
    cudaError_t r;
    e->tic(1, st);
    {
      AttnParamsTc ap;
      ap.q = e->q.as<act_t>(); ap.o = e->attn.as<act_t>();
      ap.a_row0 = l * e->Pmax;
      ap.b_row0 = to_prefix_cache ? l * e->Pmax : 0;
      AttnTcMaps maps;
      maps.ka = e->tm_kp; maps.va = e->tm_vp;
      maps.kb = to_prefix_cache ? e->tm_kp : e->tm_kown;
      maps.vb = to_prefix_cache ? e->tm_vp : e->tm_vown;
      ap.key_valid = run.any_invalid ? e->d_key_valid.as<uint8_t>() : nullptr;
      ap.tok_seq_start = e->d_seq_start.as<int>(); ap.works = e->d_works.as<AttnWorkTc>();
      ap.n_q = e->NQ; ap.n_kv = e->NKVD; ap.group = e->G; ap.scale_log2 = scale_log2; ap.q_stride = e->NQ; ap.n_works = n_works; ap.n_kv_heads = e->NKV;
      r = e->attn_version == kAttnWarpSpecialized ? launch_attention_ws<act_t>(e->tm_q, maps, ap, n_works, e->NKV, e->DH, st)
                                                   : launch_attention_tc<act_t>(maps, ap, n_works, e->NKV, e->DH, st, e->attn_version);
    }
    e->toc(st);
    if (r != cudaSuccess) return e->fail_cuda("attention launch", r);
    e->launches++;
    if (prune) {
      // last layer of a prefix run: o_proj + MLP only for the rows whose final state is read
      const int U = static_cast<int>(last_rows->size());
      if (U > e->Umax) return e->fail("internal: too many prefix units in one run");
      CKR(upload(e, e->d_idx, last_rows->data(), U * sizeof(int), st));
      gather_rows_16_kernel<<<U, 128, 0, st>>>(e->xn.as<act_t>(), e->attn.as<act_t>(), e->d_idx.as<int>(), U, e->NQ);
      CKL();
      gather_rows_f32_kernel<<<U, 256, 0, st>>>(e->prefix_last.as<float>(), e->x.as<float>(), e->d_idx.as<int>(), U, e->H);
      CKL();
      CKR(resid_gemm(e->xn.as<act_t>(), e->NQ, w.w_o.as<act_t>(), e->NQ, e->prefix_last.as<float>(), U, false));
      if (fuse) CKR(rowprep(e->prefix_last.as<float>(), U));
      else CKR(rmsnorm(e, e->xn.as<act_t>(), e->prefix_last.as<float>(), nullptr, nullptr, w.ln2.as<float>(), U, st));
      CKR(swiglu_gemm(e->xn.as<act_t>(), w.w_gu.as<act_t>(), e->act.as<act_t>(), U));
      CKR(resid_gemm(e->act.as<act_t>(), e->I, w.w_down.as<act_t>(), e->I, e->prefix_last.as<float>(), U, false));
      break;
    }
    CKR(resid_gemm(e->attn.as<act_t>(), e->NQ, w.w_o.as<act_t>(), e->NQ, e->x.as<float>(), T, fuse));
    if (!fuse) CKR(rmsnorm(e, e->xn.as<act_t>(), e->x.as<float>(), nullptr, nullptr, w.ln2.as<float>(), T, st));
    CKR(swiglu_gemm(e->xn.as<act_t>(), w.w_gu.as<act_t>(), e->act.as<act_t>(), T));
    CKR(resid_gemm(e->act.as<act_t>(), e->I, w.w_down.as<act_t>(), e->I, e->x.as<float>(), T, fuse && l + 1 < e->NL));
  }
  return 0;
}

// logp[r] = log softmax(scale * A[r] · W^T)[target[r]] for R rows, never materialising the logits.
static int lse_rows(blim_engine* e, const void* A, int lda, const void* W, int ldw, int R, int N, int K, const int* targets_dev, float scale,
                    float* logp_dev, cudaStream_t st, int b_fmt = kActFmt, int a_fmt = kActFmt) {
  if (R <= 0) return 0;
  const int n_tiles = (N + kBN - 1) / kBN;
  CKE(e->partial.reserve(static_cast<size_t>(e->Tmax) * 2 * n_tiles * sizeof(float2)));
  EpiLse::Params p{e->partial.as<float2>(), e->tgt_logit.as<float>(), targets_dev, scale};
  CKR(gemm<EpiLse>(e, A, lda, W, ldw, R, N, K, p, st, b_fmt, a_fmt));
  lse_finalize_kernel<<<(R + 7) / 8, 256, 0, st>>>(logp_dev, e->partial.as<float2>(), e->tgt_logit.as<float>(), R, 2 * n_tiles);
  CKL();
  return 0;
}

// ------------------------------------------------------------------------------------------------ batch planner
struct Item {
  int unit;     // owning unit
  int suf_len;  // suffix tokens this item runs through the decoder
  int key;      // unique-key index (result slot)
};
struct UnitPlan {
  int prefix_len;
  std::vector<int> items;  // indices into the item array
};
struct BatchUnit {
  int unit;
  int item_begin, item_end;  // range inside units[unit].items
};

// Pure host code (no engine, no device): blim_debug_plan_batches runs it on a described workload in the CPU tests.
struct PlanCaps {
  int Pmax, Tmax, Umax;  // prefix-cache rows, run tokens, prefix units per run
};
static const char* plan_batches_host(const PlanCaps* e, const std::vector<UnitPlan>& units, const std::vector<Item>& items, int max_items,
                                     std::vector<std::vector<BatchUnit>>& batches, int reserve_rows) {
  batches.clear();
  const int pcap_hard = std::min(e->Pmax, e->Tmax) - reserve_rows;  // a prefix run is also one decoder run; root rows come first
  if (pcap_hard <= 0) return "workspace too small for the shared prompt header";
  // balance: n batches of roughly equal size instead of (n-1) full ones and a small tail (small runs waste the GEMMs)
  long long tot_p = 0, tot_s = 0, tot_i = 0;
  int max_p = 0;
  long long max_s = 0;  // slack of the balanced suffix capacity: the largest piece that is placed whole (a unit, see below)
  for (const UnitPlan& up : units) {
    tot_p += up.prefix_len;
    max_p = std::max(max_p, up.prefix_len);
    long long unit_s = 0;
    for (int it : up.items) { unit_s += items[it].suf_len; max_s = std::max<long long>(max_s, items[it].suf_len); ++tot_i; }
    tot_s += unit_s;
    if (unit_s <= e->Tmax && static_cast<long long>(up.items.size()) <= max_items) max_s = std::max(max_s, unit_s);
  }
  for (const UnitPlan& up : units)
    if (up.prefix_len > e->Pmax || up.prefix_len > e->Tmax) return "a prefix sequence exceeds the workspace (raise max_prefix_tokens / max_run_tokens)";
  for (const Item& it : items)
    if (it.suf_len > e->Tmax) return "a suffix sequence exceeds max_run_tokens";
  // greedy placement in run order for a target of nb batches: capacities = the average + the largest piece placed whole,
  // so every batch but the last holds at least the average
  auto place = [&](long long nb) {
    batches.clear();
    std::vector<BatchUnit> cur;
    const int pcap = static_cast<int>(std::min<long long>(pcap_hard, (tot_p + nb - 1) / nb + max_p));
    const int scap = static_cast<int>(std::min<long long>(e->Tmax, (tot_s + nb - 1) / nb + max_s));
    long long cp = 0, cs = 0;
    int ci = 0;
    auto flush = [&]() {
      if (!cur.empty()) batches.push_back(cur);
      cur.clear(); cp = 0; cs = 0; ci = 0;
    };
    for (size_t u = 0; u < units.size(); ++u) {
      const UnitPlan& up = units[u];
      size_t pos = 0;
      // A unit's items stay in ONE batch whenever they fit an empty one: the attention tiles of a unit stack its sequences
      // in run order, so cutting a unit at a batch boundary that depends on what else is being scored would move its tile
      // (and 64-key chunk) boundaries -- the scores must not depend on the batch composition / the multi-GPU sharding.
      long long unit_s = 0;
      for (int it : up.items) unit_s += items[it].suf_len;
      const bool fits_alone = unit_s <= scap && static_cast<long long>(up.items.size()) <= max_items;
      if (fits_alone && !cur.empty() && (cs + unit_s > scap || ci + static_cast<long long>(up.items.size()) > max_items)) flush();
      while (pos < up.items.size()) {
        const int first_len = items[up.items[pos]].suf_len;   // <= scap: the slack covers the longest sequence
        if (cp + up.prefix_len > pcap || static_cast<int>(cur.size()) + 1 > e->Umax || cs + first_len > scap || ci + 1 > max_items) flush();
        BatchUnit bu{static_cast<int>(u), static_cast<int>(pos), static_cast<int>(pos)};
        cp += up.prefix_len;
        while (pos < up.items.size() && cs + items[up.items[pos]].suf_len <= scap && ci + 1 <= max_items) {
          cs += items[up.items[pos]].suf_len;
          ++ci; ++pos;
        }
        bu.item_end = static_cast<int>(pos);
        cur.push_back(bu);
      }
    }
    flush();
  };
  // (Re-planning for the number of batches this really took -- equal runs instead of full ones and a small tail -- was
  // tried: with two balanced capacities, prefix rows and suffix tokens, closing a batch on one leaves it short of the
  // average of the other, and the TVG runs of C2 went from 3 batches to 5.  A tail run costs ~5 ms of a 7 s job; left alone.)
  const long long nb = std::max<long long>(1, std::max((tot_p + pcap_hard - 1) / pcap_hard, std::max((tot_s + e->Tmax - 1) / e->Tmax, (tot_i + max_items - 1) / max_items)));
  place(nb);
  return nullptr;
}

static int plan_batches(blim_engine* e, const std::vector<UnitPlan>& units, const std::vector<Item>& items, int max_items,
                        std::vector<std::vector<BatchUnit>>& batches, int reserve_rows = 0) {
  const PlanCaps caps{e->Pmax, e->Tmax, e->Umax};
  const char* err = plan_batches_host(&caps, units, items, max_items, batches, reserve_rows);
  return err ? e->fail(err) : 0;
}

// ------------------------------------------------------------------------------------------------ TVG pooled visual rows
static int ensure_tvg_vis(blim_engine* e, cudaStream_t st) {
  if (e->tvg_vis_ready) return 0;
  if (e->n_videos <= 0) return e->fail("videos not set");
  const int n_rows = e->n_videos * e->n_clips;  // pooled rows
  CKE(e->tvg_vis.reserve(static_cast<size_t>(n_rows) * e->H * 2));
  // chunk over whole clips so every chunk fits the visual staging buffer
  const int clips_per_chunk = std::max(1, std::min(e->Pmax, e->Tmax) / e->TPC);
  for (int c0 = 0; c0 < n_rows; c0 += clips_per_chunk) {
    const int nc = std::min(clips_per_chunk, n_rows - c0);
    CKR(project(e, e->feats.as<act_t>() + static_cast<size_t>(c0) * e->TPC * e->MM, nc * e->TPC, 1, e->vis.as<act_t>(), st));
    mean_rows_kernel<<<nc, 256, 0, st>>>(e->tvg_vis.as<act_t>() + static_cast<size_t>(c0) * e->H, e->vis.as<act_t>(), nc, e->TPC, e->H);
    CKL();
  }
  e->tvg_vis_ready = true;
  return 0;
}

// ------------------------------------------------------------------------------------------------ shared prompt root
// The first R tokens of every prefix of a batch are often the same chat-template header at the same positions; being
// causal, their K/V rows are identical for all units.  They are prefilled ONCE into cache rows [0, R) (a run that only
// fills the KV cache) and copied in front of every unit's own rows, so a unit's cached prefix stays one contiguous
// segment [root copy | unit tokens] for the cascade attention of the suffix run.
static int prefill_root(blim_engine* e, const int32_t* ids, int R, const std::vector<int>& unit_bases, cudaStream_t st) {
  Run rr;
  rr.begin_seq(0, 0);
  for (int j = 0; j < R; ++j) rr.push(ids[j], j);
  rr.end_seq();
  const std::vector<int> none;
  CKR(run_decoder(e, rr, true, false, st, &none));
  const int U = static_cast<int>(unit_bases.size());
  if (U == 0) return 0;
  CKR(upload(e, e->d_idx, unit_bases.data(), U * sizeof(int), st));
  dim3 grid(static_cast<unsigned>(U), static_cast<unsigned>(e->NL), 2);
  replicate_root_rows_kernel<<<grid, 128, 0, st>>>(e->kp.as<act_t>(), e->vp.as<act_t>(), e->d_idx.as<int>(), R, e->NKVD,
                                                   static_cast<size_t>(e->Pmax) * e->NKVD);
  CKL();
  return 0;
}

// ------------------------------------------------------------------------------------------------ scoring: VTG / VTG prior
static int score_vtg(blim_engine* e, bool prior, const std::vector<std::pair<int, int>>& keys /* (v, t) unique */, float* out_unique,
                     cudaStream_t st) {
  const TextTable& tt = e->texts[BLIM_TEXTS_VTG];
  const int n_vis = e->n_clips * e->TPC;
  // units: (video, prompt group) for the likelihood, (prompt group) for the prior
  std::map<std::pair<int, int>, int> unit_of;
  std::vector<UnitPlan> units;
  std::vector<std::pair<int, int>> unit_key;
  std::vector<Item> items(keys.size());
  for (size_t i = 0; i < keys.size(); ++i) {
    const int v = keys[i].first, t = keys[i].second;
    const int g = tt.group[t];
    const std::pair<int, int> uk(prior ? -1 : v, g);
    auto it = unit_of.find(uk);
    if (it == unit_of.end()) {
      UnitPlan up;
      up.prefix_len = static_cast<int>(tt.groups[g].pre.size() + tt.groups[g].post.size()) + (prior ? 0 : n_vis);
      if (up.prefix_len == 0) return e->fail("empty VTG prompt");
      units.push_back(up);
      unit_key.push_back(uk);
      it = unit_of.emplace(uk, static_cast<int>(units.size()) - 1).first;
    }
    const int n_tok = static_cast<int>(tt.off[t + 1] - tt.off[t]);
    items[i].unit = it->second;
    items[i].suf_len = n_tok - tt.first_lab[t] - 1;  // caption tokens except the last one (its state predicts nothing)
    items[i].key = static_cast<int>(i);
    units[it->second].items.push_back(static_cast<int>(i));
  }
  // root sharing (likelihood only: the prior's prefix is a single 26-token sequence anyway); needs the kernels that
  // understand remapped K/V rows
  const bool root_enabled = !prior && e->root_share;
  int max_root = 0;
  if (root_enabled)
    for (const PromptGroup& g : tt.groups) max_root = std::max(max_root, static_cast<int>(g.pre.size()));
  std::vector<std::vector<BatchUnit>> batches;
  CKR(plan_batches(e, units, items, e->Tmax / 2, batches, max_root));

  Run run;
  for (const auto& batch : batches) {
    // ---- projector for the distinct videos of the batch
    std::map<int, int> vis_row0;  // video -> first row in e->vis
    if (!prior) {
      int rows = 0;
      for (const BatchUnit& bu : batch) {
        const int v = unit_key[bu.unit].first;
        if (!vis_row0.count(v)) { vis_row0[v] = rows; rows += n_vis; }
      }
      // one projector pass per batch: a run of consecutive video ids is projected in place, anything else (e.g. the
      // strided ids a rank owns in a multi-GPU run) is first gathered into a contiguous staging buffer
      std::vector<std::pair<int, int>> by_row;   // (first row in e->vis, video)
      for (const auto& kv : vis_row0) by_row.emplace_back(kv.second, kv.first);
      std::sort(by_row.begin(), by_row.end());
      bool contiguous = true;
      for (size_t i = 1; i < by_row.size(); ++i) contiguous = contiguous && by_row[i].second == by_row[i - 1].second + 1;
      if (contiguous) {
        CKR(project(e, e->feats.as<act_t>() + static_cast<size_t>(by_row[0].second) * n_vis * e->MM, rows, 0, e->vis.as<act_t>(), st));
      } else {
        std::vector<int> idx(rows);
        for (const auto& rv : by_row)
          for (int i = 0; i < n_vis; ++i) idx[rv.first + i] = rv.second * n_vis + i;
        CKR(upload(e, e->d_vis_idx, idx.data(), static_cast<size_t>(rows) * 4, st));
        gather_rows_16_kernel<<<rows, 128, 0, st>>>(e->vis_in.as<act_t>(), e->feats.as<act_t>(), e->d_vis_idx.as<int>(), rows, e->MM);
        CKL();
        CKR(project(e, e->vis_in.as<act_t>(), rows, 0, e->vis.as<act_t>(), st));
      }
    }
    // ---- prefix run.  With a shared root (all units of the batch use the same prompt group) the header tokens are
    //      prefilled once and only [visual rows | prompt tail] run per video.
    int R = 0;
    if (root_enabled) {
      const int g0 = unit_key[batch[0].unit].second;
      bool same = true;
      for (const BatchUnit& bu : batch) same = same && unit_key[bu.unit].second == g0;
      if (same) R = static_cast<int>(tt.groups[g0].pre.size());
    }
    std::vector<int> unit_start(batch.size()), unit_len(batch.size()), last_tok(batch.size());
    if (R > 0) {
      int base = R;
      for (size_t b = 0; b < batch.size(); ++b) {
        const PromptGroup& g = tt.groups[unit_key[batch[b].unit].second];
        unit_start[b] = base;
        unit_len[b] = R + n_vis + static_cast<int>(g.post.size());
        base += unit_len[b];
      }
      CKR(prefill_root(e, tt.groups[unit_key[batch[0].unit].second].pre.data(), R, unit_start, st));
    }
    run.clear();
    for (size_t b = 0; b < batch.size(); ++b) {
      run.unit = static_cast<int>(b);
      const PromptGroup& g = tt.groups[unit_key[batch[b].unit].second];
      int pos = 0;
      if (R > 0) {
        run.begin_seq(unit_start[b], R, unit_start[b] + R);   // attends to its root copy, writes K/V right behind it
        pos = R;
      } else {
        unit_start[b] = run.begin_seq(0, 0);
        for (int32_t id : g.pre) run.push(id, pos++);
      }
      if (!prior) {
        const int r0 = vis_row0[unit_key[batch[b].unit].first];
        for (int j = 0; j < n_vis; ++j) run.push(-1 - (r0 + j), pos++);
      } else {
        pos += n_vis;  // CPN: visual rows are invisible as keys but keep their rotary positions (modeling_videochat_flash.py:433)
      }
      for (int32_t id : g.post) run.push(id, pos++);
      run.end_seq();
      if (R == 0) unit_len[b] = run.T() - unit_start[b];
      last_tok[b] = run.T() - 1;
    }
    CKR(run_decoder(e, run, true, false, st, &last_tok));
    // ---- suffix run
    run.clear();
    std::vector<int> row_idx, targets, row_off, item_keys;
    row_off.push_back(0);
    for (size_t b = 0; b < batch.size(); ++b) {
      run.unit = static_cast<int>(b);
      const UnitPlan& up = units[batch[b].unit];
      const PromptGroup& g = tt.groups[unit_key[batch[b].unit].second];
      const int pos0 = static_cast<int>(g.pre.size() + g.post.size()) + n_vis;
      for (int ii = batch[b].item_begin; ii < batch[b].item_end; ++ii) {
        const Item& it = items[up.items[ii]];
        const int t = keys[it.key].second;
        const int64_t base = tt.off[t];
        const int fl = tt.first_lab[t];
        if (prior) run.unit = (1 << 24) + ii;   // prior: every text is its own tile group (the one unit holds all texts)
        const int q0 = run.begin_seq(unit_start[b], unit_len[b]);
        for (int j = 0; j < it.suf_len; ++j) run.push(tt.ids[base + fl + j], pos0 + j);
        run.end_seq();
        // LM rows: last prefix state predicts the first label, suffix token j predicts label fl + 1 + j
        row_idx.push_back(-1 - static_cast<int>(b));
        targets.push_back(tt.labels[base + fl]);
        for (int j = 0; j < it.suf_len; ++j) {
          row_idx.push_back(q0 + j);
          targets.push_back(tt.labels[base + fl + 1 + j]);
        }
        row_off.push_back(static_cast<int>(row_idx.size()));
        item_keys.push_back(it.key);
      }
    }
    CKR(run_decoder(e, run, false, false, st));
    // ---- fused LM head over whole items, chunks of <= Tmax rows
    const int n_items = static_cast<int>(item_keys.size());
    int i0 = 0;
    while (i0 < n_items) {
      int i1 = i0;
      while (i1 < n_items && row_off[i1 + 1] - row_off[i0] <= e->Tmax) ++i1;
      if (i1 == i0) return e->fail("a caption has more scored tokens than max_run_tokens");
      const int r0 = row_off[i0], R = row_off[i1] - r0;
      CKR(upload(e, e->d_idx, row_idx.data() + r0, R * sizeof(int), st));
      CKR(upload(e, e->d_targets, targets.data() + r0, R * sizeof(int), st));
      std::vector<int> off_local(i1 - i0 + 1);
      for (int i = i0; i <= i1; ++i) off_local[i - i0] = row_off[i] - r0;
      CKR(upload(e, e->d_row_off, off_local.data(), off_local.size() * sizeof(int), st));
      CKR(rmsnorm(e, e->lm_a.as<act_t>(), e->x.as<float>(), e->prefix_last.as<float>(), e->d_idx.as<int>(), e->norm.as<float>(), R, st));
      CKR(lse_rows(e, e->lm_a.as<act_t>(), e->H, e->lm_head.as<act_t>(), e->H, R, e->V, e->H, e->d_targets.as<int>(), 1.0f, e->logp.as<float>(), st));
      // item scores land in a staging area (reuse tgt_logit after finalize is done with it: stream ordered)
      const int P = i1 - i0;
      vtg_seq_mean_kernel<<<(P + 7) / 8, 256, 0, st>>>(e->tgt_logit.as<float>(), e->logp.as<float>(), e->d_row_off.as<int>(), P);
      CKL();
      // scatter to the unique-key slots
      CKR(upload(e, e->d_idx, item_keys.data() + i0, P * sizeof(int), st));
      {
        // out_unique[key[i]] = staged[i]  (a scatter with n_cols = 1 row = key)
        std::vector<int> zeros(P, 0);
        CKR(upload(e, e->d_targets, zeros.data(), P * sizeof(int), st));
        scatter_scores_kernel<<<(P + 255) / 256, 256, 0, st>>>(out_unique, 1, e->d_idx.as<int>(), e->d_targets.as<int>(), e->tgt_logit.as<float>(), P);
        CKL();
      }
      i0 = i1;
    }
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------ scoring: TVG / TVG prior
// keys: likelihood -> (v, t); prior -> (v, t_representative) with one key per distinct (prefix group, T0, v).
static int score_tvg(blim_engine* e, bool prior, const std::vector<std::pair<int, int>>& keys, float* out_unique, cudaStream_t st) {
  const TextTable& tt = e->texts[BLIM_TEXTS_TVG];
  if (!e->visual_head.p) return e->fail("visual_head weight not loaded");
  if (!e->vocab.p) return e->fail("video vocab not set");
  CKR(ensure_tvg_vis(e, st));
  const int NC = e->n_clips;
  // units: likelihood -> one per text (prefix = its T0 text tokens); prior -> one per distinct visible header
  std::map<std::vector<int32_t>, int> hdr_units;
  std::map<int, int> text_units;
  std::vector<UnitPlan> units;
  std::vector<int> unit_text;  // a text whose ids define the unit's prefix
  std::vector<int> unit_plen;
  std::vector<Item> items(keys.size());
  for (size_t i = 0; i < keys.size(); ++i) {
    const int t = keys[i].second;
    const int T0 = tt.img_pos[t];
    int u;
    if (!prior) {
      auto it = text_units.find(t);
      if (it == text_units.end()) {
        UnitPlan up; up.prefix_len = T0;
        units.push_back(up); unit_text.push_back(t); unit_plen.push_back(T0);
        it = text_units.emplace(t, static_cast<int>(units.size()) - 1).first;
      }
      u = it->second;
      items[i].suf_len = NC - 1;
    } else {
      const int plen = std::min(e->tvg_prefix_len, T0 - 1);  // visible header; the last text token is handled in the suffix
      std::vector<int32_t> hk(tt.ids.begin() + tt.off[t], tt.ids.begin() + tt.off[t] + plen);
      auto it = hdr_units.find(hk);
      if (it == hdr_units.end()) {
        UnitPlan up; up.prefix_len = plen;
        units.push_back(up); unit_text.push_back(t); unit_plen.push_back(plen);
        it = hdr_units.emplace(hk, static_cast<int>(units.size()) - 1).first;
      }
      u = it->second;
      items[i].suf_len = NC;  // [last text token] + NC-1 pooled visual rows
    }
    items[i].unit = u;
    items[i].key = static_cast<int>(i);
    units[u].items.push_back(static_cast<int>(i));
  }
  const bool root_enabled = !prior && e->root_share && e->tvg_prefix_len > 0;
  std::vector<std::vector<BatchUnit>> batches;
  CKR(plan_batches(e, units, items, e->Tmax / std::max(1, NC), batches, root_enabled ? e->tvg_prefix_len : 0));

  Run run;
  const float scale = 1.0f / sqrtf(static_cast<float>(e->MM));
  for (const auto& batch : batches) {
    // ---- prefix run (text tokens only, embeddings).  Likelihood: the first tvg_prefix_length tokens (system / user
    //      header + instruction) are the same for every text -> shared root, prefilled once per batch.
    int R = 0;
    if (root_enabled) {
      R = e->tvg_prefix_len;
      const int t0 = unit_text[batch[0].unit];
      for (const BatchUnit& bu : batch) {
        const int t = unit_text[bu.unit];
        if (unit_plen[bu.unit] <= R || memcmp(&tt.ids[tt.off[t]], &tt.ids[tt.off[t0]], R * sizeof(int32_t)) != 0) { R = 0; break; }
      }
    }
    std::vector<int> unit_start(batch.size()), unit_len(batch.size()), last_tok(batch.size());
    if (R > 0) {
      int base = R;
      for (size_t b = 0; b < batch.size(); ++b) {
        unit_start[b] = base;
        unit_len[b] = unit_plen[batch[b].unit];
        base += unit_len[b];
      }
      CKR(prefill_root(e, &tt.ids[tt.off[unit_text[batch[0].unit]]], R, unit_start, st));
    }
    run.clear();
    bool any_prefix = false;
    for (size_t b = 0; b < batch.size(); ++b) {
      run.unit = static_cast<int>(b);
      const int t = unit_text[batch[b].unit];
      const int plen = unit_plen[batch[b].unit];
      if (R > 0) {
        run.begin_seq(unit_start[b], R, unit_start[b] + R);
        for (int j = R; j < plen; ++j) run.push(tt.ids[tt.off[t] + j], j);
      } else {
        unit_start[b] = run.begin_seq(0, 0);
        for (int j = 0; j < plen; ++j) run.push(tt.ids[tt.off[t] + j], j);
        unit_len[b] = plen;
      }
      run.end_seq();
      last_tok[b] = run.T() - 1;
      any_prefix = any_prefix || plen > 0;
    }
    if (any_prefix) {
      const std::vector<int> none;
      CKR(run_decoder(e, run, true, false, st, prior ? &none : &last_tok));
    }
    // ---- suffix run.  Visual rows come straight from the pooled tvg_mlp table: tok_src = -1 - row, with e->vis
    //      temporarily replaced by the table (assemble reads `visual` rows by index).
    run.clear();
    std::vector<int> item_keys, item_q0, item_unit_b, item_row0;
    for (size_t b = 0; b < batch.size(); ++b) {
      run.unit = static_cast<int>(b);
      const UnitPlan& up = units[batch[b].unit];
      // TVG prior: the state of the last text token (position T0-1, CPN-masked as a key) only depends on the visible
      // header, T0 and the token id, not on the video: one 1-token sequence per distinct (T0, token) of this unit
      std::map<std::pair<int, int>, int> lt_tok;
      if (prior) {
        for (int ii = batch[b].item_begin; ii < batch[b].item_end; ++ii) {
          const int t = keys[items[up.items[ii]].key].second;
          const int T0 = tt.img_pos[t];
          if ((T0 - 1) < e->tvg_prefix_len) continue;  // visible as a key: stays inside the item's own sequence
          const std::pair<int, int> k(T0, tt.ids[tt.off[t] + T0 - 1]);
          if (lt_tok.count(k)) continue;
          run.unit = -1 - static_cast<int>(lt_tok.size());   // its own tile
          lt_tok[k] = run.begin_seq(unit_start[b], unit_len[b]);
          run.push(k.second, T0 - 1, false);
          run.end_seq();
        }
      }
      for (int ii = batch[b].item_begin; ii < batch[b].item_end; ++ii) {
        const Item& it = items[up.items[ii]];
        const int v = keys[it.key].first, t = keys[it.key].second;
        const int T0 = tt.img_pos[t];
        run.unit = prior ? (1 << 24) + v : static_cast<int>(b);   // prior: tiles stack the sequences of one video only
        int q0 = run.begin_seq(unit_start[b], unit_len[b]);
        int row0 = -1;
        if (prior) {
          if ((T0 - 1) < e->tvg_prefix_len) {
            row0 = q0;
            run.push(tt.ids[tt.off[t] + T0 - 1], T0 - 1, true);
            q0 += 1;
          } else {
            row0 = lt_tok[std::pair<int, int>(T0, tt.ids[tt.off[t] + T0 - 1])];
          }
        }
        for (int c = 0; c < NC - 1; ++c) run.push(-1 - (v * NC + c), T0 + c);
        run.end_seq();
        item_keys.push_back(it.key);
        item_q0.push_back(q0);       // first pooled visual row of the item
        item_row0.push_back(row0);   // prior: token whose state predicts clip 0
        item_unit_b.push_back(static_cast<int>(b));
      }
    }
    if (run.T() > 0) {
      // assemble with the pooled table as the visual source
      const int T = run.T();
      CKR(upload(e, e->d_tok_src, run.tok_src.data(), T * sizeof(int), st));
      assemble_tokens_kernel<<<T, 128, 0, st>>>(e->x.as<float>(), e->embed.as<bf16>(), e->tvg_vis.as<act_t>(), e->d_tok_src.as<int>(), T, e->H);
      CKL();
      CKR(run_decoder(e, run, false, true, st));
    }
    // ---- TVG head, chunks of items with NC * P <= Tmax rows, rows laid out clip-major [c][p]
    const int n_items = static_cast<int>(item_keys.size());
    const int max_p = std::max(1, e->Tmax / NC);
    for (int i0 = 0; i0 < n_items; i0 += max_p) {
      const int P = std::min(max_p, n_items - i0);
      std::vector<int> row_idx(static_cast<size_t>(NC) * P), targets(P);
      for (int p = 0; p < P; ++p) {
        const int i = i0 + p;
        for (int c = 0; c < NC; ++c) {
          int src;
          if (!prior) src = (c == 0) ? -1 - item_unit_b[i] : item_q0[i] + (c - 1);
          else src = (c == 0) ? item_row0[i] : item_q0[i] + (c - 1);
          row_idx[static_cast<size_t>(c) * P + p] = src;
        }
        targets[p] = e->video_labels[keys[item_keys[i]].first];
      }
      const int R = NC * P;
      CKR(upload(e, e->d_idx, row_idx.data(), R * sizeof(int), st));
      CKR(upload(e, e->d_targets, targets.data(), P * sizeof(int), st));
      CKR(rmsnorm(e, e->lm_a.as<act_t>(), e->x.as<float>(), e->prefix_last.as<float>(), e->d_idx.as<int>(), e->norm.as<float>(), R, st));
      EpiStore<act_t, false, false>::Params pv{e->pred.as<act_t>(), e->MM, nullptr};
      CKR((gemm<EpiStore<act_t, false, false>>(e, e->lm_a.as<act_t>(), e->H, e->visual_head.as<act_t>(), e->H, R, e->MM, e->H, pv, st)));
      for (int c = 0; c < NC; ++c) {
        CKR(lse_rows(e, e->pred.as<act_t>() + static_cast<size_t>(c) * P * e->MM, e->MM, e->vocab.as<act_t>() + static_cast<size_t>(c) * e->n_vocab * e->MM,
                     e->MM, P, e->n_vocab, e->MM, e->d_targets.as<int>(), scale, e->logp.as<float>() + static_cast<size_t>(c) * P, st));
      }
      tvg_clip_mean_kernel<<<(P + 255) / 256, 256, 0, st>>>(e->tgt_logit.as<float>(), e->logp.as<float>(), P, NC);
      CKL();
      CKR(upload(e, e->d_idx, item_keys.data() + i0, P * sizeof(int), st));
      std::vector<int> zeros(P, 0);
      CKR(upload(e, e->d_row_off, zeros.data(), P * sizeof(int), st));
      scatter_scores_kernel<<<(P + 255) / 256, 256, 0, st>>>(out_unique, 1, e->d_idx.as<int>(), e->d_row_off.as<int>(), e->tgt_logit.as<float>(), P);
      CKL();
    }
  }
  return 0;
}

extern "C" int blim_score_pairs(blim_engine* e, int kind, const int32_t* pair_v, const int32_t* pair_t, int64_t n_pairs, float* out_dev,
                                void* stream) {
  if (!e) return 1;
  if (n_pairs == 0) return 0;
  if (!pair_v || !pair_t || !out_dev || n_pairs < 0 || n_pairs > (1ll << 30)) return e->fail("bad pair arguments");
  if (kind < 0 || kind > 3) return e->fail("bad score kind");
  CKE(cudaSetDevice(e->device));
  cudaStream_t st = S(stream);
  CKR(check_ready(e, st));
  const bool is_tvg = kind == BLIM_TVG || kind == BLIM_TVG_PRIOR;
  const bool prior = kind == BLIM_VTG_PRIOR || kind == BLIM_TVG_PRIOR;
  const TextTable& tt = e->texts[is_tvg ? BLIM_TEXTS_TVG : BLIM_TEXTS_VTG];
  if (tt.n == 0) return e->fail(is_tvg ? "TVG texts not set" : "VTG texts not set");
  if (e->n_videos == 0) return e->fail("videos not set");
  if (!is_tvg && !prior && (!e->proj[0].w0.p)) return e->fail("mm_projector.mlp weights not loaded");
  // ---- unique keys (packed into 64 bits, sort + unique)
  //   VTG: (v,t)   VTG prior: (t) [all videos share n_clips]   TVG: (v,t)
  //   TVG prior: (visible-header group, T0, last text token, v): texts with the same visible header, the same length and
  //   the same last text token give the same prior for a video
  std::vector<int> hdr_group;  // TVG prior: header group of every text
  if (kind == BLIM_TVG_PRIOR) {
    std::map<std::vector<int32_t>, int> groups;
    hdr_group.resize(tt.n);
    for (int t = 0; t < tt.n; ++t) {
      const int plen = std::min(e->tvg_prefix_len, tt.img_pos[t] - 1);
      std::vector<int32_t> h(tt.ids.begin() + tt.off[t], tt.ids.begin() + tt.off[t] + plen);
      auto it = groups.find(h);
      if (it == groups.end()) it = groups.emplace(h, static_cast<int>(groups.size())).first;
      hdr_group[t] = it->second;
    }
    if (groups.size() > 1024) return e->fail("too many distinct TVG prompt headers");
  }
  auto make_key = [&](int v, int t) -> uint64_t {
    if (kind == BLIM_VTG || kind == BLIM_TVG) return (static_cast<uint64_t>(v) << 32) | static_cast<uint32_t>(t);
    if (kind == BLIM_VTG_PRIOR) return static_cast<uint32_t>(t);
    const uint64_t T0 = static_cast<uint64_t>(tt.img_pos[t]) & 0x3FFF;              // < 2^14: positions are bounded by the rotary table
    const uint64_t tok = static_cast<uint64_t>(tt.ids[tt.off[t] + tt.img_pos[t] - 1]);  // < 2^20 checked below
    // video outermost inside a header group: a video's keys are consecutive in the run, so its attention tiles only stack
    // its own sequences and its scores do not depend on which other videos are scored alongside (multi-GPU sharding)
    return (static_cast<uint64_t>(hdr_group[t]) << 54) | (static_cast<uint64_t>(v) << 34) | (T0 << 20) | tok;
  };
  if (kind == BLIM_TVG_PRIOR && (e->V > (1 << 20) || e->n_videos > (1 << 20) || e->rope_n > (1 << 14))) return e->fail("TVG prior key does not fit 64 bits");
  std::vector<std::pair<uint64_t, int>> tagged(n_pairs);
  for (int64_t i = 0; i < n_pairs; ++i) {
    const int v = pair_v[i], t = pair_t[i];
    if (v < 0 || v >= e->n_videos || t < 0 || t >= tt.n) return e->fail("pair index out of range");
    tagged[i] = std::make_pair(make_key(v, t), static_cast<int>(i));
  }
  std::sort(tagged.begin(), tagged.end());
  std::vector<std::pair<int, int>> keys;   // representative (v, t) of every unique key, in sorted key order
  std::vector<int> map(n_pairs);
  for (int64_t i = 0; i < n_pairs; ++i) {
    if (i == 0 || tagged[i].first != tagged[i - 1].first) keys.emplace_back(pair_v[tagged[i].second], pair_t[tagged[i].second]);
    map[tagged[i].second] = static_cast<int>(keys.size()) - 1;
  }
  CKE(e->uniq_scores.reserve(keys.size() * 4));
  if (is_tvg) CKR(score_tvg(e, prior, keys, e->uniq_scores.as<float>(), st));
  else CKR(score_vtg(e, prior, keys, e->uniq_scores.as<float>(), st));
  CKR(upload(e, e->d_map, map.data(), map.size() * sizeof(int), st));
  const int n = static_cast<int>(n_pairs);
  expand_scores_kernel<<<(n + 255) / 256, 256, 0, st>>>(out_dev, e->uniq_scores.as<float>(), e->d_map.as<int>(), n);
  CKL();
  return 0;
}

// dst[i] = Dst(src[i]) for 16-bit formats (the compat entry points speak bf16, the engine's activations are act_t)
template <typename Dst, typename Src>
__global__ void convert16_kernel(Dst* __restrict__ dst, const Src* __restrict__ src, size_t n) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
    dst[i] = Fmt16<Dst>::from_float(Fmt16<Src>::to_float(src[i]));
}
template <typename Dst, typename Src>
static int convert16(blim_engine* e, Dst* dst, const Src* src, size_t n, cudaStream_t st) {
  if (n == 0) return 0;
  convert16_kernel<Dst, Src><<<static_cast<unsigned>(std::min<size_t>((n + 255) / 256, 4096)), 256, 0, st>>>(dst, src, n);
  CKL();
  return 0;
}

// ------------------------------------------------------------------------------------------------ compat forward
extern "C" int blim_forward_logits(blim_engine* e, const void* embeds, const int32_t* mask_dev, int B, int L, float* logits, void* hidden,
                                   void* stream) {
  if (!e) return 1;
  if (!embeds || B <= 0 || L <= 0) return e->fail("bad forward arguments");
  CKE(cudaSetDevice(e->device));
  cudaStream_t st = S(stream);
  CKR(check_ready(e, st));
  if (L > e->Tmax) return e->fail("sequence longer than max_run_tokens");
  const int seq_per_run = std::max(1, e->Tmax / L);
  std::vector<int32_t> mask_h;
  if (mask_dev) {
    mask_h.resize(static_cast<size_t>(B) * L);
    CKE(cudaMemcpyAsync(mask_h.data(), mask_dev, mask_h.size() * 4, cudaMemcpyDeviceToHost, st));
    CKE(cudaStreamSynchronize(st));
  }
  Run run;
  for (int b0 = 0; b0 < B; b0 += seq_per_run) {
    const int nb = std::min(seq_per_run, B - b0);
    run.clear();
    for (int b = 0; b < nb; ++b) {
      run.begin_seq(0, 0);
      for (int j = 0; j < L; ++j) run.push(0, j, mask_dev ? mask_h[static_cast<size_t>(b0 + b) * L + j] != 0 : true);
      run.end_seq();
    }
    const int T = nb * L;
    const size_t n = static_cast<size_t>(T) * e->H;
    bf16_rows_to_f32_kernel<<<static_cast<unsigned>((n / 2 + 255) / 256), 256, 0, st>>>(
        e->x.as<float>(), reinterpret_cast<const bf16*>(embeds) + static_cast<size_t>(b0) * L * e->H, n);
    CKL();
    CKR(run_decoder(e, run, false, true, st));
    CKR(rmsnorm(e, e->lm_a.as<act_t>(), e->x.as<float>(), nullptr, nullptr, e->norm.as<float>(), T, st));
    if (hidden) CKR(convert16(e, reinterpret_cast<bf16*>(hidden) + static_cast<size_t>(b0) * L * e->H, e->lm_a.as<act_t>(), n, st));
    if (logits) {
      EpiStore<float, false, false>::Params p{logits + static_cast<size_t>(b0) * L * e->V, e->V, nullptr};
      CKR((gemm<EpiStore<float, false, false>>(e, e->lm_a.as<act_t>(), e->H, e->lm_head.as<act_t>(), e->H, T, e->V, e->H, p, st)));
    }
  }
  return 0;
}

extern "C" int blim_project_video(blim_engine* e, const void* feats, int n_rows, int tvg, void* out, void* stream) {
  if (!e) return 1;
  if (!feats || !out || n_rows < 0) return e->fail("bad projector arguments");
  CKE(cudaSetDevice(e->device));
  cudaStream_t st = S(stream);
  CKE(e->io_in.reserve(static_cast<size_t>(n_rows) * e->MM * 2));
  CKE(e->io_out.reserve(static_cast<size_t>(n_rows) * e->H * 2));
  CKR(convert16(e, e->io_in.as<act_t>(), reinterpret_cast<const bf16*>(feats), static_cast<size_t>(n_rows) * e->MM, st));
  CKR(project(e, e->io_in.as<act_t>(), n_rows, tvg ? 1 : 0, e->io_out.as<act_t>(), st));
  return convert16(e, reinterpret_cast<bf16*>(out), e->io_out.as<act_t>(), static_cast<size_t>(n_rows) * e->H, st);
}

extern "C" int blim_forward_visual(blim_engine* e, const void* hidden, int n_rows, void* out, void* stream) {
  if (!e) return 1;
  if (!hidden || !out || n_rows < 0) return e->fail("bad forward_visual arguments");
  if (!e->visual_head.p) return e->fail("visual_head weight not loaded");
  CKE(cudaSetDevice(e->device));
  cudaStream_t st = S(stream);
  CKE(e->io_in.reserve(static_cast<size_t>(n_rows) * e->H * 2));
  CKE(e->io_out.reserve(static_cast<size_t>(n_rows) * e->MM * 2));
  CKR(convert16(e, e->io_in.as<act_t>(), reinterpret_cast<const bf16*>(hidden), static_cast<size_t>(n_rows) * e->H, st));
  EpiStore<act_t, false, false>::Params p{e->io_out.as<act_t>(), e->MM, nullptr};
  CKR((gemm<EpiStore<act_t, false, false>>(e, e->io_in.as<act_t>(), e->H, e->visual_head.as<act_t>(), e->H, n_rows, e->MM, e->H, p, st)));
  return convert16(e, reinterpret_cast<bf16*>(out), e->io_out.as<act_t>(), static_cast<size_t>(n_rows) * e->MM, st);
}

__global__ void embed_rows_kernel(bf16* __restrict__ out, const bf16* __restrict__ embed, const int* __restrict__ ids, int n, int H) {
  const int t = blockIdx.x;
  if (t >= n) return;
  const bf16* row = embed + static_cast<size_t>(ids[t]) * H;
  for (int c = threadIdx.x * 8; c < H; c += blockDim.x * 8)
    *reinterpret_cast<uint4*>(out + static_cast<size_t>(t) * H + c) = *reinterpret_cast<const uint4*>(row + c);
}
extern "C" int blim_embed_tokens(blim_engine* e, const int32_t* ids_dev, int n, void* out, void* stream) {
  if (!e) return 1;
  if (!ids_dev || !out || n < 0) return e->fail("bad embed arguments");
  if (!e->embed.p) return e->fail("embed_tokens weight not loaded");
  if (n == 0) return 0;
  CKE(cudaSetDevice(e->device));
  embed_rows_kernel<<<n, 128, 0, S(stream)>>>(reinterpret_cast<bf16*>(out), e->embed.as<bf16>(), ids_dev, n, e->H);
  CKL();
  return 0;
}

// ------------------------------------------------------------------------------------------------ fuse / rerank / scatter
extern "C" int blim_fuse_rerank(blim_engine* e, const blim_fuse_cfg* c, const int32_t* cand_idx, const float* cand, const float* prior,
                                const float* query, const float* iv2, int n_rows, int n_cols, int k, int row0, double* fused_out,
                                int32_t* order_out, int32_t* gt_rank_out, int32_t* zero_count, void* stream) {
  if (!e) return 1;
  if (!c || !cand_idx || !iv2 || !fused_out || !order_out || !gt_rank_out || !zero_count || n_rows < 0 || n_cols <= 0 || k <= 0 || k > n_cols)
    return e->fail("bad fuse_rerank arguments");
  if (row0 < 0 || row0 + n_rows > n_cols) return e->fail("fuse_rerank: ground-truth column out of range (rows must be a slice of a square problem)");
  if (n_rows == 0) return 0;
  CKE(cudaSetDevice(e->device));
  FuseCoef f;
  f.alpha = static_cast<float>(c->alpha);
  f.c_q = static_cast<float>(c->c_query);
  f.c_q_om = static_cast<float>(1.0 - c->c_query);
  f.c_e = static_cast<float>(c->c_ens);
  f.c_e_om = static_cast<float>(1.0 - c->c_ens);
  f.c_e_d = c->c_ens;
  f.use_prior = c->use_prior && prior != nullptr;
  f.use_query = c->use_query;
  f.cpn_zero_f64 = c->cpn_zero_f64;
  const size_t smem = static_cast<size_t>(k) * (sizeof(double) + sizeof(int)) + n_cols;
  if (smem > 200 * 1024) return e->fail("fuse_rerank: row too wide for shared memory");
  if (!e->attr_fuse) {   // per engine = per device (function attributes are per device)
    CKE(cudaFuncSetAttribute(fuse_rerank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    e->attr_fuse = true;
  }
  fuse_rerank_kernel<<<n_rows, 256, smem, S(stream)>>>(f, cand_idx, cand, prior, query, iv2, n_rows, n_cols, k, row0, fused_out, order_out,
                                                        gt_rank_out, zero_count);
  CKL();
  return 0;
}

extern "C" int blim_rank_dense(blim_engine* e, const float* mat, int n_rows, int n_cols, int row0, int32_t* gt_rank_out, int32_t* zero_count,
                               void* stream) {
  if (!e) return 1;
  if (!mat || !gt_rank_out || !zero_count || n_rows < 0 || n_cols <= 0 || row0 < 0 || row0 + n_rows > n_cols) return e->fail("bad rank_dense arguments");
  if (n_rows == 0) return 0;
  CKE(cudaSetDevice(e->device));
  rank_dense_kernel<<<n_rows, 256, 0, S(stream)>>>(mat, n_rows, n_cols, row0, gt_rank_out, zero_count);
  CKL();
  return 0;
}

extern "C" int blim_topk_rows(blim_engine* e, const float* mat, int n_rows, int n_cols, int k, int32_t* idx_out, float* val_out, void* stream) {
  if (!e) return 1;
  if (!mat || !idx_out || !val_out || n_rows < 0 || n_cols <= 0 || k <= 0 || k > n_cols) return e->fail("bad topk arguments");
  if (n_rows == 0) return 0;
  CKE(cudaSetDevice(e->device));
  const int warps = 4;
  const size_t smem = static_cast<size_t>(warps) * (n_cols * sizeof(float) + ((n_cols + 31) / 32) * sizeof(unsigned));
  if (smem > 200 * 1024) return e->fail("topk: row too wide for shared memory");
  if (!e->attr_topk) {
    CKE(cudaFuncSetAttribute(topk_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    e->attr_topk = true;
  }
  topk_rows_kernel<<<(n_rows + warps - 1) / warps, warps * 32, smem, S(stream)>>>(mat, n_rows, n_cols, k, idx_out, val_out);
  CKL();
  return 0;
}

extern "C" int blim_scatter_scores(blim_engine* e, float* dense, int n_rows, int n_cols, int do_fill, float fill, const int32_t* row,
                                   const int32_t* col, const float* val, int64_t n, void* stream) {
  if (!e) return 1;
  if (!dense || n_rows <= 0 || n_cols <= 0 || n < 0 || (n > 0 && (!row || !col || !val))) return e->fail("bad scatter arguments");
  CKE(cudaSetDevice(e->device));
  if (do_fill) {
    const size_t tot = static_cast<size_t>(n_rows) * n_cols;
    fill_f32_kernel<<<static_cast<unsigned>(std::min<size_t>((tot + 255) / 256, 4096)), 256, 0, S(stream)>>>(dense, fill, tot);
    CKL();
  }
  if (n > 0) {
    scatter_scores_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, S(stream)>>>(dense, n_cols, row, col, val, static_cast<int>(n));
    CKL();
  }
  return 0;
}

extern "C" int blim_profile(blim_engine* e, int enable) {
  if (!e) return 1;
  e->profiling = enable != 0;
  return 0;
}
// Synchronises the device, sums the recorded event intervals per category (0 = tcgen05 GEMM, 1 = attention) and resets.
extern "C" int blim_profile_read(blim_engine* e, double* gemm_ms, double* attn_ms, int64_t* gemm_launches, int64_t* attn_launches) {
  if (!e) return 1;
  CKE(cudaSetDevice(e->device));
  CKE(cudaDeviceSynchronize());
  double ms[3] = {0.0, 0.0, 0.0};
  int64_t n[3] = {0, 0, 0};
  for (auto& t : e->timed) {
    float f = 0.f;
    if (cudaEventElapsedTime(&f, t.a, t.b) == cudaSuccess) { ms[t.cat] += f; n[t.cat]++; }
    e->event_pool.push_back(t.a);
    e->event_pool.push_back(t.b);
  }
  e->timed.clear();
  if (gemm_ms) *gemm_ms = ms[0];
  if (attn_ms) *attn_ms = ms[1];
  if (gemm_launches) *gemm_launches = n[0];
  if (attn_launches) *attn_launches = n[1];
  return 0;
}

// Per-kind breakdown of the same event intervals WITHOUT resetting them (call before blim_profile_read): index
// 0 QKV+RoPE, 1 o_proj, 2 gate|up+SwiGLU, 3 down_proj, 4 LM/TVG head + log-sum-exp, 5 other GEMMs (projectors, visual head),
// 6 attention, 7 RMSNorm; ms = summed device time, flops = executed 2*M*N*K of the GEMM launches (0 for 6, 7).
extern "C" int blim_profile_read_detail(blim_engine* e, int n, double* ms, double* flops, int64_t* launches) {
  if (!e) return 1;
  if (n < kProfCats || !ms || !flops || !launches) return e->fail("blim_profile_read_detail: need room for 8 categories");
  CKE(cudaSetDevice(e->device));
  CKE(cudaDeviceSynchronize());
  for (int i = 0; i < n; ++i) { ms[i] = 0.0; flops[i] = 0.0; launches[i] = 0; }
  for (auto& t : e->timed) {
    float f = 0.f;
    if (cudaEventElapsedTime(&f, t.a, t.b) != cudaSuccess) continue;
    const int c = t.cat == 0 ? t.sub : (t.cat == 1 ? kProfAttn : kProfNorm);
    ms[c] += f; flops[c] += t.flops; launches[c]++;
  }
  return 0;
}

// 1 = bf16, 2 = fp16 (same codes as blim_load_weight's dtype): the 16-bit format of the engine's activation operands
extern "C" int blim_act_dtype(void) { return kActFmt == kFmtF16 ? 2 : 1; }
extern "C" int64_t blim_kernel_launches(const blim_engine* e) { return e ? e->launches + e->gemm.launches : 0; }
extern "C" double blim_gemm_flops(const blim_engine* e) { return e ? e->flops : 0.0; }

// ------------------------------------------------------------------------------------------------ multi-GPU exchange
extern "C" int blim_comm_unique_id(void* id_out128) {
  if (!id_out128) { g_create_error = "null argument"; return 1; }
  NcclApi& api = nccl_api();
  if (!api.load()) { g_create_error = api.error; return 1; }
  NcclUniqueId id;
  const int rc = api.GetUniqueId(&id);
  if (rc != kNcclSuccess) { g_create_error = "ncclGetUniqueId: " + api.describe(rc); return 1; }
  memcpy(id_out128, id.internal, sizeof(id.internal));
  return 0;
}

extern "C" int blim_comm_init(blim_engine* e, const void* id128, int rank, int world) {
  if (!e) return 1;
  if (!id128 || world <= 0 || rank < 0 || rank >= world) return e->fail("bad communicator arguments");
  if (e->comm) return e->fail("communicator already initialised");
  NcclApi& api = nccl_api();
  if (!api.load()) return e->fail(api.error);
  CKE(cudaSetDevice(e->device));
  NcclUniqueId id;
  memcpy(id.internal, id128, sizeof(id.internal));
  const int rc = api.CommInitRank(&e->comm, world, id, rank);
  if (rc != kNcclSuccess) { e->comm = nullptr; return e->fail("ncclCommInitRank: " + api.describe(rc)); }
  e->comm_rank = rank;
  e->comm_world = world;
  return 0;
}

extern "C" int blim_comm_destroy(blim_engine* e) {
  if (!e) return 1;
  if (e->comm) {
    CKE(cudaSetDevice(e->device));
    nccl_api().CommDestroy(e->comm);
    e->comm = nullptr;
    e->comm_world = 1;
    e->comm_rank = 0;
  }
  return 0;
}

// recv[r * count .. (r + 1) * count) = rank r's send[0 .. count) for every rank r, fp32, enqueued on `stream` behind the
// scoring kernels that produce `send` (no host synchronisation).  nccl_comm: an ncclComm_t owned by the caller, or
// nullptr for the engine's own communicator (blim_comm_init).
extern "C" int blim_allgather_scores(blim_engine* e, void* nccl_comm, const float* send_dev, float* recv_dev, int64_t count_per_rank, void* stream) {
  if (!e) return 1;
  if (!send_dev || !recv_dev || count_per_rank < 0) return e->fail("bad all-gather arguments");
  if (count_per_rank == 0) return 0;
  NcclApi& api = nccl_api();
  if (!api.load()) return e->fail(api.error);
  NcclComm comm = nccl_comm ? reinterpret_cast<NcclComm>(nccl_comm) : e->comm;
  if (!comm) return e->fail("no communicator: call blim_comm_init or pass an ncclComm_t");
  CKE(cudaSetDevice(e->device));
  const int rc = api.AllGather(send_dev, recv_dev, static_cast<size_t>(count_per_rank), kNcclFloat32, comm, S(stream));
  if (rc != kNcclSuccess) return e->fail("ncclAllGather: " + api.describe(rc));
  return 0;
}

// ------------------------------------------------------------------------------------------------ debug GEMM entry
extern "C" int blim_debug_gemm(blim_engine* e, int epilogue, const void* A, const void* W, void* C, int M, int N, int K, const float* bias,
                               const int32_t* target, float scale, int cta_group, void* stream) {
  if (!e) return 1;
  if (!A || !W || !C || M <= 0 || N <= 0 || K <= 0) return e->fail("bad debug_gemm arguments");
  CKE(cudaSetDevice(e->device));
  cudaStream_t st = S(stream);
  const int saved = e->gemm.cta_group;
  if (cta_group == 1 || cta_group == 2) e->gemm.cta_group = cta_group;
  const act_t* a = reinterpret_cast<const act_t*>(A);   // operand format (blim_act_dtype), like every GEMM operand of the path
  const act_t* w = reinterpret_cast<const act_t*>(W);
  int r = 0;
  switch (epilogue) {
    case 0: { EpiStore<act_t, false, false>::Params p{reinterpret_cast<act_t*>(C), N, nullptr}; r = gemm<EpiStore<act_t, false, false>>(e, a, K, w, K, M, N, K, p, st); break; }
    case 1: { EpiStore<act_t, true, false>::Params p{reinterpret_cast<act_t*>(C), N, bias}; r = gemm<EpiStore<act_t, true, false>>(e, a, K, w, K, M, N, K, p, st); break; }
    case 2: { EpiStore<act_t, true, true>::Params p{reinterpret_cast<act_t*>(C), N, bias}; r = gemm<EpiStore<act_t, true, true>>(e, a, K, w, K, M, N, K, p, st); break; }
    case 3: { EpiStore<float, false, false>::Params p{reinterpret_cast<float*>(C), N, nullptr}; r = gemm<EpiStore<float, false, false>>(e, a, K, w, K, M, N, K, p, st); break; }
    case 4: { EpiResid::Params p{reinterpret_cast<float*>(C), N}; r = gemm<EpiResid>(e, a, K, w, K, M, N, K, p, st); break; }
    case 7: { EpiResidT<1>::Params p{reinterpret_cast<float*>(C), N}; r = gemm<EpiResidT<1>>(e, a, K, w, K, M, N, K, p, st); break; }
    case 5: { EpiSwiglu::Params p{reinterpret_cast<act_t*>(C), N / 2, nullptr}; r = gemm<EpiSwiglu>(e, a, K, w, K, M, N, K, p, st); break; }
    case 6: {
      if (M > e->Tmax) { r = e->fail("debug_gemm lse: M exceeds max_run_tokens"); break; }
      r = lse_rows(e, a, K, w, K, M, N, K, target, scale, reinterpret_cast<float*>(C), st);
      break;
    }
    default: r = e->fail("unknown epilogue");
  }
  e->gemm.cta_group = saved;
  return r;
}

// Host-only view of the batch planner (no engine, no device): which batch each suffix sequence of a described workload
// lands in.  The CPU tests pin the planner's contract with it -- capacities respected, and a unit's sequences never cut
// at a batch boundary unless the unit alone exceeds a run (what keeps the scores independent of batching / sharding).
extern "C" int blim_debug_plan_batches(int max_prefix_tokens, int max_run_tokens, int max_units, int max_items, int reserve_rows,
                                       const int32_t* unit_prefix_len, const int32_t* unit_item_count, int n_units,
                                       const int32_t* item_suf_len, int32_t* batch_of_item_out) {
  if (!unit_prefix_len || !unit_item_count || !item_suf_len || !batch_of_item_out || n_units < 0 || max_items <= 0 || max_run_tokens <= 0) return -1;
  std::vector<UnitPlan> units(static_cast<size_t>(n_units));
  std::vector<Item> items;
  for (int u = 0; u < n_units; ++u) {
    units[u].prefix_len = unit_prefix_len[u];
    for (int j = 0; j < unit_item_count[u]; ++j) {
      units[u].items.push_back(static_cast<int>(items.size()));
      items.push_back(Item{u, item_suf_len[items.size()], static_cast<int>(items.size())});
    }
  }
  const PlanCaps caps{max_prefix_tokens, max_run_tokens, max_units};
  std::vector<std::vector<BatchUnit>> batches;
  if (plan_batches_host(&caps, units, items, max_items, batches, reserve_rows)) return -1;
  for (size_t b = 0; b < batches.size(); ++b)
    for (const BatchUnit& bu : batches[b])
      for (int j = bu.item_begin; j < bu.item_end; ++j) batch_of_item_out[units[bu.unit].items[j]] = static_cast<int32_t>(b);
  return static_cast<int>(batches.size());
}

extern "C" int blim_debug_umma(blim_engine* e, const void* A, const void* B, float* C, int K, int N, int b_mn_major, uint32_t lbo, uint32_t sbo,
                               uint32_t kstep_bytes, void* stream) {
  if (!e) return 1;
  if (!A || !B || !C) return e->fail("bad debug_umma arguments");
  CKE(cudaSetDevice(e->device));
  cudaError_t r = launch_umma_probe(A, B, C, K, N, b_mn_major, lbo, sbo, kstep_bytes, S(stream));
  if (r != cudaSuccess) return e->fail_cuda("umma probe", r);
  e->launches++;
  return 0;
}

// ------------------------------------------------------------------------------------------------ feature extractor
// include/blim_vision.h: UMT ViT encoder + ToMe token merging on the same GEMM / attention kernels (one translation unit)
#include "vision.cuh"
