// tcgen05 / TMEM attention: varlen causal GQA with a shared (cascade) prefix segment (sequence description: attention.cuh;
// reference semantics: Qwen2SdpaAttention, modeling_qwen2_flash.py:685-709).  This header holds the kernels in which one
// group of threads executes the whole per-chunk chain -- used by the feature extractor (attention_tc4_kernel at head_dim
// 64, attention_tc2_kernel otherwise) and as the A/B fallback of the scoring path (BLIM_ATTN=tc2p|tc2) -- plus the work
// list builder and the tensor-map / launch helpers.  The scoring path's default is the warp-specialised kernel in
// attention_ws.cuh, which shares the work items, masks and numerics described here.
//
// One CTA = up to 128 stacked query rows (row = token * G + head-in-group of `128 / G` consecutive tokens of a
// prefix-sharing group of sequences) x one KV head.
//   S  = Q K^T      tcgen05.mma  M=128, N=16..64 (keys of the chunk), K=head_dim          -> TMEM
//   softmax         each thread reads ITS row of S from TMEM (tcgen05.ld 32x32b): running max / sum / rescale are
//                   thread-local; P (16-bit operand format) goes to a 128-byte-swizzled K-major tile in shared memory
//   O += P V        tcgen05.mma  M=128, N=head_dim, K=keys; V stays [key][head_dim] in memory = MN-major B operand
//                   (descriptor semantics pinned by tests/test_umma_probe_gpu.py); O lives in TMEM
// Keys come in 64-key chunks: first the a_len prefix keys (all visible), then ONE contiguous range of own-run keys
// [first token of the first sequence in the block, last token of the block]: key kt is visible to the query at token rt
// iff seq_start(rt) <= kt <= rt and key_valid[kt] -- sequences are contiguous in the run, so "same sequence AND causal"
// is an interval test.  K / V chunks are staged by TMA (64 keys x 64 columns boxes, SWIZZLE_128B, mbarrier completion).
// Rows of a box beyond the chunk's keys hold other (finite) cache rows and meet P = 0.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "act_type.cuh"
#include "attention.cuh"
#include "gemm_sm100.cuh"
#include "ptx_sm100.cuh"
#include "umma_probe.cuh"

namespace blim {

struct AttnWorkTc {
  int tok0;     // first token (run index) of the block
  int n_tok;    // tokens in the block (<= 128 / G)
  int a_start;  // prefix key rows [a_start, a_start + a_len) in k_a / v_a
  int a_len;
  int kb0;      // first own key needed, as a run token index (= first token of the sequence that contains tok0)
  int b_off;    // own key row of token t = t + b_off (non-zero when a prefix run writes its K/V behind replicated root rows)
  int pad1, pad2;
};
struct AttnParamsTc {
  const void* q;              // 16-bit elements in the kernel's operand format T16 (act_type.cuh)
  void* o;
  int a_row0;                 // row offset of this layer inside the tensor map of the prefix K / V buffers
  int b_row0;                 // row offset inside the tensor map of the own-run K / V buffers
  const uint8_t* key_valid;   // per own token, nullptr = all valid
  const int* tok_seq_start;   // [T] first token of the sequence each token belongs to (nullptr: no own-run keys at all)
  const AttnWorkTc* works;
  int n_q, n_kv, group;
  float scale_log2;
  int q_stride;               // elements between the Q rows of consecutive tokens (n_q, or 3 * n_q for a packed [Q|K|V] buffer)
  int n_works, n_kv_heads;    // filled by launch_attention_tc (the persistent kernel walks items = works x KV heads)
};

constexpr int kTcKeys = 64;       // keys per chunk
constexpr int kTcTmemCols = 256;  // S (64) + Oc (<= 128), power of two

// ------------------------------------------------------------------------------------------------ v2: O stays in TMEM
//   * 256 threads: two threads per query row (warps w and w+4 share a TMEM lane quadrant); each handles 32 of the chunk's
//     64 keys and half of the head_dim columns, the row maximum is exchanged through shared memory;
//   * O is accumulated by the tensor core in TMEM (Oc = P V with accumulate); it is only touched by threads when the
//     running reference maximum has to grow by more than 2^8 ("lazy rescale": tcgen05.ld -> scale -> tcgen05.st);
//   * S = Q K^T of chunk c+1 is issued right behind Oc = P V of chunk c, so it runs under the softmax of nobody and is
//     ready when the threads come back; V is double-buffered, K single-buffered, both by TMA.
constexpr int kTc2Threads = 256;

template <int DH>
constexpr int attn_tc2_smem_bytes() {
  // Q (DH/64 x 16 KB) + K chunk (DH/64 x 8 KB) + 2 x V chunk (DH/64 x 8 KB) + P (16 KB) + alignment slack
  return (DH / 64) * 16384 + 3 * (DH / 64) * 8192 + 16384 + 1024;
}

template <int DH, typename T16>
__global__ void __launch_bounds__(kTc2Threads, 2)
attention_tc2_kernel(const __grid_constant__ CUtensorMap tm_ka, const __grid_constant__ CUtensorMap tm_va,
                     const __grid_constant__ CUtensorMap tm_kb, const __grid_constant__ CUtensorMap tm_vb, const AttnParamsTc p) {
  constexpr int kSub = DH / 64;
  constexpr uint32_t kChunkBytes = kSub * 8192;
  constexpr float kGrow = 8.0f;  // lazy rescale threshold (log2 units): P stays below 2^8
  extern __shared__ uint8_t smem_raw_tc[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_tc) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* s_q = smem;
  uint8_t* s_k = s_q + kSub * 16384;
  uint8_t* s_v = s_k + kSub * 8192;          // two stages
  uint8_t* s_p = s_v + 2 * kSub * 8192;
  __shared__ uint64_t bar_s, bar_o, bar_k, bar_v[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_x[2][128];              // per-row exchange between the two threads of a row (max, then sum)

  const AttnWorkTc w = p.works[blockIdx.x];
  const int kvh = blockIdx.y;
  const int G = p.group;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int r = tid & 127, half = tid >> 7;
  // Warp 0 issues every TMA / tcgen05.mma with uniform control flow (all 32 lanes walk the issue code, so addresses and
  // descriptors are computed once in uniform registers) and only the issuing instruction is predicated on the elected lane:
  // the issue work sits on the CTA's critical path (everyone waits for warp 0 at the barriers).
  const bool lead = warp == 0 ? elect_one() != 0 : false;

  if (tid == 0) {
    mbar_init(&bar_s, 1);
    mbar_init(&bar_o, 1);
    mbar_init(&bar_k, 1);
    mbar_init(&bar_v[0], 1);
    mbar_init(&bar_v[1], 1);
    fence_barrier_init();
    tma_prefetch_desc(&tm_ka);
    tma_prefetch_desc(&tm_va);
    tma_prefetch_desc(&tm_kb);
    tma_prefetch_desc(&tm_vb);
  }
  if (warp == 0) tmem_alloc<1>(&tmem_slot, kTcTmemCols);

  const int tok_local = r / G, head = r - tok_local * G;
  const bool row_ok = tok_local < w.n_tok;
  const int rt = w.tok0 + (row_ok ? tok_local : 0);
  const int seq_lo = (row_ok && p.tok_seq_start != nullptr) ? __ldg(p.tok_seq_start + rt) : 0;
  {
    const T16* src = reinterpret_cast<const T16*>(p.q) + static_cast<size_t>(rt) * p.q_stride + (kvh * G + head) * DH;
#pragma unroll
    for (int cc = 0; cc < DH / 16; ++cc) {
      const int c = half * (DH / 16) + cc;
      cp_async16(s_q + (c >> 3) * 16384 + sw128_offset(r, (c & 7) * 8), src + c * 8, row_ok ? 16 : 0);
    }
    cp_async_commit();
  }

  const int n_a = (w.a_len + kTcKeys - 1) / kTcKeys;
  const int own_len = w.tok0 + w.n_tok - w.kb0;
  const int n_chunks = n_a + (own_len + kTcKeys - 1) / kTcKeys;

  auto chunk_keys = [&](int c, int& nk, int& key0, bool& own, int& tm_row) {
    if (c < n_a) {
      own = false;
      key0 = c * kTcKeys;
      nk = min(kTcKeys, w.a_len - key0);
      tm_row = p.a_row0 + w.a_start + key0;
    } else {
      own = true;
      key0 = w.kb0 + (c - n_a) * kTcKeys;
      nk = min(kTcKeys, w.tok0 + w.n_tok - key0);
      tm_row = p.b_row0 + key0 + w.b_off;
    }
  };
  auto stage = [&](uint8_t* dst, const CUtensorMap* tm, uint64_t* bar, int tm_row) {   // warp 0
    if (lead) {
      mbar_arrive_expect_tx(bar, kChunkBytes);
#pragma unroll
      for (int sub = 0; sub < kSub; ++sub) tma_load_2d(dst + sub * 8192, tm, bar, kvh * DH + sub * 64, tm_row);
    }
  };
  auto stage_k = [&](int c) {
    int nk, key0, row; bool own;
    chunk_keys(c, nk, key0, own, row);
    stage(s_k, own ? &tm_kb : &tm_ka, &bar_k, row);
  };
  auto stage_v = [&](int c) {
    int nk, key0, row; bool own;
    chunk_keys(c, nk, key0, own, row);
    stage(s_v + (c & 1) * kSub * 8192, own ? &tm_vb : &tm_va, &bar_v[c & 1], row);
  };
  auto issue_s = [&](int c, uint32_t tmem) {   // S = Q K^T of chunk c (warp 0)
    int nk, key0, row; bool own;
    chunk_keys(c, nk, key0, own, row);
    const int nk16 = (nk + 15) & ~15;
    mbar_wait(&bar_k, static_cast<uint32_t>(c & 1));
    tc_fence_after();
    const uint32_t idesc = make_idesc_f16kind(128, nk16, Fmt16<T16>::code, Fmt16<T16>::code, 0);
    if (lead) {
#pragma unroll
      for (int kk = 0; kk < DH / 16; ++kk) {
        const uint64_t da = make_smem_desc_sw128(smem_u32(s_q) + (kk >> 2) * 16384 + (kk & 3) * 32);
        const uint64_t db = make_smem_desc_sw128(smem_u32(s_k) + (kk >> 2) * 8192 + (kk & 3) * 32);
        umma_bf16<1>(tmem, da, db, idesc, kk != 0 ? 1u : 0u);
      }
      umma_commit(&bar_s);
    }
    __syncwarp();
  };

  cp_async_wait<0>();
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();   // barriers initialised, TMEM allocated, Q staged
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    stage_k(0);
    stage_v(0);
    if (n_chunks > 1) stage_v(1);
    issue_s(0, tmem);
  }
  const uint32_t t_row = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  const uint32_t t_s = t_row + half * 32;                 // this thread's 32 S columns
  const uint32_t t_o = t_row + 64 + half * (DH / 2);      // this thread's DH/2 O columns

  float m_ref = -INFINITY, l_part = 0.f;

  for (int c = 0; c < n_chunks; ++c) {
    int nk, key0, tm_row_unused; bool own;
    chunk_keys(c, nk, key0, own, tm_row_unused);
    const int nk16 = (nk + 15) & ~15;

    mbar_wait(&bar_s, static_cast<uint32_t>(c & 1));
    __syncwarp();   // the tcgen05.ld / .st below are warp-collective: reconverge after the per-lane spin
    tc_fence_after();
    // ---- this thread's 32 keys of the row
    float sv[32];
    if (half * 32 < nk16) {   // warp-uniform
      uint32_t raw[32];
      tmem_ld32(t_s, raw);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) sv[i] = __uint_as_float(raw[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) sv[i] = 0.f;
    }
    float cmax = -INFINITY;
    if (!own && nk == kTcKeys) {   // CTA-uniform fast path: a full chunk of the shared prefix, every key visible
#pragma unroll
      for (int i = 0; i < 32; ++i) cmax = fmaxf(cmax, sv[i]);
    } else {
      int j_lo = 0, j_hi = nk - 1;
      if (own) {
        j_lo = max(0, seq_lo - key0);
        j_hi = min(nk - 1, rt - key0);
      }
      // visible keys among this thread's 32 as a bit mask: [j_lo, j_hi] clipped to the half's window
      const int lo = max(j_lo - half * 32, 0), hi = min(j_hi - half * 32, 31);
      uint32_t vmask = (hi >= lo) ? ((0xFFFFFFFFu >> (31 - hi)) & (0xFFFFFFFFu << lo)) : 0u;
      if (own && p.key_valid != nullptr && vmask != 0u) {
        const uint8_t* kv = p.key_valid + key0 + half * 32;
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (((vmask >> i) & 1u) && kv[i] == 0) vmask &= ~(1u << i);
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float val = ((vmask >> i) & 1u) ? sv[i] : -INFINITY;
        sv[i] = val;
        cmax = fmaxf(cmax, val);
      }
    }
    s_x[half][r] = cmax;
    tc_fence_before();
    __syncthreads();   // [A] every S(c) value is in registers: the S columns and the K buffer are free
    if (warp == 0 && c + 1 < n_chunks) stage_k(c + 1);
    const float cmax_s = fmaxf(s_x[0][r], s_x[1][r]) * p.scale_log2;   // -inf * positive = -inf

    // ---- Oc = P V of the previous chunk has to be complete before P / O are touched
    if (c > 0) {
      mbar_wait(&bar_o, static_cast<uint32_t>((c - 1) & 1));
      tc_fence_after();
      if (warp == 0 && c + 1 < n_chunks) stage_v(c + 1);   // its buffer was read by chunk c-1
      __syncwarp();
    }
    // ---- lazy rescale of O (TMEM) and of the running sum
    float corr = 1.f;
    bool grow = false;
    if (c == 0) {
      m_ref = cmax_s;
    } else if (cmax_s > m_ref + kGrow) {
      grow = true;
      corr = exp2f(m_ref - cmax_s);   // 0 when nothing was visible before
      m_ref = cmax_s;
      l_part *= corr;
    }
    if (__any_sync(0xffffffffu, grow)) {
#pragma unroll
      for (int h = 0; h < DH / 64; ++h) {
        uint32_t raw[32];
        tmem_ld32(t_o + h * 32, raw);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) raw[i] = __float_as_uint(__uint_as_float(raw[i]) * corr);
        tmem_st32(t_o + h * 32, raw);
      }
      tmem_st_wait();
    }
    // ---- P for this thread's 32 keys
    const float m_use = (m_ref == -INFINITY) ? 0.f : m_ref;
    float csum = 0.f;
#pragma unroll
    for (int j8 = 0; j8 < 4; ++j8) {
      uint32_t pk[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float p0 = fast_exp2(fmaf(sv[j8 * 8 + 2 * e], p.scale_log2, -m_use));
        const float p1 = fast_exp2(fmaf(sv[j8 * 8 + 2 * e + 1], p.scale_log2, -m_use));
        csum += p0 + p1;
        pk[e] = Fmt16<T16>::pack2(p0, p1);
      }
      if (half * 32 + j8 * 8 < nk16) *reinterpret_cast<uint4*>(s_p + sw128_offset(r, half * 32 + j8 * 8)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
    l_part += csum;

    fence_proxy_async();
    tc_fence_before();
    __syncthreads();   // [B] P written, O rescaled
    if (warp == 0) {
      mbar_wait(&bar_v[c & 1], static_cast<uint32_t>((c >> 1) & 1));
      __syncwarp();
      tc_fence_after();
      const uint32_t idesc = make_idesc_f16kind(128, DH, Fmt16<T16>::code, Fmt16<T16>::code, 1);
      const uint32_t v_base = smem_u32(s_v) + (c & 1) * kSub * 8192;
      if (lead) {
        for (int kk = 0; kk < nk16 / 16; ++kk) {
          const uint64_t da = make_smem_desc_sw128(smem_u32(s_p) + kk * 32);
          const uint64_t db = make_smem_desc_raw(v_base + kk * 2048, 8192, 1024);
          umma_bf16<1>(tmem + 64, da, db, idesc, (c > 0 || kk != 0) ? 1u : 0u);
        }
        umma_commit(&bar_o);
      }
      __syncwarp();
      if (c + 1 < n_chunks) issue_s(c + 1, tmem);
    }
  }

  // ---- O / l -> bf16
  mbar_wait(&bar_o, static_cast<uint32_t>((n_chunks - 1) & 1));
  __syncwarp();
  tc_fence_after();
  s_x[half][r] = l_part;
  __syncthreads();
  const float l = s_x[0][r] + s_x[1][r];
  const float inv = l > 0.f ? 1.0f / l : 0.f;
  T16* dst = reinterpret_cast<T16*>(p.o) + static_cast<size_t>(rt) * p.n_q + (kvh * G + head) * DH + half * (DH / 2);
#pragma unroll
  for (int h = 0; h < DH / 64; ++h) {
    uint32_t raw[32];
    tmem_ld32(t_o + h * 32, raw);
    tmem_ld_wait();
    if (row_ok) {
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        uint32_t pk[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          pk[e] = Fmt16<T16>::pack2(__uint_as_float(raw[c8 * 8 + 2 * e]) * inv, __uint_as_float(raw[c8 * 8 + 2 * e + 1]) * inv);
        }
        *reinterpret_cast<uint4*>(dst + h * 32 + c8 * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) tmem_dealloc<1>(tmem, kTcTmemCols);
}

// ------------------------------------------------------------------------------------------------ v2p: persistent v2
// The scoring path's work items are short (282 prefix keys + a few own keys = 5-6 chunks), so what a v2 CTA does once --
// barrier initialisation, TMEM allocation, tensor-map prefetch, the CTA launch itself -- is a large part of its life.
// v2p launches two CTAs per SM once and lets each walk over work items (item = work x KV head, consecutive items share the
// prefix K/V in L2); the mbarrier phases simply keep counting chunks across items (g = chunks done so far + c).
template <int DH, typename T16>
__global__ void __launch_bounds__(kTc2Threads, 2)
attention_tc2p_kernel(const __grid_constant__ CUtensorMap tm_ka, const __grid_constant__ CUtensorMap tm_va,
                     const __grid_constant__ CUtensorMap tm_kb, const __grid_constant__ CUtensorMap tm_vb, const AttnParamsTc p) {
  constexpr int kSub = DH / 64;
  constexpr uint32_t kChunkBytes = kSub * 8192;
  constexpr float kGrow = 8.0f;  // lazy rescale threshold (log2 units): P stays below 2^8
  extern __shared__ uint8_t smem_raw_tc[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw_tc) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* s_q = smem;
  uint8_t* s_k = s_q + kSub * 16384;
  uint8_t* s_v = s_k + kSub * 8192;          // two stages
  uint8_t* s_p = s_v + 2 * kSub * 8192;
  __shared__ uint64_t bar_s, bar_o, bar_k, bar_v[2];
  __shared__ uint32_t tmem_slot;
  __shared__ float s_x[2][128];              // per-row exchange between the two threads of a row (max, then sum)

  AttnWorkTc w;
  int kvh = 0, g0 = 0, n_a = 0, n_chunks = 0;   // per item; g0 = chunks of the items this CTA has finished
  const int G = p.group;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int r = tid & 127, half = tid >> 7;
  // Warp 0 issues every TMA / tcgen05.mma with uniform control flow (all 32 lanes walk the issue code, so addresses and
  // descriptors are computed once in uniform registers) and only the issuing instruction is predicated on the elected lane:
  // the issue work sits on the CTA's critical path (everyone waits for warp 0 at the barriers).
  const bool lead = warp == 0 ? elect_one() != 0 : false;

  if (tid == 0) {
    mbar_init(&bar_s, 1);
    mbar_init(&bar_o, 1);
    mbar_init(&bar_k, 1);
    mbar_init(&bar_v[0], 1);
    mbar_init(&bar_v[1], 1);
    fence_barrier_init();
    tma_prefetch_desc(&tm_ka);
    tma_prefetch_desc(&tm_va);
    tma_prefetch_desc(&tm_kb);
    tma_prefetch_desc(&tm_vb);
  }
  if (warp == 0) tmem_alloc<1>(&tmem_slot, kTcTmemCols);
  const int tok_local = r / G, head = r - tok_local * G;

  auto chunk_keys = [&](int c, int& nk, int& key0, bool& own, int& tm_row) {
    if (c < n_a) {
      own = false;
      key0 = c * kTcKeys;
      nk = min(kTcKeys, w.a_len - key0);
      tm_row = p.a_row0 + w.a_start + key0;
    } else {
      own = true;
      key0 = w.kb0 + (c - n_a) * kTcKeys;
      nk = min(kTcKeys, w.tok0 + w.n_tok - key0);
      tm_row = p.b_row0 + key0 + w.b_off;
    }
  };
  auto stage = [&](uint8_t* dst, const CUtensorMap* tm, uint64_t* bar, int tm_row) {   // warp 0
    if (lead) {
      mbar_arrive_expect_tx(bar, kChunkBytes);
#pragma unroll
      for (int sub = 0; sub < kSub; ++sub) tma_load_2d(dst + sub * 8192, tm, bar, kvh * DH + sub * 64, tm_row);
    }
  };
  auto stage_k = [&](int c) {
    int nk, key0, row; bool own;
    chunk_keys(c, nk, key0, own, row);
    stage(s_k, own ? &tm_kb : &tm_ka, &bar_k, row);
  };
  auto stage_v = [&](int c) {
    int nk, key0, row; bool own;
    chunk_keys(c, nk, key0, own, row);
    stage(s_v + ((g0 + c) & 1) * kSub * 8192, own ? &tm_vb : &tm_va, &bar_v[(g0 + c) & 1], row);
  };
  auto issue_s = [&](int c, uint32_t tmem) {   // S = Q K^T of chunk c (warp 0)
    int nk, key0, row; bool own;
    chunk_keys(c, nk, key0, own, row);
    const int nk16 = (nk + 15) & ~15;
    mbar_wait(&bar_k, static_cast<uint32_t>((g0 + c) & 1));
    tc_fence_after();
    const uint32_t idesc = make_idesc_f16kind(128, nk16, Fmt16<T16>::code, Fmt16<T16>::code, 0);
    if (lead) {
#pragma unroll
      for (int kk = 0; kk < DH / 16; ++kk) {
        const uint64_t da = make_smem_desc_sw128(smem_u32(s_q) + (kk >> 2) * 16384 + (kk & 3) * 32);
        const uint64_t db = make_smem_desc_sw128(smem_u32(s_k) + (kk >> 2) * 8192 + (kk & 3) * 32);
        umma_bf16<1>(tmem, da, db, idesc, kk != 0 ? 1u : 0u);
      }
      umma_commit(&bar_s);
    }
    __syncwarp();
  };

  tc_fence_before();
  __syncthreads();   // barriers initialised, TMEM allocated
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t t_row = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  const uint32_t t_s = t_row + half * 32;                 // this thread's 32 S columns
  const uint32_t t_o = t_row + 64 + half * (DH / 2);      // this thread's DH/2 O columns
  const int n_items = p.n_works * p.n_kv_heads;

  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
  w = p.works[item % p.n_works];
  kvh = item / p.n_works;
  const bool row_ok = tok_local < w.n_tok;
  const int rt = w.tok0 + (row_ok ? tok_local : 0);
  const int seq_lo = (row_ok && p.tok_seq_start != nullptr) ? __ldg(p.tok_seq_start + rt) : 0;
  {
    const T16* src = reinterpret_cast<const T16*>(p.q) + static_cast<size_t>(rt) * p.q_stride + (kvh * G + head) * DH;
#pragma unroll
    for (int cc = 0; cc < DH / 16; ++cc) {
      const int c = half * (DH / 16) + cc;
      cp_async16(s_q + (c >> 3) * 16384 + sw128_offset(r, (c & 7) * 8), src + c * 8, row_ok ? 16 : 0);
    }
    cp_async_commit();
  }
  n_a = (w.a_len + kTcKeys - 1) / kTcKeys;
  n_chunks = n_a + (w.tok0 + w.n_tok - w.kb0 + kTcKeys - 1) / kTcKeys;
  if (warp == 0) {   // the K / V stages are free (warp 0 itself waited for the previous item's last S and P V): their TMA
    stage_k(0);      // latency overlaps the Q rows that are still in flight
    stage_v(0);
    if (n_chunks > 1) stage_v(1);
  }
  cp_async_wait<0>();
  fence_proxy_async();
  __syncthreads();   // Q staged; every thread is done with the previous item (its O rows are read, its exchange slots are free)
  if (warp == 0) issue_s(0, tmem);
  float m_ref = -INFINITY, l_part = 0.f;

  for (int c = 0; c < n_chunks; ++c) {
    int nk, key0, tm_row_unused; bool own;
    chunk_keys(c, nk, key0, own, tm_row_unused);
    const int nk16 = (nk + 15) & ~15;

    mbar_wait(&bar_s, static_cast<uint32_t>((g0 + c) & 1));
    __syncwarp();   // the tcgen05.ld / .st below are warp-collective: reconverge after the per-lane spin
    tc_fence_after();
    // ---- this thread's 32 keys of the row
    float sv[32];
    if (half * 32 < nk16) {   // warp-uniform
      uint32_t raw[32];
      tmem_ld32(t_s, raw);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) sv[i] = __uint_as_float(raw[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) sv[i] = 0.f;
    }
    float cmax = -INFINITY;
    if (!own && nk == kTcKeys) {   // CTA-uniform fast path: a full chunk of the shared prefix, every key visible
#pragma unroll
      for (int i = 0; i < 32; ++i) cmax = fmaxf(cmax, sv[i]);
    } else {
      int j_lo = 0, j_hi = nk - 1;
      if (own) {
        j_lo = max(0, seq_lo - key0);
        j_hi = min(nk - 1, rt - key0);
      }
      // visible keys among this thread's 32 as a bit mask: [j_lo, j_hi] clipped to the half's window
      const int lo = max(j_lo - half * 32, 0), hi = min(j_hi - half * 32, 31);
      uint32_t vmask = (hi >= lo) ? ((0xFFFFFFFFu >> (31 - hi)) & (0xFFFFFFFFu << lo)) : 0u;
      if (own && p.key_valid != nullptr && vmask != 0u) {
        const uint8_t* kv = p.key_valid + key0 + half * 32;
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (((vmask >> i) & 1u) && kv[i] == 0) vmask &= ~(1u << i);
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float val = ((vmask >> i) & 1u) ? sv[i] : -INFINITY;
        sv[i] = val;
        cmax = fmaxf(cmax, val);
      }
    }
    s_x[half][r] = cmax;
    tc_fence_before();
    __syncthreads();   // [A] every S(c) value is in registers: the S columns and the K buffer are free
    if (warp == 0 && c + 1 < n_chunks) stage_k(c + 1);
    const float cmax_s = fmaxf(s_x[0][r], s_x[1][r]) * p.scale_log2;   // -inf * positive = -inf

    // ---- Oc = P V of the previous chunk has to be complete before P / O are touched
    if (c > 0) {
      mbar_wait(&bar_o, static_cast<uint32_t>((g0 + c - 1) & 1));
      tc_fence_after();
      if (warp == 0 && c + 1 < n_chunks) stage_v(c + 1);   // its buffer was read by chunk c-1
      __syncwarp();
    }
    // ---- lazy rescale of O (TMEM) and of the running sum
    float corr = 1.f;
    bool grow = false;
    if (c == 0) {
      m_ref = cmax_s;
    } else if (cmax_s > m_ref + kGrow) {
      grow = true;
      corr = exp2f(m_ref - cmax_s);   // 0 when nothing was visible before
      m_ref = cmax_s;
      l_part *= corr;
    }
    if (__any_sync(0xffffffffu, grow)) {
#pragma unroll
      for (int h = 0; h < DH / 64; ++h) {
        uint32_t raw[32];
        tmem_ld32(t_o + h * 32, raw);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) raw[i] = __float_as_uint(__uint_as_float(raw[i]) * corr);
        tmem_st32(t_o + h * 32, raw);
      }
      tmem_st_wait();
    }
    // ---- P for this thread's 32 keys
    const float m_use = (m_ref == -INFINITY) ? 0.f : m_ref;
    float csum = 0.f;
#pragma unroll
    for (int j8 = 0; j8 < 4; ++j8) {
      uint32_t pk[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float p0 = fast_exp2(fmaf(sv[j8 * 8 + 2 * e], p.scale_log2, -m_use));
        const float p1 = fast_exp2(fmaf(sv[j8 * 8 + 2 * e + 1], p.scale_log2, -m_use));
        csum += p0 + p1;
        pk[e] = Fmt16<T16>::pack2(p0, p1);
      }
      if (half * 32 + j8 * 8 < nk16) *reinterpret_cast<uint4*>(s_p + sw128_offset(r, half * 32 + j8 * 8)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
    l_part += csum;

    fence_proxy_async();
    tc_fence_before();
    __syncthreads();   // [B] P written, O rescaled
    if (warp == 0) {
      mbar_wait(&bar_v[(g0 + c) & 1], static_cast<uint32_t>(((g0 + c) >> 1) & 1));
      __syncwarp();
      tc_fence_after();
      const uint32_t idesc = make_idesc_f16kind(128, DH, Fmt16<T16>::code, Fmt16<T16>::code, 1);
      const uint32_t v_base = smem_u32(s_v) + ((g0 + c) & 1) * kSub * 8192;
      if (lead) {
        for (int kk = 0; kk < nk16 / 16; ++kk) {
          const uint64_t da = make_smem_desc_sw128(smem_u32(s_p) + kk * 32);
          const uint64_t db = make_smem_desc_raw(v_base + kk * 2048, 8192, 1024);
          umma_bf16<1>(tmem + 64, da, db, idesc, (c > 0 || kk != 0) ? 1u : 0u);
        }
        umma_commit(&bar_o);
      }
      __syncwarp();
      if (c + 1 < n_chunks) issue_s(c + 1, tmem);
    }
  }

  // ---- O / l -> bf16
  mbar_wait(&bar_o, static_cast<uint32_t>((g0 + n_chunks - 1) & 1));
  __syncwarp();
  tc_fence_after();
  s_x[half][r] = l_part;
  __syncthreads();
  const float l = s_x[0][r] + s_x[1][r];
  const float inv = l > 0.f ? 1.0f / l : 0.f;
  T16* dst = reinterpret_cast<T16*>(p.o) + static_cast<size_t>(rt) * p.n_q + (kvh * G + head) * DH + half * (DH / 2);
#pragma unroll
  for (int h = 0; h < DH / 64; ++h) {
    uint32_t raw[32];
    tmem_ld32(t_o + h * 32, raw);
    tmem_ld_wait();
    if (row_ok) {
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        uint32_t pk[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          pk[e] = Fmt16<T16>::pack2(__uint_as_float(raw[c8 * 8 + 2 * e]) * inv, __uint_as_float(raw[c8 * 8 + 2 * e + 1]) * inv);
        }
        *reinterpret_cast<uint4*>(dst + h * 32 + c8 * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
    }
  }
  g0 += n_chunks;
  }   // items
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) tmem_dealloc<1>(tmem, kTcTmemCols);
}

// ------------------------------------------------------------------------------------------------ v4: dedicated issue warp
// Two S accumulators in TMEM, K and V double-buffered, and every TMA and tcgen05.mma is issued by a
// NINTH warp that does nothing else.  In v2 thread 0 issues them between its own softmax work, so warp 0 executes
// ~2x the instructions of the other warps and the whole CTA waits for it at every __syncthreads: the per-chunk critical
// path was warp 0's instruction stream (ncu: 42 % issue utilisation, 17 % tensor pipe on the ViT shapes).  Here the
// 256 softmax threads only talk to the issue warp through mbarriers:
//     bar_s[2]      S(c) complete                      (tcgen05.commit)      -> softmax threads, issue warp
//     bar_sfree[2]  every thread holds its S(c) row    (256 arrivals)        -> issue warp may overwrite that accumulator
//     bar_p         P(c) in shared memory, O rescaled  (256 arrivals)        -> issue warp starts Oc = P V
//     bar_o[2]      Oc of chunk c complete             (tcgen05.commit)      -> P tile / V stage free, O readable
// and among themselves only pairwise: the two threads of a row sit in warps w and w + 4 (same TMEM lane quadrant), which
// exchange the row maximum through shared memory and a 64-thread named barrier of their own (ids 1..4).  S(c+1) = Q K(c+1)^T is issued as
// soon as S(c) completes, i.e. it runs under the softmax of chunk c.  head_dim 64 has room for TWO P tiles: P(c+1) is
// written while Oc(c) = P(c) V(c) is still running, and a thread waits for Oc(c-1) only when it must rescale O.
constexpr int kTc4Threads = 288;

template <int DH>
constexpr int attn_tc4_smem_bytes() {
  // Q (DH/64 x 16 KB) + 2 x K + 2 x V (DH/64 x 8 KB each) + P (16 KB; two for head_dim 64) + max exchange (512 B; two
  // for head_dim 64) + barriers (96 B)
  return (DH / 64) * 16384 + 4 * (DH / 64) * 8192 + (DH == 64 ? 2 : 1) * (16384 + 512) + 96;
}

// barrier of the two warps (w, w + 4) that share TMEM lane quadrant q = w & 3
__device__ __forceinline__ void softmax_pair_sync(int q) { asm volatile("bar.sync %0, 64;" ::"r"(q + 1) : "memory"); }

template <int DH, typename T16>
__global__ void __launch_bounds__(kTc4Threads, 2)
attention_tc4_kernel(const __grid_constant__ CUtensorMap tm_ka, const __grid_constant__ CUtensorMap tm_va,
                     const __grid_constant__ CUtensorMap tm_kb, const __grid_constant__ CUtensorMap tm_vb, const AttnParamsTc p) {
  constexpr int kSub = DH / 64;
  constexpr uint32_t kChunkBytes = kSub * 8192;
  constexpr float kGrow = 8.0f;
  constexpr bool kPDouble = DH == 64;
  constexpr int kPBytes = (kPDouble ? 2 : 1) * 16384;
  constexpr int kMxBytes = (kPDouble ? 2 : 1) * 512;
  extern __shared__ __align__(1024) uint8_t smem_tc4[];
  uint8_t* s_q = smem_tc4;
  uint8_t* s_k = s_q + kSub * 16384;           // two stages
  uint8_t* s_v = s_k + 2 * kSub * 8192;        // two stages
  uint8_t* s_p = s_v + 2 * kSub * 8192;
  __nv_bfloat16* s_mx = reinterpret_cast<__nv_bfloat16*>(s_p + kPBytes);   // [2][128] row maxima of the two half-row threads
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_p + kPBytes + kMxBytes);
  uint64_t* bar_s = bars;        // [2]
  uint64_t* bar_k = bars + 2;    // [2]
  uint64_t* bar_v = bars + 4;    // [2]
  uint64_t* bar_o = bars + 6;    // [2]  Oc(c) -> bar_o[c & 1]
  uint64_t* bar_sfree = bars + 8;   // [2]
  uint64_t* bar_p = bars + 10;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 11);

  const AttnWorkTc w = p.works[blockIdx.x];
  const int kvh = blockIdx.y;
  const int G = p.group;
  const int tid = threadIdx.x, warp = tid >> 5;
  const bool issuer = warp == 8;                 // warp 8: TMA + MMA issue only
  const int r = tid & 127, half = (tid >> 7) & 1;

  if (tid == 0) {
    if (smem_u32(smem_tc4) & 1023u) __trap();   // the swizzled tiles need a 1024-byte aligned base
    for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1);
    mbar_init(&bar_sfree[0], 256);
    mbar_init(&bar_sfree[1], 256);
    mbar_init(bar_p, 256);
    fence_barrier_init();
    tma_prefetch_desc(&tm_ka);
    tma_prefetch_desc(&tm_va);
    tma_prefetch_desc(&tm_kb);
    tma_prefetch_desc(&tm_vb);
  }
  if (warp == 0) tmem_alloc<1>(tmem_slot, kTcTmemCols);

  const int tok_local = r / G, head = r - tok_local * G;
  const bool row_ok = tok_local < w.n_tok;
  const int rt = w.tok0 + (row_ok ? tok_local : 0);
  const int seq_lo = (row_ok && p.tok_seq_start != nullptr) ? __ldg(p.tok_seq_start + rt) : 0;
  if (!issuer) {
    const T16* src = reinterpret_cast<const T16*>(p.q) + static_cast<size_t>(rt) * p.q_stride + (kvh * G + head) * DH;
#pragma unroll
    for (int cc = 0; cc < DH / 16; ++cc) {
      const int c = half * (DH / 16) + cc;
      cp_async16(s_q + (c >> 3) * 16384 + sw128_offset(r, (c & 7) * 8), src + c * 8, row_ok ? 16 : 0);
    }
    cp_async_commit();
  }

  const int n_a = (w.a_len + kTcKeys - 1) / kTcKeys;
  const int own_len = w.tok0 + w.n_tok - w.kb0;
  const int n_chunks = n_a + (own_len + kTcKeys - 1) / kTcKeys;

  auto chunk_keys = [&](int c, int& nk, int& key0, bool& own, int& tm_row) {
    if (c < n_a) {
      own = false;
      key0 = c * kTcKeys;
      nk = min(kTcKeys, w.a_len - key0);
      tm_row = p.a_row0 + w.a_start + key0;
    } else {
      own = true;
      key0 = w.kb0 + (c - n_a) * kTcKeys;
      nk = min(kTcKeys, w.tok0 + w.n_tok - key0);
      tm_row = p.b_row0 + key0 + w.b_off;
    }
  };
  // The issue warp runs its loop with all 32 lanes (uniform control flow: waits, addresses and descriptors live in uniform
  // registers) and only the instruction that talks to the TMA / tensor core is predicated on the elected lane.
  const bool lead = issuer && elect_one() != 0;
  auto stage = [&](uint8_t* dst, const CUtensorMap* tm, uint64_t* bar, int tm_row) {
    if (lead) {
      mbar_arrive_expect_tx(bar, kChunkBytes);
#pragma unroll
      for (int sub = 0; sub < kSub; ++sub) tma_load_2d(dst + sub * 8192, tm, bar, kvh * DH + sub * 64, tm_row);
    }
  };
  auto stage_k = [&](int c) {
    int nk, key0, row; bool own;
    chunk_keys(c, nk, key0, own, row);
    stage(s_k + (c & 1) * kSub * 8192, own ? &tm_kb : &tm_ka, &bar_k[c & 1], row);
  };
  auto stage_v = [&](int c) {
    int nk, key0, row; bool own;
    chunk_keys(c, nk, key0, own, row);
    stage(s_v + (c & 1) * kSub * 8192, own ? &tm_vb : &tm_va, &bar_v[c & 1], row);
  };
  auto issue_s = [&](int c, uint32_t tmem) {   // S[c & 1] = Q K(c)^T (issue warp)
    int nk, key0, row; bool own;
    chunk_keys(c, nk, key0, own, row);
    const int nk16 = (nk + 15) & ~15;
    mbar_wait(&bar_k[c & 1], static_cast<uint32_t>((c >> 1) & 1));
    tc_fence_after();
    const uint32_t idesc = make_idesc_f16kind(128, nk16, Fmt16<T16>::code, Fmt16<T16>::code, 0);
    const uint32_t k_base = smem_u32(s_k) + (c & 1) * kSub * 8192;
    if (lead) {
#pragma unroll
      for (int kk = 0; kk < DH / 16; ++kk) {
        const uint64_t da = make_smem_desc_sw128(smem_u32(s_q) + (kk >> 2) * 16384 + (kk & 3) * 32);
        const uint64_t db = make_smem_desc_sw128(k_base + (kk >> 2) * 8192 + (kk & 3) * 32);
        umma_bf16<1>(tmem + (c & 1) * 64, da, db, idesc, kk != 0 ? 1u : 0u);
      }
      umma_commit(&bar_s[c & 1]);
    }
    __syncwarp();
  };

  cp_async_wait<0>();
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();   // barriers initialised, TMEM allocated, Q staged
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (issuer) {
    // ===================================================== issue warp
    {
      stage_k(0);
      stage_v(0);
      if (n_chunks > 1) { stage_k(1); stage_v(1); }
      issue_s(0, tmem);
      for (int c = 0; c < n_chunks; ++c) {
        int nk, key0, tm_row_unused; bool own;
        chunk_keys(c, nk, key0, own, tm_row_unused);
        const int nk16 = (nk + 15) & ~15;
        mbar_wait(&bar_s[c & 1], static_cast<uint32_t>((c >> 1) & 1));   // S(c) complete: its K stage is free
        if (c + 1 < n_chunks) {
          if (c >= 1) mbar_wait(&bar_sfree[(c + 1) & 1], static_cast<uint32_t>(((c - 1) >> 1) & 1));   // S(c-1) rows are in registers
          issue_s(c + 1, tmem);
        }
        if (c + 2 < n_chunks) stage_k(c + 2);
        if (c >= 1 && c + 1 < n_chunks) {
          mbar_wait(&bar_o[(c - 1) & 1], static_cast<uint32_t>(((c - 1) >> 1) & 1));   // Oc(c-1) complete: its V stage is free
          stage_v(c + 1);
        }
        mbar_wait(bar_p, static_cast<uint32_t>(c & 1));                  // P(c) written, O rescaled
        mbar_wait(&bar_v[c & 1], static_cast<uint32_t>((c >> 1) & 1));
        tc_fence_after();
        const uint32_t idesc = make_idesc_f16kind(128, DH, Fmt16<T16>::code, Fmt16<T16>::code, 1);
        const uint32_t v_base = smem_u32(s_v) + (c & 1) * kSub * 8192;
        if (lead) {
          if (nk16 == kTcKeys) {
#pragma unroll
            for (int kk = 0; kk < kTcKeys / 16; ++kk) {
              const uint64_t da = make_smem_desc_sw128(smem_u32(s_p) + (kPDouble ? (c & 1) * 16384 : 0) + kk * 32);
              const uint64_t db = make_smem_desc_raw(v_base + kk * 2048, 8192, 1024);
              umma_bf16<1>(tmem + 128, da, db, idesc, (c > 0 || kk != 0) ? 1u : 0u);
            }
          } else {
            for (int kk = 0; kk < nk16 / 16; ++kk) {
              const uint64_t da = make_smem_desc_sw128(smem_u32(s_p) + (kPDouble ? (c & 1) * 16384 : 0) + kk * 32);
              const uint64_t db = make_smem_desc_raw(v_base + kk * 2048, 8192, 1024);
              umma_bf16<1>(tmem + 128, da, db, idesc, (c > 0 || kk != 0) ? 1u : 0u);
            }
          }
          umma_commit(&bar_o[c & 1]);
        }
        __syncwarp();
      }
      // keep the CTA's shared memory alive until the last MMAs read it
      if (n_chunks > 1) mbar_wait(&bar_o[(n_chunks - 2) & 1], static_cast<uint32_t>(((n_chunks - 2) >> 1) & 1));
      mbar_wait(&bar_o[(n_chunks - 1) & 1], static_cast<uint32_t>(((n_chunks - 1) >> 1) & 1));
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();   // matches the softmax threads' final barrier (TMEM release)
    return;
  }
  // ===================================================== 256 softmax threads
  const uint32_t t_row = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  const uint32_t t_o = t_row + 128 + half * (DH / 2);      // this thread's DH/2 O columns (S0: [0,64), S1: [64,128))

  float m_ref = -INFINITY, l_part = 0.f;

  for (int c = 0; c < n_chunks; ++c) {
    int nk, key0, tm_row_unused; bool own;
    chunk_keys(c, nk, key0, own, tm_row_unused);
    const int nk16 = (nk + 15) & ~15;

    mbar_wait(&bar_s[c & 1], static_cast<uint32_t>((c >> 1) & 1));
    __syncwarp();
    tc_fence_after();
    float sv[32];
    if (half * 32 < nk16) {   // warp-uniform
      uint32_t raw[32];
      tmem_ld32(t_row + (c & 1) * 64 + half * 32, raw);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) sv[i] = __uint_as_float(raw[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) sv[i] = 0.f;
    }
    float cmax = -INFINITY;
    if (!own && nk == kTcKeys) {   // CTA-uniform fast path: a full chunk of the shared prefix, every key visible
#pragma unroll
      for (int i = 0; i < 32; ++i) cmax = fmaxf(cmax, sv[i]);
    } else {
      int j_lo = 0, j_hi = nk - 1;
      if (own) {
        j_lo = max(0, seq_lo - key0);
        j_hi = min(nk - 1, rt - key0);
      }
      // visible keys among this thread's 32 as a bit mask: [j_lo, j_hi] clipped to the half's window
      const int lo = max(j_lo - half * 32, 0), hi = min(j_hi - half * 32, 31);
      uint32_t vmask = (hi >= lo) ? ((0xFFFFFFFFu >> (31 - hi)) & (0xFFFFFFFFu << lo)) : 0u;
      if (own && p.key_valid != nullptr && vmask != 0u) {
        const uint8_t* kv = p.key_valid + key0 + half * 32;
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (((vmask >> i) & 1u) && kv[i] == 0) vmask &= ~(1u << i);
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float val = ((vmask >> i) & 1u) ? sv[i] : -INFINITY;
        sv[i] = val;
        cmax = fmaxf(cmax, val);
      }
    }
    tc_fence_before();
    mbar_arrive(&bar_sfree[c & 1]);                  // this thread's S(c) values are in registers
    __nv_bfloat16* mx = s_mx + (kPDouble ? (c & 1) * 256 : 0);
    mx[half * 128 + r] = __float2bfloat16(cmax);   // both threads of the row use the same (bf16-rounded) pair of maxima
    softmax_pair_sync(warp & 3);   // [A] the partner's maximum is visible
    const float cmax_s = fmaxf(__bfloat162float(mx[r]), __bfloat162float(mx[128 + r])) * p.scale_log2;
    if constexpr (!kPDouble) softmax_pair_sync(warp & 3);   // [A'] single exchange buffer: read before the next chunk overwrites it

    if constexpr (!kPDouble) {
      if (c > 0) {
        mbar_wait(&bar_o[(c - 1) & 1], static_cast<uint32_t>(((c - 1) >> 1) & 1));   // Oc(c-1) complete: P tile free, O readable
        __syncwarp();
        tc_fence_after();
      }
    } else if (c >= 2) {
      mbar_wait(&bar_o[c & 1], static_cast<uint32_t>(((c - 2) >> 1) & 1));           // Oc(c-2) complete: P tile (c & 1) is free
      __syncwarp();
    }
    float corr = 1.f;
    bool grow = false;
    if (c == 0) {
      m_ref = cmax_s;
    } else if (cmax_s > m_ref + kGrow) {
      grow = true;
      corr = exp2f(m_ref - cmax_s);
      m_ref = cmax_s;
      l_part *= corr;
    }
    if (__any_sync(0xffffffffu, grow)) {
      if constexpr (kPDouble) {   // O is only touched here: wait for the accumulation that is still in flight
        mbar_wait(&bar_o[(c - 1) & 1], static_cast<uint32_t>(((c - 1) >> 1) & 1));
        __syncwarp();
        tc_fence_after();
      }
#pragma unroll
      for (int h = 0; h < DH / 64; ++h) {
        uint32_t raw[32];
        tmem_ld32(t_o + h * 32, raw);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) raw[i] = __float_as_uint(__uint_as_float(raw[i]) * corr);
        tmem_st32(t_o + h * 32, raw);
      }
      tmem_st_wait();
    }
    const float m_use = (m_ref == -INFINITY) ? 0.f : m_ref;
    float csum = 0.f;
#pragma unroll
    for (int j8 = 0; j8 < 4; ++j8) {
      uint32_t pk[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float p0 = fast_exp2(fmaf(sv[j8 * 8 + 2 * e], p.scale_log2, -m_use));
        const float p1 = fast_exp2(fmaf(sv[j8 * 8 + 2 * e + 1], p.scale_log2, -m_use));
        csum += p0 + p1;
        pk[e] = Fmt16<T16>::pack2(p0, p1);
      }
      if (half * 32 + j8 * 8 < nk16)
        *reinterpret_cast<uint4*>(s_p + (kPDouble ? (c & 1) * 16384 : 0) + sw128_offset(r, half * 32 + j8 * 8)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
    l_part += csum;

    fence_proxy_async();
    tc_fence_before();
    mbar_arrive(bar_p);     // [B] this thread's part of P(c) is written, its O columns are rescaled
  }

  // ---- O / l -> bf16   (the P tile is free: reuse it to add the two half-row sums)
  if (n_chunks > 1) mbar_wait(&bar_o[(n_chunks - 2) & 1], static_cast<uint32_t>(((n_chunks - 2) >> 1) & 1));
  mbar_wait(&bar_o[(n_chunks - 1) & 1], static_cast<uint32_t>(((n_chunks - 1) >> 1) & 1));
  __syncwarp();
  tc_fence_after();
  float* s_l = reinterpret_cast<float*>(s_p);
  s_l[half * 128 + r] = l_part;
  softmax_pair_sync(warp & 3);
  const float l = s_l[r] + s_l[128 + r];
  const float inv = l > 0.f ? 1.0f / l : 0.f;
  T16* dst = reinterpret_cast<T16*>(p.o) + static_cast<size_t>(rt) * p.n_q + (kvh * G + head) * DH + half * (DH / 2);
#pragma unroll
  for (int h = 0; h < DH / 64; ++h) {
    uint32_t raw[32];
    tmem_ld32(t_o + h * 32, raw);
    tmem_ld_wait();
    if (row_ok) {
#pragma unroll
      for (int c8 = 0; c8 < 4; ++c8) {
        uint32_t pk[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          pk[e] = Fmt16<T16>::pack2(__uint_as_float(raw[c8 * 8 + 2 * e]) * inv, __uint_as_float(raw[c8 * 8 + 2 * e + 1]) * inv);
        }
        *reinterpret_cast<uint4*>(dst + h * 32 + c8 * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) tmem_dealloc<1>(tmem, kTcTmemCols);
}

// Host: blocks of 128 / G consecutive tokens over every prefix-sharing group of consecutive sequences, plus the
// per-token "first token of my sequence" table.
inline void build_attn_works_tc(const AttnSeq* seqs, int n_seqs, int group, std::vector<AttnWorkTc>& works, std::vector<int>& tok_seq_start,
                                int n_tokens) {
  works.clear();
  tok_seq_start.assign(n_tokens, 0);
  const int tpb = 128 / group;
  int s0 = 0;
  while (s0 < n_seqs) {
    int s1 = s0 + 1;
    if (seqs[s0].a_len > 0)
      while (s1 < n_seqs && seqs[s1].unit == seqs[s0].unit && seqs[s1].a_len == seqs[s0].a_len && seqs[s1].a_start == seqs[s0].a_start &&
             seqs[s1].q_start == seqs[s1 - 1].q_start + seqs[s1 - 1].q_len &&
             seqs[s1].b_start - seqs[s1].q_start == seqs[s0].b_start - seqs[s0].q_start)
        ++s1;
    for (int s = s0; s < s1; ++s)
      for (int t = 0; t < seqs[s].q_len; ++t) tok_seq_start[seqs[s].q_start + t] = seqs[s].q_start;
    const int g0 = seqs[s0].q_start, g1 = seqs[s1 - 1].q_start + seqs[s1 - 1].q_len;
    int cur = s0;
    for (int t0 = g0; t0 < g1; t0 += tpb) {
      while (seqs[cur].q_start + seqs[cur].q_len <= t0) ++cur;
      AttnWorkTc w;
      w.tok0 = t0;
      w.n_tok = (g1 - t0 < tpb) ? g1 - t0 : tpb;
      w.a_start = seqs[s0].a_start;
      w.a_len = seqs[s0].a_len;
      w.kb0 = seqs[cur].q_start;
      w.b_off = seqs[cur].b_start - seqs[cur].q_start;
      w.pad1 = w.pad2 = 0;
      works.push_back(w);
    }
    s0 = s1;
  }
}

// bf16 [rows, n_kv] K / V buffer -> tensor map with 64-column x 64-row boxes (SWIZZLE_128B)
inline bool make_kv_tmap(CUtensorMap* out, const void* base, uint64_t rows, uint64_t n_kv) {
  return make_tmap_bf16(out, base, rows, n_kv, n_kv, kTcKeys);
}

struct AttnTcMaps {
  CUtensorMap ka, va, kb, vb;
};

enum : int { kAttnPersistent = 5, kAttnPerItem = 2, kAttnIssueWarp = 4, kAttnWarpSpecialized = 6 };   // attention_tc2p / _tc2 / _tc4 / _tc5 (attention_ws.cuh)

// cudaFuncSetAttribute once per (kernel instantiation, device): function attributes are per device
template <class Kern>
inline cudaError_t attn_set_smem(Kern kern, int bytes, bool max_carveout, bool* done) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (done[dev & 63]) return cudaSuccess;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess && max_carveout) e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e == cudaSuccess) done[dev & 63] = true;
  return e;
}

template <int DH, typename T16>
inline cudaError_t launch_attention_tc2_impl(const AttnTcMaps& m, const AttnParamsTc& p, dim3 grid, cudaStream_t stream) {
  static bool done[64] = {};
  cudaError_t e = attn_set_smem(attention_tc2_kernel<DH, T16>, attn_tc2_smem_bytes<DH>(), false, done);
  if (e != cudaSuccess) return e;
  attention_tc2_kernel<DH, T16><<<grid, kTc2Threads, attn_tc2_smem_bytes<DH>(), stream>>>(m.ka, m.va, m.kb, m.vb, p);
  return cudaGetLastError();
}

// persistent v2: two CTAs per SM walk over the work items (work x KV head)
template <int DH, typename T16>
inline cudaError_t launch_attention_tc2p_impl(const AttnTcMaps& m, const AttnParamsTc& p, int n_works, int n_kv_heads, cudaStream_t stream) {
  static bool done[64] = {};
  cudaError_t e = attn_set_smem(attention_tc2p_kernel<DH, T16>, attn_tc2_smem_bytes<DH>(), false, done);
  if (e != cudaSuccess) return e;
  AttnParamsTc pp = p;
  pp.n_works = n_works;
  pp.n_kv_heads = n_kv_heads;
  int dev = 0, n_sm = 148;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n_sm = 148;
  const int items = n_works * n_kv_heads;
  dim3 pgrid(static_cast<unsigned>(std::min(items, 2 * n_sm)));
  attention_tc2p_kernel<DH, T16><<<pgrid, kTc2Threads, attn_tc2_smem_bytes<DH>(), stream>>>(m.ka, m.va, m.kb, m.vb, pp);
  return cudaGetLastError();
}

// v4 needs two 112.7 KB CTAs of 288 threads per SM; returns cudaErrorLaunchOutOfResources (caller falls back) otherwise.
template <int DH, typename T16>
inline cudaError_t launch_attention_tc4_impl(const AttnTcMaps& m, const AttnParamsTc& p, dim3 grid, cudaStream_t stream) {
  static int state[64] = {};   // per device: 0 = unknown, 1 = usable, -1 = does not fit twice
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  int& st = state[dev & 63];
  if (st == 0) {
    e = cudaFuncSetAttribute(attention_tc4_kernel<DH, T16>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_tc4_smem_bytes<DH>());
    if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_tc4_kernel<DH, T16>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    // two CTAs per SM?  (the occupancy API under-reports with large dynamic shared memory, so the two limits are checked
    // directly: 228 KB of shared memory per SM incl. 1 KB reserved per CTA, 64 K registers allocated in units of 8 per thread)
    cudaFuncAttributes fa;
    int blocks = 0, smem_sm = 0, regs_sm = 0;
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, attention_tc4_kernel<DH, T16>);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, dev);
    if (e == cudaSuccess) {
      const int smem_cta = ((attn_tc4_smem_bytes<DH>() + static_cast<int>(fa.sharedSizeBytes) + 127) & ~127) + 1024;
      const int regs_cta = kTc4Threads * ((fa.numRegs + 7) & ~7);
      blocks = std::min(smem_sm / smem_cta, regs_sm / regs_cta);
    } else {
      cudaGetLastError();
    }
    st = (e == cudaSuccess && blocks >= 2) ? 1 : -1;
    if (getenv("BLIM_DEBUG")) fprintf(stderr, "[blim] attention v4 (head_dim %d): %d CTA(s) per SM with %d B dynamic smem -> %s\n", DH, blocks,
                                      attn_tc4_smem_bytes<DH>(), st > 0 ? "used" : "falling back");
  }
  if (st < 0) return cudaErrorLaunchOutOfResources;
  attention_tc4_kernel<DH, T16><<<grid, kTc4Threads, attn_tc4_smem_bytes<DH>(), stream>>>(m.ka, m.va, m.kb, m.vb, p);
  return cudaGetLastError();
}

// T16 = operand format of Q / K / V / P / O.  version: kAttnPersistent (scoring path), kAttnPerItem, kAttnIssueWarp
// (head_dim 64 only: at 128 its 96-register budget spills; falls back to the per-item kernel when two CTAs do not fit).
template <typename T16>
inline cudaError_t launch_attention_tc(const AttnTcMaps& m, const AttnParamsTc& p, int n_works, int n_kv_heads, int head_dim,
                                       cudaStream_t stream, int version = kAttnPerItem) {
  if (n_works <= 0) return cudaSuccess;
  if (head_dim != 64 && head_dim != 128) return cudaErrorInvalidValue;
  dim3 grid(static_cast<unsigned>(n_works), static_cast<unsigned>(n_kv_heads));
  if (version == kAttnIssueWarp && head_dim == 64) {
    cudaError_t e = launch_attention_tc4_impl<64, T16>(m, p, grid, stream);
    if (e != cudaErrorLaunchOutOfResources) return e;
  }
  if (version == kAttnPersistent)
    return head_dim == 128 ? launch_attention_tc2p_impl<128, T16>(m, p, n_works, n_kv_heads, stream)
                           : launch_attention_tc2p_impl<64, T16>(m, p, n_works, n_kv_heads, stream);
  return head_dim == 128 ? launch_attention_tc2_impl<128, T16>(m, p, grid, stream) : launch_attention_tc2_impl<64, T16>(m, p, grid, stream);
}

}  // namespace blim
