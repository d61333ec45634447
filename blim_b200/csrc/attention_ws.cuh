// Warp-specialised tcgen05 attention for the scoring path ("v5"): same work items, masks and numerics as
// attention_tc2p_kernel (attention_tc.cuh; reference semantics: Qwen2SdpaAttention, modeling_qwen2_flash.py:685-709), but
// the per-chunk chain  K arrives -> S = Q K^T -> softmax -> P -> O += P V  is no longer executed by one group of threads
// that waits for every link.  A CTA is three roles that only meet at mbarriers (no __syncthreads in the main loop):
//
//   warps 0-3  softmax   ONE thread per query row (TMEM lane = row = thread): reads its 64 S values of the chunk with
//                        tcgen05.ld, keeps max / sum / lazy rescale entirely thread-local (no exchange between threads),
//                        writes its 128-byte P row to the swizzled shared-memory tile, normalises and stores O at the end
//   warp  4    loader    TMA only: the Q tile of the NEXT item (3-D box: head_dim x heads-of-the-group x tokens, i.e. the
//                        stacked (token, head) rows arrive in tile order) as soon as the last S of the current item has
//                        been issued, and K / V chunks through two-stage rings, up to two chunks ahead
//   warp  5    MMA       one thread issues every tcgen05.mma: S(c+1) is issued BEFORE P V(c), into the second S
//                        accumulator, so it runs while the softmax warps work on S(c); tcgen05.commit hands buffers back
//                        (K stage -> loader, S -> softmax, V stage + P tile -> loader / softmax, Q tile -> loader)
//
// 192 threads and 112.3 KB of shared memory per CTA, 256 TMEM columns (S0 | S1 | O): two CTAs per SM, each an independent
// pipeline.  What this removes from v2p's critical path (ncu there: tensor pipe 10 % active, issue slots 27 %): the two
// __syncthreads per chunk, the K chunk being requested only after S of the previous chunk was consumed, S(c+1) queued
// behind P V(c) with nobody working meanwhile, the softmax threads issuing TMA / MMA between their own work, and the Q tile
// being fetched by cp.async in front of every item.
// One thread per row is deliberate: a variant with two threads per row (8 softmax warps, the row maximum exchanged through
// shared memory and a 64-thread named barrier, 320 threads at 96 registers) passed the same tests and measured 256 ms of
// attention per C2 step against 204 ms for this kernel on the same box
// (profiles/r02_bench_c2_ab_attention_one_vs_two_threads_per_row_v22.log).
#pragma once
#include "attention_tc.cuh"

namespace blim {

constexpr int kTc5Threads = 192;

template <int N>
struct KeyCount { static constexpr int value = N; };   // compile-time element count of a softmax chunk body

template <int DH>
constexpr int attn_tc5_smem_bytes() {
  // Q (DH/64 x 16 KB) + 2 x K + 2 x V (DH/64 x 8 KB each) + P (16 KB) + barriers
  return (DH / 64) * 16384 + 4 * (DH / 64) * 8192 + 16384 + 256;
}

// 3-D tiled load global -> shared (c0 = innermost coordinate), completion on an mbarrier of this CTA
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* desc, uint64_t* bar, int32_t c0, int32_t c1, int32_t c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(desc)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// Q [T, n_heads * head_dim] as a 3-D tensor (head_dim, head, token); box = 64 columns x `group` heads x `tok_per_tile`
// tokens, 128-byte swizzle: the box lands as rows r = token * group + head of 128 bytes each = the stacked-row order of a tile.
inline bool make_q_tmap(CUtensorMap* out, const void* base, uint64_t tokens, int n_heads, int head_dim, int group, int tok_per_tile) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t gdim[3] = {static_cast<cuuint64_t>(head_dim), static_cast<cuuint64_t>(n_heads), tokens};
  cuuint64_t gstride[2] = {static_cast<cuuint64_t>(head_dim) * 2, static_cast<cuuint64_t>(n_heads) * head_dim * 2};
  cuuint32_t box[3] = {64, static_cast<cuuint32_t>(group), static_cast<cuuint32_t>(tok_per_tile)};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <int DH, typename T16>
__global__ void __launch_bounds__(kTc5Threads, 2)
attention_tc5_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_ka, const __grid_constant__ CUtensorMap tm_va,
                     const __grid_constant__ CUtensorMap tm_kb, const __grid_constant__ CUtensorMap tm_vb, const AttnParamsTc p) {
  constexpr int kSub = DH / 64;
  constexpr uint32_t kChunkBytes = kSub * 8192;
  constexpr float kGrow = 8.0f;   // lazy rescale threshold (log2 units): P stays below 2^8
  extern __shared__ __align__(1024) uint8_t smem_tc5[];
  uint8_t* s_q = smem_tc5;
  uint8_t* s_k = s_q + kSub * 16384;          // two stages
  uint8_t* s_v = s_k + 2 * kChunkBytes;       // two stages
  uint8_t* s_p = s_v + 2 * kChunkBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_p + 16384);
  uint64_t* bar_q_full = bars + 0;    // Q tile of the item landed                    (TMA tx)
  uint64_t* bar_q_free = bars + 1;    // last S of the item issued and complete      (tcgen05.commit)
  uint64_t* bar_o_free = bars + 2;    // softmax warps have read the item's O        (4 warp arrivals)
  uint64_t* bar_p_full = bars + 3;    // P(c) in shared memory, O rescaled           (4 warp arrivals)
  uint64_t* bar_pv_done = bars + 5;   // O += P V of chunk c complete                (tcgen05.commit)
  uint64_t* bar_k_full = bars + 7;    // [2] TMA tx
  uint64_t* bar_k_free = bars + 9;    // [2] tcgen05.commit behind S(c)
  uint64_t* bar_v_full = bars + 11;   // [2] TMA tx
  uint64_t* bar_v_free = bars + 13;   // [2] tcgen05.commit behind P V(c)
  uint64_t* bar_s_full = bars + 15;   // [2] tcgen05.commit behind S(c)
  uint64_t* bar_s_free = bars + 17;   // [2] softmax warps hold S(c) in registers    (4 warp arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = p.group;
  const int tpb = 128 / G;                       // tokens per tile (build_attn_works_tc)
  const int n_items = p.n_works * p.n_kv_heads;

  if (tid == 0) {
    if (smem_u32(smem_tc5) & 1023u) __trap();    // the swizzled tiles need a 1024-byte aligned base
    mbar_init(bar_q_full, 1); mbar_init(bar_q_free, 1); mbar_init(bar_o_free, 4); mbar_init(bar_p_full, 4); mbar_init(bar_pv_done, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_k_full[i], 1); mbar_init(&bar_k_free[i], 1); mbar_init(&bar_v_full[i], 1); mbar_init(&bar_v_free[i], 1);
      mbar_init(&bar_s_full[i], 1); mbar_init(&bar_s_free[i], 4);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_ka); tma_prefetch_desc(&tm_va); tma_prefetch_desc(&tm_kb); tma_prefetch_desc(&tm_vb);
  }
  // rows of the Q tile that no box row ever covers (128 - group * tok_per_tile of them) must hold finite numbers
  for (int i = tid; i < kSub * 16384 / 16; i += kTc5Threads) reinterpret_cast<uint4*>(s_q)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (warp == 4) tmem_alloc<1>(tmem_slot, kTcTmemCols);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  // chunk c of an item: first the a_len prefix keys (all visible), then the own-run keys [kb0, tok0 + n_tok)
  auto chunk_info = [](const AttnWorkTc& w, int n_a, int c, int& nk, int& key0, bool& own) {
    if (c < n_a) {
      own = false;
      key0 = c * kTcKeys;
      nk = min(kTcKeys, w.a_len - key0);
    } else {
      own = true;
      key0 = w.kb0 + (c - n_a) * kTcKeys;
      nk = min(kTcKeys, w.tok0 + w.n_tok - key0);
    }
  };

  if (warp == 4) {
    // ===================================================================================== loader
    if (lane == 0) {
      uint32_t g = 0, it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const AttnWorkTc w = p.works[item % p.n_works];
        const int kvh = item / p.n_works;
        const int n_a = (w.a_len + kTcKeys - 1) / kTcKeys;
        const int n_chunks = n_a + (w.tok0 + w.n_tok - w.kb0 + kTcKeys - 1) / kTcKeys;
        mbar_wait(bar_q_free, (it & 1u) ^ 1u);
        mbar_arrive_expect_tx(bar_q_full, static_cast<uint32_t>(kSub * G * tpb * 128));
#pragma unroll
        for (int sub = 0; sub < kSub; ++sub) tma_load_3d(s_q + sub * 16384, &tm_q, bar_q_full, sub * 64, kvh * G, w.tok0);
        for (int c = 0; c < n_chunks; ++c, ++g) {
          int nk, key0; bool own;
          chunk_info(w, n_a, c, nk, key0, own);
          const int row = own ? p.b_row0 + key0 + w.b_off : p.a_row0 + w.a_start + key0;
          const uint32_t st = g & 1u, ph = (g >> 1) & 1u;
          mbar_wait(&bar_k_free[st], ph ^ 1u);
          mbar_arrive_expect_tx(&bar_k_full[st], kChunkBytes);
#pragma unroll
          for (int sub = 0; sub < kSub; ++sub) tma_load_2d(s_k + st * kChunkBytes + sub * 8192, own ? &tm_kb : &tm_ka, &bar_k_full[st], kvh * DH + sub * 64, row);
          mbar_wait(&bar_v_free[st], ph ^ 1u);
          mbar_arrive_expect_tx(&bar_v_full[st], kChunkBytes);
#pragma unroll
          for (int sub = 0; sub < kSub; ++sub) tma_load_2d(s_v + st * kChunkBytes + sub * 8192, own ? &tm_vb : &tm_va, &bar_v_full[st], kvh * DH + sub * 64, row);
        }
      }
    }
  } else if (warp == 5) {
    // ===================================================================================== MMA issuer
    if (lane == 0) {
      uint32_t g = 0, it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const AttnWorkTc w = p.works[item % p.n_works];
        const int n_a = (w.a_len + kTcKeys - 1) / kTcKeys;
        const int n_chunks = n_a + (w.tok0 + w.n_tok - w.kb0 + kTcKeys - 1) / kTcKeys;
        auto nk16_of = [&](int c) {
          int nk, key0; bool own;
          chunk_info(w, n_a, c, nk, key0, own);
          return (nk + 15) & ~15;
        };
        auto issue_s = [&](int c) {   // S(c) = Q K(c)^T into accumulator (g + c) & 1
          const uint32_t gg = g + c, st = gg & 1u, ph = (gg >> 1) & 1u;
          mbar_wait(&bar_k_full[st], ph);
          mbar_wait(&bar_s_free[st], ph ^ 1u);
          tc_fence_after();
          const uint32_t idesc = make_idesc_f16kind(128, nk16_of(c), Fmt16<T16>::code, Fmt16<T16>::code, 0);
          const uint32_t k_base = smem_u32(s_k) + st * kChunkBytes;
#pragma unroll
          for (int kk = 0; kk < DH / 16; ++kk) {
            const uint64_t da = make_smem_desc_sw128(smem_u32(s_q) + (kk >> 2) * 16384 + (kk & 3) * 32);
            const uint64_t db = make_smem_desc_sw128(k_base + (kk >> 2) * 8192 + (kk & 3) * 32);
            umma_bf16<1>(tmem + st * 64, da, db, idesc, kk != 0 ? 1u : 0u);
          }
          umma_commit(&bar_s_full[st]);
          umma_commit(&bar_k_free[st]);
          if (c == n_chunks - 1) umma_commit(bar_q_free);
        };
        mbar_wait(bar_q_full, it & 1u);
        issue_s(0);
        for (int c = 0; c < n_chunks; ++c) {
          if (c + 1 < n_chunks) issue_s(c + 1);          // runs while the softmax warps work on S(c)
          const uint32_t gg = g + c, st = gg & 1u, ph = (gg >> 1) & 1u;
          if (c == 0) mbar_wait(bar_o_free, (it & 1u) ^ 1u);   // the previous item's O has been read
          mbar_wait(&bar_v_full[st], ph);
          const int nk16 = nk16_of(c);
          const uint32_t idesc = make_idesc_f16kind(128, DH, Fmt16<T16>::code, Fmt16<T16>::code, 1);
          const uint32_t v_base = smem_u32(s_v) + st * kChunkBytes;
          mbar_wait(bar_p_full, gg & 1u);
          tc_fence_after();
          for (int kk = 0; kk < nk16 / 16; ++kk) {
            const uint64_t da = make_smem_desc_sw128(smem_u32(s_p) + kk * 32);
            const uint64_t db = make_smem_desc_raw(v_base + kk * 2048, 8192, 1024);
            umma_bf16<1>(tmem + 128, da, db, idesc, (c > 0 || kk != 0) ? 1u : 0u);
          }
          umma_commit(&bar_v_free[st]);
          umma_commit(bar_pv_done);
        }
        g += n_chunks;
      }
    }
  } else {
    // ===================================================================================== softmax (one thread per row)
    const int r = tid;                                    // row of the tile = TMEM lane
    const uint32_t t_row = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    const uint32_t t_o = t_row + 128;
    const int tok_local = r / G, head = r - tok_local * G;
    uint32_t g = 0, it = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const AttnWorkTc w = p.works[item % p.n_works];
      const int kvh = item / p.n_works;
      const int n_a = (w.a_len + kTcKeys - 1) / kTcKeys;
      const int n_chunks = n_a + (w.tok0 + w.n_tok - w.kb0 + kTcKeys - 1) / kTcKeys;
      const bool row_ok = tok_local < w.n_tok;
      const int rt = w.tok0 + (row_ok ? tok_local : 0);
      const int seq_lo = (row_ok && p.tok_seq_start != nullptr) ? __ldg(p.tok_seq_start + rt) : 0;
      float m_ref = -INFINITY, l_sum = 0.f;

      for (int c = 0; c < n_chunks; ++c) {
        int nk, key0; bool own;
        chunk_info(w, n_a, c, nk, key0, own);
        const int nk16 = (nk + 15) & ~15;
        const uint32_t gg = g + c, st = gg & 1u, ph = (gg >> 1) & 1u;

        mbar_wait(&bar_s_full[st], ph);
        __syncwarp();   // tcgen05.ld is warp-collective: reconverge after the per-lane spin
        tc_fence_after();
        // ---- this thread's row of S(c): kNE = 32 or 64 elements (chunks of <= 32 keys -- the 14-key prompt root, the tail
        //      of a prefix, a few own tokens -- only pay for half a tile), thread-local max, lazy rescale, P row
        float csum = 0.f;
        auto softmax_chunk = [&](auto ne_tag) {
          constexpr int kNE = decltype(ne_tag)::value;
          float sv[kNE];
          {
            uint32_t raw[kNE];
#pragma unroll
            for (int h = 0; h < kNE / 32; ++h) tmem_ld32(t_row + st * 64 + h * 32, *reinterpret_cast<uint32_t(*)[32]>(raw + h * 32));
            tmem_ld_wait();   // one wait for both loads
#pragma unroll
            for (int i = 0; i < kNE; ++i) sv[i] = __uint_as_float(raw[i]);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_s_free[st]);   // this warp holds its S(c) rows: the accumulator may be overwritten

          float cmax = -INFINITY;
          if (!own && nk == kTcKeys) {   // CTA-uniform fast path: a full chunk of the shared prefix, every key visible
#pragma unroll
            for (int i = 0; i < kNE; ++i) cmax = fmaxf(cmax, sv[i]);
          } else {
            int j_lo = 0, j_hi = nk - 1;
            if (own) {
              j_lo = max(0, seq_lo - key0);
              j_hi = min(nk - 1, rt - key0);
            }
            uint64_t vmask = (j_hi >= j_lo) ? ((~0ull >> (63 - j_hi)) & (~0ull << j_lo)) : 0ull;
            if (own && p.key_valid != nullptr && vmask != 0ull) {
              const uint8_t* kv = p.key_valid + key0;
              for (int i = j_lo; i <= j_hi; ++i)
                if (kv[i] == 0) vmask &= ~(1ull << i);
            }
#pragma unroll
            for (int i = 0; i < kNE; ++i) {
              const float val = ((vmask >> i) & 1ull) ? sv[i] : -INFINITY;
              sv[i] = val;
              cmax = fmaxf(cmax, val);
            }
          }
          const float cmax_s = cmax * p.scale_log2;   // -inf * positive = -inf

          // ---- everything that only needs registers comes first: the reference maximum, the correction factor and the
          //      whole P row are computed while P V(c-1) is still running on the tensor pipe
          float corr = 1.f;
          bool grow = false;
          if (c == 0) {
            m_ref = cmax_s;
          } else if (cmax_s > m_ref + kGrow) {
            grow = true;
            corr = exp2f(m_ref - cmax_s);   // 0 when nothing was visible before
            m_ref = cmax_s;
            l_sum *= corr;
          }
          const float m_use = (m_ref == -INFINITY) ? 0.f : m_ref;
          uint32_t pk[kNE / 2];
#pragma unroll
          for (int j = 0; j < kNE / 2; ++j) {
            const float p0 = fast_exp2(fmaf(sv[2 * j], p.scale_log2, -m_use));
            const float p1 = fast_exp2(fmaf(sv[2 * j + 1], p.scale_log2, -m_use));
            csum += p0 + p1;
            pk[j] = Fmt16<T16>::pack2(p0, p1);
          }
          // ---- O and the P tile belong to P V(c-1) until it completes (for c == 0 the item's epilogue already waited)
          if (c > 0) {
            mbar_wait(bar_pv_done, (gg - 1u) & 1u);
            __syncwarp();
            tc_fence_after();
          }
          if (__any_sync(0xffffffffu, grow)) {   // lazy rescale of this warp's O rows in TMEM
#pragma unroll
            for (int h = 0; h < DH / 32; ++h) {
              uint32_t raw[32];
              tmem_ld32(t_o + h * 32, raw);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) raw[i] = __float_as_uint(__uint_as_float(raw[i]) * corr);
              tmem_st32(t_o + h * 32, raw);
            }
            tmem_st_wait();
          }
#pragma unroll
          for (int j8 = 0; j8 < kNE / 8; ++j8)
            if (j8 * 8 < nk16) *reinterpret_cast<uint4*>(s_p + sw128_offset(r, j8 * 8)) = make_uint4(pk[4 * j8], pk[4 * j8 + 1], pk[4 * j8 + 2], pk[4 * j8 + 3]);
          fence_proxy_async();   // generic-proxy writes of P -> visible to the tensor core
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_p_full);
        };
        if (nk16 > 32) softmax_chunk(KeyCount<64>{}); else softmax_chunk(KeyCount<32>{});   // CTA-uniform
        l_sum += csum;
      }

      // ---- O / l -> 16-bit: the whole row goes to registers first, so O is released to the next item's P V before the
      //      conversion and the global stores
      mbar_wait(bar_pv_done, (g + n_chunks - 1u) & 1u);
      __syncwarp();
      tc_fence_after();
      uint32_t oraw[DH];
#pragma unroll
      for (int h = 0; h < DH / 32; ++h) tmem_ld32(t_o + h * 32, *reinterpret_cast<uint32_t(*)[32]>(oraw + h * 32));
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_o_free);   // the next item's first P V may overwrite O
      const float inv = l_sum > 0.f ? 1.0f / l_sum : 0.f;
      if (row_ok) {
        T16* dst = reinterpret_cast<T16*>(p.o) + static_cast<size_t>(rt) * p.n_q + (kvh * G + head) * DH;
#pragma unroll
        for (int c8 = 0; c8 < DH / 8; ++c8) {
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < 4; ++e)
            pk[e] = Fmt16<T16>::pack2(__uint_as_float(oraw[c8 * 8 + 2 * e]) * inv, __uint_as_float(oraw[c8 * 8 + 2 * e + 1]) * inv);
          *reinterpret_cast<uint4*>(dst + c8 * 8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
      g += n_chunks;
    }
  }

  // ===================================================================================== teardown
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 4) tmem_dealloc<1>(tmem, kTcTmemCols);
}

// Two CTAs per SM (112.3 KB each) is what the kernel is built for; with one it still runs, at half the concurrency.
template <int DH, typename T16>
inline cudaError_t launch_attention_tc5_impl(const CUtensorMap& tm_q, const AttnTcMaps& m, const AttnParamsTc& p, int n_works, int n_kv_heads,
                                             cudaStream_t stream) {
  static bool done[64] = {};
  cudaError_t e = attn_set_smem(attention_tc5_kernel<DH, T16>, attn_tc5_smem_bytes<DH>(), true, done);
  if (e != cudaSuccess) return e;
  AttnParamsTc pp = p;
  pp.n_works = n_works;
  pp.n_kv_heads = n_kv_heads;
  int dev = 0, n_sm = 148;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) n_sm = 148;
  const int items = n_works * n_kv_heads;
  dim3 pgrid(static_cast<unsigned>(std::min(items, 2 * n_sm)));
  attention_tc5_kernel<DH, T16><<<pgrid, kTc5Threads, attn_tc5_smem_bytes<DH>(), stream>>>(tm_q, m.ka, m.va, m.kb, m.vb, pp);
  return cudaGetLastError();
}

template <typename T16>
inline cudaError_t launch_attention_ws(const CUtensorMap& tm_q, const AttnTcMaps& m, const AttnParamsTc& p, int n_works, int n_kv_heads, int head_dim,
                                       cudaStream_t stream) {
  if (n_works <= 0) return cudaSuccess;
  if (p.q_stride != p.n_q || p.group <= 0 || p.group > 128) return cudaErrorInvalidValue;   // Q rows are addressed through the 3-D tensor map
  if (head_dim == 128) return launch_attention_tc5_impl<128, T16>(tm_q, m, p, n_works, n_kv_heads, stream);
  if (head_dim == 64) return launch_attention_tc5_impl<64, T16>(tm_q, m, p, n_works, n_kv_heads, stream);
  return cudaErrorInvalidValue;
}

}  // namespace blim
