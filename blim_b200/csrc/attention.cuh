// Varlen causal GQA attention with a shared (cascade) prefix segment.
//
// Replaces Qwen2SdpaAttention's softmax(QK^T/sqrt(d) + mask)V (reference: modeling_qwen2_flash.py:685-709) and repeat_kv
// (modeling_qwen2_flash.py:192-201, never materialised here: the G query heads of a KV group are stacked along the row
// dimension of one tile).  Every sequence sees
//   segment A: a_len keys of a previously prefilled prefix (all visible) -- the video prefix shared by all captions of a
//              video (VTG), the text prefix shared by all candidate videos of a text (TVG), or the CPN-visible header;
//   segment B: its own q_len tokens, key j visible to query i iff j <= i and key_valid[j]
// which is exactly "causal AND key-valid" of the reference's 4-D additive mask (modeling_qwen2_flash.py:1019-1040) once
// the invisible tokens are dropped from the layout.  Rotary positions were applied by the QKV epilogue, so position
// gaps (CPN) need nothing here.
//
// Work decomposition: consecutive sequences that share the same prefix form a GROUP; the (token, head-in-group) rows of
// all sequences of a group are stacked and cut into 64-row blocks, one CTA per (block, kv head).  The shared prefix K/V
// is therefore streamed once per 64 stacked rows instead of once per sequence (a caption suffix alone only fills
// ~13 tokens x 7 heads = 91 rows).  Own-segment keys are visited sequence by sequence; warps whose rows do not belong to
// the sequence skip the chunk.  K/V chunks of 64 keys are double-buffered with cp.async.
//
// Attention is ~0.5 % of the path's FLOPs; the math runs on warp-level mma.sync (m16n8k16 bf16) with
// fp32 online softmax.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

namespace blim {

struct AttnSeq {
  int q_start;  // first token of the sequence in this run (row of Q / O)
  int q_len;
  int a_start;  // first prefix key row in k_a / v_a
  int a_len;
  int b_start;  // first own key row in k_b / v_b (its q_len tokens are consecutive rows)
  int unit;     // scheduling unit (video / text) the sequence belongs to: attention tiles never stack sequences of different
                // units, so a unit's numbers do not depend on which other units share its run (multi-GPU sharding, batching)
};
// 64 stacked rows of the sequences seq_first .. seq_first + n_seq - 1 (all sharing a_start / a_len); the block starts at
// row `row_first` of sequence seq_first (row = token * G + head_in_group).
struct AttnWork {
  int seq_first;
  int n_seq;
  int row_first;
  int pad;
};
struct AttnParams {
  const __nv_bfloat16* q;  // [T, n_q]
  __nv_bfloat16* o;        // [T, n_q]
  const __nv_bfloat16* k_a;
  const __nv_bfloat16* v_a;
  const __nv_bfloat16* k_b;
  const __nv_bfloat16* v_b;
  const uint8_t* key_valid;  // per own key row (indexed like k_b rows); nullptr = all valid
  const AttnSeq* seqs;
  const AttnWork* works;
  int n_q, n_kv, group;
  float scale_log2;  // log2(e) / sqrt(head_dim)
};

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// 16-byte async copy global -> shared; src_bytes = 0 zero-fills the destination.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))), "l"(gsrc),
               "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

constexpr int kAttnRows = 64;     // stacked query rows per CTA
constexpr int kAttnKeys = 64;     // keys per chunk
constexpr int kAttnThreads = 128;
constexpr int kAttnMaxSeq = 64;   // a 64-row block overlaps at most 64 sequences

template <int DH>
constexpr int attn_smem_bytes() { return 2 * 2 * kAttnKeys * (DH + 8) * 2; }  // 2 stages x (K, V)

template <int DH>
__global__ void __launch_bounds__(kAttnThreads, 3) attention_kernel(const AttnParams p) {
  constexpr int LD = DH + 8;  // padded smem row (elements): 16 B shift per row -> conflict-free ldmatrix
  constexpr int kStageElems = 2 * kAttnKeys * LD;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  __nv_bfloat16* s_kv = reinterpret_cast<__nv_bfloat16*>(smem_attn);
  __shared__ int s_row0[kAttnMaxSeq + 1];    // block-relative first row of sequence j (may be negative for j = 0)
  __shared__ int s_rows[kAttnMaxSeq];        // q_len * G
  __shared__ int s_blen[kAttnMaxSeq];        // own keys the block needs from sequence j (causal limit)
  __shared__ int s_choff[kAttnMaxSeq + 1];   // prefix sum of own-segment chunk counts
  __shared__ int s_bstart[kAttnMaxSeq];
  __shared__ int s_qstart[kAttnMaxSeq];
  __shared__ short s_row_seq[kAttnRows];     // sequence (block-relative) of each row, -1 = padding
  __shared__ short s_row_tok[kAttnRows];     // token index inside its sequence
  __shared__ short s_row_head[kAttnRows];    // head inside the KV group

  const AttnWork w = p.works[blockIdx.x];
  const int kvh = blockIdx.y;
  const int G = p.group;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gq = lane >> 2, tq = lane & 3;
  const AttnSeq sq0 = p.seqs[w.seq_first];
  const int a_start = sq0.a_start, a_len = sq0.a_len;

  if (tid == 0) {
    int row = -w.row_first, off = 0;
    for (int j = 0; j < w.n_seq; ++j) {
      const AttnSeq s = p.seqs[w.seq_first + j];
      const int rows = s.q_len * G;
      s_row0[j] = row;
      s_rows[j] = rows;
      s_bstart[j] = s.b_start;
      s_qstart[j] = s.q_start;
      const int last_local = min(kAttnRows, row + rows) - 1 - row;  // last row of sequence j inside the block
      const int bl = min(s.q_len, last_local / G + 1);
      s_blen[j] = bl;
      s_choff[j] = off;
      off += (bl + kAttnKeys - 1) / kAttnKeys;
      row += rows;
    }
    s_row0[w.n_seq] = row;
    s_choff[w.n_seq] = off;
  }
  __syncthreads();
  if (tid < kAttnRows) {
    int j = -1;
    for (int c = 0; c < w.n_seq; ++c)
      if (tid >= s_row0[c] && tid < s_row0[c] + s_rows[c]) j = c;
    s_row_seq[tid] = static_cast<short>(j);
    const int lr = j >= 0 ? tid - s_row0[j] : 0;
    s_row_tok[tid] = static_cast<short>(lr / G);
    s_row_head[tid] = static_cast<short>(lr % G);
  }
  __syncthreads();

  const int n_a = (a_len + kAttnKeys - 1) / kAttnKeys;
  const int n_chunks = n_a + s_choff[w.n_seq];

  // chunk c -> (segment, sequence j, first key k0, key count nk, K/V base pointers)
  auto chunk_info = [&](int c, int& j, int& k0, int& nk, const __nv_bfloat16*& kb, const __nv_bfloat16*& vb) {
    if (c < n_a) {
      j = -1;
      k0 = c * kAttnKeys;
      nk = min(kAttnKeys, a_len - k0);
      const size_t off = static_cast<size_t>(a_start + k0) * p.n_kv + kvh * DH;
      kb = p.k_a + off;
      vb = p.v_a + off;
    } else {
      const int cb = c - n_a;
      j = 0;
      while (cb >= s_choff[j + 1]) ++j;
      k0 = (cb - s_choff[j]) * kAttnKeys;
      nk = min(kAttnKeys, s_blen[j] - k0);
      const size_t off = static_cast<size_t>(s_bstart[j] + k0) * p.n_kv + kvh * DH;
      kb = p.k_b + off;
      vb = p.v_b + off;
    }
  };
  auto issue_load = [&](int c) {
    int j, k0, nk;
    const __nv_bfloat16 *kb, *vb;
    chunk_info(c, j, k0, nk, kb, vb);
    __nv_bfloat16* sk = s_kv + (c & 1) * kStageElems;
    __nv_bfloat16* sv = sk + kAttnKeys * LD;
    for (int idx = tid; idx < kAttnKeys * (DH / 8); idx += kAttnThreads) {
      const int r = idx / (DH / 8), cc = (idx % (DH / 8)) * 8;
      const bool ok = r < nk;
      const size_t go = ok ? static_cast<size_t>(r) * p.n_kv + cc : 0;
      cp_async16(sk + r * LD + cc, kb + go, ok ? 16 : 0);
      cp_async16(sv + r * LD + cc, vb + go, ok ? 16 : 0);
    }
    cp_async_commit();
  };

  if (n_chunks > 0) issue_load(0);

  // ---- this thread's two rows (gq and gq + 8 of the warp's 16) and its Q fragments, straight from global memory
  const int lr_lo = warp * 16 + gq, lr_hi = lr_lo + 8;
  const int seq_lo = s_row_seq[lr_lo], seq_hi = s_row_seq[lr_hi];
  const int tok_lo = s_row_tok[lr_lo], tok_hi = s_row_tok[lr_hi];
  const __nv_bfloat16* q_lo = nullptr;
  const __nv_bfloat16* q_hi = nullptr;
  if (seq_lo >= 0) q_lo = p.q + static_cast<size_t>(s_qstart[seq_lo] + tok_lo) * p.n_q + (kvh * G + s_row_head[lr_lo]) * DH;
  if (seq_hi >= 0) q_hi = p.q + static_cast<size_t>(s_qstart[seq_hi] + tok_hi) * p.n_q + (kvh * G + s_row_head[lr_hi]) * DH;
  uint32_t qf[DH / 16][4];
#pragma unroll
  for (int kk = 0; kk < DH / 16; ++kk) {
    const int c0 = kk * 16 + 2 * tq;
    qf[kk][0] = q_lo ? *reinterpret_cast<const uint32_t*>(q_lo + c0) : 0u;
    qf[kk][1] = q_hi ? *reinterpret_cast<const uint32_t*>(q_hi + c0) : 0u;
    qf[kk][2] = q_lo ? *reinterpret_cast<const uint32_t*>(q_lo + c0 + 8) : 0u;
    qf[kk][3] = q_hi ? *reinterpret_cast<const uint32_t*>(q_hi + c0 + 8) : 0u;
  }
  // sequences / tokens covered by this warp's 16 rows (for skipping own-segment chunks)
  int w_seq_min = 1 << 20, w_seq_max = -1;
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    const int sj = s_row_seq[warp * 16 + r];
    if (sj >= 0) { w_seq_min = min(w_seq_min, sj); w_seq_max = max(w_seq_max, sj); }
  }

  float o[DH / 8][4];
#pragma unroll
  for (int i = 0; i < DH / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

  for (int c = 0; c < n_chunks; ++c) {
    if (c + 1 < n_chunks) {
      issue_load(c + 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    int j, k0, nk;
    const __nv_bfloat16 *kb_unused, *vb_unused;
    chunk_info(c, j, k0, nk, kb_unused, vb_unused);
    const __nv_bfloat16* s_k = s_kv + (c & 1) * kStageElems;
    const __nv_bfloat16* s_v = s_k + kAttnKeys * LD;
    // warp-uniform skip: prefix chunks need at least one real row, own chunks a row of sequence j that can see key k0
    bool active = w_seq_max >= 0;
    if (j >= 0) {
      active = (j >= w_seq_min && j <= w_seq_max);
      if (active) {
        const int last_local = min(warp * 16 + 16, s_row0[j] + s_rows[j]) - 1 - s_row0[j];
        active = last_local >= 0 && (last_local / G) >= k0;
      }
    }
    if (active) {
      // ---- S = Q K^T  (16 rows x 64 keys per warp)
      // short chunks (own segment of a caption, tail of the prefix): key blocks beyond nk are skipped (warp-uniform)
      const int nb_lim = (nk + 15) >> 4 << 1;  // 8-key blocks to compute, rounded to the 16-key ldmatrix granularity
      float s[kAttnKeys / 8][4];
#pragma unroll
      for (int i = 0; i < kAttnKeys / 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < DH / 16; ++kk) {
#pragma unroll
        for (int nb = 0; nb < kAttnKeys / 8; nb += 2) {
          if (nb >= nb_lim) continue;
          const int key = nb * 8 + (lane & 7) + ((lane >> 4) << 3);
          const int col = kk * 16 + (((lane >> 3) & 1) << 3);
          uint32_t b0, b1, b2, b3;
          ldsm_x4(static_cast<uint32_t>(__cvta_generic_to_shared(s_k + key * LD + col)), b0, b1, b2, b3);
          mma_bf16_16816(s[nb], qf[kk], b0, b1);
          mma_bf16_16816(s[nb + 1], qf[kk], b2, b3);
        }
      }
      // ---- mask + online softmax
      float cmax[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int nb = 0; nb < kAttnKeys / 8; ++nb) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int kl = nb * 8 + 2 * tq + (e & 1);  // key inside the chunk
          const int rseq = (e < 2) ? seq_lo : seq_hi;
          bool vis = kl < nk && rseq >= 0;
          if (j >= 0) {
            const int tok = (e < 2) ? tok_lo : tok_hi;
            vis = vis && rseq == j && (k0 + kl) <= tok;
            if (vis && p.key_valid) vis = p.key_valid[s_bstart[j] + k0 + kl] != 0;
          }
          const float val = vis ? s[nb][e] * p.scale_log2 : -INFINITY;
          s[nb][e] = val;
          cmax[e >> 1] = fmaxf(cmax[e >> 1], val);
        }
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        cmax[h] = fmaxf(cmax[h], __shfl_xor_sync(0xffffffffu, cmax[h], 1));
        cmax[h] = fmaxf(cmax[h], __shfl_xor_sync(0xffffffffu, cmax[h], 2));
      }
      float corr[2], m_use[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float m_new = fmaxf(m_run[h], cmax[h]);
        m_use[h] = (m_new == -INFINITY) ? 0.f : m_new;  // nothing visible so far: everything stays exp2(-inf) = 0
        corr[h] = exp2f(m_run[h] - m_use[h]);
        m_run[h] = m_new;
      }
      float csum[2] = {0.f, 0.f};
#pragma unroll
      for (int nb = 0; nb < kAttnKeys / 8; ++nb) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float pv = exp2f(s[nb][e] - m_use[e >> 1]);
          s[nb][e] = pv;
          csum[e >> 1] += pv;
        }
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) l_run[h] = l_run[h] * corr[h] + csum[h];
#pragma unroll
      for (int i = 0; i < DH / 8; ++i) {
        o[i][0] *= corr[0]; o[i][1] *= corr[0];
        o[i][2] *= corr[1]; o[i][3] *= corr[1];
      }
      // ---- O += P V
#pragma unroll
      for (int kk = 0; kk < kAttnKeys / 16; ++kk) {
        if (2 * kk >= nb_lim) continue;
        uint32_t pa[4];
        {
          __nv_bfloat162 t0 = __floats2bfloat162_rn(s[2 * kk][0], s[2 * kk][1]);
          __nv_bfloat162 t1 = __floats2bfloat162_rn(s[2 * kk][2], s[2 * kk][3]);
          __nv_bfloat162 t2 = __floats2bfloat162_rn(s[2 * kk + 1][0], s[2 * kk + 1][1]);
          __nv_bfloat162 t3 = __floats2bfloat162_rn(s[2 * kk + 1][2], s[2 * kk + 1][3]);
          pa[0] = *reinterpret_cast<uint32_t*>(&t0);
          pa[1] = *reinterpret_cast<uint32_t*>(&t1);
          pa[2] = *reinterpret_cast<uint32_t*>(&t2);
          pa[3] = *reinterpret_cast<uint32_t*>(&t3);
        }
#pragma unroll
        for (int nb = 0; nb < DH / 8; nb += 2) {
          const int key = kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
          const int col = nb * 8 + ((lane >> 4) << 3);
          uint32_t b0, b1, b2, b3;
          ldsm_x4_t(static_cast<uint32_t>(__cvta_generic_to_shared(s_v + key * LD + col)), b0, b1, b2, b3);
          mma_bf16_16816(o[nb], pa, b0, b1);
          mma_bf16_16816(o[nb + 1], pa, b2, b3);
        }
      }
    }
    __syncthreads();  // everyone is done with this stage before the next iteration's load overwrites it
  }

  // ---- normalise and write
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 1);
    l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 2);
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int rseq = h == 0 ? seq_lo : seq_hi;
    if (rseq < 0) continue;
    const int lr = h == 0 ? lr_lo : lr_hi;
    const float inv = l_run[h] > 0.f ? 1.0f / l_run[h] : 0.f;
    const int tok = s_qstart[rseq] + (h == 0 ? tok_lo : tok_hi);
    __nv_bfloat16* dst = p.o + static_cast<size_t>(tok) * p.n_q + (kvh * G + s_row_head[lr]) * DH + 2 * tq;
#pragma unroll
    for (int nb = 0; nb < DH / 8; ++nb) {
      __nv_bfloat162 v2 = __floats2bfloat162_rn(o[nb][2 * h] * inv, o[nb][2 * h + 1] * inv);
      *reinterpret_cast<__nv_bfloat162*>(dst + nb * 8) = v2;
    }
  }
}

// Host: cut the stacked rows of every prefix-sharing group of consecutive sequences into 64-row blocks.
inline void build_attn_works(const AttnSeq* seqs, int n_seqs, int group, std::vector<AttnWork>& works) {
  works.clear();
  int s0 = 0;
  while (s0 < n_seqs) {
    int s1 = s0 + 1;
    if (seqs[s0].a_len > 0)
      while (s1 < n_seqs && seqs[s1].a_len == seqs[s0].a_len && seqs[s1].a_start == seqs[s0].a_start) ++s1;
    // rows of the group
    long long total = 0;
    for (int s = s0; s < s1; ++s) total += static_cast<long long>(seqs[s].q_len) * group;
    int cur = s0;                 // sequence containing the block's first row
    long long cur_row0 = 0;       // group row of cur's first row
    for (long long r = 0; r < total; r += kAttnRows) {
      while (cur_row0 + static_cast<long long>(seqs[cur].q_len) * group <= r) {
        cur_row0 += static_cast<long long>(seqs[cur].q_len) * group;
        ++cur;
      }
      AttnWork w;
      w.seq_first = cur;
      w.row_first = static_cast<int>(r - cur_row0);
      int n = 0;
      long long row = cur_row0;
      for (int s = cur; s < s1 && row < r + kAttnRows; ++s) {
        row += static_cast<long long>(seqs[s].q_len) * group;
        ++n;
      }
      w.n_seq = n;
      w.pad = 0;
      works.push_back(w);
    }
    s0 = s1;
  }
}

inline cudaError_t launch_attention(const AttnParams& p, int n_works, int n_kv_heads, int head_dim, cudaStream_t stream) {
  if (n_works <= 0) return cudaSuccess;
  dim3 grid(static_cast<unsigned>(n_works), static_cast<unsigned>(n_kv_heads));
  if (head_dim == 128) {
    static bool set = false;
    if (!set) {
      cudaError_t e = cudaFuncSetAttribute(attention_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_smem_bytes<128>());
      if (e != cudaSuccess) return e;
      set = true;
    }
    attention_kernel<128><<<grid, kAttnThreads, attn_smem_bytes<128>(), stream>>>(p);
  } else if (head_dim == 64) {
    static bool set = false;
    if (!set) {
      cudaError_t e = cudaFuncSetAttribute(attention_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_smem_bytes<64>());
      if (e != cudaSuccess) return e;
      set = true;
    }
    attention_kernel<64><<<grid, kAttnThreads, attn_smem_bytes<64>(), stream>>>(p);
  } else {
    return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace blim
