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
// Attention is ~0.5 % of the path's FLOPs; this kernel uses warp-level mma.sync (m16n8k16 bf16) with
// fp32 online softmax, one CTA per (sequence, 64-row block of (token, head-in-group) rows, kv head).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace blim {

struct AttnSeq {
  int q_start;  // first token of the sequence in this run (row of Q / O)
  int q_len;
  int a_start;  // first prefix key row in k_a / v_a
  int a_len;
  int b_start;  // first own key row in k_b / v_b (its q_len tokens are consecutive rows)
};
struct AttnWork {
  int seq;
  int row_block;
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

constexpr int kAttnRows = 64;   // query rows per CTA
constexpr int kAttnKeys = 64;   // keys per chunk
constexpr int kAttnThreads = 128;

template <int DH>
constexpr int attn_smem_bytes() { return 3 * kAttnRows * (DH + 8) * 2; }

template <int DH>
__global__ void __launch_bounds__(kAttnThreads) attention_kernel(const AttnParams p) {
  constexpr int LD = DH + 8;  // padded smem row (elements): 16 B shift per row -> conflict-free ldmatrix
  extern __shared__ __align__(16) uint8_t smem_attn[];
  __nv_bfloat16* s_q = reinterpret_cast<__nv_bfloat16*>(smem_attn);
  __nv_bfloat16* s_k = s_q + kAttnRows * LD;
  __nv_bfloat16* s_v = s_k + kAttnKeys * LD;

  const AttnWork w = p.works[blockIdx.x];
  const AttnSeq sq = p.seqs[w.seq];
  const int kvh = blockIdx.y;
  const int G = p.group;
  const int n_rows = sq.q_len * G;
  const int row0 = w.row_block * kAttnRows;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int gq = lane >> 2, tq = lane & 3;

  // ---- stage Q rows (row r -> token r / G, head kvh*G + r % G)
  for (int idx = tid; idx < kAttnRows * (DH / 8); idx += kAttnThreads) {
    const int r = idx / (DH / 8), c = (idx % (DH / 8)) * 8;
    const int gr = row0 + r;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (gr < n_rows) {
      const int tok = sq.q_start + gr / G, head = kvh * G + gr % G;
      val = *reinterpret_cast<const uint4*>(p.q + static_cast<size_t>(tok) * p.n_q + head * DH + c);
    }
    *reinterpret_cast<uint4*>(s_q + r * LD + c) = val;
  }
  __syncthreads();
  uint32_t qf[DH / 16][4];
  {
    const int r = warp * 16 + (lane & 15), coff = (lane >> 4) << 3;
#pragma unroll
    for (int kk = 0; kk < DH / 16; ++kk) {
      const uint32_t addr = static_cast<uint32_t>(__cvta_generic_to_shared(s_q + r * LD + kk * 16 + coff));
      ldsm_x4(addr, qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3]);
    }
  }

  float o[DH / 8][4];
#pragma unroll
  for (int i = 0; i < DH / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

  // token index (inside the sequence) of this thread's two rows
  const int r_lo = row0 + warp * 16 + gq, r_hi = r_lo + 8;
  const int tok_lo = r_lo / G, tok_hi = r_hi / G;
  // causal limit of the block: last token index touched by a valid row
  const int last_row = min(row0 + kAttnRows, n_rows) - 1;
  const int b_len = min(sq.q_len, last_row / G + 1);

  const int n_chunks_a = (sq.a_len + kAttnKeys - 1) / kAttnKeys;
  const int n_chunks_b = (b_len + kAttnKeys - 1) / kAttnKeys;

  for (int ch = 0; ch < n_chunks_a + n_chunks_b; ++ch) {
    const bool seg_b = ch >= n_chunks_a;
    const int k0 = (seg_b ? ch - n_chunks_a : ch) * kAttnKeys;       // first key of the chunk inside its segment
    const int seg_len = seg_b ? b_len : sq.a_len;
    const int nk = min(kAttnKeys, seg_len - k0);
    const __nv_bfloat16* kbase = (seg_b ? p.k_b : p.k_a) + static_cast<size_t>((seg_b ? sq.b_start : sq.a_start) + k0) * p.n_kv + kvh * DH;
    const __nv_bfloat16* vbase = (seg_b ? p.v_b : p.v_a) + static_cast<size_t>((seg_b ? sq.b_start : sq.a_start) + k0) * p.n_kv + kvh * DH;
    __syncthreads();  // previous chunk fully consumed
    for (int idx = tid; idx < kAttnKeys * (DH / 8); idx += kAttnThreads) {
      const int r = idx / (DH / 8), c = (idx % (DH / 8)) * 8;
      uint4 kv = make_uint4(0, 0, 0, 0), vv = make_uint4(0, 0, 0, 0);
      if (r < nk) {
        kv = *reinterpret_cast<const uint4*>(kbase + static_cast<size_t>(r) * p.n_kv + c);
        vv = *reinterpret_cast<const uint4*>(vbase + static_cast<size_t>(r) * p.n_kv + c);
      }
      *reinterpret_cast<uint4*>(s_k + r * LD + c) = kv;
      *reinterpret_cast<uint4*>(s_v + r * LD + c) = vv;
    }
    __syncthreads();

    // ---- S = Q K^T  (16 rows x 64 keys per warp)
    float s[kAttnKeys / 8][4];
#pragma unroll
    for (int i = 0; i < kAttnKeys / 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < DH / 16; ++kk) {
#pragma unroll
      for (int nb = 0; nb < kAttnKeys / 8; nb += 2) {
        const int key = nb * 8 + (lane & 7) + ((lane >> 4) << 3);
        const int col = kk * 16 + (((lane >> 3) & 1) << 3);
        uint32_t b0, b1, b2, b3;
        ldsm_x4(static_cast<uint32_t>(__cvta_generic_to_shared(s_k + key * LD + col)), b0, b1, b2, b3);
        mma_bf16_16816(s[nb], qf[kk], b0, b1);
        mma_bf16_16816(s[nb + 1], qf[kk], b2, b3);
      }
    }

    // ---- mask + online softmax (rows gq and gq+8 of the warp's 16)
    float cmax[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nb = 0; nb < kAttnKeys / 8; ++nb) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int kj = k0 + nb * 8 + 2 * tq + (e & 1);  // key index inside the segment
        const int tok = (e < 2) ? tok_lo : tok_hi;
        bool vis = (kj - k0) < nk;
        if (seg_b) {
          vis = vis && (kj <= tok);
          if (vis && p.key_valid) vis = p.key_valid[sq.b_start + kj] != 0;
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
      m_use[h] = (m_new == -INFINITY) ? 0.f : m_new;  // fully masked so far: keep everything at exp2(-inf) = 0
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

  // ---- normalise and write
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 1);
    l_run[h] += __shfl_xor_sync(0xffffffffu, l_run[h], 2);
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int gr = h == 0 ? r_lo : r_hi;
    if (gr >= n_rows) continue;
    const float inv = l_run[h] > 0.f ? 1.0f / l_run[h] : 0.f;
    const int tok = sq.q_start + gr / G, head = kvh * G + gr % G;
    __nv_bfloat16* dst = p.o + static_cast<size_t>(tok) * p.n_q + head * DH + 2 * tq;
#pragma unroll
    for (int nb = 0; nb < DH / 8; ++nb) {
      __nv_bfloat162 v2 = __floats2bfloat162_rn(o[nb][2 * h] * inv, o[nb][2 * h + 1] * inv);
      *reinterpret_cast<__nv_bfloat162*>(dst + nb * 8) = v2;
    }
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
