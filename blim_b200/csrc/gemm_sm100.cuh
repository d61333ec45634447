// tcgen05 / TMEM / TMA GEMM core for sm_100a:  C[M,N] = A[M,K] · W[N,K]^T  (both operands K-major bf16, fp32 accumulate in TMEM)
// with pluggable epilogues.  This one mainloop serves every dense contraction of the BLiM scoring path:
//   projector MLPs            (reference: videochat_flash/mm_projector_builder.py:156-159)  -> EpiStore (bias, bias+GELU)
//   fused QKV + bias + RoPE   (reference: modeling_qwen2_flash.py:666-679, 147-172)         -> EpiQkvRope
//   o_proj / down_proj + res. (reference: modeling_qwen2_flash.py:714, 784, 790)            -> EpiResid
//   gate|up + SwiGLU          (reference: modeling_qwen2_flash.py:188)                      -> EpiSwiglu
//   LM head + log-softmax     (reference: modeling_qwen2_flash.py:1452-1453 + retrieval_utils.py:23-33) -> EpiLse
//   TVG head                  (reference: retrieval_utils.py:104-107)                       -> EpiStore + EpiLse
//
// Kernel anatomy (one persistent CTA, or CTA pair, per SM):
//   warp 0   TMA producer      : cp.async.bulk.tensor (SWIZZLE_128B) -> kStages-deep smem ring, mbarrier full/empty
//   warp 1   MMA issuer        : one elected thread issues tcgen05.mma (M=128·cta_group, N=256, K=16), tcgen05.commit frees slots
//   warp 2   TMEM allocator    : 512 columns = 2 accumulator stages of 256 fp32 columns
//   warps 4-11 epilogue        : tcgen05.ld (one TMEM lane = one output row per thread) -> fused epilogue -> global;
//                                two warpgroups, each owns half of the tile's columns (TMEM lane quadrant = warp % 4)
// The accumulator is double-buffered in TMEM so the epilogue of tile i overlaps the mainloop of tile i+1.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "act_type.cuh"
#include "ptx_sm100.cuh"

namespace blim {

constexpr int kBM = 128;   // rows per CTA
constexpr int kBN = 256;   // columns per tile (= one UMMA N)
constexpr int kBK = 64;    // K per pipeline stage: 64 bf16 = 128 B = one swizzle row
constexpr int kUmmaK = 16;
constexpr int kGemmThreads = 384;  // 4 control warps + 8 epilogue warps
constexpr int kTmemCols = 512;

template <int kCtaGroup>
struct GemmCfg {
  static constexpr int kStages = kCtaGroup == 1 ? 4 : 6;
  static constexpr int kBRows = kBN / kCtaGroup;  // B rows staged by each CTA
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = kBRows * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarBytes = 256;
  // per epilogue warp a 4 KB slab, 1024-byte aligned: 32 rows x (64 B + 16 B pad) for the coalesced 16-bit / fp32 stores, or a
  // 128-byte-swizzled 32 x 32 fp32 tile that a TMA reduce adds to the residual stream (EpiResidTma)
  static constexpr int kOutWarpBytes = 4096;
  static constexpr int kOutStageBytes = 8 * kOutWarpBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + kOutStageBytes + kBarBytes + 1024;
};

struct GemmDims {
  int M, N, K;
  int m_tiles;   // number of (kBM*cta_group)-row tiles
  int n_tiles;   // number of kBN-column tiles
  int sb_tiles;  // m-tiles per super-block (L2 blocking of the A operand)
  int l2_hints;  // 1: A (the super-block slab, reused by every n-tile) is loaded evict_last, W (streamed once per super-block) evict_first
  int a_fmt, b_fmt;  // 16-bit formats of the operands (kFmtF16 / kFmtBF16), independent of each other (act_type.cuh)
};

// tile t -> (m_tile, n_tile): super-blocks of sb_tiles m-tiles; inside a super-block m runs fastest so the CTAs that run
// concurrently share a handful of W tiles while the A super-block stays L2-resident.
__device__ __forceinline__ void tile_coords(const GemmDims& d, int t, int& m_tile, int& n_tile) {
  const int per_sb = d.sb_tiles * d.n_tiles;
  const int sb = t / per_sb;
  const int base_m = sb * d.sb_tiles;
  const int msz = min(d.sb_tiles, d.m_tiles - base_m);
  const int r = t - sb * per_sb;
  n_tile = r / msz;
  m_tile = base_m + (r - n_tile * msz);
}

// ------------------------------------------------------------------------------------------------ epilogue helpers
// Store 32 consecutive 16-bit values (T = __nv_bfloat16 or __half) of one row (this thread's row).  `valid` = number of
// in-range columns (<= 32).
template <typename T>
__device__ __forceinline__ void store_row32_16(T* dst, const float (&v)[32], int valid) {
  if (valid >= 32) {
    uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 u;
      u.x = Fmt16<T>::pack2(v[8 * i + 0], v[8 * i + 1]);
      u.y = Fmt16<T>::pack2(v[8 * i + 2], v[8 * i + 3]);
      u.z = Fmt16<T>::pack2(v[8 * i + 4], v[8 * i + 5]);
      u.w = Fmt16<T>::pack2(v[8 * i + 6], v[8 * i + 7]);
      d4[i] = u;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < valid) dst[i] = Fmt16<T>::from_float(v[i]);
  }
}
__device__ __forceinline__ void store_row32_f32(float* dst, const float (&v)[32], int valid) {
  if (valid >= 32) {
    float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
    for (int i = 0; i < 8; ++i) d4[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else {
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < valid) dst[i] = v[i];
  }
}
__device__ __forceinline__ void tmem_ld32f(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  tmem_ld32(taddr, r);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// Exact-erf GELU (nn.GELU default) with erf from Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7, far below the bf16 rounding
// of the stored activation): one MUFU.RCP, one MUFU.EX2 and ~12 FMA-pipe instructions instead of erff's ~40 with branches --
// at K = 1024 (the ViT's fc1) the epilogue of a 256-column tile otherwise takes longer than its mainloop.
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));   // one MUFU.RCP (1 ulp), not the IEEE-rounded software path
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * z * z));
  const float h = 0.5f * p * t * e;          // 0.5 * erfc(|z|) = Phi(-|x|)
  return x * (x >= 0.f ? 1.0f - h : h);      // x * Phi(x)
}

// out[row, col .. col+31] = bf16(v) for the 32 rows of a warp (lane = row).  Row-per-thread stores scatter every warp
// instruction over 32 rows (32 half-used sectors); the warp's 32 x 64 B go through its shared-memory slab instead and
// are stored 8 rows x 64 contiguous bytes per instruction.  r0 = first row of the warp.
template <typename T>
__device__ __forceinline__ void store_tile32_16_staged(T* __restrict__ out, int ldo, int r0, int col, const float (&v)[32], int M,
                                                       uint8_t* wstage) {
  const int lane = threadIdx.x & 31;
  uint4* srow = reinterpret_cast<uint4*>(wstage + lane * 80);
#pragma unroll
  for (int i = 0; i < 4; ++i)
    srow[i] = make_uint4(Fmt16<T>::pack2(v[8 * i + 0], v[8 * i + 1]), Fmt16<T>::pack2(v[8 * i + 2], v[8 * i + 3]),
                         Fmt16<T>::pack2(v[8 * i + 4], v[8 * i + 5]), Fmt16<T>::pack2(v[8 * i + 6], v[8 * i + 7]));
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int rr = (lane >> 2) + 8 * j;
    const uint4 val = *reinterpret_cast<const uint4*>(wstage + rr * 80 + (lane & 3) * 16);
    if (r0 + rr < M) *reinterpret_cast<uint4*>(out + static_cast<size_t>(r0 + rr) * ldo + col + (lane & 3) * 8) = val;
  }
  __syncwarp();
}

// Every epilogue gets: this thread's global row (may be >= M: loads from TMEM still have to be executed warp-uniformly,
// only the global-memory side is predicated), the tile's first column n0, the TMEM address of (its lane, column 0
// of the accumulator stage) and `half` (0/1): which half of the tile's work this epilogue warpgroup owns.

// out[row, col] = act(acc + bias[col])   OutT = __nv_bfloat16, __half or float
template <typename OutT, bool kBias, bool kGelu>
struct EpiStore {
  static constexpr int kGroups = 2;
  struct Params {
    OutT* out;
    int ldo;
    const float* bias;
  };
  __device__ static void run(const Params& p, const GemmDims& d, int row, int n0, int n_tile, uint32_t taddr, int half, uint8_t* wstage) {
    const bool row_ok = row < d.M;
#pragma unroll 1
    for (int c = half * (kBN / 2); c < (half + 1) * (kBN / 2); c += 32) {
      float v[32];
      tmem_ld32f(taddr + c, v);
      const int col = n0 + c;
      const int valid = d.N - col;
      if (valid <= 0) continue;  // warp-uniform
      if constexpr (kBias) {
        if (valid >= 32) {   // warp-uniform: whole 32-column step in range
          const float4* b4 = reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 b = __ldg(b4 + i);
            v[4 * i] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < valid) v[i] += __ldg(p.bias + col + i);
        }
      }
      if constexpr (kGelu) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
      }
      if constexpr (sizeof(OutT) == 2) {
        if (valid >= 32) {   // warp-uniform
          store_tile32_16_staged<OutT>(p.out, p.ldo, row - static_cast<int>(threadIdx.x & 31), col, v, d.M, wstage);
        } else if (row_ok) {
          store_row32_16<OutT>(p.out + static_cast<size_t>(row) * p.ldo + col, v, valid);
        }
      } else if (row_ok) {
        store_row32_f32(reinterpret_cast<float*>(p.out) + static_cast<size_t>(row) * p.ldo + col, v, valid);
      }
    }
  }
};

// resid[row, col .. col+31] += v for the 32 rows of a warp (lane = row), staged through the warp's shared-memory slab so that
// every global instruction touches 8 rows x 64 contiguous bytes (full sectors) instead of 32 rows x 16 bytes: the fp32
// residual read-modify-write is what bounds the short-K GEMMs (o_proj, the ViT's proj).  r0 = first row of the warp.
__device__ __forceinline__ void resid_add_staged(float* __restrict__ resid, int ldo, int r0, int col, const float (&v)[32], int M, uint8_t* wstage) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    float4* srow = reinterpret_cast<float4*>(wstage + lane * 80);
#pragma unroll
    for (int i = 0; i < 4; ++i) srow[i] = make_float4(v[16 * h + 4 * i], v[16 * h + 4 * i + 1], v[16 * h + 4 * i + 2], v[16 * h + 4 * i + 3]);
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int rr = (lane >> 2) + 8 * j;
      const float4 a = *reinterpret_cast<const float4*>(wstage + rr * 80 + (lane & 3) * 16);
      if (r0 + rr < M) {
        float4* g = reinterpret_cast<float4*>(resid + static_cast<size_t>(r0 + rr) * ldo + col + 16 * h + (lane & 3) * 4);
        float4 r = *g;
        r.x += a.x; r.y += a.y; r.z += a.z; r.w += a.w;
        *g = r;
      }
    }
    __syncwarp();
  }
}

// resid[row, col] += acc     (fp32 residual stream, in place).  kGroups = epilogue warpgroups that share a tile.
template <int kGroupsT>
struct EpiResidT {
  static constexpr int kGroups = kGroupsT;
  struct Params {
    float* resid;
    int ldo;
  };
  __device__ static void run(const Params& p, const GemmDims& d, int row, int n0, int n_tile, uint32_t taddr, int half, uint8_t* wstage) {
    const bool row_ok = row < d.M;
#pragma unroll 1
    for (int c = half * (kBN / kGroups); c < (half + 1) * (kBN / kGroups); c += 32) {
      float v[32];
      tmem_ld32f(taddr + c, v);
      const int col = n0 + c;
      const int valid = d.N - col;
      if (valid <= 0) continue;
      if (valid >= 32) {   // warp-uniform
        resid_add_staged(p.resid, p.ldo, row - static_cast<int>(threadIdx.x & 31), col, v, d.M, wstage);
      } else if (row_ok) {
        float* dst = p.resid + static_cast<size_t>(row) * p.ldo + col;
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i < valid) dst[i] += v[i];
      }
    }
  }
};

using EpiResid = EpiResidT<2>;

// A/B variant (BLIM_RESID_TMA=1|2, off by default): resid[row, col] += acc through the TMA -- the warp's 32 x 32 fp32 piece
// is written to its shared-memory slab (128-byte swizzle) and ONE cp.reduce.async.bulk.tensor (.add, fp32) hands it to
// the L2, which performs the read-modify-write, so the epilogue warps never wait for residual rows to arrive.  Every
// element receives exactly one add per GEMM: the same IEEE sum as the in-register version (the GPU tests pass with it).
// Measured on the same box (profiles/r02_bench_c2_ab_epilogues_same_box_v22.log): o_proj 393-396 -> 399 ms per C2 step, i.e.
// no gain -- with the staged read-modify-write the epilogue already hides under the next tile's mainloop, so the in-register
// version stays the default.  Needs N % 32 == 0; rows beyond M are clipped by the tensor map.
struct EpiResidTma {
  static constexpr int kGroups = 2;
  struct Params {
    CUtensorMap tm_x;   // fp32 [M, N] residual stream, box = 32 columns x 32 rows, SWIZZLE_128B
  };
  __device__ static void run(const Params& p, const GemmDims& d, int row, int n0, int n_tile, uint32_t taddr, int half, uint8_t* wstage) {
    const int lane = threadIdx.x & 31;
    const int r0 = row - lane;
#pragma unroll 1
    for (int c = half * (kBN / 2); c < (half + 1) * (kBN / 2); c += 32) {
      float v[32];
      tmem_ld32f(taddr + c, v);
      const int col = n0 + c;
      if (col >= d.N) continue;   // warp-uniform
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the previous reduce has read the slab
      __syncwarp();
      float4* srow = reinterpret_cast<float4*>(wstage + lane * 128);
#pragma unroll
      for (int i = 0; i < 8; ++i) srow[i ^ (lane & 7)] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0 && r0 < d.M) {
        asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                     ::"l"(reinterpret_cast<uint64_t>(&p.tm_x)), "r"(smem_u32(wstage)), "r"(col), "r"(r0)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the slab outlives every read of it
    __syncwarp();
  }
};

// resid[row, col] += acc + bias[col]   (ViT blocks: attn.proj / mlp.fc2 carry a bias, vision_tower_builder.py:96,53)
struct EpiResidBias {
  static constexpr int kGroups = 2;
  struct Params {
    float* resid;
    int ldo;
    const float* bias;
  };
  __device__ static void run(const Params& p, const GemmDims& d, int row, int n0, int n_tile, uint32_t taddr, int half, uint8_t* wstage) {
    const bool row_ok = row < d.M;
#pragma unroll 1
    for (int c = half * (kBN / kGroups); c < (half + 1) * (kBN / kGroups); c += 32) {
      float v[32];
      tmem_ld32f(taddr + c, v);
      const int col = n0 + c;
      const int valid = d.N - col;
      if (valid <= 0) continue;
      if (valid >= 32) {   // warp-uniform
        const float4* b4 = reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 b = __ldg(b4 + i);
          v[4 * i] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
        }
        resid_add_staged(p.resid, p.ldo, row - static_cast<int>(threadIdx.x & 31), col, v, d.M, wstage);
      } else if (row_ok) {
        float* dst = p.resid + static_cast<size_t>(row) * p.ldo + col;
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i < valid) dst[i] += v[i] + __ldg(p.bias + col + i);
      }
    }
  }
};

// Residual add fused with the FIRST half of the following RMSNorm (reference: Qwen2DecoderLayer q2:784-789 + Qwen2RMSNorm
// q2:93-98): x_new = resid + acc is written back in fp32, its bf16 copy `xb` becomes the A operand of the next GEMM
// (un-normalised: the norm weight is folded into that GEMM's weight columns and 1/rms is applied per row in its
// epilogue), and the partial sum of squares of this thread's 128 columns goes to ssq[row, 2 * n_tile + half].
// The partials are reduced in a fixed order (rstd_rows_kernel), so results stay bit-identical under re-batching.
struct EpiResidNorm {
  static constexpr int kGroups = 2;
  struct Params {
    float* resid;          // [M, N] in / out
    act_t* xb;             // [M, N] out
    float* ssq;            // [M, 2 * n_tiles] out
    int ldo;
  };
  __device__ static void run(const Params& p, const GemmDims& d, int row, int n0, int n_tile, uint32_t taddr, int half, uint8_t* wstage) {
    const bool row_ok = row < d.M;
    float ss = 0.f;
#pragma unroll 1
    for (int c = half * (kBN / 2); c < (half + 1) * (kBN / 2); c += 32) {
      float v[32];
      tmem_ld32f(taddr + c, v);
      const int col = n0 + c;
      const int valid = d.N - col;
      if (valid <= 0) continue;
      if (row_ok) {
        float* dst = p.resid + static_cast<size_t>(row) * p.ldo + col;
        if (valid >= 32) {
          float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4 r = d4[i];
            r.x += v[4 * i]; r.y += v[4 * i + 1]; r.z += v[4 * i + 2]; r.w += v[4 * i + 3];
            d4[i] = r;
            v[4 * i] = r.x; v[4 * i + 1] = r.y; v[4 * i + 2] = r.z; v[4 * i + 3] = r.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < valid) { v[i] += dst[i]; dst[i] = v[i]; } else { v[i] = 0.f; }
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) ss = fmaf(v[i], v[i], ss);
        store_row32_16<act_t>(p.xb + static_cast<size_t>(row) * p.ldo + col, v, valid);
      }
    }
    if (row_ok) p.ssq[static_cast<size_t>(row) * (2 * d.n_tiles) + 2 * n_tile + half] = ss;
  }
};

// Fused QKV projection epilogue: + bias, rotary embedding on Q and K heads (half-split rotation, per-token position
// looked up in a host-built cos/sin table), Q -> q_out[token], K/V -> kv cache rows kv_slot[token].
// Column layout of the fused weight: [ Q heads | K heads | V heads ], every head head_dim wide; kBN is a multiple of head_dim.
// The table is stored transposed, [head_dim/2][max_pos]: consecutive tokens (= consecutive TMEM lanes = the lanes of a
// warp) have consecutive positions, so one table load of a warp touches one or two 128 B lines instead of 32.
// Work split between the two epilogue warpgroups: a tile holds 4 (32-wide rotary chunk, head) items; each warpgroup takes
// the two items that share one rotary chunk (head_dim 128: chunk = half, both heads; head_dim 64: heads 2*half, 2*half+1).
template <int kHeadDim>
struct EpiQkvRope {
  static constexpr int kGroups = 2;
  struct Params {
    act_t* q_out;  // [M, n_q]
    act_t* k_out;  // [slots, n_kv]
    act_t* v_out;  // [slots, n_kv]
    const float* bias;     // [n_q + 2 n_kv]
    const int* pos;        // [M] rotary position of each token
    const int* kv_slot;    // [M] cache row of each token (nullptr: row index)
    const float* cos_tab;  // [kHeadDim/2][max_pos]
    const float* sin_tab;
    int n_q, n_kv, max_pos;
    const float* rstd;     // [M] 1/rms of the (un-normalised) A rows, nullptr = A is already normalised
  };
  __device__ static void run(const Params& p, const GemmDims& d, int row, int n0, int n_tile, uint32_t taddr, int half, uint8_t* wstage) {
    constexpr int kHalf = kHeadDim / 2;
    static_assert(kHeadDim == 64 || kHeadDim == 128, "head_dim must be 64 or 128");
    const bool row_ok = row < d.M;
    int pos = 0, slot = 0;
    float rs = 1.f;
    if (row_ok) {
      pos = __ldg(p.pos + row);
      slot = p.kv_slot ? __ldg(p.kv_slot + row) : row;
      if (p.rstd) rs = __ldg(p.rstd + row);
    }
    const int j0 = (kHeadDim == 128) ? half * 32 : 0;          // rotary chunk inside the half-dim
    const int head0 = (kHeadDim == 128) ? 0 : half * 2;        // first of the two heads this warpgroup handles
    float cs[32], sn[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      cs[i] = __ldg(p.cos_tab + static_cast<size_t>(j0 + i) * p.max_pos + pos);
      sn[i] = __ldg(p.sin_tab + static_cast<size_t>(j0 + i) * p.max_pos + pos);
    }
#pragma unroll 1
    for (int hh = 0; hh < 2; ++hh) {
      const int h = head0 + hh;
      const int c0 = n0 + h * kHeadDim;  // first fused column of this head
      if (c0 >= d.N) break;              // warp-uniform
      act_t* dst;
      bool rope;
      if (c0 < p.n_q) {
        dst = p.q_out + static_cast<size_t>(row) * p.n_q + c0;
        rope = true;
      } else if (c0 < p.n_q + p.n_kv) {
        dst = p.k_out + static_cast<size_t>(slot) * p.n_kv + (c0 - p.n_q);
        rope = true;
      } else {
        dst = p.v_out + static_cast<size_t>(slot) * p.n_kv + (c0 - p.n_q - p.n_kv);
        rope = false;
      }
      float lo[32], hi[32];
      tmem_ld32f(taddr + h * kHeadDim + j0, lo);
      tmem_ld32f(taddr + h * kHeadDim + kHalf + j0, hi);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float a = fmaf(lo[i], rs, __ldg(p.bias + c0 + j0 + i));
        const float b = fmaf(hi[i], rs, __ldg(p.bias + c0 + kHalf + j0 + i));
        // q*cos + rotate_half(q)*sin:  first half q1*cos - q2*sin, second half q2*cos + q1*sin
        lo[i] = rope ? a * cs[i] - b * sn[i] : a;
        hi[i] = rope ? b * cs[i] + a * sn[i] : b;
      }
      // (staging these rows through shared memory like the other epilogues, with vector bias loads, measured 2 % SLOWER on
      // the same box: per-row destinations need a shuffle per piece, and the rotary math, not the stores, bounds this epilogue)
      if (row_ok) {
        store_row32_16<act_t>(dst + j0, lo, 32);
        store_row32_16<act_t>(dst + kHalf + j0, hi, 32);
      }
    }
  }
};

// SwiGLU: the fused gate|up weight is interleaved in 128-row blocks, so tile n holds gate[128n..128n+127] in columns
// 0..127 and up[128n..128n+127] in columns 128..255.  act[row, 128 n + c] = silu(gate) * up.
struct EpiSwiglu {
  static constexpr int kGroups = 2;
  struct Params {
    act_t* act;          // [M, I]
    int ldo;             // = I
    const float* rstd;   // [M] 1/rms of the (un-normalised) A rows, nullptr = A is already normalised
  };
  __device__ static void run(const Params& p, const GemmDims& d, int row, int n0, int n_tile, uint32_t taddr, int half, uint8_t* wstage) {
    const bool row_ok = row < d.M;
    const float rs = (row_ok && p.rstd) ? __ldg(p.rstd + row) : 1.f;
#pragma unroll 1
    for (int c = half * 64; c < half * 64 + 64; c += 32) {
      float g[32], u[32];
      tmem_ld32f(taddr + c, g);
      tmem_ld32f(taddr + 128 + c, u);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float x = g[i] * rs;
#ifdef BLIM_EXACT_SILU
        g[i] = (x / (1.0f + __expf(-x))) * (u[i] * rs);
#else
        g[i] = x * fast_rcp(1.0f + fast_exp2(-1.4426950408889634f * x)) * (u[i] * rs);   // silu(x) * up: one MUFU.EX2 + one MUFU.RCP
#endif
      }
      store_tile32_16_staged<act_t>(p.act, p.ldo, row - static_cast<int>(threadIdx.x & 31), n_tile * 128 + c, g, d.M, wstage);
    }
  }
};

// Fused "logits never materialised" epilogue: per row, the running (max, sum-exp) over this tile's columns of
// scale*acc, plus the scaled logit of the row's target column if it falls into this tile.
struct EpiLse {
  static constexpr int kGroups = 2;
  struct Params {
    float2* partial;    // [M, 2 * n_tiles] (max, sumexp) per half tile
    float* tgt_logit;   // [M]
    const int* target;  // [M]
    float scale;
  };
  __device__ static void run(const Params& p, const GemmDims& d, int row, int n0, int n_tile, uint32_t taddr, int half, uint8_t* wstage) {
    const bool row_ok = row < d.M;
    const int tgt = row_ok ? __ldg(p.target + row) : -1;
    constexpr float kLog2e = 1.4426950408889634f;
    float m_run = -INFINITY, s_run = 0.f, t_val = 0.f;
    bool t_hit = false;
#pragma unroll 1
    for (int c = half * (kBN / 2); c < (half + 1) * (kBN / 2); c += 32) {
      float v[32];
      tmem_ld32f(taddr + c, v);
      const int col = n0 + c;
      const int valid = d.N - col;
      if (valid <= 0) continue;
      float cm = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        v[i] = (i < valid) ? v[i] * p.scale : -INFINITY;
        cm = fmaxf(cm, v[i]);
      }
      const int tc = tgt - col;
      if (tc >= 0 && tc < 32) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i == tc) t_val = v[i];
        t_hit = true;
      }
      const float m_new = fmaxf(m_run, cm);
      float s = 0.f;
      const float m_l2 = m_new * kLog2e;
#pragma unroll
#ifdef BLIM_EXACT_LSE
      for (int i = 0; i < 32; ++i) s += exp2f((v[i] - m_new) * kLog2e);
#else
      for (int i = 0; i < 32; ++i) s += fast_exp2(fmaf(v[i], kLog2e, -m_l2));   // one FFMA + one MUFU.EX2 per logit
#endif
      s_run = s_run * exp2f((m_run - m_new) * kLog2e) + s;
      m_run = m_new;
    }
    if (row_ok) {
      p.partial[static_cast<size_t>(row) * (2 * d.n_tiles) + 2 * n_tile + half] = make_float2(m_run, s_run);
      if (t_hit) p.tgt_logit[row] = t_val;
    }
  }
};

// ------------------------------------------------------------------------------------------------ the kernel
template <class Epi, int kCtaGroup>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const GemmDims dims,
                    const __grid_constant__ typename Epi::Params ep) {   // grid constant: an epilogue may hold a tensor map (EpiResidTma)
  using Cfg = GemmCfg<kCtaGroup>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* s_a = smem;
  uint8_t* s_b = smem + Cfg::kStages * Cfg::kABytes;
  uint8_t* s_out = smem + Cfg::kStages * Cfg::kStageBytes;   // epilogue staging slabs (1024-byte aligned)
  uint64_t* bar_full = reinterpret_cast<uint64_t*>(s_out + Cfg::kOutStageBytes);
  uint64_t* bar_empty = bar_full + Cfg::kStages;
  uint64_t* bar_tfull = bar_empty + Cfg::kStages;
  uint64_t* bar_tempty = bar_tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (kCtaGroup == 2) ? cluster_ctarank() : 0u;
  const bool is_leader = cta_rank == 0;
  const int worker = blockIdx.x / kCtaGroup;       // persistent worker (CTA or CTA pair) index
  const int n_workers = gridDim.x / kCtaGroup;
  const int total_tiles = dims.m_tiles * dims.n_tiles;
  const int num_kb = dims.K / kBK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < Cfg::kStages; ++i) {
      mbar_init(&bar_full[i], 1);
      mbar_init(&bar_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bar_tfull[i], 1);
      mbar_init(&bar_tempty[i], 128 * Epi::kGroups * kCtaGroup);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<kCtaGroup>(tmem_slot, kTmemCols);
  tc_fence_before();
  if constexpr (kCtaGroup == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ===================================================== TMA producer
    int stage = 0;
    uint32_t phase = 0;
    const bool hints = dims.l2_hints != 0;
    const uint64_t pol_a = hints ? l2_policy_evict_last() : 0ull, pol_b = hints ? l2_policy_evict_first() : 0ull;
    for (int t = worker; t < total_tiles; t += n_workers) {
      int m_tile, n_tile;
      tile_coords(dims, t, m_tile, n_tile);
      const int m0 = m_tile * (kBM * kCtaGroup) + static_cast<int>(cta_rank) * kBM;
      const int n0 = n_tile * kBN + static_cast<int>(cta_rank) * Cfg::kBRows;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&bar_empty[stage], phase ^ 1);
        if (is_leader) mbar_arrive_expect_tx(&bar_full[stage], Cfg::kStageBytes * kCtaGroup);
        if (hints) {
          if constexpr (kCtaGroup == 1) {
            tma_load_2d_hint(s_a + stage * Cfg::kABytes, &tm_a, &bar_full[stage], kb * kBK, m0, pol_a);
            tma_load_2d_hint(s_b + stage * Cfg::kBBytes, &tm_b, &bar_full[stage], kb * kBK, n0, pol_b);
          } else {
            tma_load_2d_pair_hint(s_a + stage * Cfg::kABytes, &tm_a, &bar_full[stage], kb * kBK, m0, pol_a);
            tma_load_2d_pair_hint(s_b + stage * Cfg::kBBytes, &tm_b, &bar_full[stage], kb * kBK, n0, pol_b);
          }
        } else if constexpr (kCtaGroup == 1) {
          tma_load_2d(s_a + stage * Cfg::kABytes, &tm_a, &bar_full[stage], kb * kBK, m0);
          tma_load_2d(s_b + stage * Cfg::kBBytes, &tm_b, &bar_full[stage], kb * kBK, n0);
        } else {
          tma_load_2d_pair(s_a + stage * Cfg::kABytes, &tm_a, &bar_full[stage], kb * kBK, m0);
          tma_load_2d_pair(s_b + stage * Cfg::kBBytes, &tm_b, &bar_full[stage], kb * kBK, n0);
        }
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0 && is_leader) {
    // ===================================================== MMA issuer (single thread)
    const uint32_t idesc = make_idesc_f16kind(kBM * kCtaGroup, kBN, dims.a_fmt, dims.b_fmt);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = worker; t < total_tiles; t += n_workers) {
      mbar_wait(&bar_tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * kBN);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&bar_full[stage], phase);
        tc_fence_after();
        const uint64_t da = make_smem_desc_sw128(smem_u32(s_a + stage * Cfg::kABytes));
        const uint64_t db = make_smem_desc_sw128(smem_u32(s_b + stage * Cfg::kBBytes));
#pragma unroll
        for (int k = 0; k < kBK / kUmmaK; ++k) {
          // advance 16 elements = 32 B along K inside the 128 B swizzle row: +2 in the (addr >> 4) field
          umma_bf16<kCtaGroup>(d_tmem, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc,
                               (kb | k) != 0 ? 1u : 0u);
        }
        if constexpr (kCtaGroup == 1) umma_commit(&bar_empty[stage]); else umma_commit_pair(&bar_empty[stage]);
        if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
      }
      if constexpr (kCtaGroup == 1) umma_commit(&bar_tfull[acc]); else umma_commit_pair(&bar_tfull[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 4 && (warp - 4) < 4 * Epi::kGroups) {
    // ===================================================== epilogue (kGroups warpgroups x 4 warps; 4 warps = 128 TMEM lanes = 128 rows)
    const int ew = (warp - 4) & 3;       // TMEM lane quadrant (= warp % 4)
    const int half = (warp - 4) >> 2;    // which half of the tile's columns / work items
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = worker; t < total_tiles; t += n_workers) {
      int m_tile, n_tile;
      tile_coords(dims, t, m_tile, n_tile);
      const int row = m_tile * (kBM * kCtaGroup) + static_cast<int>(cta_rank) * kBM + ew * 32 + lane;
      mbar_wait(&bar_tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + static_cast<uint32_t>(acc * kBN);
      Epi::run(ep, dims, row, n_tile * kBN, n_tile, taddr, half, s_out + (warp - 4) * Cfg::kOutWarpBytes);
      tc_fence_before();
      if (is_leader) mbar_arrive(&bar_tempty[acc]); else mbar_arrive_cluster(&bar_tempty[acc], 0);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  // ===================================================== teardown
  tc_fence_before();
  if constexpr (kCtaGroup == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  if (warp == 2) tmem_dealloc<kCtaGroup>(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// bf16 row-major [rows, cols] matrix, row pitch ld elements; box = 64 columns x box_rows rows, 128 B swizzle.
// (16-bit elements: the map only moves bytes, so bf16 and fp16 tensors use the same encoding)
inline bool make_tmap_bf16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kBK), box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// fp32 row-major [rows, cols] matrix; box = 32 columns x 32 rows, 128 B swizzle (EpiResidTma)
inline bool make_tmap_f32_32x32(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) return false;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 4};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

struct GemmLaunchCtx {
  int num_sms = 148;
  int l2_hints = 0;   // BLIM_GEMM_HINTS=1: eviction-priority hints on the operand loads (see GemmDims::l2_hints)
  int sb_mb = 40;     // BLIM_GEMM_SB_MB: target size of the L2-resident A slab (MB)
  int sb_min = 2;     // BLIM_GEMM_SB_MIN: fewest m-tiles per super-block
  bool sb_auto = true;  // per-shape choice between the 40 MB slab and one m-tile at a time; BLIM_GEMM_SB_MB / _MIN switch it off
  int cta_group = 1;  // 1: one CTA per tile, 2: CTA pairs (cta_group::2, 256-row tiles)
  long long launches = 0;
  int device = 0;     // CUDA device of the owning engine
};

template <class Epi, int kCtaGroup>
inline cudaError_t launch_gemm_impl(GemmLaunchCtx& ctx, const void* A, int lda, const void* W, int ldw, int M, int N,
                                    int K, const typename Epi::Params& ep, cudaStream_t stream, int a_fmt, int b_fmt) {
  using Cfg = GemmCfg<kCtaGroup>;
  if (M <= 0 || N <= 0) return cudaSuccess;
  if (K <= 0 || (K % kBK) != 0) return cudaErrorInvalidValue;
  CUtensorMap ta, tb;
  if (!make_tmap_bf16(&ta, A, static_cast<uint64_t>(M), static_cast<uint64_t>(K), static_cast<uint64_t>(lda), kBM)) return cudaErrorInvalidValue;
  if (!make_tmap_bf16(&tb, W, static_cast<uint64_t>(N), static_cast<uint64_t>(K), static_cast<uint64_t>(ldw), Cfg::kBRows)) return cudaErrorInvalidValue;
  GemmDims d;
  d.M = M; d.N = N; d.K = K;
  const int tile_m = kBM * kCtaGroup;
  d.m_tiles = (M + tile_m - 1) / tile_m;
  d.n_tiles = (N + kBN - 1) / kBN;
  // L2 blocking (ncu DRAM bytes per launch, profiles/ + DESIGN.md 6):
  //  * W larger than L2 and many n-tiles (gate|up, LM head): keep an A super-block of ~40 MB L2-resident and stream W once
  //    per super-block;
  //  * W that fits L2 by itself (QKV, o_proj, the ViT's GEMMs), or a super-block shorter than two waves of CTAs
  //    (down_proj: 4 m-tiles x 14 n-tiles against 74 concurrent CTA pairs, so every A panel was touched by two waves):
  //    one m-tile at a time with n fastest -- the CTAs that share an A panel run in the same wave, A is read once.
  long long sb = (static_cast<long long>(ctx.sb_mb) << 20) / (static_cast<long long>(tile_m) * K * 2);
  if (sb < ctx.sb_min) sb = ctx.sb_min;
  if (sb > 64) sb = 64;
  const int total = d.m_tiles * d.n_tiles;
  int workers = ctx.num_sms / kCtaGroup;
  if (workers > total) workers = total;
  if (ctx.sb_auto) {
    const long long w_bytes = static_cast<long long>(N) * K * 2;
    if (w_bytes <= (48ll << 20) || sb * d.n_tiles < 2ll * workers) sb = 1;
  }
  d.sb_tiles = static_cast<int>(sb);
  d.l2_hints = ctx.l2_hints;
  d.a_fmt = a_fmt; d.b_fmt = b_fmt;
  auto kern = gemm_tcgen05_kernel<Epi, kCtaGroup>;
  static bool attr_set[64] = {};  // per instantiation AND per device (function attributes are per device)
  const int dev_slot = ctx.device & 63;
  if (!attr_set[dev_slot]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return e;
    attr_set[dev_slot] = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(workers * kCtaGroup));
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCtaGroup;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  ctx.launches++;
  return cudaLaunchKernelEx(&cfg, kern, ta, tb, d, ep);
}

// A [M, K] and W [N, K] are 16-bit K-major operands in the formats a_fmt / b_fmt (kFmtF16 / kFmtBF16).
template <class Epi>
inline cudaError_t launch_gemm(GemmLaunchCtx& ctx, const void* A, int lda, const void* W, int ldw, int M, int N, int K,
                               const typename Epi::Params& ep, cudaStream_t stream, int a_fmt = kFmtBF16, int b_fmt = kFmtBF16) {
  if (ctx.cta_group == 2) return launch_gemm_impl<Epi, 2>(ctx, A, lda, W, ldw, M, N, K, ep, stream, a_fmt, b_fmt);
  return launch_gemm_impl<Epi, 1>(ctx, A, lda, W, ldw, M, N, K, ep, stream, a_fmt, b_fmt);
}

}  // namespace blim
