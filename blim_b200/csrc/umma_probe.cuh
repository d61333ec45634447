// Single-CTA tcgen05 probe used by tests/test_umma_probe_gpu.py: C[128, N] = A[128, K] · B, with the operands copied to
// shared memory by ordinary threads (manual 128-byte swizzle, no TMA) -- the staging scheme of the tcgen05 attention
// kernel.  B can be given K-major ([N][K], like a weight / the K matrix of attention) or MN-major ([K][N], like the V
// matrix of attention, row = key, contiguous head_dim).  The MN-major shared-memory descriptor fields (leading / stride
// byte offsets, per-MMA K advance) are runtime arguments so that their semantics are pinned by a test rather than
// assumed.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "act_type.cuh"
#include "ptx_sm100.cuh"

namespace blim {

// Byte offset of element (row, col) of a [rows x 64] bf16 tile stored as 8-row x 128 B swizzle atoms (SWIZZLE_128B):
// the 16-byte chunk index is XORed with (row % 8).
__device__ __forceinline__ uint32_t sw128_offset(int row, int col) {
  const int chunk = col >> 3;
  return static_cast<uint32_t>((row >> 3) * 1024 + (row & 7) * 128 + ((chunk ^ (row & 7)) << 4) + (col & 7) * 2);
}

__host__ __device__ inline uint64_t make_smem_desc_raw(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}


// A: [128][K] bf16 row-major (K = 64 or 128).  B: b_mn_major ? [K][N] : [N][K] row-major, N = 64 or 128.  C: [128][N] fp32.
__global__ void __launch_bounds__(128, 1) umma_probe_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B,
                                                           float* __restrict__ C, int K, int N, int b_mn_major, uint32_t lbo, uint32_t sbo,
                                                           uint32_t kstep_bytes, int a_fmt, int b_fmt) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* s_a = smem;                 // K/64 sub-tiles of [128 x 64]
  uint8_t* s_b = smem + 2 * 16384;     // K-major: K/64 sub-tiles of [N x 64]; MN-major: N/64 sub-tiles of [K x 64]
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  // stage A
  for (int i = tid; i < 128 * (K / 8); i += 128) {
    const int r = i / (K / 8), c = (i % (K / 8)) * 8;
    const uint4 v = *reinterpret_cast<const uint4*>(A + static_cast<size_t>(r) * K + c);
    *reinterpret_cast<uint4*>(s_a + (c >> 6) * 16384 + sw128_offset(r, c & 63)) = v;
  }
  if (!b_mn_major) {
    for (int i = tid; i < N * (K / 8); i += 128) {
      const int r = i / (K / 8), c = (i % (K / 8)) * 8;
      const uint4 v = *reinterpret_cast<const uint4*>(B + static_cast<size_t>(r) * K + c);
      *reinterpret_cast<uint4*>(s_b + (c >> 6) * (N * 128) + sw128_offset(r, c & 63)) = v;
    }
  } else {
    // rows = K index (keys), 64-column sub-tiles along N
    for (int i = tid; i < K * (N / 8); i += 128) {
      const int r = i / (N / 8), c = (i % (N / 8)) * 8;
      const uint4 v = *reinterpret_cast<const uint4*>(B + static_cast<size_t>(r) * N + c);
      *reinterpret_cast<uint4*>(s_b + (c >> 6) * (K * 128) + sw128_offset(r, c & 63)) = v;
    }
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<1>(&tmem_slot, 128);
  fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor-core (async) proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_f16kind(128, N, a_fmt, b_fmt, b_mn_major);
    for (int k = 0; k < K / 16; ++k) {
      const uint64_t da = make_smem_desc_sw128(smem_u32(s_a) + (k >> 2) * 16384 + (k & 3) * 32);
      uint64_t db;
      if (!b_mn_major) db = make_smem_desc_sw128(smem_u32(s_b) + (k >> 2) * (N * 128) + (k & 3) * 32);
      else db = make_smem_desc_raw(smem_u32(s_b) + k * kstep_bytes, lbo, sbo);
      umma_bf16<1>(tmem, da, db, idesc, k != 0 ? 1u : 0u);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c = 0; c < N; c += 32) {
    uint32_t r[32];
    tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c, r);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) C[static_cast<size_t>(tid) * N + c + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) tmem_dealloc<1>(tmem, 128);
}

// b_mn_major: bit 0 = B is MN-major; bit 1 = A holds fp16 (else bf16); bit 2 = B holds fp16 -- the two operand formats of
// kind::f16 are independent, which the scoring path relies on (fp16 activations x bf16 weights).
inline cudaError_t launch_umma_probe(const void* A, const void* B, float* C, int K, int N, int b_mn_major, uint32_t lbo, uint32_t sbo,
                                     uint32_t kstep_bytes, cudaStream_t st) {
  const int a_fmt = (b_mn_major & 2) ? kFmtF16 : kFmtBF16, b_fmt = (b_mn_major & 4) ? kFmtF16 : kFmtBF16;
  b_mn_major &= 1;
  if ((K != 64 && K != 128) || (N != 64 && N != 128)) return cudaErrorInvalidValue;
  const int smem = 1024 + 4 * 16384;
  cudaError_t e = cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  umma_probe_kernel<<<1, 128, smem, st>>>(reinterpret_cast<const __nv_bfloat16*>(A), reinterpret_cast<const __nv_bfloat16*>(B), C, K, N,
                                          b_mn_major, lbo, sbo, kstep_bytes, a_fmt, b_fmt);
  return cudaGetLastError();
}

}  // namespace blim
