// Memory-bound kernels of the BLiM scoring path: sequence assembly, RMSNorm, reductions of the fused LM-head / TVG
// epilogues, weight repacking, score scatter, and the CPN + ensemble + rerank stage.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "act_type.cuh"

namespace blim {

// ------------------------------------------------------------------------------------------------ sequence assembly
// Replaces embed_tokens + the torch.cat splice of prepare_inputs_labels_for_multimodal (reference:
// modeling_videochat_flash.py:395-433) without padding: token t takes row tok_src[t] of the embedding table when
// tok_src[t] >= 0, else row (-1 - tok_src[t]) of the projected visual rows.  Output: fp32 residual stream.
// The embedding table is a weight (bf16); the visual rows are activations (act_t).
__global__ void assemble_tokens_kernel(float* __restrict__ x, const __nv_bfloat16* __restrict__ embed,
                                       const act_t* __restrict__ visual, const int* __restrict__ tok_src, int T, int H) {
  const int t = blockIdx.x;
  if (t >= T) return;
  const int src = tok_src[t];
  const bool is_tok = src >= 0;
  const uint16_t* row = is_tok ? reinterpret_cast<const uint16_t*>(embed) + static_cast<size_t>(src) * H
                               : reinterpret_cast<const uint16_t*>(visual) + static_cast<size_t>(-1 - src) * H;
  float* dst = x + static_cast<size_t>(t) * H;
  for (int c = threadIdx.x * 8; c < H; c += blockDim.x * 8) {
    const uint4 u = *reinterpret_cast<const uint4*>(row + c);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    float f[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 v = is_tok ? Fmt16<__nv_bfloat16>::unpack2(w[i]) : Fmt16<act_t>::unpack2(w[i]);
      f[2 * i] = v.x; f[2 * i + 1] = v.y;
    }
    *reinterpret_cast<float4*>(dst + c) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(dst + c + 4) = make_float4(f[4], f[5], f[6], f[7]);
  }
}

// bf16 [R, H] -> fp32 rows (compat forward: caller-provided inputs_embeds)
__global__ void bf16_rows_to_f32_kernel(float* __restrict__ x, const __nv_bfloat16* __restrict__ in, size_t n) {
  size_t i = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 2;
  if (i + 1 < n) {
    const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(in + i));
    x[i] = f.x; x[i + 1] = f.y;
  } else if (i < n) {
    x[i] = __bfloat162float(in[i]);
  }
}

// ------------------------------------------------------------------------------------------------ RMSNorm
// Reference: Qwen2RMSNorm.forward (modeling_qwen2_flash.py:93-98): fp32 variance, x * rsqrt(var + eps) cast to the model
// dtype, then multiplied by the (model dtype) weight.  Here the product is formed in fp32 and rounded ONCE to the
// activation format (the reference's intermediate cast only adds a rounding).  Row gather: row r reads
// x0[idx[r]] when idx[r] >= 0 else x1[-1 - idx[r]] (idx == nullptr -> identity on x0).
__global__ void __launch_bounds__(256) rmsnorm_kernel(act_t* __restrict__ out, const float* __restrict__ x0, const float* __restrict__ x1,
                                                      const int* __restrict__ idx, const float* __restrict__ weight, int R, int H, float eps) {
  const int r = blockIdx.x;
  if (r >= R) return;
  const float* src;
  if (idx) {
    const int i = idx[r];
    src = i >= 0 ? x0 + static_cast<size_t>(i) * H : x1 + static_cast<size_t>(-1 - i) * H;
  } else {
    src = x0 + static_cast<size_t>(r) * H;
  }
  // The row stays in registers between the two passes (up to 4 float4 per thread = H <= 4096 with 256 threads; all four
  // loads are issued before the first use, so a block has its whole 14 KB row in flight at once); wider rows re-read it.
  constexpr int kVec = 4;
  const int n_vec = H >> 2;
  const bool in_regs = n_vec <= kVec * 256;
  float4 v[kVec];
  float ss = 0.f;
  if (in_regs) {
#pragma unroll
    for (int k = 0; k < kVec; ++k) {
      const int c = threadIdx.x + k * 256;
      v[k] = c < n_vec ? __ldcs(reinterpret_cast<const float4*>(src) + c) : make_float4(0.f, 0.f, 0.f, 0.f);   // read once: streaming
    }
#pragma unroll
    for (int k = 0; k < kVec; ++k) ss += v[k].x * v[k].x + v[k].y * v[k].y + v[k].z * v[k].z + v[k].w * v[k].w;
  } else {
    for (int c = threadIdx.x; c < n_vec; c += 256) {
      const float4 t = reinterpret_cast<const float4*>(src)[c];
      ss += t.x * t.x + t.y * t.y + t.z * t.z + t.w * t.w;
    }
  }
  __shared__ float red[8];
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) tot += red[k];      // same order in every thread: one barrier, deterministic
  const float rstd = rsqrtf(tot / static_cast<float>(H) + eps);
  act_t* dst = out + static_cast<size_t>(r) * H;
  if (in_regs) {
#pragma unroll
    for (int k = 0; k < kVec; ++k) {
      const int c = threadIdx.x + k * 256;
      if (c < n_vec) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(weight) + c);
        uint2 u;
        u.x = Fmt16<act_t>::pack2((v[k].x * rstd) * w.x, (v[k].y * rstd) * w.y);
        u.y = Fmt16<act_t>::pack2((v[k].z * rstd) * w.z, (v[k].w * rstd) * w.w);
        reinterpret_cast<uint2*>(dst)[c] = u;
      }
    }
  } else {
    for (int c = threadIdx.x; c < n_vec; c += 256) {
      const float4 t = reinterpret_cast<const float4*>(src)[c];
      const float4 w = __ldg(reinterpret_cast<const float4*>(weight) + c);
      uint2 u;
      u.x = Fmt16<act_t>::pack2((t.x * rstd) * w.x, (t.y * rstd) * w.y);
      u.y = Fmt16<act_t>::pack2((t.z * rstd) * w.z, (t.w * rstd) * w.w);
      reinterpret_cast<uint2*>(dst)[c] = u;
    }
  }
}

// Fused-RMSNorm support (see EpiResidNorm in gemm_sm100.cuh).
// rstd[r] = rsqrt(sum(parts[r, :]) / H + eps): fixed-order reduction of the per-half-tile sums of squares.
__global__ void rstd_rows_kernel(float* __restrict__ rstd, const float* __restrict__ parts, int R, int n_parts, int H, float eps) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  float s = 0.f;
  for (int i = 0; i < n_parts; ++i) s += parts[static_cast<size_t>(r) * n_parts + i];
  rstd[r] = rsqrtf(s / static_cast<float>(H) + eps);
}
// Entry of a decoder run (and the pruned last layer): xb = act_t(x), rstd = 1/rms(x) for R fp32 rows.
__global__ void rowprep_kernel(act_t* __restrict__ xb, float* __restrict__ rstd, const float* __restrict__ x, int R, int H, float eps) {
  const int r = blockIdx.x;
  if (r >= R) return;
  const float* src = x + static_cast<size_t>(r) * H;
  float ss = 0.f;
  for (int c = threadIdx.x * 4; c < H; c += blockDim.x * 4) {
    const float4 v = *reinterpret_cast<const float4*>(src + c);
    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    uint2 u;
    u.x = Fmt16<act_t>::pack2(v.x, v.y);
    u.y = Fmt16<act_t>::pack2(v.z, v.w);
    *reinterpret_cast<uint2*>(xb + static_cast<size_t>(r) * H + c) = u;
  }
  __shared__ float red[32];
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) rstd[r] = rsqrtf(v / static_cast<float>(H) + eps);
  }
}
// W[n, k] *= g[k] in place (operand-format weight, fp32 norm weight): folds an RMSNorm weight into the GEMM that consumes its output.
__global__ void fold_norm_weight_kernel(act_t* __restrict__ w, const float* __restrict__ g, size_t rows, int cols) {
  const size_t n = rows * cols;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
    w[i] = Fmt16<act_t>::from_float(Fmt16<act_t>::to_float(w[i]) * g[i % cols]);
}

// copy selected fp32 rows: dst[r] = src[idx[r]]   (saves the last-prefix-token state of every prefix sequence)
__global__ void gather_rows_f32_kernel(float* __restrict__ dst, const float* __restrict__ src, const int* __restrict__ idx, int R, int H) {
  const int r = blockIdx.x;
  if (r >= R) return;
  const float* s = src + static_cast<size_t>(idx[r]) * H;
  float* d = dst + static_cast<size_t>(r) * H;
  for (int c = threadIdx.x * 4; c < H; c += blockDim.x * 4) *reinterpret_cast<float4*>(d + c) = *reinterpret_cast<const float4*>(s + c);
}

// dst[r] = src[idx[r]] for rows of `w` 16-bit elements of any format (w % 8 == 0)
__global__ void gather_rows_16_kernel(void* __restrict__ dst, const void* __restrict__ src, const int* __restrict__ idx, int R, int w) {
  const int r = blockIdx.x;
  if (r >= R) return;
  const uint16_t* s = reinterpret_cast<const uint16_t*>(src) + static_cast<size_t>(idx[r]) * w;
  uint16_t* d = reinterpret_cast<uint16_t*>(dst) + static_cast<size_t>(r) * w;
  for (int c = threadIdx.x * 8; c < w; c += blockDim.x * 8) *reinterpret_cast<uint4*>(d + c) = *reinterpret_cast<const uint4*>(s + c);
}

// TVG visual rows: mean over the `group` tokens of a clip (reference: frame_feature.mean(1), modeling_videochat_flash.py:243),
// fp32 accumulate, result in the activation format.  in [R*group, H] -> out [R, H]
__global__ void mean_rows_kernel(act_t* __restrict__ out, const act_t* __restrict__ in, int R, int group, int H) {
  const int r = blockIdx.x;
  if (r >= R) return;
  for (int c = threadIdx.x * 2; c < H; c += blockDim.x * 2) {
    float a = 0.f, b = 0.f;
    for (int g = 0; g < group; ++g) {
      const float2 f = Fmt16<act_t>::unpack2(*reinterpret_cast<const uint32_t*>(in + (static_cast<size_t>(r) * group + g) * H + c));
      a += f.x; b += f.y;
    }
    *reinterpret_cast<uint32_t*>(out + static_cast<size_t>(r) * H + c) = Fmt16<act_t>::pack2(a / group, b / group);
  }
}

// video_vocab built on the device from the stored features (reference: load_video_feature(vid).mean(1) per video,
// dataloader/base_dataset.py:33-37): vocab[c][label[v]][:] = mean over the `group` tokens of clip c of video v.
// grid = (n_videos * n_clips); feats [n_videos * n_clips * group, mm]; vocab [n_clips][n_vocab][mm], both in the activation format.
__global__ void vocab_from_feats_kernel(act_t* __restrict__ vocab, const act_t* __restrict__ feats,
                                        const int* __restrict__ labels, int n_clips, int group, int mm, int n_vocab) {
  const int v = blockIdx.x / n_clips, c = blockIdx.x % n_clips;
  const act_t* src = feats + static_cast<size_t>(blockIdx.x) * group * mm;
  act_t* dst = vocab + (static_cast<size_t>(c) * n_vocab + labels[v]) * mm;
  for (int d = threadIdx.x * 2; d < mm; d += blockDim.x * 2) {
    float a = 0.f, b = 0.f;
    for (int g = 0; g < group; ++g) {
      const float2 f = Fmt16<act_t>::unpack2(*reinterpret_cast<const uint32_t*>(src + static_cast<size_t>(g) * mm + d));
      a += f.x; b += f.y;
    }
    *reinterpret_cast<uint32_t*>(dst + d) = Fmt16<act_t>::pack2(a / group, b / group);
  }
}

// ------------------------------------------------------------------------------------------------ log-softmax reductions
// Merge the per-half-tile (max, sumexp) partials of EpiLse: logp[r] = tgt_logit[r] - (m + log(sum)).
__global__ void lse_finalize_kernel(float* __restrict__ logp, const float2* __restrict__ partial, const float* __restrict__ tgt_logit,
                                    int R, int n_tiles) {
  const int warps_per_block = blockDim.x >> 5;
  const int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  if (r >= R) return;
  const int lane = threadIdx.x & 31;
  float m = -INFINITY, s = 0.f;
  for (int i = lane; i < n_tiles; i += 32) {
    const float2 p = partial[static_cast<size_t>(r) * n_tiles + i];
    const float mn = fmaxf(m, p.x);
    if (mn == -INFINITY) continue;
    s = s * __expf(m - mn) + p.y * __expf(p.x - mn);
    m = mn;
  }
  for (int o = 16; o > 0; o >>= 1) {
    const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
    const float mn = fmaxf(m, m2);
    if (mn != -INFINITY) {
      s = s * __expf(m - mn) + s2 * __expf(m2 - mn);
      m = mn;
    }
  }
  if (lane == 0) logp[r] = tgt_logit[r] - (m + logf(s));
}

// VTG score of pair p = sum(logp rows) / count(non-zero NLL rows)   (reference: VTGCriterion, retrieval_utils.py:31-33)
__global__ void vtg_seq_mean_kernel(float* __restrict__ scores, const float* __restrict__ logp, const int* __restrict__ row_off, int P) {
  const int warps_per_block = blockDim.x >> 5;
  const int p = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  if (p >= P) return;
  const int lane = threadIdx.x & 31;
  float s = 0.f;
  int cnt = 0;
  for (int r = row_off[p] + lane; r < row_off[p + 1]; r += 32) {
    const float v = logp[r];
    s += v;
    cnt += (v != 0.f);
  }
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if (lane == 0) scores[p] = s / static_cast<float>(cnt);
}

// TVG score of pair p = mean over clips of logp[c * P + p]   (reference: TVGCriterion, retrieval_utils.py:41-43)
__global__ void tvg_clip_mean_kernel(float* __restrict__ scores, const float* __restrict__ logp, int P, int n_clips) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  float s = 0.f;
  for (int c = 0; c < n_clips; ++c) s += logp[static_cast<size_t>(c) * P + p];
  scores[p] = s / static_cast<float>(n_clips);
}

// out[i] = src[map[i]]  (expand unique-key results to every requested pair)
__global__ void expand_scores_kernel(float* __restrict__ out, const float* __restrict__ src, const int* __restrict__ map, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = src[map[i]];
}

// ------------------------------------------------------------------------------------------------ weight repacking
// dst (bf16 / fp32) <- src (fp32 = 0, bf16 = 1, fp16 = 2).  Row r of the source goes to destination row
//   interleave == 0 : dst_row0 + r
//   interleave  > 0 : (r / interleave) * 2 * interleave + half * interleave + r % interleave     (gate|up 128-row blocks)
__device__ __forceinline__ float load_as_f32(const void* src, int dtype, size_t i) {
  if (dtype == 0) return reinterpret_cast<const float*>(src)[i];
  if (dtype == 1) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(src)[i]);
  return __half2float(reinterpret_cast<const __half*>(src)[i]);
}
template <typename T>   // T = act_t (GEMM operands: weights, features) or __nv_bfloat16 (embedding table, the extractor)
__global__ void repack_rows_16_kernel(T* __restrict__ dst, const void* __restrict__ src, int dtype, int rows, int cols,
                                      int dst_row0, int interleave, int half) {
  const size_t n = static_cast<size_t>(rows) * cols;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / cols), c = static_cast<int>(i % cols);
    const int dr = interleave > 0 ? (r / interleave) * 2 * interleave + half * interleave + r % interleave : dst_row0 + r;
    dst[static_cast<size_t>(dr) * cols + c] = Fmt16<T>::from_float(load_as_f32(src, dtype, i));
  }
}
// fp32 destination.  round_bf16 != 0 rounds through bf16 first (parameters that live in the model dtype in the reference).
__global__ void repack_f32_kernel(float* __restrict__ dst, const void* __restrict__ src, int dtype, size_t n, int round_bf16) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    float v = load_as_f32(src, dtype, i);
    if (round_bf16) v = __bfloat162float(__float2bfloat16(v));
    dst[i] = v;
  }
}
// video_vocab [N_v, n_clips, MM] -> [n_clips, N_v, MM] in the activation format (one K-major B operand per clip)
__global__ void repack_vocab_kernel(act_t* __restrict__ dst, const void* __restrict__ src, int dtype, int n_v, int n_clips, int mm) {
  const size_t n = static_cast<size_t>(n_v) * n_clips * mm;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int d = static_cast<int>(i % mm);
    const int c = static_cast<int>((i / mm) % n_clips);
    const int u = static_cast<int>(i / (static_cast<size_t>(mm) * n_clips));
    dst[(static_cast<size_t>(c) * n_v + u) * mm + d] = Fmt16<act_t>::from_float(load_as_f32(src, dtype, i));
  }
}

// dst[c][r] = src[r][c]   (rotary tables: [positions, head_dim/2] -> [head_dim/2, positions])
__global__ void transpose_f32_kernel(float* __restrict__ dst, const float* __restrict__ src, int rows, int cols) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const int r = i / cols, c = i % cols;
  dst[static_cast<size_t>(c) * rows + r] = src[i];
}

// ------------------------------------------------------------------------------------------------ dense score matrices
__global__ void fill_f32_kernel(float* __restrict__ dst, float v, size_t n) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) dst[i] = v;
}
// S[row[i], col[i]] = val[i]   (reference: retrieval_utils.py:110,152)
__global__ void scatter_scores_kernel(float* __restrict__ dense, int n_cols, const int* __restrict__ row, const int* __restrict__ col,
                                      const float* __restrict__ val, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dense[static_cast<size_t>(row[i]) * n_cols + col[i]] = val[i];
}

// Replicate the K/V rows [0, n_rows) of every layer of the prefix cache to rows [dst[u], dst[u] + n_rows) for every unit u:
// the shared "root" of the prefixes (chat-template header) is prefilled once and copied in front of each unit's own rows.
// grid = (n_units, n_layers, 2 [K, V]); w = row width in elements (w % 8 == 0).
__global__ void replicate_root_rows_kernel(act_t* __restrict__ kp, act_t* __restrict__ vp, const int* __restrict__ dst,
                                           int n_rows, int w, size_t layer_stride) {
  act_t* base = (blockIdx.z == 0 ? kp : vp) + static_cast<size_t>(blockIdx.y) * layer_stride;
  const act_t* src = base;
  act_t* out = base + static_cast<size_t>(dst[blockIdx.x]) * w;
  const int n = n_rows * (w / 8);
  for (int i = threadIdx.x; i < n; i += blockDim.x) reinterpret_cast<uint4*>(out)[i] = reinterpret_cast<const uint4*>(src)[i];
}

// ------------------------------------------------------------------------------------------------ stage-1 candidates
// Per-row top-k of the InternVideo2 similarity rows (reference: sims.topk(k), retrieval_utils.py:52,117): one warp per row,
// k rounds of a warp arg-max over the row staged in shared memory (ties: lowest column first), results in descending order.
__global__ void topk_rows_kernel(const float* __restrict__ mat, int n_rows, int n_cols, int k, int* __restrict__ idx_out,
                                 float* __restrict__ val_out) {
  extern __shared__ float smem_tk[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warps = blockDim.x >> 5;
  const int row = blockIdx.x * warps + warp;
  float* vals = smem_tk + static_cast<size_t>(warp) * n_cols;
  unsigned* taken = reinterpret_cast<unsigned*>(smem_tk + static_cast<size_t>(warps) * n_cols) + static_cast<size_t>(warp) * ((n_cols + 31) / 32);
  if (row >= n_rows) return;
  for (int i = lane; i < n_cols; i += 32) vals[i] = mat[static_cast<size_t>(row) * n_cols + i];
  for (int i = lane; i < (n_cols + 31) / 32; i += 32) taken[i] = 0u;
  __syncwarp();
  for (int r = 0; r < k; ++r) {
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = lane; i < n_cols; i += 32) {
      if (taken[i >> 5] & (1u << (i & 31))) continue;
      const float v = vals[i];
      if (bi == 0x7fffffff || v > best) { best = v; bi = i; }   // strided scan visits indices in increasing order
    }
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (oi != 0x7fffffff && (bi == 0x7fffffff || ov > best || (ov == best && oi < bi))) { best = ov; bi = oi; }
    }
    if (lane == 0) {
      idx_out[static_cast<size_t>(row) * k + r] = bi;
      val_out[static_cast<size_t>(row) * k + r] = best;
      taken[bi >> 5] |= (1u << (bi & 31));
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------ CPN + ensemble + rerank
// Reference: val_one_epoch (training_utils.py:154-165) + get_recall (training_utils.py:173-221).
// numpy semantics reproduced operation by operation: Python-float coefficients meet float32 arrays as float32 scalars,
// every multiply / add / subtract is a separately rounded fp32 operation (no FMA contraction), and in the zero-shot
// text->video branch `(1 - c0) * np.zeros(...)` is a float64 array, so the remaining arithmetic of that direction runs in
// float64.  Entries outside the candidate list are the reference's -100 fill.
struct FuseCoef {
  float alpha;     // fp32(alpha)
  float c_q;       // fp32(c0 | c1)
  float c_q_om;    // fp32(1 - c)          (subtraction done in double on the host, like Python)
  float c_e;       // fp32(c2 | c3)
  float c_e_om;    // fp32(1 - c2|c3)
  double c_e_d;    // double(c2)           (only used by the float64 branch)
  int use_prior;   // cand - alpha * prior          (args.cpn and the prior matrix exists)
  int use_query;   // c_q * query + (1 - c_q) * cpn (else blim = cpn: zero-shot video->text)
  int cpn_zero_f64;  // zero-shot text->video: cpn term is a float64 zeros matrix
};

__device__ __forceinline__ double fuse_one(const FuseCoef& k, float cand, float prior, float query, float iv2) {
  if (k.cpn_zero_f64) {
    // blim = c0 * query (fp32) + (1 - c0) * 0.0 (f64)  ->  f64;  c2 * blim (f64) + (1 - c2) * iv2 (fp32 product, promoted)
    const double b = static_cast<double>(__fmul_rn(k.c_q, query)) + 0.0;
    return __dadd_rn(__dmul_rn(k.c_e_d, b), static_cast<double>(__fmul_rn(k.c_e_om, iv2)));
  }
  float cpn = cand;
  if (k.use_prior) cpn = __fsub_rn(cand, __fmul_rn(k.alpha, prior));
  float b = cpn;
  if (k.use_query) b = __fadd_rn(__fmul_rn(k.c_q, query), __fmul_rn(k.c_q_om, cpn));
  return static_cast<double>(__fadd_rn(__fmul_rn(k.c_e, b), __fmul_rn(k.c_e_om, iv2)));
}

// One CTA per row.  Inputs: compact candidate arrays [rows, k] (column id + up to three likelihood terms) and the dense
// InternVideo2 row.  Outputs: fused score of every candidate, the candidates reordered by descending fused score, the
// rank of the ground-truth column (= row index, training_utils.py:146-147) in the full fused row, and a count of exact
// zeros (the reference's "matrix absent" guard, training_utils.py:174,195).
// Ties: equal scores are ordered by descending column index (what a stable ascending argsort reversed would give);
// np.argsort's default introsort does not pin tie order, so ties are outside the bit-exactness claim.
__global__ void fuse_rerank_kernel(const FuseCoef coef, const int* __restrict__ cand_idx, const float* __restrict__ cand,
                                   const float* __restrict__ prior, const float* __restrict__ query, const float* __restrict__ iv2,
                                   int n_rows, int n_cols, int k, int row0, double* __restrict__ fused_out, int* __restrict__ order_out,
                                   int* __restrict__ gt_rank_out, int* __restrict__ zero_count) {
  extern __shared__ uint8_t smem_fr[];
  double* s_f = reinterpret_cast<double*>(smem_fr);          // [k]
  int* s_idx = reinterpret_cast<int*>(s_f + k);              // [k]
  uint8_t* s_flag = reinterpret_cast<uint8_t*>(s_idx + k);   // [n_cols]
  __shared__ int s_cnt, s_zero;
  __shared__ double s_gt;
  const int r = blockIdx.x;
  if (r >= n_rows) return;
  const int gt = row0 + r;  // ground-truth column of this row
  const float* iv2_row = iv2 + static_cast<size_t>(r) * n_cols;
  for (int j = threadIdx.x; j < n_cols; j += blockDim.x) s_flag[j] = 0;
  if (threadIdx.x == 0) { s_cnt = 0; s_zero = 0; }
  __syncthreads();
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const size_t o = static_cast<size_t>(r) * k + j;
    const int col = cand_idx[o];
    const double f = fuse_one(coef, cand ? cand[o] : -100.f, prior ? prior[o] : -100.f, query ? query[o] : -100.f, iv2_row[col]);
    s_f[j] = f;
    s_idx[j] = col;
    s_flag[col] = 1;
    fused_out[o] = f;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double g;
    if (gt < n_cols && s_flag[gt]) {
      g = 0.0;
      for (int j = 0; j < k; ++j)
        if (s_idx[j] == gt) g = s_f[j];
    } else {
      g = fuse_one(coef, -100.f, -100.f, -100.f, iv2_row[gt < n_cols ? gt : 0]);
    }
    s_gt = g;
  }
  __syncthreads();
  const double g = s_gt;
  int cnt = 0, zeros = 0;
  for (int j = threadIdx.x; j < n_cols; j += blockDim.x) {
    if (s_flag[j]) continue;
    const double f = fuse_one(coef, -100.f, -100.f, -100.f, iv2_row[j]);
    cnt += (f > g) || (f == g && j > gt);
    zeros += (f == 0.0);
  }
  for (int j = threadIdx.x; j < k; j += blockDim.x) {
    const double f = s_f[j];
    const int col = s_idx[j];
    cnt += (f > g) || (f == g && col > gt);
    zeros += (f == 0.0);
    // position of candidate j among the candidates (descending score, ties by descending column)
    int pos = 0;
    for (int q = 0; q < k; ++q) {
      const double fq = s_f[q];
      pos += (fq > f) || (fq == f && s_idx[q] > col);
    }
    order_out[static_cast<size_t>(r) * k + pos] = col;
  }
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    zeros += __shfl_xor_sync(0xffffffffu, zeros, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&s_cnt, cnt);
    atomicAdd(&s_zero, zeros);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    gt_rank_out[r] = s_cnt;
    if (s_zero) atomicAdd(zero_count, s_zero);
  }
}

// Rank of the ground-truth column in a dense score row (get_recall on an arbitrary matrix): one CTA per row.
__global__ void rank_dense_kernel(const float* __restrict__ mat, int n_rows, int n_cols, int row0, int* __restrict__ gt_rank_out,
                                  int* __restrict__ zero_count) {
  __shared__ int s_cnt, s_zero;
  const int r = blockIdx.x;
  if (r >= n_rows) return;
  if (threadIdx.x == 0) { s_cnt = 0; s_zero = 0; }
  __syncthreads();
  const float* row = mat + static_cast<size_t>(r) * n_cols;
  const int gt = row0 + r;
  const float g = row[gt];
  int cnt = 0, zeros = 0;
  for (int j = threadIdx.x; j < n_cols; j += blockDim.x) {
    const float f = row[j];
    cnt += (f > g) || (f == g && j > gt);
    zeros += (f == 0.f);
  }
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    zeros += __shfl_xor_sync(0xffffffffu, zeros, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&s_cnt, cnt);
    atomicAdd(&s_zero, zeros);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    gt_rank_out[r] = s_cnt;
    if (s_zero) atomicAdd(zero_count, s_zero);
  }
}

}  // namespace blim
