// B200-native video feature extractor (include/blim_vision.h): UMT ViT encoder + ToMe token merging.
//
// Reference: extract.py:96-110 -> VideoChatFlashQwenForCausalLM.encode_video_image(return_video_feature=True)
// (modeling_videochat_flash.py:126-181) -> UMTVisionTower.forward (vision_tower_builder.py:564-577) ->
// PretrainVisionTransformerEncoder.forward_features (vision_tower_builder.py:329-348) -> ToMe16_mlp_hd64.forward /
// merge_tokens (mm_projector_builder.py:101-154).
//
// Mapping onto the engine's kernels:
//   Conv3d patch embedding (kernel = stride = (1, P, P))   patchify_kernel (im2col) + tcgen05 GEMM, fp32 out + bias
//   + sinusoid position table                               add_pos_kernel
//   LayerNorm (norm1 / norm2 / vision_layernorm)            layernorm_kernel (fp32 statistics, bf16 or fp32 out)
//   qkv = Linear(C, 3C) with (q_bias, 0, v_bias)            tcgen05 GEMM, EpiStore<bf16, bias> into a packed [M, 3C] buffer
//   softmax(q k^T / sqrt(d)) v, all tokens of a clip        attention_tc2_kernel: the clip's K/V rows are the "shared prefix"
//                                                           segment (every key visible), no own-run keys; Q / K / V are
//                                                           read straight from the packed buffer (strided TMA maps)
//   x += proj(attn) + b, x += fc2(gelu(fc1(norm2 x)))       EpiResidBias / EpiStore<bf16, bias, GELU>, fp32 residual stream
//   bipartite soft matching + weighted merge (ToMe)         fp32 CUDA-core kernels (tome_*): index decisions are made in
//                                                           fp32 like the CPU reference, so they are reproducible bit for bit
// included at the end of engine.cu (one translation unit)
#pragma once

#include "../../include/blim_vision.h"

namespace {

// ------------------------------------------------------------------------------------------------ kernels
// frames [n_frames, 3, S, S] -> patches bf16 [n_frames * Wp * Wp, 3 * P * P]; column k = (channel, dy, dx): the flattened
// Conv3d weight [C_out, 3, 1, P, P] is the matching K-major B operand (vision_tower_builder.py:174-178, 187).
__global__ void patchify_kernel(bf16* __restrict__ out, const void* __restrict__ frames, int dtype, int n_frames, int S, int P) {
  const int Wp = S / P, K = 3 * P * P;
  const size_t n = static_cast<size_t>(n_frames) * Wp * Wp * K;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i % K);
    const size_t tok = i / K;
    const int px = static_cast<int>(tok % Wp), py = static_cast<int>((tok / Wp) % Wp);
    const size_t f = tok / (static_cast<size_t>(Wp) * Wp);
    const int c = k / (P * P), dy = (k / P) % P, dx = k % P;
    const size_t src = ((f * 3 + c) * S + py * P + dy) * S + px * P + dx;
    out[i] = __float2bfloat16(load_as_f32(frames, dtype, src));
  }
}

// x[r, :] += pos[r % rows_per_clip, :]   (vision_tower_builder.py:335)
__global__ void add_pos_kernel(float* __restrict__ x, const float* __restrict__ pos, size_t n_rows, int rows_per_clip, int C) {
  const size_t n4 = n_rows * C / 4;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t r = (i * 4) / C;
    const int c = static_cast<int>((i * 4) % C);
    float4 v = reinterpret_cast<float4*>(x)[i];
    const float4 p = *reinterpret_cast<const float4*>(pos + (r % rows_per_clip) * C + c);
    v.x += p.x; v.y += p.y; v.z += p.z; v.w += p.w;
    reinterpret_cast<float4*>(x)[i] = v;
  }
}

__device__ __forceinline__ float block_sum_256(float v, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();   // sh may still be read from a previous call
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) t += sh[w];
  return t;
}

// nn.LayerNorm over the last dimension (biased variance, fp32 statistics): one WARP per row, the row lives in registers as
// NV float4 per lane (C = 128 * NV), 8 rows per 256-thread CTA, no shared memory and no block barrier.
template <typename OutT, int NV>
__global__ void __launch_bounds__(256) layernorm_warp_kernel(OutT* __restrict__ out, const float* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ b, int rows, float eps) {
  constexpr int C = 128 * NV;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float4* src = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * C);
  float4 v[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = src[lane + 32 * i];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
    q += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / C + eps);
  const float4* w4 = reinterpret_cast<const float4*>(w);
  const float4* b4 = reinterpret_cast<const float4*>(b);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 g = __ldg(w4 + lane + 32 * i), h = __ldg(b4 + lane + 32 * i);
    const float y0 = v[i].x * rstd * g.x + h.x, y1 = v[i].y * rstd * g.y + h.y, y2 = v[i].z * rstd * g.z + h.z, y3 = v[i].w * rstd * g.w + h.w;
    if constexpr (sizeof(OutT) == 2) {
      uint2 u;
      u.x = Fmt16<__nv_bfloat16>::pack2(y0, y1);
      u.y = Fmt16<__nv_bfloat16>::pack2(y2, y3);
      reinterpret_cast<uint2*>(out + static_cast<size_t>(row) * C)[lane + 32 * i] = u;
    } else {
      reinterpret_cast<float4*>(out + static_cast<size_t>(row) * C)[lane + 32 * i] = make_float4(y0, y1, y2, y3);
    }
  }
}

// generic fallback (any C <= 4096): one 256-thread CTA per row
template <typename OutT>
__global__ void __launch_bounds__(256) layernorm_kernel(OutT* __restrict__ out, const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ b, int C, float eps) {
  __shared__ float sh[8];
  const float* row = x + static_cast<size_t>(blockIdx.x) * C;
  float v[16];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int c = threadIdx.x + i * 256;
    v[i] = c < C ? row[c] : 0.f;
    s += v[i];
  }
  const float mean = block_sum_256(s, sh) / C;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int c = threadIdx.x + i * 256;
    const float d = c < C ? v[i] - mean : 0.f;
    q += d * d;
  }
  const float rstd = rsqrtf(block_sum_256(q, sh) / C + eps);
  OutT* dst = out + static_cast<size_t>(blockIdx.x) * C;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int c = threadIdx.x + i * 256;
    if (c < C) {
      const float y = (v[i] - mean) * rstd * w[c] + b[c];
      if constexpr (sizeof(OutT) == 2) dst[c] = __float2bfloat16(y); else dst[c] = y;
    }
  }
}

template <typename OutT>
static void launch_layernorm(OutT* out, const float* x, const float* w, const float* b, int rows, int C, float eps, cudaStream_t st) {
  const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
  if (C == 1024) layernorm_warp_kernel<OutT, 8><<<grid, 256, 0, st>>>(out, x, w, b, rows, eps);
  else if (C == 128) layernorm_warp_kernel<OutT, 1><<<grid, 256, 0, st>>>(out, x, w, b, rows, eps);
  else if (C == 768) layernorm_warp_kernel<OutT, 6><<<grid, 256, 0, st>>>(out, x, w, b, rows, eps);
  else layernorm_kernel<OutT><<<rows, 256, 0, st>>>(out, x, w, b, C, eps);
}

// ---- ToMe (mm_projector_builder.py:6-130), one grid.y slice per clip
// metric = x.reshape(p, heads, dim).mean(1), then metric / metric.norm(-1)   (mm_projector_builder.py:121, 25); dim <= 128.
__global__ void tome_metric_kernel(float* __restrict__ metric, const float* __restrict__ x, int p, int C, int heads) {
  const int dim = C / heads;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= p) return;
  const float* row = x + (static_cast<size_t>(blockIdx.y) * p + warp) * C;
  float m[4];
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int d = lane + 32 * k;
    float acc = 0.f;
    if (d < dim)
      for (int h = 0; h < heads; ++h) acc += row[h * dim + d];
    m[k] = acc / heads;
    ss += m[k] * m[k];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float nrm = sqrtf(ss);
  float* dst = metric + (static_cast<size_t>(blockIdx.y) * p + warp) * dim;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int d = lane + 32 * k;
    if (d < dim) dst[d] = m[k] / nrm;
  }
}

// node_max[i], node_idx[i] = max / argmax over odd tokens j of <metric[2i], metric[2j+1]>  (mm_projector_builder.py:26-29);
// equal scores keep the lowest j.  CTA = 16 even tokens x 16 lanes, odd tokens stream through shared memory 64 at a time.
template <int DIM>
__global__ void __launch_bounds__(256) tome_match_kernel(float* __restrict__ node_max, int* __restrict__ node_idx, const float* __restrict__ metric,
                                                         int p) {
  __shared__ float sa[16][DIM];
  __shared__ float sb[64][DIM + 1];
  const int na = (p + 1) / 2, nb = p / 2;
  const float* base = metric + static_cast<size_t>(blockIdx.y) * p * DIM;
  const int ai = threadIdx.x >> 4, l16 = threadIdx.x & 15;
  const int i0 = blockIdx.x * 16;
  for (int e = threadIdx.x; e < 16 * DIM; e += 256) {
    const int r = e / DIM, d = e % DIM;
    sa[r][d] = (i0 + r < na) ? base[static_cast<size_t>(2 * (i0 + r)) * DIM + d] : 0.f;
  }
  float best = -INFINITY;
  int best_j = 0x7fffffff;
  for (int j0 = 0; j0 < nb; j0 += 64) {
    __syncthreads();
    for (int e = threadIdx.x; e < 64 * DIM; e += 256) {
      const int r = e / DIM, d = e % DIM;
      sb[r][d] = (j0 + r < nb) ? base[static_cast<size_t>(2 * (j0 + r) + 1) * DIM + d] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int jl = l16 + 16 * k;
      float acc = 0.f;
#pragma unroll 16
      for (int d = 0; d < DIM; ++d) acc = fmaf(sa[ai][d], sb[jl][d], acc);
      const int j = j0 + jl;
      if (j < nb && acc > best) { best = acc; best_j = j; }   // j ascends inside a thread: strict > keeps the first maximum
    }
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oj = __shfl_xor_sync(0xffffffffu, best_j, o);
    if (ov > best || (ov == best && oj < best_j)) { best = ov; best_j = oj; }
  }
  if (l16 == 0 && i0 + ai < na) {
    node_max[static_cast<size_t>(blockIdx.y) * na + i0 + ai] = best;
    node_idx[static_cast<size_t>(blockIdx.y) * na + i0 + ai] = best_j;
  }
}

// edge_idx = argsort(node_max, descending) with ties in index order (mm_projector_builder.py:30), then the r merged sources
// (sorted positions [0, r)) grouped by their merge target dst = node_idx[edge_idx[q]] (:34): csr_src lists the positions q
// ordered by (dst, q), csr_start / csr_end delimit every odd token's group -- scatter_add visits the sources of a
// destination in ascending q, and so does the merge kernel.  One CTA per clip, bitonic networks over n_sort >= na keys.
__global__ void __launch_bounds__(1024) tome_sort_kernel(int* __restrict__ edge, int* __restrict__ csr_src, int* __restrict__ csr_start,
                                                         int* __restrict__ csr_end, const float* __restrict__ node_max,
                                                         const int* __restrict__ node_idx, int na, int nb, int r, int n_sort) {
  extern __shared__ uint8_t sort_smem[];
  float* key = reinterpret_cast<float*>(sort_smem);
  int* idx = reinterpret_cast<int*>(sort_smem + static_cast<size_t>(n_sort) * 4);
  const float* nm = node_max + static_cast<size_t>(blockIdx.x) * na;
  for (int i = threadIdx.x; i < n_sort; i += blockDim.x) {
    key[i] = i < na ? nm[i] : -INFINITY;
    idx[i] = i < na ? i : 0x7fffffff;
  }
  __syncthreads();
  // order: a before b  <=>  key_a > key_b  or  (key_a == key_b and idx_a < idx_b)
  for (int k = 2; k <= n_sort; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < n_sort / 2; t += blockDim.x) {
        const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int hi = lo | j;
        const bool up = (lo & k) == 0;   // ascending position = "before" order in this sub-sequence
        const float ka = key[lo], kb = key[hi];
        const int ia = idx[lo], ib = idx[hi];
        const bool a_first = ka > kb || (ka == kb && ia < ib);
        if (a_first != up) { key[lo] = kb; key[hi] = ka; idx[lo] = ib; idx[hi] = ia; }
      }
      __syncthreads();
    }
  }
  const int* ni = node_idx + static_cast<size_t>(blockIdx.x) * na;
  int* key2 = reinterpret_cast<int*>(key);   // (dst << 14 | q) of the merged sources, INT_MAX padding
  for (int i = threadIdx.x; i < n_sort; i += blockDim.x) {
    if (i < na) edge[static_cast<size_t>(blockIdx.x) * na + i] = idx[i];
    key2[i] = i < r ? ((ni[idx[i]] << 14) | i) : 0x7fffffff;
  }
  for (int j = threadIdx.x; j < nb; j += blockDim.x) {
    csr_start[static_cast<size_t>(blockIdx.x) * nb + j] = 0;
    csr_end[static_cast<size_t>(blockIdx.x) * nb + j] = 0;
  }
  __syncthreads();
  for (int k = 2; k <= n_sort; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < n_sort / 2; t += blockDim.x) {
        const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int hi = lo | j;
        const bool up = (lo & k) == 0;
        const int ka = key2[lo], kb = key2[hi];
        if ((ka < kb) != up) { key2[lo] = kb; key2[hi] = ka; }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < r; i += blockDim.x) {
    const int kv = key2[i], j = kv >> 14;
    csr_src[static_cast<size_t>(blockIdx.x) * na + i] = kv & 16383;
    if (i == 0 || (key2[i - 1] >> 14) != j) csr_start[static_cast<size_t>(blockIdx.x) * nb + j] = i;
    if (i == r - 1 || (key2[i + 1] >> 14) != j) csr_end[static_cast<size_t>(blockIdx.x) * nb + j] = i + 1;
  }
}

// merge_wavg (mm_projector_builder.py:61-77) for one round: out rows [0, na - r) = the unmerged even tokens in sorted
// order (x * size) / size, rows [na - r, p - r) = every odd token plus the even tokens merged into it (scatter_add in
// sorted order, :40-43), divided by the merged size.  size_in == nullptr: all ones (first round).
__global__ void __launch_bounds__(256) tome_merge_kernel(float* __restrict__ x_out, float* __restrict__ size_out, const float* __restrict__ x_in,
                                                         const float* __restrict__ size_in, const int* __restrict__ edge,
                                                         const int* __restrict__ csr_src, const int* __restrict__ csr_start,
                                                         const int* __restrict__ csr_end, int p, int r, int C) {
  const int na = (p + 1) / 2, nb = p / 2;
  const int clip = blockIdx.y, k = blockIdx.x;
  const float* xin = x_in + static_cast<size_t>(clip) * p * C;
  const float* sin_ = size_in ? size_in + static_cast<size_t>(clip) * p : nullptr;
  const int* eg = edge + static_cast<size_t>(clip) * na;
  float* xo = x_out + (static_cast<size_t>(clip) * (p - r) + k) * C;
  if (k < na - r) {
    const int tok = 2 * eg[r + k];
    const float s = sin_ ? sin_[tok] : 1.f;
    for (int c = threadIdx.x; c < C; c += 256) xo[c] = (xin[static_cast<size_t>(tok) * C + c] * s) / s;
    if (threadIdx.x == 0) size_out[static_cast<size_t>(clip) * (p - r) + k] = s;
    return;
  }
  const int j = k - (na - r);
  const int q0 = csr_start[static_cast<size_t>(clip) * nb + j], q1 = csr_end[static_cast<size_t>(clip) * nb + j];
  const int* src = csr_src + static_cast<size_t>(clip) * na;
  const int tok = 2 * j + 1;
  const float s0 = sin_ ? sin_[tok] : 1.f;
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int c = threadIdx.x + 256 * i;
    acc[i] = c < C ? xin[static_cast<size_t>(tok) * C + c] * s0 : 0.f;
  }
  float ssum = s0;
  for (int t = q0; t < q1; ++t) {   // ascending sorted position: the order scatter_add accumulates in
    const int st = 2 * eg[src[t]];
    const float ss = sin_ ? sin_[st] : 1.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int c = threadIdx.x + 256 * i;
      if (c < C) acc[i] += xin[static_cast<size_t>(st) * C + c] * ss;
    }
    ssum += ss;
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int c = threadIdx.x + 256 * i;
    if (c < C) xo[c] = acc[i] / ssum;
  }
  if (threadIdx.x == 0) size_out[static_cast<size_t>(clip) * (p - r) + k] = ssum;
}

__global__ void f32_to_dtype_kernel(void* __restrict__ out, const float* __restrict__ in, int dtype, size_t n) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    if (dtype == 0) reinterpret_cast<float*>(out)[i] = in[i];
    else if (dtype == 1) reinterpret_cast<bf16*>(out)[i] = __float2bfloat16(in[i]);
    else reinterpret_cast<__half*>(out)[i] = __float2half(in[i]);
  }
}

struct VisBlockW {
  DevBuf ln1_w, ln1_b, ln2_w, ln2_b, w_qkv, b_qkv, w_proj, b_proj, w_fc1, b_fc1, w_fc2, b_fc2;
};

}  // namespace

struct blim_vision {
  blim_vision_cfg cfg;
  int device = 0;
  int S, P, Wp, L, FPC, C, NL, NH, DH, F, KP, TL;   // TL = tokens per clip = FPC * L
  int max_clips = 16;
  int attn_version = 2;   // BLIM_VIS_ATTN=2|3|4: attention_tc2 / tc3 / tc4 kernel
  std::string err;
  GemmLaunchCtx gemm;
  DevBuf w_patch, b_patch, lnf_w, lnf_b, pos;
  bool pos_set = false;
  std::vector<VisBlockW> blocks;
  // workspaces
  DevBuf patches, x, xn, qkv, attn, act, feat, works;
  DevBuf t_x[2], t_size[2], t_metric, t_nmax, t_nidx, t_edge, t_dst, t_cstart, t_cend;
  int works_clips = 0;
  int n_works = 0;
  int64_t launches = 0;
  double flops = 0.0;
  bool profiling = false;
  struct Timed { cudaEvent_t a, b; int cat; };
  std::vector<Timed> timed;
  std::vector<cudaEvent_t> pool;
  cudaEvent_t ev() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
  }
  void tic(int cat, cudaStream_t st) {
    if (!profiling) return;
    Timed t; t.a = ev(); t.b = ev(); t.cat = cat;
    cudaEventRecord(t.a, st);
    timed.push_back(t);
  }
  void toc(cudaStream_t st) {
    if (profiling) cudaEventRecord(timed.back().b, st);
  }
  int fail(const std::string& m) { err = m; return 1; }
  int fail_cuda(const char* what, cudaError_t e) { err = std::string(what) + ": " + cudaGetErrorString(e); return 1; }
};

static std::string g_vision_create_error;

#define VCK(expr)                                                  \
  do {                                                             \
    cudaError_t _e = (expr);                                       \
    if (_e != cudaSuccess) return v->fail_cuda(#expr, _e);         \
  } while (0)
#define VCL()                                                      \
  do {                                                             \
    v->launches++;                                                 \
    cudaError_t _e = cudaGetLastError();                           \
    if (_e != cudaSuccess) return v->fail_cuda("kernel launch", _e); \
  } while (0)

extern "C" const char* blim_vision_last_error(const blim_vision* v) { return v ? v->err.c_str() : g_vision_create_error.c_str(); }

extern "C" void blim_vision_destroy(blim_vision* v) {
  if (!v) return;
  cudaSetDevice(v->device);
  DevBuf* bufs[] = {&v->w_patch, &v->b_patch, &v->lnf_w, &v->lnf_b, &v->pos, &v->patches, &v->x, &v->xn, &v->qkv, &v->attn, &v->act, &v->feat,
                    &v->works, &v->t_x[0], &v->t_x[1], &v->t_size[0], &v->t_size[1], &v->t_metric, &v->t_nmax, &v->t_nidx, &v->t_edge, &v->t_dst, &v->t_cstart, &v->t_cend};
  for (DevBuf* b : bufs) b->release();
  for (auto& bw : v->blocks) {
    DevBuf* wb[] = {&bw.ln1_w, &bw.ln1_b, &bw.ln2_w, &bw.ln2_b, &bw.w_qkv, &bw.b_qkv, &bw.w_proj, &bw.b_proj, &bw.w_fc1, &bw.b_fc1, &bw.w_fc2, &bw.b_fc2};
    for (DevBuf* b : wb) b->release();
  }
  for (auto& t : v->timed) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
  for (auto e : v->pool) cudaEventDestroy(e);
  delete v;
}

extern "C" int blim_vision_create(const blim_vision_cfg* cfg, int device, blim_vision** out) {
  if (!cfg || !out) { g_vision_create_error = "null argument"; return 1; }
  *out = nullptr;
  auto bad = [&](const std::string& m) { g_vision_create_error = m; return 1; };
  if (cfg->patch_size <= 0 || cfg->image_size <= 0 || cfg->image_size % cfg->patch_size) return bad("image_size must be a multiple of patch_size");
  if (cfg->hidden_size <= 0 || cfg->hidden_size % 64 || cfg->hidden_size > 4096) return bad("hidden_size must be a multiple of 64, at most 4096");
  if (cfg->num_heads <= 0 || cfg->hidden_size % cfg->num_heads) return bad("hidden_size must be a multiple of num_heads");
  const int dh = cfg->hidden_size / cfg->num_heads;
  if (dh != 64 && dh != 128) return bad("head_dim must be 64 or 128");
  if (cfg->mlp_hidden_size <= 0 || cfg->mlp_hidden_size % 64) return bad("mlp_hidden_size must be a multiple of 64");
  if ((3 * cfg->patch_size * cfg->patch_size) % 64) return bad("3 * patch_size^2 must be a multiple of 64");
  if (cfg->frames_per_clip <= 0 || cfg->num_layers <= 0 || cfg->tome_tokens_per_frame <= 0) return bad("bad frames_per_clip / num_layers / tome_tokens_per_frame");
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device < 0 || device >= n_dev) return bad("no such CUDA device (the extractor has no CPU fallback)");
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return bad("cudaGetDeviceProperties failed");
  if (prop.major != 10) return bad("blim_vision needs an sm_100 (B200) device: the kernels are tcgen05 / TMEM code");
  if (cudaSetDevice(device) != cudaSuccess) return bad("cudaSetDevice failed");
  blim_vision* v = new blim_vision();
  v->cfg = *cfg;
  v->device = device;
  v->S = cfg->image_size; v->P = cfg->patch_size; v->Wp = v->S / v->P; v->L = v->Wp * v->Wp; v->FPC = cfg->frames_per_clip;
  v->C = cfg->hidden_size; v->NL = cfg->num_layers; v->NH = cfg->num_heads; v->DH = dh; v->F = cfg->mlp_hidden_size;
  v->KP = 3 * v->P * v->P; v->TL = v->FPC * v->L;
  v->max_clips = cfg->max_clips > 0 ? cfg->max_clips : 16;
  if (v->TL <= cfg->tome_tokens_per_frame * v->FPC) { delete v; return bad("a clip must have more tokens than the merging target (mm_projector_builder.py:110)"); }
  v->gemm.num_sms = prop.multiProcessorCount;
  v->gemm.device = device;
  v->gemm.cta_group = 2;
  v->attn_version = dh == 64 ? kAttnIssueWarp : kAttnPerItem;   // v4 (issue warp, two P tiles) wins at head_dim 64; at 128 its 96-register budget spills
  if (const char* a = getenv("BLIM_VIS_ATTN")) v->attn_version = atoi(a) == 4 ? kAttnIssueWarp : atoi(a) == 2 ? kAttnPerItem : v->attn_version;
  v->blocks.resize(v->NL);
  const size_t M = static_cast<size_t>(v->max_clips) * v->TL;
  const size_t na = (v->TL + 1) / 2;
  struct { DevBuf* b; size_t bytes; } ws[] = {
      {&v->patches, M * v->KP * 2}, {&v->x, M * v->C * 4}, {&v->xn, M * v->C * 2}, {&v->qkv, M * 3 * v->C * 2}, {&v->attn, M * v->C * 2},
      {&v->act, M * v->F * 2}, {&v->feat, M * v->C * 4}, {&v->pos, static_cast<size_t>(v->TL) * v->C * 4},
      {&v->t_x[0], M * v->C * 4}, {&v->t_x[1], M * v->C * 4}, {&v->t_size[0], M * 4}, {&v->t_size[1], M * 4},
      {&v->t_metric, M * (v->C / v->NH) * 4}, {&v->t_nmax, v->max_clips * na * 4}, {&v->t_nidx, v->max_clips * na * 4},
      {&v->t_edge, v->max_clips * na * 4}, {&v->t_dst, v->max_clips * na * 4}, {&v->t_cstart, v->max_clips * na * 4},
      {&v->t_cend, v->max_clips * na * 4}};
  for (auto& w : ws) {
    if (w.b->reserve(w.bytes) != cudaSuccess) {
      g_vision_create_error = "out of device memory for the extractor workspaces";
      blim_vision_destroy(v);
      return 1;
    }
  }
  *out = v;
  return 0;
}

// ------------------------------------------------------------------------------------------------ weights
static int vis_copy_bf16(blim_vision* v, DevBuf& dst, const void* src, int dtype, size_t rows, size_t cols, cudaStream_t st) {
  VCK(dst.reserve(rows * cols * 2));
  repack_rows_16_kernel<bf16><<<1024, 256, 0, st>>>(dst.as<bf16>(), src, dtype, static_cast<int>(rows), static_cast<int>(cols), 0, 0, 0);
  VCL();
  return 0;
}
static int vis_copy_f32(blim_vision* v, DevBuf& dst, size_t dst_off, size_t total, const void* src, int dtype, size_t n, cudaStream_t st) {
  if (dst.cap < total * 4) {
    VCK(dst.reserve(total * 4));
    VCK(cudaMemsetAsync(dst.p, 0, total * 4, st));
  }
  repack_f32_kernel<<<static_cast<unsigned>(std::min<size_t>((n + 255) / 256, 1024)), 256, 0, st>>>(dst.as<float>() + dst_off, src, dtype, n, 0);
  VCL();
  return 0;
}

extern "C" int blim_vision_load_weight(blim_vision* v, const char* name, const void* dev_ptr, int dtype, const int64_t* shape, int ndim, void* stream) {
  if (!v) return 1;
  if (!name || !dev_ptr || !shape || ndim <= 0 || dtype < 0 || dtype > 2) return v->fail("bad load_weight arguments");
  VCK(cudaSetDevice(v->device));
  cudaStream_t st = S(stream);
  std::string n(name);
  const std::string prefix = "model.vision_tower.vision_tower.";
  if (n.compare(0, prefix.size(), prefix) == 0) n = n.substr(prefix.size());
  size_t numel = 1;
  for (int i = 0; i < ndim; ++i) numel *= static_cast<size_t>(shape[i]);
  const size_t C = v->C, F = v->F;
  auto want = [&](size_t expect) { return numel == expect ? 0 : v->fail("shape mismatch for " + n); };
  if (n == "encoder.patch_embed.proj.weight") { if (want(C * v->KP)) return 1; return vis_copy_bf16(v, v->w_patch, dev_ptr, dtype, C, v->KP, st); }
  if (n == "encoder.patch_embed.proj.bias") { if (want(C)) return 1; return vis_copy_f32(v, v->b_patch, 0, C, dev_ptr, dtype, C, st); }
  if (n == "encoder.vision_layernorm.weight") { if (want(C)) return 1; return vis_copy_f32(v, v->lnf_w, 0, C, dev_ptr, dtype, C, st); }
  if (n == "encoder.vision_layernorm.bias") { if (want(C)) return 1; return vis_copy_f32(v, v->lnf_b, 0, C, dev_ptr, dtype, C, st); }
  const std::string bp = "encoder.blocks.";
  if (n.compare(0, bp.size(), bp) == 0) {
    const size_t dot = n.find('.', bp.size());
    if (dot == std::string::npos) return v->fail("unknown parameter " + n);
    const int li = atoi(n.substr(bp.size(), dot - bp.size()).c_str());
    if (li >= v->NL) return 0;   // blocks behind mm_vision_select_layer are never run (vision_tower_builder.py:289)
    if (li < 0) return v->fail("unknown parameter " + n);
    VisBlockW& w = v->blocks[li];
    const std::string rest = n.substr(dot + 1);
    if (rest == "norm1.weight") { if (want(C)) return 1; return vis_copy_f32(v, w.ln1_w, 0, C, dev_ptr, dtype, C, st); }
    if (rest == "norm1.bias") { if (want(C)) return 1; return vis_copy_f32(v, w.ln1_b, 0, C, dev_ptr, dtype, C, st); }
    if (rest == "norm2.weight") { if (want(C)) return 1; return vis_copy_f32(v, w.ln2_w, 0, C, dev_ptr, dtype, C, st); }
    if (rest == "norm2.bias") { if (want(C)) return 1; return vis_copy_f32(v, w.ln2_b, 0, C, dev_ptr, dtype, C, st); }
    if (rest == "attn.qkv.weight") { if (want(3 * C * C)) return 1; return vis_copy_bf16(v, w.w_qkv, dev_ptr, dtype, 3 * C, C, st); }
    if (rest == "attn.q_bias") { if (want(C)) return 1; return vis_copy_f32(v, w.b_qkv, 0, 3 * C, dev_ptr, dtype, C, st); }
    if (rest == "attn.v_bias") { if (want(C)) return 1; return vis_copy_f32(v, w.b_qkv, 2 * C, 3 * C, dev_ptr, dtype, C, st); }
    if (rest == "attn.proj.weight") { if (want(C * C)) return 1; return vis_copy_bf16(v, w.w_proj, dev_ptr, dtype, C, C, st); }
    if (rest == "attn.proj.bias") { if (want(C)) return 1; return vis_copy_f32(v, w.b_proj, 0, C, dev_ptr, dtype, C, st); }
    if (rest == "mlp.fc1.weight") { if (want(F * C)) return 1; return vis_copy_bf16(v, w.w_fc1, dev_ptr, dtype, F, C, st); }
    if (rest == "mlp.fc1.bias") { if (want(F)) return 1; return vis_copy_f32(v, w.b_fc1, 0, F, dev_ptr, dtype, F, st); }
    if (rest == "mlp.fc2.weight") { if (want(C * F)) return 1; return vis_copy_bf16(v, w.w_fc2, dev_ptr, dtype, C, F, st); }
    if (rest == "mlp.fc2.bias") { if (want(C)) return 1; return vis_copy_f32(v, w.b_fc2, 0, C, dev_ptr, dtype, C, st); }
  }
  return v->fail("unknown parameter " + n);
}

extern "C" int blim_vision_set_pos_embed(blim_vision* v, const float* table_dev, int rows, void* stream) {
  if (!v) return 1;
  if (!table_dev || rows != v->TL) return v->fail("position table must have frames_per_clip * (image_size / patch_size)^2 rows");
  VCK(cudaSetDevice(v->device));
  VCK(cudaMemcpyAsync(v->pos.p, table_dev, static_cast<size_t>(rows) * v->C * 4, cudaMemcpyDeviceToDevice, S(stream)));
  v->pos_set = true;
  return 0;
}

// ------------------------------------------------------------------------------------------------ encoder
template <class Epi>
static int vis_gemm(blim_vision* v, const bf16* A, int lda, const bf16* W, int ldw, int M, int N, int K, const typename Epi::Params& p, cudaStream_t st) {
  v->tic(0, st);
  cudaError_t r = launch_gemm<Epi>(v->gemm, A, lda, W, ldw, M, N, K, p, st);
  v->toc(st);
  if (r != cudaSuccess) return v->fail_cuda("tcgen05 gemm launch", r);
  v->flops += 2.0 * M * static_cast<double>(N) * K;
  return 0;
}

static int vis_check_weights(blim_vision* v) {
  if (!v->w_patch.p || !v->b_patch.p || !v->lnf_w.p || !v->lnf_b.p) return v->fail("patch embedding / vision_layernorm weights not loaded");
  if (!v->pos_set) return v->fail("position table not set (blim_vision_set_pos_embed)");
  for (int l = 0; l < v->NL; ++l) {
    VisBlockW& w = v->blocks[l];
    if (!w.ln1_w.p || !w.ln1_b.p || !w.ln2_w.p || !w.ln2_b.p || !w.w_qkv.p || !w.w_proj.p || !w.b_proj.p || !w.w_fc1.p || !w.b_fc1.p || !w.w_fc2.p ||
        !w.b_fc2.p)
      return v->fail("weights of encoder block " + std::to_string(l) + " not loaded");
    if (!w.b_qkv.p) {   // qkv_bias=False checkpoints: no q_bias / v_bias
      if (w.b_qkv.reserve(static_cast<size_t>(3) * v->C * 4) != cudaSuccess) return v->fail("out of device memory");
      if (cudaMemset(w.b_qkv.p, 0, static_cast<size_t>(3) * v->C * 4) != cudaSuccess) return v->fail("cudaMemset failed");
    }
  }
  return 0;
}

// attention work items: every 128-token block of a clip attends to all TL tokens of its clip ("prefix" segment only)
static int vis_works(blim_vision* v, int n_clips, cudaStream_t st) {
  if (v->works_clips == n_clips) return 0;
  std::vector<AttnWorkTc> works;
  for (int c = 0; c < n_clips; ++c)
    for (int t0 = 0; t0 < v->TL; t0 += 128) {
      AttnWorkTc w;
      w.tok0 = c * v->TL + t0;
      w.n_tok = std::min(128, v->TL - t0);
      w.a_start = c * v->TL;
      w.a_len = v->TL;
      w.kb0 = w.tok0 + w.n_tok;   // no own-run keys
      w.b_off = 0;
      w.pad1 = w.pad2 = 0;
      works.push_back(w);
    }
  VCK(v->works.reserve(works.size() * sizeof(AttnWorkTc)));
  VCK(cudaMemcpyAsync(v->works.p, works.data(), works.size() * sizeof(AttnWorkTc), cudaMemcpyHostToDevice, st));
  VCK(cudaStreamSynchronize(st));   // `works` is a host temporary
  v->n_works = static_cast<int>(works.size());
  v->works_clips = n_clips;
  return 0;
}

static int vis_encode(blim_vision* v, const void* frames, int dtype, int n_frames, float* feat_out, cudaStream_t st) {
  if (!frames || !feat_out || dtype < 0 || dtype > 2) return v->fail("bad encode arguments");
  if (n_frames <= 0 || n_frames % v->FPC) return v->fail("n_frames must be a positive multiple of frames_per_clip");
  const int n_clips = n_frames / v->FPC;
  if (n_clips > v->max_clips) return v->fail("too many clips for the workspace (max_clips)");
  if (vis_check_weights(v)) return 1;
  if (vis_works(v, n_clips, st)) return 1;
  const int M = n_clips * v->TL, C = v->C, F = v->F;
  float* x = v->x.as<float>();
  bf16* xn = v->xn.as<bf16>();
  bf16* qkv = v->qkv.as<bf16>();

  v->tic(3, st);
  patchify_kernel<<<2048, 256, 0, st>>>(v->patches.as<bf16>(), frames, dtype, n_frames, v->S, v->P);
  v->toc(st);
  VCL();
  {
    EpiStore<float, true, false>::Params p{x, C, v->b_patch.as<float>()};
    if (vis_gemm<EpiStore<float, true, false>>(v, v->patches.as<bf16>(), v->KP, v->w_patch.as<bf16>(), v->KP, M, C, v->KP, p, st)) return 1;
  }
  v->tic(3, st);
  add_pos_kernel<<<2048, 256, 0, st>>>(x, v->pos.as<float>(), static_cast<size_t>(M), v->TL, C);
  v->toc(st);
  VCL();

  AttnTcMaps maps;
  if (!make_tmap_bf16(&maps.ka, qkv + C, static_cast<uint64_t>(M), static_cast<uint64_t>(C), static_cast<uint64_t>(3 * C), kTcKeys) ||
      !make_tmap_bf16(&maps.va, qkv + 2 * C, static_cast<uint64_t>(M), static_cast<uint64_t>(C), static_cast<uint64_t>(3 * C), kTcKeys))
    return v->fail("cuTensorMapEncodeTiled failed for the packed QKV buffer");
  maps.kb = maps.ka;
  maps.vb = maps.va;
  AttnParamsTc ap;
  ap.q = qkv; ap.o = v->attn.as<bf16>();
  ap.a_row0 = 0; ap.b_row0 = 0;
  ap.key_valid = nullptr; ap.tok_seq_start = nullptr;
  ap.works = v->works.as<AttnWorkTc>();
  ap.n_q = C; ap.n_kv = C; ap.group = 1;
  ap.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(v->DH));   // qk_scale = head_dim^-0.5 (vision_tower_builder.py:77)
  ap.q_stride = 3 * C;
  ap.n_works = v->n_works; ap.n_kv_heads = v->NH;

  for (int l = 0; l < v->NL; ++l) {
    const VisBlockW& w = v->blocks[l];
    v->tic(2, st);
    launch_layernorm<bf16>(xn, x, w.ln1_w.as<float>(), w.ln1_b.as<float>(), M, C, v->cfg.ln_eps, st);
    v->toc(st);
    VCL();
    {
      EpiStore<bf16, true, false>::Params p{qkv, 3 * C, w.b_qkv.as<float>()};
      if (vis_gemm<EpiStore<bf16, true, false>>(v, xn, C, w.w_qkv.as<bf16>(), C, M, 3 * C, C, p, st)) return 1;
    }
    v->tic(1, st);
    cudaError_t r = launch_attention_tc<bf16>(maps, ap, v->n_works, v->NH, v->DH, st, v->attn_version);
    v->toc(st);
    if (r != cudaSuccess) return v->fail_cuda("attention launch", r);
    v->launches++;
    {
      EpiResidBias::Params p{x, C, w.b_proj.as<float>()};
      if (vis_gemm<EpiResidBias>(v, v->attn.as<bf16>(), C, w.w_proj.as<bf16>(), C, M, C, C, p, st)) return 1;
    }
    v->tic(2, st);
    launch_layernorm<bf16>(xn, x, w.ln2_w.as<float>(), w.ln2_b.as<float>(), M, C, v->cfg.ln_eps, st);
    v->toc(st);
    VCL();
    {
      EpiStore<bf16, true, true>::Params p{v->act.as<bf16>(), F, w.b_fc1.as<float>()};
      if (vis_gemm<EpiStore<bf16, true, true>>(v, xn, C, w.w_fc1.as<bf16>(), C, M, F, C, p, st)) return 1;
    }
    {
      EpiResidBias::Params p{x, C, w.b_fc2.as<float>()};
      if (vis_gemm<EpiResidBias>(v, v->act.as<bf16>(), F, w.w_fc2.as<bf16>(), F, M, C, F, p, st)) return 1;
    }
  }
  v->tic(2, st);
  launch_layernorm<float>(feat_out, x, v->lnf_w.as<float>(), v->lnf_b.as<float>(), M, C, v->cfg.final_ln_eps, st);
  v->toc(st);
  VCL();
  return 0;
}

extern "C" int blim_vision_encode(blim_vision* v, const void* frames_dev, int dtype, int n_frames, float* feat_out_dev, void* stream) {
  if (!v) return 1;
  VCK(cudaSetDevice(v->device));
  return vis_encode(v, frames_dev, dtype, n_frames, feat_out_dev, S(stream));
}

// ------------------------------------------------------------------------------------------------ token merging
// merge_tokens (mm_projector_builder.py:101-130): rounds of r = p // 2 until the target is within reach.
static std::vector<int> tome_schedule(int p, int target) {
  std::vector<int> rs;
  int tmp = p;
  while (tmp != target) {
    if (tmp - target <= tmp / 2) { rs.push_back(tmp - target); break; }
    rs.push_back(tmp / 2);
    tmp -= tmp / 2;
  }
  return rs;
}

static int vis_merge(blim_vision* v, const float* x_in, int b, int p, int target, float* out, int32_t* edge_out, int32_t* nidx_out, cudaStream_t st) {
  if (!x_in || !out || b <= 0) return v->fail("bad merge_tokens arguments");
  if (p <= target || target <= 0) return v->fail("merge_tokens: p must be greater than the target (mm_projector_builder.py:110)");
  if (static_cast<size_t>(b) * p > static_cast<size_t>(v->max_clips) * v->TL) return v->fail("merge_tokens: input larger than the workspace");
  const int C = v->C, dim = C / v->NH;
  const std::vector<int> rs = tome_schedule(p, target);
  const float* cur = x_in;
  const float* cur_size = nullptr;
  int pc = p;
  v->tic(4, st);
  for (size_t round = 0; round < rs.size(); ++round) {
    int r = std::min(rs[round], pc / 2);   // bipartite_soft_matching clamps r to t // 2 (:19)
    const int na = (pc + 1) / 2;
    if (static_cast<size_t>(b) * na > static_cast<size_t>(v->max_clips) * ((v->TL + 1) / 2)) return v->fail("merge_tokens: round larger than the workspace");
    tome_metric_kernel<<<dim3((pc * 32 + 255) / 256, b), 256, 0, st>>>(v->t_metric.as<float>(), cur, pc, C, v->NH);
    VCL();
    dim3 mg((na + 15) / 16, b);
    if (dim == 64) tome_match_kernel<64><<<mg, 256, 0, st>>>(v->t_nmax.as<float>(), v->t_nidx.as<int>(), v->t_metric.as<float>(), pc);
    else tome_match_kernel<128><<<mg, 256, 0, st>>>(v->t_nmax.as<float>(), v->t_nidx.as<int>(), v->t_metric.as<float>(), pc);
    VCL();
    int n_sort = 2;
    while (n_sort < na) n_sort <<= 1;
    if (n_sort > 8192) return v->fail("merge_tokens: more than 16384 tokens per clip are not supported");
    static bool sort_attr = false;
    if (!sort_attr) {
      VCK(cudaFuncSetAttribute(tome_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 8));
      sort_attr = true;
    }
    tome_sort_kernel<<<b, 1024, static_cast<size_t>(n_sort) * 8, st>>>(v->t_edge.as<int>(), v->t_dst.as<int>(), v->t_cstart.as<int>(),
                                                                       v->t_cend.as<int>(), v->t_nmax.as<float>(), v->t_nidx.as<int>(), na, pc / 2, r,
                                                                       n_sort);
    VCL();
    if (round == 0) {
      if (edge_out) VCK(cudaMemcpyAsync(edge_out, v->t_edge.p, static_cast<size_t>(b) * na * 4, cudaMemcpyDeviceToDevice, st));
      if (nidx_out) VCK(cudaMemcpyAsync(nidx_out, v->t_nidx.p, static_cast<size_t>(b) * na * 4, cudaMemcpyDeviceToDevice, st));
    }
    const bool last = round + 1 == rs.size();
    float* dst = last ? out : v->t_x[round & 1].as<float>();
    float* dst_size = v->t_size[round & 1].as<float>();
    tome_merge_kernel<<<dim3(pc - r, b), 256, 0, st>>>(dst, dst_size, cur, cur_size, v->t_edge.as<int>(), v->t_dst.as<int>(),
                                                       v->t_cstart.as<int>(), v->t_cend.as<int>(), pc, r, C);
    VCL();
    cur = dst;
    cur_size = dst_size;
    pc -= r;
  }
  v->toc(st);
  return 0;
}

extern "C" int blim_vision_merge_tokens(blim_vision* v, const float* x_dev, int b, int p, int target, float* out_dev, int32_t* edge_idx_out_dev,
                                        int32_t* node_idx_out_dev, void* stream) {
  if (!v) return 1;
  VCK(cudaSetDevice(v->device));
  return vis_merge(v, x_dev, b, p, target, out_dev, edge_idx_out_dev, node_idx_out_dev, S(stream));
}

extern "C" int blim_vision_extract(blim_vision* v, const void* frames_dev, int dtype, int n_frames, void* out_dev, int out_dtype, void* stream) {
  if (!v) return 1;
  if (!out_dev || out_dtype < 0 || out_dtype > 2) return v->fail("bad extract arguments");
  VCK(cudaSetDevice(v->device));
  cudaStream_t st = S(stream);
  if (vis_encode(v, frames_dev, dtype, n_frames, v->feat.as<float>(), st)) return 1;
  const int n_clips = n_frames / v->FPC;
  const int target = v->cfg.tome_tokens_per_frame * v->FPC;
  // the merged fp32 features land in x (the residual stream is dead after the final LayerNorm)
  if (vis_merge(v, v->feat.as<float>(), n_clips, v->TL, target, v->x.as<float>(), nullptr, nullptr, st)) return 1;
  const size_t n = static_cast<size_t>(n_clips) * target * v->C;
  f32_to_dtype_kernel<<<static_cast<unsigned>(std::min<size_t>((n + 255) / 256, 2048)), 256, 0, st>>>(out_dev, v->x.as<float>(), out_dtype, n);
  VCL();
  return 0;
}

extern "C" int64_t blim_vision_kernel_launches(const blim_vision* v) { return v ? v->launches + v->gemm.launches : 0; }
extern "C" double blim_vision_gemm_flops(const blim_vision* v) { return v ? v->flops : 0.0; }
extern "C" int blim_vision_profile(blim_vision* v, int enable) {
  if (!v) return 1;
  v->profiling = enable != 0;
  return 0;
}
extern "C" int blim_vision_profile_read(blim_vision* v, int n, double* ms, int64_t* launches) {
  if (!v) return 1;
  if (n < 5 || !ms || !launches) return v->fail("blim_vision_profile_read: need room for 5 categories");
  VCK(cudaSetDevice(v->device));
  VCK(cudaDeviceSynchronize());
  for (int i = 0; i < n; ++i) { ms[i] = 0.0; launches[i] = 0; }
  for (auto& t : v->timed) {
    float f = 0.f;
    if (cudaEventElapsedTime(&f, t.a, t.b) == cudaSuccess) { ms[t.cat] += f; launches[t.cat]++; }
    v->pool.push_back(t.a);
    v->pool.push_back(t.b);
  }
  v->timed.clear();
  return 0;
}
