// Backward pieces of the MoE block that are not GEMMs (HBM-bound element-wise / gather kernels).
// Forward definitions: aria/model/moe_lm.py:505-507 (glu), :350-364 (unpermute + weighted sum), :261-262 (top-k softmax).
#include "common.cuh"
#include "ptx.cuh"

namespace aria {

__global__ void __launch_bounds__(256) swiglu_fwd_kernel(const uint4* __restrict__ h1, uint4* __restrict__ h, int64_t rows, int vI) {
  // one thread per 8 output columns
  const int64_t total = rows * vI;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / vI;
    const int c = static_cast<int>(i - r * vI);
    const uint4 g4 = h1[r * 2 * vI + c], u4 = h1[r * 2 * vI + vI + c];
    const uint32_t g[4] = {g4.x, g4.y, g4.z, g4.w}, u[4] = {u4.x, u4.y, u4.z, u4.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float s0 = bf16r(fast_silu(bf16_lo(g[j]))), s1 = bf16r(fast_silu(bf16_hi(g[j])));
      o[j] = pack_bf16(s0 * bf16_lo(u[j]), s1 * bf16_hi(u[j]));
    }
    h[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

__global__ void __launch_bounds__(256) swiglu_bwd_kernel(const uint4* __restrict__ h1, const uint4* __restrict__ dh,
                                                         uint4* __restrict__ dh1, int64_t rows, int vI) {
  const int64_t total = rows * vI;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / vI;
    const int c = static_cast<int>(i - r * vI);
    const uint4 g4 = h1[r * 2 * vI + c], u4 = h1[r * 2 * vI + vI + c], d4 = dh[i];
    const uint32_t g[4] = {g4.x, g4.y, g4.z, g4.w}, u[4] = {u4.x, u4.y, u4.z, u4.w}, dd[4] = {d4.x, d4.y, d4.z, d4.w};
    uint32_t og[4], ou[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float dgv[2], duv[2];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const float gv = hh ? bf16_hi(g[j]) : bf16_lo(g[j]);
        const float uv = hh ? bf16_hi(u[j]) : bf16_lo(u[j]);
        const float dv = hh ? bf16_hi(dd[j]) : bf16_lo(dd[j]);
        const float sig = fast_rcp(1.0f + fast_ex2(-1.4426950408889634f * gv));
        const float silu = gv * sig;
        duv[hh] = dv * silu;
        dgv[hh] = dv * uv * (sig * (1.0f + gv * (1.0f - sig)));
      }
      og[j] = pack_bf16(dgv[0], dgv[1]);
      ou[j] = pack_bf16(duv[0], duv[1]);
    }
    dh1[r * 2 * vI + c] = make_uint4(og[0], og[1], og[2], og[3]);
    dh1[r * 2 * vI + vI + c] = make_uint4(ou[0], ou[1], ou[2], ou[3]);
  }
}

// One block per token: dy rows (scaled copies of dout[t]) and dscores (block-reduced dot products).
__global__ void __launch_bounds__(256) combine_bwd_kernel(const uint4* __restrict__ dout, const uint4* __restrict__ y,
                                                          const int32_t* __restrict__ dest_row, const __nv_bfloat16* __restrict__ scores,
                                                          uint4* __restrict__ dy, float* __restrict__ dscores, int64_t T, int vpr, int k) {
  __shared__ float red[8][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t t = blockIdx.x; t < T; t += gridDim.x) {
    float dots[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) dots[j] = 0.f;
    for (int v = threadIdx.x; v < vpr; v += blockDim.x) {
      const uint4 g4 = __ldg(dout + t * vpr + v);
      const uint32_t g[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (j < k) {
          const int r = dest_row[t * k + j];
          const float s = __bfloat162float(scores[t * k + j]);
          const uint4 y4 = __ldg(y + static_cast<int64_t>(r) * vpr + v);
          const uint32_t yy[4] = {y4.x, y4.y, y4.z, y4.w};
          uint32_t o[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float g0 = bf16_lo(g[i]), g1 = bf16_hi(g[i]);
            dots[j] += g0 * bf16_lo(yy[i]) + g1 * bf16_hi(yy[i]);
            o[i] = pack_bf16(g0 * s, g1 * s);
          }
          dy[static_cast<int64_t>(r) * vpr + v] = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int o = 16; o; o >>= 1) dots[j] += __shfl_xor_sync(0xffffffffu, dots[j], o);
    }
    __syncthreads();
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) red[warp][j] = dots[j];
    }
    __syncthreads();
    if (threadIdx.x < k) {
      float s = 0.f;
      for (int w = 0; w < (blockDim.x >> 5); ++w) s += red[w][threadIdx.x];
      dscores[t * k + threadIdx.x] = s;
    }
  }
}

__global__ void __launch_bounds__(256) router_bwd_kernel(const float* __restrict__ dscores, const __nv_bfloat16* __restrict__ scores,
                                                         const int32_t* __restrict__ top_idx, __nv_bfloat16* __restrict__ dlogits,
                                                         int64_t T, int E, int k) {
  // one warp per token: lanes clear the E logits, then lanes < k write their selected entries
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * wpb + (threadIdx.x >> 5); t < T; t += static_cast<int64_t>(gridDim.x) * wpb) {
    for (int e = lane; e < E; e += 32) dlogits[t * E + e] = __float2bfloat16_rn(0.f);
    float s = 0.f, g = 0.f;
    if (lane < k) {
      s = __bfloat162float(scores[t * k + lane]);
      g = dscores[t * k + lane];
    }
    float dot = s * g;
#pragma unroll
    for (int o = 16; o; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    __syncwarp();
    if (lane < k) dlogits[t * E + top_idx[t * k + lane]] = __float2bfloat16_rn(s * (g - dot));
  }
}


// ------------------------------------------------------------------------------------------------
// Training-mode router losses (moe_lm.py:128-166, 203-241).  The reference never returns their values: they act on the
// router only through MoEAuxLossAutoScaler.backward (:103-117), which injects d(loss)/d(logits) * main_loss_backward_scale.
//   z   = c_z * mean_t(lse_t^2)                                  -> dlogits[t,e] += s * c_z * 2 lse_t / T * p[t,e]
//   aux = c_aux * E/(T k) * sum_e mean_t(p[t,e]) * count_e       -> dlogits[t,e] += s * p[t,e] * (g_e - sum_e' p[t,e'] g_e'),
//                                                                   g_e = c_aux * E * count_e / (T k T)
// with p = softmax(logits) in fp32 (:235) and lse = logsumexp(logits).  One warp per token, E <= 256.
constexpr int AUX_MAX_E = 256;

__device__ __forceinline__ void token_softmax(const __nv_bfloat16* __restrict__ row, int E, int lane, float (&p)[AUX_MAX_E / 32], float& lse) {
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < AUX_MAX_E / 32; ++i) {
    const int e = lane + i * 32;
    p[i] = e < E ? __bfloat162float(row[e]) : -INFINITY;
    mx = fmaxf(mx, p[i]);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < AUX_MAX_E / 32; ++i) {
    p[i] = (lane + i * 32 < E) ? __expf(p[i] - mx) : 0.f;
    sum += p[i];
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.0f / sum;
#pragma unroll
  for (int i = 0; i < AUX_MAX_E / 32; ++i) p[i] *= inv;
  lse = mx + __logf(sum);
}

// dlogits (bf16, already holding the top-k softmax term from router_bwd_kernel) += loss gradients
__global__ void __launch_bounds__(256) router_aux_bwd_kernel(const __nv_bfloat16* __restrict__ logits, const int32_t* __restrict__ counts,
                                                             __nv_bfloat16* __restrict__ dlogits, int64_t T, int E, int k, float z_coeff,
                                                             float aux_coeff, float loss_scale) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  const float invT = 1.0f / static_cast<float>(T);
  float g[AUX_MAX_E / 32];
#pragma unroll
  for (int i = 0; i < AUX_MAX_E / 32; ++i) {
    const int e = lane + i * 32;
    g[i] = e < E ? aux_coeff * static_cast<float>(E) * static_cast<float>(counts[e]) * invT * invT / static_cast<float>(k) : 0.f;
  }
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * wpb + (threadIdx.x >> 5); t < T; t += static_cast<int64_t>(gridDim.x) * wpb) {
    float p[AUX_MAX_E / 32], lse;
    token_softmax(logits + t * E, E, lane, p, lse);
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < AUX_MAX_E / 32; ++i) dot += p[i] * g[i];
#pragma unroll
    for (int o = 16; o; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    const float zc = z_coeff * 2.0f * lse * invT;
#pragma unroll
    for (int i = 0; i < AUX_MAX_E / 32; ++i) {
      const int e = lane + i * 32;
      if (e < E) {
        const float add = loss_scale * p[i] * (zc + g[i] - dot);
        dlogits[t * E + e] = __float2bfloat16_rn(__bfloat162float(dlogits[t * E + e]) + add);
      }
    }
  }
}

// loss values for logging: stage 1 = per-block partial sums [grid][1 + E] (sum lse^2 | sum_t p[t,e]); stage 2 = fixed-order reduce
__global__ void __launch_bounds__(256) router_aux_partial_kernel(const __nv_bfloat16* __restrict__ logits, float* __restrict__ partial,
                                                                 int64_t T, int E) {
  __shared__ float acc[8][AUX_MAX_E + 1];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  float ps[AUX_MAX_E / 32], z = 0.f;
#pragma unroll
  for (int i = 0; i < AUX_MAX_E / 32; ++i) ps[i] = 0.f;
  for (int64_t t = static_cast<int64_t>(blockIdx.x) * wpb + w; t < T; t += static_cast<int64_t>(gridDim.x) * wpb) {
    float p[AUX_MAX_E / 32], lse;
    token_softmax(logits + t * E, E, lane, p, lse);
    z += lse * lse;
#pragma unroll
    for (int i = 0; i < AUX_MAX_E / 32; ++i) ps[i] += p[i];
  }
#pragma unroll
  for (int i = 0; i < AUX_MAX_E / 32; ++i)
    if (lane + i * 32 < E) acc[w][1 + lane + i * 32] = ps[i];
  if (lane == 0) acc[w][0] = z;
  __syncthreads();
  for (int c = threadIdx.x; c <= E; c += blockDim.x) {
    float s = 0.f;
    for (int ww = 0; ww < wpb; ++ww) s += acc[ww][c];
    partial[static_cast<int64_t>(blockIdx.x) * (E + 1) + c] = s;
  }
}

__global__ void __launch_bounds__(256) router_aux_final_kernel(const float* __restrict__ partial, const int32_t* __restrict__ counts,
                                                               float* __restrict__ losses, int nblk, int64_t T, int E, int k, float z_coeff,
                                                               float aux_coeff) {
  __shared__ float red[AUX_MAX_E + 1];
  for (int c = threadIdx.x; c <= E; c += blockDim.x) {
    float s = 0.f;
    for (int b = 0; b < nblk; ++b) s += partial[static_cast<int64_t>(b) * (E + 1) + c];
    red[c] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const float invT = 1.0f / static_cast<float>(T);
    float aux = 0.f;
    for (int e = 0; e < E; ++e) aux += red[1 + e] * invT * static_cast<float>(counts[e]);
    losses[0] = z_coeff * red[0] * invT;
    losses[1] = aux * (static_cast<float>(E) * invT / static_cast<float>(k)) * aux_coeff;
  }
}

static inline int ew_grid(int64_t n, int per_block) {
  int64_t b = (n + per_block - 1) / per_block;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

}  // namespace aria

using namespace aria;

extern "C" int aria_swiglu_fwd(const void* h1, void* h, int64_t rows, int32_t I, aria_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ARIA_CHECK_ARG(h1 && h && rows >= 0 && I > 0 && I % 8 == 0);
  if (rows == 0) return ARIA_OK;
  swiglu_fwd_kernel<<<ew_grid(rows * (I / 8), 256), 256, 0, stream>>>(static_cast<const uint4*>(h1), static_cast<uint4*>(h), rows, I / 8);
  return check_launch("swiglu_fwd_kernel");
}

extern "C" int aria_swiglu_bwd(const void* h1, const void* dh, void* dh1, int64_t rows, int32_t I, aria_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ARIA_CHECK_ARG(h1 && dh && dh1 && rows >= 0 && I > 0 && I % 8 == 0);
  if (rows == 0) return ARIA_OK;
  swiglu_bwd_kernel<<<ew_grid(rows * (I / 8), 256), 256, 0, stream>>>(static_cast<const uint4*>(h1), static_cast<const uint4*>(dh),
                                                                     static_cast<uint4*>(dh1), rows, I / 8);
  return check_launch("swiglu_bwd_kernel");
}

extern "C" int aria_combine_bwd(const void* dout, const void* y, const int32_t* dest_row, const void* scores, void* dy, float* dscores,
                                int64_t T, int32_t d, int32_t k, aria_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ARIA_CHECK_ARG(dout && y && dest_row && scores && dy && dscores && d % 8 == 0 && k >= 1 && k <= 8 && T >= 0);
  if (T == 0) return ARIA_OK;
  int64_t grid = T;
  if (grid > static_cast<int64_t>(sm_count()) * 16) grid = static_cast<int64_t>(sm_count()) * 16;
  combine_bwd_kernel<<<static_cast<int>(grid), 256, 0, stream>>>(static_cast<const uint4*>(dout), static_cast<const uint4*>(y), dest_row,
                                                                static_cast<const __nv_bfloat16*>(scores), static_cast<uint4*>(dy),
                                                                dscores, T, d / 8, k);
  return check_launch("combine_bwd_kernel");
}

extern "C" int aria_router_bwd(const float* dscores, const void* scores, const int32_t* top_idx, void* dlogits, int64_t T, int32_t E,
                               int32_t k, aria_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ARIA_CHECK_ARG(dscores && scores && top_idx && dlogits && T >= 0 && E >= 1 && k >= 1 && k <= 8);
  if (T == 0) return ARIA_OK;
  router_bwd_kernel<<<ew_grid(T, 8), 256, 0, stream>>>(dscores, static_cast<const __nv_bfloat16*>(scores), top_idx,
                                                      static_cast<__nv_bfloat16*>(dlogits), T, E, k);
  return check_launch("router_bwd_kernel");
}

extern "C" size_t aria_router_aux_workspace_bytes(int32_t E) {
  return static_cast<size_t>(sm_count()) * 4 * static_cast<size_t>(E + 1) * sizeof(float);
}

extern "C" int aria_router_aux_loss(const void* logits, const int32_t* counts, float* losses, int64_t T, int32_t E, int32_t k,
                                    float z_coeff, float aux_coeff, void* workspace, size_t ws_bytes, aria_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ARIA_CHECK_ARG(logits && counts && losses && workspace && T >= 1 && E >= 1 && E <= AUX_MAX_E && k >= 1);
  int grid = ew_grid(T, 8);
  if (grid > sm_count() * 4) grid = sm_count() * 4;
  ARIA_CHECK_ARG(ws_bytes >= static_cast<size_t>(grid) * (E + 1) * sizeof(float));
  router_aux_partial_kernel<<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(logits), static_cast<float*>(workspace), T, E);
  int rc = check_launch("router_aux_partial_kernel");
  if (rc) return rc;
  router_aux_final_kernel<<<1, 256, 0, stream>>>(static_cast<const float*>(workspace), counts, losses, grid, T, E, k, z_coeff, aux_coeff);
  return check_launch("router_aux_final_kernel");
}

extern "C" int aria_router_aux_bwd(const void* logits, const int32_t* counts, void* dlogits, int64_t T, int32_t E, int32_t k,
                                   float z_coeff, float aux_coeff, float loss_scale, aria_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ARIA_CHECK_ARG(logits && counts && dlogits && T >= 0 && E >= 1 && E <= AUX_MAX_E && k >= 1);
  if (T == 0) return ARIA_OK;
  router_aux_bwd_kernel<<<ew_grid(T, 8), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(logits), counts,
                                                          static_cast<__nv_bfloat16*>(dlogits), T, E, k, z_coeff, aux_coeff, loss_scale);
  return check_launch("router_aux_bwd_kernel");
}
