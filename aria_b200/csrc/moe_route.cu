// MoE routing / dispatch kernels (HBM-bound integer + row-copy work; no tensor cores by design).
//   route_from_logits  : torch.topk + softmax(fp32)->bf16 + histc      (aria/model/moe_lm.py:261-269)
//   build_permutation  : argsort(stable) of the flattened expert ids  (moe_lm.py:329) as a counting sort
//   permute_rows       : index_select of token rows                    (moe_lm.py:330)
//   unpermute_combine  : zeros/index_copy_/mul/sum (+ shared add)      (moe_lm.py:350-364, :576)
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace aria {

constexpr int MAX_E = 64;
constexpr int MAX_K = 8;

// One warp per token. Each lane owns experts {lane, lane+32}.  Selection = k rounds of warp arg-max with the
// tie rule "lowest expert id" (oracle/aria_oracle.py header).
__global__ void __launch_bounds__(256) route_kernel(const __nv_bfloat16* __restrict__ logits, int32_t* __restrict__ top_idx,
                                                    __nv_bfloat16* __restrict__ scores, int32_t* __restrict__ counts,
                                                    int64_t T, int E, int k) {
  __shared__ int hist[MAX_E];
  for (int i = threadIdx.x; i < MAX_E; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int64_t warp_global = static_cast<int64_t>(blockIdx.x) * warps_per_block + (threadIdx.x >> 5);
  const int64_t warp_stride = static_cast<int64_t>(gridDim.x) * warps_per_block;
  for (int64_t t = warp_global; t < T; t += warp_stride) {
    const __nv_bfloat16* row = logits + t * E;
    float v0 = lane < E ? __bfloat162float(row[lane]) : -INFINITY;
    float v1 = lane + 32 < E ? __bfloat162float(row[lane + 32]) : -INFINITY;
    bool used0 = lane >= E, used1 = lane + 32 >= E;
    float sel_v = 0.f;  // lane j holds the j-th selected logit
    int sel_i = 0;
    for (int j = 0; j < k; ++j) {
      float bv;
      int bi;
      // local best (lower index wins ties: candidate 0 has the lower index)
      if (!used0 && (used1 || v0 >= v1)) {
        bv = v0;
        bi = lane;
      } else if (!used1) {
        bv = v1;
        bi = lane + 32;
      } else {
        bv = -INFINITY;
        bi = 0x7fffffff;
      }
#pragma unroll
      for (int off = 16; off; off >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, bv, off);
        int oi = __shfl_xor_sync(0xffffffffu, bi, off);
        if (ov > bv || (ov == bv && oi < bi)) {
          bv = ov;
          bi = oi;
        }
      }
      if (bi == lane) used0 = true;
      if (bi == lane + 32) used1 = true;
      if (lane == j) {
        sel_v = bv;
        sel_i = bi;
      }
    }
    // softmax over the k selected logits in fp32 (moe_lm.py:262), max = first selected
    const float vmax = __shfl_sync(0xffffffffu, sel_v, 0);
    float e = lane < k ? expf(sel_v - vmax) : 0.f;
    float s = 0.f;
    for (int j = 0; j < k; ++j) s += __shfl_sync(0xffffffffu, e, j);
    if (lane < k) {
      top_idx[t * k + lane] = sel_i;
      scores[t * k + lane] = __float2bfloat16_rn(e / s);
      atomicAdd(&hist[sel_i], 1);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < E; i += blockDim.x)
    if (hist[i]) atomicAdd(&counts[i], hist[i]);
}

// Routing with the expert choice GIVEN (parity / replay hook, TopKRouter.forced_top_indices): scores = fp32 softmax over the
// k logits at the given ids in the given order (moe_lm.py:262), counts = histogram (:264-269).  One warp per token.
__global__ void __launch_bounds__(256) route_given_kernel(const __nv_bfloat16* __restrict__ logits, const int32_t* __restrict__ top_idx,
                                                          __nv_bfloat16* __restrict__ scores, int32_t* __restrict__ counts,
                                                          int64_t T, int E, int k) {
  __shared__ int hist[MAX_E];
  for (int i = threadIdx.x; i < MAX_E; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int64_t warp_global = static_cast<int64_t>(blockIdx.x) * warps_per_block + (threadIdx.x >> 5);
  const int64_t warp_stride = static_cast<int64_t>(gridDim.x) * warps_per_block;
  for (int64_t t = warp_global; t < T; t += warp_stride) {
    const int id = lane < k ? top_idx[t * k + lane] : 0;
    const float v = lane < k ? __bfloat162float(logits[t * E + id]) : -INFINITY;
    float vmax = v;
#pragma unroll
    for (int off = 16; off; off >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, off));
    const float e = lane < k ? expf(v - vmax) : 0.f;
    float s = 0.f;
    for (int j = 0; j < k; ++j) s += __shfl_sync(0xffffffffu, e, j);
    if (lane < k) {
      scores[t * k + lane] = __float2bfloat16_rn(e / s);
      atomicAdd(&hist[id], 1);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < E; i += blockDim.x)
    if (hist[i]) atomicAdd(&counts[i], hist[i]);
}

// Stable counting sort, one block per expert: dest_row[i] = offsets[e] + #{i' < i : id[i'] == e}.
__global__ void __launch_bounds__(1024) permutation_kernel(const int32_t* __restrict__ top_idx,
                                                           const int32_t* __restrict__ counts,
                                                           int32_t* __restrict__ offsets, int32_t* __restrict__ dest_row,
                                                           int32_t* __restrict__ src_token, int64_t n, int E, int k, int align) {
  const int e = blockIdx.x;
  __shared__ int warp_cnt[32];
  __shared__ int base_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    // each expert's block starts on a multiple of `align` rows (1 for inference; 16 in training so that the
    // wgrad GEMM's 16-row K steps never straddle two experts); pad rows keep src_token = -1
    int acc = 0;
    for (int i = 0; i < e; ++i) acc += (counts[i] + align - 1) / align * align;
    base_s = acc;
    if (e == 0) {
      int a = 0;
      for (int i = 0; i < E; ++i) {
        offsets[i] = a;
        a += (counts[i] + align - 1) / align * align;
      }
      offsets[E] = a;
    }
  }
  __syncthreads();
  int running = base_s;
  for (int64_t start = 0; start < n; start += blockDim.x) {
    const int64_t i = start + threadIdx.x;
    const bool hit = i < n && top_idx[i] == e;
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) warp_cnt[warp] = __popc(m);
    __syncthreads();
    int before = 0, total = 0;
    // 32 warps: every thread sums the (tiny) array
#pragma unroll
    for (int w = 0; w < 32; ++w) {
      const int c = warp_cnt[w];
      if (w < warp) before += c;
      total += c;
    }
    if (hit) {
      const int r = running + before + __popc(m & ((1u << lane) - 1));
      dest_row[i] = r;
      src_token[r] = static_cast<int32_t>(i / k);
    }
    running += total;
    __syncthreads();
  }
}

// Same stable counting sort for up to ~32 K ids in ONE block (the prefill / decode sizes: 768 tokens x 6 = 4608 ids): the id
// array is cut into 32 contiguous chunks, one per warp; pass 1 = per-warp histograms (match_any, no atomics), pass 2 = expert
// offsets and, per (warp, expert), the first destination row; pass 3 = each warp re-walks its chunk in order and ranks the
// ids of a 32-id group among themselves with match_any.  ~20 x less work than E blocks each scanning all ids.
__global__ void __launch_bounds__(1024) permutation_small_kernel(const int32_t* __restrict__ top_idx, int32_t* __restrict__ offsets,
                                                                 int32_t* __restrict__ dest_row, int32_t* __restrict__ src_token,
                                                                 int n, int E, int k, int align) {
  __shared__ int hist[32][MAX_E];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int chunk = ((n + 31) / 32 + 31) / 32 * 32;  // ids per warp, a multiple of 32
  const int lo = warp * chunk, hi = min(n, lo + chunk);
  for (int e = lane; e < E; e += 32) hist[warp][e] = 0;
  __syncwarp();
  for (int i0 = lo; i0 < hi; i0 += 32) {
    const int i = i0 + lane;
    const int e = i < hi ? top_idx[i] : (0x40000000 + lane);  // inactive lanes: unique keys
    const unsigned m = __match_any_sync(0xffffffffu, e);
    if (i < hi && (m & ((1u << lane) - 1)) == 0) hist[warp][e] += __popc(m);  // first lane of each key
    __syncwarp();
  }
  __syncthreads();
  __shared__ int tot_s[MAX_E], off_s[MAX_E + 1];
  if (threadIdx.x < E) {  // per-expert totals, one thread per expert (a single thread summing 32 x E entries took ~12 us)
    int tot = 0;
#pragma unroll
    for (int w = 0; w < 32; ++w) tot += hist[w][threadIdx.x];
    tot_s[threadIdx.x] = (tot + align - 1) / align * align;   // blocks start on multiples of `align` rows
  }
  __syncthreads();
  if (threadIdx.x == 0) {  // exclusive scan of E <= 64 totals
    int a = 0;
    for (int e = 0; e < E; ++e) {
      off_s[e] = a;
      a += tot_s[e];
    }
    off_s[E] = a;
  }
  __syncthreads();
  if (threadIdx.x <= E) offsets[threadIdx.x] = off_s[threadIdx.x];
  if (threadIdx.x < E) {  // hist[w][e] <- first destination row of warp w's ids of expert e
    const int e = threadIdx.x;
    int a = off_s[e];
    for (int w = 0; w < 32; ++w) {
      const int c = hist[w][e];
      hist[w][e] = a;
      a += c;
    }
  }
  __syncthreads();
  for (int i0 = lo; i0 < hi; i0 += 32) {
    const int i = i0 + lane;
    const int e = i < hi ? top_idx[i] : (0x40000000 + lane);
    const unsigned m = __match_any_sync(0xffffffffu, e);
    if (i < hi) {
      const int r = hist[warp][e] + __popc(m & ((1u << lane) - 1));
      dest_row[i] = r;
      src_token[r] = i / k;
    }
    __syncwarp();
    if (i < hi && (m & ((1u << lane) - 1)) == 0) hist[warp][e] += __popc(m);
    __syncwarp();
  }
}

// permuted[r] = x[src_token[r]]; one warp per row, 128-bit loads/stores.
__global__ void __launch_bounds__(256) permute_rows_kernel(const uint4* __restrict__ x, const int32_t* __restrict__ src_token,
                                                           uint4* __restrict__ out, int64_t rows, int vec_per_row) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int64_t r = static_cast<int64_t>(blockIdx.x) * wpb + (threadIdx.x >> 5); r < rows;
       r += static_cast<int64_t>(gridDim.x) * wpb) {
    const int st = src_token[r];
    uint4* dst = out + r * vec_per_row;
    if (st < 0) {  // alignment pad row (training-mode layout): zeros
      for (int v = lane; v < vec_per_row; v += 32) dst[v] = make_uint4(0, 0, 0, 0);
      continue;
    }
    const uint4* src = x + static_cast<int64_t>(st) * vec_per_row;
    for (int v = lane; v < vec_per_row; v += 32) dst[v] = __ldg(src + v);
  }
}

// out[t] = bf16( sum_j bf16(y[dest[t*k+j]] * s[t,j]) ) (+ shared[t] with one more bf16 rounding).
// One block per token, one 16-byte column per thread: the k gathered rows are k independent loads in flight.
__global__ void __launch_bounds__(512) combine_kernel(const uint4* __restrict__ y, const int32_t* __restrict__ dest_row,
                                                      const __nv_bfloat16* __restrict__ scores, const uint4* __restrict__ shared_out,
                                                      uint4* __restrict__ out, int64_t T, int vec_per_row, int k) {
  for (int64_t t = blockIdx.x; t < T; t += gridDim.x) {
    int rows[MAX_K];
    float sc[MAX_K];
#pragma unroll
    for (int j = 0; j < MAX_K; ++j) {
      if (j < k) {
        rows[j] = dest_row[t * k + j];
        sc[j] = __bfloat162float(scores[t * k + j]);
      }
    }
    for (int v = threadIdx.x; v < vec_per_row; v += blockDim.x) {
      uint4 q[MAX_K];
#pragma unroll
      for (int j = 0; j < MAX_K; ++j)
        if (j < k) q[j] = __ldg(y + static_cast<int64_t>(rows[j]) * vec_per_row + v);
      uint4 sh = make_uint4(0, 0, 0, 0);
      if (shared_out) sh = __ldg(shared_out + t * vec_per_row + v);
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
      for (int j = 0; j < MAX_K; ++j) {
        if (j < k) {
          const uint32_t w[4] = {q[j].x, q[j].y, q[j].z, q[j].w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            acc[2 * i] += bf16r(bf16_lo(w[i]) * sc[j]);
            acc[2 * i + 1] += bf16r(bf16_hi(w[i]) * sc[j]);
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = bf16r(acc[i]);
      if (shared_out) {
        const uint32_t w[4] = {sh.x, sh.y, sh.z, sh.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          acc[2 * i] += bf16_lo(w[i]);
          acc[2 * i + 1] += bf16_hi(w[i]);
        }
      }
      out[t * vec_per_row + v] =
          make_uint4(pack_bf16(acc[0], acc[1]), pack_bf16(acc[2], acc[3]), pack_bf16(acc[4], acc[5]), pack_bf16(acc[6], acc[7]));
    }
  }
}

__global__ void offsets_from_counts_kernel(const int64_t* counts, int32_t* offsets, int G) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int a = 0;
    for (int i = 0; i < G; ++i) {
      offsets[i] = a;
      a += static_cast<int>(counts[i]);
    }
    offsets[G] = a;
  }
}

static inline int grid_for_warps(int64_t n_warps, int wpb) {
  int64_t blocks = (n_warps + wpb - 1) / wpb;
  const int64_t cap = static_cast<int64_t>(sm_count()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace aria

using namespace aria;

extern "C" int aria_route_from_logits(const void* logits, int32_t* top_idx, void* scores, int32_t* counts, int64_t T,
                                      int32_t E, int32_t k, aria_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ARIA_CHECK_ARG(logits && top_idx && scores && counts);
  ARIA_CHECK_ARG(E >= 1 && E <= MAX_E && k >= 1 && k <= MAX_K && k <= E && T >= 0);
  if (cudaMemsetAsync(counts, 0, sizeof(int32_t) * E, stream) != cudaSuccess) return ARIA_ERR_CUDA;
  if (T == 0) return ARIA_OK;
  route_kernel<<<grid_for_warps(T, 8), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(logits), top_idx,
                                                         static_cast<__nv_bfloat16*>(scores), counts, T, E, k);
  return check_launch("route_kernel");
}

extern "C" int aria_route_given_indices(const void* logits, const int32_t* top_idx, void* scores, int32_t* counts, int64_t T,
                                        int32_t E, int32_t k, aria_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ARIA_CHECK_ARG(logits && top_idx && scores && counts);
  ARIA_CHECK_ARG(E >= 1 && E <= MAX_E && k >= 1 && k <= MAX_K && k <= E && T >= 0);
  if (cudaMemsetAsync(counts, 0, sizeof(int32_t) * E, stream) != cudaSuccess) return ARIA_ERR_CUDA;
  if (T == 0) return ARIA_OK;
  route_given_kernel<<<grid_for_warps(T, 8), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(logits), top_idx,
                                                               static_cast<__nv_bfloat16*>(scores), counts, T, E, k);
  return check_launch("route_given_kernel");
}

extern "C" int aria_router_topk(const void* x, const void* w_router, void* logits_out, int32_t* top_idx, void* scores,
                                int32_t* counts, int64_t T, int32_t d, int32_t E, int32_t k, aria_stream_t stream) {
  ARIA_CHECK_ARG(x && w_router && logits_out);
  ARIA_CHECK_ARG(E % 8 == 0);
  if (T == 0) return aria_route_from_logits(logits_out, top_idx, scores, counts, 0, E, k, stream);
  // gating (moe_lm.py:200): logits = F.linear(x, W) -> bf16, on the tensor cores
  aria_gemm_desc_t g{};
  g.a = x;
  g.lda = d;
  g.m = T;
  g.n = E;
  g.k = d;
  g.b[0] = w_router;
  g.n_seg = 1;
  g.b_layout = ARIA_B_NK;
  g.num_groups = 1;
  g.epilogue = ARIA_EPI_LINEAR;
  g.out[0] = logits_out;
  g.ldo = E;
  int rc = aria_gemm(&g, stream);
  if (rc) return rc;
  return aria_route_from_logits(logits_out, top_idx, scores, counts, T, E, k, stream);
}

extern "C" int aria_build_permutation(const int32_t* top_idx, const int32_t* counts, int32_t* offsets, int32_t* dest_row,
                                      int32_t* src_token, int64_t T, int32_t E, int32_t k, int32_t row_align,
                                      aria_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ARIA_CHECK_ARG(top_idx && counts && offsets && dest_row && src_token);
  ARIA_CHECK_ARG(E >= 1 && k >= 1 && T >= 0 && T * k < (1ll << 31) && row_align >= 1);
  if (row_align > 1) {  // src_token has T*k + E*(row_align-1) slots; pad rows stay -1
    const size_t slots = static_cast<size_t>(T) * k + static_cast<size_t>(E) * (row_align - 1);
    if (cudaMemsetAsync(src_token, 0xFF, slots * sizeof(int32_t), stream) != cudaSuccess) return ARIA_ERR_CUDA;
  }
  static const bool small_on = [] { const char* e = getenv("ARIA_PERM_SMALL"); return !(e && e[0] == '0'); }();
  if (small_on && E <= MAX_E && T * k <= 32768) {
    permutation_small_kernel<<<1, 1024, 0, stream>>>(top_idx, offsets, dest_row, src_token, static_cast<int>(T * k), E, k, row_align);
    return check_launch("permutation_small_kernel");
  }
  permutation_kernel<<<E, 1024, 0, stream>>>(top_idx, counts, offsets, dest_row, src_token, T * k, E, k, row_align);
  return check_launch("permutation_kernel");
}

extern "C" int aria_permute_rows(const void* x, const int32_t* src_token, void* permuted, int64_t rows, int32_t d,
                                 aria_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ARIA_CHECK_ARG(x && src_token && permuted && d % 8 == 0 && rows >= 0);
  if (rows == 0) return ARIA_OK;
  permute_rows_kernel<<<grid_for_warps(rows, 8), 256, 0, stream>>>(static_cast<const uint4*>(x), src_token,
                                                                   static_cast<uint4*>(permuted), rows, d / 8);
  return check_launch("permute_rows_kernel");
}

extern "C" int aria_unpermute_combine(const void* y, const int32_t* dest_row, const void* scores, const void* shared,
                                      void* out, int64_t T, int32_t d, int32_t k, aria_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ARIA_CHECK_ARG(y && dest_row && scores && out && d % 8 == 0 && k >= 1 && k <= MAX_K && T >= 0);
  if (T == 0) return ARIA_OK;
  int64_t cgrid = T;
  if (cgrid > static_cast<int64_t>(sm_count()) * 16) cgrid = static_cast<int64_t>(sm_count()) * 16;
  int cthreads = (d / 8 + 31) / 32 * 32;  // one 16-byte column per thread, whole row in one pass (d=2560 -> 320 threads)
  if (cthreads > 512) cthreads = 512;
  combine_kernel<<<static_cast<int>(cgrid), cthreads, 0, stream>>>(static_cast<const uint4*>(y), dest_row,
                                                           static_cast<const __nv_bfloat16*>(scores),
                                                           static_cast<const uint4*>(shared), static_cast<uint4*>(out), T,
                                                           d / 8, k);
  return check_launch("combine_kernel");
}

extern "C" int aria_offsets_from_counts(const int64_t* counts, int32_t* offsets, int32_t num_groups, aria_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ARIA_CHECK_ARG(counts && offsets && num_groups >= 1);
  offsets_from_counts_kernel<<<1, 32, 0, stream>>>(counts, offsets, num_groups);
  return check_launch("offsets_from_counts_kernel");
}
