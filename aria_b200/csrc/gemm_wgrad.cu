// Weight-gradient ("wgrad") grouped GEMM for sm_100a: the contraction runs over the (ragged) token rows.
//
//     dW[g][m][n] = sum_{r in group g} A[r][m] * B[r][n]        A = [rows, Md] (layer input), B = [rows, Nd] (grad of output)
//
// Backward of the reference's expert GEMMs (autograd through `gmm`, aria/model/moe_lm.py:484; weight [E, in, out]) and,
// with one group, of the nn.Linear layers (dW[out,in] = dY^T X).  Both operands are consumed MN-major straight from the
// row-major activation tensors (UMMA a_major = b_major = MN): no transposes are materialised.
//
// Ragged groups: group row offsets must be multiples of 16 (the training-mode dispatcher pads each expert's block with
// zero rows).  TMA always fetches 64-row slabs; for the last slab of a group only the 16-row UMMA K-steps that lie inside
// the group are issued, so rows of the next expert that share the slab are never multiplied.
#include <stdlib.h>

#include "gemm_common.cuh"

namespace aria {

struct WgradParams {
  int Md, Nd, G;
  int n_src;            // row groups are (source, g) pairs, source-major: offs has n_src*G+1 entries and out[g] sums over sources
  const int32_t* offs;  // [G+1], multiples of 16 (last = total rows, any)
  __nv_bfloat16* out;   // [G, Md, Nd]
};

constexpr int WG_BN = 128;
constexpr int WG_STAGES = 6;
constexpr int WG_STAGE_BYTES = 64 * BM * 2 + 64 * WG_BN * 2;  // A slab [64 rows][128 m] + B slab [64 rows][128 n]

__global__ void __launch_bounds__(GEMM_THREADS, 1)
wgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + WG_STAGES;
  uint64_t* tfull_bar = empty_bar + WG_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int i = 0; i < WG_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 2 * WG_BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int mt = (p.Md + BM - 1) / BM, nt = (p.Nd + WG_BN - 1) / WG_BN;
  const int tiles_per_group = mt * nt;
  const int total = p.G * tiles_per_group;
  // tile t -> (group, m tile, n tile); n innermost so concurrent CTAs share the A slab of the group
  auto decode = [&](int t, int& g, int& mi, int& ni) {
    g = t / tiles_per_group;
    const int r = t - g * tiles_per_group;
    mi = r / nt;
    ni = r - mi * nt;
  };
  // rows of (source s, group g): start row and number of 16-row UMMA K steps
  auto span = [&](int s, int g, int& r0, int& ksteps) {
    r0 = p.offs[s * p.G + g];
    ksteps = (p.offs[s * p.G + g + 1] - r0 + 15) / 16;
  };

  if (warp == 0) {
    if (elect_one()) {  // elect.sync: ptxas keeps the single-thread body on the uniform datapath
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        int g, mi, ni;
        decode(t, g, mi, ni);
        for (int src = 0; src < p.n_src; ++src) {
          int r0, ksteps;
          span(src, g, r0, ksteps);
          const int slabs = (ksteps + 3) / 4;
          for (int s = 0; s < slabs; ++s) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * WG_STAGE_BYTES;
            uint8_t* sb = sa + 64 * BM * 2;
            mbar_arrive_expect_tx(&full_bar[stage], WG_STAGE_BYTES);
            const int row = r0 + s * 64;
#pragma unroll
            for (int c = 0; c < BM / 64; ++c) tma_load_2d(sa + c * 8192, &tmA, &full_bar[stage], mi * BM + c * 64, row);
#pragma unroll
            for (int c = 0; c < WG_BN / 64; ++c) tma_load_2d(sb + c * 8192, &tmB, &full_bar[stage], ni * WG_BN + c * 64, row);
            if (++stage == WG_STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {  // elect.sync: ptxas keeps the single-thread body on the uniform datapath
      constexpr uint32_t idesc = make_idesc_bf16(BM, WG_BN, true, true);  // both operands MN-major
      const uint64_t dA0 = make_smem_desc(smem_u32(smem), 8192, 1024);    // stage 0, k-step 0; later ones are + (bytes >> 4)
      const uint64_t dB0 = make_smem_desc(smem_u32(smem) + 64 * BM * 2, 8192, 1024);
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
        int g, mi, ni;
        decode(t, g, mi, ni);
        const int as = it & 1;
        mbar_wait(&tempty_bar[as], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * WG_BN;
        uint32_t started = 0;
        for (int src = 0; src < p.n_src; ++src) {
          int r0, ksteps;
          span(src, g, r0, ksteps);
          const int slabs = (ksteps + 3) / 4;
          for (int s = 0; s < slabs; ++s) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint64_t da = dA0 + stage * (WG_STAGE_BYTES >> 4), db = dB0 + stage * (WG_STAGE_BYTES >> 4);
            const int kmax = min(4, ksteps - s * 4);
            for (int k = 0; k < kmax; ++k) {
              umma_bf16_ss(d_tmem, da + k * (2048 >> 4), db + k * (2048 >> 4), idesc, started);
              started = 1;
            }
            umma_commit(&empty_bar[stage]);
            if (++stage == WG_STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
        umma_commit(&tfull_bar[as]);  // with zero slabs this still arrives (nothing pending)
      }
    }
  } else {
    const int quad = warp & 3, half = (warp - 2) >> 2;
    int it = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++it) {
      int g, mi, ni;
      decode(t, g, mi, ni);
      int ksteps = 0;
      for (int src = 0; src < p.n_src; ++src) {
        int r0, ks;
        span(src, g, r0, ks);
        ksteps += ks;
      }
      const int as = it & 1;
      mbar_wait(&tfull_bar[as], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + as * WG_BN + (static_cast<uint32_t>(quad * 32) << 16);
      const int m = mi * BM + quad * 32 + lane;
      __nv_bfloat16* orow = p.out + (static_cast<int64_t>(g) * p.Md + m) * p.Nd + ni * WG_BN;
#pragma unroll 1
      for (int c = half * (WG_BN / 2); c < (half + 1) * (WG_BN / 2); c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + c, v);
        tmem_ld_wait();
        if (ksteps == 0) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0;  // empty group: the accumulator was never written
        }
        if (m < p.Md) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (ni * WG_BN + c + q * 8 + 8 <= p.Nd)
              *reinterpret_cast<uint4*>(orow + c + q * 8) =
                  make_uint4(pack_bf16(__uint_as_float(v[q * 8]), __uint_as_float(v[q * 8 + 1])),
                             pack_bf16(__uint_as_float(v[q * 8 + 2]), __uint_as_float(v[q * 8 + 3])),
                             pack_bf16(__uint_as_float(v[q * 8 + 4]), __uint_as_float(v[q * 8 + 5])),
                             pack_bf16(__uint_as_float(v[q * 8 + 6]), __uint_as_float(v[q * 8 + 7])));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 2 * WG_BN);
}

// ------------------------------------------------------------------------------------------------------------------
// 2-CTA variant: a CTA pair owns a 256 x 256 tile of dW[g]; CTA r stages feature columns [r*128, +128) of A and of B for
// every 64-row slab (UMMA M=256 across the pair, each CTA's TMEM holds its 128 rows of the accumulator).  Same ragged
// K-loop; selected when the weight matrices are large enough to fill the pairs.
constexpr int WG2_BN = 256;
constexpr int WG2_STAGE_BYTES = 64 * BM * 2 + 64 * (WG2_BN / 2) * 2;  // 32 KB per CTA per stage

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
wgrad2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const WgradParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG2_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + WG_STAGES;
  uint64_t* tfull_bar = empty_bar + WG_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB);
    for (int i = 0; i < WG_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 2 * EPI_WARPS);
    }
    fence_mbar_init();
  }
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tmem_alloc_2sm(tmem_slot, 2 * WG2_BN);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int mt = (p.Md + 2 * BM - 1) / (2 * BM), nt = (p.Nd + WG2_BN - 1) / WG2_BN;
  const int tiles_per_group = mt * nt;
  const int total = p.G * tiles_per_group;
  auto decode = [&](int t, int& g, int& mi, int& ni) {
    g = t / tiles_per_group;
    const int r = t - g * tiles_per_group;
    mi = r / nt;
    ni = r - mi * nt;
  };
  auto span = [&](int s, int g, int& r0, int& ksteps) {
    r0 = p.offs[s * p.G + g];
    ksteps = (p.offs[s * p.G + g + 1] - r0 + 15) / 16;
  };

  if (warp == 0) {
    if (elect_one()) {  // elect.sync: ptxas keeps the single-thread body on the uniform datapath
      int stage = 0;
      uint32_t phase = 0;
      for (int t = cluster_id; t < total; t += n_clusters) {
        int g, mi, ni;
        decode(t, g, mi, ni);
        const int acol = mi * 2 * BM + static_cast<int>(rank) * BM;
        const int bcol = ni * WG2_BN + static_cast<int>(rank) * (WG2_BN / 2);
        for (int src = 0; src < p.n_src; ++src) {
          int r0, ksteps;
          span(src, g, r0, ksteps);
          const int slabs = (ksteps + 3) / 4;
          for (int s = 0; s < slabs; ++s) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* sa = smem + stage * WG2_STAGE_BYTES;
            uint8_t* sb = sa + 64 * BM * 2;
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * WG2_STAGE_BYTES);
            const int row = r0 + s * 64;
#pragma unroll
            for (int c = 0; c < 2; ++c) tma_load_2d_2sm(sa + c * 8192, &tmA, &full_bar[stage], acol + c * 64, row);
#pragma unroll
            for (int c = 0; c < 2; ++c) tma_load_2d_2sm(sb + c * 8192, &tmB, &full_bar[stage], bcol + c * 64, row);
            if (++stage == WG_STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(2 * BM, WG2_BN, true, true);
      const uint64_t dA0 = make_smem_desc(smem_u32(smem), 8192, 1024);
      const uint64_t dB0 = make_smem_desc(smem_u32(smem) + 64 * BM * 2, 8192, 1024);
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int t = cluster_id; t < total; t += n_clusters, ++it) {
        int g, mi, ni;
        decode(t, g, mi, ni);
        const int as = it & 1;
        mbar_wait(&tempty_bar[as], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * WG2_BN;
        uint32_t started = 0;
        for (int src = 0; src < p.n_src; ++src) {
          int r0, ksteps;
          span(src, g, r0, ksteps);
          const int slabs = (ksteps + 3) / 4;
          for (int s = 0; s < slabs; ++s) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint64_t da = dA0 + stage * (WG2_STAGE_BYTES >> 4), db = dB0 + stage * (WG2_STAGE_BYTES >> 4);
            const int kmax = min(4, ksteps - s * 4);
            for (int k = 0; k < kmax; ++k) {
              umma_bf16_ss_2sm(d_tmem, da + k * (2048 >> 4), db + k * (2048 >> 4), idesc, started);
              started = 1;
            }
            umma_commit_2sm(&empty_bar[stage]);
            if (++stage == WG_STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
        umma_commit_2sm(&tfull_bar[as]);
      }
    }
  } else {
    const int quad = warp & 3, half = (warp - 2) >> 2;
    const uint32_t leader_tempty0 = mapa_shared(smem_u32(&tempty_bar[0]), 0);
    int it = 0;
    for (int t = cluster_id; t < total; t += n_clusters, ++it) {
      int g, mi, ni;
      decode(t, g, mi, ni);
      int ksteps = 0;
      for (int src = 0; src < p.n_src; ++src) {
        int r0, ks;
        span(src, g, r0, ks);
        ksteps += ks;
      }
      const int as = it & 1;
      mbar_wait(&tfull_bar[as], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + as * WG2_BN + (static_cast<uint32_t>(quad * 32) << 16);
      const int m = mi * 2 * BM + static_cast<int>(rank) * BM + quad * 32 + lane;
      __nv_bfloat16* orow = p.out + (static_cast<int64_t>(g) * p.Md + m) * p.Nd + ni * WG2_BN;
#pragma unroll 1
      for (int c = half * (WG2_BN / 2); c < (half + 1) * (WG2_BN / 2); c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + c, v);
        tmem_ld_wait();
        if (ksteps == 0) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0;
        }
        if (m < p.Md) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (ni * WG2_BN + c + q * 8 + 8 <= p.Nd)
              *reinterpret_cast<uint4*>(orow + c + q * 8) =
                  make_uint4(pack_bf16(__uint_as_float(v[q * 8]), __uint_as_float(v[q * 8 + 1])),
                             pack_bf16(__uint_as_float(v[q * 8 + 2]), __uint_as_float(v[q * 8 + 3])),
                             pack_bf16(__uint_as_float(v[q * 8 + 4]), __uint_as_float(v[q * 8 + 5])),
                             pack_bf16(__uint_as_float(v[q * 8 + 6]), __uint_as_float(v[q * 8 + 7])));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(leader_tempty0 + as * 8);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_2sm(tmem_base, 2 * WG2_BN);
}

}  // namespace aria

using namespace aria;

extern "C" int aria_grouped_wgrad(const void* a, int64_t lda, const void* b, int64_t ldb, void* out, const int32_t* group_offsets,
                                  int64_t rows, int64_t md, int64_t nd, int32_t num_groups, int32_t num_sources,
                                  aria_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ARIA_CHECK_ARG(a && b && out && group_offsets && rows >= 0 && md > 0 && nd > 0 && num_groups >= 1 && num_sources >= 1);
  ARIA_CHECK_ARG(md % 8 == 0 && nd % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && lda >= md && ldb >= nd);
  CUtensorMap tmA, tmB;
  // row-major [rows, Md]: inner = feature dim (the MMA's M / N), outer = rows (the contraction)
  int rc = make_tmap_2d(&tmA, a, md, rows > 0 ? rows : 1, lda * 2, 64, 64);
  if (rc) return rc;
  rc = make_tmap_2d(&tmB, b, nd, rows > 0 ? rows : 1, ldb * 2, 64, 64);
  if (rc) return rc;
  WgradParams p{};
  p.Md = static_cast<int>(md);
  p.Nd = static_cast<int>(nd);
  p.G = num_groups;
  p.n_src = num_sources;
  p.offs = group_offsets;
  p.out = static_cast<__nv_bfloat16*>(out);
  constexpr int SMEM = WG_STAGES * WG_STAGE_BYTES + 1024 + 256;
  constexpr int SMEM2 = WG_STAGES * WG2_STAGE_BYTES + 1024 + 256;
  static bool attr_set[kMaxDevices] = {}, attr_set2[kMaxDevices] = {};
  if (ensure_dynamic_smem(attr_set, wgrad_kernel, SMEM) != cudaSuccess) return ARIA_ERR_CUDA;
  if (ensure_dynamic_smem(attr_set2, wgrad2_kernel, SMEM2) != cudaSuccess) return ARIA_ERR_CUDA;
  {  // 2-CTA pairs when there are enough 256 x 256 output tiles to fill them (the large expert / MLP weight matrices)
    const int64_t tiles2 = static_cast<int64_t>(num_groups) * ((md + 2 * BM - 1) / (2 * BM)) * ((nd + WG2_BN - 1) / WG2_BN);
    static int force = -1;
    if (force < 0) {
      const char* ev = getenv("ARIA_GEMM_CTAS");
      force = ev ? atoi(ev) : 0;
    }
    if ((tiles2 >= sm_count() / 2 && md >= 256 && nd >= 256 && force != 1) || force == 2) {
      int clusters = sm_count() / 2;
      if (tiles2 < clusters) clusters = static_cast<int>(tiles2);
      wgrad2_kernel<<<clusters * 2, GEMM_THREADS, SMEM2, stream>>>(tmA, tmB, p);
      return check_launch("wgrad2_kernel");
    }
  }
  const int64_t total = static_cast<int64_t>(num_groups) * ((md + BM - 1) / BM) * ((nd + WG_BN - 1) / WG_BN);
  int grid = sm_count();
  if (total < grid) grid = static_cast<int>(total);
  wgrad_kernel<<<grid, GEMM_THREADS, SMEM, stream>>>(tmA, tmB, p);
  return check_launch("wgrad_kernel");
}
