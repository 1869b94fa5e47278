// Persistent, warp-specialised tcgen05 GEMM for sm_100a (bf16 x bf16 -> fp32 in TMEM -> bf16).
//
//   * warp 0  : TMA producer  (cp.async.bulk.tensor, SWIZZLE_128B tiles, mbarrier complete_tx)
//   * warp 1  : MMA issuer    (one elected thread, tcgen05.mma cta_group::1 M=128 x N=BN x K=16), owns TMEM
//   * warps 2-5: epilogue     (tcgen05.ld 32x32b, fused epilogue, 16-byte global stores)
//   * TMEM accumulators are double buffered (2 x BN columns) so tile i+1's MMAs overlap tile i's epilogue.
//
// One kernel template serves (a) nn.Linear-layout dense GEMMs (B = [N,K], K-major), (b) the reference's
// grouped expert GEMM (aria/model/moe_lm.py:398-428,467-484: B = [E,K,N], consumed N-contiguous through an
// MN-major UMMA descriptor — the HF weight layout is used as is, no repack) with a device-side tile
// scheduler over the expert row offsets (no host sync, cf. moe_lm.py:478), and (c) the fused epilogues:
// bias/activation/residual, SwiGLU (moe_lm.py:505-507), and RoPE + head-major scatter for q/k/v.
#include "common.cuh"
#include "ptx.cuh"

namespace aria {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB
constexpr int GEMM_THREADS = 192;

struct GemmParams {
  int M, N, K;  // N = output columns per segment
  int num_groups;
  const int32_t* group_offsets;
  int n_seg;
  int act;
  const __nv_bfloat16* bias[3];
  const __nv_bfloat16* residual;
  int64_t ldr;
  __nv_bfloat16* out[3];
  int64_t ldo;
  int head_dim, head_ld, rows_per_batch, pos0;
  int64_t stride_b, stride_h;
  int rope_mask;
  const __nv_bfloat16* rope_cos;
  const __nv_bfloat16* rope_sin;
  const int32_t* position_ids;
  int dbg_lbo, dbg_sbo, dbg_kadv;
};

// Monotonic decoder of the persistent tile index -> (group, m-tile, n-tile).  Tiles are ordered group-major,
// then n-tile, with the m-tile innermost so that CTAs running concurrently share the same weight tile.
struct TileSched {
  const int32_t* offs;
  int G, n_tiles, M;
  int g, mt_prefix, g_row0, g_rows, g_mt;
  __device__ void init(const GemmParams& p, int n_tiles_) {
    offs = p.group_offsets;
    G = p.num_groups;
    n_tiles = n_tiles_;
    M = p.M;
    g = -1;
    mt_prefix = 0;
    g_mt = 0;
    g_row0 = 0;
    g_rows = 0;
  }
  __device__ bool load_group(int gi) {
    if (gi >= G) return false;
    if (offs) {
      g_row0 = offs[gi];
      g_rows = offs[gi + 1] - g_row0;
    } else {
      g_row0 = 0;
      g_rows = M;
    }
    g_mt = (g_rows + BM - 1) / BM;
    return true;
  }
  __device__ bool decode(int t, int& grp, int& m_idx, int& n_idx, int& row0, int& rows) {
    if (g < 0) {
      g = 0;
      if (!load_group(0)) return false;
    }
    while (t >= (mt_prefix + g_mt) * n_tiles) {
      mt_prefix += g_mt;
      ++g;
      if (!load_group(g)) return false;
    }
    int r = t - mt_prefix * n_tiles;
    n_idx = r / g_mt;
    m_idx = r - n_idx * g_mt;
    grp = g;
    row0 = g_row0;
    rows = g_rows;
    return true;
  }
};

ARIA_DEVICE float act_apply(float x, int act) {
  if (act == ARIA_ACT_GELU_TANH) {
    // torch gelu(approximate="tanh"): one fp32 evaluation, rounded once by the caller
    const float k0 = 0.7978845608028654f, k1 = 0.044715f;
    float inner = k0 * (x + k1 * x * x * x);
    return 0.5f * x * (1.f + fast_tanh(inner));
  }
  if (act == ARIA_ACT_GELU_NEW) {
    // transformers NewGELUActivation evaluated op by op on bf16 tensors (aria/model/projector.py:40-45):
    // 0.5 * x * (1.0 + tanh(sqrt(2/pi) * (x + 0.044715 * pow(x, 3))))
    float p3 = bf16r(x * x * x);
    float t = bf16r(0.044715f * p3);
    t = bf16r(x + t);
    t = bf16r(0.7978845608028654f * t);
    t = bf16r(fast_tanh(t));
    t = bf16r(1.0f + t);
    float h = bf16r(0.5f * x);
    return h * t;
  }
  return x;
}

template <int BN, bool B_MN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB0,
            const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ CUtensorMap tmB2, const GemmParams p) {
  constexpr int B_STAGE_BYTES = BN * BK * 2;
  constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  constexpr int STAGES = (BN == 256) ? 4 : (BN >= 128 ? 6 : 8);
  constexpr int ACC_STRIDE = (BN <= 32) ? 32 : (BN <= 64) ? 64 : (BN <= 128) ? 128 : 256;  // TMEM columns per accumulator stage
  constexpr int TMEM_COLS = 2 * ACC_STRIDE;
  constexpr int OUT_BN = (EPI == ARIA_EPI_SWIGLU) ? BN / 2 : BN;  // output columns per tile
  static_assert(BN % 16 == 0 && BN >= 16 && BN <= 256, "invalid UMMA N");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;  // [2]
  uint64_t* tempty_bar = tfull_bar + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmA);
    prefetch_tmap(&tmB0);
    if (p.n_seg > 1 || EPI == ARIA_EPI_SWIGLU) prefetch_tmap(&tmB1);
    if (p.n_seg > 2) prefetch_tmap(&tmB2);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_out_total = p.N * (EPI == ARIA_EPI_SWIGLU ? 1 : p.n_seg);  // output columns overall
  const int n_tiles = (n_out_total + OUT_BN - 1) / OUT_BN;
  const int k_blocks = (p.K + BK - 1) / BK;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (lane == 0) {
      TileSched sched;
      sched.init(p, n_tiles);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x;; t += gridDim.x) {
        int grp, m_idx, n_idx, row0, rows;
        if (!sched.decode(t, grp, m_idx, n_idx, row0, rows)) break;
        const int a_row = row0 + m_idx * BM;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_STAGE_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
          tma_load_2d(sa, &tmA, &full_bar[stage], kb * BK, a_row);
          if constexpr (B_MN) {
            // B = [G*K, Ncols] rows k, N contiguous; one box = 64 k-rows x 64 n (8 KB), BN/64 boxes per stage
            const int krow = grp * p.K + kb * BK;
            constexpr int CH = BN / 64;
#pragma unroll
            for (int c = 0; c < CH; ++c) {
              int ncol;
              if constexpr (EPI == ARIA_EPI_SWIGLU) {
                constexpr int HALF = CH / 2;
                ncol = (c < HALF) ? (n_idx * OUT_BN + c * 64) : (p.N + n_idx * OUT_BN + (c - HALF) * 64);
              } else {
                ncol = n_idx * BN + c * 64;
              }
              tma_load_2d(sb + c * (64 * BK * 2), &tmB0, &full_bar[stage], ncol, krow);
            }
          } else {
            if constexpr (EPI == ARIA_EPI_SWIGLU) {
              tma_load_2d(sb, &tmB0, &full_bar[stage], kb * BK, n_idx * OUT_BN);
              tma_load_2d(sb + (BN / 2) * BK * 2, &tmB1, &full_bar[stage], kb * BK, n_idx * OUT_BN);
            } else {
              const int col = n_idx * BN;
              const int seg = col / p.N;
              const CUtensorMap* tb = seg == 0 ? &tmB0 : (seg == 1 ? &tmB1 : &tmB2);
              tma_load_2d(sb, tb, &full_bar[stage], kb * BK, col - seg * p.N);
            }
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, false, B_MN);
      const uint32_t a_lbo = 16, a_sbo = 1024;
      const uint32_t b_lbo = p.dbg_lbo ? p.dbg_lbo : (B_MN ? 64 * BK * 2 : 16);
      const uint32_t b_sbo = p.dbg_sbo ? p.dbg_sbo : 1024;
      const uint32_t b_kadv = p.dbg_kadv ? p.dbg_kadv : (B_MN ? 16 * 128 : 32);
      TileSched sched;
      sched.init(p, n_tiles);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = blockIdx.x;; t += gridDim.x, ++it) {
        int grp, m_idx, n_idx, row0, rows;
        if (!sched.decode(t, grp, m_idx, n_idx, row0, rows)) break;
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * ACC_STRIDE;
        for (int kb = 0; kb < k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sb = sa + A_STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            uint64_t da = make_smem_desc(sa + k * 32, a_lbo, a_sbo);
            uint64_t db = make_smem_desc(sb + k * b_kadv, b_lbo, b_sbo);
            umma_bf16_ss(d_tmem, da, db, idesc, (kb | k) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees the smem slot once these MMAs have read it
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit(&tfull_bar[as]);  // accumulator complete -> epilogue
      }
    }
  } else {
    // =========================== epilogue (warps 2..5) ===========================
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int r_in_tile = quad * 32 + lane;
    TileSched sched;
    sched.init(p, n_tiles);
    int it = 0;
    for (int t = blockIdx.x;; t += gridDim.x, ++it) {
      int grp, m_idx, n_idx, row0, rows;
      if (!sched.decode(t, grp, m_idx, n_idx, row0, rows)) break;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + as * ACC_STRIDE + (static_cast<uint32_t>(quad * 32) << 16);
      const int r_in_grp = m_idx * BM + r_in_tile;
      const bool row_ok = r_in_grp < rows;
      const int64_t grow = static_cast<int64_t>(row0) + r_in_grp;

      if constexpr (EPI == ARIA_EPI_SWIGLU) {
        // out[:, n] = bf16( bf16(silu(bf16(gate))) * bf16(up) )  — rounding points of moe_lm.py:505-507
        __nv_bfloat16* orow = p.out[0] + grow * p.ldo + n_idx * OUT_BN;
#pragma unroll 1
        for (int c = 0; c < OUT_BN; c += 32) {
          uint32_t g[32], u[32];
          tmem_ld_32x32(taddr + c, g);
          tmem_ld_32x32(taddr + OUT_BN + c, u);
          tmem_ld_wait();
          uint32_t o[16];
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            float g0 = bf16r(__uint_as_float(g[j])), g1 = bf16r(__uint_as_float(g[j + 1]));
            float u0 = bf16r(__uint_as_float(u[j])), u1 = bf16r(__uint_as_float(u[j + 1]));
            float s0 = bf16r(fast_silu(g0)), s1 = bf16r(fast_silu(g1));
            o[j >> 1] = pack_bf16(s0 * u0, s1 * u1);
          }
          if (row_ok) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int col = n_idx * OUT_BN + c + q * 8;
              if (col + 8 <= p.N)
                *reinterpret_cast<uint4*>(orow + c + q * 8) = make_uint4(o[q * 4], o[q * 4 + 1], o[q * 4 + 2], o[q * 4 + 3]);
            }
          }
        }
      } else if constexpr (EPI == ARIA_EPI_LINEAR) {
        const int col0 = n_idx * BN;
        const int seg = col0 / p.N;  // bias is per segment; out is [m, n_seg*n]
        const __nv_bfloat16* bias = p.bias[seg];
        __nv_bfloat16* orow = p.out[0] + grow * p.ldo + col0;
        const __nv_bfloat16* rrow = p.residual ? p.residual + grow * p.ldr + col0 : nullptr;
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + c, v);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int col = col0 + c + q * 8;
            if (col + 8 > n_out_total) continue;
            float x[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = __uint_as_float(v[q * 8 + j]);
            if (bias) {
              uint4 bv = *reinterpret_cast<const uint4*>(bias + (col - seg * p.N));
              const uint32_t bw[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                x[2 * j] += bf16_lo(bw[j]);
                x[2 * j + 1] += bf16_hi(bw[j]);
              }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = bf16r(x[j]);
            if (p.act != ARIA_ACT_NONE) {
#pragma unroll
              for (int j = 0; j < 8; ++j) x[j] = bf16r(act_apply(x[j], p.act));
            }
            if (row_ok) {
              if (rrow) {
                uint4 rv = *reinterpret_cast<const uint4*>(rrow + c + q * 8);
                const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  x[2 * j] += bf16_lo(rw[j]);
                  x[2 * j + 1] += bf16_hi(rw[j]);
                }
              }
              *reinterpret_cast<uint4*>(orow + c + q * 8) =
                  make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
            }
          }
        }
      } else {  // ARIA_EPI_HEADS
        const int col0 = n_idx * BN;
        const int seg = col0 / p.N;
        const int cseg0 = col0 - seg * p.N;
        const __nv_bfloat16* bias = p.bias[seg];
        const int b = static_cast<int>(grow / p.rows_per_batch);
        const int tok = static_cast<int>(grow - static_cast<int64_t>(b) * p.rows_per_batch);
        __nv_bfloat16* obase = p.out[seg] + b * p.stride_b + static_cast<int64_t>(p.pos0 + tok) * p.head_ld;
        const bool rope = (p.rope_mask >> seg) & 1;
        if (rope) {
          // BN == head_dim == 128: this tile is exactly one head. rotate-half RoPE, op-by-op bf16 rounding:
          //   out = bf16(bf16(x*cos) + bf16(rotate_half(x)*sin))
          const int head = cseg0 / p.head_dim;
          __nv_bfloat16* orow = obase + head * p.stride_h;
          const int pos = row_ok ? (p.position_ids ? p.position_ids[grow] : p.pos0 + tok) : 0;
          const __nv_bfloat16* cs = p.rope_cos + static_cast<int64_t>(pos) * p.head_dim;
          const __nv_bfloat16* sn = p.rope_sin + static_cast<int64_t>(pos) * p.head_dim;
#pragma unroll 1
          for (int c = 0; c < 64; c += 32) {
            uint32_t lo[32], hi[32];
            tmem_ld_32x32(taddr + c, lo);
            tmem_ld_32x32(taddr + 64 + c, hi);
            tmem_ld_wait();
            if (row_ok) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                uint4 c_lo = *reinterpret_cast<const uint4*>(cs + c + q * 8);
                uint4 s_lo = *reinterpret_cast<const uint4*>(sn + c + q * 8);
                uint4 c_hi = *reinterpret_cast<const uint4*>(cs + 64 + c + q * 8);
                uint4 s_hi = *reinterpret_cast<const uint4*>(sn + 64 + c + q * 8);
                const uint32_t cl[4] = {c_lo.x, c_lo.y, c_lo.z, c_lo.w}, sl[4] = {s_lo.x, s_lo.y, s_lo.z, s_lo.w};
                const uint32_t ch[4] = {c_hi.x, c_hi.y, c_hi.z, c_hi.w}, sh[4] = {s_hi.x, s_hi.y, s_hi.z, s_hi.w};
                float ol[8], oh[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  float xl = bf16r(__uint_as_float(lo[q * 8 + j]));
                  float xh = bf16r(__uint_as_float(hi[q * 8 + j]));
                  float cosl = (j & 1) ? bf16_hi(cl[j >> 1]) : bf16_lo(cl[j >> 1]);
                  float sinl = (j & 1) ? bf16_hi(sl[j >> 1]) : bf16_lo(sl[j >> 1]);
                  float cosh_ = (j & 1) ? bf16_hi(ch[j >> 1]) : bf16_lo(ch[j >> 1]);
                  float sinh_ = (j & 1) ? bf16_hi(sh[j >> 1]) : bf16_lo(sh[j >> 1]);
                  ol[j] = bf16r(xl * cosl) + bf16r(-xh * sinl);
                  oh[j] = bf16r(xh * cosh_) + bf16r(xl * sinh_);
                }
                *reinterpret_cast<uint4*>(orow + c + q * 8) =
                    make_uint4(pack_bf16(ol[0], ol[1]), pack_bf16(ol[2], ol[3]), pack_bf16(ol[4], ol[5]), pack_bf16(ol[6], ol[7]));
                *reinterpret_cast<uint4*>(orow + 64 + c + q * 8) =
                    make_uint4(pack_bf16(oh[0], oh[1]), pack_bf16(oh[2], oh[3]), pack_bf16(oh[4], oh[5]), pack_bf16(oh[6], oh[7]));
              }
            }
          }
        } else {
#pragma unroll 1
          for (int c = 0; c < BN; c += 32) {
            uint32_t v[32];
            if (c + 32 <= BN) {
              tmem_ld_32x32(taddr + c, v);
            } else {  // BN not a multiple of 32 (e.g. 144): the tail re-reads an overlapping window
              tmem_ld_32x32(taddr + BN - 32, v);
            }
            tmem_ld_wait();
            const int cbase = (c + 32 <= BN) ? c : BN - 32;
            const int qstart = (c + 32 <= BN) ? 0 : (c - cbase) / 8;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (q < qstart) continue;
              const int cs_ = cseg0 + cbase + q * 8;  // column inside the segment
              if (cs_ + 8 > p.N) continue;
              float x[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) x[j] = __uint_as_float(v[q * 8 + j]);
              if (bias) {
                uint4 bv = *reinterpret_cast<const uint4*>(bias + cs_);
                const uint32_t bw[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  x[2 * j] += bf16_lo(bw[j]);
                  x[2 * j + 1] += bf16_hi(bw[j]);
                }
              }
              if (row_ok) {
                const int head = cs_ / p.head_dim;
                const int d = cs_ - head * p.head_dim;
                *reinterpret_cast<uint4*>(obase + head * p.stride_h + d) =
                    make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
              }
            }
          }
        }
      }
      // release this accumulator stage back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
  (void)ACC_STRIDE;
}

template <int BN, bool B_MN, int EPI>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap* tmB, const GemmParams& p, int max_tiles,
                       cudaStream_t stream) {
  constexpr int B_STAGE_BYTES = BN * BK * 2;
  constexpr int STAGES = (BN == 256) ? 4 : (BN >= 128 ? 6 : 8);
  constexpr int SMEM = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 1024 /*align*/ + 256 /*barriers*/;
  auto kern = gemm_kernel<BN, B_MN, EPI>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) {
      fprintf(stderr, "aria_b200: cudaFuncSetAttribute(smem=%d) failed: %s\n", SMEM, cudaGetErrorString(e));
      return ARIA_ERR_CUDA;
    }
    attr_set = true;
  }
  int grid = sm_count();
  if (max_tiles < grid) grid = max_tiles;
  if (grid < 1) grid = 1;
  kern<<<grid, GEMM_THREADS, SMEM, stream>>>(tmA, tmB[0], tmB[1], tmB[2], p);
  return check_launch("gemm_kernel");
}

}  // namespace aria

using namespace aria;

extern "C" int aria_gemm(const aria_gemm_desc_t* d, aria_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ARIA_CHECK_ARG(d != nullptr);
  ARIA_CHECK_ARG(d->a && d->b[0] && d->out[0]);
  ARIA_CHECK_ARG(d->m >= 0 && d->n > 0 && d->k > 0);
  ARIA_CHECK_ARG(d->n % 8 == 0 && d->k % 8 == 0 && d->lda % 8 == 0);
  ARIA_CHECK_ARG(d->n_seg >= 1 && d->n_seg <= 3);
  ARIA_CHECK_ARG(d->num_groups >= 1);
  ARIA_CHECK_ARG((reinterpret_cast<uintptr_t>(d->a) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->out[0]) & 15) == 0);
  if (d->m == 0) return ARIA_OK;
  const bool b_mn = d->b_layout == ARIA_B_GKN;
  const bool swiglu = d->epilogue == ARIA_EPI_SWIGLU;
  if (b_mn) {
    ARIA_CHECK_ARG(d->n_seg == 1);
    ARIA_CHECK_ARG(d->k % BK == 0);  // k-blocks must not straddle experts in the flattened [G*K, N] view
    ARIA_CHECK_ARG(d->n % 64 == 0);
    ARIA_CHECK_ARG(d->epilogue != ARIA_EPI_HEADS);
  } else {
    ARIA_CHECK_ARG(d->num_groups == 1 || d->n_seg == 1);
    if (swiglu) ARIA_CHECK_ARG(d->n_seg == 2 && d->b[1]);
  }
  if (d->num_groups > 1) ARIA_CHECK_ARG(d->group_offsets != nullptr);

  GemmParams p{};
  p.M = static_cast<int>(d->m);
  p.N = static_cast<int>(d->n);
  p.K = static_cast<int>(d->k);
  p.num_groups = d->num_groups;
  p.group_offsets = d->group_offsets;
  p.n_seg = swiglu ? 1 : d->n_seg;
  p.act = d->act;
  for (int i = 0; i < 3; ++i) {
    p.bias[i] = static_cast<const __nv_bfloat16*>(d->bias[i]);
    p.out[i] = static_cast<__nv_bfloat16*>(d->out[i]);
  }
  p.residual = static_cast<const __nv_bfloat16*>(d->residual);
  p.ldr = d->ldr;
  p.ldo = d->ldo;
  p.head_dim = d->head_dim;
  p.head_ld = d->head_ld;
  p.rows_per_batch = d->rows_per_batch > 0 ? d->rows_per_batch : static_cast<int>(d->m);
  p.pos0 = d->pos0;
  p.stride_b = d->stride_b;
  p.stride_h = d->stride_h;
  p.rope_mask = d->rope_mask;
  p.rope_cos = static_cast<const __nv_bfloat16*>(d->rope_cos);
  p.rope_sin = static_cast<const __nv_bfloat16*>(d->rope_sin);
  p.position_ids = d->position_ids;
  p.dbg_lbo = d->dbg_lbo;
  p.dbg_sbo = d->dbg_sbo;
  p.dbg_kadv = d->dbg_kadv;

  // ---- tile shape selection
  int BN;
  if (d->epilogue == ARIA_EPI_HEADS) {
    ARIA_CHECK_ARG(d->head_dim > 0 && d->head_dim % 8 == 0 && d->head_ld >= d->head_dim && d->n % d->head_dim == 0);
    if (d->rope_mask) {
      ARIA_CHECK_ARG(d->head_dim == 128 && d->rope_cos && d->rope_sin);
      BN = 128;
    } else {
      BN = (d->n % 128 == 0) ? 128 : 144;
      ARIA_CHECK_ARG(d->n % BN == 0);
    }
    for (int s = 0; s < d->n_seg; ++s) ARIA_CHECK_ARG(d->out[s] && d->b[s]);
  } else if (swiglu) {
    BN = 128;  // 64 gate + 64 up columns per tile
    ARIA_CHECK_ARG(d->n % 64 == 0);
  } else {
    BN = 128;
    if (d->n_seg > 1) ARIA_CHECK_ARG(d->n % BN == 0);
  }

  CUtensorMap tmA, tmB[3];
  int rc = make_tmap_2d(&tmA, d->a, d->k, d->m, d->lda * 2, BK, BM);
  if (rc) return rc;
  if (b_mn) {
    const uint64_t ncols = swiglu ? 2 * d->n : d->n;
    rc = make_tmap_2d(&tmB[0], d->b[0], ncols, static_cast<uint64_t>(d->num_groups) * d->k, ncols * 2, 64, BK);
    if (rc) return rc;
    tmB[1] = tmB[0];
    tmB[2] = tmB[0];
  } else {
    const int nb = swiglu ? 2 : d->n_seg;
    const uint32_t box_rows = swiglu ? BN / 2 : BN;
    for (int s = 0; s < 3; ++s) {
      const void* ptr = s < nb ? d->b[s] : d->b[0];
      rc = make_tmap_2d(&tmB[s], ptr, d->k, d->n, d->k * 2, BK, box_rows);
      if (rc) return rc;
    }
  }

  const int out_bn = swiglu ? BN / 2 : BN;
  const int64_t n_out_total = d->n * (swiglu ? 1 : d->n_seg);
  const int64_t n_tiles = (n_out_total + out_bn - 1) / out_bn;
  const int64_t m_tiles_ub = (d->m + BM - 1) / BM + (d->num_groups > 1 ? d->num_groups : 0);
  int64_t max_tiles = n_tiles * m_tiles_ub;
  if (max_tiles > (1 << 30)) max_tiles = 1 << 30;

#define ARIA_LAUNCH(BN_, MN_, EPI_) return launch_gemm<BN_, MN_, EPI_>(tmA, tmB, p, static_cast<int>(max_tiles), stream)
  if (d->epilogue == ARIA_EPI_HEADS) {
    if (BN == 128) ARIA_LAUNCH(128, false, ARIA_EPI_HEADS);
    ARIA_LAUNCH(144, false, ARIA_EPI_HEADS);
  }
  if (swiglu) {
    if (b_mn) ARIA_LAUNCH(128, true, ARIA_EPI_SWIGLU);
    ARIA_LAUNCH(128, false, ARIA_EPI_SWIGLU);
  }
  if (b_mn) ARIA_LAUNCH(128, true, ARIA_EPI_LINEAR);
  ARIA_LAUNCH(128, false, ARIA_EPI_LINEAR);
#undef ARIA_LAUNCH
}

extern "C" int aria_grouped_gemm(const void* a, const void* b, void* out, const int32_t* group_offsets, int64_t rows,
                                 int64_t k, int64_t n, int32_t num_groups, aria_stream_t stream) {
  aria_gemm_desc_t d{};
  d.a = a;
  d.lda = k;
  d.m = rows;
  d.n = n;
  d.k = k;
  d.b[0] = b;
  d.n_seg = 1;
  d.b_layout = ARIA_B_GKN;
  d.num_groups = num_groups;
  d.group_offsets = group_offsets;
  d.epilogue = ARIA_EPI_LINEAR;
  d.out[0] = out;
  d.ldo = n;
  return aria_gemm(&d, stream);
}

extern "C" int aria_abi_version(void) { return 1; }
extern "C" const char* aria_build_arch(void) { return "sm_100a"; }
