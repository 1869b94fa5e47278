// Attention for sm_100a.
//
// NOTE (round 2): `aria_attention_fwd` now lives in attention_v3.cu.  The round-1 kernel below stays exported as
// `aria_attention_fwd_v2` (not part of the C ABI in include/) for A/B timing in scripts/; the decode kernels are current.
//
// aria_attention_fwd_v2: flash-style forward with both contractions on tcgen05 tensor cores:
//     S = Q K^T   (A = Q tile, B = K tile, both K-major SW128 in shared memory, accumulator S in TMEM)
//     O += P V    (A = P, bf16, written by the softmax warps back into TMEM over S and consumed straight from
//                  TMEM; B = V tile consumed MN-major — V is [keys, d] with d contiguous, exactly the HF cache
//                  layout — O accumulates in TMEM)
//   warp 0: TMA producer (Q once, K/V double buffered); warp 1: MMA issuer + TMEM owner;
//   warps 2-9: softmax (one query row per thread, whole S row in registers, fp32 online softmax with a
//   warp-uniform lazy rescale of O) and the final normalise + store.
// Replaces flash_attn_func / SDPA behind LLAMA_ATTENTION_CLASSES (aria/model/moe_lm.py:594) and the
// Idefics2 / nn.MultiheadAttention attention of the ViT + projector (vision_encoder.py:120,
// projector.py:93) with head_dim padded 72 -> 128.
//
// aria_attention_decode: single-query attention against the KV cache; HBM-bound, CUDA cores, split-KV.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace aria {

constexpr int AT_BM = 128;   // queries per CTA
constexpr int AT_BN = 128;   // keys per step
constexpr int AT_D = 128;    // head dim
constexpr int AT_TILE = AT_BM * AT_D * 2;  // 32 KB
constexpr int AT_HALF = AT_TILE / 2;       // one SW128 atom column: [128 rows][64 bf16]

ARIA_DEVICE float fast_exp2(float x) { return fast_ex2(x); }

// (A polynomial exp2 on the FMA pipes for a quarter of the elements — the FA4 trick — was measured and LOST 20 %.  That
// measurement predates the lean MMA-issue loop: at the time the issuing thread, not the softmax, set the pace, so the
// trick only added instructions.  To be re-measured; see profiles/r01_attention_notes.txt.)
struct AttnParams {
  int B, H, Tq, Tk;
  int out_hd;
  int hd_eff;  // head dim actually contracted / produced (multiple of 16): 128 for the LM, 80 for the 72-wide ViT heads
  float scale_log2;
  int causal;
  const uint8_t* key_mask;  // [B, Tk] 1 = masked out
  __nv_bfloat16* out;       // [B, Tq, H*out_hd]
  int n_q_tiles;
};

// ------------------------------------------------------------------------------------------------
// Two 128-row query tiles per CTA (256 queries) ping-pong on the tensor core (FA4-style):
//   TMEM: S0 | S1 | O0 | O1 (4 x 128 columns).  P_t (bf16) aliases the first 64 columns of S_t and is fed to the
//   PV MMA straight from TMEM (tcgen05.mma A-from-TMEM), so P never touches shared memory.
//   warp 0 TMA, warp 1 MMA issuer, warps 2-5 softmax of tile 0, warps 6-9 softmax of tile 1.
//   MMA issue order per key block j:  PV0(j) QK0(j+1) PV1(j) QK1(j+1)  — tile 1's softmax runs under tile 0's MMAs
//   and vice versa; tcgen05 ops execute in issue order, which is what makes the S/P aliasing safe.
constexpr int A2_THREADS = 320;
#ifndef ARIA_ATTN_P_SPLIT
#define ARIA_ATTN_P_SPLIT 2
#endif
// P is handed to the PV MMA in P_SPLIT key chunks with one mbarrier each, so the first PV k-steps run on the tensor
// core while the softmax warps are still exponentiating the rest of the row (shortens the S -> P -> PV chain).
constexpr int P_SPLIT = ARIA_ATTN_P_SPLIT;  // 1, 2 or 4
constexpr int A2_SMEM = 2 * AT_TILE /*Q*/ + 2 * AT_TILE /*K*/ + 2 * AT_TILE /*V*/ + 1024 + 256;

__global__ void __launch_bounds__(A2_THREADS, 1)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                 // [2 tiles][32 KB]
  uint8_t* sK = sQ + 2 * AT_TILE;     // [2 stages][32 KB]
  uint8_t* sV = sK + 2 * AT_TILE;     // [2 stages][32 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * AT_TILE);
  uint64_t* q_full = bars;         // [1]
  uint64_t* k_full = bars + 1;     // [2]
  uint64_t* v_full = bars + 3;     // [2]
  uint64_t* kv_empty = bars + 5;   // [2]
  uint64_t* s_full = bars + 7;     // [2] per tile
  uint64_t* p_full = bars + 13;    // [2 tiles][P_SPLIT]
  uint64_t* o_full = bars + 11;    // [2] per tile
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13 + 2 * P_SPLIT);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.x % (p.B * p.H);
  const int q_pair = p.n_q_tiles - 1 - blockIdx.x / (p.B * p.H);  // n_q_tiles = number of 256-row pairs here
  const int b = bh / p.H, h = bh % p.H;
  const int q0 = q_pair * 2 * AT_BM;
  const int pos_off = p.Tk - p.Tq;
  const bool act1 = q0 + AT_BM < p.Tq;  // second tile has at least one valid row
  const int n_kv_all = (p.Tk + AT_BN - 1) / AT_BN;
  int n_kv0 = n_kv_all, n_kv1 = act1 ? n_kv_all : 0;
  if (p.causal) {
    n_kv0 = min(n_kv_all, (pos_off + min(q0 + AT_BM, p.Tq) - 1) / AT_BN + 1);
    if (act1) n_kv1 = min(n_kv_all, (pos_off + min(q0 + 2 * AT_BM, p.Tq) - 1) / AT_BN + 1);
  }
  const int n_kv = max(n_kv0, n_kv1);

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&kv_empty[i], 1);
      mbar_init(&s_full[i], 1);
      for (int c = 0; c < P_SPLIT; ++c) mbar_init(&p_full[i * P_SPLIT + c], 128);
      mbar_init(&o_full[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {  // elect.sync: ptxas keeps the single-thread body on the uniform datapath
      mbar_arrive_expect_tx(q_full, act1 ? 2 * AT_TILE : AT_TILE);
      tma_load_4d(sQ, &tmQ, q_full, 0, q0, h, b);
      tma_load_4d(sQ + AT_HALF, &tmQ, q_full, 64, q0, h, b);
      if (act1) {
        tma_load_4d(sQ + AT_TILE, &tmQ, q_full, 0, q0 + AT_BM, h, b);
        tma_load_4d(sQ + AT_TILE + AT_HALF, &tmQ, q_full, 64, q0 + AT_BM, h, b);
      }
      for (int j = 0; j < n_kv; ++j) {
        const int s = j & 1;
        mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&k_full[s], AT_TILE);
        tma_load_4d(sK + s * AT_TILE, &tmK, &k_full[s], 0, j * AT_BN, h, b);
        tma_load_4d(sK + s * AT_TILE + AT_HALF, &tmK, &k_full[s], 64, j * AT_BN, h, b);
        mbar_arrive_expect_tx(&v_full[s], AT_TILE);
        tma_load_4d(sV + s * AT_TILE, &tmV, &v_full[s], 0, j * AT_BN, h, b);
        tma_load_4d(sV + s * AT_TILE + AT_HALF, &tmV, &v_full[s], 64, j * AT_BN, h, b);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {  // elect.sync: ptxas keeps the single-thread body on the uniform datapath
      constexpr uint32_t idesc_qk = make_idesc_bf16(AT_BM, AT_BN, false, false);
      const uint32_t idesc_pv = make_idesc_bf16(AT_BM, p.hd_eff, false, true);  // N = hd_eff output columns
      const int qk_steps = p.hd_eff / 16;                                       // columns >= hd_eff are zero padding
      // This thread's instruction stream is the critical path of the whole kernel (32 MMAs per key block for the two
      // tiles): descriptors are base + offset adds, barrier addresses plain 32-bit values.
      const uint64_t dQ0 = make_smem_desc(smem_u32(sQ), 16, 1024);
      const uint64_t dK0 = make_smem_desc(smem_u32(sK), 16, 1024);
      const uint64_t dV0 = make_smem_desc(smem_u32(sV), AT_HALF, 1024);
      const uint32_t s_full0 = smem_u32(s_full), p_full0 = smem_u32(p_full);
      auto issue_qk = [&](int t, int j) {
        const uint64_t dQ = dQ0 + t * (AT_TILE >> 4);
        const uint64_t dK = dK0 + (j & 1) * (AT_TILE >> 4);
        const uint32_t tS = tmem_base + t * 128;
#pragma unroll
        for (int k = 0; k < AT_D / 16; ++k) {
          if (k >= qk_steps) break;
          const uint32_t off = ((k >> 2) * AT_HALF + (k & 3) * 32) >> 4;
          umma_bf16_ss(tS, dQ + off, dK + off, idesc_qk, k ? 1u : 0u);
        }
        umma_commit_addr(s_full0 + t * 8);
      };
      auto issue_pv = [&](int t, int j) {
        const uint64_t dV = dV0 + (j & 1) * (AT_TILE >> 4);
        const uint32_t tP = tmem_base + t * 128;        // bf16 P aliases S_t (2 elements per 32-bit column)
        const uint32_t tO = tmem_base + 256 + t * 128;
        constexpr int KS = (AT_BN / 16) / P_SPLIT;  // k-steps per P chunk
#pragma unroll
        for (int c = 0; c < P_SPLIT; ++c) {
          mbar_wait_addr(p_full0 + (t * P_SPLIT + c) * 8, j & 1);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < KS; ++kk) {
            const int k = c * KS + kk;
            umma_bf16_ts(tO, tP + k * 8, dV + k * (2048 >> 4), idesc_pv, (j | k) ? 1u : 0u);
          }
        }
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_qk(0, 0);
      if (act1) issue_qk(1, 0);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j & 1;
        mbar_wait(&v_full[s], (j >> 1) & 1);
        bool k_next_ready = false;
        if (j < n_kv0) {
          issue_pv(0, j);
          if (j == n_kv0 - 1) umma_commit(&o_full[0]);
          if (j + 1 < n_kv0) {
            mbar_wait(&k_full[(j + 1) & 1], ((j + 1) >> 1) & 1);
            k_next_ready = true;
            tc_fence_after();
            issue_qk(0, j + 1);
          }
        }
        if (j < n_kv1) {
          issue_pv(1, j);
          if (j == n_kv1 - 1) umma_commit(&o_full[1]);
          if (j + 1 < n_kv1) {
            if (!k_next_ready) mbar_wait(&k_full[(j + 1) & 1], ((j + 1) >> 1) & 1);
            tc_fence_after();
            issue_qk(1, j + 1);
          }
        }
        umma_commit(&kv_empty[s]);  // everything that read K(j)/V(j) has been issued before this point
      }
    }
  } else {
    const int t = (warp - 2) >> 2;  // query tile of this softmax warpgroup
    const int n_kv_t = t == 0 ? n_kv0 : n_kv1;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const int q = q0 + t * AT_BM + r;
    const bool row_ok = q < p.Tq;
    const int qpos = pos_off + q;
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t tS = tmem_base + t * 128 + lane_addr;
    const uint32_t tO = tmem_base + 256 + t * 128 + lane_addr;
    const uint8_t* km = p.key_mask ? p.key_mask + static_cast<int64_t>(b) * p.Tk : nullptr;
    float m_ref = -INFINITY, l = 0.f;

    for (int j = 0; j < n_kv_t; ++j) {
      mbar_wait(&s_full[t], j & 1);  // also implies PV_t(j-1) completed (commit covers all earlier MMAs)
      tc_fence_after();
      const int k0 = j * AT_BN;
      const bool need_mask = (k0 + AT_BN > p.Tk) || (p.causal && (k0 + AT_BN - 1 > pos_off + q0 + t * AT_BM)) || km != nullptr;
      // whole S row (128 fp32) into registers with one wait
      uint32_t sr[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_32x32(tS + c * 32, sr[c]);
      tmem_ld_wait();
      if (need_mask) {  // rare path (diagonal / tail / padded keys): -inf on dead keys
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int kc = k0 + c * 32 + i;
            const bool dead = kc >= p.Tk || (p.causal && kc > qpos) || (km && kc < p.Tk && km[kc]);
            if (dead) sr[c][i] = 0xff800000u;
          }
        }
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        mx0 = fmaxf(mx0, __uint_as_float(sr[0][i]));
        mx1 = fmaxf(mx1, __uint_as_float(sr[1][i]));
        mx2 = fmaxf(mx2, __uint_as_float(sr[2][i]));
        mx3 = fmaxf(mx3, __uint_as_float(sr[3][i]));
      }
      const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      const float m_new = fmaxf(m_ref, mx * p.scale_log2);
      const bool want = (m_new - m_ref > 8.0f) || (m_ref == -INFINITY && m_new > -INFINITY);
      if (__any_sync(0xffffffffu, want)) {
        const float f = (m_ref == -INFINITY) ? 0.f : fast_exp2(m_ref - m_new);
        l *= f;
        m_ref = m_new;
        if (j > 0) {
#pragma unroll 1
          for (int c = 0; c < p.hd_eff; c += 32) {
            uint32_t v[32];
            tmem_ld_32x32(tO + c, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * f);
            tmem_st_32x32(tO + c, v);
          }
        }
      }
      const float neg_m = (m_ref == -INFINITY) ? 0.f : -m_ref;
      float l0 = 0.f, l1 = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          const float x0 = fmaf(__uint_as_float(sr[c][i]), p.scale_log2, neg_m);
          const float x1 = fmaf(__uint_as_float(sr[c][i + 1]), p.scale_log2, neg_m);
          const float p0 = fast_exp2(x0);
          const float p1 = fast_exp2(x1);
          l0 += p0;
          l1 += p1;
          pk[i >> 1] = pack_bf16(p0, p1);
        }
        // P chunk (32 keys = 16 packed columns) overwrites S columns [16c, 16c+16): S is already in registers
        tmem_st_32x16(tS + c * 16, pk);
        if ((c + 1) % (4 / P_SPLIT) == 0) {  // chunk of 128/P_SPLIT keys complete -> hand it to the PV MMA
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(&p_full[t * P_SPLIT + c / (4 / P_SPLIT)]);
        }
      }
      l += l0 + l1;
    }
    if (n_kv_t > 0) {
      mbar_wait(&o_full[t], 0);
      tc_fence_after();
      const float inv_l = l > 0.f ? 1.0f / l : 0.f;
      __nv_bfloat16* orow = p.out + (static_cast<int64_t>(b) * p.Tq + q) * (static_cast<int64_t>(p.H) * p.out_hd) + h * p.out_hd;
#pragma unroll 1
      for (int c = 0; c < p.out_hd; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tO + c, v);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (c + g * 8 + 8 <= p.out_hd) {
              float x[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) x[i] = __uint_as_float(v[g * 8 + i]) * inv_l;
              *reinterpret_cast<uint4*>(orow + c + g * 8) =
                  make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// Decode: one query per (b,h); block = 4 warps over a contiguous chunk of keys; each lane owns 4 dims.
// Partial (m, l, acc[128]) per (b,h,split) -> workspace; a second kernel merges the splits.
constexpr int DEC_SPLIT_KEYS = 256;

__global__ void __launch_bounds__(128) attn_decode_partial(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ kc,
                                                           const __nv_bfloat16* __restrict__ vc, const uint8_t* __restrict__ key_mask,
                                                           float* __restrict__ ws, int H, int Tk,
                                                           int64_t q_stride_b, int64_t q_stride_h, int64_t kv_stride_b,
                                                           int64_t kv_stride_h, float scale_log2, int splits) {
  const int bh = blockIdx.x, split = blockIdx.y;
  const int b = bh / H, h = bh % H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k_begin = split * DEC_SPLIT_KEYS, k_end = min(Tk, k_begin + DEC_SPLIT_KEYS);
  const __nv_bfloat16* kbase = kc + b * kv_stride_b + h * kv_stride_h;
  const __nv_bfloat16* vbase = vc + b * kv_stride_b + h * kv_stride_h;
  const uint8_t* km = key_mask ? key_mask + static_cast<int64_t>(b) * Tk : nullptr;
  const uint2 qv = *reinterpret_cast<const uint2*>(q + b * q_stride_b + h * q_stride_h + lane * 4);
  const float q0 = bf16_lo(qv.x) * scale_log2, q1 = bf16_hi(qv.x) * scale_log2, q2 = bf16_lo(qv.y) * scale_log2,
              q3 = bf16_hi(qv.y) * scale_log2;
  float m = -INFINITY, l = 0.f, a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (int k0 = k_begin + warp * 4; k0 < k_end; k0 += 16) {
    float s[4];
    uint2 vv[4];
    bool live[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int kk = k0 + u;
      live[u] = kk < k_end && !(km && km[kk]);
      if (live[u]) {
        const uint2 kv = __ldg(reinterpret_cast<const uint2*>(kbase + static_cast<int64_t>(kk) * AT_D + lane * 4));
        vv[u] = __ldg(reinterpret_cast<const uint2*>(vbase + static_cast<int64_t>(kk) * AT_D + lane * 4));
        s[u] = q0 * bf16_lo(kv.x) + q1 * bf16_hi(kv.x) + q2 * bf16_lo(kv.y) + q3 * bf16_hi(kv.y);
      } else {
        s[u] = 0.f;
        vv[u] = make_uint2(0, 0);
      }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
#pragma unroll
      for (int u = 0; u < 4; ++u) s[u] += __shfl_xor_sync(0xffffffffu, s[u], o);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (live[u]) {
        const float m_new = fmaxf(m, s[u]);
        const float f = exp2f(m - m_new), pw = exp2f(s[u] - m_new);
        l = l * f + pw;
        a0 = a0 * f + pw * bf16_lo(vv[u].x);
        a1 = a1 * f + pw * bf16_hi(vv[u].x);
        a2 = a2 * f + pw * bf16_lo(vv[u].y);
        a3 = a3 * f + pw * bf16_hi(vv[u].y);
        m = m_new;
      }
    }
  }
  // merge the 4 warps through shared memory
  __shared__ float sm_m[4], sm_l[4], sm_a[4][AT_D];
  if (lane == 0) {
    sm_m[warp] = m;
    sm_l[warp] = l;
  }
  sm_a[warp][lane * 4 + 0] = a0;
  sm_a[warp][lane * 4 + 1] = a1;
  sm_a[warp][lane * 4 + 2] = a2;
  sm_a[warp][lane * 4 + 3] = a3;
  __syncthreads();
  const int d = threadIdx.x;  // 128 threads = 128 dims
  float M = fmaxf(fmaxf(sm_m[0], sm_m[1]), fmaxf(sm_m[2], sm_m[3]));
  float L = 0.f, A = 0.f;
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    const float f = (sm_m[w] == -INFINITY) ? 0.f : exp2f(sm_m[w] - M);
    L += sm_l[w] * f;
    A += sm_a[w][d] * f;
  }
  float* o = ws + (static_cast<int64_t>(bh) * splits + split) * (AT_D + 2);
  o[d] = A;
  if (d == 0) {
    o[AT_D] = M;
    o[AT_D + 1] = L;
  }
}

__global__ void __launch_bounds__(128) attn_decode_merge(const float* __restrict__ ws, __nv_bfloat16* __restrict__ out, int splits) {
  const int bh = blockIdx.x, d = threadIdx.x;
  const float* base = ws + static_cast<int64_t>(bh) * splits * (AT_D + 2);
  float M = -INFINITY;
  for (int s = 0; s < splits; ++s) M = fmaxf(M, base[s * (AT_D + 2) + AT_D]);
  float L = 0.f, A = 0.f;
  for (int s = 0; s < splits; ++s) {
    const float ms = base[s * (AT_D + 2) + AT_D];
    const float f = (ms == -INFINITY) ? 0.f : exp2f(ms - M);
    L += base[s * (AT_D + 2) + AT_D + 1] * f;
    A += base[s * (AT_D + 2) + d] * f;
  }
  out[static_cast<int64_t>(bh) * AT_D + d] = __float2bfloat16_rn(L > 0.f ? A / L : 0.f);
}

static int make_tmap_heads(CUtensorMap* tm, const void* ptr, int T, int H, int B, int64_t stride_b, int64_t stride_h) {
  uint64_t dims[4] = {static_cast<uint64_t>(AT_D), static_cast<uint64_t>(T), static_cast<uint64_t>(H), static_cast<uint64_t>(B)};
  uint64_t str[3] = {static_cast<uint64_t>(AT_D) * 2, static_cast<uint64_t>(stride_h) * 2, static_cast<uint64_t>(stride_b) * 2};
  uint32_t box[4] = {64, 128, 1, 1};
  return make_tmap_bf16(tm, ptr, 4, dims, str, box);
}

}  // namespace aria

using namespace aria;

extern "C" int aria_attention_fwd_v2(const void* q, const void* k, const void* v, void* out, const uint8_t* key_mask, int32_t B,
                                  int32_t H, int32_t Tq, int32_t Tk, int64_t q_stride_b, int64_t q_stride_h,
                                  int64_t kv_stride_b, int64_t kv_stride_h, int32_t out_hd, float scale, int32_t causal,
                                  aria_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ARIA_CHECK_ARG(q && k && v && out);
  ARIA_CHECK_ARG(B > 0 && H > 0 && Tq > 0 && Tk > 0 && Tk >= (causal ? Tq : 0));
  ARIA_CHECK_ARG(out_hd > 0 && out_hd <= AT_D && out_hd % 8 == 0);
  ARIA_CHECK_ARG(q_stride_b % 8 == 0 && q_stride_h % 8 == 0 && kv_stride_b % 8 == 0 && kv_stride_h % 8 == 0);
  CUtensorMap tmQ, tmK, tmV;
  int rc = make_tmap_heads(&tmQ, q, Tq, H, B, q_stride_b, q_stride_h);
  if (rc) return rc;
  rc = make_tmap_heads(&tmK, k, Tk, H, B, kv_stride_b, kv_stride_h);
  if (rc) return rc;
  rc = make_tmap_heads(&tmV, v, Tk, H, B, kv_stride_b, kv_stride_h);
  if (rc) return rc;
  AttnParams p{};
  p.B = B;
  p.H = H;
  p.Tq = Tq;
  p.Tk = Tk;
  p.out_hd = out_hd;
  p.hd_eff = (out_hd + 15) / 16 * 16;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.causal = causal;
  p.key_mask = key_mask;
  p.out = static_cast<__nv_bfloat16*>(out);
  static bool attr_set[kMaxDevices] = {};
  if (ensure_dynamic_smem(attr_set, attn_fwd2_kernel, A2_SMEM) != cudaSuccess) return ARIA_ERR_CUDA;
  p.n_q_tiles = (Tq + 2 * AT_BM - 1) / (2 * AT_BM);  // 256-row query pairs
  const int64_t grid = static_cast<int64_t>(B) * H * p.n_q_tiles;
  ARIA_CHECK_ARG(grid < (1ll << 31));
  attn_fwd2_kernel<<<static_cast<int>(grid), A2_THREADS, A2_SMEM, stream>>>(tmQ, tmK, tmV, p);
  return check_launch("attn_fwd2_kernel");
}

extern "C" int64_t aria_attention_decode_workspace_bytes(int32_t B, int32_t H, int32_t Tk) {
  const int64_t splits = (Tk + DEC_SPLIT_KEYS - 1) / DEC_SPLIT_KEYS;
  return static_cast<int64_t>(B) * H * splits * (AT_D + 2) * sizeof(float);
}

extern "C" int aria_attention_decode(const void* q, const void* k, const void* v, void* out, const uint8_t* key_mask, int32_t B, int32_t H, int32_t Tk,
                                     int64_t q_stride_b, int64_t q_stride_h, int64_t kv_stride_b, int64_t kv_stride_h,
                                     float scale, void* workspace, int64_t workspace_bytes, aria_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ARIA_CHECK_ARG(q && k && v && out && workspace && B > 0 && H > 0 && Tk > 0 && q_stride_b % 4 == 0 && q_stride_h % 4 == 0);
  ARIA_CHECK_ARG(workspace_bytes >= aria_attention_decode_workspace_bytes(B, H, Tk));
  const int splits = (Tk + DEC_SPLIT_KEYS - 1) / DEC_SPLIT_KEYS;
  dim3 grid(B * H, splits);
  attn_decode_partial<<<grid, 128, 0, stream>>>(static_cast<const __nv_bfloat16*>(q), static_cast<const __nv_bfloat16*>(k),
                                                static_cast<const __nv_bfloat16*>(v), key_mask, static_cast<float*>(workspace), H, Tk,
                                                q_stride_b, q_stride_h, kv_stride_b, kv_stride_h, scale * 1.4426950408889634f, splits);
  int rc = check_launch("attn_decode_partial");
  if (rc) return rc;
  attn_decode_merge<<<B * H, 128, 0, stream>>>(static_cast<const float*>(workspace), static_cast<__nv_bfloat16*>(out), splits);
  return check_launch("attn_decode_merge");
}
