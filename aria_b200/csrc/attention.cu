// Attention for sm_100a.
//
// aria_attention_fwd: flash-style forward with both contractions on tcgen05 tensor cores:
//     S = Q K^T   (A = Q tile, B = K tile, both K-major SW128 in shared memory, accumulator S in TMEM)
//     O += P V    (A = P tile written by the softmax warps into SW128 shared memory, B = V tile consumed
//                  MN-major — V is [keys, d] with d contiguous, exactly the HF cache layout — O in TMEM)
//   warp 0: TMA producer (Q once, K/V double buffered); warp 1: MMA issuer + TMEM owner;
//   warps 2-5: softmax (one query row per thread: tcgen05.ld of S, fp32 online softmax with warp-uniform
//   lazy rescale of O, P -> shared memory) and the final normalise + store.
// Replaces flash_attn_func / SDPA behind LLAMA_ATTENTION_CLASSES (aria/model/moe_lm.py:594) and the
// Idefics2 / nn.MultiheadAttention attention of the ViT + projector (vision_encoder.py:120,
// projector.py:93) with head_dim padded 72 -> 128.
//
// aria_attention_decode: single-query attention against the KV cache; HBM-bound, CUDA cores, split-KV.
#include "common.cuh"
#include "ptx.cuh"

namespace aria {

constexpr int AT_BM = 128;   // queries per CTA
constexpr int AT_BN = 128;   // keys per step
constexpr int AT_D = 128;    // head dim
constexpr int AT_TILE = AT_BM * AT_D * 2;  // 32 KB
constexpr int AT_HALF = AT_TILE / 2;       // one SW128 atom column: [128 rows][64 bf16]
constexpr int AT_THREADS = 192;
constexpr int AT_SMEM = AT_TILE /*Q*/ + 2 * AT_TILE /*K*/ + 2 * AT_TILE /*V*/ + AT_TILE /*P*/ + 1024 + 256;

struct AttnParams {
  int B, H, Tq, Tk;
  int out_hd;
  float scale_log2;
  int causal;
  const uint8_t* key_mask;  // [B, Tk] 1 = masked out
  __nv_bfloat16* out;       // [B, Tq, H*out_hd]
  int n_q_tiles;
};

__global__ void __launch_bounds__(AT_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + AT_TILE;
  uint8_t* sV = sK + 2 * AT_TILE;
  uint8_t* sP = sV + 2 * AT_TILE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sP + AT_TILE);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* v_full = bars + 3;    // [2]
  uint64_t* kv_empty = bars + 5;  // [2]
  uint64_t* s_full = bars + 7;
  uint64_t* p_full = bars + 8;
  uint64_t* o_full = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // heavy (late) causal query tiles first
  const int bh = blockIdx.x % (p.B * p.H);
  const int q_tile = p.n_q_tiles - 1 - blockIdx.x / (p.B * p.H);
  const int b = bh / p.H, h = bh % p.H;
  const int q0 = q_tile * AT_BM;
  const int pos_off = p.Tk - p.Tq;  // absolute position of query 0
  int n_kv = (p.Tk + AT_BN - 1) / AT_BN;
  if (p.causal) {
    int last = pos_off + min(q0 + AT_BM, p.Tq) - 1;  // last visible key of this tile
    int lim = last / AT_BN + 1;
    if (lim < n_kv) n_kv = lim;
  }

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tmQ);
    prefetch_tmap(&tmK);
    prefetch_tmap(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base;        // columns [0,128)
  const uint32_t tO = tmem_base + 128;  // columns [128,256)

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, AT_TILE);
      tma_load_4d(sQ, &tmQ, q_full, 0, q0, h, b);
      tma_load_4d(sQ + AT_HALF, &tmQ, q_full, 64, q0, h, b);
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
    if (lane == 0) {
      constexpr uint32_t idesc_qk = make_idesc_bf16(AT_BM, AT_BN, false, false);
      constexpr uint32_t idesc_pv = make_idesc_bf16(AT_BM, AT_D, false, true);
      const uint32_t aQ = smem_u32(sQ), aP = smem_u32(sP);
      auto issue_qk = [&](int j) {
        const int s = j & 1;
        mbar_wait(&k_full[s], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t aK = smem_u32(sK + s * AT_TILE);
#pragma unroll
        for (int k = 0; k < AT_D / 16; ++k) {
          const uint32_t off = (k >> 2) * AT_HALF + (k & 3) * 32;
          umma_bf16_ss(tS, make_smem_desc(aQ + off, 16, 1024), make_smem_desc(aK + off, 16, 1024), idesc_qk, k ? 1u : 0u);
        }
        umma_commit(s_full);
      };
      mbar_wait(q_full, 0);
      issue_qk(0);
      for (int j = 0; j < n_kv; ++j) {
        const int s = j & 1;
        mbar_wait(&v_full[s], (j >> 1) & 1);
        mbar_wait(p_full, j & 1);  // P(j) in smem, O rescaled, S(j) consumed
        tc_fence_after();
        const uint32_t aV = smem_u32(sV + s * AT_TILE);
#pragma unroll
        for (int k = 0; k < AT_BN / 16; ++k) {
          // A = P: K-major, keys k*16.. -> atom (k>>2), +32 B per step; B = V: MN-major, 16 key rows = 2048 B
          const uint64_t da = make_smem_desc(aP + (k >> 2) * AT_HALF + (k & 3) * 32, 16, 1024);
          const uint64_t db = make_smem_desc(aV + k * 2048, AT_HALF, 1024);
          umma_bf16_ss(tO, da, db, idesc_pv, (j | k) ? 1u : 0u);
        }
        umma_commit(&kv_empty[s]);
        umma_commit(o_full);
        if (j + 1 < n_kv) issue_qk(j + 1);
      }
    }
  } else {
    // ------------------------------ softmax / correction / epilogue: one query row per thread
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const int q = q0 + r;
    const bool row_ok = q < p.Tq;
    const int qpos = pos_off + q;
    const uint32_t lane_addr = static_cast<uint32_t>(quad * 32) << 16;
    const uint8_t* km = p.key_mask ? p.key_mask + static_cast<int64_t>(b) * p.Tk : nullptr;
    float m_ref = -INFINITY, l = 0.f;
    uint8_t* prow = sP + r * 128;

    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int k0 = j * AT_BN;
      const bool need_mask = (k0 + AT_BN > p.Tk) || (p.causal && (k0 + AT_BN - 1 > pos_off + q0)) || km != nullptr;
      // pass 1: row max
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < AT_BN; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tS + lane_addr + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float s = __uint_as_float(v[i]);
          if (need_mask) {
            const int kc = k0 + c + i;
            const bool dead = kc >= p.Tk || (p.causal && kc > qpos) || (km && kc < p.Tk && km[kc]);
            if (dead) s = -INFINITY;
          }
          mx = fmaxf(mx, s);
        }
      }
      const float m_new = fmaxf(m_ref, mx * p.scale_log2);
      // lazy rescale, warp-uniform decision (tcgen05.ld/st are warp-collective)
      const bool want = (m_new - m_ref > 8.0f) || (m_ref == -INFINITY && m_new > -INFINITY);
      const bool do_rescale = __any_sync(0xffffffffu, want);
      if (j > 0) {
        mbar_wait(o_full, (j - 1) & 1);  // PV(j-1) done: O valid, P smem free
        tc_fence_after();
      }
      if (do_rescale) {
        const float f = (m_ref == -INFINITY) ? 0.f : exp2f(m_ref - m_new);
        l *= f;
        m_ref = m_new;
        if (j > 0) {
#pragma unroll 1
          for (int c = 0; c < AT_D; c += 32) {
            uint32_t v[32];
            tmem_ld_32x32(tO + lane_addr + c, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * f);
            tmem_st_32x32(tO + lane_addr + c, v);
          }
          tmem_st_wait();
        }
      }
      const float mref_safe = (m_ref == -INFINITY) ? 0.f : m_ref;
      // pass 2: p = exp2(s*scale - m_ref) -> bf16 -> swizzled smem (K-major SW128: 16B chunk index ^= row & 7)
      float lsum = 0.f;
#pragma unroll 1
      for (int c = 0; c < AT_BN; c += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tS + lane_addr + c, v);
        tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float s0 = __uint_as_float(v[i]), s1 = __uint_as_float(v[i + 1]);
          if (need_mask) {
            const int kc = k0 + c + i;
            const bool d0 = kc >= p.Tk || (p.causal && kc > qpos) || (km && kc < p.Tk && km[kc]);
            const bool d1 = kc + 1 >= p.Tk || (p.causal && kc + 1 > qpos) || (km && kc + 1 < p.Tk && km[kc + 1]);
            if (d0) s0 = -INFINITY;
            if (d1) s1 = -INFINITY;
          }
          const float p0 = exp2f(s0 * p.scale_log2 - mref_safe);
          const float p1 = exp2f(s1 * p.scale_log2 - mref_safe);
          const uint32_t u = pack_bf16(p0, p1);
          lsum += bf16_lo(u) + bf16_hi(u);  // sum what the tensor core will actually multiply
          pk[i >> 1] = u;
        }
        uint8_t* atom = prow + (c >> 6) * AT_HALF;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const int chunk = ((c & 63) >> 3) + ch;
          *reinterpret_cast<uint4*>(atom + ((chunk ^ (r & 7)) << 4)) =
              make_uint4(pk[ch * 4], pk[ch * 4 + 1], pk[ch * 4 + 2], pk[ch * 4 + 3]);
        }
      }
      l += lsum;
      fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc_fence_before();
      mbar_arrive(p_full);
    }
    // ------------------------------ epilogue: O / l -> bf16 -> out[b, q, h*out_hd + d]
    mbar_wait(o_full, (n_kv - 1) & 1);
    tc_fence_after();
    const float inv_l = l > 0.f ? 1.0f / l : 0.f;
    __nv_bfloat16* orow = p.out + (static_cast<int64_t>(b) * p.Tq + q) * (static_cast<int64_t>(p.H) * p.out_hd) + h * p.out_hd;
#pragma unroll 1
    for (int c = 0; c < AT_D; c += 32) {
      uint32_t v[32];
      tmem_ld_32x32(tO + lane_addr + c, v);
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
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

// ------------------------------------------------------------------------------------------------
// Decode: one query per (b,h); block = 4 warps over a contiguous chunk of keys; each lane owns 4 dims.
// Partial (m, l, acc[128]) per (b,h,split) -> workspace; a second kernel merges the splits.
constexpr int DEC_SPLIT_KEYS = 256;

__global__ void __launch_bounds__(128) attn_decode_partial(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ kc,
                                                           const __nv_bfloat16* __restrict__ vc, float* __restrict__ ws, int H, int Tk,
                                                           int64_t kv_stride_b, int64_t kv_stride_h, float scale_log2, int splits) {
  const int bh = blockIdx.x, split = blockIdx.y;
  const int b = bh / H, h = bh % H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k_begin = split * DEC_SPLIT_KEYS, k_end = min(Tk, k_begin + DEC_SPLIT_KEYS);
  const __nv_bfloat16* kbase = kc + b * kv_stride_b + h * kv_stride_h;
  const __nv_bfloat16* vbase = vc + b * kv_stride_b + h * kv_stride_h;
  const uint2 qv = *reinterpret_cast<const uint2*>(q + static_cast<int64_t>(bh) * AT_D + lane * 4);
  const float q0 = bf16_lo(qv.x) * scale_log2, q1 = bf16_hi(qv.x) * scale_log2, q2 = bf16_lo(qv.y) * scale_log2,
              q3 = bf16_hi(qv.y) * scale_log2;
  float m = -INFINITY, l = 0.f, a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (int k0 = k_begin + warp * 4; k0 < k_end; k0 += 16) {
    float s[4];
    uint2 vv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int kk = k0 + u;
      if (kk < k_end) {
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
      if (k0 + u < k_end) {
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

extern "C" int aria_attention_fwd(const void* q, const void* k, const void* v, void* out, const uint8_t* key_mask, int32_t B,
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
  p.scale_log2 = scale * 1.4426950408889634f;
  p.causal = causal;
  p.key_mask = key_mask;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.n_q_tiles = (Tq + AT_BM - 1) / AT_BM;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM) != cudaSuccess)
      return ARIA_ERR_CUDA;
    attr_set = true;
  }
  const int64_t grid = static_cast<int64_t>(B) * H * p.n_q_tiles;
  ARIA_CHECK_ARG(grid < (1ll << 31));
  attn_fwd_kernel<<<static_cast<int>(grid), AT_THREADS, AT_SMEM, stream>>>(tmQ, tmK, tmV, p);
  return check_launch("attn_fwd_kernel");
}

extern "C" int64_t aria_attention_decode_workspace_bytes(int32_t B, int32_t H, int32_t Tk) {
  const int64_t splits = (Tk + DEC_SPLIT_KEYS - 1) / DEC_SPLIT_KEYS;
  return static_cast<int64_t>(B) * H * splits * (AT_D + 2) * sizeof(float);
}

extern "C" int aria_attention_decode(const void* q, const void* k, const void* v, void* out, int32_t B, int32_t H, int32_t Tk,
                                     int64_t kv_stride_b, int64_t kv_stride_h, float scale, void* workspace,
                                     int64_t workspace_bytes, aria_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  ARIA_CHECK_ARG(q && k && v && out && workspace && B > 0 && H > 0 && Tk > 0);
  ARIA_CHECK_ARG(workspace_bytes >= aria_attention_decode_workspace_bytes(B, H, Tk));
  const int splits = (Tk + DEC_SPLIT_KEYS - 1) / DEC_SPLIT_KEYS;
  dim3 grid(B * H, splits);
  attn_decode_partial<<<grid, 128, 0, stream>>>(static_cast<const __nv_bfloat16*>(q), static_cast<const __nv_bfloat16*>(k),
                                                static_cast<const __nv_bfloat16*>(v), static_cast<float*>(workspace), H, Tk,
                                                kv_stride_b, kv_stride_h, scale * 1.4426950408889634f, splits);
  int rc = check_launch("attn_decode_partial");
  if (rc) return rc;
  attn_decode_merge<<<B * H, 128, 0, stream>>>(static_cast<const float*>(workspace), static_cast<__nv_bfloat16*>(out), splits);
  return check_launch("attn_decode_merge");
}
