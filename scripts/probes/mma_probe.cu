// Micro-benchmark of the GEMM main loop (TMA producer thread + tcgen05.mma issuer thread) in several code styles, to find
// what bounds the per-k-block time.  No epilogue work: four warps just recycle the two TMEM accumulators.
//   GUARD : 0 = `if (lane == 0)` single-thread role bodies, 1 = `if (elect_one())`
//   DESC  : 0 = build both shared-memory descriptors from scratch for every MMA, 1 = base descriptor + immediate offsets
//   UNR   : 0 = running stage index / phase bit, 1 = stage loop fully unrolled (static stage, per-stage phase bitmask)
//   KC    : 64-column k-chunks per pipeline stage (one 3-D TMA box per operand per stage)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o mma_probe mma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
#define DEV __device__ __forceinline__

DEV uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
DEV uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile("{ .reg .pred P; elect.sync _|P, 0xffffffff; selp.b32 %0, 1, 0, P; }" : "=r"(pred));
  return pred;
}
DEV void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
DEV void mbar_expect(uint32_t b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory"); }
DEV void mbar_arrive(uint32_t b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b) : "memory"); }
DEV void mbar_wait(uint32_t b, uint32_t parity) {
  uint32_t ok = 0;
  for (uint32_t spin = 0;; ++spin) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(b), "r"(parity) : "memory");
    if (ok) return;
    if (spin > (1u << 26)) { printf("timeout bar %u\n", b); asm volatile("trap;"); }
  }
}
DEV void tma3(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
DEV void umma(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }" ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
DEV void commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
DEV uint64_t make_desc(uint32_t addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

template <int GUARD, int DESC, int UNR, int KC, int AM, int BN, int STAGES>
__global__ void __launch_bounds__(192, 1) probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int tiles_per_cta, int k_blocks) {
  constexpr int A_CHUNK = AM * 128, B_CHUNK = BN * 128;
  constexpr int A_BYTES = A_CHUNK * KC, STAGE_BYTES = A_BYTES + B_CHUNK * KC;
  constexpr int TCOLS = (2 * BN < 32) ? 32 : 2 * BN;
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + (128 - AM) * 128);
  uint64_t* full = bars; uint64_t* empty = bars + STAGES; uint64_t* tfull = empty + STAGES; uint64_t* tempty = tfull + 2;
  uint32_t* slot = reinterpret_cast<uint32_t*>(tempty + 2);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(TCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  const uint32_t s0 = smem_u32(smem), full0 = smem_u32(full), empty0 = smem_u32(empty), tfull0 = smem_u32(tfull), tempty0 = smem_u32(tempty);
  const int k_steps = k_blocks / KC;
  const int arow = blockIdx.x * 128, brow = blockIdx.x * BN;

  if (warp == 0) {
    if (GUARD ? elect_one() : (lane == 0)) {
      if constexpr (UNR) {
        uint32_t ph = 0;
        for (int t = 0; t < tiles_per_cta; ++t)
          for (int kb0 = 0; kb0 < k_steps; kb0 += STAGES) {
#pragma unroll
            for (int s = 0; s < STAGES; ++s) {
              if (kb0 + s >= k_steps) break;
              mbar_wait(empty0 + s * 8, ((ph >> s) & 1) ^ 1);
              ph ^= 1u << s;
              mbar_expect(full0 + s * 8, STAGE_BYTES);
              tma3(s0 + s * STAGE_BYTES, &tmA, full0 + s * 8, 0, arow, (kb0 + s) * KC);
              tma3(s0 + s * STAGE_BYTES + A_BYTES, &tmB, full0 + s * 8, 0, brow, (kb0 + s) * KC);
            }
          }
      } else {
        int st = 0; uint32_t ph = 0;
        for (int t = 0; t < tiles_per_cta; ++t)
          for (int kb = 0; kb < k_steps; ++kb) {
            mbar_wait(empty0 + st * 8, ph ^ 1);
            mbar_expect(full0 + st * 8, STAGE_BYTES);
            tma3(s0 + st * STAGE_BYTES, &tmA, full0 + st * 8, 0, arow, kb * KC);
            tma3(s0 + st * STAGE_BYTES + A_BYTES, &tmB, full0 + st * 8, 0, brow, kb * KC);
            if (++st == STAGES) { st = 0; ph ^= 1; }
          }
      }
    }
  } else if (warp == 1) {
    if (GUARD ? elect_one() : (lane == 0)) {
      constexpr uint32_t idesc = make_idesc(128, BN);
      const uint64_t da0 = make_desc(s0), db0 = make_desc(s0 + A_BYTES);
      if constexpr (UNR) {
        uint32_t ph = 0;
        for (int t = 0; t < tiles_per_cta; ++t) {
          const int as = t & 1;
          mbar_wait(tempty0 + as * 8, ((t >> 1) & 1) ^ 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t d = tmem + as * BN;
          for (int kb0 = 0; kb0 < k_steps; kb0 += STAGES) {
#pragma unroll
            for (int s = 0; s < STAGES; ++s) {
              if (kb0 + s >= k_steps) break;
              mbar_wait(full0 + s * 8, (ph >> s) & 1);
              ph ^= 1u << s;
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
              for (int c = 0; c < KC; ++c)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma(d, da0 + ((s * STAGE_BYTES + c * A_CHUNK + k * 32) >> 4), db0 + ((s * STAGE_BYTES + c * B_CHUNK + k * 32) >> 4), idesc,
                       (kb0 + s + c + k) ? 1u : 0u);
              commit(empty0 + s * 8);
            }
          }
          commit(tfull0 + as * 8);
        }
      } else {
        int st = 0; uint32_t ph = 0;
        for (int t = 0; t < tiles_per_cta; ++t) {
          const int as = t & 1;
          mbar_wait(tempty0 + as * 8, ((t >> 1) & 1) ^ 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t d = tmem + as * BN;
          for (int kb = 0; kb < k_steps; ++kb) {
            mbar_wait(full0 + st * 8, ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t sa = s0 + st * STAGE_BYTES, sb = sa + A_BYTES;
#pragma unroll
            for (int c = 0; c < KC; ++c)
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                if constexpr (DESC) umma(d, da0 + ((st * STAGE_BYTES + c * A_CHUNK + k * 32) >> 4), db0 + ((st * STAGE_BYTES + c * B_CHUNK + k * 32) >> 4), idesc, (kb + c + k) ? 1u : 0u);
                else umma(d, make_desc(sa + c * A_CHUNK + k * 32), make_desc(sb + c * B_CHUNK + k * 32), idesc, (kb + c + k) ? 1u : 0u);
              }
            commit(empty0 + st * 8);
            if (++st == STAGES) { st = 0; ph ^= 1; }
          }
          commit(tfull0 + as * 8);
        }
      }
    }
  } else {
    for (int t = 0; t < tiles_per_cta; ++t) {
      const int as = t & 1;
      mbar_wait(tfull0 + as * 8, (t >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty0 + as * 8);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TCOLS) : "memory");
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn g_encode;
static void* g_a; static void* g_b;
static const int64_t ROWS = 148 * 128, K = 2560;
static cudaEvent_t e0, e1;

static CUtensorMap make_map(void* base, int box_rows, int kc) {
  CUtensorMap tm;
  cuuint64_t dims[3] = {64, static_cast<cuuint64_t>(ROWS), static_cast<cuuint64_t>(K / 64)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(K * 2), 128};
  cuuint32_t box[3] = {64, static_cast<cuuint32_t>(box_rows), static_cast<cuuint32_t>(kc)};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = g_encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", r); exit(1); }
  return tm;
}

template <int GUARD, int DESC, int UNR, int KC, int AM, int BN, int STAGES>
static void run(const char* note) {
  constexpr int STAGE_BYTES = (AM + BN) * 128 * KC;
  constexpr int SMEM = STAGES * STAGE_BYTES + (128 - AM) * 128 + 1024 + 1024;
  static_assert(SMEM <= 227 * 1024, "smem");
  auto kern = probe<GUARD, DESC, UNR, KC, AM, BN, STAGES>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  CUtensorMap ta = make_map(g_a, AM, KC), tb = make_map(g_b, BN, KC);
  const int tiles = 8, kb = 40;
  kern<<<148, 192, SMEM>>>(ta, tb, tiles, kb);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < 5; ++i) kern<<<148, 192, SMEM>>>(ta, tb, tiles, kb);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  CK(cudaGetLastError());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  const double us = ms * 1e3 / 5;
  const double per_kb = us / (tiles * kb);
  const double tf = 2.0 * 128 * BN * 64 / (per_kb * 1e-6) * 148 / 1e12;
  printf("guard=%d desc=%d unroll=%d KC=%d AM=%3d BN=%3d stages=%2d | %8.1f us  %6.3f us/k-block  %7.1f TFLOP/s(MMA shape)  %s\n", GUARD, DESC, UNR, KC, AM, BN,
         STAGES, us, per_kb, tf, note);
}

int main() {
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  g_encode = reinterpret_cast<EncodeFn>(fn);
  CK(cudaMalloc(&g_a, ROWS * K * 2)); CK(cudaMalloc(&g_b, ROWS * K * 2));
  CK(cudaMemset(g_a, 0, ROWS * K * 2)); CK(cudaMemset(g_b, 0, ROWS * K * 2));
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  // 128x128 tiles (32 KB per k-block)
  run<0, 0, 0, 1, 128, 128, 6>("old style");
  run<1, 0, 0, 1, 128, 128, 6>("elect");
  run<1, 1, 0, 1, 128, 128, 6>("elect + base desc");
  run<1, 1, 1, 1, 128, 128, 6>("elect + unrolled");
  run<1, 1, 0, 2, 128, 128, 3>("elect + base desc, 2 k-chunks/stage");
  run<1, 1, 1, 2, 128, 128, 3>("elect + unrolled, 2 k-chunks/stage");
  // 128x256 tiles (48 KB per k-block)
  run<0, 0, 0, 1, 128, 256, 4>("old style");
  run<1, 0, 0, 1, 128, 256, 4>("elect");
  run<1, 1, 1, 1, 128, 256, 4>("elect + unrolled");
  run<1, 1, 1, 2, 128, 256, 2>("elect + unrolled, 2 k-chunks/stage");
  // skinny: 32-row A box, 32-column tiles (8 KB per k-block): pure issue overhead
  run<0, 0, 0, 1, 32, 32, 24>("old style");
  run<1, 0, 0, 1, 32, 32, 24>("elect");
  run<1, 1, 0, 1, 32, 32, 24>("elect + base desc");
  run<1, 1, 1, 1, 32, 32, 24>("elect + unrolled");
  run<1, 1, 1, 1, 32, 32, 8>("elect + unrolled");
  run<1, 1, 0, 4, 32, 32, 6>("elect + base desc, 4 k-chunks/stage");
  run<1, 1, 1, 4, 32, 32, 6>("elect + unrolled, 4 k-chunks/stage");
  run<1, 1, 1, 2, 32, 32, 12>("elect + unrolled, 2 k-chunks/stage");
  run<1, 1, 1, 4, 32, 128, 2>("elect + unrolled, 4 k-chunks/stage");
  run<1, 1, 1, 2, 32, 128, 4>("elect + unrolled, 2 k-chunks/stage");
  return 0;
}
