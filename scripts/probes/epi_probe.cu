// Micro-benchmark for round 2: how fast can the GEMM epilogue drain one 128-row x 256-column fp32 accumulator from TMEM
// (tcgen05.ld -> + bias -> gelu_tanh -> bf16 -> 16-byte global stores), per CTA, as a function of
//   NW    epilogue warps (8 = today: 2 per TMEM lane quadrant, 12, 16)
//   PIPE  0 = ld / wait / process per 32-column chunk (today), 1 = the next chunk's tcgen05.ld is in flight while the
//         current one is processed (double-buffered registers)
// TMEM is first filled with a known pattern through tcgen05.st and the output is verified on the host, so the PIPE variant
// also answers whether ptxas keeps the in-flight destination registers untouched between the ld and the wait.
// The ViT fc1 (+bias +gelu) GEMM is epilogue-bound today: 64 us vs 44 us without the activation (profiles/r01_step_breakdown_f.txt).
// NOT YET RUN (written after the round-1 GPU budget was spent).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o epi_probe epi_probe.cu
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
#define DEV __device__ __forceinline__

DEV uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
DEV void ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
DEV void st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]),
        "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
        "r"(r[30]), "r"(r[31])
      : "memory");
}
DEV void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
DEV void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
DEV float gelu_tanh(float x) {
  float t;
  const float inner = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(inner));
  return 0.5f * x * (1.f + t);
}
DEV uint32_t pack2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
DEV void process_store(const uint32_t (&v)[32], const __nv_bfloat16* bias, __nv_bfloat16* orow, int c) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float x[8];
    const uint4 bv = *reinterpret_cast<const uint4*>(bias + c + q * 8);
    const uint32_t bw[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      x[2 * j] = __uint_as_float(v[q * 8 + 2 * j]) + __uint_as_float(bw[j] << 16);
      x[2 * j + 1] = __uint_as_float(v[q * 8 + 2 * j + 1]) + __uint_as_float(bw[j] & 0xffff0000u);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = gelu_tanh(x[j]);
    *reinterpret_cast<uint4*>(orow + c + q * 8) = make_uint4(pack2(x[0], x[1]), pack2(x[2], x[3]), pack2(x[4], x[5]), pack2(x[6], x[7]));
  }
}

constexpr int COLS = 256;

template <int NW, int PIPE>
__global__ void __launch_bounds__(NW * 32, 1) drain(const __nv_bfloat16* __restrict__ bias, __nv_bfloat16* __restrict__ out, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = slot;
  const int quad = warp & 3, part = warp >> 2;
  constexpr int NPARTS = NW / 4;
  const uint32_t taddr = tmem + (static_cast<uint32_t>(quad * 32) << 16);
  const int row = quad * 32 + lane;
  // fill: accumulator[row][col] = (row * 0.01 - 0.6) + col * 0.002   (written by the part-0 warps)
  if (part == 0) {
    for (int c = 0; c < COLS; c += 32) {
      uint32_t v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __float_as_uint((row * 0.01f - 0.6f) + (c + i) * 0.002f);
      st32(taddr + c, v);
    }
    wait_st();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  constexpr int NCH = COLS / 32;
  const int c_begin = (part * NCH / NPARTS) * 32, c_end = ((part + 1) * NCH / NPARTS) * 32;
  __nv_bfloat16* orow = out + (static_cast<size_t>(blockIdx.x) * 128 + row) * COLS;
  for (int it = 0; it < iters; ++it) {
    if constexpr (PIPE == 0) {
#pragma unroll 1
      for (int c = c_begin; c < c_end; c += 32) {
        uint32_t v[32];
        ld32(taddr + c, v);
        wait_ld();
        process_store(v, bias, orow, c);
      }
    } else {
      uint32_t v0[32], v1[32];
      ld32(taddr + c_begin, v0);
#pragma unroll 1
      for (int c = c_begin; c < c_end; c += 64) {
        wait_ld();                                   // v0 ready
        if (c + 32 < c_end) ld32(taddr + c + 32, v1);  // in flight while v0 is processed
        process_store(v0, bias, orow, c);
        if (c + 32 < c_end) {
          wait_ld();                                 // v1 ready
          if (c + 64 < c_end) ld32(taddr + c + 64, v0);
          process_store(v1, bias, orow, c + 32);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(COLS) : "memory");
}

static float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }

template <int NW, int PIPE>
static void run() {
  const int ctas = 148, iters = 50;
  __nv_bfloat16 *bias, *out;
  CK(cudaMalloc(&bias, COLS * 2));
  CK(cudaMalloc(&out, static_cast<size_t>(ctas) * 128 * COLS * 2));
  std::vector<__nv_bfloat16> hb(COLS);
  for (int i = 0; i < COLS; ++i) hb[i] = __float2bfloat16(0.01f * (i % 17) - 0.08f);
  CK(cudaMemcpy(bias, hb.data(), COLS * 2, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  drain<NW, PIPE><<<ctas, NW * 32>>>(bias, out, 2);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  drain<NW, PIPE><<<ctas, NW * 32>>>(bias, out, iters);
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  CK(cudaGetLastError());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<__nv_bfloat16> ho(static_cast<size_t>(ctas) * 128 * COLS);
  CK(cudaMemcpy(ho.data(), out, ho.size() * 2, cudaMemcpyDeviceToHost));
  double max_err = 0;
  for (int cta : {0, 73, 147})
    for (int r = 0; r < 128; ++r)
      for (int c = 0; c < COLS; ++c) {
        const float x = (r * 0.01f - 0.6f) + c * 0.002f + __bfloat162float(hb[c]);
        const float want = bf16_round(0.5f * x * (1.f + tanhf(0.7978845608028654f * (x + 0.044715f * x * x * x))));
        const float got = __bfloat162float(ho[(static_cast<size_t>(cta) * 128 + r) * COLS + c]);
        max_err = fmax(max_err, fabs(got - want));
      }
  printf("warps=%2d pipelined=%d : %6.2f us per 128x256 drain per CTA   (max abs err vs host %.4f %s)\n", NW, PIPE, ms * 1e3 / iters, max_err,
         max_err < 0.01 ? "OK" : "MISMATCH");
  CK(cudaFree(bias)); CK(cudaFree(out));
}

int main() {
  run<8, 0>();
  run<8, 1>();
  run<12, 0>();
  run<16, 0>();
  run<16, 1>();
  return 0;
}
