// Micro-benchmark: what bounds a single producer thread streaming SWIZZLE_128B TMA boxes into shared memory?
// Sweeps box rows R, k-chunks per box C (3-D tensor map: one instruction loads C 64-column chunks), pipeline depth S and
// grid size G, over a DRAM-sized matrix and an L2-resident one.  Prints GB/s per SM and ns per box.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_probe tma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0;
  long long t0 = clock64();
  while (!ok) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    if (!ok && clock64() - t0 > 2000000000LL) { printf("timeout\n"); asm volatile("trap;"); }
  }
}
__device__ __forceinline__ void tma3(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// each CTA walks row tiles t = blockIdx.x, +gridDim.x, ... < n_tiles; per tile K/64/C boxes of [C][R][64] bf16
__global__ void __launch_bounds__(64, 1) probe(const __grid_constant__ CUtensorMap tm, int R, int C, int S, int n_tiles, int kchunks, int boxes_per_stage) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  const int box_bytes = R * 128 * C;
  const int stage_bytes = box_bytes * boxes_per_stage;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S * stage_bytes);
  uint64_t* empty = full + S;
  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int steps = kchunks / (C * boxes_per_stage);
  if (threadIdx.x == 0) {
    int st = 0; uint32_t ph = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x)
      for (int s = 0; s < steps; ++s) {
        mbar_wait(&empty[st], ph ^ 1);
        mbar_expect(&full[st], stage_bytes);
        for (int b = 0; b < boxes_per_stage; ++b)
          tma3(smem + st * stage_bytes + b * box_bytes, &tm, &full[st], 0, t * R, (s * boxes_per_stage + b) * C);
        if (++st == S) { st = 0; ph ^= 1; }
      }
  } else if (threadIdx.x == 32) {
    int st = 0; uint32_t ph = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x)
      for (int s = 0; s < steps; ++s) {
        mbar_wait(&full[st], ph);
        mbar_arrive(&empty[st]);
        if (++st == S) { st = 0; ph ^= 1; }
      }
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int64_t ROWS = 100352, K = 2560;
  void* d;
  CK(cudaMalloc(&d, ROWS * K * 2));
  CK(cudaMemset(d, 1, ROWS * K * 2));
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeFn encode = reinterpret_cast<EncodeFn>(fn);
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  void* flush; CK(cudaMalloc(&flush, 256 << 20));
  printf("%-6s %4s %2s %3s %3s %4s | %9s %9s %10s %9s\n", "src", "R", "C", "bps", "S", "G", "us", "GB/s", "GB/s/SM", "ns/box");
  const int Rs[] = {32, 64, 128, 256};
  const int Cs[] = {1, 2, 4};
  for (int l2 = 0; l2 < 2; ++l2)
    for (int G : {20, 80, 148})
      for (int R : Rs)
        for (int C : Cs)
          for (int bps : {1, 2})
            for (int S : {4, 8, 24}) {
              const int box_bytes = R * 128 * C;
              if (static_cast<int64_t>(box_bytes) * bps * S > 200 * 1024) continue;
              if (S == 24 && box_bytes * bps > 8192) continue;
              if (S == 4 && box_bytes * bps < 16384) continue;
              const int64_t rows_used = l2 ? 8192 : ROWS;  // 8192 rows x 5 KB = 42 MB: L2 resident
              int n_tiles = static_cast<int>(rows_used / R);
              const int reps = l2 ? 6 : 1;
              // cap the work so each run is ~50-200 us
              CUtensorMap tm;
              cuuint64_t dims[3] = {64, static_cast<cuuint64_t>(rows_used), static_cast<cuuint64_t>(K / 64)};
              cuuint64_t strides[2] = {static_cast<cuuint64_t>(K * 2), 128};
              cuuint32_t box[3] = {64, static_cast<cuuint32_t>(R), static_cast<cuuint32_t>(C)};
              cuuint32_t es[3] = {1, 1, 1};
              CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
              if (r != CUDA_SUCCESS) { printf("encode failed %d (R=%d C=%d)\n", r, R, C); continue; }
              const size_t smem = static_cast<size_t>(box_bytes) * bps * S + 1024 + 2 * S * 8 + 64;
              if (l2) { probe<<<G, 64, smem>>>(tm, R, C, S, n_tiles, static_cast<int>(K / 64), bps); }  // warm L2
              else { CK(cudaMemsetAsync(flush, 0, 256 << 20)); }
              CK(cudaEventRecord(e0));
              for (int i = 0; i < reps; ++i) probe<<<G, 64, smem>>>(tm, R, C, S, n_tiles, static_cast<int>(K / 64), bps);
              CK(cudaEventRecord(e1));
              CK(cudaEventSynchronize(e1));
              CK(cudaGetLastError());
              float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
              const double us = ms * 1e3 / reps;
              const double bytes = static_cast<double>(rows_used) * K * 2;
              const double boxes_per_cta = static_cast<double>(n_tiles) / G * (K / 64 / C);
              printf("%-6s %4d %2d %3d %3d %4d | %9.1f %9.0f %10.1f %9.0f\n", l2 ? "L2" : "DRAM", R, C, bps, S, G, us, bytes / us / 1e3, bytes / us / 1e3 / G,
                     us * 1e3 / boxes_per_cta);
            }
  return 0;
}
