// Micro-benchmark for round 2: how much of the per-kernel fixed cost (launch gap + prologue) does programmatic dependent
// launch (PDL) hide for kernels shaped like ours - a prologue that touches no global memory (mbarrier init, TMEM allocation),
// then a short body that depends on the previous kernel's output?  Chains of N launches, four ways:
//   stream        plain <<<>>> launches on one stream
//   graph         the same chain captured in a CUDA graph
//   stream+PDL    cudaLaunchKernelEx with cudaLaunchAttributeProgrammaticStreamSerialization; the kernel runs its prologue,
//                 then `griddepcontrol.wait`, then the body; `griddepcontrol.launch_dependents` right after the prologue
//   graph+PDL     the PDL chain captured in a graph
// Variants: with / without a 512-column TMEM allocation in the prologue (a dependent CTA cannot get TMEM while the previous
// kernel's CTA on that SM still holds it - this measures how much that costs), body length ~2 us or ~10 us.
// NOT YET RUN (written after the round-1 GPU budget was spent).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o pdl_probe pdl_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <bool TMEM, bool PDL>
__global__ void __launch_bounds__(320, 1) link(const float* __restrict__ in, float* __restrict__ out, int n, int spin) {
  __shared__ uint64_t bars[16];
  __shared__ uint32_t slot;
  // ---- prologue: no global memory traffic
  if (threadIdx.x == 0)
    for (int i = 0; i < 16; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[i])), "r"(1));
  uint32_t tmem = 0;
  if (TMEM && threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();
  if (TMEM) tmem = slot;
  if (PDL) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // let the next kernel start its prologue
    asm volatile("griddepcontrol.wait;" ::: "memory");               // previous kernel's writes are visible after this
  }
  // ---- body: depends on the previous kernel's output
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float v = i < n ? in[i] : 0.f;
  const long long t0 = clock64();
  while (clock64() - t0 < spin) v = v * 1.0000001f + 1e-9f;
  if (i < n) out[i] = v;
  __syncthreads();
  if (TMEM && threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <bool TMEM, bool PDL>
static void launch_chain(cudaStream_t s, float* a, float* b, int n, int spin, int links) {
  for (int l = 0; l < links; ++l) {
    float* in = (l & 1) ? b : a;
    float* out = (l & 1) ? a : b;
    if (PDL) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(148);
      cfg.blockDim = dim3(320);
      cfg.stream = s;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      CK(cudaLaunchKernelEx(&cfg, link<TMEM, PDL>, (const float*)in, out, n, spin));
    } else {
      link<TMEM, PDL><<<148, 320, 0, s>>>(in, out, n, spin);
    }
  }
}

template <bool TMEM, bool PDL>
static void run(const char* name, int spin, bool graph) {
  const int n = 148 * 320, links = 200;
  float *a, *b;
  CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4));
  CK(cudaMemset(a, 0, n * 4)); CK(cudaMemset(b, 0, n * 4));
  cudaStream_t s; CK(cudaStreamCreate(&s));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  cudaGraphExec_t exec = nullptr;
  if (graph) {
    cudaGraph_t g;
    CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    launch_chain<TMEM, PDL>(s, a, b, n, spin, links);
    CK(cudaStreamEndCapture(s, &g));
    CK(cudaGraphInstantiate(&exec, g, 0));
  }
  for (int rep = 0; rep < 3; ++rep) {  // warm-up
    if (graph) CK(cudaGraphLaunch(exec, s)); else launch_chain<TMEM, PDL>(s, a, b, n, spin, links);
  }
  CK(cudaStreamSynchronize(s));
  CK(cudaEventRecord(e0, s));
  const int reps = 5;
  for (int rep = 0; rep < reps; ++rep) {
    if (graph) CK(cudaGraphLaunch(exec, s)); else launch_chain<TMEM, PDL>(s, a, b, n, spin, links);
  }
  CK(cudaEventRecord(e1, s));
  CK(cudaEventSynchronize(e1));
  CK(cudaGetLastError());
  float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
  printf("%-12s tmem=%d body~%5d clk %s : %7.2f us per kernel\n", name, TMEM ? 1 : 0, spin, graph ? "graph " : "stream", ms * 1e3 / (reps * links));
  CK(cudaFree(a)); CK(cudaFree(b));
}

int main() {
  for (int spin : {4000, 20000}) {  // ~2 us and ~10 us bodies at 1.9 GHz
    run<false, false>("plain", spin, false);
    run<false, false>("plain", spin, true);
    run<false, true>("PDL", spin, false);
    run<false, true>("PDL", spin, true);
    run<true, false>("plain", spin, false);
    run<true, false>("plain", spin, true);
    run<true, true>("PDL", spin, false);
    run<true, true>("PDL", spin, true);
  }
  return 0;
}
