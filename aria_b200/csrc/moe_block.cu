// aria_moe_block_fwd: the whole MoELayer.forward (aria/model/moe_lm.py:548-577) behind ONE C-ABI call.
//
//   router GEMM + top-k + softmax + histogram (moe_lm.py:190-201, 261-269)  -> aria_router_topk
//   stable counting sort by expert            (moe_lm.py:313-334)           -> aria_build_permutation, aria_permute_rows
//   fc1 grouped GEMM + glu, fc2 grouped GEMM  (moe_lm.py:467-525)           -> aria_gemm x 2 (device-resident offsets: no
//                                                                              tokens_per_expert.cpu() sync, moe_lm.py:478)
//   shared experts                            (moe_lm.py:368-395)           -> aria_gemm x 2 on `side_stream` when given: the
//                                                                              branch is independent until the final add
//   unpermute + score-weighted sum + `+= shared` (moe_lm.py:336-365, 575-576) -> aria_unpermute_combine
//
// Host-side driver only (SURVEY.md §8b "moe_block_fwd (fused driver)"): it owns the launch ORDER, the workspace carve-up and
// the fork/join of the shared-expert branch, so that the reference-side binding is one call per layer instead of nine; the
// kernels are the same ones the individual entries launch.  No allocation, no host sync, graph-capturable.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/aria_b200.h"
#include "common.cuh"

namespace {

constexpr int64_t kAlign = 256;
inline int64_t al(int64_t n) { return (n + kAlign - 1) / kAlign * kAlign; }

struct BlockWs {
  int64_t idx, scores, counts, offsets, dest, src, logits, permuted, h, y, hs, shared, total;
};

BlockWs carve(int64_t T, int32_t d, int32_t E, int32_t k, int32_t I, int32_t Is) {
  BlockWs w{};
  int64_t o = 0;
  const int64_t R = T * k;
  w.idx = o;      o += al(R * 4);
  w.scores = o;   o += al(R * 2);
  w.counts = o;   o += al(static_cast<int64_t>(E) * 4);
  w.offsets = o;  o += al(static_cast<int64_t>(E + 1) * 4);
  w.dest = o;     o += al(R * 4);
  w.src = o;      o += al(R * 4);
  w.logits = o;   o += al(T * E * 2);
  w.permuted = o; o += al(R * d * 2);
  w.h = o;        o += al(R * I * 2);
  w.y = o;        o += al(R * d * 2);
  w.hs = o;       o += al(T * Is * 2);
  w.shared = o;   o += al(T * d * 2);
  w.total = o;
  return w;
}

// fork / join events of the shared-expert branch, one pair per device (created once; recording an event is legal during
// stream capture and becomes a dependency edge of the graph)
cudaEvent_t* branch_events() {
  static cudaEvent_t ev[aria::kMaxDevices][2] = {};
  static bool made[aria::kMaxDevices] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= aria::kMaxDevices) return nullptr;
  if (!made[dev]) {
    if (cudaEventCreateWithFlags(&ev[dev][0], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&ev[dev][1], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    made[dev] = true;
  }
  return ev[dev];
}

}  // namespace

extern "C" int64_t aria_moe_block_fwd_workspace_bytes(int64_t T, int32_t d, int32_t E, int32_t k, int32_t I, int32_t I_shared) {
  if (T <= 0 || d <= 0 || E <= 0 || k <= 0 || I <= 0 || I_shared < 0) return ARIA_ERR_BAD_ARG;
  return carve(T, d, E, k, I, I_shared).total;
}

extern "C" int aria_moe_block_fwd(const void* x, const void* w_router, const void* fc1_w, const void* fc2_w, const void* gate_w,
                                  const void* up_w, const void* down_w, void* out, int64_t T, int32_t d, int32_t E, int32_t k,
                                  int32_t I, int32_t I_shared, const int32_t* forced_top_idx, void* workspace,
                                  int64_t workspace_bytes, aria_stream_t stream, aria_stream_t side_stream) {
  if (!x || !w_router || !fc1_w || !fc2_w || !out || !workspace) return ARIA_ERR_BAD_ARG;
  if (T <= 0 || d <= 0 || E <= 0 || E > 64 || k <= 0 || k > 8 || k > E || I <= 0 || I_shared < 0) return ARIA_ERR_BAD_ARG;
  if (I_shared > 0 && (!gate_w || !up_w || !down_w)) return ARIA_ERR_BAD_ARG;
  if ((reinterpret_cast<uintptr_t>(workspace) & 15) != 0) return ARIA_ERR_BAD_ARG;
  const BlockWs ws = carve(T, d, E, k, I, I_shared);
  if (workspace_bytes < ws.total) return ARIA_ERR_BAD_ARG;
  uint8_t* base = static_cast<uint8_t*>(workspace);
  auto at = [&](int64_t off) { return static_cast<void*>(base + off); };
  int32_t* idx = static_cast<int32_t*>(at(ws.idx));
  int32_t* counts = static_cast<int32_t*>(at(ws.counts));
  int32_t* offsets = static_cast<int32_t*>(at(ws.offsets));
  int32_t* dest = static_cast<int32_t*>(at(ws.dest));
  int32_t* src = static_cast<int32_t*>(at(ws.src));
  const int64_t R = T * k;
  cudaStream_t s_main = reinterpret_cast<cudaStream_t>(stream);
  cudaStream_t s_side = reinterpret_cast<cudaStream_t>(side_stream);
  int rc;

  // ---- shared experts: an independent branch (its own stream when the caller provides one)
  const bool fork = I_shared > 0 && s_side != nullptr && s_side != s_main;
  cudaEvent_t* ev = nullptr;
  auto shared_branch = [&](aria_stream_t st) -> int {
    aria_gemm_desc_t g;
    memset(&g, 0, sizeof(g));
    g.a = x; g.lda = d; g.m = T; g.n = I_shared; g.k = d;
    g.b[0] = gate_w; g.b[1] = up_w; g.n_seg = 2; g.b_layout = ARIA_B_NK; g.num_groups = 1;
    g.epilogue = ARIA_EPI_SWIGLU;
    g.out[0] = at(ws.hs); g.ldo = I_shared;
    int r = aria_gemm(&g, st);
    if (r) return r;
    memset(&g, 0, sizeof(g));
    g.a = at(ws.hs); g.lda = I_shared; g.m = T; g.n = d; g.k = I_shared;
    g.b[0] = down_w; g.n_seg = 1; g.b_layout = ARIA_B_NK; g.num_groups = 1;
    g.epilogue = ARIA_EPI_LINEAR;
    g.out[0] = at(ws.shared); g.ldo = d;
    return aria_gemm(&g, st);
  };
  if (fork) {
    ev = branch_events();
    if (!ev) return ARIA_ERR_CUDA;
    if (cudaEventRecord(ev[0], s_main) != cudaSuccess) return ARIA_ERR_CUDA;
    if (cudaStreamWaitEvent(s_side, ev[0], 0) != cudaSuccess) return ARIA_ERR_CUDA;
    if ((rc = shared_branch(side_stream))) return rc;
    if (cudaEventRecord(ev[1], s_side) != cudaSuccess) return ARIA_ERR_CUDA;
  }

  // ---- routed experts
  if (forced_top_idx) {  // parity / replay hook: expert choice given, scores = softmax over the logits at those ids
    if ((rc = aria_router_topk(x, w_router, at(ws.logits), idx, at(ws.scores), counts, T, d, E, k, stream))) return rc;
    if ((rc = aria_route_given_indices(at(ws.logits), forced_top_idx, at(ws.scores), counts, T, E, k, stream))) return rc;
    idx = const_cast<int32_t*>(forced_top_idx);
  } else {
    // (aria_router_topk always materialises the bf16 logits: top-k runs on the ROUNDED values, moe_lm.py:200,261)
    if ((rc = aria_router_topk(x, w_router, at(ws.logits), idx, at(ws.scores), counts, T, d, E, k, stream))) return rc;
  }
  if ((rc = aria_build_permutation(idx, counts, offsets, dest, src, T, E, k, 1, stream))) return rc;
  if ((rc = aria_permute_rows(x, src, at(ws.permuted), R, d, stream))) return rc;
  {
    aria_gemm_desc_t g;
    memset(&g, 0, sizeof(g));
    g.a = at(ws.permuted); g.lda = d; g.m = R; g.n = I; g.k = d;
    g.b[0] = fc1_w; g.n_seg = 1; g.b_layout = ARIA_B_GKN; g.num_groups = E; g.group_offsets = offsets;
    g.epilogue = ARIA_EPI_SWIGLU;
    g.out[0] = at(ws.h); g.ldo = I;
    if ((rc = aria_gemm(&g, stream))) return rc;
    memset(&g, 0, sizeof(g));
    g.a = at(ws.h); g.lda = I; g.m = R; g.n = d; g.k = I;
    g.b[0] = fc2_w; g.n_seg = 1; g.b_layout = ARIA_B_GKN; g.num_groups = E; g.group_offsets = offsets;
    g.epilogue = ARIA_EPI_LINEAR;
    g.out[0] = at(ws.y); g.ldo = d;
    if ((rc = aria_gemm(&g, stream))) return rc;
  }

  // ---- join + combine (+ shared)
  const void* shared = nullptr;
  if (I_shared > 0) {
    if (fork) {
      if (cudaStreamWaitEvent(s_main, ev[1], 0) != cudaSuccess) return ARIA_ERR_CUDA;
    } else if ((rc = shared_branch(stream))) {
      return rc;
    }
    shared = at(ws.shared);
  }
  return aria_unpermute_combine(at(ws.y), dest, at(ws.scores), shared, out, T, d, k, stream);
}
