"""Expert parallelism for the MoE block (BASELINE.json configs[4]).

The reference has no EP: its `TokenDispatcher` is Megatron's all-to-all dispatcher with the communication stripped
(aria/model/moe_lm.py:296-365).  Here experts are block-partitioned over the W ranks of one NVSwitch box (rank r owns
experts [r*E/W, (r+1)*E/W) — a pure dim-0 slice of the HF `experts.fc1/fc2.weight`), tokens stay data-parallel, and
each MoE layer does

    router (local) -> stable sort by expert (= by destination rank) -> all-to-all of per-(rank,expert) counts
    -> all-to-all-v of rows (dispatch) -> grouped expert MLP over rows grouped by (source rank, local expert)
    -> reverse all-to-all-v (combine) -> score-weighted sum + local shared expert

over `torch.distributed` (NCCL on NVLink 5 / NVSwitch: uniform bandwidth, so a flat all-to-all).  The grouped GEMM takes
the (source rank, local expert) groups directly (`group_mod`), so received rows are never re-sorted.  Forward only in
this round; one 4*E-byte D2H of the counts per layer is needed because NCCL's all-to-all-v takes host split sizes.

Parity: the W-rank result equals the single-device `MoELayer` on each rank's tokens (tests/test_ep_gloo.py on CPU
through the oracle backend, tests/test_gpu_ep.py on 2 GPUs).
"""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


class CudaBackend:
    """The product compute backend: our CUDA kernels through aria_b200.ops."""

    def router(self, x, w_router, k):
        from . import ops
        scores, idx, counts, _ = ops.router_topk(x, w_router, k)
        return scores, idx, counts

    def permute(self, x, idx, counts):
        from . import ops
        offsets, dest, src = ops.build_permutation(idx, counts)
        return ops.permute_rows(x, src), dest

    def grouped_mlp(self, rows, fc1, fc2, group_counts, n_local_experts):
        """rows grouped by (source rank, local expert); group_counts: int64 [W*E_loc] on the rows' device."""
        from . import ops
        off = ops.offsets_from_counts(group_counts)
        h = ops.grouped_gemm(rows, fc1, off, swiglu=True, group_mod=n_local_experts)
        return ops.grouped_gemm(h, fc2, off, group_mod=n_local_experts)

    def shared(self, x, gate_w, up_w, down_w):
        from . import ops
        return ops.linear(ops.linear_swiglu(x, gate_w, up_w), down_w)

    def combine(self, y, dest, scores, shared):
        from . import ops
        return ops.unpermute_combine(y, dest, scores, shared)


class ExpertParallelMoE:
    """Expert-parallel `MoELayer.forward` (moe_lm.py:548-577) for one layer.

    weights: dict with the reference parameter names; `experts.fc1.weight` / `experts.fc2.weight` hold ONLY this rank's
    slice [E/W, ...]; router and shared-expert weights are replicated."""

    def __init__(self, weights: dict, num_experts: int, topk: int, group=None, backend=None):
        self.w = weights
        self.E = num_experts
        self.k = topk
        self.group = group
        self.W = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        assert self.E % self.W == 0
        self.E_loc = self.E // self.W
        assert weights["experts.fc1.weight"].shape[0] == self.E_loc
        self.backend = backend or CudaBackend()

    @staticmethod
    def shard_state(full: dict, rank: int, world: int) -> dict:
        """Slice a full MoELayer state dict for `rank` (experts on dim 0; everything else replicated)."""
        E = full["experts.fc1.weight"].shape[0]
        lo, hi = rank * E // world, (rank + 1) * E // world
        out = dict(full)
        out["experts.fc1.weight"] = full["experts.fc1.weight"][lo:hi].contiguous()
        out["experts.fc2.weight"] = full["experts.fc2.weight"][lo:hi].contiguous()
        return out

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        shape = x.shape
        x2 = x.reshape(-1, shape[-1])
        be, W, E_loc = self.backend, self.W, self.E_loc
        scores, idx, counts = be.router(x2, self.w["router.weight"], self.k)
        permuted, dest = be.permute(x2, idx, counts)  # rows sorted by global expert id == by destination rank

        # per-(destination rank, local expert) counts -> everyone learns what it will receive
        send_counts = counts.to(torch.int64).view(W, E_loc)
        recv_counts = torch.empty_like(send_counts)  # [source rank, local expert]
        dist.all_to_all_single(recv_counts, send_counts, group=self.group)
        send_host = send_counts.sum(1).tolist()  # the one host sync per layer (NCCL split sizes live on the host)
        recv_host = recv_counts.sum(1).tolist()

        recv_rows = torch.empty((sum(recv_host), x2.shape[1]), dtype=x2.dtype, device=x2.device)
        dist.all_to_all_single(recv_rows, permuted, output_split_sizes=recv_host, input_split_sizes=send_host,
                               group=self.group)
        # shared expert is local work that overlaps with the exchange on the GPU (separate NCCL stream)
        shared = be.shared(x2, self.w["shared_experts.gate_proj.weight"], self.w["shared_experts.up_proj.weight"],
                           self.w["shared_experts.down_proj.weight"])
        y_recv = be.grouped_mlp(recv_rows, self.w["experts.fc1.weight"], self.w["experts.fc2.weight"],
                                recv_counts.reshape(-1).contiguous(), E_loc)
        y = torch.empty_like(permuted)
        dist.all_to_all_single(y, y_recv, output_split_sizes=send_host, input_split_sizes=recv_host, group=self.group)
        return be.combine(y, dest, scores, shared).view(shape)

    __call__ = forward


def exchange_bytes_per_layer(tokens_per_rank: int, topk: int, hidden: int, world: int) -> float:
    """Expected bytes a rank sends per direction per layer: (W-1)/W of its k*T rows leave the rank."""
    return tokens_per_rank * topk * hidden * 2 * (world - 1) / world
