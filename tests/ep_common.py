"""Shared helpers for the expert-parallel tests (CPU/gloo and GPU/nccl)."""
import os
import socket

import torch


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class OracleBackend:
    """TEST-ONLY compute backend for ExpertParallelMoE: the oracle's CPU functions, so the host-side exchange logic
    (counts all-to-all, split sizes, (source rank, expert) grouping, reverse exchange) runs under gloo without a GPU."""

    def __init__(self, k):
        self.k = k

    def router(self, x, w_router, k):
        from oracle import aria_oracle as O
        s, i, c = O.router_routing(O.router_gating(x, w_router), k)
        return s, i, c

    def permute(self, x, idx, counts):
        from oracle import aria_oracle as O
        perm, order = O.token_permutation(x, idx, self.k)
        return perm, order

    def grouped_mlp(self, rows, fc1, fc2, group_counts, n_local_experts):
        from oracle import aria_oracle as O
        out = torch.zeros(rows.shape[0], fc2.shape[-1], dtype=rows.dtype)
        off = 0
        for g, n in enumerate(group_counts.tolist()):
            if n:
                e = g % n_local_experts
                h = O.glu(rows[off:off + n] @ fc1[e])
                out[off:off + n] = h @ fc2[e]
            off += n
        return out

    def shared(self, x, gate_w, up_w, down_w):
        from oracle import aria_oracle as O
        return O.shared_expert_mlp(x, gate_w, up_w, down_w)

    def combine(self, y, order, scores, shared):
        from oracle import aria_oracle as O
        return O.token_unpermutation(y, order, scores, self.k) + shared


def ep_worker(rank, world, port, backend_name, device_kind, tc, T, dtype_name, result_dir):
    """One rank: EP forward vs the single-device layer on this rank's tokens. Writes max error to result_dir."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    import torch.distributed as dist
    from aria_b200.expert_parallel import ExpertParallelMoE
    from oracle import aria_oracle as O
    from oracle import configs as C

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_grad_enabled(False)
    dtype = getattr(torch, dtype_name)
    if device_kind == "cuda":
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        dev = torch.device("cuda", rank)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
        dev = torch.device("cpu")
    gen = torch.Generator().manual_seed(1234)
    full = {k: v.to(dtype) for k, v in C.moe_layer_state(tc, gen).items()}
    xg = torch.Generator().manual_seed(100 + rank)
    x = torch.randn(T + 3 * rank, tc["hidden_size"], generator=xg).to(dtype)  # ragged: ranks own different token counts
    want, parts = O.moe_layer(x, full, tc["moe_topk"], return_parts=True)
    shard = {k: v.to(dev) for k, v in ExpertParallelMoE.shard_state(full, rank, world).items()}
    backend = OracleBackend(tc["moe_topk"]) if backend_name == "oracle" else None
    ep = ExpertParallelMoE(shard, tc["moe_num_experts"], tc["moe_topk"], backend=backend)
    got = ep(x.to(dev)).float().cpu()
    if device_kind == "cuda":
        torch.cuda.synchronize()
    lg = parts["logits"].float().sort(1, descending=True).values
    k = tc["moe_topk"]
    safe = (lg[:, k - 1] - lg[:, k]) / lg.abs().amax(1) > 2 ** -6
    err = (got - want.float()).abs().amax(-1)
    scale = float(want.float().abs().max())
    torch.save({"err_safe": float(err[safe].max()) / scale, "err_all": float(err.max()) / scale,
                "n_safe": int(safe.sum()), "n": int(safe.numel())}, os.path.join(result_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()
