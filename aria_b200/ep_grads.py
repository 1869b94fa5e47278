"""Gradient hand-off of an expert-parallel MoE model to a data-parallel optimizer (how the expert-sharded layers
sit inside the reference's data-parallel launcher, recipes/accelerate_configs/zero2.yaml:1-23 with aria/train.py:229).

Under expert parallelism (aria_b200/expert_parallel.py) tokens stay data-parallel and the routed experts are sharded: after a
backward pass
  * replicated parameters (router, shared experts, attention, norms, embeddings) hold PER-RANK partial gradients — they need the
    usual data-parallel reduction, exactly what DDP / ZeRO-2's reduce-scatter does for every parameter of the reference;
  * the expert shards (`experts.fc1.weight[lo:hi]`, `experts.fc2.weight[lo:hi]`) hold gradients that are already COMPLETE sums over
    the tokens of ALL ranks (every token routed to a local expert was sent here by the dispatch) — reducing them again would count
    each token W times.  They only need the same normalisation as the rest: a loss averaged over the W ranks' batches means 1/W.
`sync_gradients` applies that rule in place.  With ZeRO-1/2 the optimizer states of the expert shards are then naturally
partitioned by the expert sharding itself (each rank owns its experts), the replicated parameters' states by the launcher.
Host-side logic only (torch.distributed collectives on the caller's process group); no kernels involved."""
from __future__ import annotations

from typing import Callable, Iterable, Tuple

import torch
import torch.distributed as dist


def is_expert_shard(name: str) -> bool:
    """Default rule: the reference's parameter names of the routed experts (moe_lm.py:498-503)."""
    return ".experts.fc1." in f".{name}" or ".experts.fc2." in f".{name}"


@torch.no_grad()
def sync_gradients(named_parameters: Iterable[Tuple[str, torch.nn.Parameter]], group=None, average: bool = True,
                   expert_rule: Callable[[str], bool] = is_expert_shard, bucket_bytes: int = 64 << 20) -> dict:
    """All-reduce the gradients of replicated parameters over `group` (bucketed, flat buffers), leave expert-shard gradients
    local; with `average` both are divided by the world size (the loss is a mean over ranks).  Returns a small report."""
    W = dist.get_world_size(group)
    rep, exp = [], []
    for name, p in named_parameters:
        if p.grad is None:
            continue
        (exp if expert_rule(name) else rep).append(p.grad)
    scale = 1.0 / W if average else 1.0
    n_buckets = 0
    i = 0
    while i < len(rep):
        # bucket of same-dtype grads up to bucket_bytes
        j, size, dt, dev = i, 0, rep[i].dtype, rep[i].device
        while j < len(rep) and rep[j].dtype == dt and rep[j].device == dev and (size == 0 or size + rep[j].numel() * rep[j].element_size() <= bucket_bytes):
            size += rep[j].numel() * rep[j].element_size()
            j += 1
        flat = torch.cat([g.reshape(-1) for g in rep[i:j]])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        if average:
            flat.mul_(scale)
        off = 0
        for g in rep[i:j]:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        n_buckets += 1
        i = j
    if average:
        for g in exp:
            g.mul_(scale)
    return {"world": W, "replicated_tensors": len(rep), "expert_shard_tensors": len(exp), "buckets": n_buckets}
