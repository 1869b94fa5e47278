"""CPU, world_size 2 over gloo: gradient hand-off of an expert-parallel MoE layer to a data-parallel optimizer
(aria_b200/ep_grads.py).  A toy MoE layer in plain torch: replicated router + shared MLP, experts sharded over the
ranks with an all-to-all of token rows (autograd-aware), loss = mean over ALL ranks' tokens.  After sync_gradients every rank must
hold exactly the gradients of the single-process model on the concatenated batch: replicated ones all-reduced, expert shards untouched
apart from the 1/W of the mean."""
import os
import tempfile

import torch
import torch.multiprocessing as mp

from ep_common import free_port

E, K, D, I, T = 4, 2, 16, 8, 12


def _params(seed=0):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: (torch.randn(*s, generator=g) * 0.3).double()
    return {"router.weight": r(E, D), "experts.fc1.weight": r(E, D, I), "experts.fc2.weight": r(E, I, D), "shared.weight": r(D, D)}


def _layer(x, w, experts_of=None):
    """Dense reference math: y[t] = sum_j s[t,j] * fc2[e](tanh(fc1[e](x[t]))) + shared(x[t]); experts_of = set of expert ids to use."""
    logits = x @ w["router.weight"].T
    top, idx = logits.topk(K, dim=-1)
    s = top.softmax(-1)
    y = x @ w["shared.weight"].T
    for e in range(E):
        if experts_of is not None and e not in experts_of:
            continue
        m = (idx == e)
        if m.any():
            t_ids, j_ids = m.nonzero(as_tuple=True)
            out = torch.tanh(x[t_ids] @ w["experts.fc1.weight"][e]) @ w["experts.fc2.weight"][e]
            y = y.index_add(0, t_ids, out * s[t_ids, j_ids, None])
    return y


def _worker(rank, world, port, tmp):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    import torch.distributed.nn.functional as dfn
    from aria_b200.ep_grads import sync_gradients
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = _params()
    lo, hi = rank * E // world, (rank + 1) * E // world
    p = {"router.weight": full["router.weight"].clone().requires_grad_(True),
         "shared.weight": full["shared.weight"].clone().requires_grad_(True),
         "experts.fc1.weight": full["experts.fc1.weight"][lo:hi].clone().requires_grad_(True),   # this rank's shard
         "experts.fc2.weight": full["experts.fc2.weight"][lo:hi].clone().requires_grad_(True)}
    xs = [torch.randn(T, D, generator=torch.Generator().manual_seed(100 + r)).double() for r in range(world)]
    x = xs[rank]
    # expert-parallel forward: every rank sends all its rows + routing to every owner (a dense stand-in for the dispatch: what matters
    # here is WHICH gradients end up where), owners compute their experts' contributions for all ranks' tokens, results go back
    logits = x @ p["router.weight"].T
    top, idx = logits.topk(K, dim=-1)
    s = top.softmax(-1)
    all_x = dfn.all_gather(x)                      # autograd-aware: grads flow back to the owning rank's x
    all_idx = [torch.empty_like(idx) for _ in range(world)]
    dist.all_gather(all_idx, idx)
    contrib = []                                    # my experts' outputs for rank r's (token, slot) pairs
    for r in range(world):
        out = torch.zeros(T, K, D, dtype=torch.double)
        for el in range(hi - lo):
            m = all_idx[r] == lo + el
            if m.any():
                t_ids, j_ids = m.nonzero(as_tuple=True)
                o = torch.tanh(all_x[r][t_ids] @ p["experts.fc1.weight"][el]) @ p["experts.fc2.weight"][el]
                out = out.index_put((t_ids, j_ids), o)
        contrib.append(out)
    back = dfn.all_to_all([torch.empty(T, K, D, dtype=torch.double) for _ in range(world)], contrib)   # from every owner: my rows
    y = x @ p["shared.weight"].T + (sum(back) * s[..., None]).sum(1)
    loss = (y ** 2).mean() / world               # global mean over the W ranks' equally sized batches
    loss.backward()
    rep = sync_gradients(p.items(), average=False)   # the 1/W is already in the loss; the mean variant is checked below
    # single-process reference on the concatenated batch
    ref = {k: v.clone().requires_grad_(True) for k, v in full.items()}
    yr = _layer(torch.cat(xs), ref)
    ((yr ** 2).mean()).backward()
    err = {"router": float((p["router.weight"].grad - ref["router.weight"].grad).abs().max()),
           "shared": float((p["shared.weight"].grad - ref["shared.weight"].grad).abs().max()),
           "fc1": float((p["experts.fc1.weight"].grad - ref["experts.fc1.weight"].grad[lo:hi]).abs().max()),
           "fc2": float((p["experts.fc2.weight"].grad - ref["experts.fc2.weight"].grad[lo:hi]).abs().max()),
           "forward": float((y - yr[rank * T:(rank + 1) * T]).abs().max()), "report": rep}
    # average=True divides BOTH kinds by W (per-rank mean losses)
    g0 = {k: v.grad.clone() for k, v in p.items()}
    sync_gradients(p.items(), average=True)
    err["avg_expert"] = float((p["experts.fc1.weight"].grad - g0["experts.fc1.weight"] / world).abs().max())
    err["avg_rep"] = float((p["router.weight"].grad - g0["router.weight"]).abs().max())   # identical on all ranks already: sum / W = itself
    torch.save(err, f"{tmp}/rank{rank}.pt")
    dist.destroy_process_group()


def test_expert_parallel_gradients_match_single_process_after_sync():
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(_worker, args=(2, free_port(), tmp), nprocs=2, join=True)
        for r in range(2):
            e = torch.load(f"{tmp}/rank{r}.pt")
            assert e["forward"] < 1e-12 and e["router"] < 1e-12 and e["shared"] < 1e-12 and e["fc1"] < 1e-12 and e["fc2"] < 1e-12, e
            assert e["avg_expert"] < 1e-15 and e["avg_rep"] < 1e-12, e
            assert e["report"]["replicated_tensors"] == 2 and e["report"]["expert_shard_tensors"] == 2 and e["report"]["world"] == 2


def test_expert_rule_matches_reference_parameter_names():
    from aria_b200.ep_grads import is_expert_shard
    assert is_expert_shard("language_model.model.layers.3.mlp.experts.fc1.weight")
    assert is_expert_shard("experts.fc2.weight")
    assert not is_expert_shard("language_model.model.layers.3.mlp.shared_experts.gate_proj.weight")
    assert not is_expert_shard("language_model.model.layers.3.mlp.router.weight")
