"""CPU, world_size 2 over gloo: the expert-parallel exchange logic (aria_b200/expert_parallel.py) reproduces the
single-device MoE layer on each rank's tokens.  Compute is the oracle backend (tests/ep_common.py) — the product
backend is CUDA-only; what is under test here is the host-side dispatch/combine plumbing."""
import tempfile

import pytest
import torch
import torch.multiprocessing as mp

from ep_common import ep_worker, free_port


@pytest.mark.parametrize("E,k,T", [(8, 2, 17), (64, 6, 40)])
def test_ep_forward_matches_single_device(E, k, T):
    tc = dict(hidden_size=64, moe_num_experts=E, moe_topk=k, moe_intermediate_size=32, moe_num_shared_experts=2)
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(ep_worker, args=(2, free_port(), "oracle", "cpu", tc, T, "float32", d), nprocs=2, join=True)
        for r in range(2):
            res = torch.load(f"{d}/rank{r}.pt")
            assert res["err_all"] <= 1e-5, res


def test_exchange_volume_formula():
    from aria_b200.expert_parallel import exchange_bytes_per_layer
    # 8192 tokens/rank, k=6, d=2560, W=8 -> ~220 MB per direction per layer
    assert abs(exchange_bytes_per_layer(8192, 6, 2560, 8) - 220.2e6) < 1e6
