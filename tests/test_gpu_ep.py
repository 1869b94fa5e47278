"""2-GPU (NCCL) expert-parallel forward vs the single-device oracle.  Skipped unless >= 2 GPUs are visible (the
round-end 1-GPU pytest run skips it; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_ep.py -m gpu`)."""
import tempfile

import pytest
import torch
import torch.multiprocessing as mp

from ep_common import ep_worker, free_port

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("E,k,T,d,I", [(8, 2, 50, 256, 128), (64, 6, 1024, 2560, 1664)])
def test_ep_forward_two_gpus(E, k, T, d, I):
    tc = dict(hidden_size=d, moe_num_experts=E, moe_topk=k, moe_intermediate_size=I, moe_num_shared_experts=2)
    with tempfile.TemporaryDirectory() as tmp:
        mp.spawn(ep_worker, args=(2, free_port(), "cuda", "cuda", tc, T, "bfloat16", tmp), nprocs=2, join=True)
        for r in range(2):
            res = torch.load(f"{tmp}/rank{r}.pt")
            assert res["err_safe"] <= 1e-2 and res["n_safe"] >= res["n"] // 4, res
