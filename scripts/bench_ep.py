"""cfg-5 style microbenchmark: one full-width MoE layer (d=2560, E=64, k=6, I=1664), expert-parallel FORWARD over W ranks,
8192 tokens per rank.  Launch with torchrun (nccl).  Prints one JSON line from rank 0.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/bench_ep.py"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from aria_b200.expert_parallel import ExpertParallelMoE, PeerTransport, exchange_bytes_per_layer
from aria_b200 import moe_lm

rank = int(os.environ.get("RANK", 0)); lr = int(os.environ.get("LOCAL_RANK", 0)); W = int(os.environ.get("WORLD_SIZE", 1))
dev = torch.device("cuda", lr); torch.cuda.set_device(dev)
if W > 1:
    dist.init_process_group("nccl", device_id=dev)
else:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
torch.set_grad_enabled(False)
d, E, k, I, T = 2560, 64, 6, 1664, int(os.environ.get("T_LOC", 8192))
g = torch.Generator(device=dev).manual_seed(7)
def rnd(*s): return (torch.randn(*s, generator=g, device=dev) * 0.02).bfloat16()
w = {"router.weight": rnd(E, d), "experts.fc1.weight": rnd(E // W, d, 2 * I), "experts.fc2.weight": rnd(E // W, I, d),
     "shared_experts.gate_proj.weight": rnd(2 * I, d), "shared_experts.up_proj.weight": rnd(2 * I, d),
     "shared_experts.down_proj.weight": rnd(d, 2 * I)}
gx = torch.Generator(device=dev).manual_seed(100 + rank)
x = torch.randn(T, d, generator=gx, device=dev).bfloat16()
mode = os.environ.get("EP_TRANSPORT", "nccl")
transport = PeerTransport(T, d, E, k, dev) if mode == "p2p" else None
ep = ExpertParallelMoE(w, E, k, transport=transport)
for _ in range(3): ep(x)
dist.barrier(); torch.cuda.synchronize()
n = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n): y = ep(x)
e1.record(); dist.barrier(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    flops = T * 204.8e6  # per rank, balanced
    print(json.dumps({"bench": "ep_moe_layer_forward", "transport": mode, "world": W, "tokens_per_rank": T, "ms_per_layer": float(ms),
                      "tokens_per_s_all_ranks": W * T / float(ms) * 1e3, "tflops_per_rank": flops / float(ms) / 1e9,
                      "a2a_bytes_per_direction": exchange_bytes_per_layer(T, k, d, W)}))
dist.destroy_process_group()
