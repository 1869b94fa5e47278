"""cfg 5: one full-width MoE layer, expert-parallel forward + backward, 8192 tokens per rank (torchrun, nccl)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from aria_b200.expert_parallel import ep_moe_layer_train
rank = int(os.environ.get("RANK", 0)); lr = int(os.environ.get("LOCAL_RANK", 0)); W = int(os.environ.get("WORLD_SIZE", 1))
dev = torch.device("cuda", lr); torch.cuda.set_device(dev)
if W == 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29544")
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
else:
    dist.init_process_group("nccl", device_id=dev)
d, E, k, I, T = 2560, 64, 6, 1664, int(os.environ.get("T_LOC", 8192))
g = torch.Generator(device=dev).manual_seed(7)
def rnd(*s): return (torch.randn(*s, generator=g, device=dev) * 0.02).bfloat16().requires_grad_(True)
w = {"router.weight": rnd(E, d), "experts.fc1.weight": rnd(E // W, d, 2 * I), "experts.fc2.weight": rnd(E // W, I, d),
     "shared_experts.gate_proj.weight": rnd(2 * I, d), "shared_experts.up_proj.weight": rnd(2 * I, d),
     "shared_experts.down_proj.weight": rnd(d, 2 * I)}
gx = torch.Generator(device=dev).manual_seed(100 + rank)
x = torch.randn(T, d, generator=gx, device=dev).bfloat16().requires_grad_(True)
go = torch.randn(T, d, generator=gx, device=dev).bfloat16()
def step():
    for p in list(w.values()) + [x]: p.grad = None
    ep_moe_layer_train(x, w, k).backward(go)
for _ in range(3): step()
dist.barrier(); torch.cuda.synchronize()
n = 10
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n): step()
e1.record(); dist.barrier(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    fl = 3 * T * 204.8e6
    print(json.dumps({"bench": "ep_moe_layer_fwd_bwd", "world": W, "tokens_per_rank": T, "ms_per_layer": float(ms),
                      "tokens_per_s_all_ranks": W * T / float(ms) * 1e3, "tflops_per_rank": fl / float(ms) / 1e9}))
dist.destroy_process_group()
