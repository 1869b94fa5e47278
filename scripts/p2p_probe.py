"""2-GPU probe: map a peer GPU's buffer into this process (CUDA IPC through torch's storage sharing), check that a kernel on
this GPU can read/write it, and measure P2P store bandwidth of the row-copy kernel."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
rank = int(os.environ["RANK"]); W = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", lr); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
from aria_b200 import ops
from aria_b200.peer import PeerArena
print(rank, "can access peer:", [torch.cuda.can_device_access_peer(lr, p) for p in range(W) if p != lr], flush=True)
rows, d = 49152, 2560
arena = PeerArena(rows * d * 2, dev)
buf = arena.local_view(0, (rows, d), torch.bfloat16)
peer = (rank + 1) % W
x = torch.full((rows, d), float(rank + 1), dtype=torch.bfloat16, device=dev)
src = torch.arange(rows, dtype=torch.int32, device=dev)
dist.barrier(); torch.cuda.synchronize()
ops.permute_rows_to_ptr(x, src, arena.ptr(peer))   # OUR row-copy kernel storing into the peer GPU's arena
torch.cuda.synchronize(); dist.barrier()
print(rank, "local arena now holds", float(buf.float().mean()), "(expected", float((rank - 1) % W + 1), ")", flush=True)
for _ in range(3): ops.permute_rows_to_ptr(x, src, arena.ptr(peer))
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ops.permute_rows_to_ptr(x, src, arena.ptr(peer))
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(rank, f"P2P row store: {rows * d * 2 / ms / 1e6:.1f} GB/s ({ms:.3f} ms for {rows * d * 2 / 1e6:.0f} MB)", flush=True)
dist.barrier(); dist.destroy_process_group()
