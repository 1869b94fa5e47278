"""Times only the ViT-shape attention (16 x 4900 x 4900, hd 72) with the library selected by ARIA_B200_LIB (ablation variants)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from aria_b200 import ops
dev = "cuda"
torch.manual_seed(0)
B, H, T, hd = 1, 16, 4900, 72
sets = []
for _ in range(3):
    q = torch.zeros(B, H, T, 128, device=dev, dtype=torch.bfloat16)
    k, v = torch.zeros_like(q), torch.zeros_like(q)
    for t in (q, k, v):
        t[..., :hd] = torch.randn(B, H, T, hd, device=dev).bfloat16()
    sets.append((q, k, v))
f = lambda q, k, v: ops.attention(q, k, v, T, T, hd ** -0.5, False, out_hd=hd)
for s in sets[:2]:
    f(*s)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(20):
    f(*sets[i % 3])
e1.record()
torch.cuda.synchronize()
print(f"{os.environ.get('TAG', 'default'):40s} {e0.elapsed_time(e1) / 20 * 1e3:8.1f} us", flush=True)
