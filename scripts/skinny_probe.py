"""Probe: where do ~20 us go in an M=32 dense GEMM that streams only 13 MB?  Back-to-back launches over rotating weights
(bigger than L2 in total) vs a CUDA graph of the same; run under ncu for a per-kernel duration."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from aria_b200 import ops
dev = torch.device("cuda", 0)
M = int(os.environ.get("M", 32))
N, K, R = 2560, 2560, 40          # 40 x 13 MB = 524 MB of weights > L2
x = torch.randn(M, K, device=dev).bfloat16()
ws = [(torch.randn(N, K, device=dev) * 0.02).bfloat16() for _ in range(R)]
outs = [torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(R)]
def sweep():
    for w, o in zip(ws, outs): ops.linear(x, w, out=o)
for _ in range(3): sweep()
torch.cuda.synchronize()
if os.environ.get("NCU"):
    sys.exit(0)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); sweep(); sweep(); b.record(); torch.cuda.synchronize()
print(f"eager back-to-back: {a.elapsed_time(b) / (2 * R) * 1e3:.2f} us per GEMM ({N*K*2/1e3/(a.elapsed_time(b) / (2 * R) * 1e3):.0f} GB/s)")
g = torch.cuda.CUDAGraph()
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s): sweep()
torch.cuda.current_stream().wait_stream(s)
with torch.cuda.graph(g): sweep()
g.replay(); torch.cuda.synchronize()
a.record(); g.replay(); g.replay(); b.record(); torch.cuda.synchronize()
us = a.elapsed_time(b) / (2 * R) * 1e3
print(f"graph replay      : {us:.2f} us per GEMM ({N*K*2/1e3/us:.0f} GB/s)")
