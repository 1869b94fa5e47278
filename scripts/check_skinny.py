"""Skinny-M (m <= 32) dense GEMM variant: parity vs torch and timing, ARIA_GEMM_SKINNY=0/1 A/B (run twice)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from aria_b200 import ops
torch.manual_seed(0)
dev = torch.device("cuda", 0)
def t(f, n=30):
    for _ in range(5): f()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); tot += a.elapsed_time(b)
    return tot / n * 1e3
bad = 0
for m in (1, 7, 32):
    for (n, k) in ((2560, 2560), (2560, 3328), (100352, 2560), (1152, 4304)):
        x = torch.randn(m, k, device=dev).bfloat16(); w = (torch.randn(n, k, device=dev) * 0.02).bfloat16()
        res = torch.randn(m, n, device=dev).bfloat16()
        y = ops.linear(x, w, residual=res); ref = x.float() @ w.float().T + res.float()
        err = ((y.float() - ref).abs().max() / ref.abs().max()).item()
        bad += err > 1e-2
        print(f"linear m={m} n={n} k={k} relinf {err:.2e}  {t(lambda: ops.linear(x, w)):.1f} us  ({n*k*2/1e3/t(lambda: ops.linear(x, w)):.0f} GB/s)")
    x = torch.randn(m, 2560, device=dev).bfloat16()
    g = (torch.randn(3328, 2560, device=dev) * 0.02).bfloat16(); u = (torch.randn(3328, 2560, device=dev) * 0.02).bfloat16()
    y = ops.linear_swiglu(x, g, u); ref = torch.nn.functional.silu(x.float() @ g.float().T) * (x.float() @ u.float().T)
    err = ((y.float() - ref).abs().max() / ref.abs().max()).item(); bad += err > 1e-2
    us = t(lambda: ops.linear_swiglu(x, g, u))
    print(f"swiglu m={m} relinf {err:.2e} {us:.1f} us ({2*3328*2560*2/1e3/us:.0f} GB/s)")
print("BAD" if bad else "OK")
