"""Times the dense GEMM shapes of the cfg-2 step in isolation (rotating weight copies so that every launch streams its weights
from HBM like consecutive layers do): ViT qkv / o_proj / fc1 / fc2, LM qkv(+RoPE) / o_proj / shared experts / router."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from aria_b200 import ops, _lib as L
dev = "cuda"
torch.manual_seed(0)
NW = 6
bf = torch.bfloat16
def W(*s): return [(torch.randn(*s, device=dev) * 0.02).to(bf) for _ in range(NW)]
def timeit(name, fn, flops):
    for i in range(3): fn(i % NW)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 30
    for i in range(n): fn(i % NW)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    print(f"{os.environ.get('TAG',''):10s} {name:44s} {us:8.1f} us  {flops / us / 1e6:8.1f} TF/s", flush=True)
# ---- ViT (4900 patches, d 1152, 16 heads x 72, mlp 4304)
N, dv, H, hd, I = 4900, 1152, 16, 72, 4304
x = torch.randn(N, dv, device=dev).to(bf)
xi = torch.randn(N, I, device=dev).to(bf)
res = torch.randn(N, dv, device=dev).to(bf)
wq, wk, wv, wo = W(dv, dv), W(dv, dv), W(dv, dv), W(dv, dv)
b1 = torch.randn(dv, device=dev).to(bf)
w1, w2 = W(I, dv), W(dv, I)
bI = torch.randn(I, device=dev).to(bf)
q = torch.zeros(1, H, N, 128, device=dev, dtype=bf); k = torch.zeros_like(q); v = torch.zeros_like(q)
timeit("ViT qkv_heads 4900x(3x1152)x1152 +bias", lambda i: ops.qkv_heads(x, [wq[i], wk[i], wv[i]], [b1, b1, b1], [q, k, v], hd, N), 2 * N * dv * dv * 3)
timeit("ViT o_proj 4900x1152x1152 +bias +res", lambda i: ops.linear(x, wo[i], b1, residual=res), 2 * N * dv * dv)
timeit("ViT fc1 4900x4304x1152 +bias gelu", lambda i: ops.linear(x, w1[i], bI, act=L.ACT_GELU_TANH), 2 * N * dv * I)
timeit("ViT fc2 4900x1152x4304 +bias +res", lambda i: ops.linear(xi, w2[i], b1, residual=res), 2 * N * dv * I)
# ---- LM at T = 768
T, d, Hl, Is, E = 768, 2560, 20, 3328, 64
xl = torch.randn(T, d, device=dev).to(bf)
rl = torch.randn(T, d, device=dev).to(bf)
lq, lk, lv, lo = W(d, d), W(d, d), W(d, d), W(d, d)
g, u, dn = W(Is, d), W(Is, d), W(d, Is)
wr = W(E, d)
inv = 1.0 / (5e6 ** (torch.arange(0, 128, 2, device=dev).float() / 128))
cos, sin = ops.rope_table(inv, 1024)
ql = torch.zeros(1, Hl, 1024, 128, device=dev, dtype=bf); kl = torch.zeros_like(ql); vl = torch.zeros_like(ql)
timeit("LM qkv_heads 768x(3x2560)x2560 +RoPE", lambda i: ops.qkv_heads(xl, [lq[i], lk[i], lv[i]], [None] * 3, [ql, kl, vl], 128, T, rope_mask=3, rope_cos=cos, rope_sin=sin), 2 * T * d * d * 3)
timeit("LM o_proj 768x2560x2560 +res", lambda i: ops.linear(xl, lo[i], residual=rl), 2 * T * d * d)
timeit("LM shared gate|up swiglu 768x(2x3328)x2560", lambda i: ops.linear_swiglu(xl, g[i], u[i]), 4 * T * d * Is)
hs = torch.randn(T, Is, device=dev).to(bf)
timeit("LM shared down 768x2560x3328", lambda i: ops.linear(hs, dn[i]), 2 * T * d * Is)
timeit("LM router_topk 768x64x2560", lambda i: ops.router_topk(xl, wr[i], 6), 2 * T * d * E)
