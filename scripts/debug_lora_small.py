"""Round-2 starting point: localise the LoRA failure seen on hardware at small shapes (tests/test_zz_gpu_round1_unverified.py).
Runs every op of the adapter forward/backward separately at the failing shapes (E=4, 128 rows) against torch and prints which
one is off; then the same at the passing shape (E=8, 416 rows) as a control.  ~5 s on a GPU."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from aria_b200 import ops

dev = "cuda"


def rel(a, b):
    a, b = a.float(), b.float()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


def check(tag, got, want):
    r = rel(got, want)
    print(f"  {tag:58s} rel-L2 {r:.3e} {'OK' if r < 2e-2 else '<-- OFF'}  finite={bool(torch.isfinite(got.float()).all())}", flush=True)


def run(E, K, N, counts, label):
    print(f"== {label}: E={E} in={K} out={N} counts={counts}")
    g = torch.Generator().manual_seed(0)
    rows = sum(counts)
    off_h = [0]
    for c in counts:
        off_h.append(off_h[-1] + c)
    off = torch.tensor(off_h, dtype=torch.int32, device=dev)
    x = torch.randn(rows, K, generator=g).bfloat16().to(dev)
    dy = torch.randn(rows, N, generator=g).bfloat16().to(dev)
    w = (torch.randn(E, K, N, generator=g) * 0.05).bfloat16().to(dev)
    a_pad = torch.zeros(E, K, 128, dtype=torch.bfloat16, device=dev)
    a_pad[:, :, :8] = (torch.randn(E, K, 8, generator=g) * 0.2).bfloat16().to(dev)
    b_pad = torch.zeros(E, 128, N, dtype=torch.bfloat16, device=dev)
    b_pad[:, :8] = (torch.randn(E, 8, N, generator=g) * 0.05).bfloat16().to(dev)

    def per_group(fn, width):
        out = torch.zeros(rows, width, device=dev)
        for e in range(E):
            lo, hi = off_h[e], off_h[e + 1]
            if hi > lo:
                out[lo:hi] = fn(e, lo, hi)
        return out

    base_ref = per_group(lambda e, lo, hi: x[lo:hi].float() @ w[e].float(), N)
    base = ops.grouped_gemm(x, w, off)
    check("F1 grouped_gemm(x, W)", base, base_ref)
    h = ops.grouped_gemm(x, a_pad, off)
    check("F2 grouped_gemm(x, A_pad)", h, per_group(lambda e, lo, hi: x[lo:hi].float() @ a_pad[e].float(), 128))
    out = ops.grouped_gemm(h, b_pad, off, residual=base)
    check("F3 grouped_gemm(h, B_pad, residual=base)", out, per_group(lambda e, lo, hi: h[lo:hi].float() @ b_pad[e].float(), N) + base.float())
    d_b = ops.grouped_wgrad(h, dy, off)
    check("B1 grouped_wgrad(h, dy)", d_b, torch.stack([h[off_h[e]:off_h[e + 1]].float().t() @ dy[off_h[e]:off_h[e + 1]].float() for e in range(E)]))
    dh = ops.grouped_gemm_nt(dy, b_pad, off)
    check("B2 grouped_gemm_nt(dy, B_pad)", dh, per_group(lambda e, lo, hi: dy[lo:hi].float() @ b_pad[e].float().t(), 128))
    d_a = ops.grouped_wgrad(x, dh, off)
    check("B3 grouped_wgrad(x, dh)", d_a, torch.stack([x[off_h[e]:off_h[e + 1]].float().t() @ dh[off_h[e]:off_h[e + 1]].float() for e in range(E)]))
    dx = ops.grouped_gemm_nt(dy, w, off)
    check("B4 grouped_gemm_nt(dy, W)", dx, per_group(lambda e, lo, hi: dy[lo:hi].float() @ w[e].float().t(), K))
    dx2 = ops.grouped_gemm_nt(dh, a_pad, off, residual=dx)
    check("B5 grouped_gemm_nt(dh, A_pad, residual=dx)", dx2, per_group(lambda e, lo, hi: dh[lo:hi].float() @ a_pad[e].float().t(), K) + dx.float())
    torch.cuda.synchronize()


run(4, 128, 192, [32, 0, 80, 16], "failing golden shape")
run(4, 128, 256, [48, 16, 0, 64], "failing GroupedMLP fc1 shape")
run(4, 128, 128, [48, 16, 0, 64], "failing GroupedMLP fc2 shape")
run(8, 256, 384, [32, 0, 80, 16, 48, 160, 16, 64], "passing shape (control)")
