"""Tiny drivers for ncu captures: `attn` (ViT-shape attention), `attn_causal`, `fc1` (cfg-2 fc1 grouped GEMM + SwiGLU),
`fc2`, `dense` (ViT fc1 GEMM with gelu), `dense_big` (8192^3 on the 2-CTA kernel)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from aria_b200 import ops, _lib as L
dev = "cuda"
mode = sys.argv[1]
torch.manual_seed(0)
if mode == "attn72":   # the ViT shape as the model runs it: 16 heads x 4900 x 4900, head dim 72 carried in 128-wide rows
    B, H, T, hd = 1, 16, 4900, 72
    q = torch.zeros(B, H, T, 128, device=dev, dtype=torch.bfloat16)
    k, v = torch.zeros_like(q), torch.zeros_like(q)
    for t in (q, k, v):
        t[..., :hd] = torch.randn(B, H, T, hd, device=dev).bfloat16()
    for _ in range(2):
        ops.attention(q, k, v, T, T, hd ** -0.5, False, out_hd=hd)
elif mode in ("attn", "attn_causal"):
    B, H, T = (1, 16, 4900) if mode == "attn" else (1, 20, 8192)
    q = torch.randn(B, H, T, 128, device=dev).bfloat16()
    k = torch.randn(B, H, T, 128, device=dev).bfloat16()
    v = torch.randn(B, H, T, 128, device=dev).bfloat16()
    for _ in range(2):
        ops.attention(q, k, v, T, T, 128 ** -0.5, mode == "attn_causal")
elif mode in ("fc1", "fc2"):
    E, d, I, T = 64, 2560, 1664, 768
    rows = T * 6
    g = torch.Generator().manual_seed(3)
    counts = torch.randint(50, 95, (E,), generator=g)
    counts[-1] += rows - counts.sum()
    off = torch.cat([torch.zeros(1, dtype=torch.long), counts.cumsum(0)]).to(torch.int32).to(dev)
    if mode == "fc1":
        w = (torch.randn(E, d, 2 * I, device=dev) * 0.02).bfloat16()
        a = torch.randn(rows, d, device=dev).bfloat16()
        for _ in range(3):
            ops.grouped_gemm(a, w, off, swiglu=True)
    else:
        w = (torch.randn(E, I, d, device=dev) * 0.02).bfloat16()
        a = torch.randn(rows, I, device=dev).bfloat16()
        for _ in range(3):
            ops.grouped_gemm(a, w, off)
elif mode in ("fc1_big", "wgrad_big"):
    E, d, I, T = 64, 2560, 1664, 8192
    rows = T * 6
    off = torch.arange(0, rows + 1, rows // E, dtype=torch.int32, device=dev)
    a = torch.randn(rows, d, device=dev).bfloat16()
    if mode == "fc1_big":
        w = (torch.randn(E, d, 2 * I, device=dev) * 0.02).bfloat16()
        for _ in range(3):
            ops.grouped_gemm(a, w, off, swiglu=True)
    else:
        g = torch.randn(rows, 2 * I, device=dev).bfloat16()
        for _ in range(3):
            ops.grouped_wgrad(a, g, off)
elif mode == "small":
    T, d, E, k = 8192, 2560, 64, 6
    x = torch.randn(T, d, device=dev).bfloat16()
    w = torch.ones(d, device=dev).bfloat16()
    logits = torch.randn(T, E, device=dev).bfloat16()
    for _ in range(2):
        s, i, c = ops.route_from_logits(logits, k)
        off, dest, src = ops.build_permutation(i, c)
        p = ops.permute_rows(x, src)
        o = ops.unpermute_combine(p, dest, s, x)
        n = ops.rmsnorm(x, w, 1e-5, residual=o)
elif mode == "dense":
    x = torch.randn(4900, 1152, device=dev).bfloat16()
    w = (torch.randn(4304, 1152, device=dev) * 0.02).bfloat16()
    b = torch.randn(4304, device=dev).bfloat16()
    for _ in range(3):
        ops.linear(x, w, b, act=L.ACT_GELU_TANH)
elif mode == "dense_lm":   # LM o_proj / down_proj shape at the cfg-2 prefill: 768 tokens, 2560 x 2560
    x = torch.randn(768, 2560, device=dev).bfloat16()
    w = (torch.randn(2560, 2560, device=dev) * 0.02).bfloat16()
    for _ in range(3):
        ops.linear(x, w)
elif mode == "dense_oproj":  # ViT o_proj: 4900 x 1152 x 1152 with residual
    x = torch.randn(4900, 1152, device=dev).bfloat16()
    w = (torch.randn(1152, 1152, device=dev) * 0.02).bfloat16()
    b = torch.randn(1152, device=dev).bfloat16()
    r = torch.randn(4900, 1152, device=dev).bfloat16()
    for _ in range(3):
        ops.linear(x, w, b, residual=r)
elif mode in ("ep_fc1_expert_major", "ep_fc1_source_major"):
    # the fc1 grouped GEMM of one rank of a W=2 expert-parallel cfg-2 prefill: 32 local experts x 2 source ranks = 64 fixed-capacity
    # regions with ~72 rows each; region order (expert, source) [group_mod = -W] vs (source, expert) [group_mod = E_loc]
    W_, E_loc, d, I, cap = 2, 32, 2560, 1664, 784
    G = W_ * E_loc
    g = torch.Generator().manual_seed(3)
    counts = torch.randint(50, 95, (G,), generator=g).to(torch.int32).to(dev)
    starts = (torch.arange(G, dtype=torch.int32) * cap).to(dev)
    w = (torch.randn(E_loc, d, 2 * I, device=dev) * 0.02).bfloat16()
    a = torch.randn(G * cap, d, device=dev).bfloat16()
    out = torch.empty(G * cap, I, device=dev, dtype=torch.bfloat16)
    gm = -W_ if mode == "ep_fc1_expert_major" else E_loc
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        ops.grouped_gemm_regions(a, w, starts, counts, 4608, swiglu=True, group_mod=gm, out=out)
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(5):
        flush.zero_()
        e0.record()
        ops.grouped_gemm_regions(a, w, starts, counts, 4608, swiglu=True, group_mod=gm, out=out)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print(mode, "us per launch (L2 flushed):", [round(t, 1) for t in ts], "weights MB", w.numel() * 2 / 1e6)
elif mode == "dense_big":
    x = torch.randn(8192, 8192, device=dev).bfloat16()
    w = (torch.randn(8192, 8192, device=dev) * 0.02).bfloat16()
    for _ in range(3):
        ops.linear(x, w)
torch.cuda.synchronize()
