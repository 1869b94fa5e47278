"""GPU bring-up battery (run under gpurun). Each group runs in its own subprocess so that a trapping kernel
does not poison the others.  Writes gpurun_out/bringup.log.

    python scripts/bringup.py            # all groups
    python scripts/bringup.py gemm_mn    # one group (in-process)
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

GROUPS = ["gemm_dense", "gemm_mn", "gemm_mn_sweep", "swiglu", "heads", "routing", "rowwise", "attn", "attn_decode", "perf", "perf_attn"]


def rel(a, b):
    import torch
    a, b = a.float(), b.float()
    return float((a - b).abs().max()), float(b.abs().max())


def report(name, got, want, tol):
    err, mag = rel(got, want)
    ok = err <= tol * max(mag, 1e-6)
    print(f"  {name}: max_abs_err={err:.4e} ref_max={mag:.4e} -> {'PASS' if ok else 'FAIL'}", flush=True)
    return ok


def g_gemm_dense():
    import torch
    from aria_b200 import ops, _lib as L
    torch.manual_seed(0)
    dev = "cuda"
    for (M, N, K) in [(128, 128, 64), (300, 256, 512), (4900, 4304, 1152), (1, 1024, 2560), (768, 64, 2560)]:
        x = torch.randn(M, K, device=dev).bfloat16()
        w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
        b = torch.randn(N, device=dev).bfloat16()
        r = torch.randn(M, N, device=dev).bfloat16()
        y = ops.linear(x, w)
        torch.cuda.synchronize()
        ref = x.float() @ w.float().t()
        report(f"linear {M}x{N}x{K}", y, ref.bfloat16(), 2e-2)
        y = ops.linear(x, w, bias=b, act=L.ACT_GELU_TANH, residual=r)
        torch.cuda.synchronize()
        ref2 = torch.nn.functional.gelu((ref + b.float()).bfloat16().float(), approximate="tanh").bfloat16().float() + r.float()
        report(f"linear+bias+gelu+res {M}x{N}x{K}", y, ref2.bfloat16(), 2e-2)


def _grouped_ref(a, b, counts):
    import torch
    out = torch.zeros(a.shape[0], b.shape[-1], device=a.device)
    off = 0
    for e, n in enumerate(counts):
        out[off:off + n] = a[off:off + n].float() @ b[e].float()
        off += n
    return out


def g_gemm_mn(dbg=(0, 0, 0), quiet=False):
    import torch
    from aria_b200 import ops
    torch.manual_seed(1)
    dev = "cuda"
    oks = []
    for (E, K, N, counts) in [(1, 64, 64, [128]), (1, 128, 128, [128]), (4, 256, 256, [5, 0, 300, 77]), (8, 512, 1024, [64] * 8),
                              (64, 2560, 3328, None)]:
        if counts is None:
            g = torch.Generator().manual_seed(3)
            counts = torch.randint(40, 110, (E,), generator=g).tolist()
        rows = sum(counts)
        a = torch.randn(rows, K, device=dev).bfloat16()
        b = (torch.randn(E, K, N, device=dev) * 0.05).bfloat16()
        off = torch.tensor([0] + list(torch.tensor(counts).cumsum(0)), dtype=torch.int32, device=dev)
        y = ops.grouped_gemm(a, b, off, dbg=dbg)
        torch.cuda.synchronize()
        ref = _grouped_ref(a, b, counts)
        if quiet:
            err, mag = rel(y, ref)
            oks.append(err <= 2e-2 * mag)
        else:
            oks.append(report(f"grouped E={E} K={K} N={N} rows={rows}", y, ref.bfloat16(), 2e-2))
    return all(oks)


def g_gemm_mn_sweep():
    """If the default MN-major descriptor is wrong, find the (LBO, SBO, K-advance) that is right."""
    if g_gemm_mn(quiet=True):
        print("  default descriptor OK; sweep skipped")
        return
    for lbo in (8192, 1024, 128, 16, 2048):
        for sbo in (1024, 8192, 128, 2048):
            for kadv in (2048, 32, 1024, 256):
                try:
                    ok = g_gemm_mn((lbo, sbo, kadv), quiet=True)
                except Exception as e:  # noqa
                    print(f"  lbo={lbo} sbo={sbo} kadv={kadv}: EXC {e}")
                    return
                print(f"  lbo={lbo} sbo={sbo} kadv={kadv}: {'OK' if ok else 'bad'}", flush=True)
                if ok:
                    return


def g_swiglu():
    import torch
    import torch.nn.functional as F
    from aria_b200 import ops
    torch.manual_seed(2)
    dev = "cuda"
    M, K, I = 333, 512, 256
    x = torch.randn(M, K, device=dev).bfloat16()
    gw = (torch.randn(I, K, device=dev) * 0.05).bfloat16()
    uw = (torch.randn(I, K, device=dev) * 0.05).bfloat16()
    y = ops.linear_swiglu(x, gw, uw)
    torch.cuda.synchronize()
    g = (x.float() @ gw.float().t()).bfloat16()
    u = (x.float() @ uw.float().t()).bfloat16()
    ref = (F.silu(g.float()).bfloat16().float() * u.float()).bfloat16()
    report("swiglu NK", y, ref, 2e-2)
    E, counts = 4, [100, 3, 0, 200]
    rows = sum(counts)
    a = torch.randn(rows, K, device=dev).bfloat16()
    b = (torch.randn(E, K, 2 * I, device=dev) * 0.05).bfloat16()
    off = torch.tensor([0, 100, 103, 103, 303], dtype=torch.int32, device=dev)
    y = ops.grouped_gemm(a, b, off, swiglu=True)
    torch.cuda.synchronize()
    h = _grouped_ref(a, b, counts).bfloat16()
    ref = (F.silu(h[:, :I].float()).bfloat16().float() * h[:, I:].float()).bfloat16()
    report("swiglu grouped GKN", y, ref, 2e-2)


def g_heads():
    import torch
    from aria_b200 import ops
    from oracle import aria_oracle as O
    torch.manual_seed(3)
    dev = "cuda"
    B, T, H, hd, K = 2, 70, 3, 128, 256
    x = torch.randn(B * T, K, device=dev).bfloat16()
    ws = [(torch.randn(H * hd, K, device=dev) * 0.05).bfloat16() for _ in range(3)]
    Tmax = 96
    outs = [torch.zeros(B, H, Tmax, hd, device=dev, dtype=torch.bfloat16) for _ in range(3)]
    inv_freq = (1.0 / (5e6 ** (torch.arange(0, hd, 2, dtype=torch.int64).float() / hd))).to(dev)
    cos, sin = ops.rope_table(inv_freq, Tmax)
    pos0 = 10
    ops.qkv_heads(x, ws, [None] * 3, outs, hd, T, pos0=pos0, rope_mask=0b011, rope_cos=cos, rope_sin=sin)
    torch.cuda.synchronize()
    xc = x.cpu()
    q = torch.nn.functional.linear(xc, ws[0].cpu()).view(B, T, H, hd).transpose(1, 2)
    k = torch.nn.functional.linear(xc, ws[1].cpu()).view(B, T, H, hd).transpose(1, 2)
    v = torch.nn.functional.linear(xc, ws[2].cpu()).view(B, T, H, hd).transpose(1, 2)
    pos = (torch.arange(T) + pos0)[None].expand(B, T)
    c, s = O.rope_cos_sin(pos, hd, 5e6, torch.bfloat16)
    report("rope table cos", cos[pos0:pos0 + T].cpu(), c[0], 1e-2)
    qr, kr = O.apply_rope(q, k, c, s)
    report("q rope", outs[0][:, :, pos0:pos0 + T].cpu(), qr, 2e-2)
    report("k rope", outs[1][:, :, pos0:pos0 + T].cpu(), kr, 2e-2)
    report("v", outs[2][:, :, pos0:pos0 + T].cpu(), v, 2e-2)
    # ViT-style: hd=72 padded to 128, bias, no rope
    H2, hd2, K2 = 2, 72, 144
    x = torch.randn(50, K2, device=dev).bfloat16()
    ws = [(torch.randn(H2 * hd2, K2, device=dev) * 0.05).bfloat16() for _ in range(3)]
    bs = [torch.randn(H2 * hd2, device=dev).bfloat16() for _ in range(3)]
    outs = [torch.zeros(1, H2, 50, 128, device=dev, dtype=torch.bfloat16) for _ in range(3)]
    ops.qkv_heads(x, ws, bs, outs, hd2, 50)
    torch.cuda.synchronize()
    for i in range(3):
        ref = torch.nn.functional.linear(x.cpu(), ws[i].cpu(), bs[i].cpu()).view(1, 50, H2, hd2).transpose(1, 2)
        report(f"vit heads seg{i}", outs[i][..., :hd2].cpu(), ref, 2e-2)
        print("   pad zero:", float(outs[i][..., hd2:].abs().max()))


def g_routing():
    import torch
    from aria_b200 import ops
    from oracle import aria_oracle as O
    torch.manual_seed(4)
    dev = "cuda"
    for (T, E, k, d) in [(37, 8, 2, 256), (768, 64, 6, 2560), (5000, 64, 6, 2560)]:
        logits = torch.randn(T, E).bfloat16()
        s_ref, i_ref, c_ref = O.router_routing(logits, k)
        s, i, c = ops.route_from_logits(logits.to(dev), k)
        torch.cuda.synchronize()
        print(f"  route T={T} E={E} k={k}: idx_equal={bool((i.cpu() == i_ref).all())} counts_equal={bool((c.cpu() == c_ref).all())} "
              f"score_maxdiff={float((s.cpu().float() - s_ref.float()).abs().max()):.3e}", flush=True)
        off, dest, src = ops.build_permutation(i, c)
        torch.cuda.synchronize()
        x = torch.randn(T, d).bfloat16()
        perm_ref, order = O.token_permutation(x, i_ref, k)
        inv = torch.empty_like(order)
        inv[order] = torch.arange(order.numel())
        print(f"    perm: dest_equal={bool((dest.cpu() == inv).all())} src_equal={bool((src.cpu() == order // k).all())} "
              f"offsets_ok={bool((off.cpu()[1:] == c_ref.cumsum(0)).all())}", flush=True)
        p = ops.permute_rows(x.to(dev), src)
        torch.cuda.synchronize()
        print(f"    permute_rows equal={bool((p.cpu() == perm_ref).all())}")
        y = torch.randn(T * k, d).bfloat16()
        sh = torch.randn(T, d).bfloat16()
        ref = O.token_unpermutation(y, order, s_ref, k) + sh
        out = ops.unpermute_combine(y.to(dev), dest, s, sh.to(dev))
        torch.cuda.synchronize()
        print(f"    combine maxdiff={float((out.cpu().float() - ref.float()).abs().max()):.3e} "
              f"exact_frac={float((out.cpu() == ref).float().mean()):.5f}")
        # fused gating
        w = (torch.randn(E, d) * 0.02).bfloat16()
        s2, i2, c2, lg = ops.router_topk(x.to(dev), w.to(dev), k)
        torch.cuda.synchronize()
        lref = O.router_gating(x, w)
        print(f"    gating logits maxdiff={float((lg.cpu().float() - lref.float()).abs().max()):.3e} "
              f"same_sets_frac={float((i2.cpu().sort(1).values == O.router_routing(lref, k)[1].sort(1).values).all(1).float().mean()):.4f}")


def g_rowwise():
    import torch
    import torch.nn.functional as F
    from aria_b200 import ops
    from oracle import aria_oracle as O
    torch.manual_seed(5)
    dev = "cuda"
    for d in (256, 1152, 2560):
        x = torch.randn(77, d).bfloat16()
        r = torch.randn(77, d).bfloat16()
        w = (1 + 0.1 * torch.randn(d)).bfloat16()
        b = (0.1 * torch.randn(d)).bfloat16()
        y = ops.rmsnorm(x.to(dev), w.to(dev), 1e-5)
        ref = O.rms_norm(x, w, 1e-5)
        print(f"  rmsnorm d={d}: maxdiff={float((y.cpu().float() - ref.float()).abs().max()):.3e} exact={float((y.cpu() == ref).float().mean()):.5f}")
        y, s = ops.rmsnorm(x.to(dev), w.to(dev), 1e-5, residual=r.to(dev))
        ref = O.rms_norm(x + r, w, 1e-5)
        print(f"  rmsnorm+res d={d}: maxdiff={float((y.cpu().float() - ref.float()).abs().max()):.3e} sum_exact={bool((s.cpu() == x + r).all())}")
        if d <= 2560:
            y = ops.layernorm(x.to(dev), w.to(dev), b.to(dev), 1e-6)
            ref = F.layer_norm(x, (d,), w, b, 1e-6)
            print(f"  layernorm d={d}: maxdiff={float((y.cpu().float() - ref.float()).abs().max()):.3e} exact={float((y.cpu() == ref).float().mean()):.5f}")
    ids = torch.randint(0, 100, (50,))
    ids[5:13] = 9
    table = torch.randn(100, 256).bfloat16()
    e = ops.embedding(ids.to(dev), table.to(dev))
    print("  embedding exact:", bool((e.cpu() == table[ids]).all()))
    feats = torch.randn(8, 256).bfloat16()
    cnt = torch.zeros(1, dtype=torch.int32, device=dev)
    ops.merge_image_features(ids.to(dev), 9, feats.to(dev), e, cnt)
    ref = table[ids].clone()
    ref[ids == 9] = feats
    print("  merge exact:", bool((e.cpu() == ref).all()), "count", int(cnt))
    pix = torch.randn(2, 3, 56, 56).bfloat16()
    pt = ops.im2col_patches(pix.to(dev), 14, 592)
    ref = F.unfold(pix.float(), 14, stride=14).transpose(1, 2).reshape(-1, 588).bfloat16()
    print("  im2col exact:", bool((pt.cpu()[:, :588] == ref).all()), "pad zero:", float(pt[:, 588:].abs().max()))
    pos = torch.randint(0, 100, (50,))
    y = ops.add_pos_embedding(e, pos.to(dev), table.to(dev))
    print("  add_pos exact:", bool((y.cpu() == (e.cpu() + table[pos])).all()))


def _attn_ref(q, k, v, scale, causal, mask=None):
    import torch
    s = (q.float() @ k.float().transpose(-1, -2)) * scale
    Tq, Tk = s.shape[-2:]
    if causal:
        i = torch.arange(Tq, device=s.device)[:, None] + (Tk - Tq)
        j = torch.arange(Tk, device=s.device)[None, :]
        s = s.masked_fill(j > i, float("-inf"))
    if mask is not None:
        s = s.masked_fill(mask[:, None, None, :].bool(), float("-inf"))
    p = torch.softmax(s, -1)
    return (p @ v.float()).transpose(1, 2)  # [B,Tq,H,D]


def g_attn():
    import torch
    from aria_b200 import ops
    torch.manual_seed(6)
    dev = "cuda"
    for (B, H, Tq, Tk, causal, masked, out_hd) in [(1, 1, 128, 128, False, False, 128), (1, 2, 256, 256, True, False, 128),
                                                   (2, 3, 300, 300, True, False, 128), (1, 2, 100, 420, True, False, 128),
                                                   (2, 2, 200, 333, False, True, 72), (1, 20, 768, 768, True, False, 128)]:
        q = torch.randn(B, H, Tq, 128, device=dev).bfloat16()
        k = torch.randn(B, H, Tk + 7, 128, device=dev).bfloat16()
        v = torch.randn(B, H, Tk + 7, 128, device=dev).bfloat16()
        if out_hd != 128:
            q[..., out_hd:] = 0
            k[..., out_hd:] = 0
            v[..., out_hd:] = 0
        mask = None
        if masked:
            mask = (torch.rand(B, Tk, device=dev) < 0.3).to(torch.uint8)
        scale = out_hd ** -0.5
        o = ops.attention(q, k, v, Tq, Tk, scale, causal, out_hd=out_hd, key_mask=mask)
        torch.cuda.synchronize()
        ref = _attn_ref(q, k[:, :, :Tk], v[:, :, :Tk], scale, causal, mask)[..., :out_hd].reshape(B, Tq, H * out_hd)
        report(f"attn B{B} H{H} Tq{Tq} Tk{Tk} causal={causal} mask={masked} hd={out_hd}", o, ref, 2e-2)


def g_perf_attn():
    import torch
    from aria_b200 import ops
    dev = "cuda"
    for (B, H, T, causal) in [(1, 20, 768, True), (1, 20, 8192, True), (1, 16, 4900, False), (1, 20, 32768, True)]:
        q = torch.randn(B, H, T, 128, device=dev).bfloat16()
        k = torch.randn(B, H, T, 128, device=dev).bfloat16()
        v = torch.randn(B, H, T, 128, device=dev).bfloat16()
        ms = _time(lambda: ops.attention(q, k, v, T, T, 128 ** -0.5, causal), iters=5)
        fl = 4 * B * H * T * T * 128 * (0.5 if causal else 1.0)
        print(f"  attention B{B} H{H} T{T} causal={causal}: {ms:.3f} ms = {fl / ms / 1e9:.1f} TFLOP/s", flush=True)


def g_attn_decode():
    import torch
    from aria_b200 import ops
    torch.manual_seed(7)
    dev = "cuda"
    for (B, H, Tk) in [(2, 3, 100), (32, 20, 2048)]:
        q = torch.randn(B, H, 128, device=dev).bfloat16()
        k = torch.randn(B, H, Tk + 5, 128, device=dev).bfloat16()
        v = torch.randn(B, H, Tk + 5, 128, device=dev).bfloat16()
        o = ops.attention_decode(q, k, v, Tk, 128 ** -0.5)
        torch.cuda.synchronize()
        ref = _attn_ref(q[:, :, None], k[:, :, :Tk], v[:, :, :Tk], 128 ** -0.5, False).reshape(B, H * 128)
        report(f"decode B{B} H{H} Tk{Tk}", o, ref, 2e-2)


def _time(fn, iters=20, warm=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def g_perf_gemm():
    import torch
    from aria_b200 import ops, _lib as L
    dev = "cuda"
    torch.manual_seed(8)
    for (M, N, K) in [(4096, 4096, 4096), (8192, 8192, 8192), (4900, 4304, 1152), (4900, 1152, 4304), (4900, 1152, 1152),
                      (768, 2560, 2560), (768, 3328, 2560)]:
        x = torch.randn(M, K, device=dev).bfloat16()
        w = torch.randn(N, K, device=dev).bfloat16()
        b = torch.randn(N, device=dev).bfloat16()
        ms = _time(lambda: ops.linear(x, w))
        ms_g = _time(lambda: ops.linear(x, w, b, act=L.ACT_GELU_TANH))
        ms_t = _time(lambda: x @ w.t())
        print(f"  dense {M}x{N}x{K}: {ms:.3f} ms = {2 * M * N * K / ms / 1e9:.1f} TFLOP/s  (+bias+gelu {ms_g:.3f} ms)  (cuBLAS {ms_t:.3f} ms = {2 * M * N * K / ms_t / 1e9:.1f})", flush=True)
    E, d, I = 64, 2560, 1664
    fc1 = (torch.randn(E, d, 2 * I, device=dev) * 0.02).bfloat16()
    fc2 = (torch.randn(E, I, d, device=dev) * 0.02).bfloat16()
    for T in (768, 8192):
        rows = T * 6
        counts = torch.full((E,), rows // E, dtype=torch.int64)
        off = torch.tensor([0] + counts.cumsum(0).tolist(), dtype=torch.int32, device=dev)
        a = torch.randn(rows, d, device=dev).bfloat16()
        h = torch.randn(rows, I, device=dev).bfloat16()
        ms1 = _time(lambda: ops.grouped_gemm(a, fc1, off, swiglu=True))
        ms2 = _time(lambda: ops.grouped_gemm(h, fc2, off))
        fl1, fl2 = 2 * rows * d * 2 * I, 2 * rows * I * d
        print(f"  grouped T={T}: fc1+swiglu {ms1:.3f} ms ({fl1 / ms1 / 1e9:.1f} TFLOP/s)  fc2 {ms2:.3f} ms ({fl2 / ms2 / 1e9:.1f} TFLOP/s)", flush=True)


def g_perf():
    import torch
    from aria_b200 import ops
    dev = "cuda"
    torch.manual_seed(8)
    # dense GEMM
    for (M, N, K) in [(4096, 4096, 4096), (8192, 8192, 8192), (4900, 4304, 1152), (768, 2560, 2560)]:
        x = torch.randn(M, K, device=dev).bfloat16()
        w = torch.randn(N, K, device=dev).bfloat16()
        ms = _time(lambda: ops.linear(x, w))
        ms_t = _time(lambda: x @ w.t())
        print(f"  dense {M}x{N}x{K}: {ms:.3f} ms = {2 * M * N * K / ms / 1e9:.1f} TFLOP/s   (cuBLAS {ms_t:.3f} ms = {2 * M * N * K / ms_t / 1e9:.1f})", flush=True)
    # grouped expert GEMMs at the real width
    E, d, I = 64, 2560, 1664
    fc1 = (torch.randn(E, d, 2 * I, device=dev) * 0.02).bfloat16()
    fc2 = (torch.randn(E, I, d, device=dev) * 0.02).bfloat16()
    for T in (768, 8192):
        rows = T * 6
        counts = torch.full((E,), rows // E, dtype=torch.int64)
        off = torch.tensor([0] + counts.cumsum(0).tolist(), dtype=torch.int32, device=dev)
        a = torch.randn(rows, d, device=dev).bfloat16()
        h = torch.randn(rows, I, device=dev).bfloat16()
        ms1 = _time(lambda: ops.grouped_gemm(a, fc1, off, swiglu=True))
        ms2 = _time(lambda: ops.grouped_gemm(h, fc2, off))
        fl1, fl2 = 2 * rows * d * 2 * I, 2 * rows * I * d
        by1, by2 = fc1.numel() * 2, fc2.numel() * 2
        print(f"  grouped T={T}: fc1+swiglu {ms1:.3f} ms ({fl1 / ms1 / 1e9:.1f} TFLOP/s, weights {by1 / ms1 / 1e6:.0f} GB/s)  "
              f"fc2 {ms2:.3f} ms ({fl2 / ms2 / 1e9:.1f} TFLOP/s, weights {by2 / ms2 / 1e6:.0f} GB/s)", flush=True)
    # attention
    for (B, H, T) in [(1, 20, 768), (1, 20, 8192), (1, 16, 4900)]:
        q = torch.randn(B, H, T, 128, device=dev).bfloat16()
        k = torch.randn(B, H, T, 128, device=dev).bfloat16()
        v = torch.randn(B, H, T, 128, device=dev).bfloat16()
        causal = H == 20
        ms = _time(lambda: ops.attention(q, k, v, T, T, 128 ** -0.5, causal), iters=5)
        fl = 4 * B * H * T * T * 128 * (0.5 if causal else 1.0)
        print(f"  attention B{B} H{H} T{T} causal={causal}: {ms:.3f} ms = {fl / ms / 1e9:.1f} TFLOP/s", flush=True)


def main():
    if len(sys.argv) > 1:
        name = sys.argv[1]
        print(f"== {name}", flush=True)
        globals()["g_" + name]()
        return
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "bringup.log"), "w")
    for g in GROUPS:
        t = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), g], capture_output=True, text=True, timeout=600)
            out = r.stdout + ("\n[stderr]\n" + r.stderr[-3000:] if r.returncode else "")
            out += f"\n[{g}: exit {r.returncode}, {time.time() - t:.1f}s]\n"
        except subprocess.TimeoutExpired as e:
            out = f"== {g}\nTIMEOUT\n{(e.stdout or b'').decode()[-2000:] if isinstance(e.stdout, bytes) else e.stdout}\n"
        print(out, flush=True)
        log.write(out)
        log.flush()


if __name__ == "__main__":
    main()
