"""A/B timing of the attention kernels (CUDA events, L2-cold-ish: 5 rotating input sets): round-1 kernel (attention_v2) vs the
current one, at the ViT / projector / LM shapes.  Env (read once per process): ARIA_ATTN_W=128 (padded tiles for hd 72),
ARIA_ATTN_PERSIST=0 (one unit per CTA)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from aria_b200 import ops
dev = "cuda"
torch.manual_seed(0)
tag = f"W={os.environ.get('ARIA_ATTN_W', '80')} persist={os.environ.get('ARIA_ATTN_PERSIST', '1')}"


def timeit(fn, sets, iters=20):
    for s in sets[:2]:
        fn(*s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(*sets[i % len(sets)])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3   # us


def run(name, B, H, Tq, Tk, hd, causal, nsets=3):
    sets = []
    for _ in range(nsets):
        q = torch.zeros(B, H, Tq, 128, device=dev, dtype=torch.bfloat16)
        k = torch.zeros(B, H, Tk, 128, device=dev, dtype=torch.bfloat16)
        v = torch.zeros(B, H, Tk, 128, device=dev, dtype=torch.bfloat16)
        for t in (q, k, v):
            t[..., :hd] = torch.randn(*t.shape[:-1], hd, device=dev).bfloat16()
        sets.append((q, k, v))
    pairs = Tq * Tk - (Tq * (Tq - 1) // 2 if causal else 0)
    fl = 4 * B * H * pairs * hd
    f3 = lambda q, k, v: ops.attention(q, k, v, Tq, Tk, hd ** -0.5, causal, out_hd=hd)
    f2 = lambda q, k, v: ops.attention_v2(q, k, v, Tq, Tk, hd ** -0.5, causal, out_hd=hd)
    a, b = f3(*sets[0]), f2(*sets[0])
    diff = float((a.float() - b.float()).abs().max() / b.float().abs().max())
    t3, t2 = timeit(f3, sets), timeit(f2, sets)
    print(f"{name:28s} [{tag}] v3 {t3:8.1f} us = {fl / t3 / 1e6:7.1f} TF/s | v2 {t2:8.1f} us = {fl / t2 / 1e6:7.1f} TF/s | "
          f"speedup {t2 / t3:5.2f}x | max|v3-v2|/max {diff:.2e}", flush=True)


run("ViT 16x4900x4900 hd72", 1, 16, 4900, 4900, 72, False)
run("ViT x4 images", 4, 16, 4900, 4900, 72, False, nsets=2)
run("projector 256 x 4900 hd72", 1, 16, 256, 4900, 72, False)
run("LM causal T=768 H=20", 1, 20, 768, 768, 128, True)
run("LM causal T=8192 H=20", 1, 20, 8192, 8192, 128, True)
if os.environ.get("LONG"):
    run("LM causal T=32768 H=20", 1, 20, 32768, 32768, 128, True, nsets=1)
