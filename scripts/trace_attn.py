"""Cycle accounting of CTA 0 of the ViT-shape attention (library built with -DARIA_ATTN_TRACE=1, see csrc/attention_v3.cu)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from aria_b200 import _lib as L
lib = L.load()
dev = "cuda"
torch.manual_seed(0)
B, H, T, hd = 1, 16, 4900, 72
q = torch.zeros(B, H, T, 128, device=dev, dtype=torch.bfloat16)
k, v = torch.zeros_like(q), torch.zeros_like(q)
for t in (q, k, v):
    t[..., :hd] = torch.randn(B, H, T, hd, device=dev).bfloat16()
out = torch.empty(B, T, H * hd, device=dev, dtype=torch.bfloat16)
ws = torch.zeros(4_000_000, device=dev, dtype=torch.float32)
vp = lambda t: C.c_void_p(t.data_ptr())
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(3):
    rc = lib.aria_attention_fwd(vp(q), vp(k), vp(v), vp(out), None, B, H, T, T, q.stride(0), q.stride(1), k.stride(0), k.stride(1), hd,
                                hd ** -0.5, 0, vp(ws), 148 * 256 * 84 * 4, st)
    assert rc == 0
torch.cuda.synchronize()
tr = ws[3200000:3200000 + 80].cpu().tolist()
names = {0: ("producer", ["wait q_empty", "wait kv_empty", "issue/other"]),
         16: ("issuer tile0", ["wait k_full", "wait q_full", "wait o_empty", "wait v_full", "wait p_full", "issue/other"]),
         32: ("issuer tile1", ["wait k_full", "wait q_full", "wait o_empty", "wait v_full", "wait p_full", "issue/other"]),
         48: ("softmax tile0 (row 0)", ["wait s_full", "ld S + max", "wait token", "exp + P store", "wait o_full", "epilogue", "-", "other"]),
         64: ("softmax tile1 (row 0)", ["wait s_full", "ld S + max", "wait token", "exp + P store", "wait o_full", "epilogue", "-", "other"])}
print(f"[{os.environ.get('TAG', '')}] cycle accounting of CTA 0 (3 work items: 2 whole units of 39 key blocks + a 1/9 piece)")
for base, (who, labels) in names.items():
    tot = tr[base + 12]
    parts = ", ".join(f"{l} {tr[base + i] / 1e3:.1f}k" for i, l in enumerate(labels) if l != "-")
    print(f"  {who:24s} total {tot / 1e3:8.1f}k clk | {parts}")

pc = ws[3300000:3300000 + 4 * 148].view(torch.int32).cpu().view(148, 4).long()
g0, g1, clk, smid = pc[:, 0] & 0xffffffff, pc[:, 1] & 0xffffffff, pc[:, 2] & 0xffffffff, pc[:, 3]
t0 = int(g0.min())
dur = ((g1 - g0) & 0xffffffff).float() / 1e3
print(f"  per-CTA (148): start spread {float((g0 - t0).max()) / 1e3:.1f} us; duration us min {float(dur.min()):.1f} median {float(dur.median()):.1f} max {float(dur.max()):.1f}; "
      f"kernel span {float(((g1 - t0) & 0xffffffff).max()) / 1e3:.1f} us; clk/ns {float((clk.float() / (dur * 1e3)).median()):.3f}")
order = dur.argsort()
print("  slowest CTAs (cta, sm, us):", [(int(i), int(smid[i]), round(float(dur[i]), 1)) for i in order[-6:]], " fastest:", [(int(i), int(smid[i]), round(float(dur[i]), 1)) for i in order[:4]])
