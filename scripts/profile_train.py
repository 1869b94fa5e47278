"""Per-op breakdown of one full-width MoE layer forward+backward (single GPU, 8192 tokens)."""
import collections, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from aria_b200 import ops, moe_lm, moe_train
dev = "cuda"
d, E, k, I, T = 2560, 64, 6, 1664, int(os.environ.get("T_LOC", 8192))
cfg = moe_lm.AriaMoELMConfig(hidden_size=d, moe_num_experts=E, moe_topk=k, moe_intermediate_size=I, moe_num_shared_experts=2)
layer = moe_lm.MoELayer(cfg, device=dev)
for p in layer.parameters():
    p.data.normal_(0, 0.02); p.requires_grad_(True)
x = torch.randn(1, T, d, device=dev).bfloat16().requires_grad_(True)
go = torch.randn(1, T, d, device=dev).bfloat16()
def step():
    for p in list(layer.parameters()) + [x]: p.grad = None
    moe_train.moe_layer_train(layer, x).backward(go)
for _ in range(3): step()
torch.cuda.synchronize()
events = []
names = [n for n in dir(ops) if callable(getattr(ops, n)) and not n.startswith("_") and getattr(getattr(ops, n), "__module__", "") == ops.__name__]
orig = {n: getattr(ops, n) for n in names}
depth = [0]
def wrap(n, f):
    def w(*a, **kw):
        if depth[0]: return f(*a, **kw)
        depth[0] += 1
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = f(*a, **kw); e1.record(); depth[0] -= 1
        tag = n
        if n in ("grouped_gemm", "grouped_gemm_nt", "grouped_wgrad", "matmul_kn", "linear", "linear_multi"):
            tag += str([tuple(t.shape) for t in a[:2] if hasattr(t, "shape")])
        events.append((tag, e0, e1)); return r
    return w
for n in names: setattr(ops, n, wrap(n, orig[n]))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); step(); b.record(); torch.cuda.synchronize()
for n in names: setattr(ops, n, orig[n])
agg = collections.OrderedDict()
for t, x0, x1 in events:
    c = agg.setdefault(t, [0, 0.0]); c[0] += 1; c[1] += x0.elapsed_time(x1)
print(f"fwd+bwd {a.elapsed_time(b):.2f} ms")
for t, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {ms:8.3f} ms x{c} {t}")
