"""Per-op GPU time breakdown of one cfg-2 forward (CUDA events around every ops.* call), plus a CUDA-graph replay
timing of the same forward.  Writes gpurun_out/step_breakdown.txt."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from aria_b200 import ops, _lib as L
from aria_b200.modeling_aria import AriaConfig, AriaForConditionalGeneration, init_random_
from aria_b200 import configs as C

torch.set_grad_enabled(False)
dev = torch.device("cuda", 0)
lm_layers = int(os.environ.get("LM_LAYERS", "28"))
vit_layers = int(os.environ.get("VIT_LAYERS", "27"))
cfg = C.with_layers(C.ARIA_25B, lm_layers, vit_layers)
model = AriaForConditionalGeneration(AriaConfig.from_dict(cfg), device=dev)
init_random_(model, 0)
g = torch.Generator().manual_seed(1)
pv = torch.randn(1, 3, 980, 980, generator=g).bfloat16().to(dev)
text = torch.randint(10, 100352, (512,), generator=g)
ids_h = torch.cat([text[:16], torch.full((256,), cfg["image_token_index"]), text[16:]])[None].contiguous()
ids = ids_h.to(dev)

def fwd():
    return model(ids, pv, None, num_logits_to_keep=1, input_ids_host=ids_h).logits

for _ in range(3):
    fwd()
torch.cuda.synchronize()

events = []
names = [n for n in dir(ops) if callable(getattr(ops, n)) and not n.startswith("_") and getattr(getattr(ops, n), "__module__", "") == ops.__name__]
orig = {n: getattr(ops, n) for n in names}
stack = []
def wrap(n, f):
    def w(*a, **k):
        if stack:  # nested (e.g. router_topk -> none nested in python) keep outer only
            return f(*a, **k)
        stack.append(n)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = f(*a, **k); e1.record()
        stack.pop()
        tag = n
        if n == "grouped_gemm":
            tag += "_swiglu" if k.get("swiglu") else "_plain"
        if n == "linear":
            tag += f"[{a[0].reshape(-1, a[0].shape[-1]).shape[0]}x{a[1].shape[0]}x{a[1].shape[1]}]"
        if n == "attention":
            tag += f"[Tq{a[3]},Tk{a[4]},H{a[0].shape[1]}]"
        if n == "qkv_heads":
            tag += f"[{a[0].reshape(-1, a[0].shape[-1]).shape[0]}x{a[1][0].shape[0]}x{len(a[1])}]"
        events.append((tag, e0, e1))
        return r
    return w
for n in names:
    setattr(ops, n, wrap(n, orig[n]))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
import time
t0 = time.perf_counter(); e0.record(); fwd(); e1.record(); torch.cuda.synchronize(); wall = time.perf_counter() - t0
for n in names:
    setattr(ops, n, orig[n])
agg = collections.OrderedDict()
for tag, a, b in events:
    t = a.elapsed_time(b)
    c = agg.setdefault(tag, [0, 0.0]); c[0] += 1; c[1] += t
lines = [f"forward with per-op events: gpu {e0.elapsed_time(e1):.2f} ms, wall {wall*1e3:.2f} ms, ops {len(events)}"]
tot = sum(v[1] for v in agg.values())
for tag, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"  {t:8.3f} ms {100*t/tot:5.1f}%  x{n:<4d} avg {1e3*t/n:8.1f} us  {tag}")
lines.append(f"  sum of op times {tot:.2f} ms")

# plain timing + CUDA graph
def timeit(f, n=5):
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
lines.append(f"eager forward: {timeit(fwd):.2f} ms")
try:
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fwd()
    torch.cuda.current_stream().wait_stream(s)
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        out = fwd()
    lines.append(f"cuda-graph replay: {timeit(gr.replay):.2f} ms")
    ref = fwd()
    gr.replay(); torch.cuda.synchronize()
    lines.append(f"graph output equals eager: {bool(torch.equal(out, ref))}")
except Exception as ex:
    lines.append(f"cuda graph failed: {type(ex).__name__}: {ex}")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "step_breakdown.txt"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
