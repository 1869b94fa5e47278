"""Per-op breakdown of one decode step (B=32, 2K KV, 28 layers)."""
import collections, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from aria_b200 import ops, configs as C
from aria_b200.modeling_aria import AriaConfig, AriaForConditionalGeneration, init_random_
torch.set_grad_enabled(False)
dev = torch.device("cuda", 0)
B, Tkv = 32, 2048
cfg = C.with_layers(C.ARIA_25B, 28, 1)
model = AriaForConditionalGeneration(AriaConfig.from_dict(cfg), device=dev)
init_random_(model, 0)
cache = model.language_model.new_cache(B, Tkv + 8, dev)
for t in cache.k + cache.v: t.normal_()
ids = torch.randint(10, 100352, (B, 1), device=dev)
def step():
    cache.seq_len = Tkv - 1
    return model(ids, past_key_values=cache, num_logits_to_keep=1).logits
for _ in range(3): step()
torch.cuda.synchronize()
events = []
names = [n for n in dir(ops) if callable(getattr(ops, n)) and not n.startswith("_") and getattr(getattr(ops, n), "__module__", "") == ops.__name__]
orig = {n: getattr(ops, n) for n in names}
def wrap(n, f):
    def w(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = f(*a, **k); e1.record()
        tag = n + ("_swiglu" if k.get("swiglu") else "")
        if n == "linear": tag += f"[{a[1].shape[0]}x{a[1].shape[1]}]"
        events.append((tag, e0, e1)); return r
    return w
for n in names: setattr(ops, n, wrap(n, orig[n]))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); step(); b.record(); torch.cuda.synchronize()
for n in names: setattr(ops, n, orig[n])
agg = collections.OrderedDict()
for t, x, y in events:
    c = agg.setdefault(t, [0, 0.0]); c[0] += 1; c[1] += x.elapsed_time(y)
print(f"decode step (eager, with events) {a.elapsed_time(b):.2f} ms")
for t, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {ms:7.3f} ms x{c:<3d} avg {1e3*ms/c:7.1f} us  {t}")
