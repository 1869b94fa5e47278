"""BASELINE cfg 3: Aria-25.3B decode, batch 32 at a 2K KV cache, one step (eager and CUDA-graph replay).
Reports ms/step, tokens/s and achieved HBM bytes/s against the ~66-68 GB/step streaming floor."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from aria_b200.modeling_aria import AriaConfig, AriaForConditionalGeneration, init_random_
from aria_b200 import configs as C
torch.set_grad_enabled(False)
dev = torch.device("cuda", 0)
B, Tkv = int(os.environ.get("B", 32)), int(os.environ.get("TKV", 2048))
cfg = C.with_layers(C.ARIA_25B, int(os.environ.get("LM_LAYERS", 28)), 1)
model = AriaForConditionalGeneration(AriaConfig.from_dict(cfg), device=dev)
init_random_(model, 0)
lm = model.language_model
cache = lm.new_cache(B, Tkv + 8, dev)
for t in cache.k + cache.v:
    t.normal_()
ids = torch.randint(10, 100352, (B, 1), device=dev)

def step():
    cache.seq_len = Tkv - 1     # the new token lands at position Tkv-1 -> attention over Tkv keys
    return model(ids, past_key_values=cache, num_logits_to_keep=1).logits

def timeit(f, n=10):
    for _ in range(3): f()
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): f()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
ms_eager = timeit(step)
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    step()
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    out = step()
ms_graph = timeit(g.replay)
tc = cfg["text_config"]; L_ = tc["num_hidden_layers"]
w_layer = (4 * 2560 * 2560 + 3 * 2560 * 3328 + 64 * 2560 + 64 * 3 * 2560 * 1664) * 2
expected_hit = 64 * (1 - (1 - 6 / 64) ** B)
w_layer_hit = (4 * 2560 * 2560 + 3 * 2560 * 3328 + 64 * 2560 + expected_hit * 3 * 2560 * 1664) * 2
kv = 2 * B * 20 * Tkv * 128 * 2
bytes_step = L_ * (w_layer_hit + kv) + 100352 * 2560 * 2
print(json.dumps({"bench": "decode_step", "B": B, "kv": Tkv, "layers": L_, "ms_eager": ms_eager, "ms_graph": ms_graph,
                  "tokens_per_s": B / ms_graph * 1e3, "algorithmic_GB_per_step": bytes_step / 1e9,
                  "achieved_GBps": bytes_step / ms_graph / 1e6, "finite": bool(torch.isfinite(out.float()).all())}))
