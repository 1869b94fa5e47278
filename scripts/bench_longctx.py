"""BASELINE cfg 4 (LM part): 64K-token causal prefill of the 28-layer MoE LM on one B200 (text embeddings only; the 32-frame
ViT pass is timed separately with VIT_FRAMES).  Prints time, tokens/s and the attention/MoE op breakdown."""
import collections, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from aria_b200 import ops
from aria_b200.modeling_aria import AriaConfig, AriaForConditionalGeneration, init_random_
from aria_b200 import configs as C
torch.set_grad_enabled(False)
dev = torch.device("cuda", 0)
T = int(os.environ.get("T", 65536)); layers = int(os.environ.get("LM_LAYERS", 28))
cfg = C.with_layers(C.ARIA_25B, layers, int(os.environ.get("VIT_LAYERS", 1)))
model = AriaForConditionalGeneration(AriaConfig.from_dict(cfg), device=dev)
init_random_(model, 0)
frames = int(os.environ.get("VIT_FRAMES", 0))
ids = torch.randint(10, 100352, (1, T), device=dev)
pv = None
if frames:  # cfg 4 proper: `frames` synthetic 980-px frames -> 256 image tokens each, the rest random text
    if model.vision_tower.config.num_hidden_layers == 1:
        pass
    ids[0, 64:64 + 256 * frames] = cfg["image_token_index"]
    pv = torch.randn(frames, 3, 980, 980, device=dev).bfloat16()
ids_h = ids.cpu()
def fwd():
    return model(ids, pv, None, num_logits_to_keep=1, input_ids_host=ids_h).logits
fwd(); torch.cuda.synchronize()
events = []
names = ["grouped_gemm", "attention", "linear", "linear_swiglu", "qkv_heads", "router_topk", "build_permutation", "permute_rows",
         "unpermute_combine", "rmsnorm", "embedding"]
orig = {n: getattr(ops, n) for n in names}
def wrap(n, f):
    def w(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = f(*a, **k); e1.record(); events.append((n + ("_swiglu" if k.get("swiglu") else ""), e0, e1)); return r
    return w
for n in names: setattr(ops, n, wrap(n, orig[n]))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); out = fwd(); b.record(); torch.cuda.synchronize()
for n in names: setattr(ops, n, orig[n])
ms = a.elapsed_time(b)
agg = collections.OrderedDict()
for n, x, y in events:
    c = agg.setdefault(n, [0, 0.0]); c[0] += 1; c[1] += x.elapsed_time(y)
attn_fl = layers * 5120 * T * T
moe_fl = layers * T * 153.35e6
res = {"bench": "longctx_prefill", "T": T, "layers": layers, "vit_frames": frames, "ms": ms, "tokens_per_s": T / ms * 1e3,
       "finite": bool(torch.isfinite(out.float()).all()), "mem_GB": torch.cuda.max_memory_allocated() / 1e9,
       "ops_ms": {k: round(v[1], 2) for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])},
       "attention_TFLOPs": attn_fl / agg["attention"][1] / 1e9,
       "expert_gemm_TFLOPs": moe_fl / (agg["grouped_gemm"][1] + agg["grouped_gemm_swiglu"][1]) / 1e9}
print(json.dumps(res))
