"""Stage-by-stage comparison of the tiny model on GPU vs the oracle (CPU)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F
from aria_b200 import ops
from aria_b200.modeling_aria import AriaConfig, AriaForConditionalGeneration
from oracle import aria_oracle as O, configs as C

torch.set_grad_enabled(False)
DEV = "cuda"
cfg = C.TINY
sd = C.aria_state(cfg, seed=0, dtype=torch.bfloat16)
m = AriaForConditionalGeneration(AriaConfig.from_dict(cfg), device=DEV)
m.load_state_dict({k: v.to(DEV) for k, v in sd.items()}, strict=True)
gold = torch.load(os.path.join(ROOT, "tests/golden/aria_tiny_bf16_full.pt"), weights_only=False)
ids, pv = gold["input_ids"], gold["pixel_values"]

def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))

tc = cfg["text_config"]
emb = F.embedding(ids, sd["language_model.model.embed_tokens.weight"])
feats, _ = O.vit_forward(pv, None, sd, cfg["vision_config"])
feats = O.projector_forward(feats, None, sd, cfg["projector"])
mask = (ids == cfg["image_token_index"]).unsqueeze(-1).expand_as(emb)
emb = emb.masked_scatter(mask, feats)

g_emb = ops.embedding(ids.to(DEV), m.get_input_embeddings().weight)
gf, _ = m.vision_tower(pv.to(DEV), None)
gf = m.multi_modal_projector(gf, None)
ops.merge_image_features(ids.to(DEV).reshape(-1), cfg["image_token_index"], gf.reshape(-1, gf.shape[-1]), g_emb.view(-1, g_emb.shape[-1]))
print("merged embeds rel", rel(g_emb, emb))

# feed the ORACLE's embeds to both, layer by layer
B, T, d = emb.shape
pos = torch.arange(T)[None]
lm = m.language_model
cache = lm.new_cache(B, T, DEV)
rope = lm.model.rope_tables(T, DEV)
x_o = emb
x_g, pending = emb.to(DEV), None
for i in range(tc["num_hidden_layers"]):
    p = f"language_model.model.layers.{i}."
    layer = lm.model.layers[i]
    # oracle pieces
    h_o = O.rms_norm(x_o, sd[p + "input_layernorm.weight"], tc["rms_norm_eps"])
    a_o, _ = O.lm_attention(h_o, sd, p + "self_attn.", tc["num_attention_heads"], tc["rope_theta"], pos)
    x1_o = x_o + a_o
    h2_o = O.rms_norm(x1_o, sd[p + "post_attention_layernorm.weight"], tc["rms_norm_eps"])
    mo_o, parts = O.moe_layer(h2_o, sd, tc["moe_topk"], p + "mlp.", return_parts=True)
    # gpu pieces fed with the oracle's inputs for isolation
    h_g = layer.input_layernorm(x_o.to(DEV))
    print(f"L{i} rms rel", rel(h_g, h_o))
    x1_g = layer.self_attn(h_o.to(DEV), cache, rope, residual=x_o.to(DEV))
    print(f"L{i} attn+res rel", rel(x1_g, x1_o))
    h2_g = layer.post_attention_layernorm(x1_o.to(DEV))
    print(f"L{i} rms2 rel", rel(h2_g, h2_o))
    sc, idx, cnt = layer.mlp.router(h2_o.to(DEV))
    same = (idx.cpu().long().sort(1).values == parts["top_idx"].sort(1).values).all(1)
    print(f"L{i} router same sets: {int(same.sum())}/{same.numel()}")
    if not same.all():
        bad = (~same).nonzero().flatten().tolist()
        lg = parts["logits"].float()
        for t in bad[:5]:
            print("   token", t, "oracle", parts["top_idx"][t].tolist(), "gpu", idx[t].tolist(), "sorted logits", lg[t].sort(descending=True).values[:4].tolist())
    mo_g = layer.mlp(h2_o.to(DEV))
    err = (mo_g.float().cpu() - mo_o.float()).abs().amax(-1).view(-1)
    print(f"L{i} moe rel", rel(mo_g, mo_o), "rows>1e-2:", (err > 1e-2 * mo_o.float().abs().max()).nonzero().flatten().tolist())
    x_o = x1_o + mo_o
# end-to-end chain on GPU
cache = lm.new_cache(B, T, DEV)
lg_g, _ = lm(emb.to(DEV), cache, 0)
lg_o, _ = O.lm_forward(emb, sd, tc)
err = (lg_g.float().cpu() - lg_o.float()).abs().amax(-1).view(-1)
print("LM logits rel", rel(lg_g, lg_o), "per-token max err", [round(float(e), 3) for e in err])
