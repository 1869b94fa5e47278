"""TEST INFRASTRUCTURE — model configs (plain dicts) and deterministic random-init state dicts with the
Hugging Face parameter names of the reference (`language_model.model.layers.{i}.mlp.experts.fc1.weight`
… — gptfast/scripts/convert_hf_checkpoint.py:90-107 lists the checkpoint keys).

There are no pretrained weights offline, so every parity case uses seeded random init:
Linear / expert / router / embedding weights ~ N(0, 0.02^2), norms = 1 (+ small noise so the scale
actually matters in the tests), biases ~ N(0, 0.02^2).
"""
from __future__ import annotations

import copy
from collections import OrderedDict

import torch

# Real model dimensions live with the product (aria_b200/configs.py); the oracle only re-exports them.
from aria_b200.configs import ARIA_25B, with_layers  # noqa: E402,F401

# BASELINE.json configs[0]: single AriaMoE FFN block d=256, 8 experts, top-2 (I=512 chosen & recorded).
# The LM/ViT around it are shrunk but keep the real head dims (128 / 72) the kernels are written for.
TINY = dict(
    image_token_index=9,
    text_config=dict(hidden_size=256, num_attention_heads=2, num_hidden_layers=2, moe_num_experts=8,
                     moe_topk=2, moe_intermediate_size=512, moe_num_shared_experts=2, vocab_size=512,
                     rms_norm_eps=1e-5, rope_theta=5e6),
    vision_config=dict(hidden_size=144, num_attention_heads=2, num_hidden_layers=2, intermediate_size=256,
                       patch_size=14, image_size=56, layer_norm_eps=1e-6, num_channels=3),
    projector=dict(embed_dim=144, num_heads=2, kv_dim=144, ff_dim=256, output_dim=256,
                   patch_to_query_dict={16: 8, 4: 4}),
)


def _n(gen, shape, std=0.02):
    return torch.randn(shape, generator=gen, dtype=torch.float32) * std


def moe_layer_state(tcfg, gen, prefix=""):
    d, E, I, S = tcfg["hidden_size"], tcfg["moe_num_experts"], tcfg["moe_intermediate_size"], tcfg["moe_num_shared_experts"]
    sd = OrderedDict()
    sd[prefix + "router.weight"] = _n(gen, (E, d))
    sd[prefix + "experts.fc1.weight"] = _n(gen, (E, d, 2 * I))
    sd[prefix + "experts.fc2.weight"] = _n(gen, (E, I, d))
    sd[prefix + "shared_experts.gate_proj.weight"] = _n(gen, (I * S, d))
    sd[prefix + "shared_experts.up_proj.weight"] = _n(gen, (I * S, d))
    sd[prefix + "shared_experts.down_proj.weight"] = _n(gen, (d, I * S))
    return sd


def lm_state(tcfg, gen, prefix="language_model."):
    d, V = tcfg["hidden_size"], tcfg["vocab_size"]
    sd = OrderedDict()
    sd[prefix + "model.embed_tokens.weight"] = _n(gen, (V, d))
    for i in range(tcfg["num_hidden_layers"]):
        p = f"{prefix}model.layers.{i}."
        for n in ("q_proj", "k_proj", "v_proj", "o_proj"):
            sd[p + f"self_attn.{n}.weight"] = _n(gen, (d, d))
        sd.update(moe_layer_state(tcfg, gen, p + "mlp."))
        sd[p + "input_layernorm.weight"] = 1.0 + _n(gen, (d,), 0.1)
        sd[p + "post_attention_layernorm.weight"] = 1.0 + _n(gen, (d,), 0.1)
    sd[prefix + "model.norm.weight"] = 1.0 + _n(gen, (d,), 0.1)
    sd[prefix + "lm_head.weight"] = _n(gen, (V, d))
    return sd


def vit_state(vcfg, gen, prefix="vision_tower.vision_model."):
    d, I, P, C = vcfg["hidden_size"], vcfg["intermediate_size"], vcfg["patch_size"], vcfg["num_channels"]
    n_pos = (vcfg["image_size"] // P) ** 2
    sd = OrderedDict()
    sd[prefix + "embeddings.patch_embedding.weight"] = _n(gen, (d, C, P, P))
    sd[prefix + "embeddings.patch_embedding.bias"] = _n(gen, (d,))
    sd[prefix + "embeddings.position_embedding.weight"] = _n(gen, (n_pos, d))
    for i in range(vcfg["num_hidden_layers"]):
        p = f"{prefix}encoder.layers.{i}."
        for n in ("k_proj", "v_proj", "q_proj", "out_proj"):
            sd[p + f"self_attn.{n}.weight"] = _n(gen, (d, d))
            sd[p + f"self_attn.{n}.bias"] = _n(gen, (d,))
        sd[p + "layer_norm1.weight"] = 1.0 + _n(gen, (d,), 0.1)
        sd[p + "layer_norm1.bias"] = _n(gen, (d,))
        sd[p + "mlp.fc1.weight"] = _n(gen, (I, d))
        sd[p + "mlp.fc1.bias"] = _n(gen, (I,))
        sd[p + "mlp.fc2.weight"] = _n(gen, (d, I))
        sd[p + "mlp.fc2.bias"] = _n(gen, (d,))
        sd[p + "layer_norm2.weight"] = 1.0 + _n(gen, (d,), 0.1)
        sd[p + "layer_norm2.bias"] = _n(gen, (d,))
    return sd


def projector_state(pcfg, gen, prefix="multi_modal_projector."):
    E, kv, ff, out = pcfg["embed_dim"], pcfg["kv_dim"], pcfg["ff_dim"], pcfg["output_dim"]
    Q = max(pcfg["patch_to_query_dict"].values())
    sd = OrderedDict()
    sd[prefix + "query"] = _n(gen, (Q, E))
    ca = prefix + "cross_attn."
    sd[ca + "q_proj.weight"] = _n(gen, (E, E))
    sd[ca + "k_proj.weight"] = _n(gen, (E, kv))
    sd[ca + "v_proj.weight"] = _n(gen, (E, kv))
    sd[ca + "multihead_attn.in_proj_weight"] = _n(gen, (3 * E, E))
    sd[ca + "multihead_attn.in_proj_bias"] = _n(gen, (3 * E,))
    sd[ca + "multihead_attn.out_proj.weight"] = _n(gen, (E, E))
    sd[ca + "multihead_attn.out_proj.bias"] = _n(gen, (E,))
    sd[ca + "linear.weight"] = _n(gen, (E, E))
    sd[ca + "linear.bias"] = _n(gen, (E,))
    sd[ca + "layer_norm.weight"] = 1.0 + _n(gen, (E,), 0.1)
    sd[ca + "layer_norm.bias"] = _n(gen, (E,))
    sd[ca + "ln_kv.weight"] = 1.0 + _n(gen, (kv,), 0.1)
    sd[ca + "ln_kv.bias"] = _n(gen, (kv,))
    sd[prefix + "ln_ffn.weight"] = 1.0 + _n(gen, (E,), 0.1)
    sd[prefix + "ln_ffn.bias"] = _n(gen, (E,))
    sd[prefix + "ffn.linear_in.weight"] = _n(gen, (ff, E))
    sd[prefix + "ffn.linear_out.weight"] = _n(gen, (out, ff))
    return sd


def aria_state(cfg, seed=0, dtype=torch.float32):
    gen = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    sd.update(vit_state(cfg["vision_config"], gen))
    sd.update(projector_state(cfg["projector"], gen))
    sd.update(lm_state(cfg["text_config"], gen))
    return OrderedDict((k, v.to(dtype)) for k, v in sd.items())


def state_checksum(sd) -> float:
    """Order-dependent scalar fingerprint; golden files store it to detect RNG drift across torch versions."""
    acc = 0.0
    for i, (k, v) in enumerate(sd.items()):
        acc += (i + 1) * float(v.double().abs().sum())
    return acc
