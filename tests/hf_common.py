"""Test helper: a tiny *unmodified* `transformers.models.aria.AriaForConditionalGeneration` (the HF-native class, present in the
image's transformers 5.5 and therefore on the GPU box too) with seeded random weights.  The drop-in tests install our seams on
it and compare with the same model running HF's eager path."""
import torch


def tiny_hf_aria(device="cpu", dtype=torch.float32, seed=0, layers=2):
    from transformers.models.aria.configuration_aria import AriaConfig, AriaTextConfig
    from transformers.models.aria.modeling_aria import AriaForConditionalGeneration
    text = AriaTextConfig(vocab_size=512, hidden_size=256, intermediate_size=128, num_hidden_layers=layers, num_attention_heads=2,
                          num_key_value_heads=2, max_position_embeddings=4096, rms_norm_eps=1e-5,
                          rope_parameters={"rope_type": "default", "rope_theta": 5e6}, moe_num_experts=8, moe_topk=2,
                          moe_num_shared_experts=2, pad_token_id=0, bos_token_id=1, eos_token_id=2, head_dim=128)
    vision = dict(model_type="idefics3_vision", hidden_size=144, num_attention_heads=2, num_hidden_layers=2, intermediate_size=256,
                  patch_size=14, image_size=56, layer_norm_eps=1e-6, hidden_act="gelu_pytorch_tanh")
    cfg = AriaConfig(vision_config=vision, text_config=text, projector_patch_to_query_dict={16: 8, 4: 4}, image_token_index=9)
    cfg._attn_implementation = "eager"
    torch.manual_seed(seed)
    model = AriaForConditionalGeneration(cfg)
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if p.dim() == 1 and ("norm" in name or "ln_" in name) and name.endswith("weight"):
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
            else:
                p.copy_(0.02 * torch.randn(p.shape, generator=g))
    return model.to(device=device, dtype=dtype).eval()


def tiny_inputs(seed=3, n_text=20, batch=1, pad_last_image=False):
    g = torch.Generator().manual_seed(seed)
    pv = torch.randn(batch, 3, 56, 56, generator=g)
    pm = torch.ones(batch, 56, 56, dtype=torch.bool)
    if pad_last_image:
        pm[-1, 28:, :] = False          # bottom half padded -> 2x4 = 8 valid patches of 16
    text = torch.randint(10, 512, (batch, n_text), generator=g)
    ids = torch.cat([text[:, :4], torch.full((batch, 8), 9), text[:, 4:]], dim=1)   # 8 image tokens (16 patches -> 8 queries)
    return ids, pv, pm
