"""Model dimensions of the path (plain dicts).  Aria-25.3B: LM dims from gptfast/model.py:39-54, ViT dims from
gptfast/model.py:539-551, projector query counts from aria/model/configuration_aria.py:63-66."""
import copy

ARIA_25B = dict(
    image_token_index=9,
    text_config=dict(hidden_size=2560, num_attention_heads=20, num_hidden_layers=28, moe_num_experts=64,
                     moe_topk=6, moe_intermediate_size=1664, moe_num_shared_experts=2, vocab_size=100352,
                     rms_norm_eps=1e-5, rope_theta=5e6),
    vision_config=dict(hidden_size=1152, num_attention_heads=16, num_hidden_layers=27, intermediate_size=4304,
                       patch_size=14, image_size=980, layer_norm_eps=1e-6, num_channels=3),
    projector=dict(embed_dim=1152, num_heads=16, kv_dim=1152, ff_dim=2560, output_dim=2560,
                   patch_to_query_dict={1225: 128, 4900: 256}),
)


def with_layers(cfg, lm_layers=None, vit_layers=None):
    c = copy.deepcopy(cfg)
    if lm_layers is not None:
        c["text_config"]["num_hidden_layers"] = lm_layers
    if vit_layers is not None:
        c["vision_config"]["num_hidden_layers"] = vit_layers
    return c
