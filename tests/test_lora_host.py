"""CPU: host-side contract of the LoRA mirror (aria/lora/layers.py:30-152) — parameter names/shapes as peft saves them,
argument validation, and that there is no CPU fallback."""
import pytest
import torch

from aria_b200 import lora, moe_lm


def test_lora_layer_state_dict_layout_and_init():
    base = moe_lm.GroupedGEMM(64, 96, 4)
    layer = lora.GroupedGemmLoraLayer(base, "default", r=8, lora_alpha=32)
    sd = layer.state_dict()
    assert set(sd) == {"base_layer.weight", "lora_A.default.weight", "lora_B.default.weight"}
    assert sd["lora_A.default.weight"].shape == (4, 64, 8)       # GroupedGEMM(in, r, groups).weight  (layers.py:87-89)
    assert sd["lora_B.default.weight"].shape == (4, 8, 96)       # GroupedGEMM(r, out, groups).weight (layers.py:90-92)
    assert layer.scaling["default"] == 4.0                        # lora_alpha / r (layers.py:93)
    assert float(sd["lora_B.default.weight"].float().abs().max()) == 0.0
    assert float(sd["lora_A.default.weight"].float().abs().max()) > 0.0
    trainable = {n for n, p in layer.named_parameters() if p.requires_grad}
    assert trainable == {"lora_A.default.weight", "lora_B.default.weight"}


def test_lora_layer_rejects_bad_arguments():
    base = moe_lm.GroupedGEMM(64, 96, 4)
    with pytest.raises(ValueError):   # same message as the reference (layers.py:74-77)
        lora.GroupedGemmLoraLayer(base, r=0)
    with pytest.raises(ValueError):
        lora.GroupedGemmLoraLayer(base, r=12)
    with pytest.raises(ValueError):
        lora.GroupedGemmLoraLayer(base, r=8, lora_dropout=0.1)


def test_lora_layer_has_no_cpu_path():
    base = moe_lm.GroupedGEMM(64, 96, 4)
    layer = lora.GroupedGemmLoraLayer(base, r=8)
    with pytest.raises(RuntimeError):
        layer(torch.zeros(16, 64, dtype=torch.bfloat16), torch.tensor([16, 0, 0, 0]))


def test_merge_and_unmerge_fold_the_adapter_into_the_base_weight():
    """layers.py:154-228: W += A @ B * scaling; unmerge restores W (fp32 here so the round trip is exact to rounding)."""
    base = moe_lm.GroupedGEMM(64, 96, 4)
    layer = lora.GroupedGemmLoraLayer(base, r=8, lora_alpha=32).float()
    with torch.no_grad():
        base.weight.normal_(0, 0.05)
        layer.lora_B["default"].weight.normal_(0, 0.05)
    w0 = base.weight.detach().clone()
    delta = torch.matmul(layer.lora_A["default"].weight, layer.lora_B["default"].weight) * 4.0
    layer.merge()
    assert layer.merged and torch.allclose(base.weight, w0 + delta, atol=1e-6)
    layer.merge()                                     # idempotent
    assert torch.allclose(base.weight, w0 + delta, atol=1e-6)
    layer.unmerge()
    assert not layer.merged and torch.allclose(base.weight, w0, atol=1e-6)


def test_target_module_selection_follows_the_reference_rule():
    """aria/lora/utils.py:29-64 (the reference pins it in tests/test_get_target_modules.py): substring match on qualified names,
    minus frozen towers and frozen LM layers."""
    names = ["vision_tower.vision_model.encoder.layers.0.mlp.fc1", "multi_modal_projector.ffn.linear_in",
             "language_model.model.layers.0.mlp.experts.fc1", "language_model.model.layers.0.mlp.experts.fc2",
             "language_model.model.layers.1.mlp.experts.fc1", "language_model.model.layers.1.self_attn.q_proj",
             "language_model.lm_head"]
    sel = lora.get_lora_target_modules(names, ["fc1", "fc2", "q_proj"], freeze_vit=True)
    assert sel == names[2:6]
    sel = lora.get_lora_target_modules(names, ["fc1"], freeze_llm_layers=[0])
    assert sel == [names[0], names[4]]
    assert lora.get_lora_target_modules(names, ["fc1"], freeze_vit=True, freeze_llm=True) == []
    assert lora.get_lora_target_modules(names, ["linear_in"], freeze_projector=True) == []


def test_inject_lora_wraps_only_grouped_gemms_and_keeps_checkpoint_names():
    cfg = moe_lm.AriaMoELMConfig(hidden_size=128, num_attention_heads=1, num_hidden_layers=2, vocab_size=32,
                                 moe_intermediate_size=32, moe_num_experts=4, moe_topk=2)
    model = torch.nn.Module()
    model.language_model = moe_lm.AriaMoELMForCausalLM(cfg)      # names as in AriaForConditionalGeneration
    names = [n for n, _ in model.named_modules()]
    targets = lora.get_lora_target_modules(names, ["experts.fc1", "experts.fc2", "q_proj"], freeze_llm_layers=[0])
    wrapped = lora.inject_lora(model, targets, r=8, lora_alpha=16)
    p1 = "language_model.model.layers.1.mlp.experts."
    assert wrapped == [p1 + "fc1", p1 + "fc2"]                                     # q_proj is not a GroupedGEMM
    sd = model.state_dict()
    assert p1 + "fc1.base_layer.weight" in sd                                      # peft's key layout for a wrapped module
    assert sd[p1 + "fc2.lora_B.default.weight"].shape == (4, 8, 128)
    assert "language_model.model.layers.0.mlp.experts.fc1.weight" in sd            # frozen layer untouched
    assert type(model.language_model.model.layers[0].mlp.experts.fc1) is moe_lm.GroupedGEMM
