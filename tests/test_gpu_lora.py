"""GPU: LoRA on the grouped expert GEMM (aria/lora/layers.py:30-152) — forward and the adapter / input
gradients of `aria_b200.lora.GroupedGemmLoraLayer` against fp32 autograd through the oracle's restatement."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


@pytest.mark.parametrize("r,alpha", [(8, 32), (16, 24)])   # scaling 4 (power of two: folded exactly) and 1.5
def test_grouped_gemm_lora_forward_backward_vs_oracle(r, alpha):
    from aria_b200 import lora, moe_lm
    from oracle import aria_oracle as O
    g = torch.Generator().manual_seed(r)
    E, K, N = 8, 256, 384
    counts = torch.tensor([32, 0, 80, 16, 48, 160, 16, 64])            # 16-row aligned groups (training dispatcher), one empty
    rows = int(counts.sum())
    x = torch.randn(rows, K, generator=g).bfloat16()
    w = (torch.randn(E, K, N, generator=g) * 0.05).bfloat16()
    a = (torch.randn(E, K, r, generator=g) * 0.05).bfloat16()
    b = (torch.randn(E, r, N, generator=g) * 0.05).bfloat16()
    dy = torch.randn(rows, N, generator=g).bfloat16()
    scaling = alpha / r
    # oracle: fp32 autograd on the bf16-rounded values
    x32, a32, b32 = (t.float().requires_grad_(True) for t in (x, a, b))
    with torch.enable_grad():
        want = O.grouped_gemm_lora(x32, w.float(), a32, b32, counts, scaling)
        want.backward(dy.float())
    base = moe_lm.GroupedGEMM(K, N, E, device=DEV)
    base.weight.data.copy_(w.to(DEV))
    layer = lora.GroupedGemmLoraLayer(base, "default", r=r, lora_alpha=alpha)
    assert layer.scaling["default"] == scaling
    assert float(layer.lora_B["default"].weight.abs().max()) == 0.0     # adapters start as a no-op
    layer.lora_A["default"].weight.data.copy_(a.to(DEV))
    layer.lora_B["default"].weight.data.copy_(b.to(DEV))
    assert not base.weight.requires_grad and layer.lora_A["default"].weight.requires_grad
    xg = x.to(DEV).requires_grad_(True)
    with torch.enable_grad():
        got = layer(xg, counts)                                        # counts as the reference passes them (CPU int64)
        got.backward(dy.to(DEV))
    assert _rel_l2(got.detach(), want.detach()) <= 1e-2
    assert _rel_l2(layer.lora_A["default"].weight.grad, a32.grad) <= 2e-2
    assert _rel_l2(layer.lora_B["default"].weight.grad, b32.grad) <= 2e-2
    assert _rel_l2(xg.grad, x32.grad) <= 2e-2
    assert base.weight.grad is None
    # the adapter term is visible in the output (otherwise the forward check proves nothing)
    with torch.no_grad():
        layer.disable_adapters = True
        plain = layer(x.to(DEV), counts)
    assert _rel_l2(plain, want.detach()) > 5e-3

