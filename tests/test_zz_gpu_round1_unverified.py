"""GPU tests written after the round's GPU-minutes were spent: they exercise only kernels and wrappers that the verified
tests already cover, in new combinations, but have NOT yet run on a B200.  They sit in the last-collected file and are
non-strict xfail so that an unexpected failure here is reported (xfail / XPASS) without masking the verified suite under
`pytest -x`.  Promote them into test_gpu_lora.py / test_gpu_ep.py once they have passed on hardware."""
import pytest
import torch

pytestmark = [pytest.mark.gpu,
              pytest.mark.xfail(reason="not yet run on hardware (round-1 GPU budget exhausted)", strict=False)]
DEV = "cuda"


def _rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-12))


def test_grouped_gemm_lora_against_reference_golden():
    """The bf16 fixture holds outputs and gradients of the UNMODIFIED reference layer (oracle/make_golden.py): compare the
    CUDA path with it directly (no oracle in between)."""
    import os
    from aria_b200 import lora, moe_lm
    g = torch.load(os.path.join(os.path.dirname(__file__), "golden", "lora_grouped_gemm_bf16.pt"), weights_only=False)
    E, K, N = g["w"].shape
    base = moe_lm.GroupedGEMM(K, N, E, device=DEV)
    base.weight.data.copy_(g["w"].to(DEV))
    layer = lora.GroupedGemmLoraLayer(base, "default", r=g["r"], lora_alpha=g["lora_alpha"])
    layer.lora_A["default"].weight.data.copy_(g["a"].to(DEV))
    layer.lora_B["default"].weight.data.copy_(g["b"].to(DEV))
    xg = g["x"].to(DEV).requires_grad_(True)
    with torch.enable_grad():
        out = layer(xg, g["counts"])
        out.backward(g["dy"].to(DEV))
    assert _rel_l2(out.detach(), g["out"]) <= 1e-2
    assert _rel_l2(layer.lora_A["default"].weight.grad, g["d_a"]) <= 2e-2
    assert _rel_l2(layer.lora_B["default"].weight.grad, g["d_b"]) <= 2e-2
    assert _rel_l2(xg.grad, g["dx"]) <= 2e-2
    # merged weights: same function through the plain grouped GEMM
    with torch.no_grad():
        layer.merge()
        merged = layer(g["x"].to(DEV), g["counts"])
    assert _rel_l2(merged, g["out"]) <= 2e-2


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_one_process_two_devices():
    """The reference can span GPUs inside ONE process (`device_map="auto"`, aria/inference.py:55-57; hence the
    `torch.cuda.set_device(input.device)` at moe_lm.py:483).  Every C-ABI call must honour the tensor's device: the
    dynamic-shared-memory opt-ins and the SM count are per-device state inside the library."""
    from aria_b200 import ops
    g = torch.Generator().manual_seed(0)
    x = torch.randn(300, 512, generator=g).bfloat16()
    w = (torch.randn(384, 512, generator=g) * 0.05).bfloat16()
    q = torch.randn(1, 2, 200, 128, generator=g).bfloat16()
    ref = x.float() @ w.float().t()
    outs = []
    for dev in ("cuda:0", "cuda:1", "cuda:0"):
        y = ops.linear(x.to(dev), w.to(dev))
        o = ops.attention(q.to(dev), q.to(dev), q.to(dev), 200, 200, 128 ** -0.5, True)
        torch.cuda.synchronize(dev)
        assert y.device == torch.device(dev)
        assert float((y.float().cpu() - ref).abs().max()) <= 2e-2 * float(ref.abs().max())
        outs.append(o.float().cpu())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])



def test_grouped_mlp_with_lora_adapters_forward_backward_vs_oracle():
    """`GroupedMLP` whose fc1 / fc2 were wrapped by `inject_lora` (what peft does to the reference, aria/train.py:107):
    fc1+adapter -> glu -> fc2+adapter, forward and the four adapter gradients + dx against fp32 autograd through the oracle."""
    from aria_b200 import lora, moe_lm
    from oracle import aria_oracle as O
    g = torch.Generator().manual_seed(21)
    E, d, I, r, alpha = 4, 128, 128, 8, 32
    cfg = moe_lm.AriaMoELMConfig(hidden_size=d, moe_num_experts=E, moe_topk=2, moe_intermediate_size=I)
    mlp = moe_lm.GroupedMLP(cfg, device=DEV)
    counts = torch.tensor([48, 16, 0, 64])
    rows = int(counts.sum())
    w1 = (torch.randn(E, d, 2 * I, generator=g) * 0.05).bfloat16()
    w2 = (torch.randn(E, I, d, generator=g) * 0.05).bfloat16()
    mlp.fc1.weight.data.copy_(w1.to(DEV))
    mlp.fc2.weight.data.copy_(w2.to(DEV))
    assert lora.inject_lora(mlp, ["fc1", "fc2"], r=r, lora_alpha=alpha) == ["fc1", "fc2"]
    ab = {}
    for name, K, N in (("fc1", d, 2 * I), ("fc2", I, d)):
        a = (torch.randn(E, K, r, generator=g) * 0.05).bfloat16()
        b = (torch.randn(E, r, N, generator=g) * 0.05).bfloat16()
        layer = getattr(mlp, name)
        layer.lora_A["default"].weight.data.copy_(a.to(DEV))
        layer.lora_B["default"].weight.data.copy_(b.to(DEV))
        ab[name] = (a.float().requires_grad_(True), b.float().requires_grad_(True))
    x = torch.randn(rows, d, generator=g).bfloat16()
    dy = torch.randn(rows, d, generator=g).bfloat16()
    x32 = x.float().requires_grad_(True)
    s = alpha / r
    with torch.enable_grad():
        h1 = O.grouped_gemm_lora(x32, w1.float(), *ab["fc1"], counts, s)
        want = O.grouped_gemm_lora(O.glu(h1), w2.float(), *ab["fc2"], counts, s)
        want.backward(dy.float())
    xg = x.to(DEV).requires_grad_(True)
    with torch.enable_grad():
        got = mlp(xg, counts)
        got.backward(dy.to(DEV))
    assert _rel_l2(got.detach(), want.detach()) <= 2e-2
    assert _rel_l2(xg.grad, x32.grad) <= 3e-2
    for name in ("fc1", "fc2"):
        layer = getattr(mlp, name)
        assert _rel_l2(layer.lora_A["default"].weight.grad, ab[name][0].grad) <= 3e-2, name
        assert _rel_l2(layer.lora_B["default"].weight.grad, ab[name][1].grad) <= 3e-2, name


@pytest.mark.parametrize("padded", [False, True])
def test_install_vit_layer_matches_hf_eager(padded):
    """Seam 3: transformers' own Idefics2EncoderLayer (what the reference vision tower is built from) patched by
    `install_vit` vs the same layer run by HF in fp32 eager mode; hd = 72, optional key padding as a 4-D additive mask."""
    import copy
    from transformers.models.idefics2.modeling_idefics2 import Idefics2EncoderLayer, Idefics2VisionConfig
    from aria_b200 import install
    torch.manual_seed(5)
    cfg = Idefics2VisionConfig(hidden_size=144, num_attention_heads=2, intermediate_size=256, num_hidden_layers=1,
                               hidden_act="gelu_pytorch_tanh")
    cfg._attn_implementation = "eager"
    ref_layer = Idefics2EncoderLayer(cfg).float().cuda().eval()
    for p_ in ref_layer.parameters():                       # bf16-representable weights so both sides see the same values
        p_.data = (torch.randn_like(p_) * 0.05).bfloat16().float()
    ref_layer.layer_norm1.weight.data += 1.0
    ref_layer.layer_norm2.weight.data += 1.0
    ours = copy.deepcopy(ref_layer).bfloat16()
    holder = torch.nn.ModuleList([ours])
    assert install.install_vit(holder) == 1
    B, N = 2, 200
    x = torch.randn(B, N, 144, device="cuda").bfloat16()
    mask = None
    if padded:
        valid = torch.ones(B, N, dtype=torch.bool, device="cuda")
        valid[0, 150:] = False
        mask = torch.zeros(B, 1, N, N, device="cuda").masked_fill(~valid[:, None, None, :], torch.finfo(torch.float32).min)
    with torch.no_grad():
        want = ref_layer(x.float(), mask)
        got = ours(x, mask)
    want = want[0] if isinstance(want, tuple) else want
    got = got[0] if isinstance(got, tuple) else got
    rows = slice(None) if not padded else (slice(None), slice(0, 150))   # padded query rows are don't-care downstream
    err = (got.float()[rows] - want[rows]).abs().max() / want[rows].abs().max()
    assert float(err) <= 2e-2, float(err)
