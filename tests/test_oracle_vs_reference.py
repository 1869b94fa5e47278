"""CPU, build container only: run the UNMODIFIED reference (/root/reference) live next to the oracle.
Skipped on the GPU box, where the reference tree does not exist (the golden fixtures cover it there)."""
import pytest
import torch

from oracle import aria_oracle as O
from oracle import configs as C
from oracle.ref_loader import load_reference, reference_available

pytestmark = pytest.mark.skipif(not reference_available(), reason="/root/reference not present")
torch.set_grad_enabled(False)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("T,E,k", [(1, 8, 2), (7, 8, 2), (64, 16, 6), (33, 64, 6)])
def test_moe_layer_live(dtype, T, E, k):
    ref = load_reference()
    tc = dict(hidden_size=128, moe_num_experts=E, moe_topk=k, moe_intermediate_size=64, moe_num_shared_experts=2)
    gen = torch.Generator().manual_seed(T * 131 + E)
    sd = {n: v.to(dtype) for n, v in C.moe_layer_state(tc, gen).items()}
    x = torch.randn(1, T, 128, generator=gen).to(dtype)
    layer = ref.moe_lm.MoELayer(ref.moe_lm.AriaMoELMConfig(**tc))
    layer.load_state_dict(sd, strict=True)
    layer = layer.to(dtype).eval()
    want = layer(x)
    got, parts = O.moe_layer(x, sd, k, return_parts=True)
    _, ridx, _ = layer.router(x)
    if torch.equal(ridx.sort(1).values, parts["top_idx"].sort(1).values):
        tol = 1e-6 if dtype == torch.float32 else 0.0
        assert (want.float() - got.float()).abs().max() <= tol
    else:  # a bf16 tie resolved differently by torch.topk: only tied rows may differ
        vals = parts["logits"].float().sort(1, descending=True).values
        tied = vals[:, k - 1] == vals[:, k]
        bad = (want.float() - got.float()).abs().amax(-1).view(-1) > 0
        assert not (bad & ~tied).any()


def test_installer_rebinds_reference_seams():
    """aria_b200.install.install() patches the reference MoELayer.forward and the `experts_gemm` global (no GPU needed to
    check the wiring; executing the patched path needs CUDA)."""
    from aria_b200 import install, moe_lm as ours
    ref = load_reference()
    cfg = ref.moe_lm.AriaMoELMConfig(hidden_size=128, moe_num_experts=8, moe_topk=2, moe_intermediate_size=64,
                                     moe_num_shared_experts=2, num_hidden_layers=2, num_attention_heads=1, vocab_size=64)
    model = ref.moe_lm.AriaMoELMForCausalLM(cfg)
    saved = ref.moe_lm.experts_gemm
    try:
        assert install.install(model, ref.moe_lm) == 2
        assert ref.moe_lm.experts_gemm is ours.experts_gemm
        layer = model.model.layers[0].mlp
        assert layer.forward.__func__ is install._moe_forward
        with pytest.raises(RuntimeError):   # no CPU fallback
            layer(torch.zeros(1, 4, 128, dtype=torch.bfloat16))
    finally:
        ref.moe_lm.experts_gemm = saved


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("scale", [1.0, 64.0])
def test_moe_layer_training_losses_backward_live(dtype, scale):
    """Training-mode routing (z-loss + load-balancing loss injected through MoEAuxLossAutoScaler, moe_lm.py:84-166,
    203-241): gradients of the unmodified reference vs the oracle's restatement, and vs the closed form the CUDA kernel
    implements (O.router_loss_grad)."""
    ref = load_reference()
    T, E, k, d = 48, 16, 4, 128
    # large coefficients so the loss terms are well above bf16 noise of the main gradient
    tc = dict(hidden_size=d, moe_num_experts=E, moe_topk=k, moe_intermediate_size=64, moe_num_shared_experts=2,
              moe_z_loss_coeff=0.5, moe_aux_loss_coeff=2.0)
    gen = torch.Generator().manual_seed(5)
    sd = {n: v.to(dtype) for n, v in C.moe_layer_state(tc, gen).items()}
    sd["router.weight"] = (sd["router.weight"].float() * 20).to(dtype)   # spread the logits
    x0 = torch.randn(1, T, d, generator=gen).to(dtype)
    dout = torch.randn(1, T, d, generator=gen).to(dtype)
    with torch.enable_grad():
        layer = ref.moe_lm.MoELayer(ref.moe_lm.AriaMoELMConfig(**tc))
        layer.load_state_dict(sd, strict=True)
        layer = layer.to(dtype).train()
        ref.moe_lm.MoEAuxLossAutoScaler.set_loss_scale(torch.tensor(scale))
        try:
            xr = x0.clone().requires_grad_(True)
            layer(xr).backward(dout)
        finally:
            ref.moe_lm.MoEAuxLossAutoScaler.set_loss_scale(torch.tensor(1.0))
        want = {n: p.grad for n, p in layer.named_parameters()}
        w = {n: v.clone().requires_grad_(True) for n, v in sd.items()}
        xo = x0.clone().requires_grad_(True)
        O._LossGradInjector.scale = scale
        try:
            O.moe_layer(xo, w, k, loss_coeffs=(0.5, 2.0)).backward(dout)
        finally:
            O._LossGradInjector.scale = 1.0
        # eval-mode oracle (no losses): the difference of the router gradients is the loss contribution
        w0 = {n: v.clone().requires_grad_(True) for n, v in sd.items()}
        x1 = x0.clone().requires_grad_(True)
        O.moe_layer(x1, w0, k).backward(dout)
    tol = 1e-6 if dtype == torch.float32 else 0.0
    for n in want:
        assert (want[n].float() - w[n].grad.float()).abs().max() <= tol * max(1.0, float(want[n].float().abs().max())), n
    # dx sums three branches (router, dispatch, shared expert); autograd's bf16 accumulation order is an engine detail,
    # so in bf16 it is only checked to an ulp-level tolerance (the parameter gradients above are bit-exact)
    xtol = 1e-6 if dtype == torch.float32 else 1e-2
    assert (xr.grad.float() - xo.grad.float()).abs().max() <= xtol * max(1.0, float(xr.grad.float().abs().max()))
    # closed form: d_router(train) - d_router(eval) == router_loss_grad^T @ x
    x2 = x0.reshape(T, d)
    logits = O.router_gating(x2, sd["router.weight"])
    _, _, counts = O.router_routing(logits, k)
    dl = O.router_loss_grad(logits, counts, k, 0.5, 2.0, scale)
    assert float(dl.abs().max()) > 0
    delta = w["router.weight"].grad.float() - w0["router.weight"].grad.float()
    closed = dl.t() @ x2.float()
    rel = (delta - closed).abs().max() / closed.abs().max()
    assert rel <= (1e-4 if dtype == torch.float32 else 4e-2), float(rel)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_grouped_gemm_lora_layer_live(dtype):
    """LoRA on the grouped GEMM: the unmodified `aria/lora/layers.py` (loaded under a stand-in for the two peft symbols it
    imports, see oracle/ref_loader.py) vs the oracle's restatement — forward and the gradients of A, B and x."""
    from oracle.ref_loader import load_reference_lora
    ref = load_reference()
    L = load_reference_lora()
    g = torch.Generator().manual_seed(4)
    E, K, N, r, alpha = 4, 64, 96, 8, 32
    counts = torch.tensor([16, 0, 40, 24])
    rows = int(counts.sum())
    w = (torch.randn(E, K, N, generator=g) * 0.05).to(dtype)
    a = (torch.randn(E, K, r, generator=g) * 0.05).to(dtype)
    b = (torch.randn(E, r, N, generator=g) * 0.05).to(dtype)
    x0 = torch.randn(rows, K, generator=g).to(dtype)
    dy = torch.randn(rows, N, generator=g).to(dtype)
    with torch.enable_grad():
        base = ref.moe_lm.GroupedGEMM(K, N, E)
        layer = L.GroupedGemmLoraLayer(base, "default", r=r, lora_alpha=alpha).to(dtype)
        assert layer.scaling["default"] == alpha / r
        with torch.no_grad():
            base.weight.copy_(w)
            layer.lora_A["default"].weight.copy_(a)
            layer.lora_B["default"].weight.copy_(b)
        xr = x0.clone().requires_grad_(True)
        want = layer(xr, counts)
        want.backward(dy)
        ao, bo, xo = a.clone().requires_grad_(True), b.clone().requires_grad_(True), x0.clone().requires_grad_(True)
        got = O.grouped_gemm_lora(xo, w, ao, bo, counts, alpha / r)
        got.backward(dy)
    assert torch.equal(got.detach(), want.detach())
    assert torch.equal(ao.grad, layer.lora_A["default"].weight.grad)
    assert torch.equal(bo.grad, layer.lora_B["default"].weight.grad)
    tol = 1e-6 if dtype == torch.float32 else 1e-2   # dx sums two autograd branches (accumulation order, see above)
    assert (xo.grad.float() - xr.grad.float()).abs().max() <= tol * float(xr.grad.float().abs().max())
    # merged weights (layers.py:154-213: W += A @ B * scaling) give the same function as the adapter path
    if dtype == torch.float32:
        with torch.no_grad():
            layer.merge()
            merged = layer(x0, counts)
        assert (merged - want.detach()).abs().max() <= 1e-5
        assert (base.weight - (w + torch.matmul(a, b) * (alpha / r))).abs().max() <= 1e-7


def test_install_vit_rebinds_the_reference_vision_layers():
    """Seam 3 wiring (executing the patched layers needs CUDA): every Idefics2EncoderLayer of the reference vision tower is
    patched, the tuple/tensor return convention of the host transformers version is detected, and CPU input fails loudly."""
    from aria_b200 import install
    ref = load_reference()
    vc = ref.vision_encoder.AriaVisionConfig(hidden_size=144, num_attention_heads=2, num_hidden_layers=2, intermediate_size=256,
                                             patch_size=14, image_size=56, hidden_act="gelu_pytorch_tanh")
    vc._attn_implementation = "eager"
    tower = ref.vision_encoder.AriaVisionModel(vc)
    assert install.install_vit(tower) == 2
    layer = tower.vision_model.encoder.layers[0]
    assert layer.forward.__func__ is install._vit_layer_forward and layer._aria_returns_tuple is False   # transformers 5.x here
    with pytest.raises(RuntimeError):
        layer(torch.zeros(1, 16, 144, dtype=torch.bfloat16), None)
