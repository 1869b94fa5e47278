"""GPU: seam 3 — transformers' own Idefics2EncoderLayer (what the reference vision tower is built from) patched by
`aria_b200.install.install_vit` against the same layer run by HF in fp32 eager mode.  Passed on a B200 at the end of round 1."""
import pytest
import torch

pytestmark = pytest.mark.gpu

@pytest.mark.parametrize("padded", [False, True])
def test_install_vit_layer_matches_hf_eager(padded):
    """hd = 72, optional key padding handed over as the 4-D additive mask HF builds from the patch mask."""
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
