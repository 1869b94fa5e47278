"""Seam 2: the attention core registered under transformers' attention interface.
CPU: registration wiring and loud failures.  GPU: parity of the registered function against HF's own eager attention on the
same (already rotated) q/k/v — prefill, chunked prefill over a cache, decode."""
import pytest
import torch

from aria_b200 import hf_attention


class _Stub(torch.nn.Module):
    is_causal = True
    num_key_value_groups = 1
    training = False


def test_register_is_idempotent_and_visible():
    from transformers.modeling_utils import ALL_ATTENTION_FUNCTIONS
    key = hf_attention.register()
    assert key == "aria_b200" and hf_attention.register() == key
    assert ALL_ATTENTION_FUNCTIONS[key] is hf_attention.aria_b200_attention_forward
    try:
        from transformers.masking_utils import ALL_MASK_ATTENTION_FUNCTIONS
    except ImportError:
        return
    assert key in ALL_MASK_ATTENTION_FUNCTIONS.valid_keys()


def test_unsupported_arguments_fail_loudly():
    q = torch.zeros(1, 2, 4, 128, dtype=torch.bfloat16)
    f = hf_attention.aria_b200_attention_forward
    with pytest.raises(NotImplementedError):
        f(_Stub(), q, q, q, None, dropout=0.1)
    with pytest.raises(NotImplementedError):
        f(_Stub(), q, q, q, torch.ones(1, 1, 4, dtype=torch.bool))     # neither the 2-D padding mask nor a 4-D mask
    with pytest.raises(NotImplementedError):
        f(_Stub(), q, q[:, :1], q[:, :1], None)
    with pytest.raises(NotImplementedError):
        f(_Stub(), q[..., :64], q[..., :64], q[..., :64], None)
    with pytest.raises(RuntimeError):          # CPU tensors: there is no CPU path
        f(_Stub(), q, q, q, None)


@pytest.mark.gpu
@pytest.mark.parametrize("Tq,Tk", [(200, 200), (64, 333), (1, 517)])
def test_registered_core_matches_hf_eager(Tq, Tk):
    from transformers.models.llama.modeling_llama import eager_attention_forward
    g = torch.Generator().manual_seed(Tq + Tk)
    B, H, hd = 2, 3, 128
    q = torch.randn(B, H, Tq, hd, generator=g).bfloat16().cuda()
    k = torch.randn(B, H, Tk, hd, generator=g).bfloat16().cuda()
    v = torch.randn(B, H, Tk, hd, generator=g).bfloat16().cuda()
    scaling = hd ** -0.5
    # additive causal mask for the eager reference: query i sits at absolute position Tk - Tq + i
    qpos = torch.arange(Tk - Tq, Tk, device="cuda")[:, None]
    kpos = torch.arange(Tk, device="cuda")[None, :]
    mask = torch.zeros(Tq, Tk, device="cuda").masked_fill(kpos > qpos, float("-inf"))[None, None]
    want, _ = eager_attention_forward(_Stub(), q.float(), k.float(), v.float(), mask, scaling=scaling)
    # HF hands the function transposed views: reproduce that (non-contiguous query)
    q_view = q.transpose(1, 2).contiguous().transpose(1, 2)
    got, w = hf_attention.aria_b200_attention_forward(_Stub(), q_view, k, v, None, scaling=scaling)
    assert w is None and got.shape == (B, Tq, H, hd)
    err = (got.float() - want).abs().max() / want.abs().max()
    assert float(err) <= 2e-2, float(err)
