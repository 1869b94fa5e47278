"""Drop-in installer: rebind the hot-path seams of an *instantiated reference model* (`aria.model.*`, importable in a
transformers-4.46.3 environment or through a loader like oracle/ref_loader.py) to the B200-native kernels, sharing its
parameters (no copy, HF layout untouched).

Seams:
  1. `aria.model.moe_lm.experts_gemm` (moe_lm.py:431-443)  -> `aria_b200.moe_lm.experts_gemm` (gmm-compatible)
  2. `MoELayer.forward` (moe_lm.py:548-577)                 -> fused router / dispatch / grouped GEMM / combine path
  3. decoder-layer attention (moe_lm.py:594)                -> `AriaAttention`-style forward on the module's own q/k/v/o_proj
  4. `Idefics2EncoderLayer.forward` (vision_encoder.py:120) -> fused ViT layer

(1) and (2) are wired by `install()`; (3) is `aria_b200.hf_attention.register()` (an implementation key for transformers'
attention interface — the module keeps its projections, RoPE and HF Cache); (4) is `install_vit()` (per-layer forward on the
HF module's own parameters; the embeddings / mask creation around it stay HF's).
There is no CPU fallback: the patched modules require CUDA bf16 tensors.
"""
from __future__ import annotations

import types

import torch

from . import moe_lm as _m
from . import ops


def _reject_autograd(what: str, module, *tensors):
    """The seams below are the INFERENCE path: their outputs come straight from the C ABI and carry no grad_fn, so under
    autograd they would silently cut the graph (and drop the training-mode router losses, moe_lm.py:257-272).  Refuse
    instead of training with wrong gradients; fine-tuning goes through aria_b200.moe_train / aria_b200.lora."""
    if torch.is_grad_enabled() and (module.training or any(t is not None and t.requires_grad for t in tensors)):
        raise RuntimeError(f"aria_b200.install: {what} is inference-only (no autograd through the fused kernels); call it under "
                           "torch.no_grad() with the module in eval() mode, or use aria_b200.moe_train / aria_b200.lora to train")


def _moe_forward(self, hidden_states: torch.Tensor) -> torch.Tensor:
    """Replacement for the reference `MoELayer.forward` (moe_lm.py:548-577) — and for transformers' own `AriaTextMoELayer.forward`,
    which has the same sub-modules with `router` an nn.Linear — using the module's own parameters."""
    _reject_autograd("MoELayer.forward", self, hidden_states)
    cfg = getattr(self.router, "config", None) or self.config
    shape = hidden_states.shape
    x = hidden_states.reshape(-1, shape[-1]).contiguous()
    se = self.shared_experts
    if x.is_cuda:
        # one C-ABI call for the whole block (aria_moe_block_fwd, csrc/moe_block.cu); the shared experts run on a side stream
        out = ops.moe_block_fwd(x, self.router.weight, self.experts.fc1.weight, self.experts.fc2.weight, se.gate_proj.weight,
                                se.up_proj.weight, se.down_proj.weight, cfg.moe_topk, side_stream=_m._side_stream(x.device))
        return out.view(shape)
    # kernel-by-kernel sequence (what the block entry launches); reached only by the host-logic tests, which swap `ops` for
    # oracle-backed stand-ins on CPU (tests/standin_ops.py)
    forked = _m.shared_expert_overlapped(
        lambda: ops.linear(ops.linear_swiglu(x, se.gate_proj.weight, se.up_proj.weight), se.down_proj.weight), x)
    scores, idx, counts, _ = ops.router_topk(x, self.router.weight, cfg.moe_topk)
    offsets, dest, src = ops.build_permutation(idx, counts)
    permuted = ops.permute_rows(x, src)
    h = ops.grouped_gemm(permuted, self.experts.fc1.weight, offsets, swiglu=True)
    y = ops.grouped_gemm(h, self.experts.fc2.weight, offsets)
    shared = _m.join_side(forked, x)
    return ops.unpermute_combine(y, dest, scores, shared).view(shape)


def _vit_layer_forward(self, hidden_states: torch.Tensor, attention_mask=None, *args, **kwargs):
    """Replacement for transformers' `Idefics2EncoderLayer.forward` (what `AriaVisionTransformer` is built from,
    vision_encoder.py:65-67,120), on the module's own parameters: LN -> fused q/k/v GEMM with head scatter -> non-causal attention
    (hd 72 carried in 128-wide rows, key mask) -> out_proj (+residual) -> LN -> fc1 (+bias, gelu_tanh) -> fc2 (+bias, +residual).
    `attention_mask`: None, the 4-D additive mask [B,1,N,N] HF builds from the patch mask, or a 2-D validity mask [B,N]."""
    from . import _lib as L
    _reject_autograd("Idefics2EncoderLayer.forward", self, hidden_states)
    a, m = self.self_attn, self.mlp
    B, N, _ = hidden_states.shape
    H, hd = a.num_heads, a.head_dim
    key_mask = None
    if attention_mask is not None:
        if attention_mask.dim() == 4 and attention_mask.dtype == torch.bool:      # sdpa-style mask: True = may attend
            key_mask = (~attention_mask[:, 0, 0, :]).to(torch.uint8).contiguous()
        elif attention_mask.dim() == 4:
            key_mask = (attention_mask[:, 0, 0, :] < 0).to(torch.uint8).contiguous()
        else:
            key_mask = (~attention_mask.bool()).to(torch.uint8).contiguous()
    buf = getattr(self, "_aria_qkv", None)
    if buf is None or buf[0].shape != (B, H, N, 128) or buf[0].device != hidden_states.device:
        buf = [torch.zeros(B, H, N, 128, dtype=torch.bfloat16, device=hidden_states.device) for _ in range(3)]
        self._aria_qkv = buf            # pad columns hd..127 stay zero across calls
    x = hidden_states.contiguous()
    h = ops.layernorm(x, self.layer_norm1.weight, self.layer_norm1.bias, self.layer_norm1.eps)
    ops.qkv_heads(h, [a.q_proj.weight, a.k_proj.weight, a.v_proj.weight], [a.q_proj.bias, a.k_proj.bias, a.v_proj.bias],
                  buf, hd, N)
    o = ops.attention(buf[0], buf[1], buf[2], N, N, hd ** -0.5, causal=False, out_hd=hd, key_mask=key_mask)
    x = ops.linear(o, a.out_proj.weight, a.out_proj.bias, residual=x)
    h = ops.layernorm(x, self.layer_norm2.weight, self.layer_norm2.bias, self.layer_norm2.eps)
    h = ops.linear(h, m.fc1.weight, m.fc1.bias, act=L.ACT_GELU_TANH)
    out = ops.linear(h, m.fc2.weight, m.fc2.bias, residual=x)
    return (out,) if getattr(self, "_aria_returns_tuple", False) else out


def install_vit(model) -> int:
    """Seam 3: patch every `Idefics2EncoderLayer` inside `model` (the reference's vision tower) with `_vit_layer_forward`.
    transformers 4.46 layers return a 1-tuple, 5.x layers the tensor: detected from the original signature.  Returns the number
    of layers patched.  Requires `hidden_act = gelu_pytorch_tanh` (what Aria uses); anything else is rejected."""
    import inspect
    n = 0
    for mod in model.modules():
        # Idefics2EncoderLayer: the reference's tower (vision_encoder.py:26-28,65-67); Idefics3EncoderLayer: the identical layer
        # transformers' own `models.aria` builds its tower from
        if type(mod).__name__ in ("Idefics2EncoderLayer", "Idefics3EncoderLayer") and hasattr(mod, "self_attn") and hasattr(mod, "layer_norm1"):
            act = type(getattr(mod.mlp, "activation_fn", None)).__name__
            if act not in ("GELUTanh", "PytorchGELUTanh"):
                raise NotImplementedError(f"install_vit: unsupported MLP activation {act} (the fused epilogue is gelu_pytorch_tanh)")
            mod._aria_returns_tuple = "output_attentions" in inspect.signature(type(mod).forward).parameters
            mod.forward = types.MethodType(_vit_layer_forward, mod)
            n += 1
    return n


def install(model, reference_moe_lm_module=None) -> int:
    """Patch every reference `MoELayer` inside `model` (and, if given, the reference module's global `experts_gemm`).
    Returns the number of layers patched.  Idempotent."""
    n = 0
    for mod in model.modules():
        if type(mod).__name__ in ("MoELayer", "AriaTextMoELayer") and hasattr(mod, "router") and hasattr(mod, "experts") \
                and hasattr(mod, "shared_experts"):
            mod.forward = types.MethodType(_moe_forward, mod)
            n += 1
    if reference_moe_lm_module is not None:
        reference_moe_lm_module.experts_gemm = _m.experts_gemm  # seam 1: GroupedGEMM.forward calls the module global
    return n
