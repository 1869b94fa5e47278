"""Drop-in installer: rebind the hot-path seams of an *instantiated reference model* (`aria.model.*`, importable in a
transformers-4.46.3 environment or through a loader like oracle/ref_loader.py) to the B200-native kernels, sharing its
parameters (no copy, HF layout untouched).

Seams:
  1. `aria.model.moe_lm.experts_gemm` (moe_lm.py:431-443)  -> `aria_b200.moe_lm.experts_gemm` (gmm-compatible)
  2. `MoELayer.forward` (moe_lm.py:548-577)                 -> fused router / dispatch / grouped GEMM / combine path
  3. decoder-layer attention (moe_lm.py:594)                -> `AriaAttention`-style forward on the module's own q/k/v/o_proj
  4. `Idefics2EncoderLayer.forward` (vision_encoder.py:120) -> fused ViT layer

(1) and (2) are wired by `install()`; (3) is `aria_b200.hf_attention.register()` (an implementation key for transformers'
attention interface — the module keeps its projections, RoPE and HF Cache); (4) needs the mask plumbing of the host
transformers version and is exposed as the standalone mirror in `aria_b200.vision_encoder` (load the same state dict).
There is no CPU fallback: the patched modules require CUDA bf16 tensors.
"""
from __future__ import annotations

import types

import torch

from . import moe_lm as _m
from . import ops


def _moe_forward(self, hidden_states: torch.Tensor) -> torch.Tensor:
    """Replacement for the reference `MoELayer.forward` using the reference module's own parameters."""
    cfg = self.router.config
    shape = hidden_states.shape
    x = hidden_states.reshape(-1, shape[-1])
    scores, idx, counts, _ = ops.router_topk(x, self.router.weight, cfg.moe_topk)
    offsets, dest, src = ops.build_permutation(idx, counts)
    permuted = ops.permute_rows(x, src)
    h = ops.grouped_gemm(permuted, self.experts.fc1.weight, offsets, swiglu=True)
    y = ops.grouped_gemm(h, self.experts.fc2.weight, offsets)
    se = self.shared_experts
    shared = ops.linear(ops.linear_swiglu(x, se.gate_proj.weight, se.up_proj.weight), se.down_proj.weight)
    return ops.unpermute_combine(y, dest, scores, shared).view(shape)


def install(model, reference_moe_lm_module=None) -> int:
    """Patch every reference `MoELayer` inside `model` (and, if given, the reference module's global `experts_gemm`).
    Returns the number of layers patched.  Idempotent."""
    n = 0
    for mod in model.modules():
        if type(mod).__name__ == "MoELayer" and hasattr(mod, "router") and hasattr(mod, "experts") and hasattr(mod, "shared_experts"):
            mod.forward = types.MethodType(_moe_forward, mod)
            n += 1
    if reference_moe_lm_module is not None:
        reference_moe_lm_module.experts_gemm = _m.experts_gemm  # seam 1: GroupedGEMM.forward calls the module global
    return n
