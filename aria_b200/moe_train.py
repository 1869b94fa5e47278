"""Training path of the MoE block (BASELINE cfg 5): `MoELayer` forward + backward as one torch.autograd.Function over our
CUDA kernels.  The reference gets its backward from autograd through `gmm` / ATen (aria/model/moe_lm.py:548-577); here every
gradient is an explicit kernel:

    forward (keeps h1, h, y):  router GEMM -> top-k/softmax -> 16-row-aligned stable permutation -> fc1 grouped GEMM ->
                               glu kernel -> fc2 grouped GEMM -> weighted combine (+ shared expert)
    backward:  combine_bwd (dy, dscores) -> fc2 wgrad (ragged contraction) + dgrad (weight read transposed in place)
               -> glu backward -> fc1 wgrad + dgrad -> un-permute sum -> shared-expert dgrad/wgrad -> top-k softmax backward
               -> router wgrad + dgrad

Router losses: with `router_losses=True` (the reference's `self.training` branch, moe_lm.py:257-258,271-272) the z-loss and
load-balancing-loss gradients (moe_lm.py:84-166) are added to dlogits by `aria_router_aux_bwd`, scaled by
`MoEAuxLossAutoScaler.main_loss_backward_scale`; the default (False) is eval-mode routing, which is what BASELINE cfg 5 times.
Gradients are bf16 tensors accumulated in fp32 inside the tensor-core kernels.
"""
from __future__ import annotations

import torch

from . import _lib as L
from . import ops


class MoELayerFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w_router, fc1, fc2, gate_w, up_w, down_w, topk: int, loss_coeffs=None):
        shape = x.shape
        x2 = x.reshape(-1, shape[-1]).contiguous()
        E = w_router.shape[0]
        I = fc2.shape[1]
        scores, idx, counts, _logits = ops.router_topk(x2, w_router, topk)
        offsets, dest, src = ops.build_permutation(idx, counts, row_align=16)
        xp = ops.permute_rows(x2, src)                       # [rows_pad, d], pad rows zero
        h1 = ops.grouped_gemm(xp, fc1, offsets)              # [rows_pad, 2I]
        h = ops.swiglu_fwd(h1)
        y = ops.grouped_gemm(h, fc2, offsets)                # [rows_pad, d]
        # shared expert: gate|up in one GEMM (two B segments), unfused glu so that the pre-activation is kept
        hs1 = ops.linear_multi(x2, [gate_w, up_w])
        hs = ops.swiglu_fwd(hs1)
        shared = ops.linear(hs, down_w)
        out = ops.unpermute_combine(y, dest, scores, shared)
        ctx.save_for_backward(x2, w_router, fc1, fc2, gate_w, up_w, down_w, scores, idx, offsets, dest, xp, h1, h, y, hs1, hs)
        ctx.topk = topk
        ctx.loss_coeffs = loss_coeffs
        if loss_coeffs is not None:  # keep what the loss gradients need: the bf16 logits and tokens_per_expert
            ctx.router_logits, ctx.counts = _logits, counts
        ctx.shape = shape
        return out.view(shape)

    @staticmethod
    def backward(ctx, dout):
        (x2, w_router, fc1, fc2, gate_w, up_w, down_w, scores, idx, offsets, dest, xp, h1, h, y, hs1, hs) = ctx.saved_tensors
        E, d = w_router.shape
        Is = gate_w.shape[0]
        T = x2.shape[0]
        do = dout.reshape(-1, d).contiguous()
        dense = torch.tensor([0, T], dtype=torch.int32, device=do.device)  # one 16-aligned "group" for dense wgrads
        # ---- routed experts
        dy, dscores = ops.combine_bwd(do, y, dest, scores)
        d_fc2 = ops.grouped_wgrad(h, dy, offsets)                             # [E, I, d]
        dh = ops.grouped_gemm_nt(dy, fc2, offsets)                            # dy @ fc2[e].T -> [rows, I]
        dh1 = ops.swiglu_bwd(h1, dh)
        d_fc1 = ops.grouped_wgrad(xp, dh1, offsets)                           # [E, d, 2I]
        dxp = ops.grouped_gemm_nt(dh1, fc1, offsets)                          # [rows, d]
        # ---- shared expert (out += shared: its upstream gradient is dout itself)
        d_down = ops.grouped_wgrad(do, hs, dense)[0]                          # [d, Is]
        dhs = ops.matmul_kn(do, down_w)                                       # do @ down_w  ([d, Is] read as K x N)
        dhs1 = ops.swiglu_bwd(hs1, dhs)                                       # [T, 2 Is] = [d gate | d up]
        d_gate = ops.grouped_wgrad(dhs1[:, :Is], x2, dense)[0]                # [Is, d]
        d_up = ops.grouped_wgrad(dhs1[:, Is:], x2, dense)[0]
        dx = ops.matmul_kn(dhs1[:, :Is], gate_w)
        dx = ops.matmul_kn(dhs1[:, Is:], up_w, residual=dx)
        # ---- router
        dlogits = ops.router_bwd(dscores, scores, idx, E)                     # [T, E]
        if ctx.loss_coeffs is not None:
            from .moe_lm import MoEAuxLossAutoScaler
            z_c, aux_c = ctx.loss_coeffs
            ops.router_aux_bwd(ctx.router_logits, ctx.counts, dlogits, ctx.topk, z_c, aux_c,
                               MoEAuxLossAutoScaler.main_loss_backward_scale)
        d_router = ops.grouped_wgrad(dlogits, x2, dense)[0]                   # [E, d]
        dx = ops.matmul_kn(dlogits, w_router, residual=dx)
        # ---- un-permute: dx[t] += sum_j dxp[dest[t, j]]
        ones = torch.ones_like(scores)
        dx = ops.unpermute_combine(dxp, dest, ones, dx)
        return dx.view(ctx.shape), d_router, d_fc1, d_fc2, d_gate, d_up, d_down, None, None


def moe_layer_train(layer, hidden_states: torch.Tensor, router_losses: bool = False) -> torch.Tensor:
    """Differentiable `MoELayer.forward` for an `aria_b200.moe_lm.MoELayer` whose parameters require grad.
    router_losses=True adds the training-mode z-loss / load-balancing-loss gradients (coefficients from the config)."""
    cfg = layer.router.config
    coeffs = (float(cfg.moe_z_loss_coeff), float(cfg.moe_aux_loss_coeff)) if router_losses else None
    return MoELayerFunction.apply(hidden_states, layer.router.weight, layer.experts.fc1.weight, layer.experts.fc2.weight,
                                  layer.shared_experts.gate_proj.weight, layer.shared_experts.up_proj.weight,
                                  layer.shared_experts.down_proj.weight, cfg.moe_topk, coeffs)
