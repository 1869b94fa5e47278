"""Host-side mirror of `aria/model/projector.py` (AriaProjector: cross-attention of learned queries over the
ViT tokens + FFN) on the B200-native kernels.  Parameter names = HF checkpoint keys (incl. the
nn.MultiheadAttention `in_proj_weight` / `in_proj_bias` / `out_proj`).  SURVEY.md §8(f) "next #1".
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import _lib as L
from . import ops
from .moe_lm import Linear, _param, bf16
from .vision_encoder import LayerNorm


class _MultiheadAttention(nn.Module):
    """Parameter holder for torch.nn.MultiheadAttention(embed_dim, num_heads) (projector.py:66)."""

    def __init__(self, embed_dim, num_heads, device=None):
        super().__init__()
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.in_proj_weight = _param(3 * embed_dim, embed_dim, device=device)
        self.in_proj_bias = _param(3 * embed_dim, device=device)
        self.out_proj = Linear(embed_dim, embed_dim, bias=True, device=device)


class CrossAttention(nn.Module):
    """projector.py:48-102."""

    def __init__(self, kv_dim, embed_dim, num_heads, device=None):
        super().__init__()
        self.num_heads = num_heads
        self.q_proj = Linear(embed_dim, embed_dim, device=device)
        self.k_proj = Linear(kv_dim, embed_dim, device=device)
        self.v_proj = Linear(kv_dim, embed_dim, device=device)
        self.multihead_attn = _MultiheadAttention(embed_dim, num_heads, device)
        self.linear = Linear(embed_dim, embed_dim, bias=True, device=device)
        self.layer_norm = LayerNorm(embed_dim, 1e-5, device)
        self.ln_kv = LayerNorm(kv_dim, 1e-5, device)

    def forward(self, x, hidden_states, key_mask=None):
        """x [B,N,kv] (keys/values), hidden_states [B,Q,E] (queries), key_mask [B,N] uint8 (1 = masked out)."""
        B, N, _ = x.shape
        Q, E = hidden_states.shape[1], hidden_states.shape[2]
        H = self.num_heads
        hd = E // H
        mha = self.multihead_attn
        wi, bi = mha.in_proj_weight, mha.in_proj_bias
        query = ops.linear(self.layer_norm(hidden_states), self.q_proj.weight)
        xn = self.ln_kv(x)
        key = ops.linear(xn, self.k_proj.weight)
        value = ops.linear(xn, self.v_proj.weight)
        dev = x.device
        q2 = torch.zeros(B, H, Q, 128, dtype=bf16, device=dev)
        k2 = torch.zeros(B, H, N, 128, dtype=bf16, device=dev)
        v2 = torch.zeros(B, H, N, 128, dtype=bf16, device=dev)
        # nn.MultiheadAttention's own in-projection (second projection, with bias), scattered head-major
        ops.qkv_heads(query, [wi[:E]], [bi[:E]], [q2], hd, Q)
        ops.qkv_heads(key, [wi[E:2 * E]], [bi[E:2 * E]], [k2], hd, N)
        ops.qkv_heads(value, [wi[2 * E:]], [bi[2 * E:]], [v2], hd, N)
        o = ops.attention(q2, k2, v2, Q, N, hd ** -0.5, causal=False, out_hd=hd, key_mask=key_mask)
        o = ops.linear(o, mha.out_proj.weight, mha.out_proj.bias)
        return ops.linear(o, self.linear.weight, self.linear.bias)  # dropout p=0 (projector.py:67)


class FFN(nn.Module):
    """projector.py:27-45: linear_in -> gelu_new -> linear_out (no bias)."""

    def __init__(self, embed_dim, ff_dim, output_dim, device=None):
        super().__init__()
        self.linear_in = Linear(embed_dim, ff_dim, device=device)
        self.linear_out = Linear(ff_dim, output_dim, device=device)

    def forward(self, x):
        return ops.linear(ops.linear(x, self.linear_in.weight, act=L.ACT_GELU_NEW), self.linear_out.weight)


class AriaProjector(nn.Module):
    """projector.py:105-189.  forward(x [B,N,kv_dim], attn_mask [B,N] bool (True = padding) | None) -> [B,Q,out]."""

    def __init__(self, patch_to_query_dict, embed_dim, num_heads, kv_dim, ff_dim, output_dim, device=None):
        super().__init__()
        self.patch_to_query_dict = {int(k): int(v) for k, v in patch_to_query_dict.items()}
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.query = _param(max(self.patch_to_query_dict.values()), embed_dim, device=device)
        self.cross_attn = CrossAttention(kv_dim, embed_dim, num_heads, device)
        self.ln_ffn = LayerNorm(embed_dim, 1e-5, device)
        self.ffn = FFN(embed_dim, ff_dim, output_dim, device)

    def forward(self, x: torch.Tensor, attn_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        bs = x.shape[0]
        query_num = self.patch_to_query_dict.get(x.shape[1], None)
        assert query_num is not None, f"Query number for {x.shape[1]} patches is not provided"  # projector.py:174-177
        queries = self.query[:query_num].unsqueeze(0).repeat(bs, 1, 1)
        key_mask = None if attn_mask is None else attn_mask.to(torch.uint8).contiguous()
        attention_out = self.cross_attn(x, queries, key_mask)
        return self.ffn(self.ln_ffn(attention_out))
