"""Host-side mirror of `aria/model/vision_encoder.py` (AriaVisionModel over Idefics2VisionTransformer without
post-layernorm, vision_encoder.py:58-67) on the B200-native kernels.  Parameter names = HF checkpoint keys.

Per encoder layer: LayerNorm kernel -> ONE fused q/k/v GEMM (bias, head-major scatter, head_dim 72 zero-padded to
128) -> tcgen05 attention (non-causal, key-padding mask) -> out_proj GEMM (+bias +residual) -> LayerNorm ->
fc1 GEMM (+bias + gelu_tanh) -> fc2 GEMM (+bias +residual).
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import _lib as L
from . import ops
from .moe_lm import Linear, _param, bf16


class AriaVisionConfig:
    def __init__(self, hidden_size=1152, num_attention_heads=16, num_hidden_layers=27, intermediate_size=4304,
                 patch_size=14, image_size=980, layer_norm_eps=1e-6, num_channels=3, **_ignored):
        self.hidden_size = hidden_size
        self.num_attention_heads = num_attention_heads
        self.num_hidden_layers = num_hidden_layers
        self.intermediate_size = intermediate_size
        self.patch_size = patch_size
        self.image_size = image_size
        self.layer_norm_eps = layer_norm_eps
        self.num_channels = num_channels


class LayerNorm(nn.Module):
    def __init__(self, d, eps, device=None):
        super().__init__()
        self.weight = _param(d, device=device)
        self.bias = _param(d, device=device)
        self.eps = eps

    def forward(self, x):
        return ops.layernorm(x, self.weight, self.bias, self.eps)


class _PatchEmbedding(nn.Module):
    """nn.Conv2d(3, C, k=P, s=P) holder: `weight` [C,3,P,P], `bias` [C]; run as im2col + GEMM."""

    def __init__(self, cfg, device=None):
        super().__init__()
        self.weight = _param(cfg.hidden_size, cfg.num_channels, cfg.patch_size, cfg.patch_size, device=device)
        self.bias = _param(cfg.hidden_size, device=device)
        self._packed = None

    def packed_weight(self):
        """[C, k_pad] with k = 3*P*P zero-padded to a multiple of 64 (TMA needs 16-byte row strides; 588*2 is not)."""
        w = self.weight
        if self._packed is None or self._packed[0] != w._version or self._packed[1].device != w.device:
            k = w[0].numel()
            k_pad = (k + 63) // 64 * 64
            pk = torch.zeros(w.shape[0], k_pad, dtype=bf16, device=w.device)
            pk[:, :k] = w.reshape(w.shape[0], k)
            self._packed = (w._version, pk)
        return self._packed[1]


class Idefics2VisionEmbeddings(nn.Module):
    def __init__(self, cfg, device=None):
        super().__init__()
        self.cfg = cfg
        self.patch_embedding = _PatchEmbedding(cfg, device)
        self.num_patches_per_side = cfg.image_size // cfg.patch_size
        self.position_embedding = nn.Embedding(self.num_patches_per_side ** 2, cfg.hidden_size, device=device, dtype=bf16)
        self.position_embedding.weight.requires_grad_(False)
        self._bucket_cache = {}
        self._full_ids = None

    def _buckets(self, nb: int) -> torch.Tensor:
        """Bucket index of each of `nb` valid rows (or columns): Idefics2VisionEmbeddings.forward of the transformers release the
        reference pins (4.46.3, pyproject.toml:13) — `bucketize(arange(0, 1 - 1e-6, 1 / nb), boundaries, right=True)` with fp32
        coordinates computed on the HOST.  When nb == num_patches_per_side every coordinate sits on a boundary up to fp32
        rounding, so the result depends on the exact arithmetic: this runs the very same torch CPU ops (host-side index glue,
        a few hundred bytes, cached per nb).  (transformers 5.x casts the coordinates to the pixel dtype first — bf16 moves
        half of the 70 buckets by one; a real checkpoint was trained with the fp32 ids.)"""
        t = self._bucket_cache.get(nb)
        if t is None:
            n = self.num_patches_per_side
            boundaries = torch.arange(1 / n, 1.0, 1 / n)
            step = 1 / torch.tensor(nb)                       # fp32 tensor, as `1 / nb_patches_h` is in the reference
            t = torch.bucketize(torch.arange(0, 1 - 1e-6, step), boundaries, right=True)
            self._bucket_cache[nb] = t
        return t

    def position_ids(self, patch_mask_host: Optional[torch.Tensor], B: int, device) -> torch.Tensor:
        """patch_mask_host: None (every patch valid) or the [B, Hp, Wp] bool patch mask ON THE HOST -> [B*Hp*Wp] int64 ids
        on `device`.  Follows the reference loop: ids of the nb_h x nb_w valid grid are written to the mask's True positions."""
        n = self.num_patches_per_side
        if patch_mask_host is None:
            key = (B, str(device))
            if self._full_ids is None or self._full_ids[0] != key:
                b = self._buckets(n)
                ids = (b[:, None] * n + b[None, :]).flatten()
                self._full_ids = (key, ids.repeat(B).to(device))
            return self._full_ids[1]
        pm = patch_mask_host
        pos = torch.zeros(B, pm.shape[1] * pm.shape[2], dtype=torch.int64)
        for b in range(B):
            nb_h, nb_w = int(pm[b, :, 0].sum()), int(pm[b, 0, :].sum())
            ids = (self._buckets(nb_h)[:, None] * n + self._buckets(nb_w)[None, :]).flatten()
            pos[b][pm[b].reshape(-1)] = ids      # raises on a non-rectangular mask, like the reference
        return pos.reshape(-1).to(device, non_blocking=True)

    def forward(self, pixel_values, patch_mask_host):
        w = self.patch_embedding.packed_weight()
        patches = ops.im2col_patches(pixel_values, self.cfg.patch_size, w.shape[1])
        x = ops.linear(patches, w, self.patch_embedding.bias)
        pos = self.position_ids(patch_mask_host, pixel_values.shape[0], pixel_values.device)
        x = ops.add_pos_embedding(x, pos, self.position_embedding.weight)
        return x.view(pixel_values.shape[0], -1, x.shape[-1])


class Idefics2VisionAttention(nn.Module):
    def __init__(self, cfg, device=None):
        super().__init__()
        d = cfg.hidden_size
        self.num_heads = cfg.num_attention_heads
        self.head_dim = d // self.num_heads
        self.q_proj = Linear(d, d, bias=True, device=device)
        self.k_proj = Linear(d, d, bias=True, device=device)
        self.v_proj = Linear(d, d, bias=True, device=device)
        self.out_proj = Linear(d, d, bias=True, device=device)


class Idefics2MLP(nn.Module):
    def __init__(self, cfg, device=None):
        super().__init__()
        self.fc1 = Linear(cfg.hidden_size, cfg.intermediate_size, bias=True, device=device)
        self.fc2 = Linear(cfg.intermediate_size, cfg.hidden_size, bias=True, device=device)


class Idefics2EncoderLayer(nn.Module):
    def __init__(self, cfg, device=None):
        super().__init__()
        self.self_attn = Idefics2VisionAttention(cfg, device)
        self.layer_norm1 = LayerNorm(cfg.hidden_size, cfg.layer_norm_eps, device)
        self.mlp = Idefics2MLP(cfg, device)
        self.layer_norm2 = LayerNorm(cfg.hidden_size, cfg.layer_norm_eps, device)

    def forward(self, x, qkv_buf, key_mask):
        """x [B,N,d]; qkv_buf: three zero-initialised [B,H,N,128] buffers (pad columns stay zero)."""
        B, N, d = x.shape
        a = self.self_attn
        h = self.layer_norm1(x)
        ops.qkv_heads(h, [a.q_proj.weight, a.k_proj.weight, a.v_proj.weight],
                      [a.q_proj.bias, a.k_proj.bias, a.v_proj.bias], qkv_buf, a.head_dim, N)
        o = ops.attention(qkv_buf[0], qkv_buf[1], qkv_buf[2], N, N, a.head_dim ** -0.5, causal=False,
                          out_hd=a.head_dim, key_mask=key_mask)
        x = ops.linear(o, a.out_proj.weight, a.out_proj.bias, residual=x)
        h = self.layer_norm2(x)
        h = ops.linear(h, self.mlp.fc1.weight, self.mlp.fc1.bias, act=L.ACT_GELU_TANH)
        return ops.linear(h, self.mlp.fc2.weight, self.mlp.fc2.bias, residual=x)


class Idefics2Encoder(nn.Module):
    def __init__(self, cfg, device=None):
        super().__init__()
        self.layers = nn.ModuleList([Idefics2EncoderLayer(cfg, device) for _ in range(cfg.num_hidden_layers)])


class AriaVisionTransformer(nn.Module):
    """vision_encoder.py:58-67: Idefics2VisionTransformer with post_layernorm = IdentityOp."""

    def __init__(self, cfg, device=None):
        super().__init__()
        self.config = cfg
        self.embeddings = Idefics2VisionEmbeddings(cfg, device)
        self.encoder = Idefics2Encoder(cfg, device)

    def forward(self, pixel_values, patch_mask_host: Optional[torch.Tensor] = None):
        """patch_mask_host: None = every patch valid; else the [B, Hp, Wp] bool patch mask on the HOST (position ids and the
        "is anything padded" branch are host-side index glue, as in the reference's per-image loop)."""
        B = pixel_values.shape[0]
        if patch_mask_host is not None and bool(patch_mask_host.all()):
            patch_mask_host = None
        x = self.embeddings(pixel_values, patch_mask_host)
        N = x.shape[1]
        key_mask = None
        if patch_mask_host is not None:
            key_mask = (~patch_mask_host.reshape(B, -1)).to(torch.uint8).to(x.device, non_blocking=True).contiguous()
        H = self.config.num_attention_heads
        qkv = [torch.zeros(B, H, N, 128, dtype=bf16, device=x.device) for _ in range(3)]
        for layer in self.encoder.layers:
            x = layer(x, qkv, key_mask)
        return x


class AriaVisionModel(nn.Module):
    """vision_encoder.py:70-152.  forward(pixel_values [B,3,S,S], pixel_mask [B,S,S] bool|None)
    -> (last_hidden_state [B,N,d], image_attn_mask [B,N] bool with True = padding | None)."""

    def __init__(self, cfg, device=None):
        super().__init__()
        self.config = cfg
        self.vision_model = AriaVisionTransformer(cfg, device)

    def forward(self, pixel_values: torch.Tensor, pixel_mask: Optional[torch.Tensor] = None):
        if pixel_mask is None:
            return self.vision_model(pixel_values, None), None
        # the patch mask is [B, 70, 70] bools: built where the pixel mask lives, used on the host (a device-resident mask costs
        # one small D2H copy — the reference syncs here too, `p_attn_mask.view(-1).cpu()` per image)
        pam = self._create_patch_attention_mask(pixel_mask)
        pam_host = pam.cpu() if pam.is_cuda else pam
        image_atts = torch.logical_not(pam_host.flatten(1)).to(pixel_values.device, non_blocking=True)  # vision_encoder.py:147-152
        return self.vision_model(pixel_values, pam_host), image_atts

    def _create_patch_attention_mask(self, pixel_mask):
        """vision_encoder.py:132-145 (bool glue on a [B,S,S] mask)."""
        P = self.config.patch_size
        sub = pixel_mask.unfold(1, P, P).unfold(2, P, P)
        return (sub.sum(dim=(-1, -2)) > 0).bool()
