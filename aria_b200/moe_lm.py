"""Host-side mirror of the reference's `aria/model/moe_lm.py` operator interface, running on the
B200-native kernels (libaria_b200.so).  Same class names, parameter names/shapes (HF checkpoint keys) and
argument meaning as the reference; the arithmetic is ours.

    reference                                  here
    ---------------------------------------   -------------------------------------------------------------
    TopKRouter            moe_lm.py:170-293    TopKRouter        -> aria_router_topk (tcgen05 GEMM + warp top-k)
    TokenDispatcher       moe_lm.py:297-365    TokenDispatcher   -> counting sort + 128-bit row gather / combine
    experts_gemm / gmm    moe_lm.py:431-443    experts_gemm      -> aria_grouped_gemm (no .cpu() sync)
    GroupedGEMM           moe_lm.py:446-484    GroupedGEMM       (class + `weight` [E,in,out] preserved for PEFT)
    GroupedMLP            moe_lm.py:487-525    GroupedMLP        -> fc1 with fused SwiGLU epilogue, fc2
    SharedExpertMLP       moe_lm.py:368-395    SharedExpertMLP   -> gate/up fused SwiGLU GEMM + down GEMM
    MoELayer              moe_lm.py:528-577    MoELayer
    MoEDecoderLayer       moe_lm.py:580-602    MoEDecoderLayer   (+ AriaAttention for LLAMA_ATTENTION_CLASSES[...])
    AriaMoELMModel        moe_lm.py:605-636    AriaMoELMModel
    AriaMoELMForCausalLM  moe_lm.py:639-679    AriaMoELMForCausalLM

These modules are the inference path (eval-mode routing, moe_lm.py:261-269); the differentiable MoE block incl. the
training-mode aux/z losses (moe_lm.py:84-166) lives in aria_b200/moe_train.py.  There is no CPU fallback.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import os

import torch
import torch.nn as nn

from . import _lib as L
from . import ops

bf16 = torch.bfloat16


class AriaMoELMConfig:
    """Plain-attribute stand-in for the reference `AriaMoELMConfig(LlamaConfig)` (moe_lm.py:43-80)."""

    def __init__(self, hidden_size=4096, num_attention_heads=32, num_hidden_layers=32, vocab_size=32000,
                 moe_intermediate_size=4096, moe_num_experts=8, moe_topk=2, moe_num_shared_experts=2,
                 moe_z_loss_coeff=1e-5, moe_aux_loss_coeff=1e-3, rms_norm_eps=1e-6, rope_theta=10000.0, **_ignored):
        self.hidden_size = hidden_size
        self.num_attention_heads = num_attention_heads
        self.num_hidden_layers = num_hidden_layers
        self.vocab_size = vocab_size
        self.moe_intermediate_size = moe_intermediate_size
        self.moe_num_experts = moe_num_experts
        self.moe_topk = moe_topk
        self.moe_num_shared_experts = moe_num_shared_experts
        self.moe_z_loss_coeff = moe_z_loss_coeff      # training-mode router losses (moe_lm.py:57-58); used by moe_train
        self.moe_aux_loss_coeff = moe_aux_loss_coeff
        self.rms_norm_eps = rms_norm_eps
        self.rope_theta = rope_theta
        self.head_dim = hidden_size // num_attention_heads


class MoEAuxLossAutoScaler:
    """Holder of the scale applied to the router-loss gradients (reference `MoEAuxLossAutoScaler`, moe_lm.py:84-125: the
    aux losses never enter the returned loss, their gradient is injected with this scale in backward)."""

    main_loss_backward_scale: float = 1.0

    @staticmethod
    def set_loss_scale(scale):
        MoEAuxLossAutoScaler.main_loss_backward_scale = float(scale)


def _param(*shape, device=None):
    return nn.Parameter(torch.empty(*shape, dtype=bf16, device=device), requires_grad=False)


class Linear(nn.Module):
    """Parameter holder with nn.Linear's names/layout (`weight` [out,in], optional `bias`)."""

    def __init__(self, in_features, out_features, bias=False, device=None):
        super().__init__()
        self.weight = _param(out_features, in_features, device=device)
        self.bias = _param(out_features, device=device) if bias else None

    def forward(self, x, act=L.ACT_NONE, residual=None):
        return ops.linear(x, self.weight, self.bias, act=act, residual=residual)


class RMSNorm(nn.Module):
    def __init__(self, d, eps, device=None):
        super().__init__()
        self.weight = _param(d, device=device)
        self.variance_epsilon = eps

    def forward(self, x, residual=None):
        return ops.rmsnorm(x, self.weight, self.variance_epsilon, residual)


class TopKRouter(nn.Module):
    """moe_lm.py:170-293.  forward(input[T,d]) -> (scores [T,k] bf16, top_indices [T,k], tokens_per_expert [E]).
    Indices/counts are int32 and stay on the device (the reference's int64 is an ATen artefact)."""

    def __init__(self, config, device=None):
        super().__init__()
        self.config = config
        self.weight = _param(config.moe_num_experts, config.hidden_size, device=device)
        # Parity / replay hook: when set to an int32 CUDA tensor [T, k], the expert choice of the next forward() is TAKEN
        # from it (scores and counts are still computed on the device from this router's own logits).  The forced-routing
        # parity test injects the oracle's top-k here so that a bf16 near-tie in the router cannot mask other differences.
        self.forced_top_indices: Optional[torch.Tensor] = None

    def forward(self, input: torch.Tensor):
        x = input.reshape(-1, input.shape[-1])
        if self.forced_top_indices is not None:
            logits = ops.linear(x, self.weight)                                   # gating, moe_lm.py:200
            top_indices = self.forced_top_indices
            scores, tokens_per_expert = ops.route_given_indices(logits, top_indices)
            return scores, top_indices, tokens_per_expert
        scores, top_indices, tokens_per_expert, _ = ops.router_topk(x, self.weight, self.config.moe_topk)
        return scores, top_indices, tokens_per_expert


class TokenDispatcher:
    """moe_lm.py:297-365: permutes token rows into expert-sorted order and combines them back."""

    def __init__(self, config):
        self.config = config
        self.hidden_states_shape = None
        self.reversed_input_permutation_mapping = None  # dest_row: flattened (token,slot) -> sorted row
        self.expert_offsets = None

    def token_permutation(self, hidden_states: torch.Tensor, indices: torch.Tensor,
                          tokens_per_expert: Optional[torch.Tensor] = None) -> torch.Tensor:
        self.hidden_states_shape = hidden_states.shape
        x = hidden_states.reshape(-1, hidden_states.shape[-1])
        if tokens_per_expert is None:
            raise RuntimeError("tokens_per_expert (device int32 counts from the router) is required")
        offsets, dest_row, src_token = ops.build_permutation(indices, tokens_per_expert)
        self.reversed_input_permutation_mapping = dest_row
        self.expert_offsets = offsets
        return ops.permute_rows(x, src_token)

    def token_unpermutation(self, permuted_tokens: torch.Tensor, scores: torch.Tensor,
                            shared: Optional[torch.Tensor] = None) -> torch.Tensor:
        sh = None if shared is None else shared.reshape(-1, shared.shape[-1])
        out = ops.unpermute_combine(permuted_tokens, self.reversed_input_permutation_mapping, scores, sh)
        return out.view(self.hidden_states_shape)


def _as_offsets(tokens_per_expert: torch.Tensor, num_experts: int, device) -> torch.Tensor:
    """Accept what the reference passes at moe_lm.py:478-484 (per-expert counts [E], any integer dtype, CPU or
    CUDA) or our int32 device row offsets [E+1] (TokenDispatcher.expert_offsets)."""
    t = tokens_per_expert
    if t.numel() == num_experts + 1:
        if t.dtype != torch.int32:      # (CUDA residency is enforced where the pointer is taken, ops._chk)
            raise RuntimeError("row offsets must be an int32 tensor of E+1 entries")
        return t
    if t.numel() != num_experts:
        raise RuntimeError(f"tokens_per_expert must have {num_experts} (counts) or {num_experts + 1} (offsets) entries")
    return ops.offsets_from_counts(t.to(device=device, dtype=torch.int64).contiguous())


def experts_gemm(input: torch.Tensor, weight: torch.Tensor, tokens_per_expert: torch.Tensor) -> torch.Tensor:
    """Drop-in for `grouped_gemm.ops.gmm` / `sequential_gemm` as bound at moe_lm.py:431-443:
    (input [rows,K] bf16, weight [E,K,N] bf16, tokens_per_expert [E]) -> [rows,N].
    `tokens_per_expert` may be counts [E] (reference contract) or int32 device offsets [E+1]."""
    return ops.grouped_gemm(input, weight, _as_offsets(tokens_per_expert, weight.shape[0], input.device))


gmm = experts_gemm  # name used by `from grouped_gemm.ops import gmm`


class GroupedGEMM(nn.Module):
    """moe_lm.py:446-484: `weight` [groups, in_features, out_features] (out contiguous — HF layout, untouched)."""

    def __init__(self, in_features, out_features, groups, device=None):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.groups = groups
        self.weight = _param(groups, in_features, out_features, device=device)

    def forward(self, input, tokens_per_expert):
        return experts_gemm(input, self.weight, tokens_per_expert)


class GroupedMLP(nn.Module):
    """moe_lm.py:487-525: fc1 -> glu (first half gate, second half up) -> fc2.  The glu is fused into fc1's
    epilogue with the reference's bf16 rounding points."""

    def __init__(self, config, device=None):
        super().__init__()
        self.config = config
        self.fc1 = GroupedGEMM(config.hidden_size, config.moe_intermediate_size * 2, config.moe_num_experts, device)
        self.fc2 = GroupedGEMM(config.moe_intermediate_size, config.hidden_size, config.moe_num_experts, device)

    def forward(self, permuted_tokens, tokens_per_expert):
        off = _as_offsets(tokens_per_expert, self.fc1.groups, permuted_tokens.device)
        if type(self.fc1) is GroupedGEMM and type(self.fc2) is GroupedGEMM:
            h = ops.grouped_gemm(permuted_tokens, self.fc1.weight, off, swiglu=True)
            return ops.grouped_gemm(h, self.fc2.weight, off)
        # fc1 / fc2 wrapped by an adapter (aria_b200.lora.GroupedGemmLoraLayer, what peft does to the reference's
        # GroupedGEMM modules, aria/train.py:107): call through the modules as the reference does (moe_lm.py:521-525);
        # the glu runs as its own (differentiable) kernel because the adapter term must be added before it
        from .lora import swiglu
        return self.fc2(swiglu(self.fc1(permuted_tokens, off)), off)


class SharedExpertMLP(nn.Module):
    """moe_lm.py:368-395 (LlamaMLP with intermediate = I * num_shared): down(silu(gate(x)) * up(x))."""

    def __init__(self, config, device=None):
        super().__init__()
        self.hidden_size = config.hidden_size
        self.intermediate_size = config.moe_intermediate_size * config.moe_num_shared_experts
        self.gate_proj = Linear(self.hidden_size, self.intermediate_size, device=device)
        self.up_proj = Linear(self.hidden_size, self.intermediate_size, device=device)
        self.down_proj = Linear(self.intermediate_size, self.hidden_size, device=device)

    def forward(self, x):
        h = ops.linear_swiglu(x, self.gate_proj.weight, self.up_proj.weight)
        return ops.linear(h, self.down_proj.weight)


_SIDE_STREAMS = {}


def shared_expert_overlapped(fn, like: torch.Tensor):
    """Run `fn()` (the shared-expert branch, moe_lm.py:575: independent of the routed branch until the final add) on a side
    stream of `like`'s device, forked from / joined to the current stream — under CUDA-graph capture this becomes a parallel
    branch of the graph.  At prefill sizes the branch is two latency-bound weight-streaming GEMMs that otherwise sit
    serialised between HBM-saturating expert GEMMs.  ARIA_MOE_SIDE_STREAM=0 runs it in line."""
    import os
    if os.environ.get("ARIA_MOE_SIDE_STREAM", "1") == "0" or not like.is_cuda:
        return fn()
    dev = like.device
    side = _SIDE_STREAMS.get(dev)
    if side is None:
        side = _SIDE_STREAMS[dev] = torch.cuda.Stream(device=dev)
    cur = torch.cuda.current_stream(dev)
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        out = fn()
    out.record_stream(cur)     # allocated on the side stream, consumed on the current one
    return out, side


def _side_stream(dev):
    """The per-device side stream of the shared-expert branch (None when ARIA_MOE_SIDE_STREAM=0: branch runs in line)."""
    import os
    if os.environ.get("ARIA_MOE_SIDE_STREAM", "1") == "0":
        return None
    side = _SIDE_STREAMS.get(dev)
    if side is None:
        side = _SIDE_STREAMS[dev] = torch.cuda.Stream(device=dev)
    return side


def join_side(forked, like: torch.Tensor):
    """Second half of `shared_expert_overlapped`: make the current stream wait for the side branch; returns its result."""
    if isinstance(forked, tuple):
        out, side = forked
        torch.cuda.current_stream(like.device).wait_stream(side)
        return out
    return forked


class MoELayer(nn.Module):
    """moe_lm.py:528-577.  forward(hidden_states [B,T,d]) -> [B,T,d]."""

    def __init__(self, config, device=None):
        super().__init__()
        self.router = TopKRouter(config, device)
        self.token_dispatcher = TokenDispatcher(config)
        self.experts = GroupedMLP(config, device)
        self.shared_experts = SharedExpertMLP(config, device)
        self.expert_parallel = None  # set by AriaForConditionalGeneration.enable_expert_parallel()

    def forward(self, hidden_states: torch.Tensor) -> torch.Tensor:
        if self.expert_parallel is not None:
            return self.expert_parallel(hidden_states)
        if (hidden_states.is_cuda and type(self.experts.fc1) is GroupedGEMM and type(self.experts.fc2) is GroupedGEMM
                and os.environ.get("ARIA_MOE_BLOCK", "1") != "0"):   # =0: one C-ABI call per kernel (bench.py's per-kernel table)
            # the whole block behind one C-ABI call (csrc/moe_block.cu); shared experts on the side stream of this device
            x = hidden_states.reshape(-1, hidden_states.shape[-1])
            se = self.shared_experts
            out = ops.moe_block_fwd(x, self.router.weight, self.experts.fc1.weight, self.experts.fc2.weight, se.gate_proj.weight,
                                    se.up_proj.weight, se.down_proj.weight, self.router.config.moe_topk,
                                    forced_top_idx=self.router.forced_top_indices, side_stream=_side_stream(x.device))
            return out.view(hidden_states.shape)
        # module-by-module path (adapter-wrapped experts, CPU stand-in ops of the host-logic tests)
        forked = shared_expert_overlapped(lambda: self.shared_experts(hidden_states), hidden_states)
        scores, indices, tokens_per_expert = self.router(hidden_states)
        permuted_tokens = self.token_dispatcher.token_permutation(hidden_states, indices, tokens_per_expert)
        expert_output = self.experts(permuted_tokens, self.token_dispatcher.expert_offsets)
        shared_expert_output = join_side(forked, hidden_states)
        # unpermute + score-weighted sum + `output += shared_expert_output` (moe_lm.py:573-576) in one kernel
        return self.token_dispatcher.token_unpermutation(expert_output, scores, shared_expert_output)


class KVCache:
    """Static per-layer KV cache in the HF layout [B, H, T_max, head_dim] (modeling_aria.py:49-50)."""

    def __init__(self, n_layers, B, H, T_max, hd, device):
        # rows >= seq_len are never read (the attention TMA maps cover the valid rows only), so no zero-fill
        self.k = [torch.empty(B, H, T_max, hd, dtype=bf16, device=device) for _ in range(n_layers)]
        self.v = [torch.empty(B, H, T_max, hd, dtype=bf16, device=device) for _ in range(n_layers)]
        self.q = torch.empty(B, H, T_max, hd, dtype=bf16, device=device)  # rows [seq_len, seq_len+T) used per step
        self.seq_len = 0
        self.T_max = T_max


class AriaAttention(nn.Module):
    """What `LLAMA_ATTENTION_CLASSES[config._attn_implementation]` provides at moe_lm.py:594: MHA, no bias,
    rotate-half RoPE, causal, KV cache.  q/k/v projections + RoPE + cache write are ONE GEMM launch."""

    def __init__(self, config, layer_idx, device=None):
        super().__init__()
        self.config = config
        self.layer_idx = layer_idx
        d = config.hidden_size
        self.num_heads = config.num_attention_heads
        self.head_dim = d // self.num_heads
        if self.head_dim != 128:
            raise RuntimeError("AriaAttention kernels are written for head_dim 128")
        self.q_proj = Linear(d, d, device=device)
        self.k_proj = Linear(d, d, device=device)
        self.v_proj = Linear(d, d, device=device)
        self.o_proj = Linear(d, d, device=device)

    def forward(self, hidden_states, cache: KVCache, rope, residual=None, key_mask=None, position_ids=None):
        """key_mask [B, pos0+T] uint8, 1 = key masked out (padded batch); position_ids [B*T] int32 RoPE positions
        (default: cache position pos0 + t, what LlamaModel uses when none are given)."""
        B, T, d = hidden_states.shape
        H, hd = self.num_heads, self.head_dim
        pos0 = cache.seq_len
        if pos0 + T > cache.T_max:
            # the fused epilogue stores k/v rows at pos0 + t and reads the RoPE table there: never past the cache
            raise RuntimeError(f"KV cache overflow: {pos0} cached + {T} new tokens > T_max = {cache.T_max}")
        kc, vc = cache.k[self.layer_idx], cache.v[self.layer_idx]
        q = cache.q  # staging buffer with the cache's strides: the fused epilogue scatters q, k, v with one stride pair
        cos, sin = rope
        ops.qkv_heads(hidden_states, [self.q_proj.weight, self.k_proj.weight, self.v_proj.weight], [None] * 3,
                      [q, kc, vc], hd, T, pos0=pos0, rope_mask=0b011, rope_cos=cos, rope_sin=sin, position_ids=position_ids)
        Tk = pos0 + T
        scale = hd ** -0.5
        if T == 1:
            o = ops.attention_decode(q[:, :, pos0, :], kc, vc, Tk, scale, key_mask=key_mask).view(B, 1, d)  # strided view, no copy
        else:  # prefill, or a multi-token continuation (chunked prefill): the queries are the last T of Tk positions
            o = ops.attention(q[:, :, pos0:], kc, vc, T, Tk, scale, causal=True, key_mask=key_mask)
        return ops.linear(o, self.o_proj.weight, residual=residual)


class MoEDecoderLayer(nn.Module):
    """moe_lm.py:580-602: x + attn(rms(x)); h + moe(rms(h)).  The MoE residual add is deferred into the next
    RMSNorm kernel (same bf16 rounding as the reference's separate add)."""

    def __init__(self, config, layer_idx, device=None):
        super().__init__()
        self.hidden_size = config.hidden_size
        self.self_attn = AriaAttention(config, layer_idx, device)
        self.mlp = MoELayer(config, device)
        self.input_layernorm = RMSNorm(config.hidden_size, config.rms_norm_eps, device)
        self.post_attention_layernorm = RMSNorm(config.hidden_size, config.rms_norm_eps, device)

    def forward(self, x, pending, cache, rope, key_mask=None, position_ids=None):
        """x: residual stream; pending: MoE output of the previous layer not yet added (or None)."""
        if pending is None:
            h = self.input_layernorm(x)
        else:
            h, x = self.input_layernorm(x, residual=pending)
        x = self.self_attn(h, cache, rope, residual=x, key_mask=key_mask, position_ids=position_ids)
        h = self.post_attention_layernorm(x)
        return x, self.mlp(h)


class AriaMoELMModel(nn.Module):
    """moe_lm.py:605-636."""

    def __init__(self, config, device=None):
        super().__init__()
        self.config = config
        self.embed_tokens = nn.Embedding(config.vocab_size, config.hidden_size, device=device, dtype=bf16)
        self.embed_tokens.weight.requires_grad_(False)
        self.layers = nn.ModuleList([MoEDecoderLayer(config, i, device) for i in range(config.num_hidden_layers)])
        self.norm = RMSNorm(config.hidden_size, config.rms_norm_eps, device)
        self._rope = None

    def rope_tables(self, n_pos, device):
        if self._rope is None or self._rope[0].shape[0] < n_pos or self._rope[0].device != device:
            hd = self.config.head_dim
            # LlamaRotaryEmbedding: inv_freq in fp32 exactly as transformers computes it (moe_lm.py:632)
            inv_freq = 1.0 / (self.config.rope_theta ** (torch.arange(0, hd, 2, dtype=torch.int64).float() / hd))
            self._rope = ops.rope_table(inv_freq.to(device), n_pos)
        return self._rope

    def forward(self, inputs_embeds, cache: KVCache, key_mask=None, position_ids=None):
        B, T, _ = inputs_embeds.shape
        if cache.seq_len + T > cache.T_max:
            raise RuntimeError(f"KV cache overflow: {cache.seq_len} cached + {T} new tokens > T_max = {cache.T_max}")
        rope = self.rope_tables(cache.T_max, inputs_embeds.device)
        x, pending = inputs_embeds, None
        for layer in self.layers:
            x, pending = layer(x, pending, cache, rope, key_mask, position_ids)
        cache.seq_len += T
        return x, pending  # final residual add happens inside the final norm


class AriaMoELMForCausalLM(nn.Module):
    """moe_lm.py:639-679."""

    def __init__(self, config, device=None):
        super().__init__()
        self.config = config
        self.model = AriaMoELMModel(config, device)
        self.vocab_size = config.vocab_size
        self.lm_head = Linear(config.hidden_size, config.vocab_size, device=device)

    def new_cache(self, B, T_max, device):
        c = self.config
        return KVCache(c.num_hidden_layers, B, c.num_attention_heads, T_max, c.head_dim, device)

    # moe_lm.py:663-679: the routers read the coefficients from the shared config object (used by moe_train's router losses)
    def set_z_loss_coeff(self, z_loss_coeff: float):
        self.config.moe_z_loss_coeff = z_loss_coeff

    def set_aux_loss_coeff(self, aux_loss_coeff: float):
        self.config.moe_aux_loss_coeff = aux_loss_coeff

    def get_input_embeddings(self):
        return self.model.embed_tokens

    def set_input_embeddings(self, value):
        self.model.embed_tokens = value

    def get_output_embeddings(self):
        return self.lm_head

    def set_output_embeddings(self, value):
        self.lm_head = value

    def forward(self, inputs_embeds, cache: Optional[KVCache] = None, num_logits_to_keep: int = 0, key_mask=None,
                position_ids=None):
        B, T, _ = inputs_embeds.shape
        if cache is None:
            cache = self.new_cache(B, T, inputs_embeds.device)
        x, pending = self.model(inputs_embeds, cache, key_mask, position_ids)
        if num_logits_to_keep:
            x = x[:, -num_logits_to_keep:, :].contiguous()
            pending = pending[:, -num_logits_to_keep:, :].contiguous()
        h, _ = self.model.norm(x, residual=pending)
        return self.lm_head(h), cache
