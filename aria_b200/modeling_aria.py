"""Host-side mirror of `aria/model/modeling_aria.py`: AriaForConditionalGeneration.forward()/generate() with the
Hugging Face state-dict layout, running entirely on the B200-native kernels.

forward() follows modeling_aria.py:194-335: embed -> vision tower -> projector -> masked_scatter merge -> LM.
Differences that are deliberate: no autograd / labels path (inference hot path only), the KV cache is our
static `KVCache` (HF layout [B,H,T,hd] per layer), integer index tensors stay int32 on the device.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import ops
from .moe_lm import AriaMoELMConfig, AriaMoELMForCausalLM, KVCache, bf16
from .projector import AriaProjector
from .vision_encoder import AriaVisionConfig, AriaVisionModel


class AriaConfig:
    """configuration_aria.py:31-114 (attributes used on the hot path)."""

    def __init__(self, vision_config, text_config, projector_patch_to_query_dict=None, image_token_index=32000,
                 ignore_index=-100, **_ignored):
        self.vision_config = vision_config if not isinstance(vision_config, dict) else AriaVisionConfig(**vision_config)
        self.text_config = text_config if not isinstance(text_config, dict) else AriaMoELMConfig(**text_config)
        self.projector_patch_to_query_dict = {int(k): int(v) for k, v in
                                              (projector_patch_to_query_dict or {1225: 128, 4900: 256}).items()}
        self.image_token_index = image_token_index
        self.ignore_index = ignore_index

    @classmethod
    def from_dict(cls, cfg: dict):
        """From the plain-dict configs used by oracle/configs.py and bench.py."""
        return cls(cfg["vision_config"], cfg["text_config"], cfg["projector"]["patch_to_query_dict"],
                   cfg["image_token_index"])


class AriaCausalLMOutputWithPast:
    def __init__(self, logits, past_key_values):
        self.logits = logits
        self.past_key_values = past_key_values
        self.loss = None


class AriaForConditionalGeneration(nn.Module):
    """modeling_aria.py:125-365."""

    def __init__(self, config: AriaConfig, device=None):
        super().__init__()
        self.config = config
        v, t = config.vision_config, config.text_config
        self.vision_tower = AriaVisionModel(v, device)
        self.multi_modal_projector = AriaProjector(  # build_mm_projector, modeling_aria.py:102-122
            config.projector_patch_to_query_dict, v.hidden_size, v.num_attention_heads, v.hidden_size, t.hidden_size,
            t.hidden_size, device)
        self.vocab_size = t.vocab_size
        self.language_model = AriaMoELMForCausalLM(t, device)

    # ---- modeling_aria.py:145-192: the helpers aria/train.py:70-75 and the recipes call on the model
    def freeze_vit(self):
        for p in self.vision_tower.parameters():
            p.requires_grad = False

    def freeze_projector(self):
        for p in self.multi_modal_projector.parameters():
            p.requires_grad = False

    def freeze_llm(self):
        for p in self.language_model.parameters():
            p.requires_grad = False

    def get_input_embeddings(self):
        return self.language_model.get_input_embeddings()

    def set_input_embeddings(self, value):
        self.language_model.set_input_embeddings(value)

    def get_output_embeddings(self):
        return self.language_model.get_output_embeddings()

    def set_output_embeddings(self, value):
        self.language_model.set_output_embeddings(value)

    def set_moe_z_loss_coeff(self, value):
        self.language_model.set_z_loss_coeff(value)

    def set_moe_aux_loss_coeff(self, value):
        self.language_model.set_aux_loss_coeff(value)

    def enable_expert_parallel(self, max_tokens: int, group=None):
        """Shard the routed experts of every MoE layer over the ranks of `group` (rank r serves experts
        [r*E/W, (r+1)*E/W), a dim-0 view of the HF weights) and exchange token rows over NVLink peer memory
        (aria_b200.expert_parallel.PeerTransport).  Every rank must then call forward() in lock-step with its own tokens."""
        import torch.distributed as dist
        from .expert_parallel import ExpertParallelMoE, PeerTransport
        t = self.config.text_config
        W, r = dist.get_world_size(group), dist.get_rank(group)
        tr = PeerTransport(max_tokens, t.hidden_size, t.moe_num_experts, t.moe_topk, self.device, group)
        lo, hi = r * t.moe_num_experts // W, (r + 1) * t.moe_num_experts // W
        for layer in self.language_model.model.layers:
            m = layer.mlp
            w = {"router.weight": m.router.weight, "experts.fc1.weight": m.experts.fc1.weight[lo:hi],
                 "experts.fc2.weight": m.experts.fc2.weight[lo:hi],
                 "shared_experts.gate_proj.weight": m.shared_experts.gate_proj.weight,
                 "shared_experts.up_proj.weight": m.shared_experts.up_proj.weight,
                 "shared_experts.down_proj.weight": m.shared_experts.down_proj.weight}
            m.expert_parallel = ExpertParallelMoE(w, t.moe_num_experts, t.moe_topk, group=group, transport=tr)
        self._ep_transport = tr
        return tr

    @property
    def device(self):
        return self.language_model.lm_head.weight.device

    @torch.no_grad()
    def forward(self, input_ids: torch.Tensor = None, pixel_values: Optional[torch.Tensor] = None,
                pixel_mask: Optional[torch.Tensor] = None, past_key_values: Optional[KVCache] = None,
                inputs_embeds: Optional[torch.Tensor] = None, num_logits_to_keep: int = 0,
                max_cache_len: Optional[int] = None, input_ids_host: Optional[torch.Tensor] = None,
                **_unused) -> AriaCausalLMOutputWithPast:
        """input_ids / pixel_values / pixel_mask may be HOST tensors (pinned for async copies): they are copied to
        the device on the current stream; image-token bookkeeping is then done on the host copy (no device sync).
        Device-resident input_ids cost one sync for the image-token count check, like the reference's `.item()`
        (modeling_aria.py:265)."""
        dev = self.device
        ids_host = input_ids_host  # optional host copy of device-resident ids (bookkeeping without a sync)
        if input_ids is not None and not input_ids.is_cuda:
            ids_host = input_ids
            input_ids = input_ids.to(dev, non_blocking=True)
        if pixel_values is not None and not pixel_values.is_cuda:
            pixel_values = pixel_values.to(dev, non_blocking=True)
        if inputs_embeds is None:
            inputs_embeds = ops.embedding(input_ids.contiguous(), self.get_input_embeddings().weight)

        if pixel_values is not None:
            feats, image_attn_mask = self.vision_tower(pixel_values.to(bf16), pixel_mask)
            image_features = self.multi_modal_projector(feats, image_attn_mask)
            n_image_features = image_features.shape[0] * image_features.shape[1]
            src = ids_host if ids_host is not None else input_ids
            n_image_tokens = int((src == self.config.image_token_index).sum())
            if n_image_tokens != n_image_features:
                raise ValueError(  # modeling_aria.py:268-271
                    f"Image features and image tokens do not match: tokens: {n_image_tokens}, features {n_image_features}")
            ops.merge_image_features(input_ids.reshape(-1).contiguous(), self.config.image_token_index,
                                     image_features.reshape(-1, image_features.shape[-1]),
                                     inputs_embeds.view(-1, inputs_embeds.shape[-1]))

        B, T, _ = inputs_embeds.shape
        cache = past_key_values
        if cache is None:
            cache = self.language_model.new_cache(B, max_cache_len or T, dev)
        logits, cache = self.language_model(inputs_embeds, cache, num_logits_to_keep)
        return AriaCausalLMOutputWithPast(logits, cache)

    @torch.no_grad()
    def generate(self, input_ids, pixel_values=None, pixel_mask=None, max_new_tokens: int = 16):
        """Greedy decoding (the reference goes through HF GenerationMixin, modeling_aria.py:125,337-365):
        prefill with the image, then one token per step against the KV cache (pixel inputs only at step 0)."""
        B, T = input_ids.shape
        out = self.forward(input_ids, pixel_values, pixel_mask, num_logits_to_keep=1, max_cache_len=T + max_new_tokens)
        cache = out.past_key_values
        tokens = [out.logits[:, -1].float().argmax(-1)]
        for _ in range(max_new_tokens - 1):
            step = self.forward(tokens[-1].view(B, 1), past_key_values=cache, num_logits_to_keep=1)
            tokens.append(step.logits[:, -1].float().argmax(-1))
        return torch.cat([input_ids.to(tokens[0].device), torch.stack(tokens, 1)], dim=1)


class GraphedPrefill:
    """CUDA-graph capture of one prefill `forward()` for fixed shapes (streams + graphs instead of a tracing
    compiler): the whole ViT -> projector -> merge -> LM chain has no host sync, so it is captured once and replayed.

        g = GraphedPrefill(model, input_ids_host, pixel_values_host)      # warm-up + capture
        logits = g(input_ids_host, pixel_values_host)                      # H2D copies + replay; logits on device
        logits = g.replay()                                                # inputs already resident in HBM
    """

    def __init__(self, model: "AriaForConditionalGeneration", input_ids: torch.Tensor, pixel_values: torch.Tensor,
                 num_logits_to_keep: int = 1):
        self.model = model
        dev = model.device
        self.ids_dev = torch.empty(input_ids.shape, dtype=torch.int64, device=dev)
        self.pv_dev = torch.empty(pixel_values.shape, dtype=bf16, device=dev)
        self.ids_dev.copy_(input_ids)
        self.pv_dev.copy_(pixel_values)
        self.n_image_tokens = int((input_ids == model.config.image_token_index).sum())
        ids_host = input_ids.cpu()
        kw = dict(num_logits_to_keep=num_logits_to_keep, input_ids_host=ids_host)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):
                model(self.ids_dev, self.pv_dev, None, **kw)
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.logits = model(self.ids_dev, self.pv_dev, None, **kw).logits

    def replay(self) -> torch.Tensor:
        self.graph.replay()
        return self.logits

    def __call__(self, input_ids: torch.Tensor, pixel_values: torch.Tensor) -> torch.Tensor:
        if not input_ids.is_cuda:  # same ValueError contract as forward() (modeling_aria.py:268-271), checked on the host
            n = int((input_ids == self.model.config.image_token_index).sum())
            if n != self.n_image_tokens:
                raise ValueError(f"Image features and image tokens do not match: tokens: {n}, features {self.n_image_tokens}")
        self.ids_dev.copy_(input_ids, non_blocking=True)
        self.pv_dev.copy_(pixel_values, non_blocking=True)
        return self.replay()


def init_random_(model: nn.Module, seed: int = 0, std: float = 0.02):
    """Random-init (no checkpoint offline): N(0, std^2) for matrices / embeddings / biases / queries, 1 for the
    norm scales (the only 1-D parameters named `weight`)."""
    g = torch.Generator(device=next(model.parameters()).device).manual_seed(seed)
    for name, p in model.named_parameters():
        if p.dim() == 1 and name.endswith("weight"):
            p.fill_(1.0)
        else:
            p.normal_(0.0, std, generator=g)
    return model
