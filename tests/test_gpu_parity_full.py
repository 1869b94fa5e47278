"""GPU parity at the sizes BASELINE.json quotes, on ONE GPU, against the oracle (round-1 VERDICT "parity gaps"):

  * whole model with the ORACLE's top-k ids injected (TopKRouter.forced_top_indices): EVERY token within 1e-2 of the logit
    scale — the tie filter of test_gpu_parity.py then only ever excuses genuine bf16 near-ties of the router;
  * one full-width MoE layer (d=2560, E=64, k=6, I=1664, T=768) vs the oracle, forced and free routing;
  * ViT attention at N=4900 / head_dim 72 (ragged last query pair), with and without a key mask;
  * causal LM attention at T=8192 and T=65536 on sampled query rows (fp32 oracle rows, O(T) each);
  * decode attention vs the oracle (not vs our own prefill kernel), with and without a key mask;
  * mirror: padded batches (2-D attention_mask, position_ids), chunked prefill, cache overflow, labels -> loss.

Tolerance: |got - want| <= 1e-2 * max|want| element-wise unless stated (REL; north_star "within 1e-2 relative").
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
DEV = "cuda"
REL = 1e-2


def _oracle():
    from oracle import aria_oracle as O
    from oracle import configs as C
    return O, C


def rel_inf(got, want):
    got, want = got.float().cpu(), want.float().cpu()
    return float((got - want).abs().max() / want.abs().max().clamp_min(1e-12))


# ------------------------------------------------------------------------------------------------ forced routing
def test_whole_model_with_oracle_routing_every_token_within_tolerance():
    """ViT -> projector -> merge -> 2-layer MoE LM; each MoE layer takes the oracle's expert choice for that layer.
    With routing pinned, NO token may be filtered out: all must be within REL of the logit scale."""
    O, C = _oracle()
    from aria_b200.modeling_aria import AriaConfig, AriaForConditionalGeneration
    cfg = C.TINY
    k = cfg["text_config"]["moe_topk"]
    sd = C.aria_state(cfg, seed=0, dtype=torch.bfloat16)
    model = AriaForConditionalGeneration(AriaConfig.from_dict(cfg), device=DEV)
    model.load_state_dict({n: v.to(DEV) for n, v in sd.items()}, strict=True)
    g = torch.Generator().manual_seed(7)
    S = cfg["vision_config"]["image_size"]
    worst = 0.0
    for trial in range(3):
        pv = torch.randn(2, 3, S, S, generator=g).bfloat16()
        text = torch.randint(10, cfg["text_config"]["vocab_size"], (2, 40), generator=g)
        ids = torch.cat([text[:, :5], torch.full((2, 8), cfg["image_token_index"]), text[:, 5:]], dim=1)
        rl = []
        want, _ = O.aria_forward(ids, pv, None, sd, cfg, router_logits=rl)
        for layer, lg in zip(model.language_model.model.layers, rl):
            layer.mlp.router.forced_top_indices = O.topk_lowest_index(lg, k)[1].to(torch.int32).to(DEV).contiguous()
        got = model(ids, pv, None).logits.float().cpu()
        for layer in model.language_model.model.layers:
            layer.mlp.router.forced_top_indices = None
        scale = float(want.float().abs().max())
        err = (got - want.float()).abs().amax(-1)                       # per token
        worst = max(worst, float(err.max()) / scale)
        assert float(err.max()) <= REL * scale, (trial, float(err.max()) / scale, int(err.argmax()))
        d = got - want.float()
        assert float(d.norm() / want.float().norm()) <= REL
    print(f"forced-routing whole model: worst per-token max-abs/scale {worst:.3e} over 3 x {2 * 48} tokens")


# ------------------------------------------------------------------------------------------------ full-width MoE layer
@pytest.fixture(scope="module")
def full_width_layer():
    """One Aria-25.3B MoE layer (1.74 GB bf16) + the oracle's forward at T=768 (cfg 2's token count), computed once."""
    O, C = _oracle()
    from aria_b200 import moe_lm
    tc = dict(C.ARIA_25B["text_config"])
    gen = torch.Generator().manual_seed(2024)
    sd = {n: v.bfloat16() for n, v in C.moe_layer_state(tc, gen).items()}
    x = torch.randn(1, 768, tc["hidden_size"], generator=gen).bfloat16()
    want, parts = O.moe_layer(x, sd, tc["moe_topk"], return_parts=True)
    layer = moe_lm.MoELayer(moe_lm.AriaMoELMConfig(**tc), device=DEV)
    layer.load_state_dict({n: v.to(DEV) for n, v in sd.items()}, strict=True)
    return tc, sd, x, want, parts, layer


def test_full_width_moe_layer_forced_routing_all_tokens(full_width_layer):
    tc, sd, x, want, parts, layer = full_width_layer
    layer.router.forced_top_indices = parts["top_idx"].to(torch.int32).to(DEV).contiguous()
    try:
        scores, idx, counts = layer.router(x.to(DEV))
        assert torch.equal(counts.cpu().long(), parts["counts"].long())
        assert rel_inf(scores, parts["scores"]) <= 2 ** -7          # scores from OUR logits at the oracle's ids: 1-2 bf16 ulps
        got = layer(x.to(DEV))
    finally:
        layer.router.forced_top_indices = None
    scale = float(want.float().abs().max())
    err = (got.float().cpu() - want.float()).abs().amax(-1).view(-1)
    assert float(err.max()) <= REL * scale, (float(err.max()) / scale, int(err.argmax()))
    # intermediates at full width: expert-sorted rows bit-exact, expert outputs / shared branch within tolerance
    perm = layer.token_dispatcher.token_permutation(x.to(DEV), parts["top_idx"].to(torch.int32).to(DEV).contiguous(),
                                                    parts["counts"].to(torch.int32).to(DEV))
    assert torch.equal(perm.cpu(), parts["permuted"])
    eo = layer.experts(perm, layer.token_dispatcher.expert_offsets)
    assert rel_inf(eo, parts["expert_out"]) <= REL
    assert rel_inf(layer.shared_experts(x.to(DEV)), parts["shared"]) <= REL


def test_full_width_moe_layer_free_routing(full_width_layer):
    """Own routing: expert sets equal on every token whose 6th/7th logit gap exceeds bf16 noise; flips are counted, bounded,
    and only ever exchange near-tied experts.  Tie-free tokens match the oracle within REL."""
    tc, sd, x, want, parts, layer = full_width_layer
    k = tc["moe_topk"]
    scores, idx, counts = layer.router(x.to(DEV))
    lg = parts["logits"].float()
    srt = lg.sort(1, descending=True).values
    safe = (srt[:, k - 1] - srt[:, k]) > 2 ** -6 * srt.abs().amax(1)
    same = (idx.cpu().long().sort(1).values == parts["top_idx"].sort(1).values).all(1)
    assert bool(same[safe].all()), "routing differs on a token that is not a bf16 near-tie"
    assert float((~same).float().mean()) <= 0.10
    got = layer(x.to(DEV))
    scale = float(want.float().abs().max())
    err = (got.float().cpu() - want.float()).abs().amax(-1).view(-1)
    assert float(err[safe].max()) <= REL * scale
    print(f"full-width free routing: {int((~safe).sum())}/768 near-tie tokens, {int((~same).sum())} routed differently, "
          f"max err/scale on tie-free tokens {float(err[safe].max()) / scale:.3e}, on all tokens {float(err.max()) / scale:.3e}")


# ------------------------------------------------------------------------------------------------ ViT attention, full image
@pytest.mark.parametrize("masked", [False, True])
def test_vit_attention_full_image_hd72(masked):
    """(B,H,N,hd) = (1,16,4900,72): 19 full query pairs + a ragged 20th (4900 = 19*256 + 36: second tile of the last pair
    empty), keys padded 4900 -> 39 blocks with a 36-key tail; optional key mask = right/bottom 25 % of the image padded."""
    O, _ = _oracle()
    from aria_b200 import ops
    g = torch.Generator().manual_seed(72)
    B, H, N, hd = 1, 16, 4900, 72
    q = torch.zeros(B, H, N, 128, dtype=torch.bfloat16)
    k, v = torch.zeros_like(q), torch.zeros_like(q)
    for t in (q, k, v):
        t[..., :hd] = torch.randn(B, H, N, hd, generator=g).bfloat16()
    km, add = None, None
    if masked:
        valid = torch.zeros(70, 70, dtype=torch.bool)
        valid[:52, :52] = True
        km = (~valid).reshape(1, N)
        add = torch.zeros(B, 1, 1, N, dtype=torch.bfloat16).masked_fill_(km[:, None, None, :], torch.finfo(torch.bfloat16).min)
    want = O.attention_core(q[..., :hd], k[..., :hd], v[..., :hd], hd ** -0.5, add).reshape(B, N, H * hd)
    got = ops.attention(q.to(DEV), k.to(DEV), v.to(DEV), N, N, hd ** -0.5, False, out_hd=hd,
                        key_mask=None if km is None else km.to(torch.uint8).to(DEV))
    assert torch.isfinite(got.float()).all()
    # Each output averages ~4900 values: the reference path (transformers eager attention on bf16 tensors) rounds S and P to
    # bf16 on the way, which alone puts IT ~1e-2 of the output scale away from exact arithmetic on the same bf16 inputs (measured
    # below).  The kernel keeps S and the softmax statistics in fp32, so the statement that can hold is: within the stated
    # tolerance in the L2 sense, never farther than 2.5 x REL element-wise, and at least as close to exact math as the reference.
    qf, kf, vf = (t[..., :hd].float() for t in (q, k, v))
    s = qf @ kf.transpose(2, 3) * hd ** -0.5
    if km is not None:
        s = s.masked_fill(km[:, None, None, :], float("-inf"))
    exact = (s.softmax(-1) @ vf).transpose(1, 2).reshape(B, N, H * hd)
    g32, w32 = got.float().cpu(), want.float()
    assert float((g32 - w32).norm() / w32.norm()) <= REL
    assert rel_inf(got, want) <= 2.5 * REL
    assert rel_inf(got, exact) <= 2 ** -7
    assert rel_inf(got, exact) <= rel_inf(want, exact)
    print(f"ViT attention N=4900 masked={masked}: vs reference-semantics oracle max {rel_inf(got, want):.3e} / L2 "
          f"{float((g32 - w32).norm() / w32.norm()):.3e}; vs exact: ours {rel_inf(got, exact):.3e}, oracle {rel_inf(want, exact):.3e}")


# ------------------------------------------------------------------------------------------------ long causal attention
@pytest.mark.parametrize("T,H", [(8192, 20), (65536, 4)])
def test_causal_attention_long_context_sampled_rows(T, H):
    """64K context (BASELINE cfg 4) cannot be checked densely on the CPU; every query row is independent, so sample rows —
    tile edges, the diagonal block boundaries, first/last — and check each against an fp32 oracle row (O(T) work)."""
    from aria_b200 import ops
    g = torch.Generator(device=DEV).manual_seed(T)
    q = torch.randn(1, H, T, 128, generator=g, device=DEV).bfloat16()
    k = torch.randn(1, H, T, 128, generator=g, device=DEV).bfloat16()
    v = torch.randn(1, H, T, 128, generator=g, device=DEV).bfloat16()
    got = ops.attention(q, k, v, T, T, 128 ** -0.5, True).view(T, H, 128)
    gi = torch.Generator().manual_seed(1)
    rows = sorted(set([0, 1, 127, 128, 129, 255, 256, 257, 383, 384, 511, 512, T // 2 - 1, T // 2, T - 257, T - 256, T - 129,
                       T - 128, T - 2, T - 1] + torch.randint(0, T, (44,), generator=gi).tolist()))
    qc, kc, vc = q[0].float().cpu(), k[0].float().cpu(), v[0].float().cpu()
    worst = 0.0
    for i in rows:
        s = torch.einsum("hd,hkd->hk", qc[:, i], kc[:, : i + 1]) * 128 ** -0.5
        want = torch.einsum("hk,hkd->hd", s.softmax(-1), vc[:, : i + 1])            # [H, 128]
        e = float((got[i].float().cpu() - want).abs().max() / want.abs().max())
        worst = max(worst, e)
        assert e <= REL, (i, e)
    print(f"causal T={T}: {len(rows)} sampled rows x {H} heads, worst rel err {worst:.3e}")


# ------------------------------------------------------------------------------------------------ decode attention
@pytest.mark.parametrize("B,H,Tk,masked", [(4, 20, 777, False), (32, 20, 2048, False), (3, 2, 300, True), (2, 2, 1, False)])
def test_decode_attention_vs_oracle(B, H, Tk, masked):
    O, _ = _oracle()
    from aria_b200 import ops
    g = torch.Generator().manual_seed(B * 1000 + Tk)
    q = torch.randn(B, H, 1, 128, generator=g).bfloat16()
    k = torch.randn(B, H, Tk, 128, generator=g).bfloat16()
    v = torch.randn(B, H, Tk, 128, generator=g).bfloat16()
    km, add = None, None
    if masked:   # left padding of different lengths per sequence (HF generation convention)
        km = torch.zeros(B, Tk, dtype=torch.bool)
        for b in range(B):
            km[b, : 17 * b + 5] = True
        add = torch.zeros(B, 1, 1, Tk, dtype=torch.bfloat16).masked_fill_(km[:, None, None, :], torch.finfo(torch.bfloat16).min)
    want = O.attention_core(q, k, v, 128 ** -0.5, add).reshape(B, H * 128)
    got = ops.attention_decode(q[:, :, 0].contiguous().to(DEV), k.to(DEV), v.to(DEV), Tk, 128 ** -0.5,
                               key_mask=None if km is None else km.to(torch.uint8).to(DEV))
    # same remark as for the ViT shape: over 2048 keys the reference's bf16 rounding of S / P is ~1e-2 of the output scale by itself
    s = (q.float() @ k.float().transpose(2, 3)) * 128 ** -0.5
    if km is not None:
        s = s.masked_fill(km[:, None, None, :], float("-inf"))
    exact = (s.softmax(-1) @ v.float()).reshape(B, H * 128)
    g32, w32 = got.float().cpu(), want.float()
    assert float((g32 - w32).norm() / w32.norm()) <= REL
    assert rel_inf(got, want) <= 2.5 * REL
    assert rel_inf(got, exact) <= 2 ** -7 and rel_inf(got, exact) <= rel_inf(want, exact) + 2 ** -9


# ------------------------------------------------------------------------------------------------ mirror: masks, chunks, loss
def _tiny_lm():
    O, C = _oracle()
    from aria_b200.modeling_aria import AriaConfig, AriaForConditionalGeneration
    sd = C.aria_state(C.TINY, seed=0, dtype=torch.bfloat16)
    m = AriaForConditionalGeneration(AriaConfig.from_dict(C.TINY), device=DEV)
    m.load_state_dict({n: v.to(DEV) for n, v in sd.items()}, strict=True)
    return m, sd, C.TINY


def test_mirror_padded_batch_equals_unpadded_runs():
    """Right- and left-padded batches (HF 2-D attention_mask; left padding with the position ids GenerationMixin derives) give,
    on the real tokens, the logits of each sequence run alone.  Rows are independent in every kernel except attention, so this
    is exactly the key-mask + position-id plumbing."""
    m, _, cfg = _tiny_lm()
    g = torch.Generator().manual_seed(5)
    a = torch.randint(10, 512, (1, 37), generator=g)
    b = torch.randint(10, 512, (1, 21), generator=g)
    alone_a, alone_b = m(a).logits.float().cpu(), m(b).logits.float().cpu()
    scale = float(alone_a.abs().max())
    # right padding: positions are the plain arange
    ids = torch.zeros(2, 37, dtype=torch.long)
    ids[0], ids[1, :21] = a[0], b[0]
    mask = torch.zeros(2, 37, dtype=torch.long)
    mask[0], mask[1, :21] = 1, 1
    out = m(ids, attention_mask=mask).logits.float().cpu()
    assert float((out[0] - alone_a[0]).abs().max()) <= 4e-3 * scale      # same kernels, same order: only tile-shape noise
    assert float((out[1, :21] - alone_b[0]).abs().max()) <= 4e-3 * scale
    # left padding + derived position ids (prepare_inputs_for_generation)
    ids = torch.zeros(2, 37, dtype=torch.long)
    ids[0], ids[1, 16:] = a[0], b[0]
    mask = torch.zeros(2, 37, dtype=torch.long)
    mask[0], mask[1, 16:] = 1, 1
    inputs = m.prepare_inputs_for_generation(ids, None, attention_mask=mask)
    assert torch.equal(inputs["position_ids"][1, 16:], torch.arange(21))
    out = m(**inputs).logits.float().cpu()
    assert float((out[0] - alone_a[0]).abs().max()) <= 4e-3 * scale
    assert float((out[1, 16:] - alone_b[0]).abs().max()) <= 4e-3 * scale
    # greedy generation of the left-padded batch = each prompt generated alone
    toks = m.generate(ids, attention_mask=mask, max_new_tokens=4)
    ta, tb = m.generate(a, max_new_tokens=4), m.generate(b, max_new_tokens=4)
    assert torch.equal(toks[0, -4:], ta[0, -4:]) and torch.equal(toks[1, -4:], tb[0, -4:])


def test_mirror_chunked_prefill_then_decode_and_cache_overflow():
    """prefill 20 -> a 13-token chunk -> 1-token decode against one cache == one-shot prefill of all 34 tokens (ADVICE r1:
    q[:, :, pos0:] used to be rejected as non-contiguous); a step past T_max raises instead of writing out of bounds."""
    m, sd, cfg = _tiny_lm()
    O, _ = _oracle()
    g = torch.Generator().manual_seed(9)
    ids = torch.randint(10, 512, (2, 34), generator=g)
    full = m(ids).logits.float().cpu()
    o1 = m(ids[:, :20], max_cache_len=34)
    o2 = m(ids[:, 20:33], past_key_values=o1.past_key_values)
    o3 = m(ids[:, 33:], past_key_values=o2.past_key_values)
    got = torch.cat([o1.logits, o2.logits, o3.logits], dim=1).float().cpu()
    scale = float(full.abs().max())
    assert float((got - full).abs().max()) <= REL * scale
    emb = F.embedding(ids, sd["language_model.model.embed_tokens.weight"])
    want, _ = O.lm_forward(emb, sd, cfg["text_config"])
    assert float((got - want.float()).norm() / want.float().norm()) <= 2e-2
    assert o3.past_key_values.seq_len == 34
    with pytest.raises(RuntimeError, match="KV cache overflow"):
        m(ids[:, :1], past_key_values=o3.past_key_values)


def test_mirror_labels_loss_and_rejected_arguments():
    m, sd, cfg = _tiny_lm()
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(10, 512, (2, 19), generator=g)
    mask = torch.ones(2, 19, dtype=torch.long)
    mask[1, 15:] = 0
    out = m(ids, attention_mask=mask, labels=ids)
    lg = out.logits.float()
    sm = mask[:, 1:].to(DEV) != 0
    want = F.cross_entropy(lg[:, :-1][sm], ids.to(DEV)[:, 1:][sm])
    assert abs(float(out.loss) - float(want)) <= 1e-5
    with pytest.raises(NotImplementedError):
        m(ids, output_attentions=True)
    with pytest.raises(ValueError):
        m(ids, attention_mask=torch.ones(2, 7, dtype=torch.long))
    e = torch.randn(1, 5, 256, device=DEV).bfloat16()
    keep = e.clone()
    S = cfg["vision_config"]["image_size"]
    idsi = torch.tensor([[11, 9, 9, 9, 9, 9, 9, 9, 9, 12]])
    e10 = torch.randn(1, 10, 256, device=DEV).bfloat16()
    keep10 = e10.clone()
    m(idsi, torch.randn(1, 3, S, S).bfloat16(), None, inputs_embeds=e10)
    assert torch.equal(e10, keep10), "a caller-supplied inputs_embeds must not be modified in place"
    assert torch.equal(e, keep)
