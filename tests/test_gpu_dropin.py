"""GPU: the drop-in surface exercised on UNMODIFIED model classes (round-1 VERDICT "missing 1"):

  A. the reference's own classes (aria/model/*.py, loaded by oracle/ref_loader.py from /root/reference or from the byte-for-byte
     staging in oracle/_ref/): `install.install()` on a real `MoELayer` (seam 2), `experts_gemm` rebound alone (seam 1, CPU int64
     counts as the reference passes them), and a whole reference `AriaForConditionalGeneration` with all seams
     (`install`, `install_vit`, `hf_attention.register`) — forward() prefill + cached decode steps vs the same model eager;
  B. transformers' own `models.aria.AriaForConditionalGeneration` (present in this image, same sub-module layout): seams
     installed, `forward()` logits and `generate()` (HF GenerationMixin, greedy) vs the same model eager, incl. a padded batch.

"Eager" = the unmodified model in bf16 on the same GPU (torch/ATen kernels, the reference's `sequential_gemm`, HF eager attention).
Tolerances as in tests/test_gpu_parity.py (whole-model: rel-L2 <= 1e-2, element-wise <= 2e-2 of the logit scale on tokens whose
router top-k is not a bf16 near-tie; eager and seamed paths may break such ties differently).
"""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
DEV = "cuda"
REL = 1e-2


def _ref():
    from oracle import ref_loader
    if not ref_loader.reference_available():
        pytest.skip("reference files neither under /root/reference nor staged in oracle/_ref (run oracle/build_ref.py)")
    return ref_loader.load_reference()


def _margin(router_logits, k):
    from oracle import aria_oracle as O
    return O.topk_margin([lg.float().cpu() for lg in router_logits], k)


def _check_logits(got, want, router_logits, k, max_tie_frac=0.3):
    got, want = got.float().cpu(), want.float().cpu()
    V = want.shape[-1]
    g2, w2 = got.reshape(-1, V), want.reshape(-1, V)
    safe = _margin(router_logits, k) > 2 ** -6
    assert safe.numel() == w2.shape[0]
    assert float((~safe).float().mean()) <= max_tie_frac
    scale = float(w2.abs().max())
    err = (g2 - w2).abs().amax(-1)
    rel_l2 = float((g2 - w2)[safe].norm() / w2[safe].norm())
    assert torch.isfinite(g2).all()
    assert rel_l2 <= REL, rel_l2
    assert float(err[safe].max()) <= 2 * REL * scale, float(err[safe].max()) / scale
    assert float(err.max()) <= 0.5 * scale
    return rel_l2, float(err[safe].max()) / scale, int((~safe).sum())


# ------------------------------------------------------------------------------------------------ A. reference classes
@pytest.mark.parametrize("d,E,k,I,T", [(256, 8, 2, 512, 32), (256, 64, 6, 128, 300), (2560, 64, 6, 1664, 768)])
def test_install_on_reference_moe_layer(d, E, k, I, T):
    """BASELINE cfg 1 (d=256, 8 experts, top-2), a 64-expert variant and ONE FULL-WIDTH layer: the reference `MoELayer` run by the
    reference on the GPU, then the same module object after `install.install()`; then only `experts_gemm` swapped (seam 1)."""
    ref = _ref()
    from aria_b200 import install
    from aria_b200 import moe_lm as ours
    cfg = ref.moe_lm.AriaMoELMConfig(hidden_size=d, num_attention_heads=max(2, d // 128), moe_num_experts=E, moe_topk=k,
                                     moe_intermediate_size=I, moe_num_shared_experts=2, intermediate_size=I)
    layer = ref.moe_lm.MoELayer(cfg)
    g = torch.Generator().manual_seed(d + E)
    for p_ in layer.parameters():                      # `torch.empty` + FIXME in the reference (moe_lm.py:185-188,465)
        p_.data = (torch.randn(p_.shape, generator=g) * 0.02)
    layer = layer.to(DEV, torch.bfloat16).eval()
    x = torch.randn(2, T // 2, d, generator=g).bfloat16().to(DEV)
    logits = []
    orig_routing = layer.router.routing
    layer.router.routing = lambda lg: (logits.append(lg.detach()), orig_routing(lg))[1]
    orig_gemm = ref.moe_lm.experts_gemm
    try:
        want = layer(x)                                # the reference's own forward: ATen + sequential_gemm
        # seam 1 only: the reference forward with OUR gmm (called with CPU int64 counts, moe_lm.py:478-484)
        ref.moe_lm.experts_gemm = ours.experts_gemm
        seam1 = layer(x)
    finally:
        ref.moe_lm.experts_gemm = orig_gemm
        layer.router.routing = orig_routing
    assert install.install(torch.nn.ModuleList([layer])) == 1
    got = layer(x)
    lg = logits[0].float().cpu()
    srt = lg.sort(1, descending=True).values
    safe = (srt[:, k - 1] - srt[:, k]) > 2 ** -6 * srt.abs().amax(1)
    scale = float(want.float().abs().max())
    for name, y in (("seam1", seam1), ("install", got)):
        err = (y.float() - want.float()).abs().amax(-1).view(-1).cpu()
        sel = safe if name == "install" else torch.ones_like(safe)     # seam 1 keeps the reference's own routing: every token
        assert float(err[sel].max()) <= REL * scale, (name, float(err[sel].max()) / scale)
    # the seams are inference-only: under autograd they must refuse, not cut the graph silently
    with torch.enable_grad():
        with pytest.raises(RuntimeError, match="inference-only"):
            layer(x.clone().requires_grad_(True))
        layer.train()
        with pytest.raises(RuntimeError, match="inference-only"):
            layer(x)
        layer.eval()


def _reference_tiny_model():
    ref = _ref()
    from oracle import configs as C
    from oracle.make_golden import build_reference_model
    sd = C.aria_state(C.TINY, seed=0, dtype=torch.float32)
    model = build_reference_model(ref, C.TINY, sd, torch.bfloat16).to(DEV)
    rot = model.language_model.model.rotary_emb           # keep the fp32 inv_freq on the device (see build_reference_model)
    rot.inv_freq = rot.inv_freq.float().to(DEV)
    return ref, model, C.TINY


def _hook_router_logits(model, store):
    hs = []
    for m in model.modules():
        if type(m).__name__ == "TopKRouter":
            orig = m.routing
            m.routing = (lambda o: (lambda lg: (store.append(lg.detach()), o(lg))[1]))(orig)
            hs.append((m, orig))
    return hs


def test_reference_model_with_all_seams_forward_and_cached_decode():
    """The reference `AriaForConditionalGeneration` (tiny dims, real head dims 128 / 72) — forward() with image + text, then two
    cached decode steps through its own `past_key_values` — eager vs all three seams installed on the SAME module object."""
    ref, model, cfg = _reference_tiny_model()
    from aria_b200 import hf_attention, install
    k = cfg["text_config"]["moe_topk"]
    g = torch.Generator().manual_seed(2)
    S = cfg["vision_config"]["image_size"]
    pv = torch.randn(2, 3, S, S, generator=g).bfloat16().to(DEV)
    pm = torch.ones(2, S, S, dtype=torch.bool)
    pm[1, 28:, :] = False                                  # second image: bottom half padded (8 of 16 patches are keys)
    text = torch.randint(10, cfg["text_config"]["vocab_size"], (1, 30), generator=g)
    ids = torch.cat([text[:, :3], torch.full((1, 8 + 8), cfg["image_token_index"]), text[:, 3:]], dim=1).to(DEV)
    pm = pm.to(DEV)

    def run():
        rl = []
        hooks = _hook_router_logits(model, rl)
        try:
            out = model(input_ids=ids, pixel_values=pv, pixel_mask=pm, use_cache=True)
            logits = [out.logits]
            past = out.past_key_values
            tok = out.logits[:, -1].argmax(-1, keepdim=True)
            toks = [tok]
            for _ in range(2):
                o = model(input_ids=tok, past_key_values=past, use_cache=True)
                past = o.past_key_values
                logits.append(o.logits)
                tok = o.logits[:, -1].argmax(-1, keepdim=True)
                toks.append(tok)
        finally:
            for m, orig in hooks:
                m.routing = orig
        return logits, torch.cat(toks, 1), rl

    want, want_toks, rl = run()
    n_layers = cfg["text_config"]["num_hidden_layers"]
    assert len(rl) == 3 * n_layers
    assert install.install(model, ref.moe_lm) == n_layers
    assert install.install_vit(model) == cfg["vision_config"]["num_hidden_layers"]
    model.config.text_config._attn_implementation = hf_attention.register()
    model.language_model.config._attn_implementation = hf_attention.IMPL_KEY
    got, got_toks, _ = run()
    T = ids.shape[1]
    r = _check_logits(got[0], want[0], rl[:n_layers], k)
    print(f"reference model + seams, prefill: rel-L2 {r[0]:.3e} max/scale {r[1]:.3e} ({r[2]}/{T} near-tie tokens)")
    for s in (1, 2):
        if torch.equal(got_toks[:, :s], want_toks[:, :s]):   # same prefix -> the decode step saw the same cache
            _check_logits(got[s], want[s], rl[s * n_layers:(s + 1) * n_layers], k, max_tie_frac=1.0)


# ------------------------------------------------------------------------------------------------ B. transformers' own Aria
def _install_all(model):
    from aria_b200 import hf_attention, install
    n_moe = install.install(model)
    n_vit = install.install_vit(model)
    key = hf_attention.register()
    model.config.text_config._attn_implementation = key
    model.model.language_model.config._attn_implementation = key
    return n_moe, n_vit


def _hf_router_hooks(model, store):
    return [m.router.register_forward_hook(lambda mod, inp, out: store.append(out.detach()))
            for m in model.modules() if type(m).__name__ == "AriaTextMoELayer"]


def test_hf_aria_forward_and_generate_with_seams():
    """Unmodified `transformers.models.aria.AriaForConditionalGeneration`: forward() logits and `generate()` (GenerationMixin's
    greedy loop, its DynamicCache, its prepare_inputs_for_generation) equal to the same model eager."""
    import hf_common as H
    model = H.tiny_hf_aria(DEV, torch.bfloat16)
    ids, pv, pm = H.tiny_inputs(batch=2, pad_last_image=False)
    ids, pv, pm = ids.to(DEV), pv.bfloat16().to(DEV), pm.to(DEV)
    rl = []
    hooks = _hf_router_hooks(model, rl)
    want = model(input_ids=ids, pixel_values=pv, pixel_mask=pm).logits
    for h in hooks:
        h.remove()
    gen_kw = dict(max_new_tokens=6, do_sample=False, output_scores=True, return_dict_in_generate=True)
    want_gen = model.generate(input_ids=ids, pixel_values=pv, pixel_mask=pm, **gen_kw)
    n_moe, n_vit = _install_all(model)
    assert (n_moe, n_vit) == (2, 2)
    got = model(input_ids=ids, pixel_values=pv, pixel_mask=pm).logits
    r = _check_logits(got, want, rl, model.config.text_config.moe_topk)
    print(f"HF Aria + seams, forward: rel-L2 {r[0]:.3e} max/scale {r[1]:.3e} ({r[2]} near-tie tokens)")
    got_gen = model.generate(input_ids=ids, pixel_values=pv, pixel_mask=pm, **gen_kw)
    T0 = ids.shape[1]
    for b in range(ids.shape[0]):
        for s in range(6):
            a, c = int(want_gen.sequences[b, T0 + s]), int(got_gen.sequences[b, T0 + s])
            if a != c:   # allowed only when the eager logits themselves were a near-tie between the two tokens; then stop
                sc = want_gen.scores[s][b].float()
                assert float(sc[a] - sc[c]) <= 2 * REL * float(sc.abs().max()), (b, s, a, c)
                break


def test_hf_aria_padded_batch_through_generate():
    """Two prompts of different lengths, LEFT-padded (HF generation convention): the 2-D attention_mask reaches seam 2 as the
    kernels' key mask (prefill and every decode step); tokens equal to eager generation of the same padded batch."""
    import hf_common as H
    model = H.tiny_hf_aria(DEV, torch.bfloat16, seed=1)
    g = torch.Generator().manual_seed(11)
    a = torch.randint(10, 512, (30,), generator=g)
    b = torch.randint(10, 512, (17,), generator=g)
    ids = torch.zeros(2, 30, dtype=torch.long)
    ids[0], ids[1, 13:] = a, b
    mask = torch.zeros(2, 30, dtype=torch.long)
    mask[0], mask[1, 13:] = 1, 1
    ids, mask = ids.to(DEV), mask.to(DEV)
    kw = dict(max_new_tokens=5, do_sample=False, output_scores=True, return_dict_in_generate=True)
    want = model.generate(input_ids=ids, attention_mask=mask, **kw)
    want_logits = model(input_ids=ids, attention_mask=mask).logits
    _install_all(model)
    got_logits = model(input_ids=ids, attention_mask=mask).logits
    scale = float(want_logits.float().abs().max())
    real = mask.bool().cpu()
    err = (got_logits.float() - want_logits.float()).abs().amax(-1).cpu()
    assert float(err[real].median()) <= REL * scale          # padded positions are don't-care; near-tie tokens may differ more
    assert float((err[real] <= 2 * REL * scale).float().mean()) >= 0.7
    got = model.generate(input_ids=ids, attention_mask=mask, **kw)
    for r in range(2):
        for s in range(5):
            x, y = int(want.sequences[r, 30 + s]), int(got.sequences[r, 30 + s])
            if x != y:
                sc = want.scores[s][r].float()
                assert float(sc[x] - sc[y]) <= 2 * REL * float(sc.abs().max()), (r, s, x, y)
                break
