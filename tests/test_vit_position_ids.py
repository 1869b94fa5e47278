"""CPU: NaViT position ids of the vision tower at the REAL grid (70 patches per side, 980 px).  The reference pins transformers
4.46.3, whose Idefics2VisionEmbeddings buckets fp32 coordinates `arange(0, 1 - 1e-6, 1 / nb)`; transformers 5.x casts the
coordinates to the pixel dtype first, which in bf16 moves about half of the buckets by one.  Product and oracle must
both follow the 4.46.3 arithmetic; the tiny golden fixtures (4 patches per side) cannot see the difference, hence this test."""
import pytest
import torch

N_SIDE = 70


def _ids_4463(nb_h, nb_w, n=N_SIDE):
    """The loop body of transformers 4.46.3 `Idefics2VisionEmbeddings.forward`, restated with tensor nb (as there)."""
    boundaries = torch.arange(1 / n, 1.0, 1 / n)
    nb_h, nb_w = torch.tensor(nb_h), torch.tensor(nb_w)
    fh = torch.arange(0, 1 - 1e-6, 1 / nb_h)
    fw = torch.arange(0, 1 - 1e-6, 1 / nb_w)
    bh = torch.bucketize(fh, boundaries, right=True)
    bw = torch.bucketize(fw, boundaries, right=True)
    return (bh[:, None] * n + bw).flatten()


def _mask(nb_h, nb_w, n=N_SIDE):
    m = torch.zeros(1, n, n, dtype=torch.bool)
    m[0, :nb_h, :nb_w] = True
    return m


@pytest.mark.parametrize("nb_h,nb_w", [(70, 70), (35, 35), (50, 70), (70, 35), (1, 1), (69, 2)])
def test_product_and_oracle_follow_transformers_4_46_3(nb_h, nb_w):
    from aria_b200.vision_encoder import AriaVisionConfig, Idefics2VisionEmbeddings
    from oracle import aria_oracle as O
    emb = Idefics2VisionEmbeddings(AriaVisionConfig(hidden_size=8, image_size=980, patch_size=14), device="cpu")
    pm = _mask(nb_h, nb_w)
    want = torch.zeros(N_SIDE * N_SIDE, dtype=torch.int64)
    want[pm.reshape(-1)] = _ids_4463(nb_h, nb_w)
    got = emb.position_ids(pm, 1, "cpu")
    assert torch.equal(got, want)
    assert torch.equal(O.vit_position_ids(pm, N_SIDE)[0], want)
    if (nb_h, nb_w) == (70, 70):   # the no-mask fast path is the same table
        assert torch.equal(emb.position_ids(None, 1, "cpu"), want)
        assert torch.equal(emb.position_ids(None, 2, "cpu"), want.repeat(2))


def test_ids_are_a_permutation_free_monotone_grid_and_differ_from_bf16_coordinates():
    ids = _ids_4463(70, 70).view(70, 70)
    assert int(ids.min()) == 0 and int(ids.max()) <= 70 * 70 - 1
    assert bool((ids[1:, 0] >= ids[:-1, 0]).all()) and bool((ids[0, 1:] >= ids[0, :-1]).all())
    # what the transformers-5.x formulation gives with bf16 pixel values: documented difference, must NOT be what we compute
    boundaries = torch.arange(1 / 70, 1.0, 1 / 70)
    f = (torch.arange(70, dtype=torch.float32) * (1.0 / 70)).clamp(max=1 - 1e-6).bfloat16()
    b_bf16 = torch.bucketize(f, boundaries, right=True)
    b_fp32 = torch.bucketize(torch.arange(0, 1 - 1e-6, 1 / torch.tensor(70)), boundaries, right=True)
    assert int((b_bf16 != b_fp32).sum()) >= 20


def test_non_rectangular_mask_is_rejected_like_the_reference():
    from aria_b200.vision_encoder import AriaVisionConfig, Idefics2VisionEmbeddings
    emb = Idefics2VisionEmbeddings(AriaVisionConfig(hidden_size=8, image_size=980, patch_size=14), device="cpu")
    pm = _mask(10, 10)
    pm[0, 5, 5] = False
    with pytest.raises((RuntimeError, IndexError)):
        emb.position_ids(pm, 1, "cpu")
