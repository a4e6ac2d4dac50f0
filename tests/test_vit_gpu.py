"""GPU test: tcgen05 ViT (TF32 tensor-core GEMMs, materialised attention) against the fp32 oracle restatement
(parity unpinned: the DINOv2 block arithmetic is third-party, see oracle/vit.py)."""
import numpy as np
import pytest
import torch

from oracle import vit as ovit
from oracle import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("attention", ["fused", "fused-single-cta", "materialized"])
@pytest.mark.parametrize("cfg", [dict(H=98, W=126, depth=2, dim=128, heads=2, layer=1, T=3),
                                 dict(H=112, W=140, depth=3, dim=192, heads=3, layer=1, T=2),
                                 dict(H=182, W=238, depth=1, dim=64, heads=1, layer=0, T=1)])
def test_vit_features_match_oracle(cfg, attention):
    from dino_tracker_b200.vit import DinoV2Features
    g = torch.Generator().manual_seed(3)
    sd = ovit.random_state_dict(cfg["depth"], cfg["dim"], g, n_pos=4, std=0.05)
    video = synth.random_video(cfg["T"], cfg["H"], cfg["W"], seed=4)
    ref = ovit.dino_features_video(video, sd, cfg["heads"], cfg["layer"])          # T x C x h x w
    ex = DinoV2Features(sd, heads=cfg["heads"], layer=cfg["layer"], device="cuda:0",
                        attention="fused" if attention.startswith("fused") else attention,
                        cta_pairs=attention == "fused")
    got = ex.features_chw(video).cpu()
    assert got.shape == ref.shape
    scale = ref.abs().max().item()
    err = (got - ref).abs().max().item()
    print(f"ViT[{attention}] max |diff| = {err:.3e} (max |ref| = {scale:.3f})")
    assert err <= 5e-3 * scale   # TF32 single-pass contractions (10-bit mantissa inputs), fp32 accumulation
    # cosine between corresponding tokens
    cos = torch.nn.functional.cosine_similarity(got.flatten(2), ref.flatten(2), dim=1)
    assert cos.min().item() > 0.9999


def test_pos_embed_interpolation_matches_oracle():
    from dino_tracker_b200.vit import interpolate_pos_embed
    pe = torch.randn(1, 1 + 37 * 37, 32)
    assert torch.equal(interpolate_pos_embed(pe, 67, 121), ovit.interpolate_pos_embed(pe, 67, 121))


def test_vit_is_deterministic_and_batch_invariant():
    """Frame features must not depend on which other frames share the call (frame sharding relies on it)."""
    from dino_tracker_b200.vit import DinoV2Features
    g = torch.Generator().manual_seed(5)
    sd = ovit.random_state_dict(2, 128, g, n_pos=4, std=0.05)
    video = synth.random_video(5, 98, 126, seed=6)
    ex = DinoV2Features(sd, heads=2, layer=1, device="cuda:0")
    a = ex(video).clone()
    b = ex(video).clone()
    assert torch.equal(a, b), f"non-deterministic: {(a - b).abs().max().item()}"
    c = ex(video[1:4]).clone()
    assert torch.equal(a[1:4], c), f"batch-dependent: {(a[1:4] - c).abs().max().item()}"
