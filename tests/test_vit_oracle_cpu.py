"""CPU: the oracle's DINOv2 block arithmetic (oracle/vit.py; the block code of the reference lives in
facebookresearch/dinov2 via torch.hub and is absent here) cross-checked against an independent implementation that IS in this image:
``transformers.models.dinov2``.  Same random weights, same token input -> the outputs of every block must agree to fp32
round-off.  This pins the block math (pre-LN, eps 1e-6, 1/sqrt(head_dim) scaling, LayerScale, exact-GELU MLP) to a second
source; the stride-7 patch embedding, the position-embedding interpolation and the tap point are the reference's own code
and are restated from its source."""
import pytest
import torch

from oracle import vit as ovit

tfm = pytest.importorskip("transformers")


def _hf_layer(dim, heads, sd, i):
    from oracle.make_golden import hf_dinov2_layer
    return hf_dinov2_layer(dim, heads, sd, i)


@pytest.mark.parametrize("dim,heads,tokens", [(64, 1, 50), (128, 2, 222), (192, 3, 97)])
def test_block_arithmetic_matches_transformers_dinov2(dim, heads, tokens):
    g = torch.Generator().manual_seed(5)
    depth = 2
    sd = ovit.random_state_dict(depth, dim, g, n_pos=4, std=0.08)
    for i in range(depth):   # LayerScale away from 1 so that its placement matters
        sd[f"blocks.{i}.ls1.gamma"] = 0.5 + torch.rand(dim, generator=g)
        sd[f"blocks.{i}.ls2.gamma"] = 0.5 + torch.rand(dim, generator=g)
    x = torch.randn(2, tokens, dim, generator=g)
    with torch.no_grad():
        ref = x
        got = x
        for i in range(depth):
            out = _hf_layer(dim, heads, sd, i)(ref)
            ref = out[0] if isinstance(out, (tuple, list)) else out
            got = ovit.block_forward(got, sd, i, heads)
            assert (got - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item()), i
