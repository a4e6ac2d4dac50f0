"""CPU: the oracle's DINOv2 block arithmetic (oracle/vit.py, parity UNPINNED against the reference because the block code
lives in facebookresearch/dinov2 via torch.hub) cross-checked against an independent implementation that IS in this image:
``transformers.models.dinov2``.  Same random weights, same token input -> the outputs of every block must agree to fp32
round-off.  This pins the block math (pre-LN, eps 1e-6, 1/sqrt(head_dim) scaling, LayerScale, exact-GELU MLP) to a second
source; the stride-7 patch embedding, the position-embedding interpolation and the tap point are the reference's own code
and are restated from its source."""
import pytest
import torch

from oracle import vit as ovit

tfm = pytest.importorskip("transformers")


def _hf_layer(dim, heads, sd, i):
    from transformers import Dinov2Config
    from transformers.models.dinov2.modeling_dinov2 import Dinov2Layer
    cfg = Dinov2Config(hidden_size=dim, num_attention_heads=heads, num_hidden_layers=1, mlp_ratio=4, layer_norm_eps=1e-6,
                       hidden_act="gelu", layerscale_value=1.0, use_swiglu_ffn=False, qkv_bias=True,
                       attention_probs_dropout_prob=0.0, hidden_dropout_prob=0.0, drop_path_rate=0.0)
    cfg._attn_implementation = "eager"
    layer = Dinov2Layer(cfg).eval()
    p = f"blocks.{i}."
    qkv_w, qkv_b = sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"]
    own = layer.state_dict()
    mapped = {
        "norm1.weight": sd[p + "norm1.weight"], "norm1.bias": sd[p + "norm1.bias"],
        "norm2.weight": sd[p + "norm2.weight"], "norm2.bias": sd[p + "norm2.bias"],
        "attention.attention.query.weight": qkv_w[:dim], "attention.attention.query.bias": qkv_b[:dim],
        "attention.attention.key.weight": qkv_w[dim:2 * dim], "attention.attention.key.bias": qkv_b[dim:2 * dim],
        "attention.attention.value.weight": qkv_w[2 * dim:], "attention.attention.value.bias": qkv_b[2 * dim:],
        "attention.output.dense.weight": sd[p + "attn.proj.weight"], "attention.output.dense.bias": sd[p + "attn.proj.bias"],
        "layer_scale1.lambda1": sd[p + "ls1.gamma"], "layer_scale2.lambda1": sd[p + "ls2.gamma"],
        "mlp.fc1.weight": sd[p + "mlp.fc1.weight"], "mlp.fc1.bias": sd[p + "mlp.fc1.bias"],
        "mlp.fc2.weight": sd[p + "mlp.fc2.weight"], "mlp.fc2.bias": sd[p + "mlp.fc2.bias"],
    }
    assert set(mapped) == set(own), (sorted(set(own) - set(mapped)), sorted(set(mapped) - set(own)))
    layer.load_state_dict(mapped)
    return layer


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
