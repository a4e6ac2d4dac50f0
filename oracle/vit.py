"""Oracle restatement of the DINOv2 ViT feature extractor.

Test infrastructure (see ``oracle/__init__.py``).  Pinning: the reference's own code on this
row (everything listed below from models/extractor.py and utils.py) is pinned by golden
vectors produced by the live reference (tests/golden/vit_small.npz, posembed.npz;
oracle/make_golden.py gen_vit_case / gen_posembed_case).  The block arithmetic belongs to
facebookresearch/dinov2 (torch.hub, unpinned ``main``; call site models/extractor.py:26),
which is absent from the reference tree and this image: the block math below restates the
published DINOv2 ``DinoVisionTransformer`` (pre-LN block, LayerNorm eps 1e-6, MHA with scale
head_dim**-0.5, LayerScale, MLP 4x with exact GELU) and is cross-checked against
``transformers``' Dinov2Layer; everything the reference itself defines is restated from its source: stride patch (models/extractor.py:41-55),
pos-embed interpolation (:57-85), tap point = output of block ``layer`` before the final norm
(:112-116,137-150), ImageNet normalisation, cls drop and C x h x w layout (utils.py:44-67).
State-dict keys are the DINOv2 hub names so real checkpoints load unchanged.
"""
import math

import torch
import torch.nn.functional as F

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)

CONFIGS = {  # name: (depth, dim, heads)
    "dinov2_vits14": (12, 384, 6),
    "dinov2_vitb14": (12, 768, 12),
    "dinov2_vitl14": (24, 1024, 16),
}


def random_state_dict(depth, dim, gen, n_pos=37, patch=14, ls_init=1.0, std=0.02):
    def tn(*shape):
        return torch.randn(*shape, generator=gen) * std
    sd = {"cls_token": tn(1, 1, dim), "pos_embed": tn(1, 1 + n_pos * n_pos, dim),
          "mask_token": torch.zeros(1, dim),
          "patch_embed.proj.weight": tn(dim, 3, patch, patch), "patch_embed.proj.bias": tn(dim),
          "norm.weight": torch.ones(dim), "norm.bias": torch.zeros(dim)}
    for i in range(depth):
        p = f"blocks.{i}."
        sd[p + "norm1.weight"] = 1 + tn(dim); sd[p + "norm1.bias"] = tn(dim)
        sd[p + "attn.qkv.weight"] = tn(3 * dim, dim) * 2; sd[p + "attn.qkv.bias"] = tn(3 * dim)
        sd[p + "attn.proj.weight"] = tn(dim, dim); sd[p + "attn.proj.bias"] = tn(dim)
        sd[p + "ls1.gamma"] = torch.full((dim,), ls_init) + tn(dim)
        sd[p + "norm2.weight"] = 1 + tn(dim); sd[p + "norm2.bias"] = tn(dim)
        sd[p + "mlp.fc1.weight"] = tn(4 * dim, dim); sd[p + "mlp.fc1.bias"] = tn(4 * dim)
        sd[p + "mlp.fc2.weight"] = tn(dim, 4 * dim); sd[p + "mlp.fc2.bias"] = tn(dim)
        sd[p + "ls2.gamma"] = torch.full((dim,), ls_init) + tn(dim)
    return sd


def interpolate_pos_embed(pos_embed: torch.Tensor, n_h: int, n_w: int) -> torch.Tensor:
    """models/extractor.py:57-85.  DINOv2 calls it with (w=H_img, h=W_img), so the first
    interpolated axis has ``n_h`` (=67) entries; +0.1 trick, bicubic, align_corners=False,
    recompute_scale_factor=False.  Returns 1 x (1 + n_h*n_w) x D."""
    N = pos_embed.shape[1] - 1
    dim = pos_embed.shape[-1]
    side = int(math.sqrt(N))
    if n_h * n_w == N and n_h == n_w:
        return pos_embed
    cls_pos = pos_embed[:, 0]
    patch_pos = pos_embed[:, 1:].reshape(1, side, side, dim).permute(0, 3, 1, 2)
    w0, h0 = n_h + 0.1, n_w + 0.1
    patch_pos = F.interpolate(patch_pos, scale_factor=(w0 / math.sqrt(N), h0 / math.sqrt(N)),
                              mode="bicubic", align_corners=False, recompute_scale_factor=False)
    assert int(w0) == patch_pos.shape[-2] and int(h0) == patch_pos.shape[-1]
    patch_pos = patch_pos.permute(0, 2, 3, 1).reshape(1, -1, dim)
    return torch.cat((cls_pos[:, None], patch_pos), dim=1)


def block_forward(x, sd, i, heads):
    p = f"blocks.{i}."
    B, N, D = x.shape
    hd = D // heads
    y = F.layer_norm(x, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], eps=1e-6)
    qkv = F.linear(y, sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"])
    qkv = qkv.reshape(B, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * hd ** -0.5, qkv[1], qkv[2]
    attn = torch.softmax(q @ k.transpose(-2, -1), dim=-1)
    y = (attn @ v).transpose(1, 2).reshape(B, N, D)
    y = F.linear(y, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
    x = x + y * sd[p + "ls1.gamma"]
    y = F.layer_norm(x, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], eps=1e-6)
    y = F.linear(y, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])
    y = F.gelu(y)
    y = F.linear(y, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    return x + y * sd[p + "ls2.gamma"]


def vit_tokens(frames01: torch.Tensor, sd: dict, heads: int, layer: int, stride: int = 7,
               patch: int = 14, return_all=False):
    """frames01: B x 3 x H x W in [0, 1].  Returns block-``layer`` output B x (1+h*w) x D."""
    mean = torch.tensor(IMAGENET_MEAN, device=frames01.device)[None, :, None, None]
    std = torch.tensor(IMAGENET_STD, device=frames01.device)[None, :, None, None]
    x = (frames01 - mean) / std  # torchvision Normalize, utils.py:46,55
    x = F.conv2d(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=stride)
    B, D, n_h, n_w = x.shape
    x = x.flatten(2).transpose(1, 2)
    x = torch.cat((sd["cls_token"].expand(B, -1, -1), x), dim=1)
    x = x + interpolate_pos_embed(sd["pos_embed"], n_h, n_w)
    outs = []
    for i in range(layer + 1):
        x = block_forward(x, sd, i, heads)
        outs.append(x)
    return (x, outs) if return_all else x


def dino_features_video(video01: torch.Tensor, sd: dict, heads: int, layer: int, stride: int = 7,
                        patch: int = 14) -> torch.Tensor:
    """utils.py:32-72 (facet 'tokens'): per-frame loop, cls dropped, -> T x C x h x w."""
    T, _, H, W = video01.shape
    ph, pw = 1 + (H - patch) // stride, 1 + (W - patch) // stride
    out = []
    for i in range(T):
        tok = vit_tokens(video01[i:i + 1], sd, heads, layer, stride, patch)
        out.append(tok[0, 1:].reshape(ph, pw, -1).permute(2, 0, 1))
    return torch.stack(out)
