"""Oracle restatement of Delta-DINO + feature alignment.

Test infrastructure (see ``oracle/__init__.py``).  Functional, driven by a state dict
with the reference's keys (``layers.{0,4,8,12}.{weight,bias}`` convs,
``layers.{1,5,9,13}.*`` BatchNorm, ``layers.{3,7,11}.filt`` BlurPool buffers).
"""
import torch
import torch.nn.functional as F

CONV_IDX = (0, 4, 8, 12)
BN_IDX = (1, 5, 9, 13)
DILATIONS = (1, 1, 1, 2)  # models/networks/delta_dino.py:10
BN_EPS = 1e-5


def blur_pool(x: torch.Tensor) -> torch.Tensor:
    """antialiased_cnns.BlurPool(C, stride=2) default (filt_size 4): reflect pad
    (left 1, right 2, top 1, bottom 2), depthwise outer([1,3,3,1])/64, stride 2
    (third-party adobe/antialiased-cnns; call site models/networks/delta_dino.py:44)."""
    C = x.shape[1]
    a = torch.tensor([1.0, 3.0, 3.0, 1.0], device=x.device)
    filt = a[:, None] * a[None, :]
    filt = (filt / filt.sum())[None, None].repeat(C, 1, 1, 1)
    return F.conv2d(F.pad(x, (1, 2, 1, 2), mode="reflect"), filt, stride=2, groups=C)


def delta_cnn(frames: torch.Tensor, sd: dict, bn_training: bool = False) -> torch.Tensor:
    """models/networks/delta_dino.py:22-46,53-55: [conv5x5 reflect -> BN(eval) -> ReLU ->
    BlurPool] x3, then conv5x5 dilation 2 (reflect pad 4) -> BN.  frames: B x 3 x H x W
    raw [0,1] RGB (no ImageNet normalisation: models/tracker.py:115).  ``bn_training``: BatchNorm on the batch
    statistics of ``frames`` (the module in train mode, dino_tracker.py:133-134; running statistics are not updated here)."""
    x = frames
    for li, (ci, bi, dil) in enumerate(zip(CONV_IDX, BN_IDX, DILATIONS)):
        w = sd[f"layers.{ci}.weight"]
        k = w.shape[-1]
        pad = (k + (k - 1) * (dil - 1)) // 2
        x = F.conv2d(F.pad(x, (pad,) * 4, mode="reflect"), w, sd[f"layers.{ci}.bias"], dilation=dil)
        if bn_training:
            x = F.batch_norm(x, None, None, sd[f"layers.{bi}.weight"], sd[f"layers.{bi}.bias"], training=True, eps=BN_EPS)
        else:
            x = F.batch_norm(x, sd[f"layers.{bi}.running_mean"], sd[f"layers.{bi}.running_var"],
                             sd[f"layers.{bi}.weight"], sd[f"layers.{bi}.bias"], training=False, eps=BN_EPS)
        if li < 3:
            x = blur_pool(torch.relu(x))
    return x


def align_cnn_to_vit(cnn: torch.Tensor, vit_hw, patch=14, vit_stride=7, cnn_stride=8) -> torch.Tensor:
    """models/utils.py:7-45: bilinear ``grid_sample`` (border, align_corners=True) of the CNN
    map at the ViT token centres; CNN-grid coordinate = (pixel - 0.5) / cnn_stride."""
    vh, vw = vit_hw
    ch, cw = cnn.shape[-2:]
    c_br = [(ch - 1) * cnn_stride, (cw - 1) * cnn_stride]
    vit_x = torch.arange(vw, dtype=torch.float32, device=cnn.device) * vit_stride + patch / 2.0
    vit_y = torch.arange(vh, dtype=torch.float32, device=cnn.device) * vit_stride + patch / 2.0
    gx, gy = torch.meshgrid(-1.0 - (1.0 / c_br[1]) + (2.0 * vit_x / c_br[1]),
                            -1 - (1.0 / c_br[0]) + (2.0 * vit_y / c_br[0]), indexing="xy")
    grid = torch.stack([gx, gy], dim=-1)[None].expand(cnn.shape[0], -1, -1, -1)
    return F.grid_sample(cnn, grid, mode="bilinear", padding_mode="border", align_corners=True)


def refined_features(video: torch.Tensor, dino: torch.Tensor, sd: dict, patch=14, stride=7,
                     batch=8, bn_training: bool = False) -> torch.Tensor:
    """models/tracker.py:113-129 (batches of 8 frames) -> dino + residual, T x C x h x w."""
    res = torch.zeros_like(dino)
    for i in range(0, video.shape[0], batch):
        cnn = delta_cnn(video[i:i + batch], sd, bn_training)
        res[i:i + batch] = align_cnn_to_vit(cnn, dino.shape[-2:], patch, stride, 8)
    return dino + res


def random_state_dict(channels, gen: torch.Generator, last_std=0.01) -> dict:
    """Well-conditioned synthetic weights: default-like conv init, the last conv
    ~N(0, last_std) (the reference zero-inits it, delta_dino.py:32-34), BN stats perturbed."""
    sd = {}
    for li, (ci, bi) in enumerate(zip(CONV_IDX, BN_IDX)):
        cin, cout = channels[li], channels[li + 1]
        bound = 1.0 / (cin * 25) ** 0.5
        if li == 3:
            sd[f"layers.{ci}.weight"] = torch.randn(cout, cin, 5, 5, generator=gen) * last_std
            sd[f"layers.{ci}.bias"] = torch.randn(cout, generator=gen) * last_std
        else:
            sd[f"layers.{ci}.weight"] = (torch.rand(cout, cin, 5, 5, generator=gen) * 2 - 1) * bound
            sd[f"layers.{ci}.bias"] = (torch.rand(cout, generator=gen) * 2 - 1) * bound
        sd[f"layers.{bi}.weight"] = 0.5 + torch.rand(cout, generator=gen)
        sd[f"layers.{bi}.bias"] = torch.randn(cout, generator=gen) * 0.1
        sd[f"layers.{bi}.running_mean"] = torch.randn(cout, generator=gen) * 0.1
        sd[f"layers.{bi}.running_var"] = 0.5 + torch.rand(cout, generator=gen)
        sd[f"layers.{bi}.num_batches_tracked"] = torch.tensor(0, dtype=torch.long)
        if li < 3:
            a = torch.tensor([1.0, 3.0, 3.0, 1.0])
            f = a[:, None] * a[None, :]
            sd[f"layers.{ci + 3}.filt"] = (f / f.sum())[None, None].repeat(cout, 1, 1, 1)
    return sd
