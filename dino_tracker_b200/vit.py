"""Host mirror of the reference's DINOv2 feature stage (``models/extractor.py::VitExtractor`` +
``utils.py::get_dino_features_video``) over libdinotrk.

The reference builds the backbone with ``torch.hub.load('facebookresearch/dinov2', name)``
(``models/extractor.py:26``), patches its patch-embedding stride to 7 and its position-embedding
interpolation (``:41-85``), runs the video one frame at a time and keeps the output of block ``layer``
before the final norm, cls token dropped (``:137-150``, ``utils.py:54-67``).  Here the weights come from a
DINOv2 state dict (hub key names, so real checkpoints load unchanged), the position embedding is
interpolated once at load time with the reference's formula, and every frame batch is one
``dinotrk_vit_forward`` call that writes token-major features ``[T][P][C]`` directly.
"""
import ctypes
import math

import torch
import torch.nn.functional as F

from . import _lib

CONFIGS = {  # name: (depth, dim, heads)
    "dinov2_vits14": (12, 384, 6),
    "dinov2_vitb14": (12, 768, 12),
    "dinov2_vitl14": (24, 1024, 16),
    "dinov2_vitg14": (40, 1536, 24),
}
_BLOCK_KEYS = ("norm1.weight", "norm1.bias", "attn.qkv.weight", "attn.qkv.bias", "attn.proj.weight",
               "attn.proj.bias", "ls1.gamma", "norm2.weight", "norm2.bias", "mlp.fc1.weight", "mlp.fc1.bias",
               "mlp.fc2.weight", "mlp.fc2.bias", "ls2.gamma")


def interpolate_pos_embed(pos_embed, n_h, n_w):
    """models/extractor.py:57-85 (DINOv2 passes (w=H_img, h=W_img)): bicubic, +0.1 trick, align_corners=False,
    recompute_scale_factor=False.  Host-side, once per (model, resolution)."""
    N = pos_embed.shape[1] - 1
    dim = pos_embed.shape[-1]
    side = int(math.sqrt(N))
    if n_h * n_w == N and n_h == n_w:
        return pos_embed
    cls_pos = pos_embed[:, 0]
    patch_pos = pos_embed[:, 1:].reshape(1, side, side, dim).permute(0, 3, 1, 2)
    w0, h0 = n_h + 0.1, n_w + 0.1
    patch_pos = F.interpolate(patch_pos, scale_factor=(w0 / math.sqrt(N), h0 / math.sqrt(N)), mode="bicubic",
                              align_corners=False, recompute_scale_factor=False)
    assert int(w0) == patch_pos.shape[-2] and int(h0) == patch_pos.shape[-1]
    patch_pos = patch_pos.permute(0, 2, 3, 1).reshape(1, -1, dim)
    return torch.cat((cls_pos[:, None], patch_pos), dim=1)


class DinoV2Features(torch.nn.Module):
    """``VitExtractor`` replacement: ``forward(video01)`` -> token-major features [T][P][C] on the GPU."""

    def __init__(self, state_dict, heads, layer=None, stride=7, patch=14, device="cuda:0", frames_per_call=2,
                 attention="fused", cta_pairs=True):
        super().__init__()
        self._dev = _lib.require_cuda(device)
        self._lib = _lib.load()
        sd = {k: v.detach().to(self._dev, torch.float32).contiguous() for k, v in state_dict.items()}
        self.dim = sd["cls_token"].shape[-1]
        self.heads = heads
        self.depth = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("blocks."))
        self.layer = self.depth - 1 if layer is None else layer
        self.stride, self.patch = stride, patch
        self.frames_per_call = frames_per_call
        assert attention in ("fused", "materialized")
        self.attention = attention
        self.cta_pairs = cta_pairs
        self._sd = sd
        # fused mode: weight matrices in fp16 (kind::f16 MMAs); materialized (validation) mode: fp32 / TF32
        self._f16 = attention == "fused"
        wdt = torch.float16 if self._f16 else torch.float32
        mats = ("attn.qkv.weight", "attn.proj.weight", "mlp.fc1.weight", "mlp.fc2.weight")
        pw = sd["patch_embed.proj.weight"].reshape(self.dim, -1)
        mult = 8 if self._f16 else 4
        if pw.shape[1] % mult:
            pw = F.pad(pw, (0, mult - pw.shape[1] % mult))
        self._patch_w = pw.to(wdt).contiguous()
        self._blocks = [sd[f"blocks.{i}.{k}"].to(wdt).contiguous() if k in mats else sd[f"blocks.{i}.{k}"]
                        for i in range(self.depth) for k in _BLOCK_KEYS]
        self._block_ptrs = (ctypes.c_void_p * len(self._blocks))(*[t.data_ptr() for t in self._blocks])
        self._pos_cache = {}

    @classmethod
    def from_name(cls, model_name, state_dict, **kw):
        depth, dim, heads = CONFIGS[model_name]
        return cls(state_dict, heads=heads, **kw)

    def _pos(self, n_h, n_w):
        key = (n_h, n_w)
        if key not in self._pos_cache:
            pe = interpolate_pos_embed(self._sd["pos_embed"], n_h, n_w)[0]       # (1 + P) x D
            cls_pos = (self._sd["cls_token"][0, 0] + pe[0]).contiguous()
            self._pos_cache[key] = (cls_pos, pe[1:].contiguous())
        return self._pos_cache[key]

    @torch.no_grad()
    @_lib.on_device
    def forward(self, video01):
        """video01: T x 3 x H x W in [0, 1] (any device).  Returns tpc [T][P][C] (cuda)."""
        lib = self._lib
        T, _, H, W = video01.shape
        geom = _lib.make_geom(H, W, self.patch, self.stride, 35)
        P = geom.h * geom.w
        cfg = _lib.VitConfig(self.depth, self.dim, self.heads, self.layer, self.patch, self.stride,
                             0 if self.attention == "fused" else 1, 1 if self._f16 else 0, 1 if self.cta_pairs else 0)
        cls_pos, pos = self._pos(geom.h, geom.w)
        wt = _lib.VitWeights()
        wt.patch_w, wt.patch_b = self._patch_w.data_ptr(), self._sd["patch_embed.proj.bias"].data_ptr()
        wt.cls_pos, wt.pos = cls_pos.data_ptr(), pos.data_ptr()
        wt.blocks = ctypes.cast(self._block_ptrs, ctypes.POINTER(ctypes.c_void_p))
        out = torch.empty(T, P, self.dim, device=self._dev, dtype=torch.float32)
        B = min(self.frames_per_call, T)
        ws_bytes = lib.dinotrk_vit_workspace_bytes(ctypes.byref(cfg), ctypes.byref(geom), B)
        ws = torch.empty(ws_bytes, device=self._dev, dtype=torch.uint8)
        for i in range(0, T, B):
            e = min(i + B, T)
            fr = video01[i:e].to(self._dev, torch.float32).contiguous()
            _lib.check(lib.dinotrk_vit_forward(_lib.ptr(fr), e - i, ctypes.byref(geom), ctypes.byref(cfg), ctypes.byref(wt),
                                               _lib.ptr(out[i:e]), _lib.ptr(ws), ws_bytes, _lib.stream_ptr()), "vit_forward")
        return out

    def features_chw(self, video01):
        """T x C x h x w view (the layout ``dino_embed_video.pt`` stores, utils.py:66)."""
        tpc = self.forward(video01)
        T, P, C = tpc.shape
        geom = _lib.make_geom(video01.shape[-2], video01.shape[-1], self.patch, self.stride, 35)
        return tpc.view(T, geom.h, geom.w, C).permute(0, 3, 1, 2)


@torch.no_grad()
def get_dino_features_video(video, model_name="dinov2_vitb14", facet="tokens", stride=7, layer=None,
                            device="cuda:0", state_dict=None):
    """``utils.py::get_dino_features_video`` (facet 'tokens'): T x C x h x w on the CPU like the reference
    (``utils.py:53,67``).  ``state_dict``: DINOv2 weights (the reference downloads them with torch.hub)."""
    assert facet == "tokens", "only the 'tokens' facet is on the shipped path (config/preprocessing.yaml:12)"
    assert state_dict is not None, "pass the DINOv2 state dict (no network access here)"
    ex = DinoV2Features.from_name(model_name, state_dict, layer=layer, stride=stride, device=device)
    return ex.features_chw(video).cpu()
