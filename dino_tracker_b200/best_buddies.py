"""Host mirror of ``preprocessing_dino_bb/extract_dino_best_buddies.py`` over libdinotrk.

``run(args)`` keeps the reference script's contract (``--dino-emb-path --h --w --stride --out-path``;
output ``dict['s_t'] -> {source_coords, target_coords, cos_sims}``, rows in ascending source-token
order).  The affinity matrices never reach HBM: every ordered pair runs through the tcgen05 split-fp16
GEMM with a fused top-2 epilogue, candidates are re-evaluated in exact fp32, and the mutual check
works on index vectors (``dinotrk_best_buddies_pairs`` / ``dinotrk_bb_mutual``).

Pairs shard trivially over ranks (``rank`` / ``world`` arguments): SURVEY.md 8e config 5.
"""
import ctypes
import os

import torch

from . import _lib


def token_coords(H, W, step=7, patch=14, device="cpu"):
    """``create_meshgrid`` of preprocessing_dino_bb/dino_bb_utils.py:5-15: pixel (x, y) of every token."""
    s = patch // 2
    x = torch.arange(s, W, step, device=device).float()
    y = torch.arange(s, H, step, device=device).float()
    yy, xx = torch.meshgrid(y, x, indexing="ij")
    return torch.stack([xx.reshape(-1), yy.reshape(-1)], dim=-1)


@torch.no_grad()
def nearest_neighbours(tpc, norms, geom, pairs, hi=None, lo=None, pairs_per_launch=48):
    """pairs: list of ordered (s, t).  Returns nn_idx [n_pairs][P] int32, nn_cos [n_pairs][P] fp32 (device)."""
    with torch.cuda.device(tpc.device):   # the library launches on the current device
        return _nearest_neighbours(tpc, norms, geom, pairs, hi, lo, pairs_per_launch)


def _nearest_neighbours(tpc, norms, geom, pairs, hi, lo, pairs_per_launch):
    lib = _lib.load()
    dev = tpc.device
    T, P, C = tpc.shape
    if hi is None:
        hi = torch.empty(tpc.shape, device=dev, dtype=torch.float16)
        lo = torch.empty(tpc.shape, device=dev, dtype=torch.float16)
        _lib.check(lib.dinotrk_split_fp16(_lib.ptr(tpc), _lib.ptr(hi), _lib.ptr(lo), tpc.numel(), _lib.stream_ptr()))
    feat = _lib.make_features(tpc, norms, hi, lo)
    n = len(pairs)
    nn_idx = torch.empty(n, P, device=dev, dtype=torch.int32)
    nn_cos = torch.empty(n, P, device=dev, dtype=torch.float32)
    ws_bytes = lib.dinotrk_best_buddies_workspace_bytes(min(n, pairs_per_launch), P)
    ws = torch.empty(ws_bytes, device=dev, dtype=torch.uint8)
    pt = torch.tensor(pairs, dtype=torch.int32, device=dev).reshape(-1, 2)
    for i in range(0, n, pairs_per_launch):
        e = min(i + pairs_per_launch, n)
        src = pt[i:e, 0].contiguous()
        tgt = pt[i:e, 1].contiguous()
        _lib.check(lib.dinotrk_best_buddies_pairs(ctypes.byref(feat), ctypes.byref(geom), _lib.ptr(src), _lib.ptr(tgt),
                                                  e - i, _lib.ptr(nn_idx[i:e]), _lib.ptr(nn_cos[i:e]), _lib.ptr(ws),
                                                  ws_bytes, _lib.stream_ptr()), "best_buddies_pairs")
    return nn_idx, nn_cos


@torch.no_grad()
def best_buddies(features_chw, H, W, stride=7, patch=14, device="cuda:0", rank=0, world=1, unordered_pairs=None):
    """features_chw: T x C x h x w.  Returns the reference's dict for the unordered pairs owned by this rank
    (both orientations of each): {'s_t': {...}, 't_s': {...}}."""
    dev = _lib.require_cuda(device)
    with torch.cuda.device(dev):
        return _best_buddies(features_chw, H, W, stride, patch, dev, rank, world, unordered_pairs)


def _best_buddies(features_chw, H, W, stride, patch, dev, rank, world, unordered_pairs):
    lib = _lib.load()
    T, C, h, w = features_chw.shape
    geom = _lib.make_geom((h - 1) * stride + patch, (w - 1) * stride + patch, patch, stride, 35)
    assert (geom.h, geom.w) == (h, w)
    chw = features_chw.to(dev, torch.float32).contiguous()
    tpc = torch.empty(T, h * w, C, device=dev)
    norms = torch.empty(T, h * w, device=dev)
    _lib.check(lib.dinotrk_pack_features(_lib.ptr(chw), _lib.ptr(tpc), _lib.ptr(norms), T, C, h * w, _lib.stream_ptr()))
    del chw
    if unordered_pairs is None:
        unordered_pairs = [(s, t) for s in range(T) for t in range(s + 1, T)]
    mine = unordered_pairs[rank::world]
    ordered = [p for (s, t) in mine for p in ((s, t), (t, s))]
    nn_idx, nn_cos = nearest_neighbours(tpc, norms, geom, ordered)
    n = len(ordered)
    P = h * w
    mutual = torch.empty(n, P, device=dev, dtype=torch.uint8)
    # the partner of ordered pair 2k is 2k+1 and vice versa
    partner = nn_idx.view(-1, 2, P).flip(1).reshape(n, P).contiguous()
    _lib.check(lib.dinotrk_bb_mutual(_lib.ptr(nn_idx), _lib.ptr(partner), n, P, _lib.ptr(mutual), _lib.stream_ptr()))
    coords = token_coords(H, W, stride, patch, device=dev)
    out = {}
    mutual = mutual.bool()
    for k, (s, t) in enumerate(ordered):
        mk = mutual[k]
        out[f"{s}_{t}"] = {"source_coords": coords[mk], "target_coords": coords[nn_idx[k][mk].long()],
                          "cos_sims": nn_cos[k][mk]}
    return out


def run(args):
    """Drop-in for ``extract_dino_best_buddies.run`` (same argparse namespace)."""
    feats = torch.load(args.dino_emb_path, map_location="cpu")
    bb = best_buddies(feats, args.h, args.w, stride=args.stride)
    os.makedirs(os.path.dirname(args.out_path), exist_ok=True)
    torch.save(bb, args.out_path)
    print(f"Saved best buddies to {args.out_path}")
