"""Oracle restatement of the DINO best-buddies search.

Test infrastructure (see ``oracle/__init__.py``).
"""
import torch


def token_coords(H: int, W: int, step: int = 7, patch: int = 14) -> torch.Tensor:
    """preprocessing_dino_bb/dino_bb_utils.py:5-15 -> (h*w) x 2 fp32 pixel (x, y), row-major."""
    s = patch // 2
    x = torch.arange(s, W, step).float()
    y = torch.arange(s, H, step).float()
    yy, xx = torch.meshgrid(y, x, indexing="ij")
    return torch.stack([xx.reshape(-1), yy.reshape(-1)], dim=-1)


def best_buddies_pair(fs: torch.Tensor, ft: torch.Tensor, coords: torch.Tensor):
    """preprocessing_dino_bb/extract_dino_best_buddies.py:31-50 for one ordered (s, t) pair.
    fs, ft: P x C.  Returns source_coords n x 2, target_coords n x 2, cos_sims n and the
    source/target token indices."""
    aff = torch.einsum("nc,mc->nm", fs, ft)
    aff = aff / torch.clamp(fs.norm(dim=1)[:, None] * ft.norm(dim=1)[None], min=1e-08)
    s_max = torch.argmax(aff, dim=1)
    t_max = torch.argmax(aff, dim=0)
    rng = torch.arange(fs.shape[0], device=fs.device)
    s_bb = rng == t_max[s_max]
    t_bb = s_max[s_bb]
    return {"source_coords": coords[s_bb], "target_coords": coords[t_bb],
            "cos_sims": aff[rng[s_bb], t_bb], "source_idx": rng[s_bb], "target_idx": t_bb,
            "row_argmax": s_max, "col_argmax": t_max}


def best_buddies(features: torch.Tensor, H: int, W: int, stride: int = 7) -> dict:
    """extract_dino_best_buddies.py:12-54: all ordered pairs; features T x C x h x w.
    ``H``/``W`` are the script's --h/--w arguments (the pixel extent handed to create_meshgrid)."""
    T, C, h, w = features.shape
    f = features.permute(0, 2, 3, 1).reshape(T, h * w, C)
    coords = token_coords(H, W, stride)
    out = {}
    for s in range(T):
        for t in range(T):
            if s == t:
                continue
            r = best_buddies_pair(f[s], f[t], coords)
            out[f"{s}_{t}"] = {k: r[k] for k in ("source_coords", "target_coords", "cos_sims")}
    return out
