"""Synthetic inputs for bench.py's product arm (kept outside ``oracle/`` so that the product arm never
imports the oracle).  Same constructions as oracle/synth.py."""
import numpy as np
import torch


def lattice(nx, ny, H, W, t_q, margin, jitter_seed):
    xs = np.linspace(margin, W - 1 - margin, nx, dtype=np.float32)
    ys = np.linspace(margin, H - 1 - margin, ny, dtype=np.float32)
    gx, gy = np.meshgrid(xs, ys)
    pts = np.stack([gx.reshape(-1), gy.reshape(-1)], -1)
    pts = pts + np.random.RandomState(jitter_seed).uniform(-3, 3, size=pts.shape).astype(np.float32)
    t = np.full((pts.shape[0], 1), float(t_q), dtype=np.float32)
    return torch.from_numpy(np.concatenate([pts, t], 1).astype(np.float32))


def sharp_head(seed=0):
    """Well-conditioned refiner weights with a dominant centre tap (peaked softmax, like a trained head)."""
    rs = np.random.RandomState(1000 + seed)

    def u(lo, hi, *shape):
        return torch.from_numpy(rs.uniform(lo, hi, size=shape).astype(np.float32))
    w1 = u(0.0, 0.05, 16, 1, 3, 3); w1[:, :, 1, 1] += 1.0
    w2 = u(0.0, 0.05, 1, 16, 3, 3); w2[:, :, 1, 1] += 1.0
    return {"cnn_refiner.0.weight": w1, "cnn_refiner.0.bias": u(-0.05, 0.05, 16),
            "cnn_refiner.2.weight": w2, "cnn_refiner.2.bias": u(-0.05, 0.05, 1)}
