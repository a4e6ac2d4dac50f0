"""Seeded synthetic inputs shared by the golden generator, the tests and bench.py.

Test/bench infrastructure (see ``oracle/__init__.py``).  Everything is drawn from
``numpy.random.RandomState`` (bit-stable legacy generator) so that fixtures can be
regenerated from a seed on any box.
"""
import numpy as np
import torch


def shifted_field_features(T, C, h, w, seed=0, noise=0.15, max_shift=3, smooth=True):
    """A smooth random descriptor field that translates by a few tokens per frame plus noise:
    tracks are non-trivial, most frames pass the 0.7 anchor cos-sim threshold.
    Returns (features T x C x h x w fp32, shifts T x 2 int (dy, dx))."""
    rs = np.random.RandomState(seed)
    pad = max_shift * 2 + 2
    base = rs.standard_normal((C, h + 2 * pad, w + 2 * pad)).astype(np.float32)
    if smooth:
        b = base.copy()
        b[:, 1:-1, 1:-1] = (base[:, 1:-1, 1:-1] * 0.5 + 0.125 * (base[:, :-2, 1:-1] + base[:, 2:, 1:-1]
                            + base[:, 1:-1, :-2] + base[:, 1:-1, 2:]))
        base = b
    shifts = np.zeros((T, 2), dtype=np.int64)
    for t in range(1, T):
        shifts[t] = np.clip(shifts[t - 1] + rs.randint(-1, 2, size=2), -max_shift, max_shift)
    feats = np.empty((T, C, h, w), dtype=np.float32)
    for t in range(T):
        dy, dx = shifts[t]
        feats[t] = base[:, pad + dy: pad + dy + h, pad + dx: pad + dx + w]
        feats[t] += noise * rs.standard_normal((C, h, w)).astype(np.float32)
    return torch.from_numpy(feats), torch.from_numpy(shifts)


def random_features(T, C, h, w, seed=0):
    rs = np.random.RandomState(seed)
    return torch.from_numpy(rs.standard_normal((T, C, h, w)).astype(np.float32))


def random_video(T, H, W, seed=0):
    rs = np.random.RandomState(seed)
    return torch.from_numpy(rs.random_sample((T, 3, H, W)).astype(np.float32))


def head_weights(kind="well", seed=0):
    """Head state dict (keys of models/networks/tracker_head.py:54-58).
    'well': U(0.2, 1) for all four tensors (kernel sums far from 0);
    'default': PyTorch-default-like init -> kernel sums ~0 -> huge logits -> every map takes the
               numerical-stability fallback branch; 'mixed': mixed-sign but non-degenerate sums."""
    rs = np.random.RandomState(1000 + seed)
    def u(lo, hi, *shape):
        return torch.from_numpy(rs.uniform(lo, hi, size=shape).astype(np.float32))
    if kind == "well":
        sd = {"cnn_refiner.0.weight": u(0.2, 1, 16, 1, 3, 3), "cnn_refiner.0.bias": u(0.2, 1, 16),
              "cnn_refiner.2.weight": u(0.2, 1, 1, 16, 3, 3), "cnn_refiner.2.bias": u(0.2, 1, 1)}
    elif kind == "default":
        # mixed-sign kernels whose spatial sums are tiny (like PyTorch-default init in the probe of
        # SURVEY.md 8d): normalisation blows the gain up, the softmax collapses onto a spot unrelated
        # to the pre-CNN arg-max, and the disc mass drops below 1e-8 -> fallback branch.
        def tiny_sum(o, i, total):
            r = u(-1, 1, o, i, 3, 3)
            return r - r.mean(dim=(2, 3), keepdim=True) + total / 9
        sd = {"cnn_refiner.0.weight": tiny_sum(16, 1, 0.05), "cnn_refiner.0.bias": u(-1 / 3, 1 / 3, 16),
              "cnn_refiner.2.weight": tiny_sum(1, 16, 0.05), "cnn_refiner.2.bias": u(-1 / 12, 1 / 12, 1)}
    elif kind == "mixed":
        w1 = u(-0.5, 1, 16, 1, 3, 3); w2 = u(-0.5, 1, 1, 16, 3, 3)
        sd = {"cnn_refiner.0.weight": w1, "cnn_refiner.0.bias": u(-0.2, 0.2, 16),
              "cnn_refiner.2.weight": w2, "cnn_refiner.2.bias": u(-0.2, 0.2, 1)}
    elif kind == "sharp":
        # 'well' with a large gain folded in: kernel sums stay O(1) after normalisation, but the
        # centre tap dominates -> peaked softmax (closer to a trained head)
        w1 = u(0.0, 0.05, 16, 1, 3, 3); w1[:, :, 1, 1] += 1.0
        w2 = u(0.0, 0.05, 1, 16, 3, 3); w2[:, :, 1, 1] += 1.0
        sd = {"cnn_refiner.0.weight": w1, "cnn_refiner.0.bias": u(-0.05, 0.05, 16),
              "cnn_refiner.2.weight": w2, "cnn_refiner.2.bias": u(-0.05, 0.05, 1)}
    else:
        raise ValueError(kind)
    return sd


def lattice_query_points(n_side_x, n_side_y, H, W, t_q=0, margin=20.0, jitter_seed=None):
    xs = np.linspace(margin, W - 1 - margin, n_side_x, dtype=np.float32)
    ys = np.linspace(margin, H - 1 - margin, n_side_y, dtype=np.float32)
    gx, gy = np.meshgrid(xs, ys)
    pts = np.stack([gx.reshape(-1), gy.reshape(-1)], -1)
    if jitter_seed is not None:
        pts = pts + np.random.RandomState(jitter_seed).uniform(-3, 3, size=pts.shape).astype(np.float32)
    t = np.full((pts.shape[0], 1), float(t_q), dtype=np.float32) if np.isscalar(t_q) else \
        np.asarray(t_q, dtype=np.float32).reshape(-1, 1)
    return torch.from_numpy(np.concatenate([pts, t], 1).astype(np.float32))
