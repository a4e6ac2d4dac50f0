"""Oracle restatement of the inference driver.

Test infrastructure (see ``oracle/__init__.py``).  Follows models/model_inference.py.
"""
import torch
import torch.nn.functional as F

from .tracker import (Geometry, normalize_points_for_sampling, sample_descriptors,
                      tracker_forward, unnormalize_xy)


def trajectory_input(query_point, T, start_t, end_t):
    """models/model_inference.py:8-34."""
    rest = end_t - start_t
    dev = query_point.device
    src_pts = query_point[None].repeat(rest, 1)
    frames_set_t = torch.cat([query_point[2:3].to(torch.float32),
                              torch.arange(start_t, end_t, dtype=torch.float32, device=dev)]).int()
    src_idx = torch.zeros(rest, dtype=torch.long, device=dev)
    tgt_idx = torch.arange(rest, dtype=torch.long, device=dev) + 1
    return src_pts, src_idx, tgt_idx, frames_set_t


def compute_trajectories(features, query_points, head_sd, geo: Geometry, batch_size=None,
                         faithful=False):
    """models/model_inference.py:37-74, :97-107 -> N x T x 3 (x, y, t) in pixels."""
    T = features.shape[0]
    bs = T if batch_size is None else batch_size
    out = []
    for qp in query_points.to(torch.float32):
        chunks = []
        for s in range(0, T, bs):
            e = min(s + bs, T)
            inp = trajectory_input(qp, T, s, e)
            xy = unnormalize_xy(tracker_forward(features, inp, head_sd, geo, faithful), geo)
            ts = inp[-1][1:].to(torch.float32)
            chunks.append(torch.cat([xy, ts[:, None]], dim=1))
        out.append(torch.cat(chunks, dim=0))
    return torch.stack(out)


def compute_trajectory_cos_sims(features, trajectories, query_points, geo: Geometry):
    """models/model_inference.py:110-126 -> N x T.  Note the anchor is the *predicted*
    position at the query frame, and ``F.cosine_similarity`` clamps each norm at 1e-8."""
    N, T = trajectories.shape[:2]
    pn = normalize_points_for_sampling(trajectories, geo)  # broadcast over N x T x 3
    d = sample_descriptors(features, pn.reshape(-1, 3)).reshape(N, T, -1)
    qf = query_points[:, 2].long()
    dq = d[torch.arange(N, device=d.device), qf]
    return F.cosine_similarity(dq[:, None], d, dim=-1)


def anchor_predictions(features, preds, anchor_frames, head_sd, geo: Geometry, batch_size=None,
                       faithful=False):
    """models/model_inference.py:130-154 -> M x T x 2: for every anchor frame a, the track of
    preds[i] (living in frame i) into frame a."""
    T = preds.shape[0]
    dev = preds.device
    bs = T if batch_size is None else batch_size
    out = []
    for a in anchor_frames.tolist():
        coords = []
        for i in range(0, T, bs):
            e = min(i + bs, T)
            frames_set_t = torch.cat([torch.tensor([a], device=dev), torch.arange(i, e, device=dev)]).int()
            src_idx = torch.arange(1, frames_set_t.shape[0], device=dev)
            tgt_idx = torch.zeros(frames_set_t.shape[0] - 1, dtype=torch.long, device=dev)
            inp = (preds[i:e], src_idx, tgt_idx, frames_set_t)
            coords.append(unnormalize_xy(tracker_forward(features, inp, head_sd, geo, faithful), geo))
        out.append(torch.cat(coords)[:, :2])
    if not out:
        return torch.zeros(0, T, 2, device=dev)
    return torch.stack(out)


def compute_anchor_trajectories(features, trajectories, cos_sims, head_sd, geo, anchor_th,
                                batch_size=None, faithful=False):
    """models/model_inference.py:156-165 -> {n: M_n x T x 2}."""
    N, T = trajectories.shape[:2]
    res = {}
    for n in range(N):
        anchors = torch.arange(T, device=cos_sims.device)[cos_sims[n] >= anchor_th]
        res[n] = anchor_predictions(features, trajectories[n], anchors, head_sd, geo, batch_size,
                                    faithful)
    return res


def occlusion_for_query(green, source_xy, cos_sim, anchor_th, cos_th):
    """models/model_inference.py:169-177.  green: M x T x 2, source_xy: T x 2, cos_sim: T.
    ``torch.median`` returns the lower of the two middle elements."""
    vis = cos_sim >= anchor_th
    dists = torch.norm(green - source_xy[vis][:, None], dim=-1)  # M x T
    anchor_median = torch.median(dists[:, vis], dim=0).values
    th = anchor_median.max()
    med = torch.median(dists, dim=0).values
    return (med > th) | (cos_sim < cos_th)


def compute_occlusion(trajectories, cos_sims, anchor_trajs, anchor_th, cos_th):
    """models/model_inference.py:179-200 -> N x T bool."""
    return torch.stack([occlusion_for_query(anchor_trajs[n], trajectories[n, :, :2], cos_sims[n],
                                            anchor_th, cos_th)
                        for n in range(trajectories.shape[0])])


def infer(features, query_points, head_sd, geo: Geometry, anchor_th=0.7, cos_th=0.6,
          batch_size=None, faithful=False, return_all=False):
    """models/model_inference.py:203-216 -> (N x T x 2 px, N x T bool)."""
    trajs = compute_trajectories(features, query_points, head_sd, geo, batch_size, faithful)
    cos = compute_trajectory_cos_sims(features, trajs, query_points, geo)
    anchors = compute_anchor_trajectories(features, trajs, cos, head_sd, geo, anchor_th,
                                          batch_size, faithful)
    occ = compute_occlusion(trajs, cos, anchors, anchor_th, cos_th)
    if return_all:
        return trajs[..., :2], occ, {"trajs": trajs, "cos_sims": cos, "anchors": anchors}
    return trajs[..., :2], occ
