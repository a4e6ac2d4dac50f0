"""Benchmark-video driver: every query frame of a video in ONE inference call, and a
rank-sharded multi-video launcher.

The reference's ``inference_benchmark.py:36-41`` calls ``ModelInference.infer`` once per query frame
(7-21 calls per TAP-Vid video), each call walking all phases with a small batch.  Query points are independent
of each other in every phase (``model_inference.py:97-216``), so here the points of all query frames go through
one ``dinotrk_infer`` work list -- larger correlation-GEMM groups, one host sync per video -- and are split
back per query frame; the files written are the reference's
(``trajectories_{frame_idx}.npy``: N x T x 2 fp32, ``occlusion_preds_{frame_idx}.npy``: N x T bool).
"""
import os
from typing import Callable, Dict, Mapping, Sequence

import numpy as np
import torch

from .model_inference import ModelInference
from .parallel import lpt_assign


@torch.no_grad()
def infer_query_frames(model_inference: ModelInference, query_points: Mapping[int, "np.ndarray"], batch_size=None):
    """query_points: {query frame index: N_f x 3 (x, y, t) px} as returned by the reference's
    ``get_query_points_from_benchmark_config`` (``data/tapvid.py``).  Returns
    {frame index: (trajectories N_f x T x 2 px, occlusion N_f x T bool)} -- what the per-frame loop of
    ``inference_benchmark.py:36-41`` produces, from a single inference call."""
    frames = sorted(query_points.keys())
    if not frames:
        return {}
    dev = model_inference.model._dev
    parts = [torch.as_tensor(np.asarray(query_points[f]), dtype=torch.float32).reshape(-1, 3) for f in frames]
    counts = [int(p.shape[0]) for p in parts]
    allq = torch.cat(parts, dim=0).to(dev)
    traj, occ = model_inference.infer(allq, batch_size)
    out, row = {}, 0
    for f, n in zip(frames, counts):
        out[f] = (traj[row:row + n], occ[row:row + n])
        row += n
    return out


def save_predictions(predictions: Mapping[int, tuple], trajectories_dir: str, occlusions_dir: str):
    """The two np.save lines of ``inference_benchmark.py:40-41``."""
    os.makedirs(trajectories_dir, exist_ok=True)
    os.makedirs(occlusions_dir, exist_ok=True)
    for frame_idx, (traj, occ) in predictions.items():
        np.save(os.path.join(trajectories_dir, f"trajectories_{frame_idx}.npy"), traj[..., :2].cpu().detach().numpy())
        np.save(os.path.join(occlusions_dir, f"occlusion_preds_{frame_idx}.npy"), occ.cpu().detach().numpy())


def run_videos(video_ids: Sequence, costs: Sequence[float], rank: int, world: int,
               run_one: Callable[[object], Dict[int, tuple]]):
    """Multi-video launcher: videos are dealt to ranks by longest-processing-time-first on ``costs``
    (``parallel.video_cost``: N_q * T * (T + 1) * c_map + T * c_frame); every rank runs ``run_one(video_id)`` for its share -- no data-path
    collective.  Returns {video_id: run_one(video_id)} for this rank's videos."""
    mine = lpt_assign(list(costs), world)[rank]
    return {video_ids[i]: run_one(video_ids[i]) for i in mine}
