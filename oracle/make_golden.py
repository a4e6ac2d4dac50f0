"""Generate ``tests/golden/*.npz`` by running the LIVE reference modules (build container only).

    python -m oracle.make_golden            # rewrites every fixture

Test infrastructure (see ``oracle/__init__.py``).  Inputs come from ``oracle/synth.py``
(seeded, regenerable); outputs are whatever the unmodified reference code under
``/root/reference`` returns on CPU behind the three shims of ``oracle/ref_harness.py``.
"""
import os
import sys

import numpy as np
import torch

from . import ref_harness, synth
from . import delta_dino as od

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name: geometry / sizes / seeds of every tracker+inference case
TRACK_CASES = {
    "track_small_well":  dict(H=98, W=126, T=6, C=32, seed=11, head="well", nq=(3, 2), tq=[0, 2, 5, 1, 3, 4], batch=None, noise=0.15),
    "track_small_sharp": dict(H=98, W=126, T=6, C=32, seed=12, head="sharp", nq=(3, 2), tq=[0, 2, 5, 1, 3, 4], batch=None, noise=0.15),
    "track_small_chunk": dict(H=98, W=126, T=7, C=32, seed=13, head="well", nq=(2, 2), tq=[0, 6, 3, 2], batch=3, noise=0.15),
    "track_small_fallback": dict(H=98, W=126, T=5, C=32, seed=14, head="default", nq=(2, 2), tq=[0, 1, 2, 4], batch=None, noise=0.15),
    "track_small_mixed": dict(H=98, W=126, T=5, C=32, seed=15, head="mixed", nq=(2, 2), tq=[0, 1, 2, 4], batch=None, noise=0.15),
    "track_small_noisy": dict(H=98, W=126, T=8, C=48, seed=16, head="sharp", nq=(3, 3), tq=[0, 1, 2, 3, 4, 5, 6, 7, 0], batch=None, noise=0.9),
    "track_full_fallback": dict(H=476, W=854, T=3, C=16, seed=18, head="default", nq=(2, 2), tq=[0, 1, 2, 1], batch=None, noise=0.15),
    "track_full_geom":   dict(H=476, W=854, T=4, C=32, seed=17, head="sharp", nq=(2, 2), tq=[0, 1, 2, 3], batch=None, noise=0.15),
}


def case_inputs(cfg):
    """Inputs of a tracker case, regenerable anywhere from the seeds in ``cfg``."""
    from .tracker import Geometry
    geo = Geometry(H=cfg["H"], W=cfg["W"])
    feats, _ = synth.shifted_field_features(cfg["T"], cfg["C"], geo.h, geo.w, seed=cfg["seed"],
                                            noise=cfg["noise"], max_shift=2)
    head = synth.head_weights(cfg["head"], seed=cfg["seed"])
    q = synth.lattice_query_points(cfg["nq"][0], cfg["nq"][1], cfg["H"], cfg["W"], t_q=cfg["tq"],
                                   margin=12.0, jitter_seed=cfg["seed"])
    return geo, feats, head, q


def gen_track_case(name, cfg):
    geo, feats, head, q = case_inputs(cfg)
    T = cfg["T"]
    video = torch.zeros(T, 3, cfg["H"], cfg["W"])
    model = ref_harness.build_reference_tracker(video, feats, head_sd=head,
                                                delta_channels=[3, 2, 2, 2, cfg["C"]])
    from models.model_inference import ModelInference
    from data.dataset import RangeNormalizer
    rn = RangeNormalizer(shapes=(cfg["W"], cfg["H"], T))
    with torch.no_grad():
        mi = ModelInference(model, rn, 0.7, 0.6)
        # default-initialised delta-DINO has an exactly-zero residual (zero last conv, BN(0)=0)
        assert torch.equal(model.refined_features, feats)
        trajs = mi.compute_trajectories(q, cfg["batch"])
        cos = mi.compute_trajectory_cos_sims(trajs, q)
        anch = mi.compute_anchor_trajectories(trajs, cos, cfg["batch"])
        occ = mi.compute_occlusion(trajs, cos, anch)
        traj2, occ2 = mi.infer(q, cfg["batch"])
        assert torch.equal(traj2, trajs[..., :2]) and torch.equal(occ2, occ)
        # one raw forward (normalised output) for the first query point
        from models.model_inference import generate_trajectory_input
        inp = generate_trajectory_input(q[0], model.video)
        fwd = model(inp)
    N = q.shape[0]
    anchors_pad = np.full((N, T, T, 2), np.nan, dtype=np.float32)
    n_anch = np.zeros(N, dtype=np.int64)
    for n in range(N):
        m = anch[n].shape[0]
        n_anch[n] = m
        anchors_pad[n, :m] = anch[n].numpy()
    out = dict(query_points=q.numpy(), trajectories=trajs.numpy(), cos_sims=cos.numpy(),
               anchors=anchors_pad, n_anchors=n_anch, occlusion=occ.numpy(), forward0=fwd.numpy(),
               feat_checksum=np.array([feats.double().sum().item(), feats.double().abs().sum().item()]))
    if feats.numel() * 4 < 400_000:
        out["features"] = feats.numpy()
    for k, v in head.items():
        out["head." + k] = v.numpy()
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **out)
    print(name, "anchors per query", n_anch.tolist(), "occ frac", occ.float().mean().item())


def gen_delta_case(name, H, W, T, channels, seed):
    g = torch.Generator().manual_seed(seed)
    sd = od.random_state_dict(channels, g, last_std=0.05)
    video = synth.random_video(T, H, W, seed=seed)
    from .tracker import Geometry
    geo = Geometry(H=H, W=W)
    dino = synth.random_features(T, channels[-1], geo.h, geo.w, seed=seed + 1)
    model = ref_harness.build_reference_tracker(video, dino, delta_sd=sd, delta_channels=channels)
    with torch.no_grad():
        model.cache_refined_embeddings()
        refined = model.refined_features
        cnn = video
        for layer in model.delta_dino.layers:
            cnn = layer(cnn)
    out = dict(channels=np.array(channels), seed=np.array(seed), HWT=np.array([H, W, T]))
    ref = refined.numpy()
    if ref.size * 4 < 600_000:
        out["refined"] = ref
        out["cnn_out"] = cnn.numpy()
    else:
        rs = np.random.RandomState(seed)
        idx = rs.randint(0, ref.size, size=4096)
        out["refined_idx"] = idx
        out["refined_vals"] = ref.reshape(-1)[idx]
        out["refined_sum"] = np.array([ref.astype(np.float64).sum(), np.abs(ref.astype(np.float64)).sum()])
    for k, v in sd.items():
        out["sd." + k] = v.numpy()
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **out)
    print(name, "refined", tuple(refined.shape))


def gen_bb_case(name, H, W, T, C, seed):
    import argparse
    import tempfile
    ref_harness.install("cpu")
    from preprocessing_dino_bb import extract_dino_best_buddies as bb
    from .tracker import Geometry
    geo = Geometry(H=H, W=W)
    feats, _ = synth.shifted_field_features(T, C, geo.h, geo.w, seed=seed, noise=0.5, max_shift=2)
    d = tempfile.mkdtemp()
    torch.save(feats, os.path.join(d, "f.pt"))
    args = argparse.Namespace(dino_emb_path=os.path.join(d, "f.pt"), h=H, w=W, stride=7,
                              out_path=os.path.join(d, "out", "bb.pt"))
    bb.run(args)
    res = torch.load(args.out_path)
    out = dict(HWTC=np.array([H, W, T, C]), seed=np.array(seed), features=feats.numpy())
    for k, v in res.items():
        for kk, vv in v.items():
            out[f"{k}.{kk}"] = vv.numpy()
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **out)
    print(name, {k: v["cos_sims"].shape[0] for k, v in res.items()})


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.set_num_threads(8)
    for name, cfg in TRACK_CASES.items():
        gen_track_case(name, cfg)
    gen_delta_case("delta_small", 98, 126, 3, [3, 8, 12, 16, 24], seed=21)
    gen_delta_case("delta_full_geom", 476, 854, 1, [3, 4, 4, 4, 8], seed=22)
    gen_bb_case("bb_small", 98, 126, 3, 16, seed=31)


if __name__ == "__main__":
    main()
