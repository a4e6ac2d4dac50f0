"""GPU: the in-process pipeline (ViT -> delta-DINO -> tracker) against the chained oracles."""
import numpy as np
import pytest
import torch

from oracle import delta_dino as od
from oracle import inference as oi
from oracle import synth
from oracle import vit as ovit
from oracle.tracker import Geometry

pytestmark = pytest.mark.gpu


def test_pixels_to_tracks_pipeline():
    from dino_tracker_b200 import DinoV2Features, ModelInference, build_tracker_from_video, save_dino_embed_video
    H, W, T, D, heads = 98, 126, 4, 64, 1
    geo = Geometry(H=H, W=W)
    g = torch.Generator().manual_seed(8)
    sd = ovit.random_state_dict(2, D, g, n_pos=4, std=0.08)
    video = synth.random_video(T, H, W, seed=9)
    vit = DinoV2Features(sd, heads=heads, layer=1, device="cuda:0")
    model = build_tracker_from_video(video, vit, delta_channels=[3, 8, 8, 8, D])
    dsd = od.random_state_dict([3, 8, 8, 8, D], torch.Generator().manual_seed(10), last_std=0.05)
    model.delta_dino.load_state_dict(dsd)
    head = synth.head_weights("sharp", seed=3)
    model.tracker_head.load_state_dict(head)
    mi = ModelInference(model, model.range_normalizer, 0.7, 0.6)
    # stage parity: ViT (fp16 operands, TF32-level) then delta-DINO (exact fp32 on top of the ViT output)
    ref_dino = ovit.dino_features_video(video, sd, heads, 1)
    got_dino = model.dino_embed_video.cpu()
    scale = ref_dino.abs().max().item()
    assert (got_dino - ref_dino).abs().max().item() <= 5e-3 * scale
    ref_refined = od.refined_features(video, got_dino.contiguous(), dsd)        # delta stage checked on OUR dino features
    assert (model.refined_features.cpu() - ref_refined).abs().max().item() <= 1e-4
    # tracker stage on our refined features: exact parity bar
    q = synth.lattice_query_points(2, 2, H, W, t_q=[0, 1, 2, 3], margin=14.0, jitter_seed=1)
    traj, occ = mi.infer(q.to("cuda:0"))
    t_ref, o_ref = oi.infer(model.refined_features.cpu().contiguous(), q, head, geo, 0.7, 0.6)
    assert (traj.cpu() - t_ref).abs().max().item() <= 1e-3
    assert torch.equal(occ.cpu(), o_ref)


def test_dino_embed_file_round_trip(tmp_path):
    """the ViT stage can still write the reference's dino_embed_video.pt, and Tracker loads it."""
    from dino_tracker_b200 import DinoV2Features, Tracker, save_dino_embed_video
    H, W, T, D = 98, 126, 2, 64
    sd = ovit.random_state_dict(1, D, torch.Generator().manual_seed(1), n_pos=4)
    video = synth.random_video(T, H, W, seed=2)
    vit = DinoV2Features(sd, heads=1, layer=0, device="cuda:0")
    path = str(tmp_path / "dino_embeddings" / "dino_embed_video.pt")
    save_dino_embed_video(video, vit, path)
    saved = torch.load(path)
    assert saved.shape == (T, D, 13, 17) and saved.dtype == torch.float32 and saved.device.type == "cpu"
    m = Tracker(video=video.to("cuda:0"), dino_embed_path=path, device="cuda:0", delta_channels=[3, 4, 4, 4, D])
    assert torch.equal(m.dino_embed_video.cpu(), saved)


def test_benchmark_query_frames_in_one_call(tmp_path):
    """all query frames of a benchmark video through one inference call == the reference's per-frame loop
    (inference_benchmark.py:36-41) within the parity bar (group sizes differ, so thin groups of the loop run on the
    exact-fp32 streaming kernel while the batched call runs them on the split-precision tensor GEMM), same files on disk."""
    from dino_tracker_b200 import ModelInference, Tracker, infer_query_frames, save_predictions, run_videos
    geo = Geometry(H=98, W=126)
    T, C = 6, 64
    feats, _ = synth.shifted_field_features(T, C, geo.h, geo.w, seed=21, noise=0.2, max_shift=2)
    video = synth.random_video(T, geo.H, geo.W, seed=22)
    model = Tracker(video=video, dino_embed_video=feats, device="cuda:0", delta_channels=[3, 4, 4, 4, C])
    model.tracker_head.load_state_dict(synth.head_weights("sharp", seed=5))
    mi = ModelInference(model, model.range_normalizer, 0.7, 0.6)
    qp = {}
    for f, (nx, ny) in {0: (3, 2), 2: (2, 2), 5: (4, 1)}.items():   # TAP-Vid style: {query frame: N_f x 3 (x, y, t)}
        qp[f] = synth.lattice_query_points(nx, ny, geo.H, geo.W, t_q=f, margin=14.0, jitter_seed=f).numpy()
    got = infer_query_frames(mi, qp)
    assert sorted(got) == [0, 2, 5]
    for f in qp:
        t_loop, o_loop = mi.infer(torch.from_numpy(qp[f]).to("cuda:0"))       # what the reference's loop does
        assert (got[f][0] - t_loop).abs().max().item() <= 1e-3 and torch.equal(got[f][1], o_loop)
        t_ref, o_ref = oi.infer(model.refined_features.cpu().contiguous(), torch.from_numpy(qp[f]),
                                synth.head_weights("sharp", seed=5), geo, 0.7, 0.6)
        assert (got[f][0].cpu() - t_ref).abs().max().item() <= 1e-3 and torch.equal(got[f][1].cpu(), o_ref)
    save_predictions(got, str(tmp_path / "trajectories"), str(tmp_path / "occlusions"))
    for f in qp:
        tr = np.load(tmp_path / "trajectories" / f"trajectories_{f}.npy")
        oc = np.load(tmp_path / "occlusions" / f"occlusion_preds_{f}.npy")
        assert tr.shape == (qp[f].shape[0], T, 2) and tr.dtype == np.float32 and oc.shape == (qp[f].shape[0], T) and oc.dtype == bool
    # launcher: two ranks share three videos by cost; together they cover each video once
    seen = []
    for rank in range(2):
        seen += list(run_videos(["a", "b", "c"], [5.0, 3.0, 2.0], rank, 2, lambda v: v.upper()).items())
    assert sorted(seen) == [("a", "A"), ("b", "B"), ("c", "C")]


def test_reference_entry_point_call_sequence_through_the_dropin(tmp_path):
    """The exact call sequence of the reference's entry points (dino_tracker.py::get_model :86-108, inference_grid.py::run
    :12-41, inference_benchmark.py::run :14-41) against the drop-in ``models`` package: constructor kwargs, ``.to(device)``,
    checkpoint files + ``load_weights(iter)``, ``ModelInference(model=..., range_normalizer=..., ...)``, ``model.video.shape``,
    ``infer(query_points=..., batch_size=...)``, ``[..., :2].cpu().detach().numpy()`` -- results against the oracle.
    (The scripts themselves need the reference tree, its dataset folders and config files; tests/test_dropin_surface.py
    checks by AST walk that nothing they touch is missing here.)"""
    import os
    import sys
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(ROOT, "dino_tracker_b200", "dropin"))
    for m in [k for k in list(sys.modules) if k == "models" or k.startswith("models.")]:
        del sys.modules[m]
    try:
        from models.model_inference import ModelInference          # inference_grid.py:6
        from models.tracker import Tracker                         # dino_tracker.py (from models.tracker import Tracker)
        device = "cuda:0"
        geo = Geometry(H=98, W=126)
        T, C = 5, 32
        feats, _ = synth.shifted_field_features(T, C, geo.h, geo.w, seed=61, noise=0.15, max_shift=2)
        video = synth.random_video(T, geo.H, geo.W, seed=62).to(device)
        emb_path = str(tmp_path / "dino_embeddings" / "dino_embed_video.pt")
        os.makedirs(os.path.dirname(emb_path))
        torch.save(feats, emb_path)
        ckpt = tmp_path / "models" / "dino_tracker"
        os.makedirs(ckpt)
        head = synth.head_weights("sharp", seed=61)
        dsd = od.random_state_dict([3, 64, 128, 256, C], torch.Generator().manual_seed(63), last_std=0.01)
        torch.save(head, ckpt / "tracker_head_100.pt")
        torch.save(dsd, ckpt / "delta_dino_100.pt")
        tracker_args = {"video": video, "device": device, "dino_embed_path": emb_path, "dino_patch_size": 14, "stride": 7,
                        "ckpt_path": str(ckpt), "cyc_n_frames": 4, "cyc_batch_size_per_frame": 256,
                        "cyc_fg_points_ratio": 0.7, "cyc_thresh": 4}
        model = Tracker(**tracker_args).to(device)                 # dino_tracker.py:102
        model.load_weights(100)                                    # inference_grid.py:18
        model_inference = ModelInference(model=model, range_normalizer=model.range_normalizer,
                                         anchor_cosine_similarity_threshold=0.7, cosine_similarity_threshold=0.6)
        model_video_h, model_video_w = model.video.shape[-2], model.video.shape[-1]
        assert (model_video_h, model_video_w) == (geo.H, geo.W)
        q = synth.lattice_query_points(3, 2, geo.H, geo.W, t_q=[0, 1, 2, 3, 4, 0], margin=14.0, jitter_seed=61).to(device)
        traj, occ = model_inference.infer(q, batch_size=None)      # inference_grid.py:38
        t_np, o_np = traj[..., :2].cpu().detach().numpy(), occ.cpu().detach().numpy()
        traj_b, occ_b = model_inference.infer(query_points=q, batch_size=3)   # inference_benchmark.py:38 (keyword form)
        refined = od.refined_features(video.cpu(), feats, dsd)
        t_ref, o_ref = oi.infer(refined, q.cpu(), head, geo, 0.7, 0.6)
        assert t_np.shape == (6, T, 2) and o_np.dtype == bool
        assert np.abs(t_np - t_ref.numpy()).max() <= 1e-3 and np.array_equal(o_np, o_ref.numpy())
        t_ref_b, o_ref_b = oi.infer(refined, q.cpu(), head, geo, 0.7, 0.6, batch_size=3)
        assert (traj_b.cpu() - t_ref_b).abs().max().item() <= 1e-3 and torch.equal(occ_b.cpu(), o_ref_b)
    finally:
        sys.path.pop(0)
        for m in [k for k in list(sys.modules) if k == "models" or k.startswith("models.")]:
            del sys.modules[m]
