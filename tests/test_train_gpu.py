"""Training step: the tracker forward WITH a graph and ``dinotrk_track_backward`` against autograd
through the oracle (``oracle/tracker.py`` / ``oracle/delta_dino.py`` on the GPU in exact fp32) on the same inputs.

Bars: coordinates <= 1e-3 px (the inference bar); gradients within 2e-3 of the largest entry of the oracle's gradient
(fp32 sums in a different order, atomics), stated per tensor below.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import oracle
from oracle import delta_dino as od
from oracle import synth
from oracle import tracker as ot
from oracle.tracker import Geometry

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
XY_TOL = 1e-3
GRAD_TOL = 2e-3


def _tracker(geo, feats, head, delta_channels=None, precision="fp16x3"):
    from dino_tracker_b200 import Tracker
    T, C = feats.shape[:2]
    video = synth.random_video(T, geo.H, geo.W, seed=5).to(DEV)
    m = Tracker(video=video, dino_embed_video=feats, device=DEV, delta_channels=delta_channels or [3, 4, 4, 4, C],
                corr_precision=precision)
    m.tracker_head.load_state_dict(head)
    return m


def _batch(geo, N, B, seed):
    g = torch.Generator().manual_seed(seed)
    pts = torch.rand(B, 3, generator=g) * torch.tensor([geo.W - 1.0, geo.H - 1.0, 0.0])
    src = torch.randint(0, N, (B,), generator=g)
    tgt = torch.randint(0, N, (B,), generator=g)
    labels = torch.rand(B, 2, generator=g) * 2 - 1
    return pts, src, tgt, labels


def _rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def _loss(coords, labels):
    return F.huber_loss(coords, labels, reduction="none", delta=1 / 32).mean()      # dino_tracker.py:30,411


@pytest.mark.parametrize("precision", ["fp16x3", "fp32"])
@pytest.mark.parametrize("geo,kind,C,B", [(Geometry(H=98, W=126), "well", 64, 48), (Geometry(H=98, W=126), "mixed", 32, 40),
                                          (Geometry(H=154, W=210), "sharp", 64, 64), (Geometry(H=476, W=854), "well", 64, 24)])
def test_tracker_gradients_match_oracle_autograd(geo, kind, C, B, precision):
    oracle.use_exact_fp32()
    N = 4
    feats, _ = synth.shifted_field_features(N, C, geo.h, geo.w, seed=71, noise=0.2, max_shift=2)
    head = synth.head_weights(kind, seed=71)
    pts, src, tgt, labels = _batch(geo, N, B, 72)
    fs = torch.arange(N, dtype=torch.int32)
    # oracle: autograd through the restatement
    f_o = feats.to(DEV).requires_grad_(True)
    head_o = {k: v.to(DEV).requires_grad_(True) for k, v in head.items()}
    c_o = ot.tracker_forward(f_o, (pts.to(DEV), src.to(DEV), tgt.to(DEV), fs.to(DEV)), head_o, geo)
    _loss(c_o, labels.to(DEV)).backward()
    # CUDA path
    m = _tracker(geo, feats, head, precision=precision)
    emb = feats.to(DEV).clone().requires_grad_(True)
    c = m.get_point_predictions((pts.to(DEV), src.to(DEV), tgt.to(DEV), fs.to(DEV)), emb)
    _loss(c, labels.to(DEV)).backward()
    scale = torch.tensor([geo.W - 1, geo.H - 1], device=DEV) / 2
    assert ((c.detach() - c_o.detach()).abs() * scale).max().item() <= XY_TOL
    assert _rel(emb.grad, f_o.grad) <= GRAD_TOL
    ref = {"cnn_refiner.0.weight": m.tracker_head.cnn_refiner[0].weight, "cnn_refiner.0.bias": m.tracker_head.cnn_refiner[0].bias,
           "cnn_refiner.2.weight": m.tracker_head.cnn_refiner[2].weight, "cnn_refiner.2.bias": m.tracker_head.cnn_refiner[2].bias}
    for k, p in ref.items():
        go = head_o[k].grad
        if k == "cnn_refiner.2.bias":        # softmax is shift-invariant: the exact gradient is 0, both sides carry rounding noise
            assert p.grad.abs().max().item() <= 1e-6 and go.abs().max().item() <= 1e-6
        else:
            assert _rel(p.grad, go) <= GRAD_TOL, k


def test_stability_branch_gradients_match_oracle_autograd():
    """Heads whose kernel sums are ~0 push every map onto the numerical-stability branch (tracker_head.py:87-94): the
    softmax backward then has full-map support; same kernels, full map."""
    oracle.use_exact_fp32()
    geo = Geometry(H=98, W=126)
    N, C, B = 3, 32, 24
    feats, _ = synth.shifted_field_features(N, C, geo.h, geo.w, seed=73, noise=0.2, max_shift=2)
    head = synth.head_weights("default", seed=73)
    pts, src, tgt, labels = _batch(geo, N, B, 74)
    fs = torch.arange(N, dtype=torch.int32)
    f_o = feats.to(DEV).requires_grad_(True)
    head_o = {k: v.to(DEV).requires_grad_(True) for k, v in head.items()}
    inp = (pts.to(DEV), src.to(DEV), tgt.to(DEV), fs.to(DEV))
    corr = ot.corr_maps(ot.sample_descriptors(f_o, torch.cat([ot.normalize_points_for_sampling(inp[0], geo)[:, :2],
                                                                 inp[1][:, None].float()], 1), fs.to(DEV)), f_o, inp[2], frames_set=fs.to(DEV))
    c_o, aux = ot.head_forward(torch.relu(corr), head_o, geo, return_aux=True)
    assert aux["fallback"].any()
    _loss(c_o, labels.to(DEV)).backward()
    m = _tracker(geo, feats, head, precision="fp32")
    emb = feats.to(DEV).clone().requires_grad_(True)
    c = m.get_point_predictions(inp, emb)
    _loss(c, labels.to(DEV)).backward()
    scale = torch.tensor([geo.W - 1, geo.H - 1], device=DEV) / 2
    assert ((c.detach() - c_o.detach()).abs() * scale).max().item() <= XY_TOL
    # logits are ~1e3 here: d/dlogits is a difference of O(1) softmax terms, compared at 2 % of the largest entry
    assert _rel(emb.grad, f_o.grad) <= 2e-2
    assert _rel(m.tracker_head.cnn_refiner[0].weight.grad, head_o["cnn_refiner.0.weight"].grad) <= 2e-2
    assert _rel(m.tracker_head.cnn_refiner[2].weight.grad, head_o["cnn_refiner.2.weight"].grad) <= 2e-2


def test_forward_with_graph_equals_inference_forward_and_trains():
    """Tracker.forward with gradients enabled (cached embeddings: only the head is trainable) returns what the inference
    kernels return, and optimiser steps on the Huber loss (dino_tracker.py:405-421) lower it."""
    geo = Geometry(H=154, W=210)
    N, C, B = 5, 64, 96
    feats, _ = synth.shifted_field_features(N, C, geo.h, geo.w, seed=75, noise=0.2, max_shift=2)
    m = _tracker(geo, feats, synth.head_weights("well", seed=75))
    m.cache_refined_embeddings()
    pts, src, tgt, _ = _batch(geo, 3, B, 76)
    fs = torch.tensor([4, 0, 2], dtype=torch.int32)
    inp = (pts.to(DEV), src.to(DEV), tgt.to(DEV), fs.to(DEV))
    with torch.no_grad():
        c_inf = m(inp)
    c = m(inp)
    assert c.requires_grad and torch.equal(m.frame_embeddings, m.refined_features[fs.long().to(DEV)])
    scale = torch.tensor([geo.W - 1, geo.H - 1], device=DEV) / 2
    assert ((c.detach() - c_inf).abs() * scale).max().item() <= XY_TOL
    # targets = what a differently shaped head predicts on the same maps: Adam steps on the Huber loss move towards them
    m2 = _tracker(geo, feats, synth.head_weights("sharp", seed=75))
    m2.cache_refined_embeddings()
    with torch.no_grad():
        labels = m2(inp)
    opt = torch.optim.Adam(m.tracker_head.parameters(), lr=0.02)
    losses = []
    for _ in range(8):
        opt.zero_grad()
        loss = _loss(m(inp), labels)
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert all(np.isfinite(losses)) and all(b < a for a, b in zip(losses, losses[1:])) and losses[-1] < 0.95 * losses[0], losses


@pytest.mark.parametrize("mode", ["eval", "train"])
def test_training_step_through_delta_dino(mode):
    """The whole graph of dino_tracker.py:405-411: delta-DINO (torch graph; BatchNorm on batch statistics in train mode)
    -> refined embeddings -> tracker node.  eval mode: gradients of the last convolution and of the head against autograd
    through the oracle's functional delta-DINO + tracker; train mode: finite gradients on every parameter, the
    regularisation inputs (frame / raw / residual embeddings) carry the graph."""
    oracle.use_exact_fp32()
    geo = Geometry(H=98, W=126)
    T, C, B = 6, 32, 64
    chans = [3, 8, 8, 8, C]
    feats, _ = synth.shifted_field_features(T, C, geo.h, geo.w, seed=77, noise=0.2, max_shift=2)
    head = synth.head_weights("well", seed=77)
    sd = od.random_state_dict(chans, torch.Generator().manual_seed(78), last_std=0.05)
    m = _tracker(geo, feats, head, delta_channels=chans)
    m.delta_dino.load_state_dict(sd)
    m.train(mode == "train")
    fs = torch.tensor([5, 1, 3, 0], dtype=torch.int32)
    pts, src, tgt, labels = _batch(geo, 4, B, 79)
    inp = (pts.to(DEV), src.to(DEV), tgt.to(DEV), fs.to(DEV))
    c = m(inp)
    assert m.frame_embeddings.requires_grad and m.residual_embeddings.requires_grad and not m.raw_embeddings.requires_grad
    reg = (m.frame_embeddings.norm(dim=1) / m.raw_embeddings.norm(dim=1) - 1).abs().mean()        # dino_tracker.py:136-140
    (_loss(c, labels.to(DEV)) + 1e-4 * reg).backward()
    for name, p in list(m.delta_dino.named_parameters()) + list(m.tracker_head.named_parameters()):
        assert p.grad is not None and torch.isfinite(p.grad).all(), name
    if mode == "train":
        return
    sd_o = {k: (v.to(DEV).requires_grad_(True) if v.dtype.is_floating_point and "running" not in k and "filt" not in k else v.to(DEV))
            for k, v in sd.items()}
    head_o = {k: v.to(DEV).requires_grad_(True) for k, v in head.items()}
    video = m.video
    idx = fs.long().to(DEV)
    refined = od.refined_features(video[idx], feats.to(DEV)[idx], sd_o)
    c_o = ot.tracker_forward(refined, (inp[0], inp[1], inp[2], torch.arange(4, dtype=torch.int32, device=DEV)), head_o, geo)
    reg_o = (refined.norm(dim=1) / feats.to(DEV)[idx].norm(dim=1) - 1).abs().mean()
    (_loss(c_o, labels.to(DEV)) + 1e-4 * reg_o).backward()
    scale = torch.tensor([geo.W - 1, geo.H - 1], device=DEV) / 2
    assert ((c.detach() - c_o.detach()).abs() * scale).max().item() <= XY_TOL
    for key in ("layers.12.weight", "layers.12.bias", "layers.13.weight", "layers.8.weight", "layers.0.weight"):
        got = dict(m.delta_dino.named_parameters())[key].grad
        assert _rel(got, sd_o[key].grad) <= 5e-3, key
    assert _rel(m.tracker_head.cnn_refiner[0].weight.grad, head_o["cnn_refiner.0.weight"].grad) <= GRAD_TOL


def test_sample_embeddings_gradient_matches_oracle_autograd():
    """Tracker.sample_embeddings on embeddings with a graph (contrastive losses, dino_tracker.py:215-220)."""
    geo = Geometry(H=98, W=126)
    T, C, B = 4, 48, 200
    feats = synth.random_features(T, C, geo.h, geo.w, seed=81)
    m = _tracker(geo, feats, synth.head_weights("well"))
    g = torch.Generator().manual_seed(82)
    pts = torch.rand(B, 3, generator=g) * torch.tensor([2.4, 2.4, 0.0]) - torch.tensor([1.2, 1.2, 0.0])    # some outside [-1, 1]
    pts[:, 2] = torch.randint(0, T, (B,), generator=g).float()
    wgt = torch.randn(B, C, generator=g)
    f_o = feats.to(DEV).requires_grad_(True)
    (ot.sample_descriptors(f_o, pts.to(DEV)) * wgt.to(DEV)).sum().backward()
    emb = feats.to(DEV).clone().requires_grad_(True)
    d = m.sample_embeddings(emb, pts.to(DEV))
    assert d.requires_grad
    (d * wgt.to(DEV)).sum().backward()
    assert _rel(emb.grad, f_o.grad) <= 1e-5


def test_cycle_consistent_preds_are_consistent_and_trainable():
    """models/tracker.py:182-301 on the CUDA path: the surviving points return within cyc_thresh px, the predictions of
    the returned inputs are reproduced by get_point_predictions, and the cycle loss of dino_tracker.py:346-353 reaches
    delta-DINO and the head."""
    geo = Geometry(H=98, W=126)
    T, C = 6, 32
    chans = [3, 8, 8, 8, C]
    feats, _ = synth.shifted_field_features(T, C, geo.h, geo.w, seed=83, noise=0.1, max_shift=1)
    m = _tracker(geo, feats, synth.head_weights("sharp", seed=83), delta_channels=chans)
    m.delta_dino.load_state_dict(od.random_state_dict(chans, torch.Generator().manual_seed(84), last_std=0.02))
    m.cyc_n_frames, m.cyc_batch_size_per_frame, m.cyc_thresh = 3, 64, 4
    fg = torch.zeros(T, geo.H, geo.W, device=DEV)
    fg[:, 20:70, 30:100] = 1
    fs = torch.tensor([0, 2, 3, 5], dtype=torch.int64, device=DEV)
    torch.manual_seed(85)
    pts, src, tgt, labels = _batch(geo, 4, 32, 86)
    c = m((pts.to(DEV), src.to(DEV), tgt.to(DEV), fs))           # the step's forward: its embeddings feed the cycle search
    preds = m.get_cycle_consistent_preds(fs, fg)
    n = preds["source_coords"].shape[0]
    assert n > 0 and preds["cycle_consistency_dists"].max().item() <= m.cyc_thresh
    for k in ("source_coords", "target_coords", "source_target_coords", "target_source_coords", "cycle_points"):
        assert preds[k].shape[0] == n
    assert preds["source_target_coords"].requires_grad and preds["target_source_coords"].requires_grad
    # forward predictions land where the (px) target points say, up to the bar
    tgt_px = m.range_normalizer.unnormalize(preds["target_coords"], src=(-1, 1))[:, :2]
    got_px = m.range_normalizer.unnormalize(preds["source_target_coords"].detach(), src=(-1, 1), dims=[0, 1])
    assert (tgt_px - got_px).abs().max().item() <= XY_TOL
    w = 0.8 ** preds["cycle_consistency_dists"]
    huber = torch.nn.HuberLoss(delta=1 / 32, reduction="none")
    loss = ((w[:, None] * huber(preds["source_target_coords"], preds["target_coords"][:, :2])).mean() +
            (w[:, None] * huber(preds["target_source_coords"], preds["source_coords"][:, :2])).mean()) / 2 + _loss(c, labels.to(DEV))
    loss.backward()
    assert m.delta_dino.layers[12].weight.grad.abs().max().item() > 0
    assert m.tracker_head.cnn_refiner[0].weight.grad.abs().max().item() > 0


def test_training_step_matches_reference_vectors():
    """The step the LIVE reference ran on the CPU in train mode (tests/golden/train_small.npz: model(inputs) -> Huber +
    norm regulariser -> backward, dino_tracker.py:405-427) through the drop-in Tracker on the GPU: coordinates, loss and the
    gradient of every trainable tensor and of the refined embeddings."""
    import os
    from golden_util import GOLDEN_DIR
    from oracle import make_golden as mg
    from dino_tracker_b200 import Tracker
    oracle.use_exact_fp32()                       # cuDNN convolutions of the delta-DINO graph in fp32, as the CPU reference
    g = np.load(os.path.join(GOLDEN_DIR, "train_small.npz"))
    cfg = mg.TRAIN_CASE
    geo, feats, video, head, dsd, inp, labels = mg.train_case_inputs()
    m = Tracker(video=video.to(DEV), dino_embed_video=feats, device=DEV, delta_channels=cfg["channels"])
    m.tracker_head.load_state_dict(head)
    m.delta_dino.load_state_dict(dsd)
    m.train()
    coords = m(tuple(t.to(DEV) for t in inp))
    fe = m.frame_embeddings
    fe.retain_grad()
    loss = mg.train_loss(coords, labels.to(DEV), fe, m.raw_embeddings)
    loss.backward()
    scale = np.array([geo.W - 1, geo.H - 1]) / 2
    assert (np.abs(coords.detach().cpu().numpy() - g["coords"]) * scale).max() <= XY_TOL
    assert abs(loss.item() - float(g["loss"])) <= 1e-6

    def close(got, want, rel, floor):
        return float(np.abs(got.detach().cpu().numpy() - want).max()) <= max(rel * float(np.abs(want).max()), floor)
    assert close(fe.grad, g["grad_frame_embeddings"], GRAD_TOL, 1e-9)
    for k, p in m.delta_dino.named_parameters():
        assert close(p.grad, g["grad.delta_dino." + k], 5e-3, 2e-7), k      # (conv biases before a train-mode BN: exactly 0)
    for k, p in m.tracker_head.named_parameters():
        assert close(p.grad, g["grad.tracker_head." + k], GRAD_TOL, 1e-9), k


def test_cycle_consistent_preds_match_reference_vectors():
    """models/tracker.py:182-301 of the LIVE reference on the CPU (tests/golden/cycle_small.npz) vs the drop-in: frame set and
    masks are given as host tensors, so the random draws (randint / randperm on the host generator) are the reference's --
    same sampled pixels, same survivors; tracked coordinates within the parity bar."""
    import os
    from golden_util import GOLDEN_DIR
    from oracle import make_golden as mg
    from dino_tracker_b200 import Tracker
    g = np.load(os.path.join(GOLDEN_DIR, "cycle_small.npz"))
    cfg = mg.CYC_CASE
    geo, feats, head, fg, inp = mg.cyc_case_inputs()
    # (default-initialised delta-DINO: zero residual whatever its widths; the CUDA path wants multiples of 4)
    m = Tracker(video=torch.zeros(cfg["T"], 3, cfg["H"], cfg["W"], device=DEV), dino_embed_video=feats, device=DEV,
                delta_channels=[3, 4, 4, 4, cfg["C"]])
    m.tracker_head.load_state_dict(head)
    m.cyc_n_frames, m.cyc_batch_size_per_frame = cfg["n_frames"], cfg["per_frame"]
    m.cyc_fg_points_ratio, m.cyc_thresh = cfg["fg_ratio"], cfg["thresh"]
    with torch.no_grad():
        m.cache_refined_embeddings()
        m((inp[0].to(DEV), inp[1].to(DEV), inp[2].to(DEV), inp[3]))
        torch.manual_seed(cfg["rng"])
        preds = m.get_cycle_consistent_preds(inp[3], fg)          # host frame set + host masks: the reference's draws
    got = {k: v.detach().cpu().numpy() for k, v in preds.items()}
    assert got["source_coords"].shape == g["source_coords"].shape            # same survivors
    assert np.abs(got["source_coords"] - g["source_coords"]).max() <= 1e-6   # same sampled pixels, same normalisation
    to_px = np.array([geo.W - 1, geo.H - 1]) / 2
    assert (np.abs(got["target_coords"][:, :2] - g["target_coords"][:, :2]) * to_px).max() <= XY_TOL
    assert np.abs(got["target_coords"][:, 2] - g["target_coords"][:, 2]).max() <= 1e-6
    assert (np.abs(got["source_target_coords"] - g["source_target_coords"]) * to_px).max() <= XY_TOL
    # the way back starts from the (1e-3 px different) forward prediction: same bar plus that offset
    assert (np.abs(got["target_source_coords"] - g["target_source_coords"]) * to_px).max() <= 2 * XY_TOL
    assert np.abs(got["cycle_points"] - g["cycle_points"]).max() <= 2 * XY_TOL
    assert np.abs(got["cycle_consistency_dists"] - g["cycle_consistency_dists"]).max() <= 3 * XY_TOL
