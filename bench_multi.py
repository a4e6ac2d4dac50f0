"""Multi-GPU blocks of bench.py (BASELINE.json configs 3, 4, 5).  One process per GPU; every block reports
device/wall times as the MAX over ranks, a digest of its results (identical for every world size: sharding never changes a
bit) and the limiter.  The reference has no multi-GPU path at all (inference_grid.py:9 hard-codes one device).

  config3  the 30 TAP-Vid-DAVIS video shapes (dino_tracker_b200/data/davis_shapes.json), videos dealt to ranks by
           longest-processing-time-first on N_q*T*(T+1)*c_map + T*c_frame; per video: ViT-L/14@15 + delta-DINO on its T
           frames, then ONE inference call over all its query frames.  No data-path collective.
  config4  one long video (T=250, 1024 query points): frames block-sharded (ViT + delta-DINO per rank), refined features
           stored into every GPU's buffer from inside the delta-DINO epilogue (CUDA-IPC peer stores over NVLink = the
           all-gather), query points sharded, results gathered.
  config5  best buddies over T=100 frames: features replicated, unordered frame pairs dealt round-robin.
"""
import hashlib
import json
import os
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))


def _digest(*tensors):
    h = hashlib.sha1()
    for t in tensors:
        h.update(t.detach().contiguous().cpu().numpy().tobytes())
    return h.hexdigest()[:16]


def _max_over_ranks(dist, dev, *vals):
    t = torch.tensor(vals, device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def _vit_large(dev, seed=7):
    """ViT-L/14 truncated at the tap block (15), random-init weights of the named architecture."""
    from dino_tracker_b200.vit import CONFIGS, DinoV2Features
    depth, dim, heads = CONFIGS["dinov2_vitl14"]
    layer = 15
    g = torch.Generator(device=dev).manual_seed(seed)

    def rn(*shape, std=0.02):
        return torch.randn(*shape, device=dev, generator=g) * std
    sd = {"cls_token": rn(1, 1, dim), "pos_embed": rn(1, 1 + 37 * 37, dim), "patch_embed.proj.weight": rn(dim, 3, 14, 14),
          "patch_embed.proj.bias": rn(dim)}
    for i in range(layer + 1):
        p = f"blocks.{i}."
        sd.update({p + "norm1.weight": 1 + rn(dim), p + "norm1.bias": rn(dim), p + "attn.qkv.weight": rn(3 * dim, dim),
                   p + "attn.qkv.bias": rn(3 * dim), p + "attn.proj.weight": rn(dim, dim), p + "attn.proj.bias": rn(dim),
                   p + "ls1.gamma": 1 + rn(dim), p + "norm2.weight": 1 + rn(dim), p + "norm2.bias": rn(dim),
                   p + "mlp.fc1.weight": rn(4 * dim, dim), p + "mlp.fc1.bias": rn(4 * dim),
                   p + "mlp.fc2.weight": rn(dim, 4 * dim), p + "mlp.fc2.bias": rn(dim), p + "ls2.gamma": 1 + rn(dim)})
    return DinoV2Features(sd, heads=heads, layer=layer, device=dev, frames_per_call=2)


def _delta_weights(model, seed=5):
    """Deterministic non-trivial delta-DINO weights (identical on every rank)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in sorted(model.delta_dino.named_parameters()):
            if p.dim() > 1:      # convolution kernels
                fan = p[0].numel()
                p.copy_(((torch.rand(p.shape, generator=g) * 2 - 1) / fan ** 0.5 * (0.05 if "layers.12" in name else 1.0)).to(p.device))
            elif name.split(".")[1] in ("0", "4", "8", "12"):   # convolution biases (default init draws from the global RNG)
                p.copy_(((torch.rand(p.shape, generator=g) * 2 - 1) * 0.05).to(p.device))


def _frames(T, seed, dev):
    g = torch.Generator(device=dev).manual_seed(seed)
    return torch.rand(T, 3, 476, 854, device=dev, generator=g)


# ------------------------------------------------------------------------------------------------ config 3
def davis_shapes():
    return json.load(open(os.path.join(ROOT, "dino_tracker_b200", "data", "davis_shapes.json")))["videos"]


def video_cost(T, nq, c_map, c_frame):
    from dino_tracker_b200.parallel import video_cost as vc
    return vc(T, nq, c_map, c_frame)


def config3(dist, rank, world, dev, bench, c_map, c_frame, with_vit=True):
    from bench_inputs import lattice
    from dino_tracker_b200 import ModelInference, Tracker, infer_query_frames
    from dino_tracker_b200.parallel import lpt_assign
    vids = davis_shapes()
    costs = [video_cost(v["T"], v["n_query_points"], c_map, c_frame if with_vit else 0.0) for v in vids]
    mine = lpt_assign(costs, world)[rank]
    vit = _vit_large(dev) if with_vit else None
    digests, t_feat, t_track = {}, 0.0, 0.0
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t_start = time.perf_counter()
    for i in mine:
        v = vids[i]
        T, nq, nf = v["T"], v["n_query_points"], v["n_query_frames"]
        feats = bench.synth_video_features(T, 1024, dev, 3000 + v["video_idx"], 0.25)   # what the tracker sees
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        frames = _frames(T, 4000 + v["video_idx"], dev)
        if vit is not None:
            vit_out = vit(frames)                                                       # ViT-L/14@15 on the video's frames
            del vit_out
        model = Tracker(video=frames, dino_embed_video=feats, device=dev, delta_channels=[3, 64, 128, 256, 1024])
        _delta_weights(model)
        model.tracker_head.load_state_dict(bench.head_weights_for("sharp"))
        mi = ModelInference(model, model.range_normalizer, 0.7, 0.6)                    # delta-DINO over all frames
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        side = int(nq ** 0.5) + 1
        pts = lattice(side, side, bench.H, bench.W, 0, 30.0, v["video_idx"])[:nq].numpy()
        qp = {}
        for j in range(nq):                                                             # strided query frames 0, 5, 10, ...
            f = min((j % nf) * 5, T - 1)
            qp.setdefault(f, []).append([pts[j, 0], pts[j, 1], f])
        res = infer_query_frames(mi, qp)                                                # one call for all query frames
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        t_feat += t1 - t0
        t_track += t2 - t1
        digests[v["video_idx"]] = _digest(*[x for f in sorted(res) for x in res[f]])
        del model, mi, feats, frames, res
    torch.cuda.synchronize()
    mine_s = time.perf_counter() - t_start
    all_d = [digests]
    loads = [(mine_s, t_feat, t_track, sum(costs[i] for i in mine))]
    if dist is not None:
        all_d = [None] * world
        dist.all_gather_object(all_d, digests)
        loads = [None] * world
        dist.all_gather_object(loads, (mine_s, t_feat, t_track, sum(costs[i] for i in mine)))
    if rank != 0:
        return None
    merged = {}
    for d in all_d:
        merged.update(d)
    makespan = max(l[0] for l in loads)
    total = sum(l[0] for l in loads)
    return {"videos": len(vids), "frames": sum(v["T"] for v in vids), "query_points": sum(v["n_query_points"] for v in vids),
            "corr_maps_upper_bound": sum(v["n_query_points"] * v["T"] * (v["T"] + 1) for v in vids),
            "makespan_s": makespan, "sum_over_ranks_s": total, "ideal_s": total / world, "balance": total / world / makespan,
            "videos_per_s": len(vids) / makespan, "query_points_per_s": sum(v["n_query_points"] for v in vids) / makespan,
            "per_rank": [{"wall_s": l[0], "feature_stage_s": l[1], "tracker_s": l[2], "predicted_s": l[3]} for l in loads],
            "cost_model": {"c_map_s": c_map, "c_frame_s": c_frame if with_vit else 0.0,
                           "formula": "N_q*T*(T+1)*c_map + T*c_frame (LPT assignment)"},
            "feature_stage": "ViT-L/14@15 + delta-DINO [3,64,128,256,1024] per video" if with_vit else "delta-DINO only",
            "limiter": "the largest rank load under LPT (videos are indivisible; no collective)",
            "results_digest": hashlib.sha1(json.dumps(sorted(merged.items())).encode()).hexdigest()[:16]}


# ------------------------------------------------------------------------------------------------ config 4
def config4(dist, rank, world, dev, bench, T=250, nq=1024, with_vit=True):
    from dino_tracker_b200 import ModelInference, Tracker, parallel as par
    C, P = 1024, bench.P
    side = int(round(nq ** 0.5))
    q = bench.query_lattice(side * side, 0).to(dev)
    q[:, 2] = (torch.arange(q.shape[0], device=dev) * 7) % T            # query frames spread over the video
    s, e = par.frame_shard(T, world, rank)
    vit = _vit_large(dev) if with_vit else None
    # what the tracker sees: a translating field (every frame qualifies as an anchor: the worst case, 64 M maps)
    dino_all = bench.synth_video_features(T, C, dev, 777, 0.25)         # T x C x h x w (deterministic, same on every rank)
    dino_mine = dino_all[s:e].permute(0, 2, 3, 1).reshape(e - s, P, C).contiguous()
    del dino_all
    frames = _frames(T, 778, dev)[s:e].contiguous()                     # this rank's frames
    pbuf = par.PeerFeatureBuffer(T, P, C, rank, world)
    full = pbuf.tensor
    m = Tracker(video=frames, dino_embed_video=torch.zeros(1), device=dev, delta_channels=[3, 64, 128, 256, C],
                _adopt_tpc=dino_mine)
    _delta_weights(m)
    m.tracker_head.load_state_dict(bench.head_weights_for("sharp"))
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    vit_sum = 0.0
    if vit is not None:
        vo = vit(frames)                                                # ViT-L/14@15 on this rank's frames
        vit_sum = float(vo.double().sum().item())
        del vo
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    # delta-DINO on this rank's frames; every refined row goes to this rank's slice AND to every peer's buffer from inside
    # the producing kernel (the all-gather)
    _, norms_mine = m.delta_dino.refine_tpc(frames, dino_mine, m._geom, out=full[s:e], peer_ptrs=pbuf.peer_ptrs, first_frame=s)
    pbuf.sync()
    t2 = time.perf_counter()
    # tracker over all frames for this rank's query rows
    m._refined_tpc = full
    # per-token norms of the gathered video through the library's own kernel on every world size (bit-identical inputs to
    # the tracker whatever the sharding)
    from dino_tracker_b200 import _lib as _l
    norms_all = torch.empty(T, P, device=dev, dtype=torch.float32)
    _l.check(_l.load().dinotrk_token_norms(_l.ptr(full), _l.ptr(norms_all), T, C, P, _l.stream_ptr()), "token_norms")
    m._refined_norms = norms_all
    m.video = torch.zeros(T, 3, 2, 2, device=dev)                      # (only its frame count is read from here on)
    mi = ModelInference.__new__(ModelInference)
    torch.nn.Module.__init__(mi)
    mi.model, mi.range_normalizer = m, m.range_normalizer
    mi.anchor_cosine_similarity_threshold, mi.cosine_similarity_threshold = 0.7, 0.6
    qs, qe = par.query_shard(q.shape[0], world, rank)
    traj, occ = mi.infer(q[qs:qe])
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    if dist is not None:
        traj = par.gather_rows(traj.contiguous(), q.shape[0], world, rank)
        occ = par.gather_rows(occ.to(torch.uint8), q.shape[0], world, rank).bool()
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    from dino_tracker_b200 import _lib
    stats = _lib.infer_stats()
    # the same exchange through NCCL, for the achieved-bandwidth figure (in place on the already complete buffer)
    ag_s = None
    if dist is not None:
        dist.barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        par.allgather_frames(full, T, world, rank)
        a1.record()
        torch.cuda.synchronize()
        ag_s = a0.elapsed_time(a1) / 1e3
    feat_digest = _digest(full[:: max(T // 10, 1), ::997])
    vit_s, delta_s, track_s, gather_s, total_s, ag_max = _max_over_ranks(dist, dev, t1 - t0, t2 - t1, t3 - t2, t4 - t3, t4 - t0,
                                                                         ag_s or 0.0)
    vit_sums = [vit_sum]
    if dist is not None:
        vit_sums = [None] * world
        dist.all_gather_object(vit_sums, vit_sum)
    out = None
    if rank == 0:
        recv_gb = (T - (e - s)) * P * C * 4 / 1e9
        out = {"T": T, "query_points": int(q.shape[0]), "C": C, "frames_per_rank": e - s, "query_points_per_rank": qe - qs,
               "vit_s": vit_s, "delta_dino_plus_fused_allgather_s": delta_s, "tracker_s": track_s, "result_gather_s": gather_s,
               "total_s": total_s, "query_points_per_s": q.shape[0] / total_s,
               "anchor_maps_rank0": stats["anchor_maps"], "anchor_pipeline": stats["pipeline"],
               "allgather": {"how": "peer stores from the delta-DINO epilogue (CUDA IPC over NVLink), then stream sync + barrier",
                             "received_GB_per_gpu": recv_gb,
                             "nccl_all_gather_into_tensor_s": ag_max if dist is not None else None,
                             "nccl_GBps_received_per_gpu": (recv_gb / ag_max) if (dist is not None and ag_max > 0) else None,
                             "nvlink5_peak_GBps_per_direction": 900.0},
               "results_digest": _digest(traj, occ), "features_digest": feat_digest,
               "vit_output_checksums": [round(x, 3) for x in vit_sums],
               "limiter": "the tracker on the rank's query shard (64.3 M correlation maps / world) and the ViT on its frame shard; "
                          "the feature exchange is a few ms"}
    pbuf.sync()
    del m, mi, full
    pbuf.close()
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------ config 5
def config5(dist, rank, world, dev, bench, T=100, C=1024):
    from dino_tracker_b200 import _lib
    from dino_tracker_b200.best_buddies import nearest_neighbours
    feats = bench.synth_video_features(T, C, dev, 99, 0.5)               # replicated: same tensor on every rank
    tpc = feats.permute(0, 2, 3, 1).reshape(T, bench.P, C).contiguous()
    norms = tpc.norm(dim=2).contiguous()
    del feats
    geom = _lib.make_geom(bench.H, bench.W)
    unordered = [(s, t) for s in range(T) for t in range(s + 1, T)]
    mine = unordered[rank::world]
    ordered = [p for (s, t) in mine for p in ((s, t), (t, s))]
    nearest_neighbours(tpc, norms, geom, ordered[:4])                    # warm-up
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    nn_idx, nn_cos = nearest_neighbours(tpc, norms, geom, ordered)
    e1.record()
    torch.cuda.synchronize()
    (s_max,) = _max_over_ranks(dist, dev, e0.elapsed_time(e1) / 1e3)
    # order-independent digest of all (pair, nearest-neighbour index) results
    mine_sum = {f"{a}_{b}": int(nn_idx[k].long().sum().item()) for k, (a, b) in enumerate(ordered)}
    sums = [mine_sum]
    if dist is not None:
        sums = [None] * world
        dist.all_gather_object(sums, mine_sum)
    del tpc, norms, nn_idx, nn_cos
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    merged = {}
    for d in sums:
        merged.update(d)
    n_ordered = T * (T - 1)
    return {"T": T, "C": C, "ordered_pairs": n_ordered, "seconds": s_max, "ordered_pairs_per_s": n_ordered / s_max,
            "algorithmic_tflops": 2.0 * bench.P ** 2 * C * n_ordered / s_max / 1e12,
            "results_digest": hashlib.sha1(json.dumps(sorted(merged.items())).encode()).hexdigest()[:16],
            "limiter": "the affinity GEMMs (tensor-bound); pairs are independent, features replicated, no collective"}
