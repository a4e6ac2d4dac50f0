"""BASELINE config 4 driver: one long synthetic video, frames block-sharded over the ranks of one node.
Each rank runs ViT + delta-DINO for its frames straight into its slice of the full [T][P][C] buffer, ONE in-place
NCCL all-gather makes every frame available everywhere, query points are sharded, results gathered.
Rank 0 re-computes everything alone and checks that the sharded run is identical.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
         tools/run_config4.py --T 24 --nq 64
"""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from bench_inputs import sharp_head  # noqa: E402


def random_vit_sd(depth, dim, dev, seed=3):
    g = torch.Generator(device=dev).manual_seed(seed)

    def rn(*s, std=0.05):
        return torch.randn(*s, device=dev, generator=g) * std
    sd = {"cls_token": rn(1, 1, dim), "pos_embed": rn(1, 1 + 16, dim), "patch_embed.proj.weight": rn(dim, 3, 14, 14),
          "patch_embed.proj.bias": rn(dim)}
    for i in range(depth):
        p = f"blocks.{i}."
        sd.update({p + "norm1.weight": 1 + rn(dim), p + "norm1.bias": rn(dim), p + "attn.qkv.weight": rn(3 * dim, dim),
                   p + "attn.qkv.bias": rn(3 * dim), p + "attn.proj.weight": rn(dim, dim), p + "attn.proj.bias": rn(dim),
                   p + "ls1.gamma": 1 + rn(dim), p + "norm2.weight": 1 + rn(dim), p + "norm2.bias": rn(dim),
                   p + "mlp.fc1.weight": rn(4 * dim, dim), p + "mlp.fc1.bias": rn(4 * dim),
                   p + "mlp.fc2.weight": rn(dim, 4 * dim), p + "mlp.fc2.bias": rn(dim), p + "ls2.gamma": 1 + rn(dim)})
    return sd


def run():
    ap = argparse.ArgumentParser()
    ap.add_argument("--T", type=int, default=24)
    ap.add_argument("--nq", type=int, default=64)
    ap.add_argument("--dim", type=int, default=256)
    ap.add_argument("--depth", type=int, default=2)
    ap.add_argument("--check", type=int, default=1)
    ap.add_argument("--fused", type=int, default=0, help="1: all-gather fused into the delta-DINO epilogue (peer stores)")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    from dino_tracker_b200 import ModelInference, Tracker, parallel as par
    from dino_tracker_b200.vit import DinoV2Features
    T, N, C = a.T, a.nq, a.dim
    H, W = bench.H, bench.W
    g = torch.Generator().manual_seed(11)
    video = torch.rand(T, 3, H, W, generator=g)            # same video on every rank (a real run would read its own frames)
    sd = random_vit_sd(a.depth, C, dev)
    vit = DinoV2Features(sd, heads=C // 64, layer=a.depth - 1, device=dev)
    q = bench.query_lattice(N, 0).to(dev)
    q[:, 2] = torch.arange(N, device=dev) % T

    def build_tracker(feats_tpc_chw):
        torch.manual_seed(5)   # identical delta-DINO weights (incl. default-initialised biases) on every rank / build
        m = Tracker(video=video.to(dev), dino_embed_video=feats_tpc_chw, device=dev, delta_channels=[3, 16, 16, 16, C])
        for p_ in m.delta_dino.parameters():
            torch.nn.init.normal_(p_, std=0.05) if p_.dim() > 1 else None
        m.tracker_head.load_state_dict(sharp_head(0))
        return m

    P = 67 * 121
    pbuf = par.PeerFeatureBuffer(T, P, C, rank, world) if a.fused else None
    full = pbuf.tensor if a.fused else torch.zeros(T, P, C, device=dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s, e = par.frame_shard(T, world, rank)
    # stage 0/1 for this rank's frames: ViT features, then delta-DINO refinement, written into the shared buffer slice
    dino_local = vit(video[s:e])                                             # [e-s][P][C]
    m_local = build_tracker(dino_local.view(e - s, 67, 121, C).permute(0, 3, 1, 2))
    m_local.video = video[s:e].to(dev)
    if a.fused:
        # refined rows go to this rank's slice AND, from inside the producing kernel, to every peer's buffer
        m_local.delta_dino.refine_tpc(video[s:e].to(dev), m_local._dino_tpc, m_local._geom, out=full[s:e],
                                      peer_ptrs=pbuf.peer_ptrs, first_frame=s)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        pbuf.sync()
    else:
        m_local.cache_refined_embeddings()
        full[s:e] = m_local._refined_tpc
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        par.allgather_frames(full, T, world, rank)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    # tracker over all frames, this rank's query rows
    m_all = build_tracker(torch.zeros(T, C, 67, 121))
    m_all._refined_tpc = full
    m_all._refined_norms = full.norm(dim=2).contiguous()
    mi = ModelInference.__new__(ModelInference)
    torch.nn.Module.__init__(mi)
    mi.model, mi.range_normalizer = m_all, m_all.range_normalizer
    mi.anchor_cosine_similarity_threshold, mi.cosine_similarity_threshold = 0.7, 0.6
    qs, qe = par.query_shard(N, world, rank)
    traj, occ = mi.infer(q[qs:qe])
    traj = par.gather_rows(traj.contiguous(), N, world, rank)
    occ = par.gather_rows(occ.to(torch.uint8), N, world, rank).bool()
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    res = {"world": world, "T": T, "nq": N, "C": C, "fused_allgather": bool(a.fused), "features_s": t1 - t0, "allgather_s": t2 - t1,
           "allgather_GB": full.numel() * 4 / 1e9, "track_s": t3 - t2}
    if a.check and rank == 0:
        dino_all = vit(video)
        m_ref = build_tracker(dino_all.view(T, 67, 121, C).permute(0, 3, 1, 2))
        mi_ref = ModelInference(m_ref, m_ref.range_normalizer, 0.7, 0.6)
        t_ref, o_ref = mi_ref.infer(q)
        res["max_feature_diff"] = (m_ref._refined_tpc - full).abs().max().item()
        res["max_traj_diff"] = (t_ref - traj).abs().max().item()
        res["occ_mismatch"] = int((o_ref != occ).sum())
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    run()
