"""Small driver for ncu captures: one infer step of the bench workload (corr GEMM, head window, sample),
one corr_stream probe pass (Q_b = 1, all T frames) and one ViT-L block on one frame (GEMMs + fused attention)."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from bench_inputs import sharp_head  # noqa: E402


def main():
    from dino_tracker_b200 import ModelInference, Tracker, _lib
    from dino_tracker_b200.vit import DinoV2Features
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    dev = "cuda:0"
    T, C, nq = 50, 1024, 256
    if what in ("all", "infer", "stream"):
        feats = bench.synth_video_features(T, C, dev, 1234, 0.25)
        video = torch.zeros(T, 3, bench.H, bench.W, device=dev)
        m = Tracker(video=video, dino_embed_video=feats, device=dev, delta_channels=[3, 4, 4, 4, C])
        m.tracker_head.load_state_dict(sharp_head(0))
        mi = ModelInference(m, m.range_normalizer, 0.7, 0.6)
        q = bench.query_lattice(nq, 0).to(dev)
        if what in ("all", "infer"):
            mi.infer(q); mi.infer(q)
        if what in ("all", "stream"):
            class A: pass
            a = A(); a.T, a.C = T, C
            bench.stream_probe(m, mi, _lib.load(), _lib, a, bench.measured_peaks())
    if what in ("all", "vit"):
        g = torch.Generator(device=dev).manual_seed(7)
        dim, heads = 1024, 16
        def rn(*s): return torch.randn(*s, device=dev, generator=g) * 0.02
        sd = {"cls_token": rn(1, 1, dim), "pos_embed": rn(1, 1 + 37 * 37, dim), "patch_embed.proj.weight": rn(dim, 3, 14, 14),
              "patch_embed.proj.bias": rn(dim)}
        p = "blocks.0."
        sd.update({p + "norm1.weight": 1 + rn(dim), p + "norm1.bias": rn(dim), p + "attn.qkv.weight": rn(3 * dim, dim),
                   p + "attn.qkv.bias": rn(3 * dim), p + "attn.proj.weight": rn(dim, dim), p + "attn.proj.bias": rn(dim),
                   p + "ls1.gamma": 1 + rn(dim), p + "norm2.weight": 1 + rn(dim), p + "norm2.bias": rn(dim),
                   p + "mlp.fc1.weight": rn(4 * dim, dim), p + "mlp.fc1.bias": rn(4 * dim),
                   p + "mlp.fc2.weight": rn(dim, 4 * dim), p + "mlp.fc2.bias": rn(dim), p + "ls2.gamma": 1 + rn(dim)})
        ex = DinoV2Features(sd, heads=heads, layer=0, device=dev, frames_per_call=1)
        fr = torch.rand(1, 3, bench.H, bench.W, device=dev, generator=g)
        ex(fr); ex(fr)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
