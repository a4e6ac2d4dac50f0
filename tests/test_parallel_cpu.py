"""CPU suite: the N>1 host logic (frame / query / pair sharding, in-place all-gather, result gather) under
``gloo`` with world_size 2 and 3."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, T, N):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dino_tracker_b200 import parallel as par
    torch.manual_seed(0)
    truth = torch.randn(T, 5, 4)                        # "refined features" every rank would compute
    full = torch.zeros_like(truth)

    def refine_block(s, e):
        full[s:e] = truth[s:e]

    def infer_rows(qs, qe):                               # a deterministic function of ALL frames
        rows = torch.arange(qs, qe, dtype=torch.float32)
        traj = rows[:, None, None] + full.sum(dim=(1, 2))[None, :, None].expand(qe - qs, T, 2)
        occ = (traj[..., 0] > 0)
        return traj.contiguous(), occ

    traj, occ = par.sharded_long_video_infer(T, N, world, rank, refine_block, infer_rows, full)
    assert torch.equal(full, truth), "all-gather did not reproduce the full feature video"
    exp = torch.arange(N, dtype=torch.float32)[:, None, None] + truth.sum(dim=(1, 2))[None, :, None].expand(N, T, 2)
    assert torch.equal(traj, exp) and torch.equal(occ, exp[..., 0] > 0)
    dist.destroy_process_group()


@pytest.mark.parametrize("world,T,N", [(2, 8, 5), (2, 7, 6), (3, 10, 7)])
def test_frame_sharded_pipeline_under_gloo(world, T, N):
    mp.spawn(_worker, args=(world, _free_port(), T, N), nprocs=world, join=True)


def test_shards_and_lpt():
    from dino_tracker_b200 import parallel as par
    for T in (1, 7, 50, 250):
        for w in (1, 2, 4, 8):
            blocks = [par.frame_shard(T, w, r) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == T
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            sizes = [e - s for s, e in blocks]
            assert max(sizes) - min(sizes) <= 1
    costs = [34 * 50, 104 * 376, 60 * 100, 80 * 200, 50 * 50, 90 * 300, 40 * 120, 70 * 70, 100 * 100]
    assign = par.lpt_assign(costs, 4)
    assert sorted(i for a in assign for i in a) == list(range(len(costs)))
    loads = [sum(costs[i] for i in a) for a in assign]
    assert max(loads) <= 1.5 * (sum(costs) / 4) or max(loads) == max(costs)


def test_run_videos_covers_every_video_once():
    """Multi-video launcher: LPT assignment, every video on exactly one rank, heaviest videos spread first."""
    from dino_tracker_b200.benchmark import run_videos
    ids = [f"v{i}" for i in range(7)]
    costs = [50 * 256, 30 * 100, 90 * 400, 10 * 10, 60 * 256, 24 * 50, 80 * 300]
    for world in (1, 2, 3, 8):
        seen, loads = [], []
        for rank in range(world):
            r = run_videos(ids, costs, rank, world, lambda v: costs[ids.index(v)])
            seen += list(r)
            loads.append(sum(r.values()))
        assert sorted(seen) == sorted(ids)
        if world == 2:
            assert max(loads) <= 0.6 * sum(costs)     # balanced: neither rank carries more than 60 %
