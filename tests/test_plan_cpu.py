"""Host logic of dinotrk_infer without a GPU: the chunk planner (dinotrk_infer_plan) that cuts the trajectory phase and the
anchor phase into chunks of correlation maps and groups them by target frame.  Properties checked for random sizes:
every work item appears exactly once and in order, maps are numbered consecutively inside a chunk, chunks respect the map
budget, the number of chunks stays inside the bound that sizes the workspace."""
import ctypes

import numpy as np
import pytest

from dino_tracker_b200 import _lib

STREAM_MAX_M = 8   # groups with at most this many descriptors use the streaming kernel (csrc/corr.cuh)


def plan(kind, T, N, cnt, chunk):
    lib = _lib.load()
    gcap = T + 2
    max_chunks = int(lib.dinotrk_infer_max_chunks(T, N, chunk))
    groups = np.zeros((max_chunks, 5, gcap), dtype=np.int32)
    meta = np.zeros((max_chunks, 4), dtype=np.int32)
    n = ctypes.c_int(0)
    cnt_arr = None if cnt is None else np.ascontiguousarray(cnt, dtype=np.int32)
    rc = lib.dinotrk_infer_plan(kind, T, N, None if cnt_arr is None else cnt_arr.ctypes.data_as(ctypes.c_void_p), chunk,
                                groups.ctypes.data_as(ctypes.c_void_p), meta.ctypes.data_as(ctypes.c_void_p), max_chunks,
                                ctypes.byref(n))
    assert rc == 0, _lib.last_error() if hasattr(_lib, "last_error") else rc
    assert n.value <= max_chunks
    return groups[:n.value], meta[:n.value]


def check_chunk_invariants(groups, meta, chunk):
    for k in range(groups.shape[0]):
        used, maxm, ng, no_thin = (int(x) for x in meta[k])
        f, r, m, map0, item = (groups[k, j, :ng] for j in range(5))
        assert 0 < used <= chunk and ng >= 1
        assert int(m.sum()) == used and int(m.max()) == maxm and (m > 0).all()
        assert np.array_equal(map0, np.concatenate([[0], np.cumsum(m)[:-1]]))      # maps numbered consecutively
        assert bool(no_thin) == bool((m > STREAM_MAX_M).all())
        if k + 1 < groups.shape[0]:
            assert used == chunk                                                   # only the last chunk may be partial


@pytest.mark.parametrize("seed", range(12))
def test_trajectory_phase_plan(seed):
    rs = np.random.RandomState(seed)
    T, N = int(rs.randint(1, 40)), int(rs.randint(1, 300))
    chunk = int(rs.choice([1, 7, 64, 256, 1000, 4096, 32768]))
    groups, meta = plan(0, T, N, None, chunk)
    check_chunk_invariants(groups, meta, chunk)
    seen = []
    for k in range(groups.shape[0]):
        ng = int(meta[k, 2])
        for j in range(ng):
            t, row0, m = int(groups[k, 0, j]), int(groups[k, 1, j]), int(groups[k, 2, j])
            assert int(groups[k, 4, j]) == 0
            seen += [(t, n) for n in range(row0, row0 + m)]
    assert seen == [(t, n) for t in range(T) for n in range(N)]                     # every (frame, query) once, frame-major


@pytest.mark.parametrize("seed", range(12))
def test_anchor_phase_plan(seed):
    rs = np.random.RandomState(100 + seed)
    T, N = int(rs.randint(1, 30)), int(rs.randint(1, 200))
    cnt = rs.randint(0, N + 1, size=T)
    if seed % 4 == 0:
        cnt[rs.randint(0, T, size=max(1, T // 2))] = 0                              # frames without anchors
    chunk = int(rs.choice([1, 5, 50, 256, 4096, 32768]))
    groups, meta = plan(1, T, N, cnt, chunk)
    check_chunk_invariants(groups, meta, chunk)
    seen = []
    for k in range(groups.shape[0]):
        ng = int(meta[k, 2])
        for j in range(ng):
            a, row0, m, map0, item0 = (int(groups[k, i, j]) for i in range(5))
            assert row0 == map0                                                      # descriptor rows are per chunk
            seen += [(a, u) for u in range(item0, item0 + m)]
    assert seen == [(a, u) for a in range(T) for u in range(int(cnt[a]) * T)]       # every (anchor frame, item) once, in order
    assert groups.shape[0] == -(-sum(int(c) * T for c in cnt) // chunk)              # ceil(total / chunk) chunks


def test_plan_is_empty_without_anchors_and_counts_only():
    groups, meta = plan(1, 5, 9, np.zeros(5, dtype=np.int32), 64)
    assert groups.shape[0] == 0
    lib = _lib.load()
    n = ctypes.c_int(-1)
    assert lib.dinotrk_infer_plan(0, 50, 256, None, 32768, None, None, 0, ctypes.byref(n)) == 0   # count only
    assert n.value == 1                                                              # 12 800 maps fit one 32 768-map chunk
    cnt = np.full(50, 256, dtype=np.int32)
    assert lib.dinotrk_infer_plan(1, 50, 256, cnt.ctypes.data_as(ctypes.c_void_p), 32768, None, None, 0, ctypes.byref(n)) == 0
    assert n.value == 20 and n.value <= lib.dinotrk_infer_max_chunks(50, 256, 32768)  # 640 000 maps
