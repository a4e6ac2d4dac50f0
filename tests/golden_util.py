"""Helpers shared by the CPU (oracle-vs-golden) and GPU (CUDA-vs-oracle/golden) tests."""
import os

import numpy as np
import torch

from oracle import make_golden, synth
from oracle.tracker import Geometry

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TRACK_CASES = make_golden.TRACK_CASES


def load_track_case(name):
    """Returns (geo, features, head_sd, golden dict).  Inputs are regenerated from the seed and
    checked against the fixture's checksum (or the stored copy when it is small)."""
    cfg = TRACK_CASES[name]
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    geo, feats, head, q = make_golden.case_inputs(cfg)
    if "features" in g:
        assert np.array_equal(g["features"], feats.numpy()), "seeded inputs drifted from fixture"
    cs = np.array([feats.double().sum().item(), feats.double().abs().sum().item()])
    assert np.allclose(cs, g["feat_checksum"], rtol=1e-12), "seeded inputs drifted from fixture"
    assert np.array_equal(g["query_points"], q.numpy())
    for k, v in head.items():
        assert np.array_equal(g["head." + k], v.numpy())
    return cfg, geo, feats, head, g
