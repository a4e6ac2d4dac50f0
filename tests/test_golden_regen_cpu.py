"""Where the live reference is present (the build container), re-run it and check that the committed training-step and
cycle-consistency fixtures are what it returns today, bit for bit (`python -m oracle.make_golden` rewrites every fixture;
these two are fast enough for the CPU suite).  Skipped where /root/reference does not exist (the GPU box)."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_harness

from golden_util import GOLDEN_DIR

pytestmark = pytest.mark.skipif(not ref_harness.reference_available(), reason="live reference not present")


def _same(a, b):
    return set(a.files) == set(b.files) and all(np.array_equal(a[k], b[k], equal_nan=True) for k in a.files)


@pytest.mark.parametrize("name,gen", [("train_small", "gen_train_case"), ("cycle_small", "gen_cycle_case")])
def test_fixture_regenerates_from_the_live_reference(name, gen, tmp_path, monkeypatch):
    from oracle import make_golden as mg
    torch.set_num_threads(8)
    monkeypatch.setattr(mg, "GOLDEN_DIR", str(tmp_path))
    getattr(mg, gen)(name)
    assert _same(np.load(os.path.join(GOLDEN_DIR, name + ".npz")), np.load(os.path.join(str(tmp_path), name + ".npz")))
