"""CPU suite: the C-ABI library builds, loads and exports everything include/dinotrk.h declares;
the product path refuses to run without CUDA (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "dinotrk.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dinotrk_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from dino_tracker_b200 import _lib
    lib = ctypes.CDLL(_lib.lib_path())
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/dinotrk.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), "ctypes SIGNATURES out of sync with include/dinotrk.h"
    assert _lib.load().dinotrk_version() >= 100


def test_geom_helper_matches_reference_token_grid():
    from dino_tracker_b200 import _lib
    g = _lib.make_geom(476, 854)
    assert (g.h, g.w) == (67, 121)  # models/extractor.py:171-177
    g = _lib.make_geom(98, 126)
    assert (g.h, g.w) == (13, 17)
    with pytest.raises(_lib.DinotrkError):
        _lib.make_geom(8, 8)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_path_fails_loudly_without_cuda():
    from dino_tracker_b200 import Tracker, _lib
    with pytest.raises(_lib.DinotrkError):
        Tracker(video=torch.zeros(2, 3, 98, 126), dino_embed_video=torch.zeros(2, 8, 13, 17), device="cuda:0")
    with pytest.raises(_lib.DinotrkError):
        Tracker(video=torch.zeros(2, 3, 98, 126), dino_embed_video=torch.zeros(2, 8, 13, 17), device="cpu")


def test_dropin_models_package_resolves_to_b200_classes():
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); import models.tracker as t, models.model_inference as m; "
            "import dino_tracker_b200 as d; assert t.Tracker is d.Tracker and m.ModelInference is d.ModelInference; "
            "assert callable(m.generate_trajectory_input); print('ok')") % os.path.join(ROOT, "dino_tracker_b200", "dropin")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr
