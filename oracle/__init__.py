"""CPU oracle for the DINO-Tracker inference hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import anything from this package, and only as the checker
or as the timed CPU baseline -- never as the thing shipped.  The product path
(``dino_tracker_b200``) never imports it and fails loudly if its CUDA library is
missing.

Every function is a plain PyTorch-fp32 (CPU) restatement of one reference
function and cites the reference ``file:line`` it follows (paths are relative to
the reference repo root, AssafSinger94/dino-tracker @ 5b0f2b0).

Pinning status (see DESIGN.md "Oracle"):
  * tracker / inference / delta-DINO / best-buddies rows:
    pinned.  ``oracle/make_golden.py`` imports the *live reference modules* in
    the build container (behind the three in-memory shims of
    ``oracle/ref_harness.py``), runs them on seeded synthetic inputs and commits
    inputs-by-seed + reference outputs under ``tests/golden/``; the CPU test
    suite checks this oracle against those vectors.
  * ViT row (a1): pinned for everything the reference itself defines, anchored on a
    second source for the rest.  ``tests/golden/vit_small.npz`` and
    ``posembed.npz`` come from the LIVE reference (``utils.get_dino_features_video``
    + ``models/extractor.VitExtractor``: normalisation, re-strided patch convolution,
    position-embedding fix, block hooks, tap point, cls drop, layout) with only
    ``torch.hub.load`` replaced: the DINOv2 network itself is a third-party module
    (facebookresearch/dinov2, fetched by torch.hub at an unpinned ref,
    ``models/extractor.py:26``) that is absent from the reference tree and from this
    image and cannot be downloaded.  The stand-in exposes the DinoVisionTransformer
    surface the extractor touches and runs ``transformers``' Dinov2Layer blocks;
    ``oracle/vit.py`` -- which restates the published DINOv2 block -- reproduces the
    stored features exactly and agrees with Dinov2Layer block by block
    (``tests/test_vit_oracle_cpu.py``).  What stays unpinned: that the hub
    checkpoint's own block code equals the published definition both other
    implementations follow.
"""


def use_exact_fp32():
    """fp32 arithmetic for the oracle on a CUDA device: TF32 off for matmuls and convolutions (the
    tensor-core shortcuts PyTorch may otherwise take), so that ``oracle.*`` on ``cuda`` is the same
    fp32 computation as on the CPU up to summation order."""
    import torch
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        torch.set_float32_matmul_precision("highest")
    except Exception:  # pragma: no cover
        pass
