"""CPU-side checks of the drop-in boundary: the library builds, loads and exports every
symbol include/mmloam_b200.h declares; no compute call is made (no GPU here)."""
import ctypes
import os
import re

import pytest


def test_library_exports_every_declared_symbol(mm):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "mmloam_b200.h")).read()
    declared = set(re.findall(r"\b(mml_[a-z0-9_]+)\s*\(", header))
    declared -= {"mml_ctx", "mml_est_params"}
    assert len(declared) >= 28
    lib = mm.load_library()
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, f"symbols declared in the header but not exported: {missing}"
    assert set(mm.EXPORTS) <= declared


def test_no_cpu_fallback(mm):
    """Without a CUDA device context creation must fail loudly, never fall back."""
    lib = mm.load_library()
    h = ctypes.c_void_p()
    rc = lib.mml_ctx_create(0, 1, ctypes.byref(h))
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        assert rc == 0
        lib.mml_ctx_destroy(h)
    else:
        assert rc == -2  # MML_ERR_NO_DEVICE
        with pytest.raises(mm.MmlError):
            mm.Context(0)


def test_product_never_imports_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "multi-modal-loam_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp", ".sh")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                # comments may say "bit-identical to the CPU oracle"; code may not import, link or call it
                bad = re.findall(r"import\s+oracle|from\s+oracle|oracle/|oracle\.py|liboracle|mmloam_oracle|\borc_[a-z]", src)
                assert not bad, f"{f} depends on the oracle ({bad}): the product path must not use oracle/"


def test_est_params_default(mm):
    p = mm.est_params()
    assert (p.max_outer, p.max_inner) == (5, 10)
    assert p.lidar_m == 1.5e-3 and p.plan_weight_tan == 0.0
    assert (p.thres0, p.thres1, p.thres2) == (25.0, 10.0, 1.0)
