"""The C-ABI library loads and exports every symbol declared in include/baryon_painter_b200.h
(no compute calls: this runs without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from baryon_painter_b200 import build, _lib
    build.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    from baryon_painter_b200 import _lib
    header = open(os.path.join(ROOT, "include", "baryon_painter_b200.h")).read()
    declared = set(re.findall(r"\b(bp_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


def test_version_and_error_string(lib):
    assert lib.bp_version() == 100
    assert isinstance(lib.bp_last_error(), bytes)


def test_struct_layouts_match_header():
    """ctypes mirrors of the header structs have the C layout the header implies."""
    from baryon_painter_b200 import _lib
    assert ctypes.sizeof(_lib.LayerDesc) == 64 and _lib.LayerDesc.weight.offset == 40
    assert ctypes.sizeof(_lib.CvaeDesc) == 120 and _lib.CvaeDesc.prior.offset == 40 and _lib.CvaeDesc.q_x_in.offset == 88
    assert ctypes.sizeof(_lib.TransformParams) == 40 and _lib.TransformParams.k_in.offset == 24


def test_create_fails_loudly_without_gpu(lib):
    """No CPU fallback: without an sm_100 device creation raises (on a GPU box this is skipped)."""
    import numpy as np
    from baryon_painter_b200 import arch, synthetic
    from baryon_painter_b200.painter import CVAEPainter
    if lib.bp_device_count() > 0:
        pytest.skip("GPU present")
    A = arch.fiducial_cvae_architecture(64)
    p = CVAEPainter(architecture=A, precision="fp32")
    with pytest.raises(RuntimeError, match="no CUDA device|no CPU path"):
        p.model.load_state_dict(synthetic.synthetic_cvae_state_dict(A, seed=0))
    with pytest.raises(ValueError, match="no CPU path"):
        CVAEPainter(architecture=A, compute_device="cpu")
