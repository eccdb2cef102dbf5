"""TEST INFRASTRUCTURE -- imports the *unmodified* reference from /root/reference.

Only usable in the build container (the GPU box has no /root/reference); used by
``oracle/make_golden.py`` to generate the committed golden vectors and by the
``not gpu`` tests that pin ``oracle/cvae_oracle.py`` against the reference itself.

The reference cannot be imported as-is here: it pulls in
``matplotlib``, ``cosmotools``, ``pyccl`` and ``astropy`` at module import time, none of
which is installed and none of which the paint path uses.  Empty stub modules are put
in ``sys.modules`` for exactly those names; no reference file is modified or copied.
"""

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("BARYON_PAINTER_REFERENCE", "/root/reference")

_STUBS = ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "cosmotools", "cosmotools.utils",
          "cosmotools.power_spectrum_tools", "cosmotools.plotting", "pyccl", "astropy",
          "astropy.io", "astropy.io.fits")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "baryon_painter"))


def import_reference():
    """-> the reference ``baryon_painter`` package (painter, models, process_SLICS loaded)."""
    if not available():
        raise RuntimeError("reference checkout not found at %s" % REFERENCE_ROOT)
    for name in _STUBS:
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
    sys.modules["cosmotools.utils"].rebin_2d = None          # data_transforms.py:5 (unused name)
    for parent, child in (("matplotlib", "pyplot"), ("cosmotools", "utils"),
                          ("cosmotools", "power_spectrum_tools"), ("cosmotools", "plotting"),
                          ("astropy", "io"), ("astropy.io", "fits")):
        setattr(sys.modules[parent], child, sys.modules[parent + "." + child])
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import baryon_painter                                    # noqa: F401
    import baryon_painter.painter                            # noqa: F401
    import baryon_painter.models.cvae                        # noqa: F401
    import baryon_painter.process_SLICS                      # noqa: F401
    import baryon_painter.utils.data_transforms              # noqa: F401
    import baryon_painter.utils.datasets                     # noqa: F401
    return sys.modules["baryon_painter"]


def reference_painter(architecture, state_dict, stats):
    """A reference ``CVAEPainter`` on CPU with ``state_dict`` loaded and the transforms rebuilt
    with the reference's own factories (``scripts/CVAE_single_scale.py:34-65``), bound with
    ``datasets.compile_transform``.  The forward transform is wrapped with a float32 cast
    (NumPy>=2 promotion shim)."""
    import contextlib
    import io
    import numpy as np
    bp = import_reference()
    dt, ds = bp.utils.data_transforms, bp.utils.datasets
    with contextlib.redirect_stdout(io.StringIO()):
        painter = bp.painter.CVAEPainter(architecture=architecture, compute_device="cpu")
    painter.model.load_state_dict(state_dict)
    rc, rc_inv = dt.create_range_compress_transforms(
        k_values={"dm": 4.0, "pressure": 4}, modes={"dm": "shift-log", "pressure": "shift-log"}, eps=1e-4)
    fwd = ds.compile_transform(dt.chain_transformations([rc, dt.atleast_3d]), stats)
    inv = ds.compile_transform(dt.chain_transformations([dt.squeeze, rc_inv]), stats)
    painter.transform = lambda x, field=None, z=None: np.asarray(fwd(x, field, z), dtype=np.float32)
    painter.inverse_transform = inv
    painter.input_field, painter.label_fields = "dm", ["pressure"]
    return painter
