"""TEST INFRASTRUCTURE -- golden painted planes from the UNMODIFIED reference ``process_SLICS``.

    python -m oracle.make_golden_slics          (build container only: needs /root/reference, ~1.1 GB of /tmp)

Writes synthetic SLICS-format files (one 12288^2 mass plane, two 7745^2 delta maps, the random-shift table)
to a temporary directory, runs the reference's ``process_SLICS`` on them with ``StubPainter`` and a 48-pixel
tile, and stores the painted planes in tests/golden/slics_small.npz together with the seeds, so that the
tests can rebuild the same planes in memory (``plane_source``) without the files.
"""
import contextlib
import io
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shims, slics_oracle          # noqa: E402

CASE = dict(LOS=7, tile_size=100.0, n_pixel_tile=48, z_SLICS=[0.042, 0.130, 0.221], z_slice=[0.04, 0.13, 0.22],
            delta_size=[60.0, 150.0, 260.0], seeds=[11, 12, 13],
            shifts=[[0.81, 0.13], [0.35, 0.62], [0.07, 0.44]])


def main():
    bp = ref_shims.import_reference()
    ps = bp.process_SLICS
    tmp = tempfile.mkdtemp(prefix="slics_golden_")
    try:
        c = CASE
        # reference reads shifts[::-1][i]
        np.savetxt(os.path.join(tmp, f"random_shift_LOS{c['LOS']}"), np.array(c["shifts"])[::-1])
        for i, z in enumerate(c["z_SLICS"]):
            if c["delta_size"][i] < c["tile_size"]:
                axes = ["xy", "xz", "yz"][i % 3]
                content = slics_oracle.synthetic_massplane_file_content(c["seeds"][i])
                with open(os.path.join(tmp, f"{z:.3f}proj_half_finer_{axes}.dat_LOS{c['LOS']}"), "wb") as f:
                    np.zeros(1, np.float32).tofile(f)
                    content.tofile(f)
            else:
                slics_oracle.synthetic_delta_file_content(c["seeds"][i]).tofile(
                    os.path.join(tmp, f"{z:.3f}delta.dat_bicubic_LOS{c['LOS']}"))
        with contextlib.redirect_stdout(io.StringIO()):
            planes = ps.process_SLICS(slics_oracle.StubPainter(), c["tile_size"], c["n_pixel_tile"], c["LOS"], c["z_SLICS"],
                                      c["delta_size"], tmp, tmp, tmp, c["z_slice"], verbose=False)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    out = {f"plane{i}": np.asarray(p, np.float64) for i, p in enumerate(planes)}
    for k, v in CASE.items():
        out["case_" + k] = np.asarray(v)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "slics_small.npz"), **out)
    print({k: v.shape for k, v in out.items() if k.startswith("plane")})


if __name__ == "__main__":
    main()
