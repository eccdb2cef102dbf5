"""Per-tile parity of the CUDA path against the committed golden vectors (GPU).  Usage:
    python tools/parity_report.py fp16 [128|512]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baryon_painter_b200 import synthetic          # noqa: E402
from baryon_painter_b200.painter import CVAEPainter      # noqa: E402


def rel_l2(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.sqrt(((a - b) ** 2).sum() / max((b ** 2).sum(), 1e-300)))


def main(precision, tile):
    g = np.load(os.path.join(ROOT, "tests", "golden", "cvae_t%d.npz" % tile))
    n = len(g["z"])
    p = CVAEPainter.synthetic(tile_size=tile, seed=int(g["seed"]), precision=precision, max_batch=8)
    tiles = synthetic.synthetic_dm_tiles(n, tile, seed0=int(g["tiles_seed0"]))
    zs = g["z"]
    out = p.paint_batch(tiles, z=zs, eps=g["eps"])
    xm = p.paint_batch(tiles, z=zs, eps=g["eps"], inverse_transform=False)
    for i in range(min(n, len(g["painted_E"]))):
        ref = g["painted_E"][i]
        err = np.abs(out[i].astype(np.float64) - ref)
        j = np.unravel_index(np.argmax(err), err.shape)
        # x_mu implied by the golden painted tile: x = ln(p/sigma + 1)/4
        sig = p.inverse_transform.gpu_params("pressure", float(zs[i]))[1]
        xref = np.log(ref.astype(np.float64) / sig + 1) / 4
        xe = np.abs(xm[i, 0] - xref)
        print("%s t%d tile %d z=%.2f: painted rel-L2 %.3e  max|err| %.3e at %s (ref %.3e)  x_mu rel-L2 %.3e max|dx| %.3e  "
              "ref max %.3e, x max %.3f" % (precision, tile, i, zs[i], rel_l2(out[i], ref), err.max(), j, ref[j],
                                             rel_l2(xm[i, 0], xref), xe.max(), ref.max(), xref.max()))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "fp16", int(sys.argv[2]) if len(sys.argv) > 2 else 128)
