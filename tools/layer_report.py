"""Per-layer parity report against tests/golden/cvae_t64_layers.npz (GPU).  Usage:
    python tools/layer_report.py fp16"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baryon_painter_b200 import arch, synthetic          # noqa: E402
from baryon_painter_b200.painter import CVAEPainter      # noqa: E402


def rel_l2(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.sqrt(((a - b) ** 2).sum() / max((b ** 2).sum(), 1e-300)))


def main(precision):
    g = np.load(os.path.join(ROOT, "tests", "golden", "cvae_t64_layers.npz"))
    p = CVAEPainter.synthetic(tile_size=64, seed=int(g["seed"]), precision=precision, max_batch=8)
    p.model.net.set_debug(True)
    tiles = synthetic.synthetic_dm_tiles(1, 64, seed0=int(g["tiles_seed0"]))
    out = p.paint(tiles[0], z=float(g["z"][0]), eps=g["eps"][0])
    print("precision", precision, "painted rel-L2", rel_l2(out, g["painted_E"][0]))
    for sid, name in enumerate(("prior_network", "p_z_in", "p_y_z_in", "p_mu_out")):
        specs = p.model.stacks[name]
        keys = sorted((k for k in g.files if k.startswith("tap:" + name + ".")), key=lambda s: int(s.split(".")[-1]))
        li = -1
        for k in keys:
            li += 1
            while specs[li].res == arch.RES_OPEN:
                li += 1
            ref = g[k]
            if name == "p_mu_out" and li == len(specs) - 1:
                ref = p.inverse_transform(ref[None], field="pressure", z=float(g["z"][0]))[None]
            got = p.model.net.read_activation(sid, li, (1, *ref.shape))[0]
            info = p.model.net.layer_info(sid, li)
            print("%-14s L%-2d %-18s tensor=%d  rel-L2 %.3e   |ref| %.3f |got| %.3f  nan=%d" % (
                name, li, k, info["tensor"], rel_l2(got, ref), np.abs(ref).mean(), np.abs(got).mean(),
                int(np.isnan(got).sum())))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "fp16")
