"""Per-layer relative L2 error of the CUDA path against the committed t64 layer-boundary goldens.
    python tools/layer_errors.py fp32 fp32-ffma fp16"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baryon_painter_b200 import arch, synthetic          # noqa: E402
from baryon_painter_b200.painter import CVAEPainter      # noqa: E402


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.sqrt(((a - b) ** 2).sum() / max((b ** 2).sum(), 1e-300)))


g = np.load(os.path.join(ROOT, "tests", "golden", "cvae_t64_layers.npz"))
for prec in sys.argv[1:] or ["fp32"]:
    p = CVAEPainter.synthetic(tile_size=64, seed=int(g["seed"]), precision=prec, max_batch=8)
    p.model.net.set_debug(True)
    tiles = synthetic.synthetic_dm_tiles(1, 64, seed0=int(g["tiles_seed0"]))
    out = p.paint(tiles[0], z=float(g["z"][0]), eps=g["eps"][0])
    print("== %s: painted %.3e" % (prec, rel(out, g["painted_E"][0])))
    for name, sid in (("prior_network", 0), ("p_z_in", 1), ("p_y_z_in", 2), ("p_mu_out", 3)):
        specs = p.model.stacks[name]
        keys = sorted((k for k in g.files if k.startswith("tap:" + name + ".")), key=lambda s: int(s.split(".")[-1]))
        li = -1
        for k in keys:
            li += 1
            while specs[li].res == arch.RES_OPEN:
                li += 1
            ref = g[k]
            got = p.model.net.read_activation(sid, li, (1, *ref.shape))[0]
            if name == "p_mu_out" and li == len(specs) - 1:
                ref = p.inverse_transform(ref[None], field="pressure", z=float(g["z"][0]))[None]
            print("   %-22s layer %2d  rel-L2 %.3e   max|ref| %.3g" % (k, li, rel(got, ref), np.abs(ref).max()))
    mu, lv = p.model.net.cvae_read_prior(1)
    print("   z_mu %.3e  z_log_var %.3e" % (rel(mu[0], g["z_mu"][0][0]), rel(lv[0], g["z_log_var"][0][0])))
