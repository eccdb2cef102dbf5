#!/usr/bin/env python
"""Per-layer timing of the CGAN generator (device buffers resident)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from baryon_painter_b200 import synthetic  # noqa: E402
from baryon_painter_b200.painter import CGANPainter  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
g = CGANPainter.synthetic(device="cuda:0", max_batch=n)
net = g.model.net
tiles = synthetic.synthetic_dm_tiles(min(n, 8), 512, seed0=0)
tiles = np.concatenate([tiles] * (n // len(tiles) + 1))[:n]
g.paint_batch(tiles, z=0.0)
net.set_profile(True)
g.paint_batch(tiles, z=0.0)
torch.cuda.synchronize()
tot = 0
rows = []
for li in range(len(g.model.specs)):
    info = net.layer_info(0, li)
    ms, cnt = net.read_profile(0, li)
    rows.append((li, info, ms, cnt))
    tot += ms
print("flops/tile %.3f G, chunk %d" % (net.flops_per_tile / 1e9, net.chunk))
for li, info, ms, cnt in rows:
    tf = info["flops"] * n / (ms * 1e-3) / 1e12 if ms else 0
    print("L%-2d kind=%d k%d s%d %3d->%3d @%3dx%-3d tensor=%d %8.3f ms %5.1f%% %7.1f TFLOP/s launches=%d" % (
        li, info["kind"], info["kernel"], info["stride"], info["cin"], info["cout"], info["H"], info["W"], info["tensor"], ms,
        100 * ms / tot, tf, cnt))
print("total %.2f ms for %d tiles -> %.0f tiles/s (layers only)" % (tot, n, n / tot * 1e3))
