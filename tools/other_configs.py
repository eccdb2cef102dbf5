#!/usr/bin/env python
"""Throughput of BASELINE.json configs 3 and 4 through the public API (host buffers in and out):
   config 3: fiducial CGAN generator, 256 synthetic tiles across z in {0, 0.5, 1}
   config 4: CVAE variance maps, 64 latent draws per tile.
Prints one JSON line per config (not the bench.py contract line: these are reported in DESIGN.md)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import baryon_painter_b200 as bp  # noqa: E402
from baryon_painter_b200 import synthetic  # noqa: E402
from baryon_painter_b200.painter import CGANPainter, CVAEPainter  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
base = synthetic.synthetic_dm_tiles(16, 512, seed0=0)
tiles = bp.pinned_empty((n, 512, 512))
tiles[...] = np.concatenate([base] * (n // 16 + 1))[:n] * np.linspace(0.8, 1.25, n, dtype=np.float32)[:, None, None]
zs = np.array([0.0, 0.5, 1.0])[np.arange(n) % 3]


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


g = CGANPainter.synthetic(device="cuda:0", max_batch=n)
out = bp.pinned_empty((n, 512, 512))
t = timed(lambda: g.paint_batch(tiles, z=zs, out=out))
print(json.dumps({"config": "CGAN generator, %d tiles, z in {0,0.5,1}, host buffers" % n, "tiles_per_s": n / t, "ms": t * 1e3}))
del g
torch.cuda.empty_cache()
c = CVAEPainter.synthetic(compute_device="cuda:0", max_batch=256)
nv = 16
t = timed(lambda: c.paint_variance(tiles[:nv], z=0.0, n_draws=64, seed=1), reps=2)
print(json.dumps({"config": "CVAE variance maps, %d tiles x 64 draws, host buffers" % nv, "draws_per_s": nv * 64 / t,
                  "tiles_per_s": nv / t, "ms": t * 1e3}))
