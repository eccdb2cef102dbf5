#!/usr/bin/env python
"""Latency of the reference-shaped call: one tile in, one tile out (CVAEPainter.paint, painter.py:371-392)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from baryon_painter_b200 import synthetic  # noqa: E402
from baryon_painter_b200.painter import CVAEPainter  # noqa: E402

for mb in (1, 16):
    p = CVAEPainter.synthetic(compute_device="cuda:0", max_batch=mb)
    tile = synthetic.synthetic_dm_tiles(1, 512, seed0=0)[0]
    for _ in range(5):
        p.paint(tile, z=0.3)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 50
    for _ in range(n):
        p.paint(tile, z=0.3)
    dt = (time.perf_counter() - t0) / n
    print("paint() one 512x512 tile, max_batch=%d: %.2f ms per call (%.0f tiles/s)" % (mb, dt * 1e3, 1 / dt))
