#!/usr/bin/env python
"""Pinned host<->device copy bandwidth of the box (the ceiling of bench.py's e2e figure)."""
import time
import torch

n = 256 * 512 * 512
h_in = torch.empty(n, dtype=torch.float32).pin_memory()
h_out = torch.empty(n, dtype=torch.float32).pin_memory()
d_a = torch.empty(n, dtype=torch.float32, device="cuda")
d_b = torch.empty(n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
gb = n * 4 / 1e9


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_a.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_b, non_blocking=True)


def both():
    h2d(); d2h()


t = timed(h2d); print("H2D  %.2f ms  %.1f GB/s" % (t * 1e3, gb / t))
t = timed(d2h); print("D2H  %.2f ms  %.1f GB/s" % (t * 1e3, gb / t))
t = timed(both); print("both %.2f ms  %.1f GB/s each direction" % (t * 1e3, gb / t))
