"""TEST INFRASTRUCTURE -- numpy restatement of the device's counter-based standard-normal generator.

The reference draws the latent noise with ``torch.randn`` (reference baryon_painter/models/cvae.py:64), which is
not reproducible across devices; the CUDA path's BP_LATENT_SEED mode replaces it by
``eps_i = box_muller(splitmix64(seed ^ splitmix64(offset + i)))`` (csrc/bp_f32.cu ``counter_normal``).  This file
restates that generator so a test can hand the SAME eps to the oracle's ``sample_P`` and compare per-pixel
mean / variance maps (BASELINE.json configs[3]); the distribution itself is checked against N(0, 1) separately.
Only tests/ may import this.
"""
import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x):
    x = (np.asarray(x, np.uint64) + np.uint64(0x9E3779B97F4A7C15)) & _M
    x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M
    x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M
    return x ^ (x >> np.uint64(31))


def counter_normal(seed, offset, n):
    """eps[i], i < n: the draws the device makes for counters offset .. offset + n - 1 (float32 arithmetic as on
    the device up to the last ulp of logf / cospif)."""
    with np.errstate(over="ignore"):
        ctr = (np.uint64(offset) + np.arange(n, dtype=np.uint64)) & _M
        r = splitmix64(np.uint64(seed & 0xFFFFFFFFFFFFFFFF) ^ splitmix64(ctr))
    u1 = ((r >> np.uint64(40)).astype(np.float32) + np.float32(1)) * np.float32(1.0 / 16777216.0)
    u2 = ((r >> np.uint64(8)) & np.uint64(0xFFFFFF)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    rad = np.sqrt(np.float32(-2) * np.log(u1, dtype=np.float32), dtype=np.float32)
    return (rad * np.cos(2.0 * np.pi * u2.astype(np.float64)).astype(np.float32)).astype(np.float32)


def variance_draw_eps(seed, n_tiles, n_draws, latent_hw):
    """eps[d][t] (n_draws, n_tiles, 1, h, w) of ``paint_variance(tiles[:n_tiles], n_draws=, seed=)`` for one group of
    tiles (n_tiles <= the plan chunk): counter = (d * n_tiles + t) * h*w + p  (csrc/bp_net.cu,
    bp_cvae_paint_variance_host)."""
    h, w = latent_hw
    e = counter_normal(seed, 0, n_draws * n_tiles * h * w)
    return e.reshape(n_draws, n_tiles, 1, h, w)
