"""TEST INFRASTRUCTURE -- numpy restatement of ``scipy.ndimage.zoom(tile, zoom, order=3, mode="reflect"|"mirror")``
as the reference calls it for every lightcone tile (baryon_painter/process_SLICS.py:200, :213; the algorithm lives
in scipy: scipy/ndimage/src/ni_splines.c (spline prefilter) and ni_interpolation.c (NI_ZoomShift)).  It documents
what ``csrc/bp_zoom.cu`` computes and is pinned against scipy itself in tests/test_zoom.py.  Only tests import it."""
import numpy as np

POLE = np.sqrt(3.0) - 2.0


def _filter_line(c, mirror):
    """In-place cubic B-spline prefilter of one float64 line (ni_splines.c: apply_filter, _init_causal_*, _init_anticausal_*)."""
    n = len(c)
    if n < 2:
        return
    z = POLE
    c *= (1 - z) * (1 - 1 / z)
    if mirror:
        z_n_1 = z ** (n - 1)
        z_i, s = z, c[0] + z_n_1 * c[n - 1]
        for i in range(1, n - 1):
            s += z_i * (c[i] + z_n_1 * c[n - 1 - i])
            z_i *= z
        c[0] = s / (1 - z_n_1 * z_n_1)
    else:
        z_n = z ** n
        c0 = c[0]
        z_i, s = z, c[0] + z_n * c[n - 1]
        for i in range(1, n):
            s += z_i * (c[i] + z_n * c[n - 1 - i])
            z_i *= z
        c[0] = s * (z / (1 - z_n * z_n)) + c0
    for i in range(1, n):
        c[i] += z * c[i - 1]
    if mirror:
        c[n - 1] = (z * c[n - 2] + c[n - 1]) * z / (z * z - 1)
    else:
        c[n - 1] *= z / (z - 1)
    for i in range(n - 2, -1, -1):
        c[i] = z * (c[i + 1] - c[i])


def _fold(idx, n, mirror):
    """Boundary extension of a coefficient index (ni_interpolation.c, NI_ZoomShift offsets)."""
    if 0 <= idx < n:
        return idx
    if n <= 1:
        return 0
    if mirror:
        s2 = 2 * n - 2
        if idx < 0:
            idx = s2 * int(-idx / s2) + idx
            return idx + s2 if idx <= 1 - n else -idx
        idx -= s2 * int(idx / s2)
        return s2 - idx if idx >= n else idx
    s2 = 2 * n
    if idx < 0:
        if idx < -s2:
            idx += s2 * int(-idx / s2)
        return idx + s2 if idx < -n else -idx - 1
    idx -= s2 * int(idx / s2)
    return s2 - idx - 1 if idx >= n else idx


def _weights(x):
    f = np.floor(x)
    t = x - f
    t1 = 1 - t
    return int(f) - 1, np.array([t1 ** 3 / 6, (t * t * (t - 2) * 3 + 4) / 6, (t1 * t1 * (t1 - 2) * 3 + 4) / 6, t ** 3 / 6])


def zoom(tile, out_side, mode):
    """scipy.ndimage.zoom(tile, out_side / tile.shape[0], mode=mode) for a square float tile."""
    mirror = {"reflect": False, "mirror": True}[mode]
    c = np.array(tile, np.float64)
    n = c.shape[0]
    for r in range(n):
        _filter_line(c[r], mirror)
    for q in range(n):
        col = c[:, q].copy()
        _filter_line(col, mirror)
        c[:, q] = col
    scale = (n - 1) / (out_side - 1) if out_side > 1 else 0.0
    idx, wts = [], []
    for o in range(out_side):
        s, w = _weights(o * scale)
        idx.append([_fold(s + k, n, mirror) for k in range(4)])
        wts.append(w)
    idx, wts = np.array(idx), np.array(wts)
    rows = np.einsum("ok,okx->ox", wts, c[idx])                # (out, n): interpolate along y
    out = np.einsum("pk,opk->op", wts, rows[:, idx])           # (out, out): then along x
    return out.astype(np.asarray(tile).dtype)
