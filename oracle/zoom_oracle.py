"""TEST INFRASTRUCTURE -- numpy restatement of ``scipy.ndimage.zoom(tile, zoom, order=3, mode="reflect"|"mirror")``
as the reference calls it for every lightcone tile (baryon_painter/process_SLICS.py:200, :213; the algorithm lives
in scipy: scipy/ndimage/src/ni_splines.c (spline prefilter) and ni_interpolation.c (NI_ZoomShift)).  It documents
what ``csrc/bp_zoom.cu`` computes and is pinned against scipy itself in tests/test_zoom.py.  Only tests import it."""
import numpy as np

POLES = {3: (np.sqrt(3.0) - 2.0,),
         5: (np.sqrt(67.5 - np.sqrt(4436.25)) + np.sqrt(26.25) - 6.5, np.sqrt(67.5 + np.sqrt(4436.25)) - np.sqrt(26.25) - 6.5)}


def _filter_line(c, mirror, poles):
    """In-place B-spline prefilter of one float64 line (ni_splines.c: apply_filter, _init_causal_*, _init_anticausal_*)."""
    n = len(c)
    if n < 2:
        return
    gain = 1.0
    for z in poles:
        gain *= (1 - z) * (1 - 1 / z)
    c *= gain
    for z in poles:
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


def _beta5(y):
    y = abs(y)
    if y < 1:
        return y * y * (y * y * (0.25 - y / 12) - 0.5) + 0.55
    if y < 2:
        return y * (y * (y * (y * (y / 24 - 0.375) + 1.25) - 1.75) + 0.625) + 0.425
    if y < 3:
        return (3 - y) ** 5 / 120
    return 0.0


def _weights(x, order):
    f = np.floor(x)
    t = x - f
    if order == 3:
        t1 = 1 - t
        return int(f) - 1, np.array([t1 ** 3 / 6, (t * t * (t - 2) * 3 + 4) / 6, (t1 * t1 * (t1 - 2) * 3 + 4) / 6, t ** 3 / 6])
    return int(f) - 2, np.array([_beta5(t + 2 - k) for k in range(6)])


def zoom(tile, out_side, mode, order=3):
    """scipy.ndimage.zoom(tile, out_side / tile.shape[0], order=order, mode=mode) for a square float tile."""
    mirror = {"reflect": False, "mirror": True}[mode]
    c = np.array(tile, np.float64)
    n = c.shape[0]
    for r in range(n):
        _filter_line(c[r], mirror, POLES[order])
    for q in range(n):
        col = c[:, q].copy()
        _filter_line(col, mirror, POLES[order])
        c[:, q] = col
    scale = (n - 1) / (out_side - 1) if out_side > 1 else 0.0
    idx, wts = [], []
    for o in range(out_side):
        s, w = _weights(o * scale, order)
        idx.append([_fold(s + k, n, mirror) for k in range(order + 1)])
        wts.append(w)
    idx, wts = np.array(idx), np.array(wts)
    rows = np.einsum("ok,okx->ox", wts, c[idx])                # (out, n): interpolate along y
    out = np.einsum("pk,opk->op", wts, rows[:, idx])           # (out, out): then along x
    return out.astype(np.asarray(tile).dtype)
