"""TEST INFRASTRUCTURE -- CPU restatement of the lightcone tile loop and its synthetic inputs.

``NumpyBackend`` is the numpy counterpart of ``baryon_painter_b200.process_SLICS.DeviceBackend`` (paint loop,
``painted_plane[slice] += w*tile; weight_plane[slice] += w``, ``plane = painted/weight`` in float64, as reference
baryon_painter/process_SLICS.py:201-220).  Only tests, ``oracle/make_golden_slics.py`` and ``bench.py``'s CPU legs
may import it; the product path (``DeviceBackend``) never does.
"""

import numpy as np

from baryon_painter_b200.process_SLICS import MASS_NORM, make_weight_map


class StubPainter:
    """Deterministic stand-in with the reference ``paint`` signature (process_SLICS is duck-typed)."""
    compute_device = None

    def paint(self, input, z=0.0, transform=True, inverse_transform=True):
        return (np.sqrt(np.abs(np.asarray(input, np.float32))) * np.float32(1.0 + z)).astype(np.float32)


class NumpyBackend:
    def new_planes(self, n_pixel_plane):
        return np.zeros((2, n_pixel_plane, n_pixel_plane))

    def paint(self, painter, tiles, z, batch):
        return np.stack([np.asarray(painter.paint(input=t, z=z, transform=True, inverse_transform=True)) for t in tiles])

    def accumulate(self, planes, painted, origins, falloff, sigma):
        T = painted.shape[1]
        w = make_weight_map((T, T), falloff=falloff, sigma=sigma)
        for p, (y0, x0) in zip(painted, origins):
            planes[0, y0:y0 + T, x0:x0 + T] += w * p
            planes[1, y0:y0 + T, x0:x0 + T] += w

    def reduce(self, planes_list, dst, group):
        import torch
        import torch.distributed as dist
        flat = torch.from_numpy(np.concatenate([p.reshape(-1) for p in planes_list]))
        dist.reduce(flat, dst=dst, op=dist.ReduceOp.SUM, group=group)
        out, o = [], 0
        flat = flat.numpy()
        for p in planes_list:
            out.append(flat[o:o + p.size].reshape(p.shape))
            o += p.size
        return out

    def finalize(self, planes):
        return planes[0] / planes[1]

    finalize_device = finalize

    def crop(self, tile, shift, tile_relative_size):
        from baryon_painter_b200.process_SLICS import get_tile
        return get_tile(np.asarray(tile), shift=shift, tile_relative_size=tile_relative_size)

    def send(self, plane, dst, group):
        import torch
        import torch.distributed as dist
        dist.send(torch.from_numpy(np.ascontiguousarray(plane, np.float64)), dst=dst, group=group)

    def recv(self, shape, src, group):
        import torch
        import torch.distributed as dist
        t = torch.empty(shape, dtype=torch.float64)
        dist.recv(t, src=src, group=group)
        return t.numpy()

    def new_map(self, resolution):
        return np.zeros((resolution, resolution))

    def zoom_accumulate(self, y_map, plane, scale, order):
        import scipy.ndimage
        plane = np.asarray(plane, np.float64)
        d = np.where(np.isnan(plane), 0.0, plane) * scale
        y_map += scipy.ndimage.zoom(d, zoom=y_map.shape[0] / plane.shape[0], order=order, mode="mirror")

    def to_host(self, t):
        return np.asarray(t)


# ---- synthetic SLICS-like inputs (block-constant log-normal fields: cheap to make at the reference's fixed
# plane sizes, 7745^2 delta maps and 12288^2 mass planes) ---------------------------------------------------
def synthetic_delta_file_content(seed, n=7745, block=32):
    """float32 (n, n): what a ``*delta.dat_bicubic_LOS*`` file holds (mean about -96 + 1/MASS_NORM)."""
    nb = (n + block - 1) // block
    base = np.random.default_rng(seed).lognormal(-0.5, 1.0, (nb, nb)).astype(np.float32)
    dens = np.repeat(np.repeat(base, block, axis=0), block, axis=1)[:n, :n]
    return (dens / np.float32(MASS_NORM) - np.float32(96)).astype(np.float32)


def delta_plane_from_file_content(content):
    """the rescaling of reference process_SLICS.py:187-189 (in float32, in place, on the transposed view)"""
    delta = np.array(content.T, dtype=np.float32, copy=True)
    delta += 96
    delta *= MASS_NORM
    return delta


def synthetic_massplane_file_content(seed, n=12288, block=32):
    base = np.random.default_rng(seed).lognormal(-0.5, 1.0, (n // block, n // block)).astype(np.float32)
    dens = np.repeat(np.repeat(base, block, axis=0), block, axis=1)
    return (dens / np.float32(MASS_NORM)).astype(np.float32)


def massplane_from_file_content(content):
    plane = np.array(content.T, dtype=np.float32, copy=True)
    plane *= MASS_NORM
    return plane
