"""Tile extraction / cubic-spline zoom: the numpy restatement is pinned against scipy
(CPU), the CUDA kernels against scipy on the GPU box."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _field(n, seed):
    import scipy.ndimage
    g = scipy.ndimage.gaussian_filter(np.random.default_rng(seed).standard_normal((n, n)), 2.0, mode="wrap")
    return np.exp(g / g.std() - 0.5).astype(np.float32)


@pytest.mark.parametrize("mode", ["reflect", "mirror"])
@pytest.mark.parametrize("side,out", [(37, 64), (100, 64), (64, 64), (5, 16)])
def test_zoom_oracle_matches_scipy(mode, side, out):
    import scipy.ndimage
    from oracle import zoom_oracle
    tile = _field(side, side)
    ref = scipy.ndimage.zoom(tile, zoom=out / side, mode=mode)
    assert ref.shape == (out, out)
    got = zoom_oracle.zoom(tile, out, mode)
    np.testing.assert_allclose(got, ref, rtol=2e-6, atol=2e-6)


@pytest.mark.parametrize("side,out", [(40, 29), (23, 64)])
def test_zoom_oracle_order5_matches_scipy(side, out):
    """create_y_map's call: float64 plane, order 5, mode mirror (reference process_SLICS.py:63)."""
    import scipy.ndimage
    from oracle import zoom_oracle
    plane = _field(side, 7).astype(np.float64)
    ref = scipy.ndimage.zoom(plane, zoom=out / side, order=5, mode="mirror")
    assert ref.shape == (out, out)
    np.testing.assert_allclose(zoom_oracle.zoom(plane, out, "mirror", order=5), ref, rtol=1e-11, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("order", [3, 5])
def test_gpu_zoom_accumulate_matches_scipy(order):
    import scipy.ndimage
    import torch
    from baryon_painter_b200 import _lib
    dev = torch.device("cuda:0")
    y_ref = np.zeros((96, 96))
    y_dev = torch.zeros((96, 96), dtype=torch.float64, device=dev)
    for k, side in enumerate((61, 150)):
        plane = _field(side, 10 + k).astype(np.float64)
        plane[3, 5] = np.nan                                   # create_y_map zeroes NaNs first
        scale = 0.5 + k
        y_ref += scale * scipy.ndimage.zoom(np.where(np.isnan(plane), 0.0, plane), zoom=96 / side, order=order, mode="mirror")
        d = torch.from_numpy(plane).to(dev)
        _lib.zoom_accumulate(0, d.data_ptr(), side, 96, order, "mirror", scale, y_dev.data_ptr(),
                             torch.cuda.current_stream(dev).cuda_stream)
        torch.cuda.synchronize()
    np.testing.assert_allclose(y_dev.cpu().numpy(), y_ref, rtol=1e-10, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("order,side,out", [(5, 1100, 300), (3, 1100, 300), (5, 641, 97)])
def test_gpu_zoom_accumulate_long_lines_match_scipy(order, side, out):
    """planes longer than 512 samples take the segment-parallel sweeps (both recursions of every pole fused in shared
    memory, two poles for the quintic spline create_lightcone asks for): same map as scipy in float64"""
    import scipy.ndimage
    import torch
    from baryon_painter_b200 import _lib
    dev = torch.device("cuda:0")
    plane = _field(side, 31).astype(np.float64)
    plane[7, 9] = np.nan
    y_ref = 1.75 * scipy.ndimage.zoom(np.where(np.isnan(plane), 0.0, plane), zoom=out / side, order=order, mode="mirror")
    y_dev = torch.zeros((out, out), dtype=torch.float64, device=dev)
    d = torch.from_numpy(plane).to(dev)
    _lib.zoom_accumulate(0, d.data_ptr(), side, out, order, "mirror", 1.75, y_dev.data_ptr(),
                         torch.cuda.current_stream(dev).cuda_stream)
    torch.cuda.synchronize()
    np.testing.assert_allclose(y_dev.cpu().numpy(), y_ref, rtol=1e-10, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["reflect", "mirror"])
def test_gpu_zoom_tiles_match_scipy(mode):
    import scipy.ndimage
    import torch
    from baryon_painter_b200 import _lib
    from baryon_painter_b200 import process_SLICS as ps
    plane = _field(300, 3)
    dev = torch.device("cuda:0")
    d_plane = torch.from_numpy(plane).to(dev)
    rel = 0.37
    shifts = [(0.0, 0.0), (0.9, 0.15), (0.5, 0.95)]                 # the last two wrap around the plane edge
    side = int(plane.shape[0] * rel)
    org = torch.tensor([[int(plane.shape[0] * a), int(plane.shape[1] * b)] for a, b in shifts], dtype=torch.int32, device=dev)
    out = torch.empty((len(shifts), 128, 128), dtype=torch.float32, device=dev)
    _lib.zoom_tiles(0, d_plane.data_ptr(), plane.shape[0], plane.shape[1], org.data_ptr(), side, len(shifts), 128, mode,
                    out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
    got = out.cpu().numpy()
    for t, sh in enumerate(shifts):
        tile = ps.get_tile(plane, shift=sh, tile_relative_size=rel)
        ref = scipy.ndimage.zoom(tile, zoom=128 / tile.shape[0], mode=mode)
        # float64 spline arithmetic on both sides, float32 result: a few ulp
        np.testing.assert_allclose(got[t], ref, rtol=3e-6, atol=3e-6)


@pytest.mark.gpu
def test_gpu_plane_prepare_bit_identical_to_numpy():
    """Plane preprocessing on the device (reference process_SLICS.py:157-159, :187-189): transpose, `+= 96`,
    `*= MASS_NORM` -- the same two float32 roundings as numpy, so bit-identical."""
    from baryon_painter_b200 import process_SLICS as ps
    be = ps.DeviceBackend("cuda:0")
    raw = (np.random.default_rng(3).standard_normal((301, 257)) * 40).astype(np.float32)
    delta = raw.copy().T
    delta += 96
    delta *= ps.MASS_NORM
    got = be.prepare_plane(raw, 96.0, ps.MASS_NORM).cpu().numpy()
    assert got.shape == (257, 301) and np.array_equal(got, delta)
    mass = raw.T * np.float32(ps.MASS_NORM)
    assert np.array_equal(be.prepare_plane(raw, 0.0, np.float32(ps.MASS_NORM)).cpu().numpy(), mass)
    # and the prepared plane feeds the tile extraction without leaving the device
    tiles = be.extract_tiles(be.prepare_plane(raw, 96.0, ps.MASS_NORM), [(0.1, 0.8)], 0.4, 64, "reflect")
    import scipy.ndimage
    t = ps.get_tile(delta, shift=(0.1, 0.8), tile_relative_size=0.4)
    ref = scipy.ndimage.zoom(t, zoom=64 / t.shape[0], mode="reflect")
    np.testing.assert_allclose(tiles[0].cpu().numpy(), ref, rtol=3e-6, atol=3e-6 * np.abs(ref).max())


@pytest.mark.gpu
@pytest.mark.parametrize("mode,side", [("reflect", 1357), ("mirror", 777), ("reflect", 513)])
def test_gpu_zoom_long_lines_match_scipy(mode, side):
    """Crops longer than 512 samples run the segment-parallel prefilter (csrc/bp_zoom.cu: truncated boundary sums,
    256-sample segments warmed up over the recursion's horizon); same scipy parity as the sequential kernels."""
    import scipy.ndimage
    import torch
    from baryon_painter_b200 import _lib
    from baryon_painter_b200 import process_SLICS as ps
    n = 1600
    plane = np.tile(_field(400, 5), (4, 4))
    dev = torch.device("cuda:0")
    d_plane = torch.from_numpy(plane).to(dev)
    shifts = [(0.0, 0.0), (0.7, 0.9)]
    rel = side / n
    assert int(n * rel) == side
    org = torch.tensor([[int(n * a), int(n * b)] for a, b in shifts], dtype=torch.int32, device=dev)
    out = torch.empty((2, 256, 256), dtype=torch.float32, device=dev)
    _lib.zoom_tiles(0, d_plane.data_ptr(), n, n, org.data_ptr(), side, 2, 256, mode, out.data_ptr(),
                    torch.cuda.current_stream(dev).cuda_stream)
    got = out.cpu().numpy()
    for t, sh in enumerate(shifts):
        tile = ps.get_tile(plane, shift=sh, tile_relative_size=rel)
        ref = scipy.ndimage.zoom(tile, zoom=256 / tile.shape[0], mode=mode)
        np.testing.assert_allclose(got[t], ref, rtol=3e-6, atol=3e-6)
    # quintic projection of a long plane (two poles: four out-of-place sweeps per axis)
    p64 = plane[:side, :side].astype(np.float64)
    y = torch.zeros((300, 300), dtype=torch.float64, device=dev)
    d64 = torch.from_numpy(p64).to(dev)
    _lib.zoom_accumulate(0, d64.data_ptr(), side, 300, 5, "mirror", 2.5, y.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
    ref = 2.5 * scipy.ndimage.zoom(p64, zoom=300 / side, order=5, mode="mirror")
    np.testing.assert_allclose(y.cpu().numpy(), ref, rtol=1e-10, atol=1e-12)
