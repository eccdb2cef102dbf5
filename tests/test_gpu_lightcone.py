"""Lightcone path on the GPU: device stitching (C ABI bp_stitch_*) against the reference golden, and the full
tile loop -- batched CUDA painting + device stitching -- against the oracle painter + numpy restatement."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, rel_l2

pytestmark = pytest.mark.gpu


def test_device_stitch_matches_reference_golden():
    from oracle import slics_oracle as so
    from baryon_painter_b200 import process_SLICS as ps
    from test_lightcone import _case
    g, args = _case()
    planes = ps.process_SLICS(so.StubPainter(), **args)            # default backend: device
    for i, p in enumerate(planes):
        assert p.shape == g[f"plane{i}"].shape and p.dtype == np.float64
        # float64 atomics: summation order differs from the host loop in the last bits only
        assert np.allclose(p, g[f"plane{i}"], rtol=1e-12, atol=0), i


class _FixedEps:
    """paint with eps = 0 (latent = z_mu) so the device and oracle runs see the same latents"""

    def __init__(self, painter):
        self.p = painter
        self.compute_device = painter.compute_device

    def paint_batch_device(self, tiles, z=0.0, out=None):
        import torch
        eps = torch.zeros((tiles.shape[0], *self.p.model.dim_z[1:]), dtype=torch.float32, device=tiles.device)
        return self.p.paint_batch_device(tiles, z=z, eps=eps, out=out)


class _OraclePainter:
    compute_device = None

    def __init__(self, tile, seed):
        from oracle.cvae_oracle import CVAEOracle
        from baryon_painter_b200 import arch, synthetic, transforms
        A = arch.fiducial_cvae_architecture(tile)
        self.o = CVAEOracle(A, synthetic.synthetic_cvae_state_dict(A, seed=seed))
        self.stats = transforms.fiducial_stats()
        self.lat = (tile // 32, tile // 32)

    def paint(self, input, z=0.0, transform=True, inverse_transform=True):
        return self.o.paint(np.asarray(input, np.float32), float(z), self.stats, eps=np.zeros((1, 1, *self.lat), np.float32))


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("fp16", 1e-2)])
def test_tile_loop_vs_oracle(precision, tol):
    from oracle import slics_oracle as so
    from baryon_painter_b200 import process_SLICS as ps
    from baryon_painter_b200.painter import CVAEPainter
    tile = 64
    rng = np.random.default_rng(5)
    import scipy.ndimage

    def field(n):     # smooth positive density (a cubic zoom of white noise overshoots below zero -> log of negatives)
        g = scipy.ndimage.gaussian_filter(rng.standard_normal((n, n)), 4.0, mode="wrap")
        return np.exp(g / g.std() - 0.5).astype(np.float32)

    planes_in = {1: field(400), 2: field(500)}
    args = dict(tile_size=100.0, n_pixel_tile=tile, LOS=1, z_SLICS=[0.2, 0.5], delta_size=[150.0, 230.0],
                delta_path=None, massplane_path=None, shifts_path=np.zeros((2, 2)), z_slice=[0.21, 0.52], verbose=False,
                plane_source=lambda i, kind: planes_in[i + 1])
    dev = ps.process_SLICS(_FixedEps(CVAEPainter.synthetic(tile_size=tile, seed=4, precision=precision, max_batch=16)),
                           batch=16, **args)
    ref = ps.process_SLICS(_OraclePainter(tile, 4), backend=so.NumpyBackend(), **args)
    assert [p.shape for p in dev] == [(96, 96), (147, 147)]
    for d, r in zip(dev, ref):
        assert rel_l2(d, r) <= tol
    # projected map
    cosmo = ps.FlatLCDM()
    yd = ps.create_y_map(dev, [0.21, 0.52], 64, 10.0, cosmo, verbose=False)
    yr = ps.create_y_map(ref, [0.21, 0.52], 64, 10.0, cosmo, verbose=False)
    assert rel_l2(yd, yr) <= tol
    # the same projection on the device (quintic spline, as scripts/create_lightcone.py asks for) against host scipy
    for order in (3, 5):
        y_host = ps.create_y_map(dev, [0.21, 0.52], 64, 10.0, cosmo, order=order, verbose=False)
        y_dev = ps.create_y_map(dev, [0.21, 0.52], 64, 10.0, cosmo, order=order, verbose=False, backend=ps.DeviceBackend("cuda:0"))
        assert rel_l2(y_dev, y_host) <= 1e-10


def test_paint_lightcone_device_equals_two_step():
    """paint_lightcone (planes stay on the device, projected there, quintic spline) against
    create_y_map(process_SLICS(...)) of the same painter; and a second run re-using the backend with fresh numpy
    planes of the same shape (a stale uploaded plane must never be matched by address)."""
    from baryon_painter_b200 import process_SLICS as ps
    from baryon_painter_b200.painter import CVAEPainter
    import scipy.ndimage
    tile = 64
    rng = np.random.default_rng(11)

    def field(n):
        g = scipy.ndimage.gaussian_filter(rng.standard_normal((n, n)), 4.0, mode="wrap")
        return np.exp(g / g.std() - 0.5).astype(np.float32)

    painter = _FixedEps(CVAEPainter.synthetic(tile_size=tile, seed=4, precision="fp16", max_batch=16))
    be = ps.DeviceBackend("cuda:0")
    cosmo = ps.FlatLCDM()
    for trial in range(2):
        planes_in = [field(300), field(400), field(400)]
        args = dict(tile_size=100.0, n_pixel_tile=tile, LOS=1, z_SLICS=[0.1, 0.2, 0.5], delta_size=[60.0, 150.0, 230.0],
                    delta_path=None, massplane_path=None, shifts_path=rng.random((3, 2)), z_slice=[0.11, 0.21, 0.52],
                    plane_source=lambda i, kind: np.array(planes_in[i]))      # a fresh array on every call
        planes = ps.process_SLICS(painter, batch=16, verbose=False, backend=be, **args)
        ref = ps.create_y_map(planes, args["z_SLICS"], 80, 10.0, cosmo, order=5, verbose=False)
        y = ps.paint_lightcone(painter, resolution=80, map_size=10.0, cosmo=cosmo, order=5, verbose=False, batch=16,
                               backend=be, **args)
        assert y.shape == (80, 80) and rel_l2(y, ref) <= 1e-9, (trial, rel_l2(y, ref))


def _sharded_case(rng_seed=21):
    import scipy.ndimage
    rng = np.random.default_rng(rng_seed)

    def field(n):
        g = scipy.ndimage.gaussian_filter(rng.standard_normal((n, n)), 4.0, mode="wrap")
        return np.exp(g / g.std() - 0.5).astype(np.float32)

    planes_in = [field(300), field(400), field(400), field(400), field(400)]
    args = dict(tile_size=100.0, n_pixel_tile=64, LOS=1, z_SLICS=[0.1, 0.2, 0.3, 0.4, 0.5],
                delta_size=[60.0, 150.0, 150.0, 230.0, 230.0], delta_path=None, massplane_path=None,
                shifts_path=rng.random((5, 2)), z_slice=[0.11, 0.21, 0.31, 0.41, 0.52])
    return planes_in, args


def _device_rank_main(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    sys.path.insert(0, os.path.join(root, "tests"))
    import torch.distributed as dist
    from baryon_painter_b200 import process_SLICS as ps
    from baryon_painter_b200.painter import CVAEPainter
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    planes_in, args = _sharded_case()
    args["plane_source"] = lambda i, kind: np.array(planes_in[i])           # a fresh numpy array on every call
    painter = _FixedEps(CVAEPainter.synthetic(tile_size=64, seed=4, precision="fp16", max_batch=16))
    be = ps.DeviceBackend("cuda:0")
    planes = ps.process_SLICS(painter, batch=16, verbose=False, backend=be, rank=rank, world_size=world, **args)
    y = ps.paint_lightcone(painter, resolution=80, map_size=10.0, cosmo=ps.FlatLCDM(), order=5, verbose=False, batch=16,
                           backend=be, rank=rank, world_size=world, **args)
    if rank == 0:
        q.put(([np.asarray(p) for p in planes], np.asarray(y)))
    else:
        assert planes is None and y is None
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_device_backend_two_ranks():
    """world_size 2 with the DEVICE backend and numpy planes: whole planes dealt to two ranks (both on
    cuda:0, gloo group, exchanges staged through the host), a rank that skips planes, the backend re-used across
    process_SLICS and paint_lightcone.  Equals the single-process run."""
    import socket
    import torch.multiprocessing as mp
    from baryon_painter_b200 import process_SLICS as ps
    from baryon_painter_b200.painter import CVAEPainter
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_device_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    planes, y = q.get(timeout=900)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    planes_in, args = _sharded_case()
    args["plane_source"] = lambda i, kind: np.array(planes_in[i])
    painter = _FixedEps(CVAEPainter.synthetic(tile_size=64, seed=4, precision="fp16", max_batch=16))
    ref = ps.process_SLICS(painter, batch=16, verbose=False, **args)
    assert len(planes) == len(ref) == 5
    for a, b in zip(planes, ref):
        assert a.shape == b.shape and np.allclose(a, b, rtol=1e-10, atol=0)
    y_ref = ps.create_y_map(ref, args["z_SLICS"], 80, 10.0, ps.FlatLCDM(), order=5, verbose=False)
    assert rel_l2(y, y_ref) <= 1e-9
