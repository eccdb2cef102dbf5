"""Lightcone tiling / stitching host logic (CPU): known answers generated from the reference's own
process_SLICS.py (tests/golden/tiling.json, slics_small.npz; oracle/make_golden*.py), the reference's
tests/test_SLICS_tiling.py::test_generate_tiling cases, and the rank-sharded assembly over gloo."""
import json
import os
import socket
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

from baryon_painter_b200 import process_SLICS as ps


def test_generate_tiling_reference_cases():
    # reference tests/test_SLICS_tiling.py:72-87
    for n_plane, n_tile, ov, n in [(512, 256, 0.0, 2), (512, 250, 0.0, 3), (512, 256, 0.5, 3), (512, 128, 0.0, 4)]:
        origins, slices = ps.generate_tiling(n_plane, n_tile, ov)
        assert len(origins) == n
    origins, slices = ps.generate_tiling(512, 32, 0.33)
    cover = np.zeros((512, 512), int)
    for row in slices:
        for sl in row:
            cover[sl] += 1
    assert cover.min() >= 1


def test_tiling_known_answers():
    g = json.load(open(os.path.join(GOLDEN, "tiling.json")))
    for c in g["generate_tiling"]:
        origins, slices = ps.generate_tiling(c["n_pixel_plane"], c["n_pixel_tile"], c["min_tile_overlap"])
        assert np.array_equal(origins, np.array(c["origins"]))
        assert [r[0][0].start for r in slices] == c["starts"]
        assert all(sl[0].stop - sl[0].start == c["n_pixel_tile"] for r in slices for sl in r)
    w = ps.make_weight_map((512, 512), falloff=0.05, sigma=0.5)
    gw = g["weight_512"]
    assert w[0, 0] == gw["corner"] and w[0, 256] == gw["edge"] and w[256, 256] == gw["centre"]
    assert np.array_equal(w[:30, 256], np.array(gw["row0"])) and w.sum() == pytest.approx(gw["sum"], rel=1e-13)
    w64 = ps.make_weight_map((64, 64), falloff=0.1, sigma=1)
    assert np.array_equal(w64[:8, 32], np.array(g["weight_64"]["row0"]))
    assert w64.sum() == pytest.approx(g["weight_64"]["sum"], rel=1e-13)
    m = np.random.default_rng(g["get_tile_seed"]).standard_normal(tuple(g["get_tile_shape"])).astype(np.float32)
    for c in g["get_tile"]:
        t = ps.get_tile(m, c["shift"], c["tile_relative_size"], c["expansion_factor"])
        assert list(t.shape) == c["shape"]
        assert [float(t[0, 0]), float(t[0, -1]), float(t[-1, 0]), float(t[-1, -1])] == c["corner"]
        assert float(t.astype(np.float64).sum()) == pytest.approx(c["sum"], rel=1e-12)
    with pytest.raises(ValueError):
        ps.get_tile(m, (0, 0), 0.5, 0.5)


def _case():
    from oracle import slics_oracle as so
    g = np.load(os.path.join(GOLDEN, "slics_small.npz"))
    seeds, dsize = g["case_seeds"], g["case_delta_size"]
    tile_size = float(g["case_tile_size"])
    cache = {}

    def plane_source(i, kind):
        if i not in cache:
            if kind == "mass":
                cache[i] = so.massplane_from_file_content(so.synthetic_massplane_file_content(int(seeds[i])))
            else:
                cache[i] = so.delta_plane_from_file_content(so.synthetic_delta_file_content(int(seeds[i])))
        return cache[i]

    args = dict(tile_size=tile_size, n_pixel_tile=int(g["case_n_pixel_tile"]), LOS=int(g["case_LOS"]),
                z_SLICS=list(g["case_z_SLICS"]), delta_size=list(dsize), delta_path=None, massplane_path=None,
                shifts_path=g["case_shifts"], z_slice=list(g["case_z_slice"]), verbose=False, plane_source=plane_source)
    return g, args


def test_process_slics_matches_reference_golden():
    """same tiles, same stitching arithmetic as the reference (numpy backend = oracle restatement): bit-identical"""
    from oracle import slics_oracle as so
    g, args = _case()
    planes = ps.process_SLICS(so.StubPainter(), backend=so.NumpyBackend(), **args)
    assert len(planes) == 3
    for i, p in enumerate(planes):
        assert p.shape == g[f"plane{i}"].shape and p.dtype == np.float64
        assert np.array_equal(p, g[f"plane{i}"]), i
    with pytest.raises(ValueError, match="Shapes of z_SLICS and z_slice"):
        ps.process_SLICS(so.StubPainter(), backend=so.NumpyBackend(), **{**args, "z_slice": [0.0]})


def _rank_main(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from oracle import slics_oracle as so
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    g, args = _case()
    planes = ps.process_SLICS(so.StubPainter(), backend=so.NumpyBackend(), rank=rank, world_size=world, **args)
    if rank == 0:
        q.put([np.asarray(p) for p in planes])
    else:
        assert planes is None
    dist.barrier()
    dist.destroy_process_group()


def test_plan_planes_lpt():
    """planes are dealt whole, by cost, longest first (a SLICS line of sight: 1, 1, 4, 9, ..., 144 tiles per plane)"""
    costs = [1, 1, 4, 9, 9, 16, 25, 36, 49, 64, 81, 100, 121, 121, 144]
    assert ps.plan_planes(costs, 1) == [0] * 15
    for w in (2, 4, 8):
        owner = ps.plan_planes(costs, w)
        load = [sum(c for c, o in zip(costs, owner) if o == r) for r in range(w)]
        assert sum(load) == 781 and max(load) <= max(144, 781 / w * 1.08)       # LPT: within a few % of the bound
        assert owner == ps.plan_planes(costs, w)                                   # deterministic
    assert max(sum(c for c, o in zip(costs, ps.plan_planes(costs, 8)) if o == r) for r in range(8)) == 144


def _lightcone_rank_main(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from oracle import slics_oracle as so
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    g, args = _case()
    args = {k: v for k, v in args.items() if k != "verbose"}
    res = ps.paint_lightcone(so.StubPainter(), resolution=48, map_size=10.0, cosmo=ps.FlatLCDM(), order=3, verbose=False,
                             backend=so.NumpyBackend(), rank=rank, world_size=world, drop_planes=1, **args)
    if rank == 0:
        q.put([np.asarray(m) for m in res])
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


def test_paint_lightcone_sharded_gloo():
    """the fused path: every rank paints its planes and projects them into its partial y map; ONE reduce of the map.
    Equals create_y_map(process_SLICS(...)) of the unsharded run (also with the first plane dropped)."""
    import torch.multiprocessing as mp
    from oracle import slics_oracle as so
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_lightcone_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    y_map, y_drop = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    g, args = _case()
    planes = ps.process_SLICS(so.StubPainter(), backend=so.NumpyBackend(), **args)
    z = args["z_SLICS"]
    ref = ps.create_y_map(planes, z, 48, 10.0, ps.FlatLCDM(), order=3, verbose=False)
    ref_drop = ps.create_y_map(planes[1:], z[1:], 48, 10.0, ps.FlatLCDM(), order=3, verbose=False)
    assert np.allclose(y_map, ref, rtol=1e-12, atol=0) and np.allclose(y_drop, ref_drop, rtol=1e-12, atol=0)
    # single process, same call
    one = ps.paint_lightcone(so.StubPainter(), resolution=48, map_size=10.0, cosmo=ps.FlatLCDM(), order=3, verbose=False,
                             backend=so.NumpyBackend(), **{k: v for k, v in args.items() if k != "verbose"})
    assert np.allclose(one, ref, rtol=1e-12, atol=0)


def test_process_slics_sharded_gloo():
    """world_size 2: whole planes dealt by cost, each rank paints only its own, finished planes sent to rank 0"""
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    planes = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    g, _ = _case()
    for i, p in enumerate(planes):
        assert np.allclose(p, g[f"plane{i}"], rtol=1e-13, atol=0), i


def test_y_map_projection():
    """create_y_map: linear in the planes, NaN -> 0, flat-LCDM distances of the SLICS cosmology"""
    cosmo = ps.FlatLCDM()
    assert abs(float(cosmo.comoving_distance(1 / 2.0)) * cosmo.h - 2309.0) < 25      # chi(z=1) ~ 2.3 Gpc/h
    assert abs(float(cosmo.scale_factor_of_chi(cosmo.comoving_distance(0.4))) - 0.4) < 1e-4
    rng = np.random.default_rng(0)
    a, b = rng.random((40, 40)), rng.random((56, 56))
    z = [0.04, 0.13]
    y1 = ps.create_y_map([a, b], z, 32, 10.0, cosmo, order=3, verbose=False)
    y2 = ps.create_y_map([2 * a, 2 * b], z, 32, 10.0, cosmo, order=3, verbose=False)
    assert y1.shape == (32, 32) and np.allclose(y2, 2 * y1, rtol=1e-12)
    a_nan = a.copy()
    a_nan[3, 4] = np.nan
    a0 = a.copy()
    a0[3, 4] = 0
    assert np.array_equal(ps.create_y_map([a_nan, b], z, 32, 10.0, cosmo, verbose=False),
                          ps.create_y_map([a0, b], z, 32, 10.0, cosmo, verbose=False))
