"""GPU parity of the CGAN generator path (BASELINE.json configs[2]) against the oracle restatement.

The PainterGAN sources and the trained generator are not part of the reference checkout,
so this parity is UNPINNED: the oracle is a torch restatement of trained_models/README.md:116-128 + g_struc.pickle
and both sides share the same seeded synthetic weights.  Tolerances as for the CVAE: fp32 1e-4, 16-bit 1e-2."""
import os

import numpy as np
import pytest

from conftest import rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("fp16", 1e-2)])
def test_cgan_vs_oracle(precision, tol):
    import torch
    from oracle.cvae_oracle import CGANOracle
    from baryon_painter_b200 import arch, synthetic, transforms
    from baryon_painter_b200.painter import CGANPainter
    torch.set_num_threads(os.cpu_count())
    tile, nres = 128, 3
    layers = arch.fiducial_cgan_architecture(n_res_blocks=nres)
    sd = synthetic.synthetic_cgan_state_dict(layers, seed=5)
    orc = CGANOracle(layers, sd)
    stats = transforms.fiducial_stats()
    p = CGANPainter(device="cuda:0", precision=precision, max_batch=4, tile_size=tile, layers=layers, state_dict=sd)
    tiles = synthetic.synthetic_dm_tiles(6, tile, seed0=900)
    zs = np.array([0.0, 0.5, 1.0, 0.0, 0.5, 1.0])
    out = p.paint_batch(tiles, z=zs)                       # 6 tiles, max_batch 4: two host batches
    assert out.shape == (6, tile, tile) and out.dtype == np.float32
    for i in range(6):
        ref = orc.paint(tiles[i], float(zs[i]), stats)
        assert rel_l2(out[i], ref) <= tol, (i, rel_l2(out[i], ref))
    # network output before the inverse transform, single-tile entry point
    raw = p.paint(tiles[0], z=0.5, inverse_transform=False)
    ref = orc.paint(tiles[0], 0.5, stats, inverse_transform=False)
    assert raw.shape == (1, 1, tile, tile)
    assert rel_l2(raw, ref) <= (1e-5 if precision == "fp32" else 3e-3)
    with pytest.raises(ValueError, match="Shape mismatch between input and model"):
        p.paint(np.ones((32, 32), np.float32), z=0.0)


def test_cgan_fiducial_shape_vs_oracle():
    """BASELINE.json configs[2] shape: the full 9-block generator on 512x512 tiles (100.7 GFLOP per tile), fp16,
    z over {0, 0.5, 1}, against the oracle restatement: <= 1e-2 relative L2 per tile; deterministic (no latent)."""
    import torch
    from oracle.cvae_oracle import CGANOracle
    from baryon_painter_b200 import arch, synthetic, transforms
    from baryon_painter_b200.painter import CGANPainter
    torch.set_num_threads(os.cpu_count())
    layers = arch.fiducial_cgan_architecture()
    sd = synthetic.synthetic_cgan_state_dict(layers, seed=1)
    orc = CGANOracle(layers, sd)
    stats = transforms.fiducial_stats()
    p = CGANPainter(device="cuda:0", precision="fp16", max_batch=4, tile_size=512, layers=layers, state_dict=sd)
    tiles = synthetic.synthetic_dm_tiles(3, 512, seed0=30)
    zs = [0.0, 0.5, 1.0]
    a = p.paint_batch(tiles, z=zs)
    b = p.paint_batch(tiles, z=zs)
    assert a.shape == (3, 512, 512) and np.all(np.isfinite(a)) and np.array_equal(a, b)
    for i in range(3):
        ref = orc.paint(tiles[i], zs[i], stats)
        assert rel_l2(a[i], ref) <= 1e-2, (i, rel_l2(a[i], ref))


def test_cgan_host_pipeline_and_out_buffer():
    """CGAN host path: pipelined chunks, caller-supplied (page-locked) result buffer, sigma(z) per distinct redshift --
    same numbers whichever way the batch is cut."""
    import baryon_painter_b200 as bp
    from baryon_painter_b200 import synthetic
    from baryon_painter_b200.painter import CGANPainter
    tile, n = 64, 37
    g = CGANPainter.synthetic(tile_size=tile, device="cuda:0", precision="fp16", max_batch=64, n_res_blocks=2)
    tiles = synthetic.synthetic_dm_tiles(8, tile, seed0=3)
    tiles = np.ascontiguousarray(np.concatenate([tiles] * 5)[:n])
    zs = np.array([0.0, 0.5, 1.0])[np.arange(n) % 3]
    ref = g.paint_batch(tiles, z=zs)
    out = bp.pinned_empty(tiles.shape)
    got = g.paint_batch(tiles, z=zs, out=out)
    assert got is out and np.array_equal(out, ref)
    one = np.stack([g.paint(tiles[i], z=float(zs[i])) for i in (0, 5, 36)])
    assert np.array_equal(one, ref[[0, 5, 36]])
