"""GPU parity of the CVAE paint path (through the C ABI) against the committed golden vectors
(generated from the unmodified reference by oracle/make_golden.py) and against the oracle
restatement on fresh seeded inputs.

Tolerances (BASELINE.json north_star): fp32 path <= 1e-4 relative L2 per tile; 16-bit path <= 1e-2
relative L2 per tile and auto/cross power spectra within 1 % (asserted exactly so; see TOL below).
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN, rel_l2

pytestmark = pytest.mark.gpu

# painted-tile tolerances, exactly as north_star states them: fp32 path <= 1e-4, 16-bit path <= 1e-2 relative L2 per
# tile (and auto / cross power spectra within 1 %, asserted below at 128^2 and at 512^2).  The shipped 16-bit path has
# fp16 operands (fp32 accumulation in TMEM; same tensor rate as bf16, 8x finer rounding).  bf16 operands stay
# selectable but are NOT the shipped 16-bit path: through the inverse transform p = (exp(4x) - 1) sigma an absolute
# error dx of the network output becomes a relative error 4 dx, and bf16's 8-bit mantissa through 25 layers gives
# dx ~ 5e-3: the network output x_mu meets 1e-2, painted tiles do not, and the suite asserts no looser painted
# bound for it (DESIGN.md "Precision").
# "fp32" is the fp32-accurate tensor-core path (split-precision fp16 operands, fp32 accumulation in TMEM); "fp32-ffma"
# the scalar FFMA kernels kept as an on-device cross-check.  Both carry north_star's fp32 bound.
TOL = {"fp32": 1e-4, "fp32-ffma": 1e-4, "fp16": 1e-2}
TOL_XMU = {"fp32": 1e-5, "fp32-ffma": 1e-5, "fp16": 2e-3, "bf16": 1e-2}
# layer-boundary tensors: fp32 accumulates ~1e-6 per layer; fp16 ~ 4e-4, bf16 ~ 3e-3 per layer
TOL_LAYER = {"fp32": 2e-5, "fp32-ffma": 2e-5, "fp16": 3e-3, "bf16": 2e-2}
# (z_mu, z_log_var) of the t64 fixture are 2x2 maps with one or two non-zero entries after the ReLU, so a single
# element's rounding through the four 16-bit prior layers is the whole norm
TOL_PRIOR = {"fp32": 2e-5, "fp32-ffma": 2e-5, "fp16": 3e-3, "bf16": 5e-2}
PRECISIONS = ["fp32", "fp16"]
ALL_FORMATS = ["fp32", "fp32-ffma", "fp16", "bf16"]


def _painter(tile, seed, precision, max_batch=8):
    from baryon_painter_b200.painter import CVAEPainter
    return CVAEPainter.synthetic(tile_size=tile, seed=seed, precision=precision, max_batch=max_batch)


def _tiles(g, n):
    from baryon_painter_b200 import synthetic
    return synthetic.synthetic_dm_tiles(n, int(g["tile_size"]), seed0=int(g["tiles_seed0"]))


@pytest.mark.parametrize("precision", ALL_FORMATS)
def test_layer_boundaries_t64(precision):
    from baryon_painter_b200 import arch
    g = np.load(os.path.join(GOLDEN, "cvae_t64_layers.npz"))
    p = _painter(64, int(g["seed"]), precision)
    p.model.net.set_debug(True)
    tiles = _tiles(g, 1)
    out = p.paint(tiles[0], z=float(g["z"][0]), eps=g["eps"][0])
    if precision in TOL:
        assert rel_l2(out, g["painted_E"][0]) <= TOL[precision]
    stack_ids = {"prior_network": 0, "p_z_in": 1, "p_y_z_in": 2, "p_mu_out": 3}
    worst = 0.0
    for name, sid in stack_ids.items():
        specs = p.model.stacks[name]
        # golden taps are keyed by the Sequential index of the last module of each conv group
        keys = sorted((k for k in g.files if k.startswith("tap:" + name + ".")), key=lambda s: int(s.split(".")[-1]))
        # residual blocks contribute one tap (after the block) for two convolutions
        li = -1
        for k in keys:
            li += 1
            while specs[li].res == arch.RES_OPEN:
                li += 1
            s = specs[li]
            ref = g[k]
            got = p.model.net.read_activation(sid, li, (1, *ref.shape))[0]
            if name == "p_mu_out" and li == len(specs) - 1:
                # the inverse transform is fused into the last layer's epilogue
                ref = p.inverse_transform(ref[None], field="pressure", z=float(g["z"][0]))[None]
            e = rel_l2(got, ref)
            worst = max(worst, e)
            assert e <= TOL_LAYER[precision], (k, li, e)
    mu, lv = p.model.net.cvae_read_prior(1)
    assert rel_l2(mu[0], g["z_mu"][0][0]) <= TOL_PRIOR[precision]
    assert rel_l2(lv[0], g["z_log_var"][0][0]) <= TOL_PRIOR[precision] or np.abs(g["z_log_var"]).max() < 1e-6
    print("worst layer rel-L2", precision, worst)


@pytest.mark.parametrize("precision", PRECISIONS + ["fp32-ffma"])
def test_golden_t128(precision):
    g = np.load(os.path.join(GOLDEN, "cvae_t128.npz"))
    p = _painter(128, int(g["seed"]), precision)
    tiles = _tiles(g, 4)
    zs = g["z"]
    # batched, both latent modes
    out_e = p.paint_batch(tiles, z=zs, eps=g["eps"])
    out_l = p.paint_batch(tiles, z=zs, latents=g["eps"])
    for i in range(4):
        assert rel_l2(out_e[i], g["painted_E"][i]) <= TOL[precision], ("E", i)
        assert rel_l2(out_l[i], g["painted_L"][i]) <= TOL[precision], ("L", i)
    # one tile at a time through paint(), the reference entry point
    for i in range(4):
        o = p.paint(tiles[i], z=float(zs[i]), eps=g["eps"][i])
        assert o.shape == (128, 128) and o.dtype == np.float32
        assert rel_l2(o, g["painted_E"][i]) <= TOL[precision]


@pytest.mark.parametrize("precision", PRECISIONS)
def test_golden_t512(precision):
    g = np.load(os.path.join(GOLDEN, "cvae_t512.npz"))
    p = _painter(512, int(g["seed"]), precision)
    tiles = _tiles(g, 2)
    o_e = p.paint(tiles[0], z=float(g["z"][0]), eps=g["eps"][0])
    o_l = p.paint(tiles[1], z=float(g["z"][1]), latent=g["eps"][1])
    assert rel_l2(o_e, g["painted_E"][0]) <= TOL[precision]
    assert rel_l2(o_l, g["painted_L"][0]) <= TOL[precision]


@pytest.mark.parametrize("precision", ALL_FORMATS)
def test_vs_oracle_fresh_inputs(precision):
    """same seeded weights/tiles/latents through the oracle (CPU) and the CUDA path; ragged batch
    (5 tiles with chunking) and redshifts outside the stats table (clamped)."""
    import torch
    from oracle.cvae_oracle import CVAEOracle, pseudo_Pofk
    from baryon_painter_b200 import arch, synthetic, transforms
    torch.set_num_threads(os.cpu_count())
    tile = 128
    A = arch.fiducial_cvae_architecture(tile)
    sd = synthetic.synthetic_cvae_state_dict(A, seed=11)
    stats = transforms.fiducial_stats()
    orc = CVAEOracle(A, sd)
    p = _painter(tile, 11, precision, max_batch=3)          # forces 2 host batches for 5 tiles
    tiles = synthetic.synthetic_dm_tiles(5, tile, seed0=500)
    zs = np.array([-0.1, 0.06, 0.5, 1.9, 3.0])
    eps = synthetic.synthetic_latents(5, (tile // 32, tile // 32), seed=9)
    ref = orc.paint_batch(tiles, zs, stats, eps=eps)
    out = p.paint_batch(tiles, z=zs, eps=eps)
    for i in range(5):
        assert np.isfinite(out[i]).all()
        if precision in TOL:
            assert rel_l2(out[i], ref[i]) <= TOL[precision], i
    # network output before the inverse transform
    xm = p.paint_batch(tiles, z=zs, eps=eps, inverse_transform=False)
    for i in range(5):
        ref_x = orc.paint(tiles[i], float(zs[i]), stats, eps=eps[i:i + 1], inverse_transform=False)
        assert rel_l2(xm[i], ref_x) <= TOL_XMU[precision], i
    if precision == "bf16":
        return
    # power spectra (north_star: within 1 % for the 16-bit path)
    for i in range(5):
        assert_spectra_within_1pct(tiles[i], out[i], ref[i])


def assert_spectra_within_1pct(dm, out, ref):
    """north_star: pressure auto-power and DM-pressure cross-power spectrum each within 1 % of the reference's
    (estimator: reference validation_plotting.py:120-121 settings, restated in oracle.pseudo_Pofk)."""
    from oracle.cvae_oracle import pseudo_Pofk
    k, pa_ref = pseudo_Pofk(ref, ref)
    _, pa = pseudo_Pofk(out, out)
    _, pdm = pseudo_Pofk(dm, dm)
    _, px_ref = pseudo_Pofk(dm, ref)
    _, px = pseudo_Pofk(dm, out)
    assert np.max(np.abs(pa / pa_ref - 1)) <= 0.01
    # cross spectrum: relative where it is significantly non-zero, and everywhere within 1 % of the
    # amplitude sqrt(P_dm P_p) (bins where the cross-correlation changes sign have no relative error)
    amp = np.sqrt(pdm * pa_ref)
    assert np.max(np.abs(px - px_ref) / amp) <= 0.01
    big = np.abs(px_ref) > 0.1 * amp
    assert not big.any() or np.max(np.abs(px[big] / px_ref[big] - 1)) <= 0.01


def test_benchmark_shape_chunk_vs_oracle():
    """The configuration bench.py times: fiducial 512^2 network, fp16, many tiles in ONE plan chunk (64 here: the
    same kernels, tilings and chunked launch path as the 256-tile step; bench.py repeats this check on its own
    256-tile batch and prints it as `parity`).  8 tiles spread over the chunk are compared with the oracle:
    <= 1e-2 relative L2 per tile and auto / cross spectra within 1 % (north_star), on mixed redshifts."""
    import torch
    from oracle.cvae_oracle import CVAEOracle
    from baryon_painter_b200 import arch, synthetic, transforms
    torch.set_num_threads(os.cpu_count())
    tile, n = 512, 64
    A = arch.fiducial_cvae_architecture(tile)
    sd = synthetic.synthetic_cvae_state_dict(A, seed=0)
    stats = transforms.fiducial_stats()
    orc = CVAEOracle(A, sd)
    p = _painter(tile, 0, "fp16", max_batch=n)
    assert p.model.net.chunk >= n                      # one chunk
    base = synthetic.synthetic_dm_tiles(8, tile, seed0=7000)
    tiles = np.ascontiguousarray(np.concatenate([base] * (n // 8)) * np.linspace(0.8, 1.25, n, dtype=np.float32)[:, None, None])
    eps = synthetic.synthetic_latents(n, (tile // 32, tile // 32), seed=3)
    zs = np.array([0.0, 0.25, 0.5, 1.0])[np.arange(n) % 4]
    out = p.paint_batch(tiles, z=zs, eps=eps)
    assert np.isfinite(out).all()
    worst = 0.0
    for i in (0, 9, 18, 27, 36, 45, 54, 63):
        ref = orc.paint(tiles[i], float(zs[i]), stats, eps=eps[i:i + 1])
        e = rel_l2(out[i], ref)
        worst = max(worst, e)
        assert e <= TOL["fp16"], (i, e)
        assert_spectra_within_1pct(tiles[i], out[i], ref)
    print("worst rel-L2 over 8 of 64 tiles (512^2, fp16, one chunk): %.2e" % worst)


def test_api_semantics():
    p = _painter(64, 3, "fp32")
    from baryon_painter_b200 import synthetic
    t = synthetic.synthetic_dm_tiles(1, 64, seed0=1)[0]
    with pytest.raises(ValueError, match="Shape mismatch between input and model"):
        p.paint(np.ones((32, 32), np.float32), z=0.0)
    with pytest.raises(ValueError, match="Shape mismatch between input and model"):
        p.paint(t, z=0.0, transform=False)                 # reference: (H,W) without atleast_3d fails
    eps = np.zeros((1, 2, 2), np.float32)
    full = p.paint(t, z=0.3, eps=eps)
    assert full.shape == (64, 64) and full.dtype == np.float32
    # transform=False takes the transformed (1,H,W) input; inverse_transform=False returns (1,1,H,W)
    y = p.transform(t, field="dm", z=0.3)
    assert y.shape == (1, 64, 64) and y.dtype == np.float32
    raw = p.paint(y, z=0.3, transform=False, inverse_transform=False, eps=eps)
    assert raw.shape == (1, 1, 64, 64)
    back = p.inverse_transform(raw, field="pressure", z=0.3)
    assert rel_l2(back, full) <= 1e-5
    # input is not mutated, output is a fresh array
    t0 = t.copy()
    a = p.paint(t, z=0.0, eps=eps)
    b = p.paint(t, z=0.0, eps=eps)
    assert np.array_equal(t, t0) and a is not b and np.array_equal(a, b)
    # unseeded paint draws a latent on the device: two calls differ, same seed agrees
    c = p.paint(t, z=0.0)
    d = p.paint(t, z=0.0)
    assert not np.array_equal(c, d)
    assert np.array_equal(p.paint(t, z=0.0, seed=5), p.paint(t, z=0.0, seed=5))
    # empty batch
    assert p.paint_batch(np.zeros((0, 64, 64), np.float32), z=0.0).shape == (0, 64, 64)


def test_checkpoint_roundtrip(tmp_path):
    """save_state_to_file / load_state_from_file keep the (model_state, model_meta) contract."""
    import torch
    from baryon_painter_b200.painter import CVAEPainter
    p = _painter(64, 3, "fp32")
    files = (str(tmp_path / "model_state"), str(tmp_path / "model_meta"))
    p.save_state_to_file(files)
    sd = torch.load(files[0])
    assert len(sd) == 179
    q = CVAEPainter(files, precision="fp32")
    for k in ("L", "n_grid", "tile_L", "n_tile", "tile_size", "input_field", "label_fields", "scale_to_SLICS"):
        assert getattr(q, k) == getattr(p, k)
    from baryon_painter_b200 import synthetic
    t = synthetic.synthetic_dm_tiles(1, 64, seed0=4)[0]
    eps = np.ones((1, 2, 2), np.float32)
    assert np.array_equal(p.paint(t, z=0.7, eps=eps), q.paint(t, z=0.7, eps=eps))
    with pytest.raises(ValueError):
        q.load_state_from_file("model_state")
    bad = dict(sd)
    bad.pop("p_mu_out.0.weight")
    with pytest.raises(RuntimeError, match="Missing key"):
        q.model.load_state_dict(bad)


def test_seed_mode_rng_is_standard_normal_and_restated():
    """BP_LATENT_SEED replaces torch.randn (reference cvae.py:64) by a counter-based generator.  (1) its draws are
    N(0,1): moments, tail fractions and a Kolmogorov-Smirnov distance over 2^20 draws; distinct seeds / offsets are
    uncorrelated; (2) oracle/rng_oracle.py restates it (the variance-map parity below feeds those eps to the oracle)."""
    from math import erf, sqrt
    from baryon_painter_b200 import _lib
    from oracle.rng_oracle import counter_normal
    n = 1 << 20
    e = _lib.rng_normal(1234, 77, n).astype(np.float64)
    assert abs(e.mean()) < 4.0 / np.sqrt(n) and abs(e.var() - 1) < 5e-3
    assert abs((e ** 3).mean()) < 1e-2 and abs((e ** 4).mean() - 3) < 3e-2
    for t, frac in ((1.0, 0.3173105), (2.0, 0.0455003), (3.0, 0.0026998)):
        assert abs((np.abs(e) > t).mean() / frac - 1) < 0.05
    xs = np.sort(e)
    cdf = 0.5 * (1 + np.vectorize(erf)(xs[::64] / sqrt(2)))
    assert np.max(np.abs(cdf - (np.arange(n)[::64] + 0.5) / n)) < 2.5e-3          # KS: ~1.4/sqrt(n) at 5 %
    e2 = _lib.rng_normal(1235, 77, n).astype(np.float64)
    e3 = _lib.rng_normal(1234, 77 + n, n).astype(np.float64)
    assert abs((e * e2).mean()) < 5e-3 and abs((e * e3).mean()) < 5e-3 and abs((e[:-1] * e[1:]).mean()) < 5e-3
    r = counter_normal(1234, 77, 4096)
    assert np.max(np.abs(r - e[:4096])) < 2e-6


@pytest.mark.parametrize("precision,n_draws,tol_mean,tol_var", [("fp32", 24, 1e-4, 1e-3), ("fp16", 64, 1e-2, 3e-2)])
def test_variance_maps_vs_oracle(precision, n_draws, tol_mean, tol_var):
    """BASELINE.json configs[3]: per-pixel mean / variance of the painted pressure over latent draws, against the
    oracle = mean / population variance of that many `sample_P` paints (reference cvae.py:149-162) fed the SAME eps
    (the device's counter generator restated in oracle/rng_oracle.py).  Tolerances: the per-tile bound of the
    precision for the mean map; the variance map is a difference of nearly equal numbers, so it is held to
    10x / 3x that (relative L2 over the map)."""
    import torch
    from oracle.cvae_oracle import CVAEOracle
    from oracle.rng_oracle import variance_draw_eps
    from baryon_painter_b200 import arch, synthetic, transforms
    torch.set_num_threads(os.cpu_count())
    tile, n, seed = 128, 3, 17
    A = arch.fiducial_cvae_architecture(tile)
    sd = synthetic.synthetic_cvae_state_dict(A, seed=4)
    stats = transforms.fiducial_stats()
    orc = CVAEOracle(A, sd)
    p = _painter(tile, 4, precision, max_batch=64)
    tiles = synthetic.synthetic_dm_tiles(n, tile, seed0=800)
    zs = [0.0, 0.5, 1.0]
    mean, var = p.paint_variance(tiles, z=zs, n_draws=n_draws, seed=seed)
    assert mean.shape == var.shape == (n, tile, tile) and np.all(var >= 0) and var.max() > 0
    eps = variance_draw_eps(seed, n, n_draws, (tile // 32, tile // 32))
    for t in range(n):
        draws = np.stack([orc.paint(tiles[t], zs[t], stats, eps=eps[d, t:t + 1]) for d in range(n_draws)]).astype(np.float64)
        assert rel_l2(mean[t], draws.mean(0)) <= tol_mean, (t, rel_l2(mean[t], draws.mean(0)))
        assert rel_l2(var[t], draws.var(0)) <= tol_var, (t, rel_l2(var[t], draws.var(0)))
    mean2, var2 = p.paint_variance(tiles, z=zs, n_draws=n_draws, seed=seed)
    assert np.array_equal(mean, mean2) and np.array_equal(var, var2)


def test_bit_identical_across_processes():
    """Which tensor-core formulation a layer runs with comes from the shipped table (or the cost model), never from
    a per-process timing: two fresh processes paint bit-identical tiles (the load-time tuner made
    results process-dependent)."""
    import hashlib
    import subprocess
    import sys
    code = ("import sys, hashlib, numpy as np; sys.path.insert(0, %r);"
            "from baryon_painter_b200 import synthetic; from baryon_painter_b200.painter import CVAEPainter;"
            "p = CVAEPainter.synthetic(tile_size=256, seed=1, precision='fp16', max_batch=8);"
            "t = synthetic.synthetic_dm_tiles(5, 256, seed0=60); e = synthetic.synthetic_latents(5, (8, 8), seed=2);"
            "o = p.paint_batch(t, z=[0.0, 0.2, 0.5, 1.0, 2.0], eps=e);"
            "print('DIGEST', hashlib.sha256(np.ascontiguousarray(o).tobytes()).hexdigest())") % os.path.dirname(os.path.dirname(GOLDEN))
    digests = []
    for _ in range(2):
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        digests.append([l for l in r.stdout.splitlines() if l.startswith("DIGEST")][-1])
    assert digests[0] == digests[1]


@pytest.mark.parametrize("n", [1, 17, 100])
def test_host_pipeline_matches_device_call(n):
    """The host entry point splits a batch into ramped pipeline chunks (16, 48, 64, ..., 48, 16) with copies overlapped
    on side streams; the device entry point runs it as whole plan chunks.  A tile's result must not depend on how
    the batch was cut: bit-identical, for pageable and for page-locked buffers."""
    import torch
    import baryon_painter_b200 as bp
    from baryon_painter_b200 import synthetic
    from baryon_painter_b200.painter import CVAEPainter
    tile = 64
    p = CVAEPainter.synthetic(tile_size=tile, seed=2, precision="fp16", max_batch=128)
    tiles = synthetic.synthetic_dm_tiles(8, tile, seed0=40)
    tiles = np.ascontiguousarray(np.concatenate([tiles] * 13)[:n] * np.linspace(0.7, 1.3, n, dtype=np.float32)[:, None, None])
    eps = synthetic.synthetic_latents(n, (tile // 32, tile // 32), seed=9)
    zs = np.linspace(0.0, 1.0, n)
    host = p.paint_batch(tiles, z=zs, eps=eps)
    pin_in, pin_out = bp.pinned_empty(tiles.shape), bp.pinned_empty(tiles.shape)
    pin_in[...] = tiles
    p.paint_batch(pin_in, z=zs, eps=eps, out=pin_out)
    dev = p.paint_batch_device(torch.from_numpy(tiles).cuda(), z=zs, eps=torch.from_numpy(eps).cuda()).cpu().numpy()
    assert np.isfinite(dev).all()
    assert np.array_equal(host, dev)
    assert np.array_equal(pin_out, dev)


def test_variance_maps_single_draw():
    """one draw per tile is that draw, with zero variance (the draw-batched passes of the 16-bit engine are checked
    against per-draw oracle paints in test_variance_maps_vs_oracle)"""
    from baryon_painter_b200 import synthetic
    p = _painter(64, 3, "fp16")
    tiles = synthetic.synthetic_dm_tiles(3, 64, seed0=8)
    zs = [0.0, 0.5, 1.0]
    m, v = p.paint_variance(tiles, z=zs, n_draws=1, seed=5)
    assert np.all(v == 0) and np.all(np.isfinite(m))
    m10, v10 = p.paint_variance(tiles, z=zs, n_draws=10, seed=5)          # chunk 16 // 3 tiles -> 5 draws per pass
    assert np.all(np.isfinite(m10)) and np.all(v10 >= 0) and v10.max() > 0


@pytest.mark.parametrize("precision,tol,tol_z", [("fp32", 1e-4, 2e-5), ("fp32-ffma", 1e-4, 2e-5), ("fp16", 1e-2, 5e-3)])
def test_elbo_vs_reference_golden(precision, tol, tol_z):
    """the recognition network Q and the ELBO forward (reference cvae.py:68-80, 122-147) on the
    device, against the reference's own CVAE.forward (golden): ELBO, KL term, log-likelihood, Q's (z_mu, z_log_var).
    Both entry points: already-transformed pairs (CVAEModel.forward) and raw tiles (CVAEPainter.elbo)."""
    g = np.load(os.path.join(GOLDEN, "cvae_t64_elbo.npz"))
    p = _painter(64, int(g["seed"]), precision)
    zs = g["z"]
    elbo = p.model.forward(g["x"], g["y"], aux_label=zs.astype(np.float32), eps=g["eps"])
    assert abs(float(elbo) - float(g["ELBO"])) <= tol * abs(float(g["ELBO"]))
    assert abs(float(p.model.KL_term) - float(g["KL_term"])) <= 10 * tol * abs(float(g["KL_term"])) + 1e-6
    assert abs(float(p.model.log_likelihood[0]) - float(g["log_likelihood"][0])) <= tol * abs(float(g["log_likelihood"][0]))
    assert rel_l2(p.model.z_mu.numpy(), g["z_mu"]) <= tol_z and rel_l2(p.model.z_log_var.numpy(), g["z_log_var"]) <= 10 * tol_z
    assert p.model.get_stats()[0] == float(elbo) and p.model.get_stats_labels()[:2] == ["ELBO", "KL_term"]
    from baryon_painter_b200 import synthetic
    dm = synthetic.synthetic_dm_tiles(3, 64, seed0=int(g["dm_seed0"]))
    e2, kl2, ll2 = p.elbo(g["pressure"], dm, z=zs, eps=g["eps"])
    assert abs(e2 - float(g["ELBO"])) <= 2 * tol * abs(float(g["ELBO"]))


def test_async_stream_matches_sync_calls():
    """paint_batch_async (two I/O slots, whole-chunk launches, copies of neighbouring batches overlapped) paints exactly
    what paint_batch paints, batch after batch, also when a slot is reused before its ticket was waited on."""
    import baryon_painter_b200 as bp
    from baryon_painter_b200 import synthetic
    from baryon_painter_b200.painter import CVAEPainter
    tile, n, nbatch = 64, 24, 5
    p = CVAEPainter.synthetic(tile_size=tile, seed=2, precision="fp16", max_batch=32)
    base = synthetic.synthetic_dm_tiles(n, tile, seed0=70)
    eps = synthetic.synthetic_latents(n, (tile // 32, tile // 32), seed=4)
    batches, outs, ref = [], [], []
    for b in range(nbatch):
        t = bp.pinned_empty(base.shape)
        t[...] = base * np.float32(0.8 + 0.1 * b)
        batches.append(t)
        outs.append(bp.pinned_empty(base.shape))
        ref.append(p.paint_batch(t, z=0.1 * b, eps=eps).copy())
    tickets = [p.paint_batch_async(batches[b], z=0.1 * b, eps=eps, out=outs[b]) for b in range(nbatch)]   # slots reused unwaited
    for b, tk in enumerate(tickets):
        got = tk.wait()
        assert got is outs[b] and np.array_equal(got, ref[b]), b
    # seed mode and latent mode go through the same slots
    a = p.paint_batch_async(batches[0], z=0.0, seed=7).wait().copy()
    assert np.array_equal(a, p.paint_batch(batches[0], z=0.0, seed=7))
    with pytest.raises(ValueError, match="Shape mismatch"):
        p.paint_batch_async(np.zeros((2, 32, 32), np.float32))


@pytest.mark.parametrize("tile,precision", [(96, "fp16"), (96, "fp32"), (160, "fp16"), (256, "fp16")])
def test_other_tile_sizes_vs_oracle(tile, precision):
    """tile sizes other than the fixtures' powers of two (the reference network is fully convolutional: any multiple of
    32): widths that are not powers of two take the general index arithmetic of the front kernels, partial M-tiles
    and layer shapes without an entry in the formulation table; ragged batch of 3."""
    import torch
    from oracle.cvae_oracle import CVAEOracle
    from baryon_painter_b200 import arch, synthetic, transforms
    torch.set_num_threads(os.cpu_count())
    A = arch.fiducial_cvae_architecture(tile)
    sd = synthetic.synthetic_cvae_state_dict(A, seed=21)
    stats = transforms.fiducial_stats()
    orc = CVAEOracle(A, sd)
    p = _painter(tile, 21, precision, max_batch=2)
    tiles = synthetic.synthetic_dm_tiles(3, tile, seed0=900)
    zs = np.array([0.0, 0.4, 1.2])
    eps = synthetic.synthetic_latents(3, (tile // 32, tile // 32), seed=5)
    ref = orc.paint_batch(tiles, zs, stats, eps=eps)
    out = p.paint_batch(tiles, z=zs, eps=eps)
    assert out.shape == (3, tile, tile)
    for i in range(3):
        assert rel_l2(out[i], ref[i]) <= TOL[precision], (i, rel_l2(out[i], ref[i]))
