"""TEST INFRASTRUCTURE -- generate tests/golden/* from the UNMODIFIED reference.

Run in the build container (needs /root/reference):

    python -m oracle.make_golden

For every fixture the reference's own ``CVAEPainter.paint`` / ``CVAE.sample_P`` /
``CVAE.prior`` (imported through ``oracle/ref_shims.py``) is run on seeded synthetic
weights (``baryon_painter_b200.synthetic``), seeded tiles and seeded latents, the oracle
restatement (``oracle/cvae_oracle.py``) is asserted to agree with it bit-for-bit, and the
reference outputs are written out.  The fixtures travel to the GPU box; the reference
does not.

Fixtures
--------
cvae_t64_layers.npz   tile 64 network, one tile: every layer-boundary tensor of the paint
                      path (after prior, after p_z_in, after each p_y_z_in stage, after
                      p_mu_out), mode E
cvae_t128.npz         tile 128 network, 4 tiles x z in {0, 0.3, 1.0, 2.5}: painted tiles in
                      mode E (eps supplied) and mode L (latent supplied), (mu, log-var)
cvae_t512.npz         fiducial 512 network, 2 tiles: painted tiles, mode E at z=0 and
                      mode L at z=0.7
cvae_t64_elbo.npz     tile 64 network, 3 (pressure, dark matter) pairs: CVAE.forward's ELBO, KL term,
                      log-likelihood and the recognition network's (z_mu, z_log_var)
fiducial_meta.json    plain keys + stats table + architecture lifted from the shipped
                      trained_models/CVAE/fiducial/model_meta
tiling.json           reference generate_tiling / make_weight_map / get_tile known answers
"""

import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import cvae_oracle, ref_shims                        # noqa: E402
from baryon_painter_b200 import arch, meta, synthetic, transforms  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def state_digest(sd):
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(np.ascontiguousarray(v.numpy()).tobytes())
    return h.hexdigest()


def ref_mode_E(painter, tile, z, eps):
    """reference paint() with torch.randn replaced by ``eps``."""
    m = painter.model
    m.train(False)
    with torch.no_grad():
        y = painter.transform(tile, field="dm", z=z)
        y = torch.tensor(y.reshape(1, *y.shape))
        aux = torch.tensor(z, dtype=y.dtype)
        mu, lv = m.prior(y, aux)
        lat = mu + torch.tensor(eps).view(1, *mu.size()) * (torch.exp(lv / 2) + m.min_z_var)
        lat = lat.view(-1, *m.dim_z)
        pred = m.sample_P(y, aux_label=aux, z=lat.numpy()).cpu().numpy()
    return painter.inverse_transform(pred, field="pressure", z=z), mu.numpy(), lv.numpy(), pred


def ref_mode_L(painter, tile, z, latent):
    m = painter.model
    m.train(False)
    with torch.no_grad():
        y = painter.transform(tile, field="dm", z=z)
        y = torch.tensor(y.reshape(1, *y.shape))
        aux = torch.tensor(z, dtype=y.dtype)
        pred = m.sample_P(y, aux_label=aux, z=latent).cpu().numpy()
    return painter.inverse_transform(pred, field="pressure", z=z), pred


def make_cvae(tile_size, n_tiles, zs, seed, with_layers=False):
    A = arch.fiducial_cvae_architecture(tile_size)
    sd = synthetic.synthetic_cvae_state_dict(A, seed=seed)
    stats = transforms.fiducial_stats()
    painter = ref_shims.reference_painter(A, sd, stats)
    oracle = cvae_oracle.CVAEOracle(A, sd)
    lat_hw = A["dim_z"][1:]
    tiles = synthetic.synthetic_dm_tiles(n_tiles, tile_size, seed0=100 * seed)
    eps = synthetic.synthetic_latents(n_tiles, lat_hw, seed=1)
    out = {"seed": np.int64(seed), "tile_size": np.int64(tile_size), "z": np.asarray(zs, np.float64),
           "eps": eps, "state_digest": np.array(state_digest(sd))}
    pe, pl, mus, lvs = [], [], [], []
    for i in range(n_tiles):
        z = float(zs[i])
        p_e, mu, lv, _ = ref_mode_E(painter, tiles[i], z, eps[i:i + 1])
        p_l, _ = ref_mode_L(painter, tiles[i], z, eps[i:i + 1])
        # the restatement must agree with the reference exactly
        o_e = oracle.paint(tiles[i], z, stats, eps=eps[i:i + 1])
        o_l = oracle.paint(tiles[i], z, stats, latent=eps[i:i + 1])
        assert np.array_equal(o_e, p_e) and np.array_equal(o_l, p_l), "oracle != reference"
        assert p_e.dtype == np.float32
        pe.append(p_e); pl.append(p_l); mus.append(mu[0]); lvs.append(lv[0])
    # the unseeded reference entry point itself, with torch's RNG pinned
    torch.manual_seed(1234)
    p_ref = painter.paint(tiles[0], z=float(zs[0]))
    torch.manual_seed(1234)
    e0 = torch.randn(size=(1, 1, 1, *lat_hw)).numpy().reshape(1, 1, *lat_hw)
    assert np.array_equal(p_ref, oracle.paint(tiles[0], float(zs[0]), stats, eps=e0))
    out.update(painted_E=np.stack(pe), painted_L=np.stack(pl), z_mu=np.stack(mus), z_log_var=np.stack(lvs))
    if with_layers:
        taps = []
        y = cvae_oracle.forward_transform(tiles[0], float(zs[0]), stats)[None]
        oracle.sample_P(y, np.float32(zs[0]), eps=eps[0:1], taps=taps)
        # cross-check a few taps against forward hooks on the reference modules
        got = {}
        hooks = [mod.register_forward_hook(lambda m_, i_, o_, n=name: got.__setitem__(n, o_.detach().clone()))
                 for name, mod in painter.model.named_modules() if name.count(".") == 1]
        ref_mode_E(painter, tiles[0], float(zs[0]), eps[0:1])
        for h in hooks:
            h.remove()
        keep = {}
        specs = arch.cvae_stacks(A)
        boundary = set()
        for name in ("prior_network", "p_z_in", "p_y_z_in", "p_mu_out"):
            layers = {"prior_network": A["prior_z_y"], "p_z_in": A["p_z_in"], "p_y_z_in": A["p_y_z_in"],
                      "p_mu_out": A["p_y_z_out"][0]}[name]
            for idx, layer in enumerate(layers):
                nxt = layers[idx + 1][0].lower() if idx + 1 < len(layers) else None
                tag = layer[0].lower()
                # keep the tensor after each conv(+bn)(+act) group and after each residual block
                if tag in ("relu", "prelu", "softplus", "residual block") or \
                        (tag == "batchnorm" and nxt not in ("relu", "prelu", "softplus")):
                    boundary.add("%s.%d" % (name, idx))
        for k, v in taps:
            if k in got:
                assert torch.equal(got[k], v), k
            if k in boundary or k in ("z_mu", "z_log_var", "latent"):
                keep["tap:" + k] = v.numpy()[0]
        out.update(keep)
        out["y0"] = y[0]
    out["tiles_seed0"] = np.int64(100 * seed)
    return out


def make_elbo(tile_size, n_tiles, zs, seed):
    """reference ``CVAE.forward`` (cvae.py:122-147) with ``torch.randn`` pinned to ``eps``: ELBO, KL term,
    log-likelihood and Q's (z_mu, z_log_var) of a batch of (pressure, dark-matter) tile pairs."""
    A = arch.fiducial_cvae_architecture(tile_size)
    sd = synthetic.synthetic_cvae_state_dict(A, seed=seed)
    stats = transforms.fiducial_stats()
    painter = ref_shims.reference_painter(A, sd, stats)
    oracle = cvae_oracle.CVAEOracle(A, sd)
    lat_hw = A["dim_z"][1:]
    dm = synthetic.synthetic_dm_tiles(n_tiles, tile_size, seed0=100 * seed + 50)
    # a pressure-like positive field correlated with the dark matter
    pr = (0.05 * dm ** 1.5 * np.random.default_rng(seed).lognormal(0.0, 0.3, dm.shape)).astype(np.float32)
    eps = synthetic.synthetic_latents(n_tiles, lat_hw, seed=3)
    y = np.stack([painter.transform(dm[i], field="dm", z=float(zs[i])) for i in range(n_tiles)]).astype(np.float32)
    x = np.stack([painter.transform(pr[i], field="pressure", z=float(zs[i])) for i in range(n_tiles)]).astype(np.float32)
    m = painter.model
    m.train(False)
    real_randn = torch.randn
    try:
        torch.randn = lambda size, device=None: torch.tensor(eps).view(*size)      # sample_z's only draw
        with torch.no_grad():
            elbo = m.forward(torch.tensor(x), torch.tensor(y), torch.tensor(np.asarray(zs, np.float32)))
    finally:
        torch.randn = real_randn
    o = oracle.forward(x, y, np.asarray(zs, np.float32), eps)
    assert float(elbo) == o["ELBO"] and float(m.KL_term) == o["KL_term"], "oracle != reference"
    assert np.array_equal(m.z_mu.numpy(), o["z_mu"]) and np.array_equal(m.z_log_var.numpy(), o["z_log_var"])
    return {"seed": np.int64(seed), "tile_size": np.int64(tile_size), "z": np.asarray(zs, np.float64), "eps": eps,
            "dm_seed0": np.int64(100 * seed + 50), "pressure": pr, "x": x, "y": y,
            "ELBO": np.float64(float(elbo)), "KL_term": np.float64(float(m.KL_term)),
            "log_likelihood": m.log_likelihood.numpy().astype(np.float64),
            "z_mu": m.z_mu.numpy(), "z_log_var": m.z_log_var.numpy()}


def make_meta_json():
    path = os.path.join(ref_shims.REFERENCE_ROOT, "trained_models", "CVAE", "fiducial", "model_meta")
    d = meta.read_model_meta(path)
    t, it = d["transform"], d["inverse_transform"]
    js = {k: d[k] for k in ("L", "n_grid", "tile_L", "n_tile", "tile_size", "input_field",
                            "label_fields", "scale_to_SLICS")}
    js["stats"] = {f: {repr(z): {s: [repr(float(v)), type(v).__name__] for s, v in sv.items()}
                       for z, sv in t.stats[f].items()} for f in t.stats}
    js["k_values"], js["modes"], js["eps"] = t.k_values, t.modes, t.eps
    js["steps"] = {"transform": t.steps, "inverse_transform": it.steps}
    js["architecture_repr_sha256"] = hashlib.sha256(repr(d["model_architecture"]).encode()).hexdigest()
    js["architecture_equals_fiducial_builder"] = d["model_architecture"] == arch.fiducial_cvae_architecture()
    return js


def make_tiling_json():
    bp = ref_shims.import_reference()
    ps = bp.process_SLICS
    cases = []
    for n_plane, n_tile, ov in [(512, 256, 0.0), (512, 250, 0.0), (512, 256, 0.5), (512, 128, 0.0),
                                (512, 32, 0.33), (564, 512, 0.5), (1129, 512, 0.5), (3271, 512, 0.5),
                                (700, 512, 0.5), (2343, 512, 0.5)]:
        origins, slices = ps.generate_tiling(n_plane, n_tile, ov)
        cases.append({"n_pixel_plane": n_plane, "n_pixel_tile": n_tile, "min_tile_overlap": ov,
                      "origins": [float(o) for o in origins],
                      "starts": [int(s[0].start) for s in slices[0]] and [int(r[0][0].start) for r in slices]})
    w = ps.make_weight_map((512, 512), falloff=0.05, sigma=0.5)
    w64 = ps.make_weight_map((64, 64), falloff=0.1, sigma=1)
    rng = np.random.default_rng(7)
    m = rng.standard_normal((97, 97)).astype(np.float32)
    tiles = []
    for shift, rel, ef in [((0.0, 0.0), 0.5, 1), ((0.7, 0.9), 0.5, 1), ((0.3, 0.6), 0.2, 2.5), ((0.95, 0.05), 0.33, 1)]:
        t = ps.get_tile(m, shift, rel, ef)
        tiles.append({"shift": shift, "tile_relative_size": rel, "expansion_factor": ef,
                      "shape": list(t.shape), "sum": float(t.astype(np.float64).sum()),
                      "corner": [float(t[0, 0]), float(t[0, -1]), float(t[-1, 0]), float(t[-1, -1])]})
    return {"generate_tiling": cases,
            "weight_512": {"corner": float(w[0, 0]), "edge": float(w[0, 256]), "centre": float(w[256, 256]),
                           "row0": [float(v) for v in w[:30, 256]], "sum": float(w.sum())},
            "weight_64": {"row0": [float(v) for v in w64[:8, 32]], "sum": float(w64.sum())},
            "get_tile_seed": 7, "get_tile_shape": [97, 97], "get_tile": tiles}


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    np.savez(os.path.join(GOLDEN, "cvae_t64_layers.npz"), **make_cvae(64, 1, [0.5], seed=3, with_layers=True))
    np.savez(os.path.join(GOLDEN, "cvae_t128.npz"), **make_cvae(128, 4, [0.0, 0.3, 1.0, 2.5], seed=2))
    full = make_cvae(512, 2, [0.0, 0.7], seed=0)
    full["painted_E"] = full["painted_E"][:1]          # tile 0, mode E, z=0
    full["painted_L"] = full["painted_L"][1:]          # tile 1, mode L, z=0.7
    np.savez(os.path.join(GOLDEN, "cvae_t512.npz"), **full)
    np.savez(os.path.join(GOLDEN, "cvae_t64_elbo.npz"), **make_elbo(64, 3, [0.0, 0.5, 1.2], seed=5))
    with open(os.path.join(GOLDEN, "fiducial_meta.json"), "w") as f:
        json.dump(make_meta_json(), f, indent=1, sort_keys=True)
    with open(os.path.join(GOLDEN, "tiling.json"), "w") as f:
        json.dump(make_tiling_json(), f, indent=1, sort_keys=True)
    for fn in sorted(os.listdir(GOLDEN)):
        print("%-28s %9d bytes" % (fn, os.path.getsize(os.path.join(GOLDEN, fn))))


if __name__ == "__main__":
    main()
