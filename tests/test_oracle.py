"""The oracle (oracle/cvae_oracle.py) against the committed golden vectors and -- when the
reference checkout is present (build container only) -- against the unmodified reference."""
import hashlib
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, rel_l2
from oracle import cvae_oracle, ref_shims
from baryon_painter_b200 import arch, synthetic, transforms


def _digest(sd):
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(np.ascontiguousarray(v.numpy()).tobytes())
    return h.hexdigest()


@pytest.mark.parametrize("name,n", [("cvae_t64_layers.npz", 1), ("cvae_t128.npz", 4)])
def test_oracle_reproduces_golden(name, n):
    g = np.load(os.path.join(GOLDEN, name))
    tile = int(g["tile_size"])
    A = arch.fiducial_cvae_architecture(tile)
    sd = synthetic.synthetic_cvae_state_dict(A, seed=int(g["seed"]))
    assert _digest(sd) == str(g["state_digest"]), "synthetic weight generator drifted"
    stats = transforms.fiducial_stats()
    orc = cvae_oracle.CVAEOracle(A, sd)
    tiles = synthetic.synthetic_dm_tiles(n, tile, seed0=int(g["tiles_seed0"]))
    torch.set_num_threads(os.cpu_count())
    for i in range(n):
        z = float(g["z"][i])
        e = orc.paint(tiles[i], z, stats, eps=g["eps"][i:i + 1])
        l = orc.paint(tiles[i], z, stats, latent=g["eps"][i:i + 1])
        # same torch build -> bit-exact; another torch/oneDNN build may reorder sums
        assert rel_l2(e, g["painted_E"][i]) < 1e-5 and rel_l2(l, g["painted_L"][i]) < 1e-5
        assert e.dtype == np.float32 and e.shape == (tile, tile)


def test_oracle_layer_taps_match_golden():
    g = np.load(os.path.join(GOLDEN, "cvae_t64_layers.npz"))
    A = arch.fiducial_cvae_architecture(64)
    sd = synthetic.synthetic_cvae_state_dict(A, seed=int(g["seed"]))
    orc = cvae_oracle.CVAEOracle(A, sd)
    taps = []
    orc.sample_P(g["y0"][None], np.float32(g["z"][0]), eps=g["eps"][0:1], taps=taps)
    got = dict(taps)
    n = 0
    for k in g.files:
        if k.startswith("tap:"):
            assert rel_l2(got[k[4:]].numpy()[0], g[k]) < 1e-5, k
            n += 1
    assert n >= 20


def test_golden_512_digest_and_shapes():
    g = np.load(os.path.join(GOLDEN, "cvae_t512.npz"))
    sd = synthetic.synthetic_cvae_state_dict(arch.fiducial_cvae_architecture(512), seed=int(g["seed"]))
    assert _digest(sd) == str(g["state_digest"])
    assert g["painted_E"].shape == (1, 512, 512) and g["painted_L"].shape == (1, 512, 512)
    assert np.all(np.isfinite(g["painted_E"])) and g["painted_E"].min() >= 0


def test_parameter_count_matches_published():
    """notebooks/validation_plots.ipynb cell 5: 1 662 961 trainable parameters."""
    stacks = arch.cvae_stacks(arch.fiducial_cvae_architecture(512))
    schema = arch.state_dict_schema(stacks)
    n = sum(int(np.prod(s)) for k, (s, dt) in schema.items()
            if dt == "float32" and not k.endswith(("running_mean", "running_var")))
    assert n == 1662961
    assert len(schema) == 179


def test_flop_count_matches_survey():
    stacks = arch.cvae_stacks(arch.fiducial_cvae_architecture(512))
    f_prior, _ = arch.stack_flops(stacks["prior_network"], 512, 512)
    f_pz, _ = arch.stack_flops(stacks["p_z_in"], 16, 16)
    f_pyz, hw = arch.stack_flops(stacks["p_y_z_in"], 512, 512)
    f_mu, _ = arch.stack_flops(stacks["p_mu_out"], *hw)
    total = f_prior + f_pz + f_pyz + f_mu
    assert abs(total / 1e9 - 20.254) < 0.01, total
    gen = arch.flatten_stack(arch.fiducial_cgan_architecture(), "generator")
    f_gan, _ = arch.stack_flops(gen, 512, 512)
    assert abs(f_gan / 1e9 - 100.7) < 0.5, f_gan


@pytest.mark.skipif(not ref_shims.available(), reason="reference checkout not present")
def test_oracle_equals_reference_bitwise():
    A = arch.fiducial_cvae_architecture(64)
    sd = synthetic.synthetic_cvae_state_dict(A, seed=5)
    stats = transforms.fiducial_stats()
    painter = ref_shims.reference_painter(A, sd, stats)
    orc = cvae_oracle.CVAEOracle(A, sd)
    tile = synthetic.synthetic_dm_tiles(1, 64, seed0=77)[0]
    for z in (0.0, 0.31, 2.2):
        torch.manual_seed(42)
        ref = painter.paint(tile, z=z)
        torch.manual_seed(42)
        eps = torch.randn(size=(1, 1, 1, 2, 2)).numpy().reshape(1, 1, 2, 2)
        assert np.array_equal(ref, orc.paint(tile, z, stats, eps=eps))
    # reference schema == ours, in order
    assert list(painter.model.state_dict().keys()) == list(
        arch.state_dict_schema(arch.cvae_stacks(A)).keys())
    assert painter.model.count_parameters() == sum(
        int(np.prod(v.shape)) for k, v in sd.items()
        if v.dtype == torch.float32 and not k.endswith(("running_mean", "running_var")))


@pytest.mark.skipif(not ref_shims.available(), reason="reference checkout not present")
def test_fiducial_architecture_equals_shipped_file():
    import ast
    path = os.path.join(ref_shims.REFERENCE_ROOT, "trained_models/CVAE/fiducial/architecture.txt")
    assert ast.literal_eval(open(path).read()) == arch.fiducial_cvae_architecture(512)


def test_cgan_oracle_runs_small():
    layers = arch.fiducial_cgan_architecture(n_res_blocks=2)
    sd = synthetic.synthetic_cgan_state_dict(layers, seed=1)
    orc = cvae_oracle.CGANOracle(layers, sd)
    tile = synthetic.synthetic_dm_tiles(1, 64, seed0=3)[0]
    out = orc.paint(tile, 0.5, transforms.fiducial_stats())
    assert out.shape == (64, 64) and out.dtype == np.float32 and np.all(np.isfinite(out))


def test_oracle_forward_matches_reference_golden():
    """Q + ELBO restatement (oracle.forward) against the reference's own CVAE.forward (golden written by
    oracle/make_golden.py with torch.randn pinned): bit-identical on this torch build."""
    import torch
    torch.set_num_threads(4)
    g = np.load(os.path.join(GOLDEN, "cvae_t64_elbo.npz"))
    A = arch.fiducial_cvae_architecture(int(g["tile_size"]))
    o = cvae_oracle.CVAEOracle(A, synthetic.synthetic_cvae_state_dict(A, seed=int(g["seed"])))
    r = o.forward(g["x"], g["y"], g["z"].astype(np.float32), g["eps"])
    assert r["ELBO"] == pytest.approx(float(g["ELBO"]), rel=1e-6) and r["KL_term"] == pytest.approx(float(g["KL_term"]), rel=1e-6)
    assert np.allclose(r["z_mu"], g["z_mu"], rtol=1e-5, atol=1e-7) and np.allclose(r["log_likelihood"], g["log_likelihood"], rtol=1e-6)
