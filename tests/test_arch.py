"""Architecture flattening, state_dict schema, BN folding (host logic; CPU only)."""
import numpy as np
import pytest
import torch

from baryon_painter_b200 import arch, synthetic
from oracle import cvae_oracle


def test_flatten_fiducial_layer_table():
    st = arch.cvae_stacks(arch.fiducial_cvae_architecture(512))
    assert [len(st[k]) for k in ("prior_network", "p_z_in", "p_y_z_in", "p_mu_out")] == [4, 3, 15, 3]
    pyz = st["p_y_z_in"]
    assert (pyz[0].cin, pyz[0].cout, pyz[0].k, pyz[0].pad) == (3, 16, 5, 2)
    assert [s.res for s in pyz[4:12]] == [1, 2] * 4
    assert pyz[4].w_key == "p_y_z_in.12.res_block.0.weight" and pyz[5].bn_prefix == "p_y_z_in.12.res_block.4"
    assert pyz[12].kind == "convT" and pyz[12].w_key == "p_y_z_in.16.weight"
    mu = st["p_mu_out"]
    assert [s.act for s in mu] == ["prelu", "prelu", "softplus"] and mu[0].act_key == "p_mu_out.1.weight"
    assert st["p_z_in"][1].out_hw(32, 32) == (128, 128) and st["prior_network"][1].out_hw(256, 256) == (64, 64)


def test_schema_key_examples():
    sc = arch.state_dict_schema(arch.cvae_stacks(arch.fiducial_cvae_architecture(512)))
    assert sc["p_y_z_in.16.weight"][0] == (128, 64, 4, 4)
    assert sc["p_mu_out.0.weight"][0] == (8, 16, 7, 7) and sc["p_mu_out.1.weight"][0] == (1,)
    assert sc["prior_network.1.num_batches_tracked"] == ((), "int64")
    assert list(sc)[0] == "q_x_in.0.weight" and list(sc)[-1] == "prior_network.10.num_batches_tracked"


def test_strict_state_dict_errors():
    A = arch.fiducial_cvae_architecture(64)
    sc = arch.state_dict_schema(arch.cvae_stacks(A))
    sd = synthetic.synthetic_cvae_state_dict(A, seed=0)
    arch.check_state_dict(sc, sd)
    bad = dict(sd); bad["extra.weight"] = torch.zeros(1)
    with pytest.raises(RuntimeError, match="Unexpected key"):
        arch.check_state_dict(sc, bad)
    bad = dict(sd); bad["p_mu_out.0.weight"] = torch.zeros(8, 16, 5, 5)
    with pytest.raises(RuntimeError, match="size mismatch"):
        arch.check_state_dict(sc, bad)


def test_bn_fold_equals_unfolded_oracle():
    """conv*scale+shift (folded) == BatchNorm2d(eval)(conv) for every layer of p_y_z_in on a small input."""
    import torch.nn.functional as F
    A = arch.fiducial_cvae_architecture(64)
    sd = synthetic.synthetic_cvae_state_dict(A, seed=4)
    specs = arch.cvae_stacks(A)["p_y_z_in"]
    folded = arch.fold_stack(specs, sd)
    x = torch.from_numpy(np.random.default_rng(0).standard_normal((2, 3, 64, 64)).astype(np.float32))
    ref = cvae_oracle.run_sequential(A["p_y_z_in"], sd, "p_y_z_in", x)
    h, skip = x.double(), None
    for f in folded:
        s = f.spec
        w = torch.from_numpy(f.weight).double()
        if s.res == arch.RES_OPEN:
            skip = h
        if s.kind == "conv":
            y = F.conv2d(h, w, stride=s.stride, padding=s.pad)
        else:
            y = F.conv_transpose2d(h, w, stride=s.stride, padding=s.pad, output_padding=s.out_pad)
        y = y * torch.from_numpy(f.scale).double().view(1, -1, 1, 1) + torch.from_numpy(f.shift).double().view(1, -1, 1, 1)
        if s.res == arch.RES_CLOSE:
            y = y + skip
        assert s.act == "relu"
        h = F.relu(y)
    err = ((h - ref.double()).norm() / ref.double().norm()).item()
    assert err < 1e-5, err


def test_unsupported_layers_raise():
    with pytest.raises(NotImplementedError):
        arch.flatten_stack([("linear", {"in_features": 3, "out_features": 2})], "x")
    with pytest.raises(RuntimeError, match="ill-formed"):
        arch.flatten_stack([("conv", {}, 3)], "x")
    with pytest.raises(NotImplementedError):
        arch.cvae_stacks({"type": "Type-2"})


def test_cgan_layer_table():
    g = arch.flatten_stack(arch.fiducial_cgan_architecture(), "generator")
    assert len(g) == 3 + 18 + 3
    assert (g[0].k, g[0].pad, g[0].b_key) == (9, 4, None) and g[1].b_key == "generator.3.bias"
    assert g[-3].kind == "convT" and g[-3].out_pad == 1 and g[-3].out_hw(128, 128) == (256, 256)
    assert g[-1].act == "tanh" and g[-1].bn_prefix is None and g[3].act == "leaky" and g[4].res == 2
