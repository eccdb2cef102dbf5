"""model_meta compat reader / writer and the host-side transforms."""
import json
import os
import pickle

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import cvae_oracle, ref_shims
from baryon_painter_b200 import arch, meta, transforms


def test_fiducial_stats_match_extracted_meta():
    js = json.load(open(os.path.join(GOLDEN, "fiducial_meta.json")))
    stats = transforms.fiducial_stats()
    for field in ("dm", "pressure"):
        assert [repr(z) for z in stats[field]] == list(js["stats"][field].keys())
        for z, sv in stats[field].items():
            for s in ("mean", "var"):
                val, typ = js["stats"][field][repr(z)][s]
                assert repr(float(sv[s])) == val and type(sv[s]).__name__ == typ
    assert js["architecture_equals_fiducial_builder"] is True
    assert js["k_values"] == {"dm": 4.0, "pressure": 4} and js["eps"] == 1e-4
    assert js["L"] == 400 and js["n_grid"] == 2048 and js["tile_size"] == 512 and js["n_tile"] == 4
    assert js["input_field"] == "dm" and js["label_fields"] == ["pressure"] and js["scale_to_SLICS"] is True


@pytest.mark.skipif(not ref_shims.available(), reason="reference checkout not present")
def test_read_shipped_model_meta():
    path = os.path.join(ref_shims.REFERENCE_ROOT, "trained_models/CVAE/fiducial/model_meta")
    d = meta.read_model_meta(path)
    assert set(d) == set(meta.META_KEYS)
    assert d["model_architecture"] == arch.fiducial_cvae_architecture(512)
    t, it = d["transform"], d["inverse_transform"]
    assert t.steps == ("transform", "atleast_3d") and it.steps == ("squeeze", "inv_transform")
    assert t.is_fusable("dm") and it.is_fusable("pressure")
    # numerically identical to the reference factories bound to the same stats
    bp = ref_shims.import_reference()
    dt, ds = bp.utils.data_transforms, bp.utils.datasets
    rc, rc_inv = dt.create_range_compress_transforms(k_values={"dm": 4.0, "pressure": 4},
                                                     modes={"dm": "shift-log", "pressure": "shift-log"}, eps=1e-4)
    fwd = ds.compile_transform(dt.chain_transformations([rc, dt.atleast_3d]), t.stats)
    inv = ds.compile_transform(dt.chain_transformations([dt.squeeze, rc_inv]), t.stats)
    x = np.random.default_rng(0).lognormal(-0.5, 1, (32, 32)).astype(np.float32)
    for z in (-1.0, 0.0, 0.1, 0.125, 0.9, 2.0, 2.5):
        a, b = t(x, field="dm", z=z), np.asarray(fwd(x, "dm", z), np.float32)
        assert a.shape == (1, 32, 32) and a.dtype == np.float32 and np.array_equal(a, b)
        p = np.abs(np.random.default_rng(1).normal(0.5, 0.3, (1, 1, 32, 32))).astype(np.float32)
        assert np.array_equal(it(p, field="pressure", z=z), inv(p, "pressure", z))


@pytest.mark.skipif(not ref_shims.available(), reason="reference checkout not present")
def test_read_cgan_transform_pickles():
    from baryon_painter_b200.painter import _load_pickled_fn
    base = os.path.join(ref_shims.REFERENCE_ROOT, "trained_models/CGAN/fiducial")
    t = meta._rebind(_load_pickled_fn(os.path.join(base, "transform.pickle")))
    it = meta._rebind(_load_pickled_fn(os.path.join(base, "inv_transform.pickle")))
    assert t.modes["dm"] == "shift-log-cam" and t.k_values["dm"] == [4.0, 1.0]
    assert t.gpu_params("dm", 0.0)[0] == 1 and it.gpu_params("pressure", 0.0)[2:] == (4.0, 1.0)
    ft, fit = transforms.fiducial_transforms("cgan")
    x = np.random.default_rng(0).lognormal(-0.5, 1, (16, 16)).astype(np.float32)
    assert np.array_equal(t(x, field="dm", z=0.4), ft(x, field="dm", z=0.4))


def test_meta_roundtrip_plain_pickle(tmp_path):
    t, it = transforms.fiducial_transforms("cvae")
    d = {"L": 400, "n_grid": 2048, "tile_L": 100.0, "n_tile": 4, "tile_size": 512, "input_field": "dm",
         "label_fields": ["pressure"], "scale_to_SLICS": True, "transform": t, "inverse_transform": it,
         "model_architecture": arch.fiducial_cvae_architecture(512)}
    fn = str(tmp_path / "model_meta")
    meta.write_model_meta(fn, d)
    assert set(pickle.load(open(fn, "rb"))) == set(meta.META_KEYS)     # loads without dill
    e = meta.read_model_meta(fn)
    assert e["model_architecture"] == d["model_architecture"] and e["L"] == 400
    x = np.random.default_rng(0).lognormal(-0.5, 1, (8, 8)).astype(np.float32)
    assert np.array_equal(e["transform"](x, field="dm", z=0.3), t(x, field="dm", z=0.3))


def test_meta_reader_refuses_arbitrary_globals(tmp_path):
    fn = str(tmp_path / "evil")
    with open(fn, "wb") as f:
        pickle.dump(os.system, f)
    with pytest.raises(pickle.UnpicklingError):
        meta.read_model_meta(fn)


def test_transforms_agree_with_oracle_restatement():
    stats = transforms.fiducial_stats()
    t, it = transforms.fiducial_transforms("cvae")
    x = np.random.default_rng(3).lognormal(-0.5, 1, (16, 16)).astype(np.float32)
    for z in (0.0, 0.2, 1.3, 5.0):
        assert np.array_equal(t(x, field="dm", z=z), cvae_oracle.forward_transform(x, z, stats))
        y = np.abs(x[None, None]) * 0.3
        assert np.array_equal(it(y, field="pressure", z=z), cvae_oracle.inverse_transform(y, z, stats))
        # round trip at the tolerance idea of reference tests/test_dataset.py:80-83 (2e-5*sigma absolute)
        sig = t.sigma("dm", z)
        ti = transforms.CompiledTransform(stats, t.k_values, t.modes, inverse=True)
        back = ti(t(x, field="dm", z=z), field="dm", z=z)
        assert np.allclose(back, x, atol=2e-5 * sig * max(1.0, x.max()))
    assert t.sigma("dm", -3.0) == t.sigma("dm", 0.0) and t.sigma("dm", 9.0) == t.sigma("dm", 2.0)


def test_restricted_unpickler_refuses_escape_chains(tmp_path):
    """a crafted model_meta must not reach os.system through dill's _import_module / _get_attr or
    builtins.getattr; module names are an exact whitelist ('numpy_evil' does not pass a prefix test)."""
    import io
    import pickle
    import pickletools  # noqa: F401
    from baryon_painter_b200 import meta

    def stream(ops):
        return b"\x80\x02" + ops + b"."

    def glob(mod, name):
        return b"c" + mod.encode() + b"\n" + name.encode() + b"\n"

    def ustr(s):
        return b"X" + len(s).to_bytes(4, "little") + s.encode()

    # dill._dill._import_module('numpy.lib._npyio_impl')
    bad1 = stream(glob("dill._dill", "_import_module") + ustr("numpy.lib._npyio_impl") + b"\x85R")
    # getattr(_import_module('numpy'), 'os')
    bad2 = stream(glob("builtins", "getattr") + glob("dill._dill", "_import_module") + ustr("numpy") + b"\x85R" +
                  ustr("os") + b"\x86R")
    # dill._dill._import_module('numpy_evil')
    bad3 = stream(glob("dill._dill", "_import_module") + ustr("numpy_evil") + b"\x85R")
    # getattr on a real object (a dict)
    bad4 = stream(glob("builtins", "getattr") + b"}" + ustr("get") + b"\x86R")
    # a global outside the whitelist
    bad5 = stream(glob("os", "system") + ustr("true") + b"\x85R")
    for blob in (bad1, bad2, bad3, bad4, bad5):
        with pytest.raises(pickle.UnpicklingError):
            meta._MetaUnpickler(io.BytesIO(blob)).load()
    # the whitelisted numpy reconstructors still resolve through the same doors
    ok = stream(glob("builtins", "getattr") + glob("dill._dill", "_import_module") + ustr("numpy") + b"\x85R" +
                ustr("dtype") + b"\x86R")
    assert meta._MetaUnpickler(io.BytesIO(ok)).load() is np.dtype
