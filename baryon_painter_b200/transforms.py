"""Range-compress data transforms of the paint path (host-side description).

Mirrors reference ``baryon_painter/utils/data_transforms.py``:
``create_range_compress_transforms`` (:51-110) with ``interpolate_z`` (:52-64),
``chain_transformations`` (:44-49), ``atleast_3d`` (:112-116), ``squeeze``
(:118-119) and the binder ``datasets.compile_transform`` (``datasets.py:8-13``).

A :class:`CompiledTransform` is what ``painter.transform`` /
``painter.inverse_transform`` hold.  It is callable with the reference signature
``f(x, field=None, z=None)`` (numpy, used when a caller applies the transform
by hand) and it exposes ``gpu_params(field, z)`` so the painter can hand the
elementwise formula to the CUDA kernels, which fuse it into the first and last
passes over the tile instead of running it on the host.

Only the modes the two fiducial models use are implemented on the GPU:
``shift-log`` (CVAE) and ``shift-log-cam`` (CGAN pickles; decoded in SURVEY.md
App. C).  Other modes raise ``NotImplementedError`` at ``gpu_params`` time.
"""

import collections

import numpy as np

GPU_MODES = ("shift-log", "shift-log-cam")


def normalise_stats(stats):
    """``stats[field][z] = {'mean','var'}`` -> OrderedDict keyed by float z (sorted as stored)."""
    out = collections.OrderedDict()
    for field, per_z in stats.items():
        out[field] = collections.OrderedDict()
        for z, s in per_z.items():
            out[field][float(z)] = {"mean": s["mean"], "var": s["var"]}
    return out


def interpolate_z(stats_field, z):
    """Reference ``interpolate_z`` (data_transforms.py:52-64), arithmetic unchanged."""
    z_list = list(stats_field.keys())
    idx = np.searchsorted(z_list, z, side="right")
    if idx >= len(z_list):
        return stats_field[z_list[-1]]
    elif idx <= 0:
        return stats_field[z_list[0]]
    w = (z - z_list[idx - 1]) / (z_list[idx] - z_list[idx - 1])
    names = stats_field[z_list[0]].keys()
    return {s: w * stats_field[z_list[idx]][s] + (1 - w) * stats_field[z_list[idx - 1]][s]
            for s in names}


class CompiledTransform:
    """Bound transform ``f(x, field=None, z=None)`` (reference datasets.py:13)."""

    def __init__(self, stats, k_values, modes, eps=1e-4, sqrt_of_mean=False, inverse=False,
                 steps=None, field=None, z=None):
        self.stats = normalise_stats(stats)
        self.k_values = dict(k_values)
        self.modes = dict(modes)
        self.eps = float(eps)
        self.sqrt_of_mean = bool(sqrt_of_mean)
        self.inverse = bool(inverse)
        if steps is None:
            steps = ("squeeze", "inv_transform") if inverse else ("transform", "atleast_3d")
        self.steps = tuple(steps)
        self.field, self.z = field, z

    # -- parameters the CUDA kernels need ------------------------------------------------
    def sigma(self, field, z):
        """sqrt(var interpolated in z); float64 on the host like the reference."""
        return float(np.sqrt(np.float64(interpolate_z(self.stats[field], z)["var"])))

    def gpu_params(self, field, z):
        """-> (mode_id, sigma, k0, k1) for the fused elementwise kernels.

        mode 0 ``shift-log``:      fwd ln(x/s+1)/k0        inv (exp(k0*x)-1)*s
        mode 1 ``shift-log-cam``:  fwd ln(x/s+1)/k0 - k1   inv (exp((x+k1)*k0)-1)*s
        """
        mode = self.modes[field].lower()
        if mode not in GPU_MODES:
            raise NotImplementedError("transform mode %r has no CUDA kernel" % (mode,))
        k = self.k_values[field]
        if mode == "shift-log":
            return 0, self.sigma(field, z), float(k), 0.0
        return 1, self.sigma(field, z), float(k[0]), float(k[1])

    def is_fusable(self, field):
        want = ("squeeze", "inv_transform") if self.inverse else ("transform", "atleast_3d")
        return (self.steps == want and field in self.modes and
                self.modes[field].lower() in GPU_MODES and not self.sqrt_of_mean)

    # -- numpy evaluation (reference arithmetic) -----------------------------------------
    def _range_compress(self, x, field, z):
        k = self.k_values[field]
        mode = self.modes[field].lower()
        std = np.sqrt(interpolate_z(self.stats[field], z)["var"])
        if not self.inverse:
            if mode == "shift-log":
                return np.log(x / std + 1) / k
            if mode == "shift-log-cam":
                return np.log(x / std + 1) / k[0] - k[1]
        else:
            if mode == "shift-log":
                return (np.exp(x * k) - 1) * std
            if mode == "shift-log-cam":
                return (np.exp((x + k[1]) * k[0]) - 1) * std
        raise ValueError(f"Mode '{mode}' not supported.")

    def __call__(self, x, field=None, z=None):
        field = self.field if field is None else field
        z = self.z if z is None else z
        for step in self.steps:
            if step in ("transform", "inv_transform"):
                x = self._range_compress(x, field, z)
            elif step == "atleast_3d":
                x = x.reshape(1, *x.shape) if x.ndim == 2 else x
            elif step == "squeeze":
                x = x.squeeze()
            else:
                raise ValueError("unknown transform step %r" % (step,))
        if not self.inverse and isinstance(x, np.ndarray) and x.dtype == np.float64:
            # NumPy>=2 promotes float32/np.float64 to float64; under the authors' NumPy 1.x
            # the forward transform stayed float32 and the model requires it.
            x = x.astype(np.float32)
        return x

    def __repr__(self):
        return "CompiledTransform(%s, modes=%r, k=%r)" % (
            "inverse" if self.inverse else "forward", self.modes, self.k_values)


# Statistics of the fiducial BAHAMAS training set as stored in
# trained_models/CVAE/fiducial/model_meta (extracted with
# meta.read_model_meta by oracle/make_golden.py and cross-checked by
# tests/test_meta.py against tests/golden/fiducial_meta.json).
FIDUCIAL_Z = (0.0, 0.125, 0.25, 0.375, 0.5, 0.75, 1.0, 1.25, 1.5, 1.75, 2.0)
FIDUCIAL_DM = (
    (1.0017759225706175, 1.4725093809115477), (1.001683667841899, 1.1928380647223897),
    (1.0015942710663628, 0.9748087314972294), (1.0015036167914264, 0.8048288134017273),
    (1.0014201645585088, 0.6690492139014439), (1.0012747519006033, 0.47435083706743403),
    (1.001140656737299, 0.345349378108309), (1.0010352554239428, 0.25663857441187393),
    (1.000943229331479, 0.19418252392874455), (1.000864692778035, 0.14947816356834498),
    (1.0007993028092281, 0.11647592444540457))
FIDUCIAL_PRESSURE = (
    (0.04423535, 0.13492714), (0.041155286, 0.10697187), (0.037526328, 0.06813702),
    (0.033997055, 0.04863641), (0.030573525, 0.028984208), (0.024689011, 0.015448382),
    (0.019772898, 0.006693993), (0.015634593, 0.0030250712), (0.01233014, 0.0014460934),
    (0.009684066, 0.0007323309), (0.00752851, 0.0003842623))


def fiducial_stats():
    """The 11-redshift stats table; dm values float64, pressure values float32 (App. B)."""
    stats = collections.OrderedDict(dm=collections.OrderedDict(), pressure=collections.OrderedDict())
    for z, (m, v) in zip(FIDUCIAL_Z, FIDUCIAL_DM):
        stats["dm"][z] = {"mean": np.float64(m), "var": np.float64(v)}
    for z, (m, v) in zip(FIDUCIAL_Z, FIDUCIAL_PRESSURE):
        stats["pressure"][z] = {"mean": np.float32(m), "var": np.float32(v)}
    return stats


def fiducial_transforms(model="cvae"):
    """(transform, inverse_transform) of the fiducial CVAE (``CVAE_single_scale.py:34-65``)
    or CGAN (``trained_models/CGAN/fiducial/{transform,inv_transform}.pickle``)."""
    if model == "cvae":
        kw = dict(k_values={"dm": 4.0, "pressure": 4},
                  modes={"dm": "shift-log", "pressure": "shift-log"}, eps=1e-4)
    elif model == "cgan":
        kw = dict(k_values={"dm": [4.0, 1.0], "pressure": [4.0, 1.0]},
                  modes={"dm": "shift-log-cam", "pressure": "shift-log-cam"}, eps=1e-4)
    else:
        raise ValueError(model)
    stats = fiducial_stats()
    return (CompiledTransform(stats, inverse=False, **kw),
            CompiledTransform(stats, inverse=True, **kw))
