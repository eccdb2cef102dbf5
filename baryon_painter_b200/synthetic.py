"""Seeded synthetic checkpoints and input tiles.

The trained blobs (``trained_models/CVAE/fiducial/model_state`` and the CGAN ``.cp``)
are not part of the reference checkout (``.MISSING_LARGE_BLOBS``), so benchmarks and
parity tests run on a seeded ``state_dict`` with the identical key schema
(SURVEY.md App. D).  BatchNorm running statistics, affine parameters and PReLU slopes
are randomised -- a freshly initialised BatchNorm is the identity and would hide
folding bugs -- and convolution weights use a fan-in scaling so that activations stay
O(1) through the 25 layers and the painted pressure has a sane dynamic range.

Only numpy's ``default_rng`` is used, so the same dict is produced on every box; torch
is the container (``torch.save`` -> ``model_state``).
"""

import collections

import numpy as np

from . import arch as _arch


def _fan_in(spec):
    if spec.kind == "conv":
        return spec.cin * spec.k * spec.k
    return spec.cin * (spec.k // spec.stride) ** 2 if spec.k % spec.stride == 0 else \
        spec.cin * spec.k * spec.k / float(spec.stride ** 2)


def synthetic_state_arrays(stacks, seed=0, gain=1.0):
    """OrderedDict key -> numpy array following ``arch.state_dict_schema`` order."""
    rng = np.random.default_rng(seed)
    out = collections.OrderedDict()
    for name, specs in stacks.items():
        for s in specs:
            shape = (s.cout, s.cin, s.k, s.k) if s.kind == "conv" else (s.cin, s.cout, s.k, s.k)
            std = gain * np.sqrt(2.0 / _fan_in(s))
            if not s.bn_prefix and s.cout == 1:
                # un-normalised output heads: keep the network output x_mu inside the trained model's range
                # (x_mu < 3, i.e. p / sigma_p < e^12): the inverse transform exp(4 x) turns an absolute error dx
                # into a relative error 4 dx of the painted pressure, so a single x_mu > 3.5 pixel would carry a
                # tile's whole L2 norm (measured with std *= 0.5: x_mu up to 4.2; with 0.35: <= 2.6 on all fixtures)
                std *= 0.35
            out[s.w_key] = (rng.standard_normal(shape) * std).astype(np.float32)
            if s.b_key:
                out[s.b_key] = (rng.standard_normal(s.cout) * 0.05).astype(np.float32)
            if s.bn_prefix:
                if s.res == _arch.RES_CLOSE:     # keep the residual trunk from blowing up
                    gamma, beta = rng.uniform(0.2, 0.5, s.cout), -0.15 + 0.1 * rng.standard_normal(s.cout)
                elif s.cout <= 2:                # (mu, log-var) heads: keep both alive after the ReLU
                    gamma, beta = rng.uniform(0.7, 1.3, s.cout), rng.uniform(0.2, 0.6, s.cout)
                else:
                    gamma, beta = rng.uniform(0.7, 1.3, s.cout), 0.1 + 0.2 * rng.standard_normal(s.cout)
                out[s.bn_prefix + ".weight"] = gamma.astype(np.float32)
                out[s.bn_prefix + ".bias"] = beta.astype(np.float32)
                out[s.bn_prefix + ".running_mean"] = (0.2 * rng.standard_normal(s.cout)).astype(np.float32)
                out[s.bn_prefix + ".running_var"] = rng.uniform(0.6, 1.6, s.cout).astype(np.float32)
                out[s.bn_prefix + ".num_batches_tracked"] = np.array(1000 + len(out), dtype=np.int64)
            if s.act_key:
                out[s.act_key] = np.array([rng.uniform(0.1, 0.4)], dtype=np.float32)
    return out


def synthetic_cvae_state_dict(architecture=None, seed=0):
    """Seeded CVAE ``state_dict`` (torch tensors) for ``architecture`` (default: fiducial)."""
    import torch
    if architecture is None:
        architecture = _arch.fiducial_cvae_architecture()
    arrays = synthetic_state_arrays(_arch.cvae_stacks(architecture), seed=seed)
    return collections.OrderedDict((k, torch.from_numpy(v.copy()) if v.ndim else torch.tensor(int(v)))
                                   for k, v in arrays.items())


def synthetic_cgan_state_dict(layers=None, seed=0):
    import torch
    if layers is None:
        layers = _arch.fiducial_cgan_architecture()
    stacks = collections.OrderedDict(generator=_arch.flatten_stack(layers, "generator"))
    arrays = synthetic_state_arrays(stacks, seed=seed)
    return collections.OrderedDict((k, torch.from_numpy(v.copy()) if v.ndim else torch.tensor(int(v)))
                                   for k, v in arrays.items())


def synthetic_dm_tiles(n, tile_size=512, seed0=0):
    """Positive DM tiles, one ``default_rng(seed0+i).lognormal(-0.5, 1.0)`` field per tile
    (mean ~1 like ``scale_to_SLICS`` densities)."""
    out = np.empty((n, tile_size, tile_size), np.float32)
    for i in range(n):
        out[i] = np.random.default_rng(seed0 + i).lognormal(-0.5, 1.0, (tile_size, tile_size))
    return out


def synthetic_latents(n, latent_hw=(16, 16), seed=1):
    """``default_rng(seed).standard_normal((n,1,h,w))`` float32."""
    return np.random.default_rng(seed).standard_normal((n, 1, *latent_hw)).astype(np.float32)
