"""Drop-in painters: ``CVAEPainter`` / ``CGANPainter`` with the reference ``paint()`` API.

Host-side mirror of reference ``baryon_painter/painter.py`` (``Painter`` ABC :16-30,
``CVAEPainter.__init__`` :34-47, ``paint`` :371-392, ``save_state_to_file`` :395-418,
``load_state_from_file`` :421-445) and of the external ``GAN_Painter`` the reference's
``scripts/create_lightcone.py:41-54`` instantiates.  Same constructor arguments, attributes,
checkpoint files and exceptions; the arithmetic runs in ``libbaryon_painter_b200.so``
(hand-written sm_100a kernels behind the C ABI of ``include/baryon_painter_b200.h``).
PyTorch is only the checkpoint container (``torch.load`` / ``torch.save``).

Extensions over the reference (additive):
  * ``paint(..., latent=, eps=)``  caller-supplied latents:
      ``latent`` IS z (prior network skipped, reference ``sample_P(z=...)``), ``eps`` replaces
      the ``torch.randn`` draw in ``sample_z``.
  * ``paint_batch``   N tiles per call (reference ``paint`` is strictly batch-1).
  * ``paint_variance``  per-pixel mean/variance over latent draws.
"""

import collections
import os

import numpy as np

from . import _lib, arch as _arch, meta as _meta, transforms as _tf

# "fp16": 16-bit operands on the tcgen05 tensor cores (default); "bf16": same kernels, bf16 operands;
# "fp32": fp32-accurate on the same tensor-core kernels (split-precision fp16 operands, fp32 accumulation);
# "fp32-ffma": scalar FFMA kernels (cross-check).  See DESIGN.md "Precision".
DEFAULT_PRECISION = os.environ.get("BARYON_PAINTER_PRECISION", "fp16")


def _device_index(compute_device):
    dev = str(compute_device)
    if dev == "cuda":
        return 0
    if dev.startswith("cuda:"):
        return int(dev.split(":", 1)[1])
    raise ValueError("compute_device %r: this implementation runs on B200 GPUs only ('cuda:N'); "
                     "there is no CPU path" % (compute_device,))


class Painter:
    """Abstract base class for a baryon painter (reference painter.py:16-30)."""

    def __init__(self):
        raise NotImplementedError("This is an abstract base class.")

    def load_state_from_file(self, filename):
        raise NotImplementedError("This is an abstract base class.")

    def paint(self, input, **kwargs):
        raise NotImplementedError("This is an abstract base class.")


class CVAEModel:
    """Inference half of reference ``models.cvae.CVAE`` (cvae.py:8-61, 82-120, 149-162) over a
    ``bp_net``.  Holds the state_dict (all 179 tensors, including the training-only ``q_*``
    sub-networks, so checkpoints round-trip) and the compiled device network."""

    def __init__(self, architecture, device="cuda:0", precision=None, max_batch=16):
        print("CVAE with {} architecture.".format(architecture["type"]))
        self.architecture = architecture
        self.dim_x = tuple(architecture["dim_x"])
        self.dim_y = tuple(architecture["dim_y"])
        self.dim_z = tuple(architecture["dim_z"])
        self.L = architecture["L"] if "L" in architecture else 1
        self.n_x_features = architecture["n_x_features"]
        self.use_aux_label = architecture["aux_label"]
        self.min_z_var = architecture["min_z_var"] if "min_z_var" in architecture else 1e-7
        self.stacks = _arch.cvae_stacks(architecture)
        if "p_var_out" in self.stacks:
            raise NotImplementedError("variance head (len(p_y_z_out) > 1) is not on the fiducial paint path")
        if self.stacks.get("p_y_in"):
            raise NotImplementedError("p_y_in other than None is not supported")
        if not self.use_aux_label:
            raise NotImplementedError("architectures without the redshift aux label are not supported")
        if self.dim_y[0] != 1 or self.dim_x[0] != 1 or self.dim_z[0] != 1:
            raise NotImplementedError("multi-channel dim_x/dim_y/dim_z")
        self.schema = _arch.state_dict_schema(self.stacks)
        self.device = str(device)
        self.precision = precision or DEFAULT_PRECISION
        self.max_batch = int(max_batch)
        self.predict_var = False
        self._state = None
        self._net = None
        self.training = False

    # -- torch.nn.Module look-alikes the reference callers touch ---------------------------
    def train(self, mode=True):
        self.training = False if not mode else self.training
        return self

    def eval(self):
        return self.train(False)

    def state_dict(self):
        if self._state is None:
            raise RuntimeError("model has no weights loaded")
        return collections.OrderedDict(self._state)

    def count_parameters(self):
        return int(sum(int(np.prod(shape)) for k, (shape, dt) in self.schema.items()
                       if dt == "float32" and not k.endswith(("running_mean", "running_var"))))

    def load_state_dict(self, state_dict, strict=True):
        _arch.check_state_dict(self.schema, state_dict)
        self._state = collections.OrderedDict((k, state_dict[k]) for k in self.schema)
        self._build()

    def _build(self):
        if self._net is not None:
            self._net.close()
        folded = {name: _arch.fold_stack(self.stacks[name], self._state)
                  for name in ("prior_network", "p_z_in", "p_y_z_in", "p_mu_out", "q_x_in", "q_y_in", "q_out")
                  if name in self.stacks}
        self._net = _lib.Net.create_cvae(folded, self.dim_y[1:], self.dim_z[1:], self.min_z_var, self.precision,
                                         self.max_batch, _device_index(self.device),
                                         self.architecture.get("likelihood_scaling", 1.0))

    @property
    def net(self):
        if self._net is None:
            raise RuntimeError("model has no weights loaded")
        return self._net

    def ensure_batch(self, n):
        if n > self.max_batch:
            self.max_batch = int(n)
            self._build()

    # -- batched inference on already-transformed inputs ------------------------------------
    def sample_P(self, y, return_var=False, aux_label=None, z=None, eps=None, seed=0):
        """``y`` (N,1,H,W) transformed input, ``aux_label`` (N,) -> x_mu (N,1,H,W) as a torch CPU
        tensor (reference cvae.py:149-162).  ``z`` given = the latent itself."""
        import torch
        y = np.asarray(y.detach().cpu() if hasattr(y, "detach") else y, np.float32)
        n = y.shape[0]
        aux = np.broadcast_to(np.asarray(aux_label.detach().cpu() if hasattr(aux_label, "detach") else aux_label,
                                         np.float32).reshape(-1), (n,))
        self.ensure_batch(n)
        if z is not None:
            mode, lat = _lib.BP_LATENT_GIVEN, np.asarray(z, np.float32).reshape(n, *self.dim_z[1:])
        elif eps is not None:
            mode, lat = _lib.BP_LATENT_EPS, np.asarray(eps, np.float32).reshape(n, *self.dim_z[1:])
        else:
            mode, lat = _lib.BP_LATENT_SEED, None
        out = self.net.cvae_paint_host(y.reshape(n, *self.dim_y[1:]), lat, mode, seed,
                                       (None, None, aux, 1.0, 0.0, 1.0, 0.0), 0)
        return torch.from_numpy(out.reshape(n, 1, *self.dim_y[1:]))

    def forward(self, x, y, aux_label=None, eps=None, seed=0):
        """Evidence lower bound of a batch of already-transformed (x, y) pairs (reference cvae.py:122-147, with the
        recognition network Q :68-80), evaluated on the device.  ``eps`` (N,1,h,w) replaces the ``torch.randn`` draw of
        ``sample_z``.  Sets ``ELBO``, ``KL_term``, ``log_likelihood``, ``z_mu``, ``z_log_var`` like the reference."""
        import torch
        x = np.asarray(x.detach().cpu() if hasattr(x, "detach") else x, np.float32)
        y = np.asarray(y.detach().cpu() if hasattr(y, "detach") else y, np.float32)
        n = y.shape[0]
        aux = np.broadcast_to(np.asarray(aux_label.detach().cpu() if hasattr(aux_label, "detach") else aux_label,
                                         np.float32).reshape(-1), (n,))
        self.ensure_batch(n)
        mode = _lib.BP_LATENT_SEED if eps is None else _lib.BP_LATENT_EPS
        stats, mu, lv = self.net.cvae_elbo_host(x.reshape(n, *self.dim_x[1:]), y.reshape(n, *self.dim_y[1:]),
                                                None if eps is None else np.asarray(eps, np.float32).reshape(n, *self.dim_z[1:]),
                                                mode, seed, (None, None, aux, 1.0, 0.0, 1.0, 0.0), 0)
        self.ELBO, self.KL_term = torch.tensor(stats[0]), torch.tensor(stats[1])
        self.log_likelihood = torch.tensor([stats[2]])
        self.z_mu, self.z_log_var = torch.from_numpy(mu[:, None]), torch.from_numpy(lv[:, None])
        return self.ELBO

    __call__ = forward

    def get_stats(self):
        return (self.ELBO.item(), -self.KL_term.item(), *self.log_likelihood.numpy())

    def get_stats_labels(self):
        return ["ELBO", "KL_term"] + ["log_likelihood_{}".format(i) for i in range(self.n_x_features)]

    def prior(self, y, aux_label=None):
        import torch
        self.sample_P(y, aux_label=aux_label, eps=np.zeros((np.asarray(y).shape[0], *self.dim_z), np.float32))
        mu, lv = self.net.cvae_read_prior(np.asarray(y).shape[0])
        return torch.from_numpy(mu[:, None]), torch.from_numpy(lv[:, None])


class _Ticket:
    """Handle of an enqueued ``paint_batch_async`` batch."""

    def __init__(self, net, slot, out, keep):
        self._net, self._slot, self._out, self._keep = net, slot, out, keep

    def wait(self):
        """block until the batch's painted tiles are in the ``out`` array; returns it"""
        if self._net is not None:
            self._net.wait(self._slot)
            self._net, self._keep = None, None
        return self._out


class CVAEPainter(Painter):
    def __init__(self, filename=None, training_data_set=None, test_data_set=None, architecture="test",
                 compute_device="cuda:0", precision=None, max_batch=16, seed=None):
        self.precision = precision or DEFAULT_PRECISION
        self.max_batch = int(max_batch)
        self._seed = int(seed) if seed is not None else int.from_bytes(os.urandom(8), "little")
        self._calls = 0
        self.transform = None
        self.inverse_transform = None
        self.input_field, self.label_fields = "dm", ["pressure"]
        if filename is not None:
            self.load_state_from_file(filename, compute_device)
        else:
            self.architecture = architecture
            self.compute_device = compute_device
            _device_index(compute_device)
            self.model = CVAEModel(architecture, self.compute_device, self.precision, self.max_batch)
        self.training_data = training_data_set
        self.test_data = test_data_set

    # -- training entry points of the reference are out of scope -------------
    def train(self, *args, **kwargs):
        raise NotImplementedError("training is not part of the B200 paint path; train with the reference "
                                  "and load the resulting (model_state, model_meta) here")

    validate = train

    # -- paint ------------------------------------------------------------------------------------
    def _next_seed(self):
        self._calls += 1
        return (self._seed + 0x9E3779B97F4A7C15 * self._calls) & (2 ** 64 - 1)

    def _sigmas(self, zs, transform, inverse_transform):
        """Per-tile sigma(z) of the fused transforms (interpolated on the host in fp64 like the reference,
        data_transforms.py:52-64) -- once per distinct redshift: a batch is typically one lens plane."""
        s_in = s_out = None
        tp = [1.0, 0.0, 1.0, 0.0]
        uz, inv = np.unique(np.asarray(zs, np.float64), return_inverse=True)
        if transform:
            p = [self.transform.gpu_params(self.input_field, float(z)) for z in uz]
            s_in = np.array([q[1] for q in p], np.float32)[inv]
            tp[0], tp[1] = p[0][2], p[0][3]
        if inverse_transform:
            p = [self.inverse_transform.gpu_params(self.label_fields[0], float(z)) for z in uz]
            s_out = np.array([q[1] for q in p], np.float32)[inv]
            tp[2], tp[3] = p[0][2], p[0][3]
        return s_in, s_out, tp

    def paint_batch(self, tiles, z=0.0, latents=None, eps=None, seed=None, transform=True,
                    inverse_transform=True, out=None):
        """Paint N tiles.  ``tiles`` (N,H,W); ``z`` scalar or (N,).  ``latents`` (N,1,h,w) are used
        as the latent z directly (prior network skipped); ``eps`` (N,1,h,w) replaces the normal draw of
        ``sample_z``; with neither, eps is drawn on the device from ``seed``.
        ``out``: optional float32 (N,H,W) result buffer; with page-locked ``tiles`` / ``out``
        (``baryon_painter_b200.pinned_empty``) the copies are direct DMA transfers overlapped with the kernels.
        Returns float32 (N,H,W) [(N,1,H,W) if ``inverse_transform=False``]."""
        tiles = np.asarray(tiles)
        n = tiles.shape[0]
        if n == 0:
            hw = tuple(self.model.dim_y[1:])
            return np.empty((0, *hw) if inverse_transform else (0, 1, *hw), np.float32)
        zs = np.broadcast_to(np.asarray(z, np.float64).reshape(-1), (n,))
        use_t = bool(transform) and self.transform is not None
        use_i = bool(inverse_transform) and self.inverse_transform is not None
        if use_i and len(self.label_fields) > 1:
            raise NotImplementedError("Painting with more than one output field is not supported yet.")
        fuse_t = use_t and isinstance(self.transform, _tf.CompiledTransform) and \
            self.transform.is_fusable(self.input_field)
        fuse_i = use_i and isinstance(self.inverse_transform, _tf.CompiledTransform) and \
            self.inverse_transform.is_fusable(self.label_fields[0])
        if use_t and not fuse_t:       # caller-installed transform: apply it as the reference would
            tiles = np.stack([np.asarray(self.transform(t, field=self.input_field, z=float(zz)), np.float32)
                              for t, zz in zip(tiles, zs)])
        y_shape = tiles.shape[1:] if tiles.ndim == 4 else (1, *tiles.shape[1:])
        if tuple(y_shape) != tuple(self.model.dim_y):
            raise ValueError(f"Shape mismatch between input and model: {tiles.shape[1:]} vs {self.model.dim_y}")
        tiles = np.ascontiguousarray(tiles.reshape(n, *self.model.dim_y[1:]), np.float32)   # no copy if already so
        if latents is not None and eps is not None:
            raise ValueError("give either latents or eps, not both")
        lat_shape = (n, *self.model.dim_z[1:])
        if latents is not None:
            mode, lat = _lib.BP_LATENT_GIVEN, np.ascontiguousarray(latents, np.float32).reshape(lat_shape)
        elif eps is not None:
            mode, lat = _lib.BP_LATENT_EPS, np.ascontiguousarray(eps, np.float32).reshape(lat_shape)
        else:
            mode, lat = _lib.BP_LATENT_SEED, None
        seed = self._next_seed() if seed is None else int(seed)
        s_in, s_out, tp = self._sigmas(zs, fuse_t, fuse_i)
        flags = (_lib.BP_FLAG_TRANSFORM if fuse_t else 0) | (_lib.BP_FLAG_INVERSE if fuse_i else 0)
        aux = zs.astype(np.float32)
        if out is None:
            out = np.empty((n, *self.model.dim_y[1:]), np.float32)
        elif out.shape != (n, *self.model.dim_y[1:]) or out.dtype != np.float32 or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous float32 array of shape %r" % ((n, *self.model.dim_y[1:]),))
        mb = self.model.max_batch
        for i0 in range(0, n, mb):
            sl = slice(i0, min(n, i0 + mb))
            self.model.net.cvae_paint_host(
                tiles[sl], None if lat is None else lat[sl], mode, seed + i0,
                (None if s_in is None else s_in[sl], None if s_out is None else s_out[sl], aux[sl], *tp), flags,
                out=out[sl])
        if use_i and not fuse_i:
            return np.stack([self.inverse_transform(o.reshape(1, 1, *o.shape), field=self.label_fields[0],
                                                    z=float(zz)) for o, zz in zip(out, zs)])
        if not use_i:
            return out.reshape(n, 1, *out.shape[1:])
        return out

    def paint_batch_async(self, tiles, z=0.0, latents=None, eps=None, seed=None, out=None):
        """``paint_batch`` for a STREAM of batches: enqueues the batch and returns a ticket; ``ticket.wait()`` returns the
        painted tiles.  Consecutive calls take three device I/O slots in turn; with two batches kept outstanding the
        upload of the next batch and the download of the previous one overlap this batch's kernels (which run as whole
        plan chunks)::

            tickets = []
            for batch, result in zip(batches, results):                 # page-locked arrays (pinned_empty)
                tickets.append(painter.paint_batch_async(batch, z=z, eps=eps, out=result))
                if len(tickets) > 2:
                    tickets.pop(0).wait()
            for t in tickets:
                t.wait()

        ``tiles`` (n <= max_batch) and ``out`` must be page-locked float32 arrays (``baryon_painter_b200.pinned_empty``) and
        stay untouched until the wait; the fused fiducial transforms are applied."""
        tiles = np.asarray(tiles)
        n = tiles.shape[0]
        if tuple(tiles.shape[1:]) != tuple(self.model.dim_y[1:]) or tiles.dtype != np.float32 or not tiles.flags.c_contiguous:
            raise ValueError(f"Shape mismatch between input and model: {tiles.shape[1:]} vs {self.model.dim_y} "
                             "(C-contiguous float32 tiles expected)")
        if out is None:
            out = _lib.pinned_empty(tiles.shape)
        elif out.shape != tiles.shape or out.dtype != np.float32 or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous float32 array of shape %r" % (tiles.shape,))
        self.model.ensure_batch(n)
        zs = np.broadcast_to(np.asarray(z, np.float64).reshape(-1), (n,))
        s_in, s_out, tp = self._sigmas(zs, True, True)
        lat_shape = (n, *self.model.dim_z[1:])
        if latents is not None:
            mode, lat = _lib.BP_LATENT_GIVEN, np.ascontiguousarray(latents, np.float32).reshape(lat_shape)
        elif eps is not None:
            mode, lat = _lib.BP_LATENT_EPS, np.ascontiguousarray(eps, np.float32).reshape(lat_shape)
        else:
            mode, lat = _lib.BP_LATENT_SEED, None
        seed = self._next_seed() if seed is None else int(seed)
        slot = getattr(self, "_slot", 0)
        self._slot = (slot + 1) % 3
        keep = self.model.net.cvae_paint_host_async(tiles, lat, mode, seed, (s_in, s_out, zs.astype(np.float32), *tp),
                                                    _lib.BP_FLAG_TRANSFORM | _lib.BP_FLAG_INVERSE, out, slot)
        return _Ticket(self.model.net, slot, out, keep)

    def paint_batch_device(self, tiles, z=0.0, eps=None, latents=None, seed=None, out=None):
        """Device-resident variant of ``paint_batch`` for callers that keep tiles on the GPU (the lightcone
        loop): ``tiles`` float32 CUDA tensor (N,H,W) on this painter's device, result in ``out`` (allocated if
        None).  Transforms are the fused fiducial ones."""
        import torch
        n = int(tiles.shape[0])
        if tuple(tiles.shape[1:]) != tuple(self.model.dim_y[1:]):
            raise ValueError(f"Shape mismatch between input and model: {tuple(tiles.shape[1:])} vs {self.model.dim_y}")
        if not (tiles.is_cuda and tiles.dtype == torch.float32 and tiles.is_contiguous()):
            raise ValueError("tiles must be a contiguous float32 CUDA tensor")
        if out is None:
            out = torch.empty_like(tiles)
        fusable = isinstance(self.transform, _tf.CompiledTransform) and self.transform.is_fusable(self.input_field) and \
            isinstance(self.inverse_transform, _tf.CompiledTransform) and \
            self.inverse_transform.is_fusable(self.label_fields[0]) and len(self.label_fields) == 1
        if not fusable:
            # no transforms in the checkpoint, or caller-installed ones the kernels cannot fuse: same semantics as
            # paint_batch (the caller's callables applied on the host as the reference would), staged through the host
            res = self.paint_batch(tiles.cpu().numpy(), z=z,
                                   latents=None if latents is None else latents.cpu().numpy(),
                                   eps=None if eps is None else eps.cpu().numpy(), seed=seed)
            out.copy_(torch.from_numpy(np.ascontiguousarray(res, np.float32).reshape(out.shape)))
            return out
        zs = np.broadcast_to(np.asarray(z, np.float64).reshape(-1), (n,))
        s_in, s_out, tp = self._sigmas(zs, True, True)
        flags = _lib.BP_FLAG_TRANSFORM | _lib.BP_FLAG_INVERSE
        aux = zs.astype(np.float32)
        if latents is not None:
            mode, lat = _lib.BP_LATENT_GIVEN, latents
        elif eps is not None:
            mode, lat = _lib.BP_LATENT_EPS, eps
        else:
            mode, lat = _lib.BP_LATENT_SEED, None
        if lat is not None and not (lat.is_cuda and lat.dtype == torch.float32 and lat.is_contiguous()):
            raise ValueError("latents / eps must be contiguous float32 CUDA tensors")
        seed = self._next_seed() if seed is None else int(seed)
        stream = torch.cuda.current_stream(tiles.device).cuda_stream
        mb = self.model.max_batch
        for i0 in range(0, n, mb):
            i1 = min(n, i0 + mb)
            self.model.net.cvae_paint_device(tiles[i0:i1].data_ptr(), 0 if lat is None else lat[i0:i1].data_ptr(), mode,
                                             seed + i0, (s_in[i0:i1], s_out[i0:i1], aux[i0:i1], *tp), flags,
                                             out[i0:i1].data_ptr(), i1 - i0, stream)
        return out

    def paint(self, input, z=0.0, transform=True, inverse_transform=True, latent=None, eps=None, seed=None):
        """Reference ``paint`` (painter.py:371-392): one tile (H,W) [or (1,H,W) with
        ``transform=False``] at redshift ``z`` -> float32 (H,W) [(1,1,H,W) if
        ``inverse_transform=False``]."""
        self.model.train(False)
        input = np.asarray(input)
        use_t = bool(transform) and self.transform is not None
        y_shape = input.shape if (input.ndim != 2 or not use_t) else (1, *input.shape)   # atleast_3d
        if tuple(y_shape) != tuple(self.model.dim_y):
            raise ValueError(f"Shape mismatch between input and model: {input.shape} vs {self.model.dim_y}")
        out = self.paint_batch(input.reshape(1, *input.shape[-2:]), z=z,
                               latents=None if latent is None else np.asarray(latent).reshape(1, *self.model.dim_z),
                               eps=None if eps is None else np.asarray(eps).reshape(1, *self.model.dim_z),
                               seed=seed, transform=transform, inverse_transform=inverse_transform)
        if inverse_transform and self.inverse_transform is not None:
            return out[0]
        return out.reshape(1, 1, *out.shape[-2:])

    def elbo(self, x_tiles, y_tiles, z=0.0, eps=None, seed=0):
        """Evidence lower bound of raw (pressure, dark matter) tile pairs at redshift(s) ``z``: the dataset transforms
        of both fields fused on the device, then ``CVAE.forward`` (what the reference's ``validate`` evaluates per
        batch, painter.py:295-367).  Returns ``(ELBO, KL_term, log_likelihood)``."""
        x_tiles, y_tiles = np.asarray(x_tiles, np.float32), np.asarray(y_tiles, np.float32)
        n = y_tiles.shape[0]
        zs = np.broadcast_to(np.asarray(z, np.float64).reshape(-1), (n,))
        # forward transforms of BOTH fields: sigma_in for the input field, sigma_out for the label field
        p_in = [self.transform.gpu_params(self.input_field, float(v)) for v in zs]
        p_x = [self.transform.gpu_params(self.label_fields[0], float(v)) for v in zs]
        s_in = np.array([q[1] for q in p_in], np.float32)
        s_x = np.array([q[1] for q in p_x], np.float32)
        self.model.ensure_batch(n)
        mode = _lib.BP_LATENT_SEED if eps is None else _lib.BP_LATENT_EPS
        stats, mu, lv = self.model.net.cvae_elbo_host(
            x_tiles, y_tiles, None if eps is None else np.asarray(eps, np.float32).reshape(n, *self.model.dim_z[1:]), mode, seed,
            (s_in, s_x, zs.astype(np.float32), p_in[0][2], p_in[0][3], p_x[0][2], p_x[0][3]), _lib.BP_FLAG_TRANSFORM)
        return float(stats[0]), float(stats[1]), float(stats[2])

    def paint_variance(self, tiles, z=0.0, n_draws=64, seed=0):
        """Per-pixel mean and variance of the painted pressure over ``n_draws`` latent draws per
        tile (BASELINE config 4).  Returns (mean, var), float32 (N,H,W) each."""
        tiles = np.ascontiguousarray(tiles, np.float32)
        n = tiles.shape[0]
        if tuple(tiles.shape[1:]) != tuple(self.model.dim_y[1:]):
            raise ValueError(f"Shape mismatch between input and model: {tiles.shape[1:]} vs {self.model.dim_y}")
        zs = np.broadcast_to(np.asarray(z, np.float64).reshape(-1), (n,))
        s_in, s_out, tp = self._sigmas(zs, True, True)
        mean = np.empty_like(tiles)
        var = np.empty_like(tiles)
        mb = self.model.max_batch
        for i0 in range(0, n, mb):
            sl = slice(i0, min(n, i0 + mb))
            mean[sl], var[sl] = self.model.net.cvae_paint_variance_host(
                tiles[sl], (s_in[sl], s_out[sl], zs[sl].astype(np.float32), *tp), n_draws, seed + i0)
        return mean, var

    # -- checkpoints --------------------------------------------------------------------------------
    def save_state_to_file(self, filename, mode="model_state_dict+metadata"):
        import torch
        if not isinstance(filename, (tuple, list)):
            raise ValueError("filename needs to be a tuple of (state_filename, meta_filename).")
        d = {k: getattr(self, k) for k in ("L", "n_grid", "tile_L", "n_tile", "tile_size", "input_field",
                                           "label_fields", "scale_to_SLICS")}
        d["transform"] = self.transform
        d["inverse_transform"] = self.inverse_transform
        d["model_architecture"] = self.architecture
        _meta.write_model_meta(filename[1], d)
        torch.save(self.model.state_dict(), filename[0])

    def load_state_from_file(self, filename, compute_device="cuda:0"):
        import torch
        if not isinstance(filename, (tuple, list)):
            raise ValueError("filename needs to be a tuple of (state_filename, meta_filename).")
        self.compute_device = compute_device
        _device_index(compute_device)
        state_dict = torch.load(filename[0], map_location="cpu")
        d = _meta.read_model_meta(filename[1])
        self.model = CVAEModel(d["model_architecture"], self.compute_device, self.precision, self.max_batch)
        self.model.load_state_dict(state_dict)
        self.architecture = d["model_architecture"]
        self.L = d["L"]
        self.n_grid = d["n_grid"]
        self.tile_L = d["tile_L"]
        self.n_tile = d["n_tile"]
        self.tile_size = d["tile_size"]
        self.input_field = d["input_field"]
        self.label_fields = d["label_fields"]
        self.scale_to_SLICS = d["scale_to_SLICS"]
        self.transform = d["transform"] if "transform" in d else None
        self.inverse_transform = d["inverse_transform"] if "inverse_transform" in d else None

    @classmethod
    def synthetic(cls, tile_size=512, seed=0, compute_device="cuda:0", precision=None, max_batch=16):
        """Painter with the fiducial architecture, seeded synthetic weights and the fiducial
        transforms (the shipped weights are not in the reference checkout)."""
        from . import synthetic as _syn
        A = _arch.fiducial_cvae_architecture(tile_size)
        p = cls(architecture=A, compute_device=compute_device, precision=precision, max_batch=max_batch)
        p.model.load_state_dict(_syn.synthetic_cvae_state_dict(A, seed=seed))
        p.transform, p.inverse_transform = _tf.fiducial_transforms("cvae")
        p.L, p.n_grid, p.tile_L, p.n_tile, p.tile_size, p.scale_to_SLICS = 400, 2048, 100.0, 4, tile_size, True
        return p


class CGANModel:
    def __init__(self, layers, tile_hw, device, precision, max_batch):
        self.layers = layers
        self.specs = _arch.flatten_stack(layers, "generator")
        self.schema = _arch.state_dict_schema(collections.OrderedDict(generator=self.specs))
        self.tile_hw, self.device = tuple(tile_hw), str(device)
        self.precision, self.max_batch = precision or DEFAULT_PRECISION, int(max_batch)
        self.dim_y = (1, *self.tile_hw)
        self._state, self._net = None, None

    def state_dict(self):
        return collections.OrderedDict(self._state)

    def load_state_dict(self, state_dict):
        state_dict = fold_spectral_norm(state_dict)
        _arch.check_state_dict(self.schema, state_dict)
        self._state = collections.OrderedDict((k, state_dict[k]) for k in self.schema)
        if self._net is not None:
            self._net.close()
        self._net = _lib.Net.create_cgan(_arch.fold_stack(self.specs, self._state), self.tile_hw, self.precision,
                                         self.max_batch, _device_index(self.device))

    @property
    def net(self):
        if self._net is None:
            raise RuntimeError("model has no weights loaded")
        return self._net


def fold_spectral_norm(state_dict):
    """``torch.nn.utils.spectral_norm`` checkpoints store ``weight_orig`` / ``weight_u`` (/``weight_v``);
    the inference weight is ``weight_orig / sigma`` with ``sigma = u^T W v`` (SURVEY.md App. C iii)."""
    import torch
    out = collections.OrderedDict()
    for k, v in state_dict.items():
        if k.endswith("weight_orig"):
            base = k[:-len("weight_orig")]
            w = v.detach().cpu().to(torch.float64)
            wm = w.reshape(w.shape[0], -1)
            u = state_dict[base + "weight_u"].detach().cpu().to(torch.float64)
            if base + "weight_v" in state_dict:
                vv = state_dict[base + "weight_v"].detach().cpu().to(torch.float64)
            else:
                vv = torch.nn.functional.normalize(wm.t() @ u, dim=0, eps=1e-12)
            sigma = torch.dot(u, wm @ vv)
            out[base + "weight"] = (w / sigma).to(torch.float32)
        elif k.endswith(("weight_u", "weight_v")):
            continue
        else:
            out[k] = v
    return out


class CGANPainter(Painter):
    """Drop-in for the external ``GAN_Painter(parts_folder, checkpoint_file=, device=)`` the
    reference drives through ``paint(input, z=, transform=, inverse_transform=)``
    (scripts/create_lightcone.py:41-54; notebooks/validation_plots.ipynb cell 17).

    ``parts_folder`` holds ``transform.pickle`` / ``inv_transform.pickle`` / ``z_transform.pickle``
    (and ``g_struc.pickle``, whose layer table is restated in ``arch.fiducial_cgan_architecture``);
    ``checkpoint_file`` is a ``torch.save``d generator state_dict (optionally nested under
    ``"generator"`` / ``"g"`` / ``"state_dict"``; spectral-norm parametrisation is folded).
    The PainterGAN sources and the trained ``.cp`` are not part of the reference checkout, so this
    path is validated against the restated oracle only (CGAN parity unpinned)."""

    def __init__(self, parts_folder=None, checkpoint_file=None, device="cuda:0", precision=None, max_batch=16,
                 tile_size=512, layers=None, state_dict=None):
        self.compute_device = device
        _device_index(device)
        self.input_field, self.label_fields = "dm", ["pressure"]
        self.z_shift = 1.0
        self.transform, self.inverse_transform = _tf.fiducial_transforms("cgan")
        if parts_folder is not None:
            for key, fn in (("transform", "transform.pickle"), ("inverse_transform", "inv_transform.pickle")):
                path = os.path.join(parts_folder, fn)
                if os.path.exists(path):
                    setattr(self, key, _meta._rebind(_load_pickled_fn(path)))
        self.model = CGANModel(layers or _arch.fiducial_cgan_architecture(), (tile_size, tile_size), device,
                               precision, max_batch)
        if checkpoint_file is not None:
            import torch
            state_dict = torch.load(checkpoint_file, map_location="cpu")
            for key in ("generator", "g", "G", "state_dict", "model"):
                if isinstance(state_dict, dict) and key in state_dict and isinstance(state_dict[key], dict):
                    state_dict = state_dict[key]
        if state_dict is not None:
            self.model.load_state_dict(state_dict)

    def paint_batch(self, tiles, z=0.0, transform=True, inverse_transform=True, out=None):
        """Paint N tiles; ``z`` scalar or (N,).  ``out``: optional float32 (N,H,W) result buffer (page-locked
        ``tiles`` / ``out`` are transferred by DMA, overlapped with the kernels)."""
        tiles = np.asarray(tiles)
        n = tiles.shape[0]
        if tuple(tiles.shape[-2:]) != self.model.tile_hw:
            raise ValueError(f"Shape mismatch between input and model: {tiles.shape[1:]} vs {self.model.dim_y}")
        zs = np.broadcast_to(np.asarray(z, np.float64).reshape(-1), (n,))
        tiles = np.ascontiguousarray(tiles.reshape(n, *self.model.tile_hw), np.float32)
        tp = [1.0, 0.0, 1.0, 0.0]
        s_in = s_out = None
        uz, inv = np.unique(zs, return_inverse=True)          # sigma(z) once per distinct redshift
        if transform:
            p = [self.transform.gpu_params(self.input_field, float(zz)) for zz in uz]
            s_in, tp[0], tp[1] = np.array([q[1] for q in p], np.float32)[inv], p[0][2], p[0][3]
        if inverse_transform:
            p = [self.inverse_transform.gpu_params(self.label_fields[0], float(zz)) for zz in uz]
            s_out, tp[2], tp[3] = np.array([q[1] for q in p], np.float32)[inv], p[0][2], p[0][3]
        flags = (_lib.BP_FLAG_TRANSFORM if transform else 0) | (_lib.BP_FLAG_INVERSE if inverse_transform else 0)
        aux = (zs - self.z_shift).astype(np.float32)
        if out is None:
            out = np.empty((n, *self.model.tile_hw), np.float32)
        elif out.shape != (n, *self.model.tile_hw) or out.dtype != np.float32 or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous float32 array of shape %r" % ((n, *self.model.tile_hw),))
        mb = self.model.max_batch
        for i0 in range(0, n, mb):
            sl = slice(i0, min(n, i0 + mb))
            self.model.net.cgan_paint_host(
                tiles[sl], (None if s_in is None else s_in[sl], None if s_out is None else s_out[sl], aux[sl], *tp),
                flags, out=out[sl])
        return out if inverse_transform else out.reshape(n, 1, *out.shape[1:])

    def paint(self, input, z=0.0, transform=True, inverse_transform=True):
        input = np.asarray(input)
        out = self.paint_batch(input.reshape(1, *input.shape[-2:]), z=z, transform=transform,
                               inverse_transform=inverse_transform)
        return out[0] if inverse_transform else out.reshape(1, 1, *out.shape[-2:])

    @classmethod
    def synthetic(cls, tile_size=512, seed=0, device="cuda:0", precision=None, max_batch=16, n_res_blocks=9):
        from . import synthetic as _syn
        layers = _arch.fiducial_cgan_architecture(n_res_blocks)
        return cls(device=device, precision=precision, max_batch=max_batch, tile_size=tile_size, layers=layers,
                   state_dict=_syn.synthetic_cgan_state_dict(layers, seed=seed))


def _load_pickled_fn(path):
    import io
    with open(path, "rb") as f:
        return _meta._MetaUnpickler(io.BytesIO(f.read())).load()
