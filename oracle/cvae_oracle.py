"""TEST INFRASTRUCTURE -- CPU restatement of the reference paint path (the oracle).

This file is the checker, never the product: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  It restates,
with ``torch.nn.functional`` on the CPU (the same arithmetic library the reference's
``torch.nn`` modules dispatch to) and numpy, exactly what the reference computes:

* ``build_sequential``       reference ``baryon_painter/models/utils.py:114-157``
* ``ResidualBlock.forward``  reference ``models/utils.py:35-38``
* ``merge_aux_label``        reference ``models/utils.py:159-182``
* ``CVAE.prior/sample_z/sample_prior/P/sample_P``
                             reference ``models/cvae.py:82-95, 63-66, 97-100, 103-120, 149-162``
* ``CVAEPainter.paint``      reference ``painter.py:371-392``
* shift-log transforms + ``interpolate_z``
                             reference ``utils/data_transforms.py:52-64, 66-76, 88-98, 112-119``

Parity pin: ``oracle/make_golden.py`` runs the *unmodified* reference (imported from
/root/reference through ``oracle/ref_shims.py``) and this restatement on the same seeded
weights/tiles/latents; ``tests/test_oracle.py`` asserts they agree bit-for-bit here and
checks the committed ``tests/golden/*.npz`` produced by that script.  The reference's own
test-suite holds no golden vector for this path, so the reference run
itself is the pin.

The CGAN generator below has no reference implementation to pin against (external
PainterGAN): **CGAN parity unpinned**.
"""

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# transforms (numpy, reference arithmetic)
# --------------------------------------------------------------------------------------

def interpolate_z(stats_field, z):
    z_list = list(stats_field.keys())
    idx = np.searchsorted(z_list, z, side="right")
    if idx >= len(z_list):
        return stats_field[z_list[-1]]
    elif idx <= 0:
        return stats_field[z_list[0]]
    w = (z - z_list[idx - 1]) / (z_list[idx] - z_list[idx - 1])
    return {s: w * stats_field[z_list[idx]][s] + (1 - w) * stats_field[z_list[idx - 1]][s]
            for s in stats_field[z_list[0]].keys()}


def forward_transform(x, z, stats, k=4.0, field="dm", shift=0.0):
    """``ln(x/sigma + 1)/k [- shift]`` then atleast_3d; cast to float32."""
    std = np.sqrt(interpolate_z(stats[field], z)["var"])
    y = np.log(x / std + 1) / k - shift if shift else np.log(x / std + 1) / k
    if y.ndim == 2:
        y = y.reshape(1, *y.shape)
    return np.asarray(y, dtype=np.float32)


def inverse_transform(x, z, stats, k=4, field="pressure", shift=0.0):
    """squeeze then ``(exp((x [+ shift]) * k) - 1) * sigma``."""
    std = np.sqrt(interpolate_z(stats[field], z)["var"])
    x = x.squeeze()
    if shift:
        return (np.exp((x + shift) * k) - 1) * std
    return (np.exp(x * k) - 1) * std


# --------------------------------------------------------------------------------------
# layer-list interpreter
# --------------------------------------------------------------------------------------

def run_sequential(layers, sd, prefix, x, taps=None):
    """Apply the tagged layer list to ``x`` with weights ``sd[prefix.idx.*]`` (eval mode).
    ``taps`` (optional list) receives ``(key, tensor)`` after every top-level entry."""
    if layers is None:
        return x
    for idx, layer in enumerate(layers):
        name = layer[0].lower()
        cfg = layer[1] if len(layer) == 2 else None
        key = "%s.%d" % (prefix, idx)
        if name == "conv":
            x = F.conv2d(x, sd[key + ".weight"], sd.get(key + ".bias") if cfg.get("bias", True) else None,
                         stride=cfg.get("stride", 1), padding=cfg.get("padding", 0))
        elif name == "transp conv":
            x = F.conv_transpose2d(x, sd[key + ".weight"],
                                   sd.get(key + ".bias") if cfg.get("bias", True) else None,
                                   stride=cfg.get("stride", 1), padding=cfg.get("padding", 0),
                                   output_padding=cfg.get("output_padding", 0))
        elif name == "batchnorm":
            x = F.batch_norm(x, sd[key + ".running_mean"], sd[key + ".running_var"],
                             sd[key + ".weight"], sd[key + ".bias"], training=False,
                             momentum=0.1, eps=cfg.get("eps", 1e-5))
        elif name == "relu":
            x = F.relu(x)
        elif name == "leaky relu":
            x = F.leaky_relu(x, cfg)
        elif name == "prelu":
            x = F.prelu(x, sd[key + ".weight"])
        elif name == "tanh":
            x = torch.tanh(x)
        elif name == "sigmoid":
            x = torch.sigmoid(x)
        elif name == "softplus":
            x = F.softplus(x)                      # beta=1, threshold=20
        elif name == "residual block":
            inner, act = cfg
            h = run_sequential(inner, sd, key + ".res_block", x)
            x = torch.add(h, x)
            a = act[0].lower() if act[0] is not None else None
            if a == "relu":
                x = F.relu(x)
            elif a == "leaky relu":
                x = F.leaky_relu(x, act[1])
            elif a is not None:
                raise NotImplementedError(act[0])
        elif name == "flatten":
            x = x.view(x.size(0), -1)
        elif name == "unflatten":
            x = x.view(x.size(0), *cfg)
        else:
            raise NotImplementedError(name)
        if taps is not None:
            taps.append((key, x))
    return x


def merge_aux_label(y, aux_label):
    if aux_label.dim() == 0 or aux_label.dim() == 1:
        aux_label = aux_label.reshape(-1, 1)
    if aux_label.shape[0] != y.shape[0]:
        raise ValueError("aux_label batch size needs to match that of y")
    aux = aux_label.reshape(*aux_label.shape, 1, 1).expand((*aux_label.shape, *y.shape[-2:]))
    return torch.cat((y, aux), dim=1)


class CVAEOracle:
    """Inference half of the reference CVAE on CPU tensors."""

    def __init__(self, architecture, state_dict, dtype=torch.float32):
        self.arch = architecture
        self.dtype = dtype
        self.sd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in state_dict.items()}
        self.dim_y, self.dim_z = tuple(architecture["dim_y"]), tuple(architecture["dim_z"])
        self.min_z_var = architecture.get("min_z_var", 1e-7)

    def prior(self, y, aux):
        h = run_sequential(self.arch["prior_z_y"], self.sd, "prior_network", merge_aux_label(y, aux))
        return h[:, 0], h[:, 1]

    def sample_z(self, z_mu, z_log_var, eps):
        z = z_mu + eps * (torch.exp(z_log_var / 2) + self.min_z_var)
        return z.view(-1, *self.dim_z)

    def P(self, z, y, aux, taps=None):
        ym = merge_aux_label(y, aux)
        h_y = run_sequential(self.arch["p_y_in"], self.sd, "p_y_in", ym)
        h_z = run_sequential(self.arch["p_z_in"], self.sd, "p_z_in", z, taps)
        h = torch.cat([h_z, h_y], dim=1)
        h = run_sequential(self.arch["p_y_z_in"], self.sd, "p_y_z_in", h, taps)
        return run_sequential(self.arch["p_y_z_out"][0], self.sd, "p_mu_out", h, taps)

    def sample_P(self, y, aux, latent=None, eps=None, taps=None):
        """``latent`` given -> prior skipped (reference cvae.py:151-154, mode L);
        ``eps`` given -> prior run, ``z = mu + eps*(exp(lv/2)+min_z_var)`` (mode E);
        neither -> eps drawn with torch.randn like the reference."""
        with torch.no_grad():
            y = torch.as_tensor(y, dtype=self.dtype)
            aux = torch.as_tensor(aux, dtype=self.dtype)
            if latent is not None:
                z = torch.as_tensor(latent, dtype=self.dtype).view(-1, *self.dim_z)
            else:
                mu, lv = self.prior(y, aux)
                if taps is not None:
                    taps.append(("z_mu", mu)); taps.append(("z_log_var", lv))
                if eps is None:
                    eps = torch.randn(size=(1, *mu.size()))
                eps = torch.as_tensor(eps, dtype=self.dtype).view(1, *mu.size())
                z = self.sample_z(mu, lv, eps)
            if taps is not None:
                taps.append(("latent", z))
            return self.P(z, y, aux, taps)

    def Q(self, x, y, aux):
        """recognition network (reference cvae.py:68-80) -> (z_mu, z_log_var)"""
        h_x = run_sequential(self.arch["q_x_in"], self.sd, "q_x_in", x)
        h_y = run_sequential(self.arch["q_y_in"], self.sd, "q_y_in", merge_aux_label(y, aux))
        h = run_sequential(self.arch["q_x_y_out"], self.sd, "q_out", torch.cat([h_x, h_y], dim=1))
        return h[:, 0], h[:, 1]

    def forward(self, x, y, aux, eps):
        """evidence lower bound of a batch of transformed (x, y) pairs (reference cvae.py:122-147; fixed variance,
        L = 1, beta_KL = 1, ``torch.randn`` replaced by ``eps``) -> dict(ELBO, KL_term, log_likelihood, z_mu, z_log_var)"""
        import math
        with torch.no_grad():
            x = torch.as_tensor(x, dtype=self.dtype)
            y = torch.as_tensor(y, dtype=self.dtype)
            aux = torch.as_tensor(aux, dtype=self.dtype)
            z_mu, z_log_var = self.Q(x, y, aux)
            z = self.sample_z(z_mu, z_log_var, torch.as_tensor(eps, dtype=self.dtype).view(1, *z_mu.size()))
            M = x.size(0)
            pm, plv = self.prior(y, aux)
            pv = torch.exp(plv)
            KL = 0.5 / M * torch.sum((pm - z_mu) ** 2 / pv + torch.exp(z_log_var) / pv + plv - z_log_var - 1)
            x_mu = self.P(z, y, aux)
            ll = -0.5 * math.log(2 * math.pi) + 1 / M * (-0.5 * (x - x_mu) ** 2).sum(dim=[3, 2, 0])
            scaling = self.arch.get("likelihood_scaling", 1.0)
            return dict(ELBO=float(-KL + scaling * ll.sum()), KL_term=float(KL), log_likelihood=ll.numpy().astype(np.float64),
                        z_mu=z_mu.numpy(), z_log_var=z_log_var.numpy())

    def paint(self, tile, z, stats, latent=None, eps=None, transform=True, inverse_transform=True):
        """One tile, reference ``paint`` semantics -> float32 (H, W)."""
        y = forward_transform(tile, z, stats) if transform else tile
        y = y.reshape(1, *y.shape)
        if y.shape != (1, *self.dim_y):
            raise ValueError(f"Shape mismatch between input and model: {tile.shape} vs {self.dim_y}")
        pred = self.sample_P(y, np.float32(z), latent=latent, eps=eps).to(torch.float32).numpy()
        if inverse_transform:
            return inverse_transform_np(pred, z, stats)
        return pred

    def paint_batch(self, tiles, zs, stats, latents=None, eps=None):
        out = np.empty((len(tiles), *self.dim_y[1:]), np.float32)
        for i in range(len(tiles)):
            out[i] = self.paint(tiles[i], float(zs[i]), stats,
                                latent=None if latents is None else latents[i:i + 1],
                                eps=None if eps is None else eps[i:i + 1])
        return out


def inverse_transform_np(pred, z, stats):
    return np.asarray(inverse_transform(pred, z, stats), dtype=np.float32)


# --------------------------------------------------------------------------------------
# CGAN generator (restated from trained_models/README.md:116-128 + g_struc.pickle;
# parity unpinned -- no reference source or weights)
# --------------------------------------------------------------------------------------

class CGANOracle:
    def __init__(self, layers, state_dict, dtype=torch.float32):
        self.layers, self.dtype = layers, dtype
        self.sd = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in state_dict.items()}

    def generator(self, x):
        with torch.no_grad():
            return run_sequential(self.layers, self.sd, "generator", torch.as_tensor(x, dtype=self.dtype))

    def paint(self, tile, z, stats, transform=True, inverse_transform=True, z_shift=1.0):
        """``[x', z-1]`` -> generator -> tanh -> ``(exp((y+1)*4)-1)*sigma_p`` (App. C)."""
        y = forward_transform(tile, z, stats, k=4.0, shift=1.0) if transform else tile
        y = torch.as_tensor(y.reshape(1, *y.shape), dtype=self.dtype)
        x = merge_aux_label(y, torch.as_tensor(np.float32(z - z_shift), dtype=self.dtype))
        pred = self.generator(x).to(torch.float32).numpy()
        if inverse_transform:
            return np.asarray(globals()["inverse_transform"](pred, z, stats, k=4.0, shift=1.0), np.float32)
        return pred


# --------------------------------------------------------------------------------------
# pseudo power spectrum (restated; cosmotools.power_spectrum_tools.pseudo_Pofk is not
# vendored -- SURVEY.md section 8c(ii); the 1 % criterion is a ratio, so only consistency
# between the two sides matters)
# --------------------------------------------------------------------------------------

def pseudo_Pofk(A, B, L=100.0, n_k_bin=20, logspaced=True):
    n = A.shape[0]
    fa, fb = np.fft.rfft2(A), np.fft.rfft2(B)
    p = (fa * np.conj(fb)).real * (L / n ** 2) ** 2
    kx = np.fft.fftfreq(n, d=L / n) * 2 * np.pi
    ky = np.fft.rfftfreq(n, d=L / n) * 2 * np.pi
    k = np.sqrt(kx[:, None] ** 2 + ky[None, :] ** 2)
    k_min, k_max = 2 * np.pi / L, 2 * np.pi / L * n / 2
    edges = np.logspace(np.log10(k_min), np.log10(k_max), n_k_bin + 1) if logspaced \
        else np.linspace(k_min, k_max, n_k_bin + 1)
    w = np.ones_like(k)
    w[:, 1:-1] = 2.0                                  # rfft2 half-plane multiplicity
    idx = np.digitize(k.ravel(), edges) - 1
    ok = (idx >= 0) & (idx < n_k_bin)
    num = np.bincount(idx[ok], (p * w).ravel()[ok], n_k_bin)
    den = np.bincount(idx[ok], w.ravel()[ok], n_k_bin)
    kc = np.bincount(idx[ok], (k * w).ravel()[ok], n_k_bin)
    good = den > 0
    return kc[good] / den[good], num[good] / den[good]
