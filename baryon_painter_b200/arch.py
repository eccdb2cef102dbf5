"""Architecture dictionaries -> flat conv-layer plans with folded batch-norm.

The reference describes each sub-network as a list of string-tagged layers that
``build_sequential`` (reference ``baryon_painter/models/utils.py:114-157``) turns
into a ``torch.nn.Sequential``; the list index of every entry becomes part of the
``state_dict`` key.  This module walks the same lists without
building torch modules and produces, for every convolution, one
:class:`LayerSpec` that carries the fused epilogue the CUDA kernels implement:

    out = act( conv(x, W) * scale + shift  [+ skip] )

with ``scale = gamma / sqrt(running_var + eps)`` and
``shift = beta - running_mean * scale (+ bias * scale)`` (BatchNorm2d in eval mode),
``act`` one of none / relu / leaky-relu / prelu(1) / softplus(beta=1, thr=20) / tanh /
sigmoid, and ``skip`` the input of the enclosing ``ResidualBlock``
(reference ``models/utils.py:22-38``: ``act(block(x) + x)``).
"""

import collections
from dataclasses import dataclass, field

import numpy as np

ACT_IDS = {"none": 0, "relu": 1, "leaky": 2, "prelu": 3, "softplus": 4, "tanh": 5, "sigmoid": 6}
RES_NONE, RES_OPEN, RES_CLOSE = 0, 1, 2
# registration order of the sub-networks in CVAE.__init__ (reference cvae.py:24-50)
CVAE_STACKS = (("q_x_in", "q_x_in"), ("q_y_in", "q_y_in"), ("q_out", "q_x_y_out"),
               ("p_y_in", "p_y_in"), ("p_z_in", "p_z_in"), ("p_y_z_in", "p_y_z_in"),
               ("p_mu_out", "p_y_z_out:0"), ("p_var_out", "p_y_z_out:1"),
               ("prior_network", "prior_z_y"))


@dataclass
class LayerSpec:
    kind: str                 # "conv" | "convT"
    cin: int
    cout: int
    k: int
    stride: int
    pad: int
    out_pad: int = 0
    w_key: str = ""
    b_key: str = None         # conv bias key or None
    bn_prefix: str = None     # "<seq>.<idx>" of the BatchNorm2d or None
    bn_eps: float = 1e-5
    act: str = "none"
    act_param: float = 0.0    # leaky slope (prelu slope is read from act_key)
    act_key: str = None       # PReLU weight key
    res: int = RES_NONE

    def out_hw(self, h, w):
        if self.kind == "conv":
            f = lambda n: (n + 2 * self.pad - self.k) // self.stride + 1
        else:
            f = lambda n: (n - 1) * self.stride - 2 * self.pad + self.k + self.out_pad
        return f(h), f(w)

    def macs_per_out(self):
        """MACs per sample given input HxW are computed by the caller; this is Cin*Cout*k*k."""
        return self.cin * self.cout * self.k * self.k


@dataclass
class FoldedLayer:
    spec: LayerSpec
    weight: np.ndarray        # float32, PyTorch layout: conv (Cout,Cin,k,k); convT (Cin,Cout,k,k)
    scale: np.ndarray         # float32 (Cout,)
    shift: np.ndarray         # float32 (Cout,)
    act_param: float


def _square(v, what):
    if isinstance(v, (tuple, list)):
        if len(set(v)) != 1:
            raise NotImplementedError("non-square %s %r" % (what, v))
        return int(v[0])
    return int(v)


def flatten_stack(layers, prefix):
    """Architecture list -> [LayerSpec].  ``prefix`` is the Sequential's attribute name."""
    if layers is None:
        return []
    specs = []
    for idx, layer in enumerate(layers):
        if len(layer) == 2:
            name, config = layer
        elif len(layer) == 1:
            name, config = layer[0], None
        else:
            raise RuntimeError("Layer definition ill-formed: {}.".format(layer))
        name = name.lower()
        key = "%s.%d" % (prefix, idx)
        if name in ("conv", "transp conv"):
            extra = set(config) - {"in_channels", "out_channels", "kernel_size", "padding", "stride",
                                   "bias", "output_padding"}
            if extra:
                raise NotImplementedError("conv option(s) %s not supported" % sorted(extra))
            specs.append(LayerSpec(
                kind="conv" if name == "conv" else "convT",
                cin=int(config["in_channels"]), cout=int(config["out_channels"]),
                k=_square(config["kernel_size"], "kernel"), stride=_square(config.get("stride", 1), "stride"),
                pad=_square(config.get("padding", 0), "padding"),
                out_pad=_square(config.get("output_padding", 0), "output_padding"),
                w_key=key + ".weight", b_key=(key + ".bias") if config.get("bias", True) else None))
        elif name == "batchnorm":
            if not specs or specs[-1].bn_prefix is not None or specs[-1].act != "none":
                raise NotImplementedError("batchnorm must directly follow a convolution")
            if specs[-1].cout != int(config["num_features"]):
                raise ValueError("batchnorm width does not match the convolution before it")
            specs[-1].bn_prefix = key
            specs[-1].bn_eps = float(config.get("eps", 1e-5))
        elif name in ("relu", "leaky relu", "prelu", "tanh", "sigmoid", "softplus"):
            if not specs or specs[-1].act != "none":
                raise NotImplementedError("activation must follow a convolution (+batchnorm)")
            s = specs[-1]
            if name == "leaky relu":
                s.act, s.act_param = "leaky", float(config)
            elif name == "prelu":
                s.act, s.act_key = "prelu", key + ".weight"
            else:
                s.act = name
        elif name == "residual block":
            inner = flatten_stack(config[0], key + ".res_block")
            if not inner:
                raise NotImplementedError("empty residual block")
            if inner[-1].act != "none":
                raise NotImplementedError("activation before the residual add is not supported")
            for s in inner:
                if s.res != RES_NONE:
                    raise NotImplementedError("nested residual blocks")
            if inner[0].cin != inner[-1].cout:
                raise ValueError("residual block changes the channel count")
            a = config[1]
            aname = a[0].lower() if a[0] is not None else "none"
            if aname == "relu":
                inner[-1].act = "relu"
            elif aname == "leaky relu":
                inner[-1].act, inner[-1].act_param = "leaky", float(a[1])
            elif aname != "none":
                raise NotImplementedError("Layer {} not supported yet!".format(a[0]))
            if len(inner) == 1:
                raise NotImplementedError("single-convolution residual block")
            inner[0].res, inner[-1].res = RES_OPEN, RES_CLOSE
            specs.extend(inner)
        elif name in ("flatten", "unflatten"):
            continue                      # views only; shapes are checked by the painter
        else:
            raise NotImplementedError("Layer {} not supported yet!".format(name))
    return specs


def _stack_layers(architecture, arch_key):
    if ":" in arch_key:
        k, i = arch_key.split(":")
        seq = architecture.get(k)
        return seq[int(i)] if seq is not None and len(seq) > int(i) else None
    return architecture.get(arch_key)


def cvae_stacks(architecture):
    """-> OrderedDict attr_name -> [LayerSpec] for every sub-network present."""
    if architecture["type"] != "Type-1":
        raise NotImplementedError("Architecture {} not supported yet!".format(architecture["type"]))
    out = collections.OrderedDict()
    for attr, arch_key in CVAE_STACKS:
        layers = _stack_layers(architecture, arch_key)
        if layers is not None:
            out[attr] = flatten_stack(layers, attr)
    return out


def state_dict_schema(stacks):
    """OrderedDict key -> (shape, dtype str), in ``model.state_dict()`` order (App. D)."""
    schema = collections.OrderedDict()
    for specs in stacks.values():
        for s in specs:
            shape = (s.cout, s.cin, s.k, s.k) if s.kind == "conv" else (s.cin, s.cout, s.k, s.k)
            schema[s.w_key] = (shape, "float32")
            if s.b_key:
                schema[s.b_key] = ((s.cout,), "float32")
            if s.bn_prefix:
                for p in ("weight", "bias", "running_mean", "running_var"):
                    schema["%s.%s" % (s.bn_prefix, p)] = ((s.cout,), "float32")
                schema[s.bn_prefix + ".num_batches_tracked"] = ((), "int64")
            if s.act_key:
                schema[s.act_key] = ((1,), "float32")
    return schema


def _np(t):
    return t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)


def check_state_dict(schema, state_dict):
    """Strict key/shape check, the error ``load_state_dict`` raises (reference painter.py:432)."""
    missing = [k for k in schema if k not in state_dict]
    unexpected = [k for k in state_dict if k not in schema]
    bad = [k for k in schema if k in state_dict and tuple(_np(state_dict[k]).shape) != tuple(schema[k][0])]
    if missing or unexpected or bad:
        msg = "Error(s) in loading state_dict for CVAE:"
        if missing:
            msg += "\n\tMissing key(s) in state_dict: %s." % ", ".join('"%s"' % k for k in missing)
        if unexpected:
            msg += "\n\tUnexpected key(s) in state_dict: %s." % ", ".join('"%s"' % k for k in unexpected)
        for k in bad:
            msg += "\n\tsize mismatch for %s: %s vs %s." % (k, tuple(_np(state_dict[k]).shape), schema[k][0])
        raise RuntimeError(msg)


def fold_stack(specs, state_dict):
    """Fold BatchNorm (eval) and bias into per-channel scale/shift; read PReLU slopes."""
    folded = []
    for s in specs:
        w = np.ascontiguousarray(_np(state_dict[s.w_key]), dtype=np.float32)
        scale = np.ones(s.cout, np.float64)
        shift = np.zeros(s.cout, np.float64)
        if s.b_key:
            shift = shift + _np(state_dict[s.b_key]).astype(np.float64)
        if s.bn_prefix:
            g = _np(state_dict[s.bn_prefix + ".weight"]).astype(np.float64)
            b = _np(state_dict[s.bn_prefix + ".bias"]).astype(np.float64)
            m = _np(state_dict[s.bn_prefix + ".running_mean"]).astype(np.float64)
            v = _np(state_dict[s.bn_prefix + ".running_var"]).astype(np.float64)
            inv = g / np.sqrt(v + s.bn_eps)
            shift = (shift - m) * inv + b
            scale = scale * inv
        act_param = s.act_param
        if s.act_key:
            a = _np(state_dict[s.act_key]).reshape(-1)
            if a.size != 1:
                raise NotImplementedError("per-channel PReLU")
            act_param = float(a[0])
        folded.append(FoldedLayer(s, w, scale.astype(np.float32), shift.astype(np.float32), act_param))
    return folded


def stack_flops(specs, h, w):
    """2*MACs per sample of a stack applied to an (h, w) input (SURVEY.md App. A counting:
    conv Cin*Cout*k^2 per output pixel; convT Cin*Cout*k^2 per *input* pixel)."""
    total = 0
    for s in specs:
        oh, ow = s.out_hw(h, w)
        px = oh * ow if s.kind == "conv" else h * w
        total += 2 * px * s.macs_per_out()
        h, w = oh, ow
    return total, (h, w)


# ---------------------------------------------------------------------------------------
# The fiducial architectures
# ---------------------------------------------------------------------------------------

def _block(cin, cout, type="conv", scale=1, kernel=3, bias=False, batchnorm=True, activation="relu"):
    """One conv + optional BN + activation group (reference models/utils.py:40-77)."""
    if scale == 1:
        kps = {"kernel_size": kernel, "padding": (kernel - 1) // 2, "stride": 1}
    else:
        kps = {"kernel_size": 2 * scale, "padding": scale // 2, "stride": scale}
    out = [(type, {"in_channels": cin, "out_channels": cout, **kps, "bias": bias})]
    if batchnorm:
        out.append(("batchnorm", {"num_features": cout}))
    tag = {"relu": ("ReLU",), "prelu": ("prelu",), "softplus": ("softplus",), None: None}[activation]
    if tag:
        out.append(tag)
    return out


def _chain(cin, channels, scales, type="conv"):
    out = []
    for c, s in zip(channels, scales):
        out += _block(cin, c, type=type, scale=s)
        cin = c
    return out


def _res(c):
    inner = _block(c, c, kernel=3) + _block(c, c, kernel=3, activation=None)
    return ("residual block", (inner, ("ReLU",)))


def fiducial_cvae_architecture(tile_size=512):
    """The fiducial CVAE dict (``scripts/CVAE_single_scale.py:98-138`` with the variance head
    dropped, as shipped in ``trained_models/CVAE/fiducial/architecture.txt``).  ``tile_size``
    other than 512 gives the same network on smaller tiles (latent = tile/32), used by tests."""
    if tile_size % 32:
        raise ValueError("tile_size must be a multiple of 32")
    dim_z = (1, tile_size // 32, tile_size // 32)
    down = lambda cin: _chain(cin, [8, 16, 32], [2, 4, 4])
    return {
        "type": "Type-1", "dim_x": (1, tile_size, tile_size), "dim_y": (1, tile_size, tile_size),
        "dim_z": dim_z, "n_x_features": 1, "aux_label": True,
        "prior_z_y": down(2) + _block(32, 2, kernel=5) + [("unflatten", (2, *dim_z))],
        "q_x_in": down(1), "q_y_in": down(2),
        "q_x_y_out": _block(64, 2, kernel=5) + [("unflatten", (2, *dim_z))],
        "p_y_in": None,
        "p_z_in": _chain(1, [1, 1, 1], [2, 4, 4], type="transp conv"),
        "p_y_z_in": _block(3, 16, kernel=5) + _chain(16, [32, 64, 128], [2, 2, 2])
                    + [_res(128) for _ in range(4)]
                    + _chain(128, [64, 32, 16], [2, 2, 2], type="transp conv"),
        "p_y_z_out": (_block(16, 8, kernel=7, batchnorm=False, activation="prelu")
                      + _block(8, 1, kernel=5, batchnorm=False, activation="prelu")
                      + _block(1, 1, kernel=3, batchnorm=False, activation="softplus"),),
        "min_x_var": 1e-07, "min_z_var": 1e-07, "L": 1,
    }


def fiducial_cgan_architecture(n_res_blocks=9):
    """Generator of the fiducial CGAN as a layer list in the same tagged format
    (restated from ``trained_models/README.md:116-128`` and ``g_struc.pickle``,
    SURVEY.md App. C; the PainterGAN source is not part of the reference tree)."""
    def cbl(type, cin, cout, k, s, p, bias, op=0, bn=True, act=("Leaky ReLU", 0.2)):
        cfg = {"in_channels": cin, "out_channels": cout, "kernel_size": k, "padding": p,
               "stride": s, "bias": bias}
        if op:
            cfg["output_padding"] = op
        out = [(type, cfg)]
        if bn:
            out.append(("batchnorm", {"num_features": cout}))
        if act:
            out.append(act)
        return out

    def res(c):
        inner = (cbl("conv", c, c, 3, 1, 1, False) + cbl("conv", c, c, 3, 1, 1, False, act=None))
        return ("residual block", (inner, ("Leaky ReLU", 0.2)))

    return (cbl("conv", 2, 32, 9, 1, 4, False) + cbl("conv", 32, 64, 3, 2, 1, True)
            + cbl("conv", 64, 128, 3, 2, 1, True) + [res(128) for _ in range(n_res_blocks)]
            + cbl("transp conv", 128, 64, 3, 2, 1, True, op=1)
            + cbl("transp conv", 64, 32, 3, 2, 1, True, op=1)
            + cbl("conv", 32, 1, 9, 1, 4, True, bn=False, act=("tanh",)))
