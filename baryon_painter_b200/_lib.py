"""ctypes binding of ``libbaryon_painter_b200.so`` (C ABI in ``include/baryon_painter_b200.h``).

The library is built in-tree by ``baryon_painter_b200.build`` (``nvcc`` for sm_100a).  There is
no fallback of any kind: if the shared object is missing or no sm_100 GPU is visible every entry
point raises.
"""

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libbaryon_painter_b200.so")

BP_OK, BP_E_INVALID, BP_E_UNSUPPORTED, BP_E_CUDA, BP_E_NO_DEVICE, BP_E_NOMEM = 0, -1, -2, -3, -4, -5
BP_CONV, BP_CONVT = 0, 1
BP_PREC_F32, BP_PREC_BF16, BP_PREC_F16, BP_PREC_F32_FFMA = 0, 1, 2, 3
BP_LATENT_GIVEN, BP_LATENT_EPS, BP_LATENT_SEED = 0, 1, 2
BP_FLAG_TRANSFORM, BP_FLAG_INVERSE = 1, 2
PRECISIONS = {"fp32": BP_PREC_F32, "f32": BP_PREC_F32, "float32": BP_PREC_F32, "fp32-ffma": BP_PREC_F32_FFMA,
              "bf16": BP_PREC_BF16, "bfloat16": BP_PREC_BF16,
              "fp16": BP_PREC_F16, "f16": BP_PREC_F16, "float16": BP_PREC_F16}

c_float_p = ctypes.POINTER(ctypes.c_float)

# every symbol declared in include/baryon_painter_b200.h (checked by tests/test_cabi.py)
EXPORTS = ("bp_device_count", "bp_cvae_create", "bp_cgan_create", "bp_net_destroy", "bp_cvae_paint",
           "bp_cvae_paint_host", "bp_cvae_read_prior", "bp_cgan_paint", "bp_cgan_paint_host",
           "bp_cvae_paint_variance_host", "bp_stitch_accumulate", "bp_stitch_finalize", "bp_zoom_tiles", "bp_zoom_accumulate", "bp_plane_prepare",
           "bp_net_set_debug", "bp_net_read_activation", "bp_net_set_profile", "bp_net_read_profile",
           "bp_net_layer_info", "bp_launch_count", "bp_net_flops_per_tile", "bp_net_chunk",
           "bp_tuning_set", "bp_tuning_get", "bp_tuning_mode", "bp_rng_normal_host", "bp_cvae_elbo_host", "bp_cvae_paint_host_async", "bp_net_wait",
           "bp_last_error", "bp_version")

# which tensor-core formulation / tiling every layer runs with (see include/baryon_painter_b200.h, "formulation table")
TUNING_TABLE = os.path.join(_HERE, "tuning_table.txt")


class LayerDesc(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("cin", ctypes.c_int32), ("cout", ctypes.c_int32),
                ("kernel", ctypes.c_int32), ("stride", ctypes.c_int32), ("pad", ctypes.c_int32),
                ("out_pad", ctypes.c_int32), ("act", ctypes.c_int32), ("act_param", ctypes.c_float),
                ("res", ctypes.c_int32), ("weight", c_float_p), ("scale", c_float_p), ("shift", c_float_p)]


class CvaeDesc(ctypes.Structure):
    _fields_ = [("tile_h", ctypes.c_int32), ("tile_w", ctypes.c_int32), ("latent_h", ctypes.c_int32),
                ("latent_w", ctypes.c_int32), ("min_z_var", ctypes.c_float),
                ("n_prior", ctypes.c_int32), ("n_p_z_in", ctypes.c_int32), ("n_p_y_z_in", ctypes.c_int32),
                ("n_p_mu_out", ctypes.c_int32),
                ("prior", ctypes.POINTER(LayerDesc)), ("p_z_in", ctypes.POINTER(LayerDesc)),
                ("p_y_z_in", ctypes.POINTER(LayerDesc)), ("p_mu_out", ctypes.POINTER(LayerDesc)),
                ("n_q_x_in", ctypes.c_int32), ("n_q_y_in", ctypes.c_int32), ("n_q_out", ctypes.c_int32),
                ("q_x_in", ctypes.POINTER(LayerDesc)), ("q_y_in", ctypes.POINTER(LayerDesc)),
                ("q_out", ctypes.POINTER(LayerDesc)), ("likelihood_scaling", ctypes.c_float)]


class TransformParams(ctypes.Structure):
    _fields_ = [("sigma_in", c_float_p), ("sigma_out", c_float_p), ("aux", c_float_p),
                ("k_in", ctypes.c_float), ("shift_in", ctypes.c_float),
                ("k_out", ctypes.c_float), ("shift_out", ctypes.c_float)]


_lib = None


def load():
    """Load the shared library (once).  Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "baryon_painter_b200: %s is missing -- build it with `python -m baryon_painter_b200.build` "
            "(nvcc, sm_100a).  There is no CPU or PyTorch fallback for the paint path." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, u64, f32 = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64, ctypes.c_float
    lib.bp_device_count.restype = i32
    lib.bp_version.restype = i32
    lib.bp_last_error.restype = ctypes.c_char_p
    lib.bp_launch_count.restype = ctypes.c_int64
    lib.bp_launch_count.argtypes = [i32]
    lib.bp_net_flops_per_tile.restype = ctypes.c_double
    lib.bp_net_flops_per_tile.argtypes = [vp]
    lib.bp_net_chunk.restype = ctypes.c_int
    lib.bp_net_chunk.argtypes = [vp]
    lib.bp_cvae_create.argtypes = [ctypes.POINTER(CvaeDesc), i32, i32, i32, ctypes.POINTER(vp)]
    lib.bp_cgan_create.argtypes = [ctypes.POINTER(LayerDesc), i32, i32, i32, i32, i32, i32, ctypes.POINTER(vp)]
    lib.bp_net_destroy.argtypes = [vp]
    lib.bp_net_destroy.restype = None
    tpp = ctypes.POINTER(TransformParams)
    lib.bp_cvae_paint.argtypes = [vp, vp, vp, i32, u64, tpp, i32, vp, i32, vp]
    lib.bp_cvae_paint_host.argtypes = [vp, vp, vp, i32, u64, tpp, i32, vp, i32]
    lib.bp_cvae_read_prior.argtypes = [vp, vp, vp, i32]
    lib.bp_cgan_paint.argtypes = [vp, vp, tpp, i32, vp, i32, vp]
    lib.bp_cgan_paint_host.argtypes = [vp, vp, tpp, i32, vp, i32]
    lib.bp_cvae_paint_variance_host.argtypes = [vp, vp, tpp, i32, u64, vp, vp, i32]
    lib.bp_stitch_accumulate.argtypes = [vp, vp, i32, vp, vp, i32, i32, f32, f32, vp]
    lib.bp_stitch_finalize.argtypes = [vp, vp, vp, ctypes.c_size_t, vp]
    lib.bp_zoom_tiles.argtypes = [i32, vp, i32, i32, vp, i32, i32, i32, i32, vp, vp]
    lib.bp_zoom_accumulate.argtypes = [i32, vp, i32, i32, i32, i32, ctypes.c_double, vp, vp]
    lib.bp_plane_prepare.argtypes = [i32, vp, i32, i32, f32, f32, vp, vp]
    lib.bp_net_set_debug.argtypes = [vp, i32]
    lib.bp_net_read_activation.argtypes = [vp, i32, i32, vp, ctypes.c_size_t]
    lib.bp_net_set_profile.argtypes = [vp, i32]
    lib.bp_net_read_profile.argtypes = [vp, i32, i32, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(i32)]
    lib.bp_net_layer_info.argtypes = [vp, i32, i32, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(i32)]
    lib.bp_tuning_set.argtypes = [ctypes.c_char_p]
    lib.bp_tuning_get.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
    lib.bp_tuning_mode.argtypes = [i32, i32]
    lib.bp_rng_normal_host.argtypes = [i32, u64, u64, vp, ctypes.c_size_t]
    lib.bp_cvae_paint_host_async.argtypes = [vp, vp, vp, i32, u64, tpp, i32, vp, i32, i32]
    lib.bp_net_wait.argtypes = [vp, i32]
    lib.bp_cvae_elbo_host.argtypes = [vp, vp, vp, vp, i32, u64, tpp, i32, i32, vp, vp, vp]
    _lib = lib
    # the shipped formulation table: every process builds the same kernels for the same layer, so painted tiles
    # are bit-identical across processes (BARYON_PAINTER_TUNING_TABLE points at another table; "" = none)
    path = os.environ.get("BARYON_PAINTER_TUNING_TABLE", TUNING_TABLE)
    if path and os.path.exists(path):
        with open(path, "rb") as f:
            check(lib.bp_tuning_set(f.read()))
    return lib


def check(rc):
    """Map a BP_E_* status to the exception the reference raises in the same situation."""
    if rc == BP_OK:
        return
    msg = load().bp_last_error().decode("utf-8", "replace")
    if rc == BP_E_INVALID:
        raise ValueError(msg)
    if rc == BP_E_UNSUPPORTED:
        raise NotImplementedError(msg)
    if rc == BP_E_NOMEM:
        raise MemoryError(msg)
    raise RuntimeError(msg)


def _fptr(a):
    return a.ctypes.data_as(c_float_p) if a is not None else c_float_p()


def make_layer_descs(folded):
    """[arch.FoldedLayer] -> (ctypes array of LayerDesc, keep-alive list of numpy arrays)."""
    from . import arch
    arr = (LayerDesc * max(len(folded), 1))()
    keep = []
    for i, f in enumerate(folded):
        s = f.spec
        w = np.ascontiguousarray(f.weight, np.float32)
        sc = np.ascontiguousarray(f.scale, np.float32)
        sh = np.ascontiguousarray(f.shift, np.float32)
        keep += [w, sc, sh]
        arr[i] = LayerDesc(BP_CONV if s.kind == "conv" else BP_CONVT, s.cin, s.cout, s.k, s.stride, s.pad,
                           s.out_pad, arch.ACT_IDS[s.act], float(f.act_param), int(s.res),
                           _fptr(w), _fptr(sc), _fptr(sh))
    return arr, keep


class Net:
    """Owner of one ``bp_net*``."""

    def __init__(self, handle, kind, tile_hw, latent_hw, precision, max_batch, device):
        self.handle, self.kind = handle, kind
        self.tile_hw, self.latent_hw = tuple(tile_hw), tuple(latent_hw) if latent_hw else None
        self.precision, self.max_batch, self.device = precision, max_batch, device

    @classmethod
    def create_cvae(cls, stacks, tile_hw, latent_hw, min_z_var, precision, max_batch, device, likelihood_scaling=1.0):
        lib = load()
        keep, arrs = [], {}
        for name in ("prior_network", "p_z_in", "p_y_z_in", "p_mu_out", "q_x_in", "q_y_in", "q_out"):
            arrs[name], k = make_layer_descs(stacks.get(name, []))
            keep += k
        have_q = all(stacks.get(k) for k in ("q_x_in", "q_y_in", "q_out"))
        d = CvaeDesc(tile_hw[0], tile_hw[1], latent_hw[0], latent_hw[1], float(min_z_var),
                     len(stacks.get("prior_network", [])), len(stacks["p_z_in"]), len(stacks["p_y_z_in"]),
                     len(stacks["p_mu_out"]), arrs["prior_network"], arrs["p_z_in"], arrs["p_y_z_in"],
                     arrs["p_mu_out"],
                     len(stacks["q_x_in"]) if have_q else 0, len(stacks["q_y_in"]) if have_q else 0,
                     len(stacks["q_out"]) if have_q else 0, arrs["q_x_in"], arrs["q_y_in"], arrs["q_out"],
                     float(likelihood_scaling))
        h = ctypes.c_void_p()
        check(lib.bp_cvae_create(ctypes.byref(d), PRECISIONS[precision], int(max_batch), int(device),
                                 ctypes.byref(h)))
        return cls(h, "cvae", tile_hw, latent_hw, precision, max_batch, device)

    @classmethod
    def create_cgan(cls, layers, tile_hw, precision, max_batch, device):
        lib = load()
        arr, keep = make_layer_descs(layers)
        h = ctypes.c_void_p()
        check(lib.bp_cgan_create(arr, len(layers), tile_hw[0], tile_hw[1], PRECISIONS[precision], int(max_batch),
                                 int(device), ctypes.byref(h)))
        return cls(h, "cgan", tile_hw, None, precision, max_batch, device)

    def close(self):
        if self.handle:
            load().bp_net_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def flops_per_tile(self):
        return load().bp_net_flops_per_tile(self.handle)

    @property
    def chunk(self):
        return int(load().bp_net_chunk(self.handle))

    # -- helpers -------------------------------------------------------------------------
    @staticmethod
    def _tp(sigma_in, sigma_out, aux, k_in, shift_in, k_out, shift_out):
        f = lambda a: None if a is None else np.ascontiguousarray(a, np.float32)
        sigma_in, sigma_out, aux = f(sigma_in), f(sigma_out), f(aux)
        tp = TransformParams(_fptr(sigma_in), _fptr(sigma_out), _fptr(aux), k_in, shift_in, k_out, shift_out)
        return tp, (sigma_in, sigma_out, aux)

    def cvae_paint_host(self, tiles, latent, mode, seed, tparams, flags, out=None):
        """tiles float32 [n,H,W] (host) -> float32 [n,H,W].  ``out``: optional C-contiguous float32 result
        buffer (e.g. page-locked, see ``pinned_empty``), else a fresh array."""
        lib = load()
        n = tiles.shape[0]
        tiles = np.ascontiguousarray(tiles, np.float32)
        if out is None:
            out = np.empty((n, *self.tile_hw), np.float32)
        elif out.dtype != np.float32 or not out.flags.c_contiguous or out.shape != (n, *self.tile_hw):
            raise ValueError("out must be a C-contiguous float32 array of shape %r" % ((n, *self.tile_hw),))
        lat = None if latent is None else np.ascontiguousarray(latent, np.float32)
        tp, keep = self._tp(*tparams)
        check(lib.bp_cvae_paint_host(self.handle, tiles.ctypes.data, lat.ctypes.data if lat is not None else None,
                                     mode, ctypes.c_uint64(seed & (2 ** 64 - 1)), ctypes.byref(tp), flags,
                                     out.ctypes.data, n))
        return out

    def cvae_paint_host_async(self, tiles, latent, mode, seed, tparams, flags, out, slot):
        """Enqueue one batch on I/O slot ``slot`` (see bp_cvae_paint_host_async); returns the objects that must stay
        alive until ``wait(slot)``."""
        n = tiles.shape[0]
        tp, keep = self._tp(*tparams)
        check(load().bp_cvae_paint_host_async(self.handle, tiles.ctypes.data, None if latent is None else latent.ctypes.data,
                                              mode, ctypes.c_uint64(seed & (2 ** 64 - 1)), ctypes.byref(tp), flags,
                                              out.ctypes.data, n, slot))
        return (tiles, latent, out, keep)

    def wait(self, slot):
        check(load().bp_net_wait(self.handle, slot))

    def cvae_paint_device(self, tiles_ptr, latent_ptr, mode, seed, tparams, flags, out_ptr, n, stream=0):
        tp, keep = self._tp(*tparams)
        check(load().bp_cvae_paint(self.handle, tiles_ptr, latent_ptr, mode, ctypes.c_uint64(seed & (2 ** 64 - 1)),
                                   ctypes.byref(tp), flags, out_ptr, n, stream))

    def cvae_read_prior(self, n):
        mu = np.empty((n, *self.latent_hw), np.float32)
        lv = np.empty((n, *self.latent_hw), np.float32)
        check(load().bp_cvae_read_prior(self.handle, mu.ctypes.data, lv.ctypes.data, n))
        return mu, lv

    def cvae_paint_variance_host(self, tiles, tparams, n_draws, seed):
        n = tiles.shape[0]
        tiles = np.ascontiguousarray(tiles, np.float32)
        mean = np.empty((n, *self.tile_hw), np.float32)
        var = np.empty((n, *self.tile_hw), np.float32)
        tp, keep = self._tp(*tparams)
        check(load().bp_cvae_paint_variance_host(self.handle, tiles.ctypes.data, ctypes.byref(tp), int(n_draws),
                                                 ctypes.c_uint64(seed & (2 ** 64 - 1)), mean.ctypes.data,
                                                 var.ctypes.data, n))
        return mean, var

    def cvae_elbo_host(self, x_tiles, y_tiles, eps, mode, seed, tparams, flags):
        """(ELBO, KL_term, log_likelihood), z_mu, z_log_var of a batch (reference CVAE.forward, cvae.py:122-147)."""
        n = x_tiles.shape[0]
        x_tiles = np.ascontiguousarray(x_tiles, np.float32)
        y_tiles = np.ascontiguousarray(y_tiles, np.float32)
        eps = None if eps is None else np.ascontiguousarray(eps, np.float32)
        stats = np.zeros(3, np.float64)
        mu = np.empty((n, *self.latent_hw), np.float32)
        lv = np.empty((n, *self.latent_hw), np.float32)
        tp, keep = self._tp(*tparams)
        check(load().bp_cvae_elbo_host(self.handle, x_tiles.ctypes.data, y_tiles.ctypes.data,
                                       None if eps is None else eps.ctypes.data, mode, ctypes.c_uint64(seed & (2 ** 64 - 1)),
                                       ctypes.byref(tp), flags, n, stats.ctypes.data, mu.ctypes.data, lv.ctypes.data))
        return stats, mu, lv

    def cgan_paint_host(self, tiles, tparams, flags, out=None):
        n = tiles.shape[0]
        tiles = np.ascontiguousarray(tiles, np.float32)
        if out is None:
            out = np.empty((n, *self.tile_hw), np.float32)
        elif out.dtype != np.float32 or not out.flags.c_contiguous or out.shape != (n, *self.tile_hw):
            raise ValueError("out must be a C-contiguous float32 array of shape %r" % ((n, *self.tile_hw),))
        tp, keep = self._tp(*tparams)
        check(load().bp_cgan_paint_host(self.handle, tiles.ctypes.data, ctypes.byref(tp), flags, out.ctypes.data, n))
        return out

    def cgan_paint_device(self, tiles_ptr, tparams, flags, out_ptr, n, stream=0):
        tp, keep = self._tp(*tparams)
        check(load().bp_cgan_paint(self.handle, tiles_ptr, ctypes.byref(tp), flags, out_ptr, n, stream))

    def set_debug(self, on):
        check(load().bp_net_set_debug(self.handle, int(bool(on))))

    def set_profile(self, on):
        check(load().bp_net_set_profile(self.handle, int(bool(on))))

    def layer_info(self, stack, layer):
        fl = ctypes.c_double()
        geom = (ctypes.c_int * 10)()
        check(load().bp_net_layer_info(self.handle, stack, layer, ctypes.byref(fl), geom))
        keys = ("kind", "cin", "cout", "kernel", "stride", "H", "W", "OH", "OW", "tensor")
        d = dict(zip(keys, list(geom)))
        d["flops"] = fl.value
        return d

    def read_profile(self, stack, layer):
        ms, cnt = ctypes.c_double(), ctypes.c_int()
        check(load().bp_net_read_profile(self.handle, stack, layer, ctypes.byref(ms), ctypes.byref(cnt)))
        return ms.value, cnt.value

    def read_activation(self, stack, layer, shape):
        out = np.empty(shape, np.float32)
        check(load().bp_net_read_activation(self.handle, stack, layer, out.ctypes.data, out.size))
        return out


def tuning_mode(on, log=False):
    check(load().bp_tuning_mode(int(bool(on)), int(bool(log))))


def tuning_get():
    lib = load()
    n = lib.bp_tuning_get(None, 0)
    buf = ctypes.create_string_buffer(n + 1)
    lib.bp_tuning_get(buf, n + 1)
    return buf.value.decode()


def tuning_set(text):
    check(load().bp_tuning_set(text.encode() if isinstance(text, str) else text))


def rng_normal(seed, offset, n, device=0):
    """The first ``n`` draws, from counter ``offset``, of the device's BP_LATENT_SEED standard-normal generator."""
    out = np.empty(int(n), np.float32)
    check(load().bp_rng_normal_host(int(device), int(seed) & (2 ** 64 - 1), int(offset), out.ctypes.data, out.size))
    return out


def launch_count(reset=False):
    return int(load().bp_launch_count(int(bool(reset))))


def stitch_accumulate(num_ptr, den_ptr, n_pixel_plane, tiles_ptr, origins_ptr, n, tile_size, falloff, sigma, stream=0):
    """Device pointers: float64 numerator / denominator planes, float32 tiles [n,T,T], int32 origins [n,2]."""
    check(load().bp_stitch_accumulate(num_ptr, den_ptr, int(n_pixel_plane), tiles_ptr, origins_ptr, int(n),
                                      int(tile_size), float(falloff), float(sigma), stream))


def stitch_finalize(num_ptr, den_ptr, plane_ptr, n_pixels, stream=0):
    check(load().bp_stitch_finalize(num_ptr, den_ptr, plane_ptr, int(n_pixels), stream))


ZOOM_MODES = {"reflect": 0, "mirror": 1}       # BP_ZOOM_REFLECT / BP_ZOOM_MIRROR


def zoom_tiles(device, plane_ptr, plane_h, plane_w, origins_ptr, side, n, out_side, mode, out_ptr, stream=0):
    """Device pointers: float32 plane [plane_h, plane_w], int32 origins [n,2] (row0, col0 of each periodic
    side x side crop), float32 out [n,out_side,out_side] = scipy.ndimage.zoom(crop, out_side/side, mode=mode)."""
    check(load().bp_zoom_tiles(int(device), plane_ptr, int(plane_h), int(plane_w), origins_ptr, int(side), int(n),
                               int(out_side), ZOOM_MODES[mode], out_ptr, stream))


def zoom_accumulate(device, plane_ptr, side, out_side, order, mode, scale, map_ptr, stream=0):
    """Device pointers: float64 plane [side, side], float64 map [out_side, out_side];
    map += scale * scipy.ndimage.zoom(nan_to_zero(plane), out_side / side, order=order, mode=mode)."""
    check(load().bp_zoom_accumulate(int(device), plane_ptr, int(side), int(out_side), int(order), ZOOM_MODES[mode],
                                    float(scale), map_ptr, stream))


def plane_prepare(device, raw_ptr, rows, cols, add, mul, out_ptr, stream=0):
    """Device pointers: raw float32 [rows, cols] -> out float32 [cols, rows] = (raw.T + add) * mul (float32 roundings
    of numpy's in-place ``+=`` / ``*=``)."""
    check(load().bp_plane_prepare(int(device), raw_ptr, int(rows), int(cols), float(add), float(mul), out_ptr, stream))


def pinned_empty(shape, dtype=np.float32):
    """Page-locked numpy array (torch owns the allocation): the host entry points DMA straight from / into
    such buffers instead of staging through the library's own pinned memory."""
    import torch
    t = torch.empty(tuple(int(v) for v in np.atleast_1d(shape)), dtype=getattr(torch, np.dtype(dtype).name),
                    pin_memory=True)
    a = t.numpy()
    _PINNED_KEEPALIVE[a.ctypes.data] = t
    return a


_PINNED_KEEPALIVE = {}
