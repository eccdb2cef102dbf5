"""Lightcone tiling and stitching -- drop-in for reference ``baryon_painter/process_SLICS.py``.

Same public names and call contract as the reference (``process_SLICS`` :128-226, ``get_tile`` :68-83,
``make_weight_map`` :85-99, ``generate_tiling`` :102-126, ``create_y_map`` :12-66); reference quirks are kept
(Q1 the ``min_tiling_overlap`` argument is ignored and 0.5 is used, Q7 ``mirror`` zoom for
the mass-plane branch and ``reflect`` for the delta branch, Q9 truncating pixel arithmetic).

What changes is where the work runs:
  * all tiles of a plane are painted in one batched call (``painter.paint_batch`` when the painter has it,
    else the reference's one-tile ``paint`` loop), tiles staying on the device;
  * the Gaussian-edge weighted accumulation ``sum w*p / sum w`` runs on the device in float64 through the C ABI
    (``bp_stitch_accumulate`` / ``bp_stitch_finalize``);
  * with ``world_size > 1`` the (plane, tile) work items are dealt round-robin to the ranks, every rank keeps
    partial (numerator, denominator) planes, and ONE ``torch.distributed.reduce`` per run (NCCL over
    NVLink on a GPU box) assembles them on rank 0 -- no collective while painting.

Additive keyword arguments (all optional): ``plane_source`` (callable returning in-memory planes instead of
the SLICS files), ``rank`` / ``world_size`` / ``group`` (sharding), ``batch`` (tiles per paint call),
``backend`` (test seam; the default device backend has no CPU path).
"""

import os

import numpy as np

pi = np.pi

N_PIXEL_DELTA = 7745                 # reference process_SLICS.py:142-144
N_PIXEL_MASSPLANE = 4096 * 3
MASSPLANE_SIZE = 505                 # Mpc/h
MASS_NORM = 1 / (3072 ** 3 / 2 / 12288 ** 2)


# ---------------------------------------------------------------------------------------------------
# tiling geometry (host; pure index arithmetic)
# ---------------------------------------------------------------------------------------------------
def get_tile(m, shift, tile_relative_size, expansion_factor=1):
    """Periodic crop of ``m`` (reference :68-83): origin ``int(n*shift)``, side
    ``int(n*tile_relative_size*expansion_factor)``, centred expansion, wrap-around indexing."""
    if expansion_factor < 1:
        raise ValueError("Expension factors < 1 not supported.")
    n = m.shape[0]
    side = int(n * tile_relative_size * expansion_factor)
    pad = int(n * tile_relative_size * (expansion_factor - 1) / 2)
    rows = (int(n * shift[0]) - pad + np.arange(side)) % m.shape[0]
    cols = (int(n * shift[1]) - pad + np.arange(side)) % m.shape[1]
    return m[np.ix_(rows, cols)]


def _edge_profile(n, fp, sigma):
    """1-D factor of the blend weights: Gaussian ramp ``exp(-d^2 / 2 (fp sigma)^2)`` over the ``fp`` pixels at
    both ends (d = distance to the first interior pixel); overlapping ramps multiply."""
    g = np.ones(n)
    if fp > 0:
        d = fp - np.arange(fp)
        ramp = np.exp(-0.5 * d ** 2 / (fp * sigma) ** 2)
        g[:fp] *= ramp
        g[::-1][:fp] *= ramp
    return g


def make_weight_map(tile_shape, falloff=0.05, sigma=1):
    """Blend weights (reference :85-99): outer product of the row and column edge profiles; the ramp length
    ``int(tile_shape[0]*falloff)`` is taken from the first axis for both, as in the reference."""
    fp = int(tile_shape[0] * falloff)
    return _edge_profile(tile_shape[0], fp, sigma)[:, None] * _edge_profile(tile_shape[1], fp, sigma)[None, :]


def generate_tiling(n_pixel_plane, n_pixel_tile, min_tile_overlap=0.5):
    """Tile origins (fractions of the plane) and the destination slices (reference :102-126)."""
    t = n_pixel_tile / n_pixel_plane
    n_inner = 0
    if t < 1 - t + t * min_tile_overlap:
        step = t * (1 - min_tile_overlap)
        gap = 1 - 2 * t + t * min_tile_overlap
        n_inner = 1 if gap <= step else int(np.ceil((gap - step) / step)) + 1
    origins = np.linspace(0, 1 - t, n_inner + 2, endpoint=True)
    starts = [int(o * n_pixel_plane) for o in origins]
    slices = [[np.s_[a:a + n_pixel_tile, b:b + n_pixel_tile] for b in starts] for a in starts]
    return origins, slices


# ---------------------------------------------------------------------------------------------------
# backends
# ---------------------------------------------------------------------------------------------------
class DeviceBackend:
    """Painting + stitching on one B200 (torch only as the device-memory container)."""

    def __init__(self, device=None):
        import torch
        from . import _lib
        _lib.load()                                        # fails loudly without the CUDA library
        if not torch.cuda.is_available():
            raise RuntimeError("baryon_painter_b200.process_SLICS needs a CUDA device; there is no CPU path")
        self.torch, self._lib = torch, _lib
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())

    def new_planes(self, n_pixel_plane):
        z = self.torch.zeros((2, n_pixel_plane, n_pixel_plane), dtype=self.torch.float64, device=self.device)
        return z

    def prepare_plane(self, raw, add, mul):
        """(raw.T + add) * mul in float32 on the device (the reference's plane preprocessing, :157-159 / :187-189):
        raw (rows, cols) float32 host array -> (cols, rows) float32 device tensor."""
        torch = self.torch
        d_raw = torch.from_numpy(np.ascontiguousarray(raw, np.float32)).to(self.device)
        out = torch.empty((raw.shape[1], raw.shape[0]), dtype=torch.float32, device=self.device)
        self._lib.plane_prepare(self.device.index or 0, d_raw.data_ptr(), raw.shape[0], raw.shape[1], add, mul, out.data_ptr(),
                                torch.cuda.current_stream(self.device).cuda_stream)
        # no synchronisation: every kernel of this backend is enqueued on torch's current stream, and torch's caching
        # allocator only hands a freed block (`d_raw`) to later work of that same stream
        return out

    def extract_tiles(self, plane, shifts, tile_relative_size, n_pixel_tile, mode, expansion_factor=1):
        """``get_tile`` + ``scipy.ndimage.zoom(tile, n_pixel_tile / side, mode=mode)`` for a list of shifts, on the
        device (csrc/bp_zoom.cu): (n, n_pixel_tile, n_pixel_tile) float32 device tensor.  The plane is uploaded once
        per plane; reference process_SLICS.py:68-83, :200, :213."""
        torch = self.torch
        if expansion_factor < 1:
            raise ValueError("Expension factors < 1 not supported.")
        # the uploaded plane is cached against a STRONG reference to the caller's array (compared with `is`): an
        # id() key could match a different, later array that reuses a freed array's address
        if isinstance(plane, torch.Tensor):                 # already on the device (prepare_plane)
            self._plane_dev, self._plane_ref = plane.contiguous(), plane
        elif getattr(self, "_plane_ref", None) is not plane:
            self._plane_dev = torch.from_numpy(np.ascontiguousarray(plane, np.float32)).to(self.device)
            self._plane_ref = plane
        n = plane.shape[0]
        side = int(n * tile_relative_size * expansion_factor)
        pad = int(n * tile_relative_size * (expansion_factor - 1) / 2)
        org = np.array([[int(n * sh[0]) - pad, int(n * sh[1]) - pad] for sh in shifts], np.int32).reshape(-1, 2)
        out = torch.empty((len(org), n_pixel_tile, n_pixel_tile), dtype=torch.float32, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        per_call = max(1, int(2 ** 30 // (8 * side * side)))           # float64 workspace <= 1 GiB
        for i0 in range(0, len(org), per_call):
            o = torch.from_numpy(org[i0:i0 + per_call]).to(self.device)
            self._lib.zoom_tiles(self._plane_dev.device.index or 0, self._plane_dev.data_ptr(), plane.shape[0], plane.shape[1],
                                 o.data_ptr(), side, o.shape[0], n_pixel_tile, mode, out[i0:i0 + per_call].data_ptr(), stream)
        return out

    def new_map(self, resolution):
        return self.torch.zeros((resolution, resolution), dtype=self.torch.float64, device=self.device)

    def zoom_accumulate(self, y_map, plane, scale, order):
        """y_map += scale * scipy.ndimage.zoom(nan_to_zero(plane), resolution / side, order=order, mode="mirror")"""
        torch = self.torch
        d = plane if isinstance(plane, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(plane, np.float64)).to(self.device)
        self._lib.zoom_accumulate(self.device.index or 0, d.data_ptr(), d.shape[0], y_map.shape[0], order, "mirror", scale,
                                  y_map.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream)

    def paint(self, painter, tiles, z, batch):
        """(n, T, T) tiles (host array or device tensor) -> painted tiles as a device tensor."""
        torch = self.torch
        n = tiles.shape[0]
        on_device = isinstance(tiles, torch.Tensor)
        out = torch.empty((n, *tiles.shape[1:]), dtype=torch.float32, device=self.device)
        if on_device and not hasattr(painter, "paint_batch_device"):
            tiles, on_device = tiles.cpu().numpy(), False
        if hasattr(painter, "paint_batch_device"):
            for i0 in range(0, n, batch):
                sl = slice(i0, min(n, i0 + batch))
                chunk = tiles[sl].contiguous() if on_device else \
                    torch.from_numpy(np.ascontiguousarray(tiles[sl], np.float32)).to(self.device)
                painter.paint_batch_device(chunk, z=z, out=out[sl])
        elif hasattr(painter, "paint_batch"):
            for i0 in range(0, n, batch):
                sl = slice(i0, min(n, i0 + batch))
                out[sl] = torch.from_numpy(np.ascontiguousarray(painter.paint_batch(tiles[sl], z=z), np.float32)).to(self.device)
        else:                                              # any duck-typed painter (reference contract)
            for i in range(n):
                p = painter.paint(input=tiles[i], z=z, transform=True, inverse_transform=True)
                out[i] = torch.from_numpy(np.ascontiguousarray(p, np.float32)).to(self.device)
        return out

    def accumulate(self, planes, painted, origins, falloff, sigma):
        org = self.torch.tensor(np.asarray(origins, np.int32).reshape(-1, 2), dtype=self.torch.int32, device=self.device)
        self._lib.stitch_accumulate(planes[0].data_ptr(), planes[1].data_ptr(), planes.shape[1], painted.data_ptr(),
                                    org.data_ptr(), painted.shape[0], painted.shape[1], falloff, sigma,
                                    self.torch.cuda.current_stream(self.device).cuda_stream)

    def reduce(self, planes_list, dst, group):
        import torch.distributed as dist
        flat = self.torch.cat([p.reshape(-1) for p in planes_list])
        dist.reduce(flat, dst=dst, op=dist.ReduceOp.SUM, group=group)
        out, o = [], 0
        for p in planes_list:
            out.append(flat[o:o + p.numel()].reshape(p.shape))
            o += p.numel()
        return out

    def finalize(self, planes):
        plane = self.torch.empty_like(planes[0])
        self._lib.stitch_finalize(planes[0].data_ptr(), planes[1].data_ptr(), plane.data_ptr(), plane.numel(),
                                  self.torch.cuda.current_stream(self.device).cuda_stream)
        return plane.cpu().numpy()

    def to_host(self, t):
        return t.cpu().numpy()


# ---------------------------------------------------------------------------------------------------
# plane sources
# ---------------------------------------------------------------------------------------------------
def _massplane_file(massplane_path, z, i, LOS):
    axes = ["xy", "xz", "yz"][i % 3]
    return os.path.join(massplane_path, f"{z:.3f}proj_half_finer_{axes}.dat_LOS{LOS}")


def _delta_file(delta_path, z, LOS):
    return os.path.join(delta_path, f"{z:.3f}delta.dat_bicubic_LOS{LOS}")


def _load_massplane(massplane_path, z, i, LOS, be=None):
    """Mass plane (reference :157-159): 12288^2 float32 after one header word, transposed, times MASS_NORM.  With a
    device backend the transpose and the scaling run on the GPU (csrc/bp_zoom.cu: plane_prepare_kernel) and the
    plane stays there for the tile extraction."""
    fn = _massplane_file(massplane_path, z, i, LOS)
    raw = np.fromfile(fn, dtype=np.float32)[1:].reshape(N_PIXEL_MASSPLANE, -1)
    if be is not None and hasattr(be, "prepare_plane"):
        return fn, be.prepare_plane(raw, 0.0, np.float32(MASS_NORM))
    return fn, raw.T * np.float32(MASS_NORM)


def _load_delta(delta_path, z, LOS, SLICS_density, be=None):
    """Delta plane (reference :187-189): 7745^2 float32, transposed, ``+= 96`` (mean of the mass plane), ``*= MASS_NORM``."""
    if SLICS_density:
        import astropy.io.fits as fits
        fn = os.path.join(delta_path, f"{z:.3f}density_LOS{LOS}.fits")
        with fits.open(fn) as hdu:
            delta = hdu[0].data.T
        return fn, delta * (MASS_NORM / 64)
    fn = _delta_file(delta_path, z, LOS)
    raw = np.fromfile(fn, dtype=np.float32).reshape(N_PIXEL_DELTA, -1)
    if be is not None and hasattr(be, "prepare_plane"):
        return fn, be.prepare_plane(raw, 96.0, MASS_NORM)
    delta = raw.T
    delta += 96          # mean of the mass plane
    delta *= MASS_NORM
    return fn, delta


def _zoom(tile, n_pixel_tile, mode):
    import scipy.ndimage
    return scipy.ndimage.zoom(tile, zoom=n_pixel_tile / tile.shape[0], mode=mode)


# ---------------------------------------------------------------------------------------------------
# process_SLICS
# ---------------------------------------------------------------------------------------------------
def process_SLICS(painter,
                  tile_size, n_pixel_tile,
                  LOS, z_SLICS, delta_size, delta_path, massplane_path, shifts_path,
                  z_slice,
                  min_tiling_overlap=0.5, verbose=True,
                  SLICS_density=False,
                  regularise=False,
                  regularise_std=None,
                  return_problematic_tiles=False,
                  plane_source=None, rank=0, world_size=1, group=None, batch=64, backend=None):
    """Paint every lightcone plane (reference :128-226) and return the list of painted planes (float64
    arrays; on ranks other than 0 of a sharded run: None).

    ``plane_source(i, kind)`` with ``kind`` in ``{"mass", "delta"}``, when given, returns the (already rescaled)
    plane ``i`` as an array instead of reading the SLICS files, and ``shifts_path`` may then be an array of
    per-plane ``(x, y)`` shifts."""
    if len(z_SLICS) != len(z_slice):
        raise ValueError("Shapes of z_SLICS and z_slice need to match!")
    if regularise_std is not None:
        # the reference branch is broken (undefined name); refuse instead of guessing
        raise NotImplementedError("regularise_std is not supported (the reference branch raises NameError)")
    be = backend if backend is not None else DeviceBackend(getattr(painter, "compute_device", None))
    say = print if (verbose and rank == 0) else (lambda *a, **k: None)

    results = []          # per plane: ("mass", device tile or None, crop args) | ("delta", planes tensor)
    item = 0              # global work-item counter: item % world_size == rank paints it
    for i in range(len(z_SLICS)):
        say(f"Processing z={z_SLICS[i]:.3f}")
        if delta_size[i] < tile_size:
            say("  Tile bigger than delta plane, using mass planes.")
            mine = (item % world_size) == rank
            item += 1
            painted = None
            if mine:
                if plane_source is not None:
                    plane = plane_source(i, "mass")
                    shift = np.asarray(shifts_path)[i]
                else:
                    shifts = np.loadtxt(os.path.join(shifts_path, f"random_shift_LOS{LOS}"))[::-1]
                    fn, plane = _load_massplane(massplane_path, z_SLICS[i], i, LOS, be if not SLICS_density else None)
                    say(f"  Loading {fn}.")
                    shift = shifts[i]
                say("  Extracting tile.")
                if hasattr(be, "extract_tiles") and not SLICS_density:
                    tile = be.extract_tiles(plane, [shift], delta_size[i] / MASSPLANE_SIZE, n_pixel_tile, "mirror",
                                            expansion_factor=tile_size / delta_size[i])
                else:
                    tile = get_tile(plane, shift=shift, tile_relative_size=delta_size[i] / MASSPLANE_SIZE,
                                    expansion_factor=tile_size / delta_size[i])
                    if SLICS_density:
                        tile = tile - tile.min()
                    tile = _zoom(tile, n_pixel_tile, "mirror")[None]
                say("  Painting on tile.")
                painted = be.to_host(be.paint(painter, tile, z_slice[i], batch))[0]
                c = (1 - delta_size[i] / tile_size) / 2
                painted = get_tile(painted, shift=(c, c), tile_relative_size=delta_size[i] / tile_size)
            results.append(("mass", painted))
            continue
        if plane_source is not None:
            delta = plane_source(i, "delta")
        else:
            fn, delta = _load_delta(delta_path, z_SLICS[i], LOS, SLICS_density, be)
        n_pixel_plane = int(delta_size[i] / tile_size * n_pixel_tile)
        origins, slices = generate_tiling(n_pixel_plane=n_pixel_plane, n_pixel_tile=n_pixel_tile, min_tile_overlap=0.5)
        say(f"  Using {len(origins)} tiles (on each side)")
        planes = be.new_planes(n_pixel_plane)
        on_device = hasattr(be, "extract_tiles")      # crop + cubic-spline zoom on the GPU
        tiles, shifts, dest = [], [], []
        for j, xs in enumerate(origins):
            for k, ys in enumerate(origins):
                mine = (item % world_size) == rank
                item += 1
                if not mine:
                    continue
                if on_device:
                    shifts.append((xs, ys))
                else:
                    tile = get_tile(delta, shift=(xs, ys), tile_relative_size=tile_size / delta_size[i])
                    tiles.append(_zoom(tile, n_pixel_tile, "reflect"))
                dest.append((slices[j][k][0].start, slices[j][k][1].start))
                say(f"    Painting on tile {j + 1}-{k + 1}")
        if dest:
            if on_device:
                tiles = be.extract_tiles(delta, shifts, tile_size / delta_size[i], n_pixel_tile, "reflect")
            else:
                tiles = np.stack(tiles).astype(np.float32, copy=False)
            painted = be.paint(painter, tiles, z_slice[i], batch)
            be.accumulate(planes, painted, dest, 0.05, 0.5)
        results.append(("delta", planes))

    # ---- assemble: one reduce of all partial planes, then plane = numerator / denominator
    delta_idx = [i for i, r in enumerate(results) if r[0] == "delta"]
    if world_size > 1:
        import torch.distributed as dist
        if delta_idx:
            reduced = be.reduce([results[i][1] for i in delta_idx], 0, group)
            for i, p in zip(delta_idx, reduced):
                results[i] = ("delta", p)
        mass = [None] * world_size
        dist.gather_object([(i, r[1]) for i, r in enumerate(results) if r[0] == "mass" and r[1] is not None],
                           mass if rank == 0 else None, dst=0, group=group)
        if rank != 0:
            return (None, []) if return_problematic_tiles else None
        for part in mass:
            for i, p in part:
                results[i] = ("mass", p)
    painted_planes = [be.finalize(r[1]) if r[0] == "delta" else np.asarray(r[1], np.float64) for r in results]
    if return_problematic_tiles:
        return painted_planes, []
    return painted_planes


# ---------------------------------------------------------------------------------------------------
# Compton-y projection
# ---------------------------------------------------------------------------------------------------
class FlatLCDM:
    """Minimal flat LCDM background (matter + Lambda) standing in for the pyccl calls of the reference
    (``comoving_angular_distance``, ``scale_factor_of_chi``); distances in Mpc.  Defaults: the SLICS cosmology
    of reference scripts/create_lightcone.py:92-98."""

    def __init__(self, Omega_m=0.2905, h=0.6898):
        self.Omega_m, self.h = float(Omega_m), float(h)
        a = np.linspace(1.0, 1 / 5.0, 4097)
        Ez = np.sqrt(self.Omega_m / a ** 3 + (1 - self.Omega_m))
        integrand = 299792.458 / (100 * self.h) / (a ** 2 * Ez)
        chi = np.concatenate([[0.0], np.cumsum(0.5 * (integrand[1:] + integrand[:-1]) * -np.diff(a))])
        self._a, self._chi = a, chi

    def comoving_distance(self, a):
        return np.interp(-np.asarray(a, float), -self._a, self._chi)

    def scale_factor_of_chi(self, chi):
        return np.interp(np.asarray(chi, float), self._chi, self._a)


def _cosmo_funcs(cosmo):
    if isinstance(cosmo, FlatLCDM):
        return cosmo.h, cosmo.comoving_distance, cosmo.scale_factor_of_chi
    import pyccl as ccl                                        # reference path (pyccl Cosmology object)
    return (cosmo.cosmo.params.h, lambda a: ccl.comoving_angular_distance(cosmo, a),
            lambda chi: ccl.scale_factor_of_chi(cosmo, chi))


def create_y_map(painted_planes, z, resolution, map_size, cosmo, order=3, verbose=True, backend=None):
    """Project painted pressure planes to a Compton-y map (reference :12-66): per plane NaN -> 0, physical
    prefactor, spline zoom to ``resolution`` (mode ``mirror``), sum.  ``backend``: a ``DeviceBackend`` runs the
    zoom-and-accumulate on the GPU (orders 3 and 5; csrc/bp_zoom.cu), default: host scipy as the reference."""
    import scipy.integrate
    import scipy.ndimage
    h, dist_of_a, a_of_chi = _cosmo_funcs(cosmo)
    slab = 252.5 / h
    d_A = np.array(dist_of_a(1 / (1 + np.array(z))), float) - slab / 2
    if d_A[0] < 0:
        d_A[0] = 0
    d_A = np.append(d_A, d_A[-1] + slab)
    theta_pix = map_size / resolution * pi / 180

    def mean_pixel_area(lo, hi):
        f = lambda chi: (chi * float(a_of_chi(chi)) * theta_pix) ** 2
        return scipy.integrate.quad(f, lo, hi)[0] / (hi - lo)

    A_pix_eff = np.array([mean_pixel_area(d_A[i], d_A[i + 1]) for i in range(len(z))])
    mpc, eV, cm = 3.086e22, 1.60218e-19, 0.01
    Xe, Xi = 1.17, 1.08
    V_c = (400 / h / 2048 * mpc / cm) ** 3                     # cell volume in cm^3
    y_fac = 8.125561e-16 * eV * mpc ** -2                      # sigma_T / m_e c^2 in Mpc^2 eV^-1
    on_device = backend is not None and hasattr(backend, "zoom_accumulate") and order in (3, 5) and \
        all(p.ndim == 2 and p.shape[0] == p.shape[1] and int(round(p.shape[0] * (resolution / p.shape[0]))) == resolution
            for p in painted_planes)
    y_map = backend.new_map(resolution) if on_device else np.zeros((resolution, resolution))
    for i, plane in enumerate(painted_planes):
        zoom_factor = resolution / plane.shape[0]
        fac = V_c * (Xe + Xi) / Xe * y_fac / A_pix_eff[i] / zoom_factor ** 2
        if verbose:
            print(f"z : {z[i]:0.3f}, plane shape: {plane.shape}, zoom_factor: {zoom_factor:0.3f}")
        if on_device:
            backend.zoom_accumulate(y_map, plane, fac, order)       # NaN -> 0, zoom, scale, += on the device
        else:
            d = np.where(np.isnan(plane), 0.0, plane) * fac
            y_map += scipy.ndimage.zoom(d, zoom=zoom_factor, order=order, mode="mirror")
    return backend.to_host(y_map) if on_device else y_map
