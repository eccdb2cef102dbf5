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
  * with ``world_size > 1`` whole planes are dealt to the ranks by cost (``plan_planes``: 1 ... 144 tiles per plane);
    a rank reads, uploads (page-locked staging, one plane ahead), paints and stitches only its own planes -- no
    collective while painting.  ``paint_lightcone`` also projects each plane into the rank's partial Compton-y map
    on its device and assembles the result with ONE reduction of that map (19 MB); ``process_SLICS`` (the
    reference's contract: the planes themselves) sends each finished plane to rank 0 point to point.

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
    ``int(n*tile_relative_size*expansion_factor)``, centred expansion, wrap-around indexing.  A crop that wraps at
    most once per axis is assembled from (up to four) contiguous slices -- the same elements as the reference's
    ``take(..., mode="wrap")``, without a gather over a 600 MB plane."""
    if expansion_factor < 1:
        raise ValueError("Expension factors < 1 not supported.")
    n = m.shape[0]
    side = int(n * tile_relative_size * expansion_factor)
    pad = int(n * tile_relative_size * (expansion_factor - 1) / 2)
    r0 = (int(n * shift[0]) - pad) % m.shape[0]
    c0 = (int(n * shift[1]) - pad) % m.shape[1]
    if side <= m.shape[0] and side <= m.shape[1]:
        rs = [(r0, min(m.shape[0], r0 + side))] + ([(0, r0 + side - m.shape[0])] if r0 + side > m.shape[0] else [])
        cs = [(c0, min(m.shape[1], c0 + side))] + ([(0, c0 + side - m.shape[1])] if c0 + side > m.shape[1] else [])
        if len(rs) == 1 and len(cs) == 1:
            return np.array(m[rs[0][0]:rs[0][1], cs[0][0]:cs[0][1]])
        return np.block([[m[a:b, c:d] for c, d in cs] for a, b in rs])
    rows = (r0 + np.arange(side)) % m.shape[0]
    cols = (c0 + np.arange(side)) % m.shape[1]
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
        import threading
        self.torch, self._lib = torch, _lib
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self._copy_stream, self._staged, self._keep, self._stage_lock = None, {}, None, threading.Lock()

    def new_planes(self, n_pixel_plane):
        z = self.torch.zeros((2, n_pixel_plane, n_pixel_plane), dtype=self.torch.float64, device=self.device)
        return z

    def stage(self, host):
        """Host array -> device tensor through a page-locked buffer and an asynchronous copy on a side stream (the
        caller's stream only waits for the copy's event): the upload of plane k+1 overlaps the painting of plane k
        when this runs on ``_PlaneFeed``'s worker thread.  Device tensors pass through."""
        torch = self.torch
        if isinstance(host, torch.Tensor):
            return host
        host = np.ascontiguousarray(host)
        with self._stage_lock:
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(self.device)
            pinned = torch.from_numpy(host)
            if not pinned.is_pinned():                      # (baryon_painter_b200.pinned_empty arrays are used in place)
                pinned = torch.empty(host.shape, dtype=getattr(torch, host.dtype.name), pin_memory=True)
                pinned.numpy()[...] = host
            with torch.cuda.stream(self._copy_stream):
                dev = pinned.to(self.device, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
            self._staged[id(dev)] = (pinned, ev)            # keep the pinned source alive until the copy is consumed
        return dev

    def _consume(self, t):
        """make the current stream wait for the asynchronous upload of ``t`` (if it came from ``stage``)"""
        item = self._staged.pop(id(t), None)
        if item is not None:
            self.torch.cuda.current_stream(self.device).wait_event(item[1])
            t.record_stream(self.torch.cuda.current_stream(self.device))
            self._keep = item[0]                            # the pinned buffer may go once the NEXT plane is consumed
        return t

    def prepare_plane(self, raw, add, mul):
        """(raw.T + add) * mul in float32 on the device (the reference's plane preprocessing, :157-159 / :187-189):
        raw (rows, cols) float32 host array -> (cols, rows) float32 device tensor."""
        torch = self.torch
        d_raw = self._consume(self.stage(np.asarray(raw, np.float32)))
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
        if isinstance(plane, torch.Tensor):                 # already on the device (prepare_plane / stage)
            self._consume(plane)
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

    @staticmethod
    def _host_staged(group):
        """gloo has no device point-to-point / reduce: such a group (CPU-side tests, a box without NCCL) is served
        through host copies; NCCL groups exchange device memory directly"""
        import torch.distributed as dist
        return dist.get_backend(group) == "gloo"

    def reduce(self, tensors, dst, group):
        """ONE sum-reduction of a list of same-dtype device tensors to rank ``dst`` (NCCL over NVLink on a GPU box)"""
        import torch.distributed as dist
        flat = self.torch.cat([p.reshape(-1) for p in tensors])
        if self._host_staged(group):
            host = flat.cpu()
            dist.reduce(host, dst=dst, op=dist.ReduceOp.SUM, group=group)
            flat.copy_(host)
        else:
            dist.reduce(flat, dst=dst, op=dist.ReduceOp.SUM, group=group)
        out, o = [], 0
        for p in tensors:
            out.append(flat[o:o + p.numel()].reshape(p.shape))
            o += p.numel()
        return out

    def send(self, plane, dst, group):
        import torch.distributed as dist
        plane = plane.contiguous()
        dist.send(plane.cpu() if self._host_staged(group) else plane, dst=dst, group=group)

    def recv(self, shape, src, group):
        import torch.distributed as dist
        staged = self._host_staged(group)
        t = self.torch.empty(shape, dtype=self.torch.float64, device="cpu" if staged else self.device)
        dist.recv(t, src=src, group=group)
        return t.to(self.device) if staged else t

    def finalize_device(self, planes):
        """plane = numerator / denominator (reference :222), staying on the device"""
        plane = self.torch.empty_like(planes[0])
        self._lib.stitch_finalize(planes[0].data_ptr(), planes[1].data_ptr(), plane.data_ptr(), plane.numel(),
                                  self.torch.cuda.current_stream(self.device).cuda_stream)
        return plane

    def finalize(self, planes):
        return self.finalize_device(planes).cpu().numpy()

    def crop(self, tile, shift, tile_relative_size):
        """``get_tile`` (periodic crop, expansion factor 1) of a device tile, as float64"""
        n = tile.shape[0]
        side = int(n * tile_relative_size)
        rows = (int(n * shift[0]) + self.torch.arange(side, device=self.device)) % tile.shape[0]
        cols = (int(n * shift[1]) + self.torch.arange(side, device=self.device)) % tile.shape[1]
        return tile[rows][:, cols].to(self.torch.float64)

    def to_host(self, t):
        return t.cpu().numpy() if isinstance(t, self.torch.Tensor) else np.asarray(t)


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


def _massplane_view(massplane_path, z, i, LOS):
    """The mass plane as the reference indexes it (``raw[1:].reshape(12288, -1).T``, :157-158) WITHOUT reading the
    file: a transposed view of a memory map, so that ``get_tile`` touches only the pages of its crop."""
    fn = _massplane_file(massplane_path, z, i, LOS)
    mm = np.memmap(fn, dtype=np.float32, mode="r", offset=4)
    return mm.reshape(N_PIXEL_MASSPLANE, -1).T


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
def plan_planes(costs, world_size):
    """Owner rank of every plane: longest-processing-time-first list scheduling of the per-plane costs
    (``plane_cost``; a SLICS line of sight has 1, 1, 4, 9, ..., 144 tiles per plane, reference
    process_SLICS.py:177-220).  Deterministic: ties go to the lower rank / lower plane index."""
    owner, load = [0] * len(costs), [0.0] * max(1, world_size)
    for i in sorted(range(len(costs)), key=lambda k: (-costs[k], k)):
        r = min(range(len(load)), key=lambda q: (load[q], q))
        owner[i] = r
        load[r] += costs[i]
    return owner


def _plane_geometry(delta_size_i, tile_size, n_pixel_tile):
    """(kind, tiles per side, pixels per side of the painted plane) of one lightcone slice (reference :149, :191-194)."""
    if delta_size_i < tile_size:
        return "mass", 1, int(n_pixel_tile * (delta_size_i / tile_size))
    n_pixel_plane = int(delta_size_i / tile_size * n_pixel_tile)
    origins, _ = generate_tiling(n_pixel_plane=n_pixel_plane, n_pixel_tile=n_pixel_tile, min_tile_overlap=0.5)
    return "delta", len(origins), n_pixel_plane


def plane_cost(kind, tiles_per_side, delta_size_i, tile_size):
    """Device seconds one slice costs its owner, for ``plan_planes``: painting is per tile (36 us on a B200); the tile
    extraction runs the cubic-spline prefilter over every source pixel of every (overlapping) crop, 1.6e-11 s per crop
    pixel (measured: 43 ms of an 88 ms line of sight with 2.7 G crop pixels, bench.py --config lightcone), which is
    nearly the same for every delta plane however many tiles it is cut into; plus the upload of the plane itself."""
    n = tiles_per_side ** 2
    if kind == "mass":
        side = N_PIXEL_MASSPLANE * delta_size_i / MASSPLANE_SIZE * (tile_size / delta_size_i)
        upload = side * side * 4 / 50e9 + 10e-3            # the crop only; ~10 ms of host indexing to make it
    else:
        side = N_PIXEL_DELTA * tile_size / delta_size_i
        upload = N_PIXEL_DELTA ** 2 * 4 / 50e9
    return n * (36e-6 + 1.6e-11 * side * side) + upload


class _MassCrop:
    """the ``get_tile`` crop of a mass plane (reference :165-167), taken on the host before the upload"""

    def __init__(self, tile):
        self.tile = tile


class _PlaneFeed:
    """Produces this rank's planes one ahead of the painter: while plane k is being painted, a worker thread reads
    plane k+1 (file or ``plane_source``) and hands it to ``backend.stage`` (page-locked buffer + asynchronous upload
    on a side stream for the device backend) -- the reference reads each file synchronously before touching it
    (process_SLICS.py:157-159, :187-189)."""

    def __init__(self, indices, load):
        import threading
        self._idx, self._load = list(indices), load
        self._slots = {}
        self._lock = threading.Condition()
        self._next = 0
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def _run(self):
        for k, i in enumerate(self._idx):
            with self._lock:
                while k > self._next + 1:                    # at most one plane ahead of the consumer
                    self._lock.wait()
            try:
                item = ("ok", self._load(i))
            except BaseException as e:                       # surfaced in get()
                item = ("err", e)
            with self._lock:
                self._slots[i] = item
                self._lock.notify_all()

    def get(self, i):
        with self._lock:
            while i not in self._slots:
                self._lock.wait()
            kind, val = self._slots.pop(i)
            self._next += 1
            self._lock.notify_all()
        if kind == "err":
            raise val
        return val


class _StageClock:
    """Optional per-stage wall clock of the lightcone loop (synchronises the device around every stage, so it is for
    attribution runs only -- ``bench.py --config lightcone --stages``)."""

    def __init__(self, sync):
        self.t, self._sync = {}, sync

    def __call__(self, name):
        clock = self

        class _Ctx:
            def __enter__(self):
                import time
                clock._sync()
                self.t0 = time.perf_counter()

            def __exit__(self, *a):
                import time
                clock._sync()
                clock.t[name] = clock.t.get(name, 0.0) + time.perf_counter() - self.t0
        return _Ctx()


class _NoClock:
    def __call__(self, name):
        import contextlib
        return contextlib.nullcontext()


def _paint_plane(i, plane, painter, be, tile_size, n_pixel_tile, delta_size, z_slice, shift, SLICS_density, batch, say,
                 tile_filter=None, clock=_NoClock()):
    """One lightcone slice -> its painted plane (backend array: device tensor / numpy, float64).  Mass-plane branch
    (reference :149-176): one tile, expanded crop, painted, cropped back.  Delta branch (:177-220): tiled, painted in
    batches, blended with the Gaussian-edge weights.  ``tile_filter(j, k)``: paint only those tiles (partial planes)."""
    if delta_size[i] < tile_size:
        say("  Extracting tile.")
        if isinstance(plane, _MassCrop):
            # the periodic crop was taken on the host (only ~4 % of a 12288^2 mass plane is needed: 24 MB instead of
            # 604 MB cross PCIe); the cubic-spline resampling runs on the device
            if hasattr(be, "extract_tiles"):
                with clock("extract"):
                    tile = be.extract_tiles(plane.tile, [(0.0, 0.0)], 1.0, n_pixel_tile, "mirror")
            else:
                tile = _zoom(be.to_host(plane.tile), n_pixel_tile, "mirror")[None]
        elif hasattr(be, "extract_tiles") and not SLICS_density:
            with clock("extract"):
                tile = be.extract_tiles(plane, [shift], delta_size[i] / MASSPLANE_SIZE, n_pixel_tile, "mirror",
                                        expansion_factor=tile_size / delta_size[i])
        else:
            plane = be.to_host(plane)
            tile = get_tile(plane, shift=shift, tile_relative_size=delta_size[i] / MASSPLANE_SIZE,
                            expansion_factor=tile_size / delta_size[i])
            if SLICS_density:
                tile = tile - tile.min()
            tile = _zoom(tile, n_pixel_tile, "mirror")[None]
        say("  Painting on tile.")
        with clock("paint"):
            painted = be.paint(painter, tile, z_slice[i], batch)
        c = (1 - delta_size[i] / tile_size) / 2
        return be.crop(painted[0], shift=(c, c), tile_relative_size=delta_size[i] / tile_size), None
    n_pixel_plane = int(delta_size[i] / tile_size * n_pixel_tile)
    origins, slices = generate_tiling(n_pixel_plane=n_pixel_plane, n_pixel_tile=n_pixel_tile, min_tile_overlap=0.5)
    say(f"  Using {len(origins)} tiles (on each side)")
    planes = be.new_planes(n_pixel_plane)
    on_device = hasattr(be, "extract_tiles")      # crop + cubic-spline zoom on the GPU
    tiles, shifts, dest = [], [], []
    for j, xs in enumerate(origins):
        for k, ys in enumerate(origins):
            if tile_filter is not None and not tile_filter(j, k):
                continue
            if on_device:
                shifts.append((xs, ys))
            else:
                tile = get_tile(be.to_host(plane), shift=(xs, ys), tile_relative_size=tile_size / delta_size[i])
                tiles.append(_zoom(tile, n_pixel_tile, "reflect"))
            dest.append((slices[j][k][0].start, slices[j][k][1].start))
            say(f"    Painting on tile {j + 1}-{k + 1}")
    if dest:
        if on_device:
            with clock("extract"):
                tiles = be.extract_tiles(plane, shifts, tile_size / delta_size[i], n_pixel_tile, "reflect")
        else:
            tiles = np.stack(tiles).astype(np.float32, copy=False)
        with clock("paint"):
            painted = be.paint(painter, tiles, z_slice[i], batch)
        with clock("stitch"):
            be.accumulate(planes, painted, dest, 0.05, 0.5)
    return None, planes


def _run_lightcone(painter, tile_size, n_pixel_tile, LOS, z_SLICS, delta_size, delta_path, massplane_path, shifts_path,
                   z_slice, verbose, SLICS_density, regularise_std, plane_source, rank, world_size, batch, backend,
                   on_plane, clock=_NoClock()):
    """Shared driver of ``process_SLICS`` and ``paint_lightcone``: planes are dealt to the ranks whole (``plan_planes``),
    every rank loads ONLY its planes (one ahead, ``_PlaneFeed``), paints and stitches them on its device and calls
    ``on_plane(i, plane)`` with each finished float64 plane.  No collective in here."""
    if len(z_SLICS) != len(z_slice):
        raise ValueError("Shapes of z_SLICS and z_slice need to match!")
    if regularise_std is not None:
        # the reference branch is broken (undefined name); refuse instead of guessing
        raise NotImplementedError("regularise_std is not supported (the reference branch raises NameError)")
    be = backend if backend is not None else DeviceBackend(getattr(painter, "compute_device", None))
    say = print if (verbose and rank == 0) else (lambda *a, **k: None)
    geom = [_plane_geometry(delta_size[i], tile_size, n_pixel_tile) for i in range(len(z_SLICS))]
    owner = plan_planes([plane_cost(g[0], g[1], delta_size[i], tile_size) for i, g in enumerate(geom)], world_size)
    mine = [i for i in range(len(z_SLICS)) if owner[i] == rank]

    shifts_cache = []

    def mass_shift(i):
        if plane_source is not None:
            return np.asarray(shifts_path)[i]
        if not shifts_cache:
            shifts_cache.append(np.loadtxt(os.path.join(shifts_path, f"random_shift_LOS{LOS}"))[::-1])
        return shifts_cache[0][i]

    def load(i):
        kind = geom[i][0]
        stage = be.stage if hasattr(be, "stage") else (lambda a: a)
        if kind == "mass" and not SLICS_density:
            # crop on the host, upload the crop (see _MassCrop)
            if plane_source is not None:
                plane = plane_source(i, kind)
                if not isinstance(plane, np.ndarray):
                    return plane                          # a device tensor: crop it where it is
            else:
                plane = _massplane_view(massplane_path, z_SLICS[i], i, LOS)
            crop = get_tile(plane, shift=mass_shift(i), tile_relative_size=delta_size[i] / MASSPLANE_SIZE,
                            expansion_factor=tile_size / delta_size[i])
            if plane_source is None:
                crop = crop * np.float32(MASS_NORM)       # the reference scales the whole plane first: same float32 products
            return _MassCrop(stage(np.ascontiguousarray(crop, np.float32)))
        if plane_source is not None:
            return stage(plane_source(i, kind))
        if kind == "mass":
            return _load_massplane(massplane_path, z_SLICS[i], i, LOS, None)[1]
        return _load_delta(delta_path, z_SLICS[i], LOS, SLICS_density, be)[1]

    feed = _PlaneFeed(mine, load)
    for i in mine:
        say(f"Processing z={z_SLICS[i]:.3f}")
        with clock("wait_plane"):
            plane = feed.get(i)
        shift = None
        if geom[i][0] == "mass":
            say("  Tile bigger than delta plane, using mass planes.")
            shift = mass_shift(i)
        cropped, planes = _paint_plane(i, plane, painter, be, tile_size, n_pixel_tile, delta_size, z_slice, shift,
                                       SLICS_density, batch, say, clock=clock)
        with clock("project"):
            on_plane(i, cropped if planes is None else be.finalize_device(planes))
        del plane
    return be, owner, geom


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
    per-plane ``(x, y)`` shifts.

    Sharding (``world_size > 1``): whole planes are dealt to the ranks by cost (``plan_planes``); a rank reads, paints
    and stitches only its own planes; afterwards every finished plane is sent to rank 0 point to point.  Callers
    that want the Compton-y map rather than the planes should use ``paint_lightcone``: it projects on the owning
    rank and moves one map instead of every plane."""
    done = {}
    be, owner, geom = _run_lightcone(painter, tile_size, n_pixel_tile, LOS, z_SLICS, delta_size, delta_path,
                                     massplane_path, shifts_path, z_slice, verbose, SLICS_density, regularise_std,
                                     plane_source, rank, world_size, batch, backend, lambda i, p: done.__setitem__(i, p))
    if world_size > 1:
        for i in range(len(z_SLICS)):                       # plane i: owner -> rank 0, same order on every rank
            if owner[i] == 0:
                continue
            if rank == owner[i]:
                be.send(done.pop(i), 0, group)
            elif rank == 0:
                done[i] = be.recv((geom[i][2], geom[i][2]), owner[i], group)
        if rank != 0:
            return (None, []) if return_problematic_tiles else None
    painted_planes = [np.asarray(be.to_host(done[i]), np.float64) for i in range(len(z_SLICS))]
    if return_problematic_tiles:
        return painted_planes, []
    return painted_planes


def paint_lightcone(painter, tile_size, n_pixel_tile, LOS, z_SLICS, delta_size, delta_path, massplane_path, shifts_path,
                    z_slice, resolution, map_size, cosmo, order=5, verbose=True, SLICS_density=False, plane_source=None,
                    rank=0, world_size=1, group=None, batch=64, backend=None, drop_planes=None, keep_planes=False,
                    stage_times=None):
    """``process_SLICS`` + ``create_y_map`` (reference scripts/create_lightcone.py:106-128) as one sharded pass:
    every rank paints and stitches its own planes and adds each, with the plane's physical prefactor, to ITS partial
    Compton-y map on its device (the projection is linear in the planes); ONE reduction of the ``resolution``^2
    float64 map (19 MB at 1549^2, NCCL over NVLink on a GPU box) assembles the result on rank 0 -- no painted plane
    ever leaves the GPU that made it.  Returns ``y_map`` on rank 0 (None elsewhere); with ``drop_planes=n`` a pair
    ``(y_map, y_map_without_the_first_n_planes)``; ``keep_planes=True`` adds the rank's own planes (dict index ->
    host array) as the last element."""
    n = len(z_SLICS)
    sizes = [_plane_geometry(delta_size[i], tile_size, n_pixel_tile)[2] for i in range(n)]
    fac = y_map_factors(z_SLICS, resolution, map_size, cosmo, sizes)
    fac_drop = y_map_factors(z_SLICS[drop_planes:], resolution, map_size, cosmo, sizes[drop_planes:]) \
        if drop_planes is not None else None
    be = backend if backend is not None else DeviceBackend(getattr(painter, "compute_device", None))
    maps = [be.new_map(resolution)] + ([be.new_map(resolution)] if drop_planes is not None else [])
    kept = {}

    def on_plane(i, plane):
        if verbose and rank == 0:
            print(f"z : {z_SLICS[i]:0.3f}, plane shape: {tuple(plane.shape)}, zoom_factor: {resolution / plane.shape[0]:0.3f}")
        be.zoom_accumulate(maps[0], plane, fac[i], order)
        if drop_planes is not None and i >= drop_planes:
            be.zoom_accumulate(maps[1], plane, fac_drop[i - drop_planes], order)
        if keep_planes:
            kept[i] = np.asarray(be.to_host(plane), np.float64)

    clock = _NoClock()
    if stage_times is not None and hasattr(be, "torch"):
        clock = _StageClock(lambda: be.torch.cuda.synchronize(be.device))
    _run_lightcone(painter, tile_size, n_pixel_tile, LOS, z_SLICS, delta_size, delta_path, massplane_path, shifts_path,
                   z_slice, verbose, SLICS_density, None, plane_source, rank, world_size, batch, be, on_plane, clock=clock)
    if world_size > 1:
        with clock("reduce"):
            maps = be.reduce(maps, 0, group)                # the one collective of the run
    if stage_times is not None and isinstance(clock, _StageClock):
        stage_times.update(clock.t)
    out = None
    if rank == 0:
        out = tuple(np.asarray(be.to_host(m), np.float64) for m in maps)
        out = out[0] if drop_planes is None else out
    if keep_planes:
        return (out, kept) if drop_planes is None else (*(out or (None, None)), kept)
    return out


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


def y_map_factors(z, resolution, map_size, cosmo, plane_sizes):
    """Per-plane prefactor of the Compton-y projection (reference :14-62): cell volume, electron / ion number
    ratios, sigma_T / m_e c^2, the slab's mean pixel area and the flux-conserving 1 / zoom^2."""
    import scipy.integrate
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
    return [V_c * (Xe + Xi) / Xe * y_fac / A_pix_eff[i] / (resolution / plane_sizes[i]) ** 2 for i in range(len(z))]


def create_y_map(painted_planes, z, resolution, map_size, cosmo, order=3, verbose=True, backend=None):
    """Project painted pressure planes to a Compton-y map (reference :12-66): per plane NaN -> 0, physical
    prefactor, spline zoom to ``resolution`` (mode ``mirror``), sum.  ``backend``: a ``DeviceBackend`` runs the
    zoom-and-accumulate on the GPU (orders 3 and 5; csrc/bp_zoom.cu), default: host scipy as the reference."""
    import scipy.ndimage
    facs = y_map_factors(z, resolution, map_size, cosmo, [p.shape[0] for p in painted_planes])
    on_device = backend is not None and hasattr(backend, "zoom_accumulate") and order in (3, 5) and \
        all(p.ndim == 2 and p.shape[0] == p.shape[1] and int(round(p.shape[0] * (resolution / p.shape[0]))) == resolution
            for p in painted_planes)
    y_map = backend.new_map(resolution) if on_device else np.zeros((resolution, resolution))
    for i, plane in enumerate(painted_planes):
        zoom_factor = resolution / plane.shape[0]
        fac = facs[i]
        if verbose:
            print(f"z : {z[i]:0.3f}, plane shape: {plane.shape}, zoom_factor: {zoom_factor:0.3f}")
        if on_device:
            backend.zoom_accumulate(y_map, plane, fac, order)       # NaN -> 0, zoom, scale, += on the device
        else:
            d = np.where(np.isnan(plane), 0.0, plane) * fac
            y_map += scipy.ndimage.zoom(d, zoom=zoom_factor, order=order, mode="mirror")
    return backend.to_host(y_map) if on_device else y_map
