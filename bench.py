#!/usr/bin/env python
"""Headline benchmark: tiles/s painted by the fiducial CVAE (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision bf16|fp32] [--tiles 256]
    python bench.py --impl reference ...      # the reference's CPU paint path on the host cores

One step = one pass of the paint hot path over one batch of `--tiles` synthetic 512x512 DM tiles
(seeded log-normal fields, fixed eps latents, z = 0) with seeded synthetic weights of the fiducial
architecture (the trained blob is not in the reference checkout).  `value` times the device path
with inputs resident in HBM (C ABI `bp_cvae_paint`, device pointers, CUDA events on the launching
stream); `e2e` times the public `CVAEPainter.paint_batch` with host buffers, host<->device copies
inside the timed region.  Under torchrun every rank paints its own `--tiles` tiles (weak scaling;
the path has no collective), the time is the max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOPS_PER_TILE = 20.254e9          # SURVEY.md section 8d / App. A (2*MACs of the 25 layers)
TILE = 512


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"bf16_tflops": p["bf16_tflops_sustained"], "bf16_tflops_burst": p["bf16_tflops"],
                "hbm_gbs": p["hbm_gbs"], "source": "measured (MEASURED_PEAKS.json, sustained)"}
    except Exception:
        return {"bf16_tflops": 1400.0, "bf16_tflops_burst": 1590.0, "hbm_gbs": 6650.0,
                "source": "fallback (B200_PROFILING.md)"}


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md): NVML every 5 ms when
    the bindings are there (a timed region is ~0.2 s), else the recipe's nvidia-smi query every 0.2 s."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NVML_REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()
        self.sm, self.mx, self.reasons, self.source = [], [], set(), "nvidia-smi"
        self._h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.mx.append(float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)))
            self.source = "nvml"
        except Exception:
            self._h = None

    def _sample_nvml(self):
        nv = self._nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
        get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        mask = int(get(self._h))
        for bit, name in self.NVML_REASONS.items():
            if mask & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
        if out.returncode != 0:
            return
        r = [c.strip() for c in out.stdout.strip().split(",")]
        if r and r[0].replace(".", "").isdigit():
            self.sm.append(float(r[0]))
        if len(r) > 1 and r[1].replace(".", "").isdigit():
            self.mx.append(float(r[1]))
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        if len(r) >= 7:
            self.reasons.update(names[i] for i in range(4) if r[3 + i] == "Active")

    def run(self):
        while not self._stop_evt.is_set():
            try:
                if self._h is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._stop_evt.wait(0.005 if self._h is not None else 0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": self.source}


def make_cpu_arm(n_tiles, threads):
    """Everything the CPU arm needs, built OUTSIDE any timed region: the oracle model (port of the reference's
    torch-CPU modules, bit-identical to them in the build container), the synthetic tiles and the latents."""
    import torch
    from oracle.cvae_oracle import CVAEOracle
    from baryon_painter_b200 import arch, synthetic, transforms
    torch.set_num_threads(threads)
    A = arch.fiducial_cvae_architecture(TILE)
    orc = CVAEOracle(A, synthetic.synthetic_cvae_state_dict(A, seed=0))
    return orc, transforms.fiducial_stats(), synthetic.synthetic_dm_tiles(n_tiles, TILE), synthetic.synthetic_latents(n_tiles)


def cpu_paint_loop(arm, n_tiles):
    """The reference's CPU paint path as the reference would run a batch: a Python loop of batch-1 paint() calls.
    Only the paint() calls are timed."""
    orc, stats, tiles, eps = arm
    t0 = time.perf_counter()
    for i in range(n_tiles):
        orc.paint(tiles[i % len(tiles)], 0.0, stats, eps=eps[i % len(tiles):i % len(tiles) + 1])
    return time.perf_counter() - t0


def cpu_baseline(n_tiles, threads, warmup=1):
    arm = make_cpu_arm(n_tiles, threads)
    cpu_paint_loop(arm, warmup)
    dt = cpu_paint_loop(arm, n_tiles)
    return n_tiles / dt, dt


def workload_config(n, world):
    """The `config` object of both arms: BASELINE.json configs[1] at `n` tiles per GPU."""
    return {"workload": "fiducial CVAE batched paint, %d synthetic 512x512 tiles per GPU, fixed eps "
                        "latents, z=0 (BASELINE.json configs[1])" % n,
            "weights": "seeded synthetic state_dict, fiducial architecture (trained blob absent)",
            "l2": "inputs+outputs per step %.0f MB > 126 MB L2" % (2 * n * TILE * TILE * 4 / 1e6),
            "parallelism": "tiles sharded over %d GPU(s), no collective" % world}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count()
    sample = max(1, min(args.tiles, 64))
    arm = make_cpu_arm(sample, threads)            # model + inputs built once, outside the timed loop
    for _ in range(args.warmup):
        cpu_paint_loop(arm, 1)
    dt, n = 0.0, 0
    for _ in range(args.steps):
        dt += cpu_paint_loop(arm, sample)
        n += sample
    v = n / dt
    print(json.dumps({
        "impl": "reference", "metric": "tiles/sec painted (fiducial CVAE)", "value": v, "unit": "tiles/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.tiles, max(1, args.gpus)),
        "cpu_baseline": {"value": v, "unit": "tiles/s", "cores": threads, "kind": "port",
                         "sample": "%d batch-1 paint() calls per step (oracle port of the reference torch-CPU "
                                   "path, bit-identical to it in the build container)" % sample},
        "e2e": {"value": v, "unit": "tiles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


LIGHTCONE_FIELDS = ("metric", "value", "unit", "n_gpus", "scaling", "ms_per_step", "tiles_per_s", "e2e",
                    "stages_s_max_over_ranks", "map_check")


class _ZeroEps:
    """painter wrapper: eps = 0, so that the painted line of sight does not depend on per-rank call counters"""

    def __init__(self, painter):
        self.p = painter
        self.compute_device = painter.compute_device

    def paint_batch_device(self, tiles, z=0.0, out=None):
        import torch
        eps = torch.zeros((tiles.shape[0], *self.p.model.dim_z[1:]), dtype=torch.float32, device=tiles.device)
        return self.p.paint_batch_device(tiles, z=z, eps=eps, out=out)


def measure_lightcone(args, rank, world, local, precision):
    """BASELINE.json configs[4]: one full synthetic line of sight -- 15 lightcone slices (2 mass planes of 12288^2,
    13 delta planes of 7745^2 pixels, 781 tiles of 512^2) tiled, painted, stitched and projected to a 1549^2 Compton-y
    map (reference scripts/create_lightcone.py:106-128), planes dealt to the N ranks by cost, one NCCL reduction of
    the map.  A step = one line of sight; planes are page-locked host arrays made before the timed region (the
    stand-in for the SLICS files: disk reads are not timed, the host->device upload of every plane is)."""
    import torch
    import torch.distributed as dist
    import baryon_painter_b200 as bp
    from baryon_painter_b200 import process_SLICS as ps
    from baryon_painter_b200.painter import CVAEPainter
    n_z, h, tile_size, n_pixel_tile = args.planes, 0.6898, 100.0, TILE
    cosmo = ps.FlatLCDM(Omega_m=0.2905, h=h)
    chi = 252.5 / h * (np.arange(n_z) + 0.5)
    z_SLICS = 1 / cosmo.scale_factor_of_chi(chi) - 1
    delta_size = chi * h * 10 / 180 * np.pi
    z_slice = np.array([1 / float(cosmo.scale_factor_of_chi(252.5 / h * i)) - 1 for i in range(n_z)])
    shifts = np.random.default_rng(0).random((n_z, 2))
    geom = [ps._plane_geometry(delta_size[i], tile_size, n_pixel_tile) for i in range(n_z)]
    owner = ps.plan_planes([ps.plane_cost(g[0], g[1], delta_size[i], tile_size) for i, g in enumerate(geom)], world)
    n_tiles = sum(g[1] ** 2 for g in geom)
    scale = args.plane_scale              # 1.0 = the SLICS plane sizes
    planes = {}
    for i in range(n_z):                  # this rank's planes only, page-locked, block-constant log-normal densities
        if owner[i] != rank:
            continue
        # smooth, strictly positive, periodic log-normal density (mean ~1): a 512^2 Gaussian-filtered field tiled up to
        # the plane size (the cubic-spline resampling of a field with sharp edges overshoots below zero, and the
        # shift-log transform of a negative density is NaN)
        import scipy.ndimage
        n = int((ps.N_PIXEL_MASSPLANE if geom[i][0] == "mass" else ps.N_PIXEL_DELTA) * scale)
        g = scipy.ndimage.gaussian_filter(np.random.default_rng(100 + i).standard_normal((512, 512)), 3.0, mode="wrap")
        base = np.exp(g / g.std() - 0.5).astype(np.float32)
        reps = (n + 511) // 512
        buf = bp.pinned_empty((n, n))
        buf[...] = np.tile(base, (reps, reps))[:n, :n]
        planes[i] = buf
    painter = CVAEPainter.synthetic(tile_size=TILE, seed=0, compute_device="cuda:%d" % local, precision=precision,
                                    max_batch=args.tiles)
    be = ps.DeviceBackend("cuda:%d" % local)
    kw = dict(tile_size=tile_size, n_pixel_tile=n_pixel_tile, LOS=0, z_SLICS=z_SLICS, delta_size=delta_size, delta_path=None,
              massplane_path=None, shifts_path=shifts, z_slice=z_slice, resolution=1549, map_size=10.0, cosmo=cosmo, order=5,
              verbose=False, plane_source=lambda i, kind: planes[i], rank=rank, world_size=world, batch=args.tiles, backend=be)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    y_map = None
    lc_steps = max(1, min(args.steps, 3)) if args.config != "lightcone" else args.steps
    for _ in range(max(1, min(args.warmup, 1 if args.config != "lightcone" else args.warmup))):
        y_map = ps.paint_lightcone(painter, **kw)
    times = []
    sampler = ClockSampler(local)
    sampler.start()
    _lib_launches = __import__("baryon_painter_b200._lib", fromlist=["x"])
    _lib_launches.launch_count(reset=True)
    for _ in range(lc_steps):
        barrier()
        t0 = time.perf_counter()
        y_map = ps.paint_lightcone(painter, **kw)
        barrier()
        times.append(time.perf_counter() - t0)
    launches = _lib_launches.launch_count(reset=True)
    clocks = sampler.stop()
    stages = {}
    ps.paint_lightcone(painter, stage_times=stages, **kw)          # one more pass with per-stage synchronised clocks
    y_fixed = ps.paint_lightcone(_ZeroEps(painter), **kw)          # latents = prior mean: the same map at every N
    # the same line of sight with this rank's planes already resident in HBM (the main metric's `value` convention:
    # inputs on the device when the timed region starts; the timing above, uploads inside, is its `e2e`)
    dev_planes = {i: torch.from_numpy(p).to("cuda:%d" % local) for i, p in planes.items()}
    kw_res = dict(kw, plane_source=lambda i, kind: dev_planes[i])
    ps.paint_lightcone(painter, **kw_res)
    times_res = []
    for _ in range(lc_steps):
        barrier()
        t0 = time.perf_counter()
        ps.paint_lightcone(painter, **kw_res)
        barrier()
        times_res.append(time.perf_counter() - t0)
    y_fixed_res = ps.paint_lightcone(_ZeroEps(painter), **kw_res)
    del dev_planes
    t = torch.tensor([float(np.median(times)), float(np.sum(times)), float(np.median(times_res))], dtype=torch.float64, device="cuda")
    st = torch.tensor([stages.get(k, 0.0) for k in ("wait_plane", "extract", "paint", "stitch", "project", "reduce")],
                      dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(st, op=dist.ReduceOp.MAX)
    if rank != 0:
        return None
    med, med_res = float(t[0]), float(t[2])
    assert y_map is not None and y_map.shape == (1549, 1549) and np.isfinite(y_map).all() and y_map.max() > 0, \
        (np.isfinite(y_map).mean(), float(np.nanmax(y_map)))
    # planes resident on the device take the device-side crop of the mass planes instead of the host-side one: same map
    res_err = float(np.sqrt(((y_fixed_res - y_fixed) ** 2).sum() / (y_fixed ** 2).sum()))
    assert res_err <= 1e-9, res_err
    # uploaded per line of sight: whole delta planes, of the mass planes only the (expanded) tile crop
    plane_bytes = int(sum(((ps.N_PIXEL_MASSPLANE * tile_size / ps.MASSPLANE_SIZE if g[0] == "mass" else ps.N_PIXEL_DELTA)
                           * scale) ** 2 * 4 for g in geom))
    loads = [sum(g[1] ** 2 for g, o in zip(geom, owner) if o == r) for r in range(world)]
    return {
        "metric": "lightcone lines of sight/sec (create_lightcone: %d planes, %d tiles -> 1549^2 y map)" % (n_z, n_tiles),
        "value": 1.0 / med_res, "unit": "LOS/s", "n_gpus": world, "steps": lc_steps, "warmup": 1,
        "ms_per_step": 1e3 * med_res, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": {"fp16": "f16", "bf16": "bf16"}.get(precision, precision), "data": "synthetic",
        "tiles_per_s": n_tiles / med_res,
        "e2e": {"value": 1.0 / med, "unit": "LOS/s", "ms_per_step": 1e3 * med, "h2d_bytes_per_step": plane_bytes,
                "d2h_bytes_per_step": 1549 * 1549 * 8,
                "api": "process_SLICS.paint_lightcone with page-locked host planes (mass planes: the crop only is "
                       "uploaded), every plane uploaded inside the timed region, y map read back"},
        "resident_vs_host_map_rel_l2": res_err,
        "map_check": {"what": "y map painted with eps = 0 (independent of how planes are dealt to ranks): compare "
                              "across n_gpus", "sum": float(y_fixed.sum()), "l2": float(np.sqrt((y_fixed ** 2).sum())),
                      "max": float(y_fixed.max())},
        "config": {"workload": "full create_lightcone line of sight (BASELINE.json configs[4]): %d synthetic slices, "
                               "plane pixels x%.2f of SLICS (12288^2 mass / 7745^2 delta), %d tiles of 512^2, quintic "
                               "projection to 1549^2" % (n_z, scale, n_tiles),
                   "parallelism": "whole planes dealt to %d rank(s) by cost (tiles per rank: %s), one NCCL reduce of "
                                  "the 19 MB map" % (world, loads),
                   "timed": "host wall clock between barriers + device synchronisation, max over ranks, median of "
                            "%d lines of sight; `value`: planes resident in each rank's HBM, `e2e`: plane uploads "
                            "(page-locked) inside; the stage split is of the e2e run" % lc_steps},
        "stages_s_max_over_ranks": dict(zip(("wait_plane", "extract", "paint", "stitch", "project", "reduce"),
                                            [float(v) for v in st])),
        "gpu_launches": int(launches), "clocks": clocks}


def run_lightcone(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    line = measure_lightcone(args, rank, world, local, args.precision)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="paint", help="paint (BASELINE configs[1], default) | lightcone (configs[4])")
    ap.add_argument("--planes", type=int, default=15, help="lightcone: number of slices")
    ap.add_argument("--plane-scale", type=float, default=1.0, help="lightcone: plane pixel count relative to SLICS")
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--tiles", type=int, default=256)
    ap.add_argument("--precision", default=os.environ.get("BARYON_PAINTER_PRECISION", "fp16"))
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--cpu-tiles", type=int, default=256, help="tiles in the bounded cpu_baseline sample (about 10 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle comparison of the timed batch")
    ap.add_argument("--no-fp32", action="store_true", help="skip the secondary fp32-accurate measurement")
    ap.add_argument("--no-extra", action="store_true", help="skip the CGAN / variance-map / lightcone secondary fields")
    ap.add_argument("--no-lightcone", action="store_true", help="skip the lightcone secondary field")
    ap.add_argument("--profile-layers", action="store_true", help="print the per-layer timing table to stderr")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.config == "lightcone":
        if args.steps == 20:
            args.steps = 3
        return run_lightcone(args)

    import torch
    import torch.distributed as dist
    from baryon_painter_b200 import _lib, synthetic
    from baryon_painter_b200.painter import CVAEPainter

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    warmup = max(args.warmup, 3)
    n = args.tiles
    painter = CVAEPainter.synthetic(tile_size=TILE, seed=0, compute_device="cuda:%d" % local,
                                    precision=args.precision, max_batch=n)
    net = painter.model.net
    tiles_h = synthetic.synthetic_dm_tiles(min(n, 16), TILE, seed0=1000 * rank)
    tiles_h = np.concatenate([tiles_h] * ((n + len(tiles_h) - 1) // len(tiles_h)))[:n]
    # distinct per-tile content without 256 RNG passes: scale the 16 base fields
    tiles_h = np.ascontiguousarray(tiles_h * np.linspace(0.8, 1.25, n, dtype=np.float32)[:, None, None])
    eps_h = synthetic.synthetic_latents(n, (TILE // 32, TILE // 32), seed=1 + rank)
    zs = np.zeros(n)
    s_in, s_out, tp = painter._sigmas(zs, True, True)
    tparams = (s_in, s_out, zs.astype(np.float32), *tp)
    flags = _lib.BP_FLAG_TRANSFORM | _lib.BP_FLAG_INVERSE
    d_tiles = torch.from_numpy(tiles_h).cuda()
    d_eps = torch.from_numpy(eps_h).cuda()
    d_out = torch.empty((n, TILE, TILE), dtype=torch.float32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        net.cvae_paint_device(d_tiles.data_ptr(), d_eps.data_ptr(), _lib.BP_LATENT_EPS, 0, tparams, flags,
                              d_out.data_ptr(), n, stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    _lib.launch_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count(reset=True)
    clocks = sampler.stop()
    if not np.all(np.isfinite(d_out[:2].cpu().numpy())):
        raise RuntimeError("non-finite painted tiles")

    # ---- parity of exactly what was timed (outside the timed region): 8 tiles spread over this rank's 256-tile
    # chunk, painted by the timed step's last iteration, against the oracle on the same inputs
    parity = None
    if rank == 0 and not args.no_parity:
        from oracle.cvae_oracle import CVAEOracle
        from baryon_painter_b200 import arch, transforms
        torch.set_num_threads(os.cpu_count())
        A = arch.fiducial_cvae_architecture(TILE)
        orc = CVAEOracle(A, synthetic.synthetic_cvae_state_dict(A, seed=0))
        stats = transforms.fiducial_stats()
        tol = {"fp32": 1e-4}.get(args.precision, 1e-2)
        idx = sorted(set(int(v) for v in np.linspace(0, n - 1, min(n, 8))))
        errs = []
        for i in idx:
            ref = orc.paint(tiles_h[i], 0.0, stats, eps=eps_h[i:i + 1]).astype(np.float64)
            got = d_out[i].cpu().numpy().astype(np.float64)
            errs.append(float(np.sqrt(((got - ref) ** 2).sum() / (ref ** 2).sum())))
        parity = {"tiles": len(idx), "of": n, "max_rel_l2": max(errs), "tol": tol, "ok": bool(max(errs) <= tol),
                  "oracle": "oracle/cvae_oracle.py (bit-identical to the reference's torch-CPU paint in the build container)"}
        if not parity["ok"]:
            raise RuntimeError("parity of the benchmarked batch failed: %r" % (parity,))

    # ---- end to end through the public API: host buffers in and out (page-locked, as the contract's
    # "pinned host memory"), every step copies its inputs H2D and its painted tiles D2H inside the timed region
    import baryon_painter_b200 as bp
    tiles_p = [bp.pinned_empty(tiles_h.shape) for _ in range(3)]
    out_h = [bp.pinned_empty(tiles_h.shape) for _ in range(3)]
    for b in tiles_p:
        b[...] = tiles_h
    # (a) one synchronous call per step: the batch is cut into pipeline chunks inside the call
    for _ in range(2):
        painter.paint_batch(tiles_p[0], z=0.0, eps=eps_h, out=out_h[0])
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        painter.paint_batch(tiles_p[0], z=0.0, eps=eps_h, out=out_h[0])
    torch.cuda.synchronize()
    e2e_sync_s = time.perf_counter() - t0
    # (b) the streaming call (paint_batch_async): every step still copies its 256 input tiles host -> device and its 256
    # painted tiles device -> host inside the timed region; the copies of neighbouring steps overlap this step's kernels
    def stream_steps(k):
        pending = []
        for i in range(k):
            pending.append(painter.paint_batch_async(tiles_p[i % 3], z=0.0, eps=eps_h, out=out_h[i % 3]))
            if len(pending) > 2:
                pending.pop(0).wait()
        for tk in pending:
            tk.wait()
    stream_steps(3)
    barrier()
    t0 = time.perf_counter()
    stream_steps(args.steps)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if not all(np.all(np.isfinite(o[-1])) for o in out_h):
        raise RuntimeError("non-finite painted tiles (host path)")
    if not (np.array_equal(out_h[0], out_h[1]) and np.array_equal(out_h[0], out_h[2])):
        raise RuntimeError("the I/O slots painted different tiles from the same inputs")

    t = torch.tensor([ms, e2e_s * 1e3, e2e_sync_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms, e2e_sync_ms = float(t[0]), float(t[1]), float(t[2])
    total_tiles = n * args.steps * world

    # ---- per-layer attribution (separate pass; event pairs around every layer launch)
    pk = peaks()
    roof = None
    if rank == 0:
        net.set_profile(True)
        step()
        torch.cuda.synchronize()
        rows = []
        for sidx, name in enumerate(("prior_network", "p_z_in", "p_y_z_in", "p_mu_out")):
            for li in range(len(painter.model.stacks[name])):
                info = net.layer_info(sidx, li)
                t_ms, cnt = net.read_profile(sidx, li)
                rows.append((name, li, info, t_ms, cnt))
        net.set_profile(False)
        tot = sum(r[3] for r in rows) or 1.0
        dom = [r for r in rows if r[2]["kernel"] == 3 and r[2]["cin"] == 128 and r[2]["cout"] == 128]
        dom_ms = sum(r[3] for r in dom)
        dom_fl = sum(r[2]["flops"] for r in dom) * n
        dom_launches = sum(r[4] for r in dom) or 1
        ach = dom_fl / (dom_ms * 1e-3) / 1e12 if dom_ms else 0.0
        whole = total_tiles / world * FLOPS_PER_TILE / (ms * 1e-3) / 1e12
        traffic = traffic_step = None
        try:   # DRAM bytes per launch of this kernel from the committed ncu capture (profiles/)
            with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
                tj = json.load(f)
            traffic = tj["dram_bytes_per_launch"] if tj.get("tiles_per_launch") == painter.model.net.chunk else None
            traffic_step = tj.get("whole_step_wconv_dram_bytes")
        except Exception:
            pass
        roof = {"bound": "tensor", "kernel": "wconv_kernel: 3x3 128->128 @64x64 residual-block convolution (8 layers)",
                "achieved": ach, "peak": pk["bf16_tflops"], "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops"],
                "peak_burst": pk["bf16_tflops_burst"], "frac_burst": ach / pk["bf16_tflops_burst"],
                "note": "the kernel is timed in one isolated profiled step (clocks near burst): frac_burst is the "
                        "honest fraction for it; whole_net is timed over the long sustained loop: frac (sustained peak)",
                "traffic": traffic, "traffic_whole_step": traffic_step,
                "traffic_note": "DRAM bytes (ncu dram__bytes_read + write) per launch of this kernel, and of all window-GEMM "
                                "launches of one 256-tile step (24.5 GB = 3.7 ms at the measured copy bandwidth): the 512^2 / "
                                "256^2 layers are HBM-bound, profiles/r02_summary.md",
                "peak_source": pk["source"], "launch_ms": dom_ms / dom_launches,
                "share_of_step": dom_ms / tot,
                "whole_net": {"achieved": whole, "frac": whole / pk["bf16_tflops"],
                              "frac_burst": whole / pk["bf16_tflops_burst"], "flops_per_tile": FLOPS_PER_TILE}}
        if args.profile_layers:
            for name, li, info, t_ms, cnt in rows:
                tf = info["flops"] * n / (t_ms * 1e-3) / 1e12 if t_ms else 0
                print("%-14s %2d  k%d s%d %3d->%3d @%3dx%-3d tensor=%d  %8.3f ms  %5.1f%%  %7.1f TFLOP/s" % (
                    name, li, info["kernel"], info["stride"], info["cin"], info["cout"], info["H"], info["W"],
                    info["tensor"], t_ms, 100 * t_ms / tot, tf), file=sys.stderr)

    # ---- second driver-visible number: the fp32-accurate tensor-core path (BASELINE configs[1]: "bf16 and fp32") on
    # the same 256 tiles, device-resident, with its own parity check at north_star's fp32 tolerance
    fp32 = None
    if rank == 0 and world == 1 and not args.no_fp32 and args.precision != "fp32":
        p32 = CVAEPainter.synthetic(tile_size=TILE, seed=0, compute_device="cuda:%d" % local, precision="fp32", max_batch=n)
        net32 = p32.model.net
        d_out32 = torch.empty_like(d_out)

        def step32():
            net32.cvae_paint_device(d_tiles.data_ptr(), d_eps.data_ptr(), _lib.BP_LATENT_EPS, 0, tparams, flags,
                                    d_out32.data_ptr(), n, stream)
        for _ in range(3):
            step32()
        torch.cuda.synchronize()
        k32 = max(3, min(args.steps, 5))
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(k32):
            step32()
        f1.record()
        torch.cuda.synchronize()
        ms32 = f0.elapsed_time(f1) / k32
        errs32 = []
        if not args.no_parity:
            for i in sorted(set(int(v) for v in np.linspace(0, n - 1, min(n, 4)))):
                ref = orc.paint(tiles_h[i], 0.0, stats, eps=eps_h[i:i + 1]).astype(np.float64)
                got = d_out32[i].cpu().numpy().astype(np.float64)
                errs32.append(float(np.sqrt(((got - ref) ** 2).sum() / (ref ** 2).sum())))
        fp32 = {"value": n / (ms32 * 1e-3), "unit": "tiles/s", "ms_per_step": ms32, "steps": k32,
                "dtype": "f32-accurate: fp16 (hi, lo) split operands on tcgen05, fp32 accumulation",
                "frac_of_bf16_roofline": n / (ms32 * 1e-3) * FLOPS_PER_TILE / 1e12 / peaks()["bf16_tflops"],
                "parity": {"tiles": len(errs32), "max_rel_l2": max(errs32) if errs32 else None, "tol": 1e-4,
                           "ok": bool(errs32 and max(errs32) <= 1e-4)} if errs32 else None}
        if errs32 and max(errs32) > 1e-4:
            raise RuntimeError("fp32 parity of the benchmarked batch failed: %r" % (fp32,))
        del p32, net32

    # ---- the other BASELINE.json configurations as secondary, driver-visible fields (N = 1 only)
    extra = {}
    if rank == 0 and world == 1 and not args.no_extra:
        from baryon_painter_b200.painter import CGANPainter
        # configs[2]: CGAN generator, 256 tiles, z cycling over {0, 0.5, 1} (100.7 GFLOP per tile)
        nc = min(n, 256)
        g = CGANPainter.synthetic(tile_size=TILE, seed=0, device="cuda:%d" % local, precision=args.precision, max_batch=nc)
        zc = np.array([0.0, 0.5, 1.0])[np.arange(nc) % 3]
        uz, inv = np.unique(zc, return_inverse=True)
        pin = [g.transform.gpu_params("dm", float(v)) for v in uz]
        pout = [g.inverse_transform.gpu_params("pressure", float(v)) for v in uz]
        tpc = (np.array([q[1] for q in pin], np.float32)[inv], np.array([q[1] for q in pout], np.float32)[inv],
               (zc - g.z_shift).astype(np.float32), pin[0][2], pin[0][3], pout[0][2], pout[0][3])
        d_outc = torch.empty((nc, TILE, TILE), dtype=torch.float32, device="cuda")

        def stepc():
            g.model.net.cgan_paint_device(d_tiles.data_ptr(), tpc, flags, d_outc.data_ptr(), nc, stream)
        for _ in range(3):
            stepc()
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kc = 5
        c0.record()
        for _ in range(kc):
            stepc()
        c1.record()
        torch.cuda.synchronize()
        msc = c0.elapsed_time(c1) / kc
        par = None
        if not args.no_parity:
            from oracle.cvae_oracle import CGANOracle
            from baryon_painter_b200 import arch as _arch
            layers = _arch.fiducial_cgan_architecture()
            corc = CGANOracle(layers, synthetic.synthetic_cgan_state_dict(layers, seed=0))
            errs = []
            for i in (0, nc - 1):
                ref = corc.paint(tiles_h[i], float(zc[i]), stats).astype(np.float64)
                got = d_outc[i].cpu().numpy().astype(np.float64)
                errs.append(float(np.sqrt(((got - ref) ** 2).sum() / (ref ** 2).sum())))
            par = {"tiles": 2, "max_rel_l2": max(errs), "tol": 1e-2, "ok": bool(max(errs) <= 1e-2),
                   "oracle": "restatement of trained_models/README.md + g_struc.pickle (parity unpinned: no PainterGAN source)"}
        extra["cgan"] = {"value": nc / (msc * 1e-3), "unit": "tiles/s", "ms_per_step": msc, "tiles": nc, "z": [0.0, 0.5, 1.0],
                         "flops_per_tile": 100.7e9, "frac_of_bf16_roofline": nc / (msc * 1e-3) * 100.7e9 / 1e12 / pk["bf16_tflops"],
                         "parity": par}
        del g, d_outc
        # configs[3]: 64 latent draws per tile -> per-pixel mean / variance maps (host buffers)
        nv, nd = 16, 64
        painter.paint_variance(tiles_h[:nv], z=0.0, n_draws=4, seed=1)
        t0 = time.perf_counter()
        mean_v, var_v = painter.paint_variance(tiles_h[:nv], z=0.0, n_draws=nd, seed=1)
        dtv = time.perf_counter() - t0
        extra["variance"] = {"value": nv * nd / dtv, "unit": "draws/s", "tiles": nv, "draws_per_tile": nd,
                             "finite": bool(np.isfinite(mean_v).all() and (var_v >= 0).all()),
                             "note": "host tiles in, mean / variance maps out; parity vs oracle draws: tests/test_gpu_cvae.py"}
        # configs[4]: one full synthetic line of sight on this GPU
        if not args.no_lightcone:
            del painter
            lc = measure_lightcone(args, 0, 1, local, args.precision)
            extra["lightcone"] = {k: lc[k] for k in LIGHTCONE_FIELDS}
    if world > 1 and not args.no_extra:
        # configs[3] at N GPUs: every rank resamples its own 16 tiles 64 times (weak scaling, no exchange); time = max over ranks
        nv, nd = 16, 64
        painter.paint_variance(tiles_h[:nv], z=0.0, n_draws=4, seed=1)
        dist.barrier()
        t0 = time.perf_counter()
        mean_v, var_v = painter.paint_variance(tiles_h[:nv], z=0.0, n_draws=nd, seed=1 + rank)
        tv = torch.tensor([time.perf_counter() - t0, 0.0 if (np.isfinite(mean_v).all() and (var_v >= 0).all()) else 1.0],
                          dtype=torch.float64, device="cuda")
        dist.all_reduce(tv, op=dist.ReduceOp.MAX)
        if rank == 0:
            extra["variance"] = {"value": world * nv * nd / float(tv[0]), "unit": "draws/s", "n_gpus": world, "tiles": world * nv,
                                 "draws_per_tile": nd, "finite": bool(float(tv[1]) == 0.0),
                                 "note": "host tiles in, mean / variance maps out, every rank its own tiles; parity vs oracle "
                                         "draws: tests/test_gpu_cvae.py"}
    if world > 1 and not args.no_extra and not args.no_lightcone:
        # configs[4] at N GPUs: ONE line of sight sharded over all ranks (strong scaling; the N = 1 line carries the
        # single-GPU time, map_check must agree between them)
        del painter
        lc = measure_lightcone(args, rank, world, local, args.precision)
        if rank == 0:
            extra["lightcone"] = {k: lc[k] for k in LIGHTCONE_FIELDS}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, dt = cpu_baseline(args.cpu_tiles, os.cpu_count())
        cpu = {"value": v, "unit": "tiles/s", "cores": os.cpu_count(), "kind": "port",
               "sample": "%d of the %d tiles, batch-1 paint() loop as the reference runs it (%.1f s); oracle port of "
                         "the reference torch-CPU path" % (args.cpu_tiles, n, dt)}
    if rank == 0:
        line = {
            "metric": "tiles/sec painted (fiducial CVAE)", "value": total_tiles / (ms * 1e-3), "unit": "tiles/s",
            "n_gpus": world, "steps": args.steps, "warmup": warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"bf16": "bf16", "fp16": "f16", "fp32": "f32 (fp16 hi/lo split operands)", "fp32-ffma": "f32"}.get(args.precision, args.precision),
            "data": "synthetic",
            "config": workload_config(n, world),
            "e2e": {"value": total_tiles / (e2e_ms * 1e-3), "unit": "tiles/s",
                    "h2d_bytes_per_step": int(tiles_h.nbytes + eps_h.nbytes), "d2h_bytes_per_step": int(out_h[0].nbytes),
                    "api": "CVAEPainter.paint_batch_async (stream of batches, three I/O slots, two batches outstanding; page-locked fp32 host tiles in "
                           "and out every step)",
                    "sync_call_value": total_tiles / (e2e_sync_ms * 1e-3),
                    "sync_call_api": "CVAEPainter.paint_batch (one blocking call per step)"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "parity": parity, "fp32": fp32, **extra}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
