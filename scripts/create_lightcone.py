#!/usr/bin/env python
"""Lightcone driver -- drop-in for reference scripts/create_lightcone.py: same flags, same defaults (all values
parsed as strings and cast later), same outputs (``<output-file>.npy``,
``<output-file>_drop_<n>.npy``, optional pickle of the painted planes), painting on B200s.

Additive flags: ``--device``, ``--precision``, ``--batch``, ``--synthetic`` (seeded planes and a flat-LCDM
background instead of the SLICS files and pyccl: runs on a box with neither; ``--SLICS-base-path`` /
``--SLICS-LOS`` may then be omitted), ``--synthetic-plane-pixels``.  Under ``torchrun --nproc-per-node N`` the
(plane, tile) work items are sharded over the N GPUs of the box and rank 0 assembles and writes the map.

    python scripts/create_lightcone.py --synthetic --n-plane 4 --output-file /tmp/y_map
    python scripts/create_lightcone.py --CVAE-path trained_models/CVAE/fiducial --SLICS-base-path ... --SLICS-LOS 500 \
        --output-file y_map_LOS500
"""
import argparse
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pi = np.pi

if __name__ == "__main__":
    parser = argparse.ArgumentParser()
    parser.add_argument("--model-type", default="CVAE")
    parser.add_argument("--CVAE-path")

    parser.add_argument("--CGAN-module-path")          # accepted for compatibility; the generator runs natively
    parser.add_argument("--CGAN-parts-path")
    parser.add_argument("--CGAN-checkpoint")

    parser.add_argument("--SLICS-base-path")
    parser.add_argument("--SLICS-LOS")

    parser.add_argument("--n-plane", default=15)
    parser.add_argument("--tile-overlap", default=0.2)

    parser.add_argument("--output-resolution", default=7745 // 5)

    parser.add_argument("--drop-planes")
    parser.add_argument("--output-file", required=True)
    parser.add_argument("--output-file-planes")

    parser.add_argument("--device")
    parser.add_argument("--precision")
    parser.add_argument("--batch", default=64)
    parser.add_argument("--synthetic", action="store_true")
    parser.add_argument("--synthetic-plane-pixels", default=2048)

    args = parser.parse_args()
    if not args.synthetic and (args.SLICS_base_path is None or args.SLICS_LOS is None):
        parser.error("--SLICS-base-path and --SLICS-LOS are required (unless --synthetic)")

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    device = args.device or "cuda:%d" % int(os.environ.get("LOCAL_RANK", "0"))
    say = print if rank == 0 else (lambda *a, **k: None)

    import torch
    torch.cuda.set_device(torch.device(device))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(device))

    import baryon_painter_b200.painter
    import baryon_painter_b200.process_SLICS

    batch = int(args.batch)
    if args.model_type == "CVAE":
        say("Using CVAE.")
        if args.CVAE_path is not None:
            painter = baryon_painter_b200.painter.CVAEPainter((os.path.join(args.CVAE_path, "model_state"),
                                                               os.path.join(args.CVAE_path, "model_meta")),
                                                              compute_device=device, precision=args.precision,
                                                              max_batch=batch)
        else:
            say("No --CVAE-path: seeded synthetic weights of the fiducial architecture.")
            painter = baryon_painter_b200.painter.CVAEPainter.synthetic(compute_device=device, precision=args.precision,
                                                                         max_batch=batch)
    elif args.model_type == "CGAN":
        say("Using GAN")
        if args.CGAN_checkpoint is not None:
            painter = baryon_painter_b200.painter.CGANPainter(args.CGAN_parts_path, checkpoint_file=args.CGAN_checkpoint,
                                                              device=device, precision=args.precision, max_batch=batch)
        else:
            say("No --CGAN-checkpoint: seeded synthetic generator weights.")
            painter = baryon_painter_b200.painter.CGANPainter.synthetic(device=device, precision=args.precision,
                                                                         max_batch=batch)
    else:
        parser.error("Only CVAE and CGAN are supported for --model-type.")

    output_file = args.output_file
    say(f"Writing result to {output_file}.npy")
    n_drop = None
    if args.drop_planes is not None:
        n_drop = int(args.drop_planes)
        output_file_drop = output_file + f"_drop_{n_drop}"
        say(f"Writing result to {output_file_drop}.npy")

    n_z = int(args.n_plane)
    h = 0.6898
    plane_source = None
    if args.synthetic:
        cosmo_SLICS = baryon_painter_b200.process_SLICS.FlatLCDM(Omega_m=0.2905, h=h)
        LOS = int(args.SLICS_LOS) if args.SLICS_LOS is not None else 0
        chi = 252.5 / h * (np.arange(n_z) + 0.5)                                  # slab mid-planes, Mpc
        z_SLICS = 1 / cosmo_SLICS.scale_factor_of_chi(chi) - 1
        d_A_SLICS = chi * h
        z_slice = np.array([1 / float(cosmo_SLICS.scale_factor_of_chi(252.5 / h * i)) - 1 for i in range(n_z)])
        delta_path = massplane_path = None
        shifts_path = np.random.default_rng(0).random((n_z, 2))
        npx = int(args.synthetic_plane_pixels)

        def plane_source(i, kind):
            # smooth, strictly positive, periodic log-normal density (mean ~1): a 512^2 Gaussian-filtered field tiled
            # up to the plane size (a cubic-spline zoom of white noise would overshoot below zero and the shift-log
            # transform would see negative densities; filtering the full plane costs seconds per plane on the host)
            import scipy.ndimage
            nb = min(512, npx)
            g = scipy.ndimage.gaussian_filter(np.random.default_rng(100 + i).standard_normal((nb, nb)), 3.0, mode="wrap")
            base = np.exp(g / g.std() - 0.5).astype(np.float32)
            reps = (npx + nb - 1) // nb
            return np.ascontiguousarray(np.tile(base, (reps, reps))[:npx, :npx])
    else:
        SLICS_base_path = args.SLICS_base_path
        LOS = int(args.SLICS_LOS)
        say(f"Looking in {SLICS_base_path} for SLICS files.")
        say(f"Processing LOS{LOS}.")
        delta_path = os.path.join(SLICS_base_path, "delta")
        massplane_path = os.path.join(SLICS_base_path, "massplanes")
        shifts_path = os.path.join(SLICS_base_path, "random_shifts")
        delta_filenames = glob.glob(os.path.join(delta_path, f"*delta.dat_bicubic_LOS{LOS}"))
        if len(delta_filenames) == 0:
            raise RuntimeError(f"LOS {LOS} isn't complete.")
        z_SLICS = np.array(sorted(float(n[:n.find("delta")]) for n in (os.path.split(f)[1] for f in delta_filenames)))
        say("SLICS redshifts:", z_SLICS)
        try:
            import pyccl as ccl
            cosmo_SLICS = ccl.Cosmology(Omega_c=(1 - 0.7095 - 0.0473), Omega_b=0.0473, Omega_k=0, h=h, sigma8=0.826,
                                        n_s=0.969, m_nu=0.0)
            d_A_SLICS = ccl.comoving_angular_distance(cosmo_SLICS, 1 / (1 + z_SLICS)) * h
            z_slice = np.array([1 / ccl.scale_factor_of_chi(cosmo_SLICS, 252.5 / h * i) - 1 for i in range(len(z_SLICS))])
        except ImportError:
            cosmo_SLICS = baryon_painter_b200.process_SLICS.FlatLCDM(Omega_m=0.2905, h=h)
            d_A_SLICS = cosmo_SLICS.comoving_distance(1 / (1 + z_SLICS)) * h
            z_slice = np.array([1 / float(cosmo_SLICS.scale_factor_of_chi(252.5 / h * i)) - 1 for i in range(len(z_SLICS))])

    tile_overlap = float(args.tile_overlap)
    say(f"Painting {n_z} out of {len(z_SLICS)} planes.")
    say(f"Using an overlap of {tile_overlap}.")

    import time
    t_start = time.perf_counter()
    ps = baryon_painter_b200.process_SLICS
    common = dict(tile_size=100.0, n_pixel_tile=512, LOS=LOS, z_SLICS=z_SLICS[:n_z],
                  delta_size=d_A_SLICS[:n_z] * 10 / 180 * pi, delta_path=delta_path, massplane_path=massplane_path,
                  shifts_path=shifts_path, z_slice=z_slice[:n_z], verbose=rank == 0, plane_source=plane_source, rank=rank,
                  world_size=world, batch=batch)
    output_resolution = int(args.output_resolution)
    if args.output_file_planes is None:
        # the planes themselves are not wanted: every rank projects its own planes into its partial y map on its GPU
        # and ONE reduction of the map assembles the result (reference :106-128 as a single sharded pass)
        res = ps.paint_lightcone(painter, resolution=output_resolution, map_size=10.0, cosmo=cosmo_SLICS, order=5,
                                 drop_planes=n_drop, **common)
        say(f"Painted, stitched and projected {n_z} planes on {world} GPU(s) in {time.perf_counter() - t_start:.2f} s.")
        if rank == 0:
            if n_drop is None:
                np.save(output_file, res)
            else:
                np.save(output_file, res[0])
                np.save(output_file_drop, res[1])
    else:
        painted_planes = ps.process_SLICS(painter, min_tiling_overlap=tile_overlap, regularise=False, regularise_std=None,
                                          **common)
        t_painted = time.perf_counter()
        say(f"Painted and stitched {n_z} planes on {world} GPU(s) in {t_painted - t_start:.2f} s.")
        if rank == 0:
            be = ps.DeviceBackend(device)       # quintic zoom + sum on the GPU
            y_map = ps.create_y_map(painted_planes, z_SLICS[:n_z], resolution=output_resolution, map_size=10.0,
                                    cosmo=cosmo_SLICS, order=5, backend=be)
            np.save(output_file, y_map)
            say(f"Projected the y map ({output_resolution} px) in {time.perf_counter() - t_painted:.2f} s.")
            if n_drop is not None:
                y_map = ps.create_y_map(painted_planes[n_drop:], z_SLICS[n_drop:n_z], resolution=output_resolution,
                                        map_size=10.0, cosmo=cosmo_SLICS, order=5, backend=be)
                np.save(output_file_drop, y_map)
            import pickle
            with open(args.output_file_planes, "wb") as f:
                pickle.dump(painted_planes, f)
    if world > 1:
        dist.destroy_process_group()
