"""Produce ``tuning_table.txt``: time every tensor-core formulation / tiling of every layer on THIS GPU.

    python -m baryon_painter_b200.tune [--out baryon_painter_b200/tuning_table.txt] [--log]

Run on a B200.  Net creation normally looks each layer up in the shipped table (deterministic); here
``bp_tuning_mode(1)`` makes it time the candidates on a full chunk instead and record the winners for the
production shapes: the fiducial CVAE (512^2, 256-tile chunk) and the CGAN generator; fp16, bf16 and the
split-precision fp32 path.  The
result is data to commit, so that every process afterwards paints bit-identical tiles.
"""
import argparse
import sys

from . import _lib


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=_lib.TUNING_TABLE)
    ap.add_argument("--log", action="store_true")
    ap.add_argument("--formats", default="fp16,bf16,fp32")
    ap.add_argument("--no-cgan", action="store_true")
    args = ap.parse_args()
    from .painter import CGANPainter, CVAEPainter
    _lib.load()
    _lib.tuning_set("")
    _lib.tuning_mode(True, args.log)
    for prec in args.formats.split(","):
        p = CVAEPainter.synthetic(tile_size=512, seed=0, precision=prec, max_batch=256)
        del p
        if not args.no_cgan:
            g = CGANPainter.synthetic(tile_size=512, seed=0, precision=prec, max_batch=64)
            del g
    _lib.tuning_mode(False)
    text = ("# formulation table written by `python -m baryon_painter_b200.tune` on a B200 (see include/baryon_painter_b200.h)\n"
            "# key G Jy N mode Wt T_r gl nbst ms\n" + _lib.tuning_get())
    with open(args.out, "w") as f:
        f.write(text)
    sys.stdout.write(text)


if __name__ == "__main__":
    main()
