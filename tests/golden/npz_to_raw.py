#!/usr/bin/env python
"""Expand tests/golden/case_*.npz into raw little-endian binaries a Julia (or C) program can read without any package:

    python tests/golden/npz_to_raw.py [outdir]          # default: tests/golden/raw/

For every case `<outdir>/<case>/` holds one file per array, `<name>.f64` / `<name>.i64` / `<name>.u64` (column-major,
Julia's own order), and `meta.txt`: `nx ny nz topology fill upwind rho_scalar use_rho3d`.  These are the twins
`tests/golden/pin_with_julia.jl` diffs the real OceanTransportMatrixBuilder.jl against."""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent


def main(out):
    out.mkdir(parents=True, exist_ok=True)
    for npz in sorted(HERE.glob("case_*.npz")):
        d = np.load(npz)
        case = out / npz.stem
        case.mkdir(exist_ok=True)
        nx, ny, nz = d["volcello"].shape
        (case / "meta.txt").write_text(f"{nx} {ny} {nz} {str(d['topology'])} {float(d['fill'])!r} {int(d['upwind'])} "
                                       f"{float(d['rho_scalar'])!r} {int(d['use_rho3d'])}\n")
        for k in d.files:
            a = d[k]
            if a.ndim == 0:
                continue
            ext = {"f": "f64", "i": "i64", "u": "u64"}[a.dtype.kind]
            a.astype({"f64": "<f8", "i64": "<i8", "u64": "<u8"}[ext]).ravel(order="F").tofile(case / f"{k}.{ext}")
        print(f"{case}: {len(d.files)} arrays")


if __name__ == "__main__":
    main(Path(sys.argv[1]) if len(sys.argv) > 1 else HERE / "raw")
