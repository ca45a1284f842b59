"""BASELINE configs[4]: twelve monthly-climatology matrices on one grid (C2 shape), on ONE GPU here
(`bench.py --gpus N` deals one matrix per GPU for the multi-GPU version).  Geometry and indices are
made once; per month only umo/vmo and mlotst change.  Prints per-month times of (a) everything rebuilt
on the device and (b) the reference's caching hook (TκH and TκVdeep passed back as pre-built operators,
/root/reference/src/matrixbuilding.jl:133-143)."""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import otmb_b200.api as A
from otmb_b200 import synthetic
from _util import fields

months = int(sys.argv[1]) if len(sys.argv) > 1 else 12
ctx = A.Context(0)
oc0 = synthetic.make_config("C2", seed=0)
f0 = fields(oc0)
gm = A.makegridmetrics(areacello=f0["areacello"], volcello=f0["volcello"], lon=f0["lon"], lat=f0["lat"], lev=f0["lev"],
                       lon_vertices=f0["lon_vertices"], lat_vertices=f0["lat_vertices"], ctx=ctx)
ix = A.makeindices(gm.v3D, ctx=ctx)
data = []
for m in range(months):
    oc = synthetic.make_config("C2", seed=m)          # same grid and mask, new fluxes and mixed layer
    data.append((fields(oc), oc.mlotst))
keep = first = None
for label, reuse in (("rebuild all four operators", False), ("TκH, TκVdeep passed back pre-built (pageable copies)", True),
                     ("TκH, TκVdeep passed back pre-built (the page-locked arrays the first call returned)", 2)):
    t_ff, t_tm, t_dev = [], [], []
    for m, (f, ml) in enumerate(data):
        phi = tm = None                                   # drop last month's results: their pinned buffers are reused
        t0 = time.perf_counter()
        phi = A.facefluxesfrommasstransport(umo=f["umo"], vmo=f["vmo"], gridmetrics=gm, indices=ix, ctx=ctx)
        t1 = time.perf_counter()
        src = first if reuse == 2 else keep
        kw = dict(TκH=src.TκH, TκVdeep=src.TκVdeep) if (reuse and src is not None) else {}
        tm = A.transportmatrix(ϕ=phi, mlotst=ml, gridmetrics=gm, indices=ix, ρ=1035.0, ctx=ctx, **kw)
        t2 = time.perf_counter()
        t_ff.append(t1 - t0); t_tm.append(t2 - t1); t_dev.append(ctx.last_build_ms())
        if keep is None:
            keep, first = A.TransportMatrices(*[m.copy() for m in tm]), tm
    print(f"{label}: per month facefluxes {1e3 * np.median(t_ff):.1f} ms (host arrays in/out), transportmatrix "
          f"{1e3 * np.median(t_tm):.1f} ms end to end of which device assembly {np.median(t_dev):.3f} ms; nnz(T) {tm.T.nnz}; "
          f"transportmatrix month by month {[round(1e3 * t, 1) for t in t_tm]}")
