"""Trace of one otmb_transportmatrix_stream call (measurement build: OTMB_NVCC_EXTRA=-DOTMB_AB, OTMB_STREAM_TRACE=1)."""
import ctypes as C
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import otmb_b200.api as A
from otmb_b200 import _lib, synthetic
from _util import fields

oc = synthetic.make_config("C2", seed=0)
f = fields(oc)
ctx = A.Context(0)
lib = ctx.lib
gm = A.makegridmetrics(areacello=f["areacello"], volcello=f["volcello"], lon=f["lon"], lat=f["lat"], lev=f["lev"],
                       lon_vertices=f["lon_vertices"], lat_vertices=f["lat_vertices"], ctx=ctx)
N = ctx.resident["N"]
phi = A.facefluxesfrommasstransport(umo=f["umo"], vmo=f["vmo"], gridmetrics=gm, indices=None, ctx=ctx)
hin = [ctx.pinned_empty(gm.v3D.size, np.float64, gm.v3D.shape, "F") for _ in range(6)]
for a, k in zip(hin, A.FACES):
    a[...] = getattr(phi, k)
ml = ctx.pinned_empty(oc.mlotst.size, np.float64, oc.mlotst.shape, "F")
ml[...] = oc.mlotst
caps = [N * w for w in (7, 7, 5, 3, 3)]
outs = [(ctx.pinned_empty(N + 1, np.int64), ctx.pinned_empty(caps[m], np.int64), ctx.pinned_empty(caps[m], np.float64)) for m in range(5)]
ptrs = (C.c_void_p * 6)(*[a.ctypes.data for a in hin])
out_ptrs = [(C.c_void_p * 5)(*[outs[m][q].ctypes.data for m in range(5)]) for q in range(3)]
prm = _lib.TMParams(500.0, 0.1, 1.0e-5, 1035.0, 1, 0, 0, 0)
nnz = (C.c_int64 * 5)()
for nslabs in [int(x) for x in (sys.argv[1:] or ["8"])]:
    os.environ.pop("OTMB_STREAM_TRACE", None)
    ts = []
    for it in range(8):
        if it == 7 and os.environ.get("TRACE"):
            os.environ["OTMB_STREAM_TRACE"] = "1"
        t = time.perf_counter()
        ctx.check(lib.otmb_transportmatrix_stream(ctx.h, C.byref(prm), ptrs, A._ptr(ml), None, nslabs, (C.c_int64 * 5)(*caps), *out_ptrs, nnz))
        ts.append(1e3 * (time.perf_counter() - t))
    print(f"nslabs={nslabs}: ms per call {['%.2f' % x for x in ts]}", flush=True)
