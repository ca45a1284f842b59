"""Small driver for ncu: set up the C2 case once and run a few device-resident transportmatrix builds."""
import ctypes as C
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import otmb_b200.api as A
from otmb_b200 import _lib, synthetic
from _util import fields

path = sys.argv[1] if len(sys.argv) > 1 else "fused"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = sys.argv[3] if len(sys.argv) > 3 else "C2"
ctx = A.Context(0)
oc = synthetic.make_config(cfg, seed=0)
f = fields(oc)
gm = A.makegridmetrics(areacello=f["areacello"], volcello=f["volcello"], lon=f["lon"], lat=f["lat"], lev=f["lev"],
                       lon_vertices=f["lon_vertices"], lat_vertices=f["lat_vertices"], ctx=ctx)
A.facefluxesfrommasstransport(umo=f["umo"], vmo=f["vmo"], gridmetrics=gm, indices=None, ctx=ctx)
ctx.check(ctx.lib.otmb_set_mlotst(ctx.h, A._ptr(A._f64(oc.mlotst))))
ctx.check(ctx.lib.otmb_set_rho3d(ctx.h, None))
prm = _lib.TMParams(500.0, 0.1, 1.0e-5, 1035.0, 1, 0, _lib.PATH[path], 0)
nnz = (C.c_int64 * 5)()
for _ in range(steps):
    ctx.check(ctx.lib.otmb_transportmatrix_build(ctx.h, C.byref(prm), nnz))
    print("build ms", ctx.last_build_ms(), list(nnz))
