"""Kernel / step time of the device-resident C2 build for a list of environment settings (measurement builds).
    python profiles/kbench.py "OTMB_V4_LWET_AHEAD=0" "OTMB_V4_LWET_AHEAD=296" ...
Every setting: 5 warm-up builds, then 30 timed back to back (ms_per_step) and 30 with per-launch events (kernel_ms)."""
import ctypes as C
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import otmb_b200.api as A
from otmb_b200 import _lib, synthetic
from _util import fields

cfg = "C2"
ctx = A.Context(0)
lib = ctx.lib
oc = synthetic.make_config(cfg, seed=0)
f = fields(oc)
gm = A.makegridmetrics(areacello=f["areacello"], volcello=f["volcello"], lon=f["lon"], lat=f["lat"], lev=f["lev"],
                       lon_vertices=f["lon_vertices"], lat_vertices=f["lat_vertices"], ctx=ctx)
A.facefluxesfrommasstransport(umo=f["umo"], vmo=f["vmo"], gridmetrics=gm, indices=None, ctx=ctx)
ctx.check(lib.otmb_set_mlotst(ctx.h, A._ptr(A._f64(oc.mlotst))))
ctx.check(lib.otmb_set_rho3d(ctx.h, None))
prm = _lib.TMParams(500.0, 0.1, 1.0e-5, 1035.0, 1, 0, 0, 0)
nnz = (C.c_int64 * 5)()
step = lambda: ctx.check(lib.otmb_transportmatrix_build(ctx.h, C.byref(prm), nnz))
for rep in range(2):
    for setting in (sys.argv[1:] or ["BASE=1"]):
        k, v = setting.split("=")
        os.environ[k] = v
        for _ in range(5):
            step()
        ctx.check(lib.otmb_set_build_timing(ctx.h, 0))
        ctx.check(lib.otmb_synchronize(ctx.h))
        ctx.check(lib.otmb_timer_start(ctx.h))
        for _ in range(30):
            step()
        ms = C.c_float()
        ctx.check(lib.otmb_timer_stop(ctx.h, C.byref(ms)))
        ctx.check(lib.otmb_set_build_timing(ctx.h, 1))
        ks = []
        for _ in range(30):
            step()
            ks.append(ctx.last_build_ms())
        ks.sort()
        print(f"{setting:28s} step {ms.value / 30:.4f} ms   kernel median {ks[15]:.4f}  min {ks[0]:.4f}", flush=True)
        os.environ.pop(k, None)
