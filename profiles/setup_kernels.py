"""Runs the set-up pipeline (makeindices, makegridmetrics, facefluxes) a few times on a named grid: the program an ncu
launch list is taken of (`ncu --metrics gpu__time_duration.sum`), to read the set-up kernels' own durations."""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import otmb_b200.api as A
from otmb_b200 import synthetic
from _util import fields

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
oc = synthetic.make_config(cfg, seed=0)
f = fields(oc)
ctx = A.Context(0)
for rep in range(3):
    for plain in (False, True):
        os.environ.pop("OTMB_FACEFLUX_PLAIN", None)          # read by measurement builds only (-DOTMB_AB):
        if plain:                                            # the plain kernel instead of the bulk-copy tile kernel
            os.environ["OTMB_FACEFLUX_PLAIN"] = "1"
        gm = A.makegridmetrics(areacello=f["areacello"], volcello=f["volcello"], lon=f["lon"], lat=f["lat"], lev=f["lev"],
                               lon_vertices=f["lon_vertices"], lat_vertices=f["lat_vertices"], ctx=ctx)
        A.facefluxesfrommasstransport(umo=f["umo"], vmo=f["vmo"], gridmetrics=gm, indices=None, ctx=ctx)
print("done", cfg, ctx.resident["N"])
