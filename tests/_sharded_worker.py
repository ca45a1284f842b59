"""Worker of the world_size-2 gloo test (launched with torch.distributed.run by test_sharded.py):
the sharded driver over a real process group, compute by the oracle test double."""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import numpy as np
import torch.distributed as dist

import otmb_b200  # noqa: F401
from otmb_b200 import sharded, synthetic
from oracle import oracle as O
from _sharded_double import NAMES, OracleSlab, oracle_gridmetrics


def main():
    backend = os.environ.get("OTMB_SHARDED_BACKEND", "gloo")     # "nccl": the CUDA slab path, one GPU per rank
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if backend == "nccl":
        import torch
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        factory, shape = sharded.CudaSlab, (90, 45, 20)
    else:
        dist.init_process_group("gloo")
        factory, shape = OracleSlab, (14, 10, 7)
    ex = sharded.TorchExchange(local_rank if backend == "nccl" else None)
    oc = synthetic.make_ocean(*shape, "tripolar", seed=3, land_frac=0.25)
    gm = oracle_gridmetrics(oc)
    full, segs, info = sharded.transportmatrix_sharded(exchange=ex, gridmetrics=gm, mlotst=oc.mlotst, ρ=1035.0, umo=oc.umo,
                                                       vmo=oc.vmo, FillValue=oc.fill, slab_factory=factory, device=local_rank)
    assert sum(info["counts"]) == info["N"] and len(info["slabs"]) == ex.size
    assert segs["T"].col0 == sum(info["counts"][:ex.rank])
    if ex.rank == 0:
        phi = O.facefluxes(oc.umo, oc.vmo, gm.v3D, gm.gridtopology.kind, oc.fill)
        stack = lambda d: np.asfortranarray(np.stack([d[k] for k in O.DIRS], axis=-1))
        want = O.transportmatrix(phi, oc.mlotst, gm.v3D, gm.thkcello, gm.area2D, gm.zt, stack(gm.edge_length_2D),
                                 stack(gm.distance_to_neighbour_2D), gm.gridtopology.kind, 1035.0)
        for name, oname in NAMES.items():
            g, w = getattr(full, name), want[oname]
            assert np.array_equal(g.indptr + 1, w.colptr) and np.array_equal(g.indices + 1, w.rowval), name
            assert np.array_equal(g.data.view(np.int64), w.nzval.view(np.int64)), name
        print(f"SHARDED-{backend.upper()}-OK ranks={ex.size} slabs={info['slabs']} counts={info['counts']}")
    else:
        assert full is None
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
