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


def check(full, gm, oc, ex_rank, tag, info):
    phi = O.facefluxes(oc.umo, oc.vmo, gm.v3D, gm.gridtopology.kind, oc.fill)
    stack = lambda d: np.asfortranarray(np.stack([d[k] for k in O.DIRS], axis=-1))
    want = O.transportmatrix(phi, oc.mlotst, gm.v3D, gm.thkcello, gm.area2D, gm.zt, stack(gm.edge_length_2D),
                             stack(gm.distance_to_neighbour_2D), gm.gridtopology.kind, 1035.0)
    for name, oname in NAMES.items():
        g, w = getattr(full, name), want[oname]
        assert np.array_equal(g.indptr + 1, w.colptr) and np.array_equal(g.indices + 1, w.rowval), name
        assert np.array_equal(g.data.view(np.int64), w.nzval.view(np.int64)), name
    print(f"SHARDED-{tag}-OK {info}")


def main_native(local_rank):
    """OTMB_SHARDED_BACKEND=native (needs one GPU per rank): the in-library driver — NCCL communicator created from
    a broadcast id, chunk-pipelined carry chain, collective build — against the oracle, bit for bit."""
    import torch
    import otmb_b200.api as A
    torch.cuda.set_device(local_rank)
    dist.init_process_group("gloo")                      # host-side rendezvous only: id broadcast + result gather
    rank, size = dist.get_rank(), dist.get_world_size()
    for shape, chunks in (((90, 45, 20), 0), ((37, 11, 7), 3), ((64, 33, 9), 64)):
        box = [sharded.NativeSharded.unique_id() if rank == 0 else None]      # an id makes ONE communicator
        dist.broadcast_object_list(box, src=0)
        oc = synthetic.make_ocean(*shape, "tripolar", seed=3, land_frac=0.25)
        gm = oracle_gridmetrics(oc)
        ns = sharded.NativeSharded(gridmetrics=gm, rank=rank, nranks=size, id_bytes=box[0], device=local_rank)
        ns.set_masstransport(oc.umo, oc.vmo, oc.fill)
        ns.facefluxes(chunks)
        ns.build(oc.mlotst, 1035.0)
        segs = ns.segments()
        ex = sharded.TorchExchange(None)
        nnz_all = ex.allgather_ints(ns.slab.nnz)
        full = sharded._gather_full(ex, segs, nnz_all, ns.N)
        if rank == 0:
            check(full, gm, oc, rank, "NATIVE", f"ranks={size} shape={shape} chunks={chunks} slabs={ns.slabs}")
        # an error on one rank (NaN density in the deepest slab only) is returned on every rank
        rho = np.array(oc.rho3d, order="F")
        rho[:, :, -1] = np.where(np.isnan(rho[:, :, -1]), rho[:, :, -1], np.nan)
        deepest_wet = np.isfinite(oc.rho3d[:, :, -1]).any()
        try:
            ns.build(oc.mlotst, rho)
            assert not deepest_wet, "expected an error on every rank"
        except A.OTMBError as e:
            assert e.code == 5, (rank, e.code, str(e))     # the reference's ρ check comes first, whichever rank saw it
        ns.close()
        dist.barrier()
    dist.destroy_process_group()


def main():
    backend = os.environ.get("OTMB_SHARDED_BACKEND", "gloo")     # "nccl": the CUDA slab path, one GPU per rank
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if backend == "native":
        return main_native(local_rank)
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
