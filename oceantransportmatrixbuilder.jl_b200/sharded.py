"""k-slab sharding of ONE transport matrix across GPUs (BASELINE config 4).

The reference is single-process; what is sharded here is its loop over wet cells
(`for 𝑖 in eachindex(Lwet)`, /root/reference/src/matrixbuilding.jl:237, 348, 450).  Wet ranks
are ordered k-slowest (`Lwet = L[wet3D]`, :14-16), so a contiguous range of LEVELS is a
contiguous block of rows/columns of every matrix.  Rank r owns the levels [k0, k1) chosen so
that every rank holds about N/R wet cells, keeps one halo level on either side resident, and
assembles the CSC *columns* of its cells with global row indices; the complete matrix is the
concatenation of the ranks' (rowval, nzval) segments, with each local colptr shifted by the
number of entries of the lower ranks.

Exchanges between ranks (everything else is rank-local):
  1. all-gather of one integer per rank — owned wet cells -> global wet-rank offsets;
  2. the face-flux continuity scan (/root/reference/src/velocities.jl:234-243) runs bottom-up
     and is a floating-point recurrence, ϕtop[k] = ((((ϕtop[k+1] + w) + s) - e) - n): it cannot be
     re-associated without changing bits, so the slabs form a chain — the rank below hands its
     top plane (nx*ny doubles) to the rank above (NCCL send/recv between device buffers; this is
     the only halo that moves between GPUs).  Skipped when the caller passes ϕ itself;
  3. all-gather of five integers per rank — nnz per matrix -> colptr offsets;
  4. optionally a gather of the finished segments on rank 0 (tests / small cases; in production
     every rank copies its segment to its place in the host arrays).

`Exchange` is the plumbing (torch.distributed with NCCL or gloo, or threads inside one process
for the single-GPU emulation used by the GPU tests); `CudaSlab` is the compute (libotmb.so, one
context per rank).  Both are passed in, so the host logic can be exercised without a GPU by a
test double for the compute — the product default fails loudly without the CUDA library.
"""
from __future__ import annotations

import ctypes as C
import queue
import threading
from dataclasses import dataclass

import numpy as np
import scipy.sparse as sp

from . import _lib as _L
from . import api as A

MATS = A.MATRICES


# ------------------------------------------------------------------------------------------
# partition
# ------------------------------------------------------------------------------------------
def wet_per_level(v3D) -> np.ndarray:
    """Wet cells per level, wet <=> !isnan(v3D) (src/matrixbuilding.jl:14)."""
    v = np.asarray(v3D)
    return (~np.isnan(v)).sum(axis=(0, 1)).astype(np.int64)


def plan_slabs(wet_levels, nranks):
    """Contiguous level ranges [(k0, k1), ...], one per rank, every rank at least one level, cut
    where the cumulative wet count is closest to r*N/R.  Deterministic: every rank computes the
    same plan from the same counts."""
    wet_levels = np.asarray(wet_levels, dtype=np.int64)
    nz = len(wet_levels)
    if not 1 <= nranks <= nz:
        raise ValueError(f"need 1 <= ranks <= number of levels, got {nranks} ranks for {nz} levels")
    cum = np.concatenate([[0], np.cumsum(wet_levels)])
    total = cum[-1]
    cuts = [0]
    for r in range(1, nranks):
        lo = cuts[-1] + 1                      # at least one level for rank r-1
        hi = nz - (nranks - r)                 # and for every later rank
        target = total * r / nranks
        k = lo + int(np.argmin(np.abs(cum[lo:hi + 1] - target)))
        cuts.append(k)
    cuts.append(nz)
    return [(cuts[r], cuts[r + 1]) for r in range(nranks)]


# ------------------------------------------------------------------------------------------
# exchange plumbing
# ------------------------------------------------------------------------------------------
class TorchExchange:
    """torch.distributed process group (backend nccl: device buffers over NVLink; gloo: host buffers)."""

    def __init__(self, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank, self.size = dist.get_rank(), dist.get_world_size()
        self.on_device = dist.get_backend() == "nccl"
        self.device = torch.device("cuda", device if device is not None else torch.cuda.current_device()) \
            if self.on_device else torch.device("cpu")

    def allgather_ints(self, values):
        t = self.torch.tensor(list(values), dtype=self.torch.int64, device=self.device)
        out = [self.torch.empty_like(t) for _ in range(self.size)]
        self.dist.all_gather(out, t)
        return [[int(x) for x in o.tolist()] for o in out]

    def new_plane(self, n):
        """Buffer of n doubles the backend can send/receive: (buffer, pointer, is_device, numpy view or None)."""
        t = self.torch.zeros(n, dtype=self.torch.float64, device=self.device)
        return t, t.data_ptr(), self.on_device, (None if self.on_device else t.numpy())

    def send(self, dst, plane):
        if self.on_device:
            self.torch.cuda.synchronize()      # the library wrote the plane on its own stream
        self.dist.send(plane[0], dst)

    def recv(self, src, plane):
        self.dist.recv(plane[0], src)
        if self.on_device:
            self.torch.cuda.synchronize()

    def gather_arrays(self, arr, dst=0):
        """Variable-length gather of a 1-D numpy array to rank dst (tests / small cases)."""
        objs = [None] * self.size if self.rank == dst else None
        self.dist.gather_object(arr, objs, dst=dst)
        return objs


class ThreadExchange:
    """Ranks as threads of one process (one context per rank, possibly all on one GPU): the
    emulation the single-GPU tests use.  Host buffers."""

    class _Shared:
        def __init__(self, size):
            self.size = size
            self.barrier = threading.Barrier(size)
            self.slots = [None] * size
            self.queues = {(s, d): queue.Queue() for s in range(size) for d in range(size)}

    def __init__(self, shared, rank):
        self.shared, self.rank, self.size = shared, rank, shared.size
        self.on_device = False

    def allgather_ints(self, values):
        self.shared.slots[self.rank] = [int(v) for v in values]
        self.shared.barrier.wait()
        out = [list(s) for s in self.shared.slots]
        self.shared.barrier.wait()
        return out

    def new_plane(self, n):
        a = np.zeros(n, dtype=np.float64)
        return a, a.ctypes.data, False, a

    def send(self, dst, plane):
        self.shared.queues[(self.rank, dst)].put(plane[0].copy())

    def recv(self, src, plane):
        plane[0][...] = self.shared.queues[(src, self.rank)].get(timeout=120)

    def gather_arrays(self, arr, dst=0):
        self.shared.slots[self.rank] = arr
        self.shared.barrier.wait()
        out = list(self.shared.slots) if self.rank == dst else None
        self.shared.barrier.wait()
        return out


# ------------------------------------------------------------------------------------------
# compute on one slab (CUDA)
# ------------------------------------------------------------------------------------------
class CudaSlab:
    """One libotmb.so context restricted to the levels [k0, k1) (otmb_set_slab)."""

    def __init__(self, shape, topology, k0, k1, device=0, ctx=None):
        self.ctx = ctx or A.Context(device)
        self.lib = self.ctx.lib
        self.shape, self.k0, self.k1 = tuple(shape), k0, k1
        nx, ny, nz = self.shape
        self.ctx.check(self.lib.otmb_set_grid(self.ctx.h, nx, ny, nz, _L.TOPO[topology]))
        self.ctx.check(self.lib.otmb_set_slab(self.ctx.h, k0, k1))
        self.ctx.resident.clear()
        self.n_owned = 0

    def makeindices(self, v3D):
        N = C.c_int64()
        self.ctx.check(self.lib.otmb_makeindices(self.ctx.h, A._ptr(A._f64(v3D)), C.byref(N)))
        own, up = C.c_int64(), C.c_int64()
        self.ctx.check(self.lib.otmb_slab_counts(self.ctx.h, C.byref(own), C.byref(up)))
        self.n_owned = own.value
        return own.value, up.value

    def set_rank_offset(self, w0):
        self.ctx.check(self.lib.otmb_set_rank_offset(self.ctx.h, int(w0)))

    def set_metrics(self, gm):
        stack = lambda d: np.asfortranarray(np.stack([A._f64(d[k]) for k in A.DIRS], axis=-1))
        zt = np.ascontiguousarray(gm.zt, dtype=np.float64)
        self.ctx.check(self.lib.otmb_set_gridmetrics(
            self.ctx.h, A._ptr(A._f64(gm.area2D)), A._ptr(A._f64(gm.thkcello)), A._ptr(zt), A._ptr(stack(gm.edge_length_2D)),
            A._ptr(stack(gm.distance_to_neighbour_2D)), None, None, None))

    def facefluxes(self, umo, vmo, fill, carry_in, carry_out, outputs=None):
        """carry_in / carry_out: planes from Exchange.new_plane, or None at the chain's ends."""
        valid = (C.c_int32 * 2)()
        on_device = int(bool((carry_in or carry_out or (None, None, False))[2]))
        outs = [A._ptr(o) for o in outputs] if outputs is not None else [None] * 6
        self.ctx.check(self.lib.otmb_facefluxes_slab(
            self.ctx.h, A._ptr(A._f64(umo)), A._ptr(A._f64(vmo)), float(fill),
            C.c_void_p(carry_in[1]) if carry_in else None, C.c_void_p(carry_out[1]) if carry_out else None,
            on_device, valid, *outs))
        return bool(valid[0]), bool(valid[1])

    def set_facefluxes(self, phi):
        arrs = [A._f64(getattr(phi, k) if not isinstance(phi, dict) else phi[k]) for k in A.FACES]
        ptrs = (C.c_void_p * 6)(*[a.ctypes.data for a in arrs])
        self.ctx.check(self.lib.otmb_set_facefluxes(self.ctx.h, ptrs))

    def build(self, mlotst, rho, kH, kVML, kVdeep, upwind, upload=True):
        """Device assembly of this slab's columns; returns the five nnz.  upload=False re-runs the
        kernel on the resident inputs (benchmark loop)."""
        lib, ctx = self.lib, self.ctx
        if upload:
            ctx.check(lib.otmb_set_mlotst(ctx.h, A._ptr(A._f64(mlotst))))
            ctx.check(lib.otmb_set_rho3d(ctx.h, None if np.isscalar(rho) else A._ptr(A._f64(rho))))
        prm = _L.TMParams(float(kH), float(kVML), float(kVdeep), float(rho) if np.isscalar(rho) else 0.0, int(bool(upwind)), 0,
                          _L.PATH["fused"], 0)
        nnz = (C.c_int64 * 5)()
        ctx.check(lib.otmb_transportmatrix_build(ctx.h, C.byref(prm), nnz))
        self.nnz = [int(x) for x in nnz]
        return self.nnz

    def fetch(self):
        out = {}
        for m, name in enumerate(MATS):
            cp = np.empty(self.n_owned + 1, np.int64)
            rv, nz = np.empty(self.nnz[m], np.int64), np.empty(self.nnz[m], np.float64)
            self.ctx.check(self.lib.otmb_transportmatrix_fetch(self.ctx.h, m, A._ptr(cp), A._ptr(rv), A._ptr(nz)))
            out[name] = (cp, rv, nz)
        return out

    def transportmatrix(self, mlotst, rho, kH, kVML, kVdeep, upwind):
        self.build(mlotst, rho, kH, kVML, kVdeep, upwind)
        return self.fetch()


# ------------------------------------------------------------------------------------------
# the sharded driver (runs on every rank)
# ------------------------------------------------------------------------------------------
@dataclass
class ShardedCSC:
    """This rank's columns [col0, col0 + ncols) of an N x N CSC matrix, 0-based.  colptr already
    carries the global entry offset, so the ranks' (colptr[:-1], rowval, nzval) concatenate."""
    N: int
    col0: int
    colptr: np.ndarray
    rowval: np.ndarray
    nzval: np.ndarray

    @property
    def ncols(self):
        return len(self.colptr) - 1


def prepare_sharded(*, exchange, gridmetrics, umo=None, vmo=None, FillValue=None, ϕ=None, slab_factory=CudaSlab, device=0):
    """Steps 1-2 of the sharded assembly: slab plan, wet-rank offsets, metrics, face fluxes (chained
    continuity scan).  Returns (slab, w0, N, info); the slab is ready for build()."""
    ex = exchange
    v3D = gridmetrics.v3D
    nx, ny, nz = v3D.shape
    slabs = plan_slabs(wet_per_level(v3D), ex.size)
    k0, k1 = slabs[ex.rank]
    slab = slab_factory((nx, ny, nz), gridmetrics.gridtopology.kind, k0, k1, device=device)

    # 1. wet index offsets
    n_owned, _ = slab.makeindices(v3D)
    counts = [c[0] for c in ex.allgather_ints([n_owned])]
    w0, N = sum(counts[:ex.rank]), sum(counts)
    slab.set_rank_offset(w0)
    slab.set_metrics(gridmetrics)

    # 2. face fluxes: the continuity scan chains the slabs from the sea floor up
    if ϕ is not None:
        slab.set_facefluxes(ϕ)
    else:
        P = nx * ny
        carry_in = ex.new_plane(P) if ex.rank < ex.size - 1 else None
        carry_out = ex.new_plane(P) if ex.rank > 0 else None
        if carry_in is not None:
            ex.recv(ex.rank + 1, carry_in)
        vu, vv = slab.facefluxes(umo, vmo, FillValue, carry_in, carry_out)
        if carry_out is not None:
            ex.send(ex.rank - 1, carry_out)
        flags = ex.allgather_ints([int(vu), int(vv)])
        if not any(f[0] for f in flags) or not any(f[1] for f in flags):
            raise A.OTMBError(_L.ERR_ALL_FILL, "AssertionError: all umo/vmo values are NaN or FillValue")
    return slab, w0, N, dict(slabs=slabs, counts=counts, N=N)


def transportmatrix_sharded(*, exchange, gridmetrics, mlotst, ρ, umo=None, vmo=None, FillValue=None, ϕ=None, κH=500.0,
                            κVML=0.1, κVdeep=1.0e-5, upwind=True, slab_factory=CudaSlab, device=0, gather=True):
    """transportmatrix (src/matrixbuilding.jl:128-150) for ONE matrix sharded over exchange.size ranks.
    Pass either (umo, vmo, FillValue) — facefluxes run sharded too — or a precomputed ϕ.
    Returns (TransportMatrices of scipy CSC on rank 0 / None elsewhere when gather=True, dict of
    ShardedCSC segments, info)."""
    ex = exchange
    slab, w0, N, info = prepare_sharded(exchange=ex, gridmetrics=gridmetrics, umo=umo, vmo=vmo, FillValue=FillValue, ϕ=ϕ,
                                        slab_factory=slab_factory, device=device)

    # 3. this rank's columns
    local = slab.transportmatrix(mlotst, ρ, κH, κVML, κVdeep, upwind)

    # 4. entry offsets -> global colptr
    nnz_all = ex.allgather_ints([len(local[name][1]) for name in MATS])
    segs = {}
    for m, name in enumerate(MATS):
        cp, rv, nzv = local[name]
        off = sum(nnz_all[r][m] for r in range(ex.rank))
        segs[name] = ShardedCSC(N, w0, cp + off, rv, nzv)
    info.update(nnz=nnz_all)

    full = None
    if gather:
        mats = []
        for name in MATS:
            s = segs[name]
            cps = ex.gather_arrays(s.colptr[:-1])
            rvs = ex.gather_arrays(s.rowval)
            nzs = ex.gather_arrays(s.nzval)
            if ex.rank == 0:
                total = sum(nnz_all[r][MATS.index(name)] for r in range(ex.size))
                colptr = np.concatenate(cps + [np.array([total], np.int64)])
                mats.append(A._csc(N, colptr, np.concatenate(rvs), np.concatenate(nzs)))
        if ex.rank == 0:
            full = A.TransportMatrices(*mats)
    return full, segs, info


def run_threaded(nranks, fn):
    """Run fn(exchange) on `nranks` threads of this process (single-GPU emulation of the ranks).
    Returns the list of results by rank; re-raises the first exception."""
    shared = ThreadExchange._Shared(nranks)
    results, errors = [None] * nranks, [None] * nranks

    def work(r):
        try:
            results[r] = fn(ThreadExchange(shared, r))
        except BaseException as e:      # noqa: BLE001 - reported to the caller below
            errors[r] = e
            shared.barrier.abort()

    threads = [threading.Thread(target=work, args=(r,)) for r in range(nranks)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for e in errors:
        if e is not None and not isinstance(e, threading.BrokenBarrierError):
            raise e
    for e in errors:
        if e is not None:
            raise e
    return results
