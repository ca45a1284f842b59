"""Row-slab sharding of ONE transport matrix across GPUs (BASELINE config 4).

The reference is single-process; what is sharded here is its loop over wet cells
(`for 𝑖 in eachindex(Lwet)`, /root/reference/src/matrixbuilding.jl:237, 348, 450).  Wet ranks
are ordered by linear index (`Lwet = L[wet3D]`, :14-16), so a contiguous range of GRID ROWS
(row = j + ny*k) is a contiguous block of rows/columns of every matrix.  Rank r owns the rows
[row_cuts[r], row_cuts[r+1]) chosen so that every rank holds N/R wet cells to within one grid
row (`otmb_plan_slabs`; cuts may fall inside a level), keeps one level of halo on either side
resident, and assembles the CSC *columns* of its cells with global row indices; the complete
matrix is the concatenation of the ranks' (rowval, nzval) segments, with each local colptr
shifted by the number of entries of the lower ranks.

Exchanges between ranks (everything else is rank-local):
  1. all-gather of one integer per rank — owned wet cells -> global wet-rank offsets;
  2. the face-flux continuity scan (/root/reference/src/velocities.jl:234-243) runs bottom-up
     and is a floating-point recurrence, ϕtop[k] = ((((ϕtop[k+1] + w) + s) - e) - n): it cannot be
     re-associated without changing bits, so the slabs form a chain — the rank below hands the
     plane of its topmost ϕtop values (nx*ny doubles) to the rank above.  Skipped when the caller
     passes ϕ itself;
  3. all-gather of status + five integers per rank — nnz per matrix -> colptr offsets, and a
     failure on any rank is raised on every rank;
  4. optionally a gather of the finished segments on rank 0 (tests / small cases; in production
     every rank copies its segment to its place in the host arrays).

Two drivers:
  * `transportmatrix_sharded_native` — the production path: steps 1-3 run INSIDE libotmb.so over its own
    NCCL communicator (csrc/comm.cu: chunk-pipelined carry plane on the context's stream, no torch, no host
    synchronisation between the ranks' kernels).  The host only distributes the 128-byte communicator id.
  * `transportmatrix_sharded` — the same logic with the exchanges done by the host through an `Exchange`
    object (torch.distributed gloo / NCCL, or threads inside one process for the single-GPU emulation the GPU
    tests use) and the compute class injected, so the host logic can be exercised without a GPU by a test
    double for the compute.  The product default is `CudaSlab` and fails loudly without the library.
"""
from __future__ import annotations

import ctypes as C
import queue
import threading
from dataclasses import dataclass

import numpy as np

from . import _lib as _L
from . import api as A

MATS = A.MATRICES


# ------------------------------------------------------------------------------------------
# partition
# ------------------------------------------------------------------------------------------
def plan_slabs(v3D, nranks, level_cuts_only=False):
    """Contiguous grid-row ranges [(row0, row1), ...], one per rank, cut where the cumulative wet count is
    closest to r*N/R (`otmb_plan_slabs`: host code of libotmb.so, no GPU needed).  Deterministic: every rank
    computes the same plan from the same array.  Returns (slabs, wet_per_rank)."""
    v = A._f64(v3D)
    nx, ny, nz = v.shape
    cuts = (C.c_int64 * (nranks + 1))()
    wet = (C.c_int64 * nranks)()
    st = _L.load().otmb_plan_slabs(A._ptr(v), nx, ny, nz, int(nranks), int(bool(level_cuts_only)), cuts, wet)
    if st != _L.OK:
        units = nz if level_cuts_only else ny * nz
        raise ValueError(f"need 1 <= ranks <= {units} ({'levels' if level_cuts_only else 'grid rows'}), got {nranks}")
    return [(int(cuts[r]), int(cuts[r + 1])) for r in range(nranks)], [int(w) for w in wet]


def column_levels(rows, ny):
    """Owned levels [k_begin[j], k_end[j]) of the columns of grid row-in-level j, for the slab rows = (row0, row1)
    (the same arithmetic as k_faceflux, csrc/faceflux.cu)."""
    j = np.arange(ny)
    return (rows[0] - j + ny - 1) // ny, (rows[1] - j + ny - 1) // ny


# ------------------------------------------------------------------------------------------
# exchange plumbing (host-driven driver)
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

    def broadcast_bytes(self, data, src=0):
        box = [data if self.rank == src else None]
        self.dist.broadcast_object_list(box, src=src)
        return box[0]


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
    """One libotmb.so context restricted to the grid rows [row0, row1) (otmb_set_slab_rows)."""

    def __init__(self, shape, topology, row0, row1, device=0, ctx=None):
        self.ctx = ctx or A.Context(device)
        self.lib = self.ctx.lib
        self.shape, self.rows = tuple(shape), (int(row0), int(row1))
        nx, ny, nz = self.shape
        self.ctx.check(self.lib.otmb_set_grid(self.ctx.h, nx, ny, nz, _L.TOPO[topology]))
        self.ctx.check(self.lib.otmb_set_slab_rows(self.ctx.h, int(row0), int(row1)))
        self.ctx.resident.clear()
        self.n_owned = 0
        self.nnz = [0] * 5

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

    def upload_build_inputs(self, mlotst, rho):
        self.ctx.check(self.lib.otmb_set_mlotst(self.ctx.h, A._ptr(A._f64(mlotst))))
        self.ctx.check(self.lib.otmb_set_rho3d(self.ctx.h, None if np.isscalar(rho) else A._ptr(A._f64(rho))))

    def params(self, rho, kH, kVML, kVdeep, upwind):
        return _L.TMParams(float(kH), float(kVML), float(kVdeep), float(rho) if np.isscalar(rho) else 0.0, int(bool(upwind)), 0,
                           _L.PATH["fused"], 0)

    def build(self, mlotst, rho, kH, kVML, kVdeep, upwind, upload=True):
        """Device assembly of this slab's columns; returns the five nnz.  upload=False re-runs the
        kernel on the resident inputs (benchmark loop)."""
        if upload:
            self.upload_build_inputs(mlotst, rho)
        prm = self.params(rho, kH, kVML, kVdeep, upwind)
        nnz = (C.c_int64 * 5)()
        self.ctx.check(self.lib.otmb_transportmatrix_build(self.ctx.h, C.byref(prm), nnz))
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
# the sharded drivers (run on every rank)
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


def _agree(ex, step):
    """Run the rank-local callable `step`; all-gather (status, code) BEFORE anyone moves on to the next
    send/recv or collective, and raise the first failure on EVERY rank (the failing rank re-raises its own
    exception, the others an OTMBError naming it).  Returns step()'s result."""
    result, err = None, None
    try:
        result = step()
    except Exception as e:      # noqa: BLE001 - reported on every rank below
        err = e
    codes = ex.allgather_ints([0 if err is None else int(getattr(err, "code", -1)) or -1])
    bad = [(r, c[0]) for r, c in enumerate(codes) if c[0] != 0]
    if err is not None:
        raise err
    if bad:
        r, code = bad[0]
        msg = _L.load().otmb_status_string(code).decode() if code > 0 else "error"
        raise A.OTMBError(code, f"rank {r}: {msg}")
    return result


def prepare_sharded(*, exchange, gridmetrics, umo=None, vmo=None, FillValue=None, ϕ=None, slab_factory=CudaSlab, device=0,
                    level_cuts_only=False):
    """Steps 1-2 of the sharded assembly: slab plan, wet-rank offsets, metrics, face fluxes (chained
    continuity scan).  Returns (slab, w0, N, info); the slab is ready for build()."""
    ex = exchange
    v3D = gridmetrics.v3D
    nx, ny, nz = v3D.shape
    slabs, wet_plan = plan_slabs(v3D, ex.size, level_cuts_only)
    row0, row1 = slabs[ex.rank]
    slab = slab_factory((nx, ny, nz), gridmetrics.gridtopology.kind, row0, row1, device=device)

    # 1. wet index offsets
    n_owned, _ = _agree(ex, lambda: slab.makeindices(v3D))
    counts = [c[0] for c in ex.allgather_ints([n_owned])]
    assert counts == wet_plan, "device wet counts differ from the host plan"
    w0, N = sum(counts[:ex.rank]), sum(counts)

    def offsets_and_metrics():
        slab.set_rank_offset(w0)
        slab.set_metrics(gridmetrics)
    _agree(ex, offsets_and_metrics)

    # 2. face fluxes: the continuity scan chains the slabs from the sea floor up
    if ϕ is not None:
        _agree(ex, lambda: slab.set_facefluxes(ϕ))
    else:
        P = nx * ny
        carry_in = ex.new_plane(P) if ex.rank < ex.size - 1 else None
        carry_out = ex.new_plane(P) if ex.rank > 0 else None
        if carry_in is not None:
            ex.recv(ex.rank + 1, carry_in)
        err, valid = None, (False, False)
        try:
            valid = slab.facefluxes(umo, vmo, FillValue, carry_in, carry_out)
        except Exception as e:      # noqa: BLE001
            err = e
        if carry_out is not None:
            ex.send(ex.rank - 1, carry_out)        # also after a failure: the ranks above must not block in recv
        flags = ex.allgather_ints([int(valid[0]), int(valid[1]), 0 if err is None else int(getattr(err, "code", -1)) or -1])
        if err is not None:
            raise err
        for r, f in enumerate(flags):
            if f[2] != 0:
                raise A.OTMBError(f[2], f"rank {r}: {_L.load().otmb_status_string(f[2]).decode() if f[2] > 0 else 'error'}")
        if not any(f[0] for f in flags) or not any(f[1] for f in flags):
            raise A.OTMBError(_L.ERR_ALL_FILL, "AssertionError: all umo/vmo values are NaN or FillValue")
    return slab, w0, N, dict(slabs=slabs, counts=counts, N=N)


def _gather_full(ex, segs, nnz_all, N):
    mats = []
    for m, name in enumerate(MATS):
        s = segs[name]
        cps = ex.gather_arrays(s.colptr[:-1])
        rvs = ex.gather_arrays(s.rowval)
        nzs = ex.gather_arrays(s.nzval)
        if ex.rank == 0:
            total = sum(nnz_all[r][m] for r in range(ex.size))
            colptr = np.concatenate(cps + [np.array([total], np.int64)])
            mats.append(A._csc(N, colptr, np.concatenate(rvs), np.concatenate(nzs)))
    return A.TransportMatrices(*mats) if ex.rank == 0 else None


def transportmatrix_sharded(*, exchange, gridmetrics, mlotst, ρ, umo=None, vmo=None, FillValue=None, ϕ=None, κH=500.0,
                            κVML=0.1, κVdeep=1.0e-5, upwind=True, slab_factory=CudaSlab, device=0, gather=True,
                            level_cuts_only=False):
    """transportmatrix (src/matrixbuilding.jl:128-150) for ONE matrix sharded over exchange.size ranks, the
    exchanges done by the host.  Pass either (umo, vmo, FillValue) — facefluxes run sharded too — or a precomputed ϕ.
    Returns (TransportMatrices of scipy CSC on rank 0 / None elsewhere when gather=True, dict of
    ShardedCSC segments, info).  An error of any rank (NaN checks, dry-neighbour flux, CUDA) is raised on every rank."""
    ex = exchange
    slab, w0, N, info = prepare_sharded(exchange=ex, gridmetrics=gridmetrics, umo=umo, vmo=vmo, FillValue=FillValue, ϕ=ϕ,
                                        slab_factory=slab_factory, device=device, level_cuts_only=level_cuts_only)

    # 3. this rank's columns
    local = _agree(ex, lambda: slab.transportmatrix(mlotst, ρ, κH, κVML, κVdeep, upwind))

    # 4. entry offsets -> global colptr
    nnz_all = ex.allgather_ints([len(local[name][1]) for name in MATS])
    segs = {}
    for m, name in enumerate(MATS):
        cp, rv, nzv = local[name]
        off = sum(nnz_all[r][m] for r in range(ex.rank))
        segs[name] = ShardedCSC(N, w0, cp + off, rv, nzv)
    info.update(nnz=nnz_all)
    full = _gather_full(ex, segs, nnz_all, N) if gather else None
    return full, segs, info


class NativeSharded:
    """The production driver: one slab context per rank whose exchanges run inside libotmb.so over its own NCCL
    communicator (csrc/comm.cu).  `id_bytes`: the 128-byte id from `NativeSharded.unique_id()` on one rank, handed
    to every rank by the host program."""

    @staticmethod
    def unique_id():
        buf = (C.c_uint8 * 128)()
        st = _L.load().otmb_comm_unique_id(buf)
        if st != _L.OK:
            raise A.OTMBError(st, _L.load().otmb_status_string(st).decode())
        return bytes(buf)

    def __init__(self, *, gridmetrics, rank, nranks, id_bytes, device=0, level_cuts_only=False):
        v3D = gridmetrics.v3D
        nx, ny, nz = v3D.shape
        self.rank, self.size = rank, nranks
        self.slabs, self.wet_plan = plan_slabs(v3D, nranks, level_cuts_only)
        self.slab = CudaSlab((nx, ny, nz), gridmetrics.gridtopology.kind, *self.slabs[rank], device=device)
        ctx, lib = self.slab.ctx, self.slab.lib
        self.ctx, self.lib = ctx, lib
        idb = (C.c_uint8 * 128).from_buffer_copy(id_bytes) if id_bytes is not None else None
        ctx.check(lib.otmb_comm_init(ctx.h, nranks, rank, idb))
        N, w0, own = C.c_int64(), C.c_int64(), C.c_int64()
        ctx.check(lib.otmb_sharded_makeindices(ctx.h, A._ptr(A._f64(v3D)), C.byref(N), C.byref(w0), C.byref(own)))
        self.N, self.w0, self.slab.n_owned = N.value, w0.value, own.value
        assert own.value == self.wet_plan[rank], "device wet count differs from the host plan"
        self.slab.set_metrics(gridmetrics)

    def set_masstransport(self, umo, vmo, fill):
        self.ctx.check(self.lib.otmb_set_masstransport(self.ctx.h, A._ptr(A._f64(umo)), A._ptr(A._f64(vmo)), float(fill)))

    def facefluxes(self, nchunks=0, outputs=None):
        outs = [A._ptr(o) for o in outputs] if outputs is not None else [None] * 6
        self.ctx.check(self.lib.otmb_sharded_facefluxes(self.ctx.h, int(nchunks), *outs))

    def facefluxes_enqueue(self, nchunks=0):
        self.ctx.check(self.lib.otmb_sharded_facefluxes_enqueue(self.ctx.h, int(nchunks)))

    def build(self, mlotst, rho, kH=500.0, kVML=0.1, kVdeep=1.0e-5, upwind=True, upload=True):
        if upload:
            self.slab.upload_build_inputs(mlotst, rho)
        prm = self.slab.params(rho, kH, kVML, kVdeep, upwind)
        loc, before, total = (C.c_int64 * 5)(), (C.c_int64 * 5)(), (C.c_int64 * 5)()
        self.ctx.check(self.lib.otmb_sharded_transportmatrix_build(self.ctx.h, C.byref(prm), loc, before, total))
        self.slab.nnz = [int(x) for x in loc]
        self.nnz_before, self.nnz_total = [int(x) for x in before], [int(x) for x in total]
        return self.slab.nnz

    def segments(self):
        """This rank's columns as ShardedCSC (colptr carries the global entry offset)."""
        local = self.slab.fetch()
        return {name: ShardedCSC(self.N, self.w0, local[name][0] + self.nnz_before[m], local[name][1], local[name][2])
                for m, name in enumerate(MATS)}

    def close(self):
        self.lib.otmb_comm_free(self.ctx.h)
        self.ctx.close()


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
