"""Host-side mirror of the reference's exported API, forwarding to libotmb.so.

The reference is a Julia package whose public surface is seven exported functions
(/root/reference/src/OceanTransportMatrixBuilder.jl:31-36).  Julia is not available in this
image, so this Python module plays the role of the thin host shim (the Julia `ccall`
version of the same shim is in INTEGRATION.md): same function names, same keyword
arguments (including the Unicode ones: ϕ, ρ, κH, κVML, κVdeep, TκH, ...), same argument
meaning, same error behaviour and messages, same result fields.  All numerical work is
done by the CUDA kernels behind the C ABI; nothing here computes on the CPU beyond the
host-only steps the reference itself performs on tiny inputs (missing/_FillValue -> NaN
cleaning, vertexpermutation, getgridtopology — SURVEY.md §8a row A3).

Arrays follow the reference's layout: Fortran-ordered `(nx, ny, nz)`, `(nx, ny)`,
`(4, nx, ny)`.  Sparse results are `scipy.sparse.csc_matrix` with float64 data and int64
indices (the SparseMatrixCSC{Float64,Int64} of the reference, 0-based).
"""
from __future__ import annotations

import ctypes as C
import weakref
from collections import namedtuple
from dataclasses import dataclass

import numpy as np
import scipy.sparse as sp

from . import _lib as _L

DIRS = ("south", "east", "north", "west")                          # src/gridcellgeometry.jl:304
FACES = ("east", "west", "north", "south", "top", "bottom")        # src/velocities.jl:245-252
MATRICES = ("T", "Tadv", "TκH", "TκVML", "TκVdeep")                # src/matrixbuilding.jl:149


class OTMBError(RuntimeError):
    """An error of the reference (same message) or of the CUDA layer."""

    def __init__(self, code, message):
        super().__init__(message)
        self.code = code


class Field:
    """Stand-in for the YAXArray the reference receives: `.data` plus `.properties`
    (the reference reads `.properties["_FillValue"]`, src/gridcellgeometry.jl:270, src/velocities.jl:120)."""

    def __init__(self, data, properties=None):
        self.data = np.asanyarray(data)
        self.properties = dict(properties or {})

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.data, dtype=dtype)


def _data(x):
    return x.data if isinstance(x, Field) else np.asanyarray(x)


def _props(x):
    return getattr(x, "properties", {}) or {}


def _f64(a):
    return np.asfortranarray(a, dtype=np.float64)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _freeze(*arrays):
    """Mark arrays this module hands out as read-only.  Device residency is keyed on object identity, so an array
    whose content is resident must not change behind the library's back: an in-place update of a result of
    makegridmetrics / facefluxes raises in numpy instead of silently assembling from the stale device copy.
    Arrays the CALLER owns (writeable) are never trusted to be unchanged: they are uploaded on every call, like
    the reference, which reads its arguments at call time."""
    for a in arrays:
        a.flags.writeable = False


def _same_frozen(a, b):
    return a is b and isinstance(a, np.ndarray) and not a.flags.writeable


class Context:
    """One otmb_ctx (one GPU).  Holds the device-resident grid, indices, metrics and ϕ."""

    def __init__(self, device=0):
        self.lib = _L.load()
        h = C.c_void_p()
        st = self.lib.otmb_create(C.byref(h), int(device))
        if st != _L.OK:
            raise OTMBError(st, self.lib.otmb_status_string(st).decode())
        self.h = h
        self.device = device
        self.resident = {}       # name -> object whose content is resident on the device
        self._fin = weakref.finalize(self, self.lib.otmb_destroy, h)

    def check(self, st):
        if st != _L.OK:
            msg = self.lib.otmb_last_error(self.h).decode() or self.lib.otmb_status_string(st).decode()
            raise OTMBError(st, msg)

    def close(self):
        self._fin()

    # results land in page-locked host memory (full PCIe rate instead of the ~5 GB/s of a pageable copy).
    # Buffers return to a per-context pool when the numpy array that wraps them is garbage collected,
    # so a loop over months reuses the same memory.
    def pinned_empty(self, n, dtype, shape=None, order="C"):
        dtype = np.dtype(dtype)
        nbytes = max(int(n) * dtype.itemsize, 8)
        free = self.__dict__.setdefault("_pinned_free", [])          # [(capacity, ptr)], best fit within 1.5x
        fit = [e for e in free if nbytes <= e[0] <= nbytes + nbytes // 2 + (1 << 20)]
        if fit:
            entry = min(fit)
            free.remove(entry)
        else:
            p = C.c_void_p()
            st = self.lib.otmb_host_alloc(C.byref(p), nbytes)
            if st != _L.OK:                      # out of pinnable memory: an ordinary array still works
                a = np.empty(int(n), dtype)
                return a if shape is None else a.reshape(shape, order=order)
            entry = (nbytes, p.value)
        buf = (C.c_char * entry[0]).from_address(entry[1])
        arr = np.frombuffer(buf, dtype=dtype, count=int(n))
        weakref.finalize(buf, free.append, entry)   # `arr` (and every view of it) keeps `buf` alive
        return arr if shape is None else arr.reshape(shape, order=order)

    # measurement helpers
    def launches(self):
        n = C.c_int64()
        self.check(self.lib.otmb_launch_count(self.h, C.byref(n)))
        return n.value

    def last_build_ms(self):
        ms = C.c_float()
        self.check(self.lib.otmb_last_build_ms(self.h, C.byref(ms)))
        return ms.value


_contexts = {}


def default_context(device=0) -> Context:
    if device not in _contexts:
        _contexts[device] = Context(device)
    return _contexts[device]


# ------------------------------------------------------------------------------------------
# host-only pieces of makegridmetrics (tiny inputs; the reference does them on the host too)
# ------------------------------------------------------------------------------------------
def _clean_missing(x, fills):
    """src/gridcellgeometry.jl:269-280: missing/nothing/0/_FillValue -> NaN.  `replace` matches with
    isequal, so -0.0 is kept."""
    d = _data(x)
    if np.ma.isMaskedArray(d):                      # `missing`
        d = np.ma.filled(d.astype(np.float64), np.nan)
    a = np.array(d, dtype=np.float64, order="F")
    bad = (a == 0.0) & ~np.signbit(a)
    for fv in fills:
        bad |= a == np.float64(fv)
    a[bad] = np.nan
    return a


def vertexpermutation(lon_vertices, lat_vertices):
    """src/gridcellgeometry.jl:158-178 (0-based)."""
    lonv, latv = np.asarray(lon_vertices), np.asarray(lat_vertices)
    assert lonv.shape[0] == latv.shape[0] == 4
    pts = [(lonv[v, 0, 0], latv[v, 0, 0]) for v in range(4)]
    pe = {(lonv[v, 1, 0], latv[v, 1, 0]) for v in range(4)}
    pn = {(lonv[v, 0, 1], latv[v, 0, 1]) for v in range(4)}
    idx_e = [q for q, p in enumerate(pts) if p in pe]
    idx_n = [q for q, p in enumerate(pts) if p in pn]
    (i3,) = [q for q in idx_e if q in idx_n]
    (i2,) = [q for q in idx_e if q != i3]
    (i4,) = [q for q in idx_n if q != i3]
    (i1,) = [q for q in range(4) if q not in (i2, i3, i4)]
    return [i1, i2, i3, i4]


@dataclass(frozen=True)
class GridTopology:
    """BipolarGridTopology / TripolarGridTopology / UnknownGridTopology, src/gridtopology.jl:1-16."""
    kind: str
    nx: int
    ny: int
    nz: int


def getgridtopology(lon_vertices, lat_vertices, lev) -> GridTopology:
    """src/gridtopology.jl:33-53."""
    lonv, latv = np.asarray(lon_vertices), np.asarray(lat_vertices)
    nx, ny, nz = lonv.shape[1], lonv.shape[2], len(lev)
    NPlon, NPlat = lonv[2:4, :, -1], latv[2:4, :, -1]
    if np.all(NPlat == 90):
        return GridTopology("bipolar", nx, ny, nz)
    rot = lambda a: a[::-1, ::-1]
    d = np.mod(NPlon - rot(NPlon) + 180, 360) - 180
    lon_ok = np.linalg.norm(d) <= np.spacing(180.0)
    a, b = NPlat, rot(NPlat)
    lat_ok = np.linalg.norm(a - b) <= np.sqrt(np.finfo(float).eps) * max(np.linalg.norm(a), np.linalg.norm(b))
    if lon_ok and lat_ok:
        return GridTopology("tripolar", nx, ny, nz)
    import warnings
    warnings.warn("Unknown grid topology detected. Things might not work as expected.\n"
                  "See `getgridtopology` function to see what failed the checks")
    return GridTopology("unknown", nx, ny, nz)


GridMetrics = namedtuple("GridMetrics", "area2D v3D thkcello lon_vertices lat_vertices lon lat Z3D zt edge_length_2D "
                                        "distance_to_edge_2D distance_to_neighbour_2D gridtopology")
Indices = namedtuple("Indices", "wet3D L Lwet N Lwet3D C")
FaceFluxes = namedtuple("FaceFluxes", "east west north south top bottom")
TransportMatrices = namedtuple("TransportMatrices", "T Tadv TκH TκVML TκVdeep")

_owner = weakref.WeakValueDictionary()     # id(result array) -> Context that produced it


def _ctx_of(*objs, ctx=None):
    if ctx is not None:
        return ctx
    for o in objs:
        c = _owner.get(id(o))
        if c is not None:
            return c
    return default_context()


class _Lazy:
    """LinearIndices / CartesianIndices of the reference (lazy, no storage), 1-based."""

    def __init__(self, shape, cartesian):
        self.shape, self.cartesian = tuple(shape), cartesian

    def __getitem__(self, key):
        nx, ny, _ = self.shape
        if self.cartesian:
            L = int(key) - 1
            return (L % nx + 1, (L // nx) % ny + 1, L // (nx * ny) + 1)
        i, j, k = key
        return i + nx * (j - 1) + nx * ny * (k - 1)


def _ensure_indices(ctx, v3D, topo_kind):
    """Make v3D / mask / ranks resident (otmb_set_grid + otmb_makeindices) unless they already are."""
    key = ("v3D", id(v3D), topo_kind)
    if ctx.resident.get("v3D_key") == key and _same_frozen(ctx.resident.get("v3D"), v3D):
        return ctx.resident["N"]
    nx, ny, nz = v3D.shape
    ctx.check(ctx.lib.otmb_set_grid(ctx.h, nx, ny, nz, _L.TOPO[topo_kind]))
    N = C.c_int64()
    ctx.check(ctx.lib.otmb_makeindices(ctx.h, _ptr(v3D), C.byref(N)))
    ctx.resident.clear()
    ctx.resident.update(v3D=v3D, v3D_key=key, N=N.value)
    return N.value


def makegridmetrics(*, areacello, volcello, lon, lat, lev, lon_vertices, lat_vertices, ctx=None) -> GridMetrics:
    """makegridmetrics, src/gridcellgeometry.jl:265-311."""
    ctx = ctx or default_context()
    fills = [x.properties["_FillValue"] for x in (areacello, volcello) if "_FillValue" in _props(x)]
    v3D = _clean_missing(volcello, fills)
    area2D = _clean_missing(areacello, fills)
    zt = np.array(_data(lev), dtype=np.float64, order="C", copy=True)      # own copy: it is frozen below
    lat_a, lon_a = _f64(_data(lat)), _f64(_data(lon))
    lonv, latv = _f64(_data(lon_vertices)), _f64(_data(lat_vertices))
    perm = vertexpermutation(lonv, latv)
    lonv, latv = np.asfortranarray(lonv[perm]), np.asfortranarray(latv[perm])
    topo = getgridtopology(lonv, latv, zt)
    nx, ny, nz = v3D.shape
    _ensure_indices(ctx, v3D, topo.kind)
    thk, Z3D = np.empty_like(v3D), np.empty_like(v3D)
    edge = np.empty((nx, ny, 4), order="F")
    dedge = np.empty((nx, ny, 4), order="F")
    dnbr = np.empty((nx, ny, 4), order="F")
    ctx.check(ctx.lib.otmb_gridmetrics(ctx.h, _ptr(area2D), _ptr(lon_a), _ptr(lat_a), _ptr(lonv), _ptr(latv), _ptr(zt),
                                       _ptr(thk), _ptr(Z3D), _ptr(edge), _ptr(dedge), _ptr(dnbr)))
    _freeze(v3D, area2D, thk, Z3D, zt, edge, dedge, dnbr)      # resident: see _freeze
    as_dict = lambda a: {d: a[:, :, q] for q, d in enumerate(DIRS)}
    gm = GridMetrics(area2D, v3D, thk, lonv, latv, lon_a, lat_a, Z3D, zt, as_dict(edge), as_dict(dedge), as_dict(dnbr), topo)
    ctx.resident["metrics_src"] = (gm.area2D, gm.thkcello, gm.zt, gm.edge_length_2D, gm.distance_to_neighbour_2D)
    _owner[id(v3D)] = ctx
    return gm


def makeindices(v3D, ctx=None, topology="bipolar") -> Indices:
    """makeindices(v3D), src/matrixbuilding.jl:10-24.  Lwet3D uses 0 for `missing`."""
    v3D = v3D if (isinstance(v3D, np.ndarray) and v3D.flags.f_contiguous and v3D.dtype == np.float64) else _f64(v3D)
    ctx = _ctx_of(v3D, ctx=ctx)
    kind = ctx.resident["v3D_key"][2] if _same_frozen(ctx.resident.get("v3D"), v3D) else topology
    N = _ensure_indices(ctx, v3D, kind)
    M = v3D.size
    chunks = np.zeros((M + 63) // 64, np.uint64)
    Lwet = np.empty(N, np.int64)
    Lwet3D = np.empty(v3D.shape, np.int64, order="F")
    ctx.check(ctx.lib.otmb_get_indices(ctx.h, _ptr(chunks), _ptr(Lwet), _ptr(Lwet3D)))
    wet3D = np.unpackbits(chunks.view(np.uint8), bitorder="little")[:M].astype(bool).reshape(v3D.shape, order="F")
    ix = Indices(wet3D, _Lazy(v3D.shape, False), Lwet, N, Lwet3D, _Lazy(v3D.shape, True))
    _owner[id(Lwet)] = ctx
    return ix


def _ensure_grid(ctx, gridmetrics):
    _ensure_indices(ctx, gridmetrics.v3D, gridmetrics.gridtopology.kind)


def facefluxesfrommasstransport(*, umo, vmo, gridmetrics, indices, ctx=None) -> FaceFluxes:
    """facefluxesfrommasstransport, src/velocities.jl:118-130 (+ facefluxes :190-255)."""
    fill = umo.properties["_FillValue"]                  # KeyError like the reference
    fv = vmo.properties["_FillValue"]
    assert (fill == fv) or (fill != fill and fv != fv), "AssertionError: isequal(FillValue, vmo.properties[\"_FillValue\"])"
    return facefluxes(_f64(_data(umo)), _f64(_data(vmo)), gridmetrics, indices, FillValue=fill, ctx=ctx)


def facefluxes(umo, vmo, gridmetrics, indices, *, FillValue, ctx=None) -> FaceFluxes:
    """facefluxes, src/velocities.jl:190-255 (umo/vmo are not modified, unlike nofluxboundaries!)."""
    ctx = _ctx_of(gridmetrics.v3D, ctx=ctx)
    _ensure_grid(ctx, gridmetrics)
    umo, vmo = _f64(umo), _f64(vmo)
    shape = gridmetrics.v3D.shape
    out = [ctx.pinned_empty(gridmetrics.v3D.size, np.float64, shape, "F") for _ in range(6)]
    ctx.check(ctx.lib.otmb_facefluxes(ctx.h, _ptr(umo), _ptr(vmo), float(FillValue), *[_ptr(o) for o in out]))
    _freeze(*out)                                   # resident as ϕ: see _freeze
    phi = FaceFluxes(*out)
    ctx.resident["phi_arrays"] = tuple(out)
    return phi


def _transportmatrix_stream(ctx, N, phi, mlotst, ρ, κH, κVML, κVdeep, upwind, nslabs=0, pageable=False, caps=None):
    """otmb_transportmatrix_stream: result arrays sized by the upper bound N x {7,7,5,3,3} (page-locked, pooled), the
    matrices are views of their first nnz entries."""
    lib = ctx.lib
    ml = _data(mlotst)
    ml = _f64(np.ma.filled(ml.astype(np.float64), np.nan) if np.ma.isMaskedArray(ml) else ml)
    rho3 = None if np.isscalar(ρ) else _f64(_data(ρ))
    prm = _L.TMParams(float(κH), float(κVML), float(κVdeep), float(ρ) if np.isscalar(ρ) else 0.0, int(bool(upwind)), 0,
                      _L.PATH["fused"], 0)
    caps = caps or [N * w for w in (7, 7, 5, 3, 3)]
    new = (lambda n, dt: np.empty(max(int(n), 1), dt)) if pageable else ctx.pinned_empty
    arrays = [(new(N + 1, np.int64), new(caps[m], np.int64), new(caps[m], np.float64)) for m in range(5)]
    ptrs = [(C.c_void_p * 5)(*[_ptr(arrays[m][q]).value for m in range(5)]) for q in range(3)]
    phi_ptrs = (C.c_void_p * 6)(*[a.ctypes.data for a in phi])
    nnz = (C.c_int64 * 5)()
    ctx.check(lib.otmb_transportmatrix_stream(ctx.h, C.byref(prm), phi_ptrs, _ptr(ml), _ptr(rho3), int(nslabs),
                                              (C.c_int64 * 5)(*caps), *ptrs, nnz))
    ctx.resident["phi_arrays"] = None
    return TransportMatrices(*[_csc(N, arrays[m][0], arrays[m][1][:nnz[m]], arrays[m][2][:nnz[m]]) for m in range(5)])


def _metric_dicts_same(a, b):
    return a is b and all(isinstance(v, np.ndarray) and not v.flags.writeable for v in b.values())


def _csc(n, colptr, rowval, nzval):
    m = sp.csc_matrix((n, n), dtype=np.float64)
    m.data, m.indices, m.indptr = nzval, rowval, colptr      # keep Int64 indices like the reference
    return m


def transportmatrix(*, ϕ, mlotst, gridmetrics, indices, ρ, κH=500.0, κVML=0.1, κVdeep=1.0e-5, Tadv=None, TκH=None,
                    TκVML=None, TκVdeep=None, upwind=True, path="fused", ctx=None) -> TransportMatrices:
    """transportmatrix, src/matrixbuilding.jl:128-150.  `path` selects the device strategy
    ("fused", "fused2", "coo"); all give identical results."""
    ctx = _ctx_of(gridmetrics.v3D, ctx=ctx)
    lib = ctx.lib
    _ensure_grid(ctx, gridmetrics)
    N = ctx.resident["N"]
    # grid metrics: resident if these are the very objects makegridmetrics returned on this ctx
    src = ctx.resident.get("metrics_src")
    # (the dicts hold read-only views of frozen arrays; a caller who rebinds an entry gets a new dict value -> upload)
    same = src is not None and _same_frozen(src[0], gridmetrics.area2D) and _same_frozen(src[1], gridmetrics.thkcello) \
        and _same_frozen(src[2], gridmetrics.zt) and _metric_dicts_same(src[3], gridmetrics.edge_length_2D) \
        and _metric_dicts_same(src[4], gridmetrics.distance_to_neighbour_2D)
    if not same:
        stack = lambda d: np.asfortranarray(np.stack([_f64(d[k]) for k in DIRS], axis=-1))
        edge, dnbr = stack(gridmetrics.edge_length_2D), stack(gridmetrics.distance_to_neighbour_2D)
        zt = np.ascontiguousarray(gridmetrics.zt, dtype=np.float64)
        ctx.check(lib.otmb_set_gridmetrics(ctx.h, _ptr(_f64(gridmetrics.area2D)), _ptr(_f64(gridmetrics.thkcello)), _ptr(zt),
                                           _ptr(edge), _ptr(dnbr), _ptr(_f64(gridmetrics.Z3D)), _ptr(_f64(gridmetrics.lon)),
                                           _ptr(_f64(gridmetrics.lat))))
        ctx.resident["metrics_src"] = None          # caller-owned arrays: never assumed unchanged
    preset = {1: Tadv, 2: TκH, 3: TκVML, 4: TκVdeep}
    mask = 32 | sum(1 << m for m, v in preset.items() if v is None)   # bit 5: the mask is explicit
    get = (lambda k: ϕ[k]) if isinstance(ϕ, dict) else (lambda k: getattr(ϕ, k))
    res = ctx.resident.get("phi_arrays")
    phi_resident = res is not None and all(_same_frozen(get(k), b) for k, b in zip(FACES, res))
    if path == "fused" and all(v is None for v in preset.values()) and not phi_resident and N > 0:
        # caller-owned ϕ and all four operators to build: upload, assembly and copy-out overlap inside ONE call
        return _transportmatrix_stream(ctx, N, [_f64(get(k)) for k in FACES], mlotst, ρ, κH, κVML, κVdeep, upwind)
    if mask & 2:
        arrs = tuple(get(k) for k in FACES)
        if not phi_resident:
            arrs = tuple(_f64(a) for a in arrs)
            ptrs = (C.c_void_p * 6)(*[a.ctypes.data for a in arrs])
            ctx.check(lib.otmb_set_facefluxes(ctx.h, ptrs))
            ctx.resident["phi_arrays"] = None
        if np.isscalar(ρ):
            ctx.check(lib.otmb_set_rho3d(ctx.h, None))
        else:
            ctx.check(lib.otmb_set_rho3d(ctx.h, _ptr(_f64(_data(ρ)))))
    if mask & 8:
        ml = _data(mlotst)
        ml = np.ma.filled(ml.astype(np.float64), np.nan) if np.ma.isMaskedArray(ml) else ml
        ctx.check(lib.otmb_set_mlotst(ctx.h, _ptr(_f64(ml))))
    for m, v in preset.items():
        if v is not None:
            v = sp.csc_matrix(v)
            # no copies when the operator already has the API's types (e.g. a result of an earlier call, still page-locked)
            cp, rv, nz = (np.ascontiguousarray(v.indptr, np.int64), np.ascontiguousarray(v.indices, np.int64),
                          np.ascontiguousarray(v.data, np.float64))
            ctx.check(lib.otmb_set_operator(ctx.h, m, len(rv), _ptr(cp), _ptr(rv), _ptr(nz), 0))
    prm = _L.TMParams(float(κH), float(κVML), float(κVdeep), float(ρ) if np.isscalar(ρ) else 0.0, int(bool(upwind)), 0,
                      _L.PATH[path], mask)
    nnz = (C.c_int64 * 5)()
    ctx.check(lib.otmb_transportmatrix_build(ctx.h, C.byref(prm), nnz))
    # every result that was built comes back through ONE pipelined fetch (Int32 indices on the link, widened on the host)
    arrays, fmask = {}, 0
    for m in range(5):
        if m >= 1 and preset[m] is not None:
            continue
        arrays[m] = (ctx.pinned_empty(N + 1, np.int64), ctx.pinned_empty(nnz[m], np.int64), ctx.pinned_empty(nnz[m], np.float64))
        fmask |= 1 << m
    ptrs = [(C.c_void_p * 5)(*[_ptr(arrays[m][q]).value if m in arrays else None for m in range(5)]) for q in range(3)]
    ctx.check(lib.otmb_transportmatrix_fetch_all(ctx.h, fmask, *ptrs))
    out = [_csc(N, *arrays[m]) if m in arrays else preset[m] for m in range(5)]
    return TransportMatrices(*out)


# ---- velocities <-> mass fluxes (src/velocities.jl:10-108, 140-151; src/gridcellgeometry.jl:50-140) ----------
def _haversine_host(p, q, radius=6371000.0):
    """Distances.haversine on two (lon, lat) points in degrees — host side, used only to classify ONE cell."""
    import math
    dlon, dlat = math.radians(q[0] - p[0]), math.radians(q[1] - p[1])
    a = math.sin(dlat / 2) ** 2 + math.cos(math.radians(p[1])) * math.cos(math.radians(q[1])) * math.sin(dlon / 2) ** 2
    return 2 * (radius * math.asin(min(math.sqrt(a), 1.0)))


def _midpointonsphere(A_, B_):
    """src/gridcellgeometry.jl:249-255."""
    if abs(A_[0] - B_[0]) < 180:
        return ((A_[0] + B_[0]) / 2, (A_[1] + B_[1]) / 2)
    return ((A_[0] + B_[0]) / 2 + 180, (A_[1] + B_[1]) / 2)


ArakawaGrid = namedtuple("ArakawaGrid", "kind u_pos v_pos")


def getarakawagrid(u_lon, u_lat, v_lon, v_lat, gridmetrics) -> ArakawaGrid:
    """getarakawagrid, src/gridcellgeometry.jl:50-95: where the velocity points of cell (1,1) sit."""
    lon, lat, lonv, latv = gridmetrics.lon, gridmetrics.lat, gridmetrics.lon_vertices, gridmetrics.lat_vertices
    u_point, v_point = (float(u_lon[0, 0]), float(u_lat[0, 0])), (float(v_lon[0, 0]), float(v_lat[0, 0]))
    cell = {"C": (float(lon[0, 0]), float(lat[0, 0]))}
    for q, name in enumerate(("SW", "SE", "NE", "NW")):
        cell[name] = (float(lonv[q, 0, 0]), float(latv[q, 0, 0]))
    cell["S"] = _midpointonsphere(cell["SW"], cell["SE"])
    cell["N"] = _midpointonsphere(cell["NE"], cell["NW"])
    cell["W"] = _midpointonsphere(cell["SW"], cell["NW"])
    cell["E"] = _midpointonsphere(cell["SE"], cell["NE"])
    u_pos = min(cell, key=lambda k: _haversine_host(cell[k], u_point))      # findmin keeps the first minimum
    v_pos = min(cell, key=lambda k: _haversine_host(cell[k], v_point))
    if u_pos == v_pos == "C":
        return ArakawaGrid("A", u_pos, v_pos)
    if u_pos == v_pos and u_pos in ("NE", "NW", "SE", "SW"):
        return ArakawaGrid("B", u_pos, v_pos)
    if u_pos in ("E", "W") and v_pos in ("N", "S"):
        return ArakawaGrid("C", u_pos, v_pos)
    raise OTMBError(_L.ERR_BADARG, "Unknown Arakawa grid type")


def interpolateontodefaultCgrid(u, u_lon, u_lat, v, v_lon, v_lat, gridmetrics, ctx=None):
    """interpolateontodefaultCgrid, src/gridcellgeometry.jl:103-140.  Returns (u, u_lon, u_lat, v, v_lon, v_lat)."""
    grid = getarakawagrid(np.asarray(u_lon), np.asarray(u_lat), np.asarray(v_lon), np.asarray(v_lat), gridmetrics)
    if grid.kind == "C":
        return u, u_lon, u_lat, v, v_lon, v_lat
    if grid.kind == "A":
        raise OTMBError(_L.ERR_BADARG, "Interpolation not implemented for A-grid type")
    if not (grid.u_pos == grid.v_pos == "NE"):
        raise OTMBError(_L.ERR_BADARG, f"Interpolation not implemented for this B-grid({grid.u_pos},{grid.v_pos}) type")
    fill = u.properties["_FillValue"]
    ctx = _ctx_of(gridmetrics.v3D, ctx=ctx)
    _ensure_grid(ctx, gridmetrics)
    ua, va = _f64(_data(u)), _f64(_data(v))
    u2, v2 = np.empty_like(ua), np.empty_like(va)
    ctx.check(ctx.lib.otmb_bgrid_to_cgrid(ctx.h, _ptr(ua), _ptr(va), float(fill), _ptr(u2), _ptr(v2)))
    lonv, latv = gridmetrics.lon_vertices, gridmetrics.lat_vertices
    mid = np.vectorize(lambda a0, a1, b0, b1: _midpointonsphere((a0, a1), (b0, b1)))
    u2_lon, u2_lat = mid(lonv[2], latv[2], lonv[1], latv[1])        # midpoints of the east side (NE, SE)
    v2_lon, v2_lat = mid(lonv[3], latv[3], lonv[2], latv[2])        # midpoints of the north side (NW, NE)
    return u2, u2_lon, u2_lat, v2, v2_lon, v2_lat


def _vel_common(gridmetrics, ctx):
    ctx = _ctx_of(gridmetrics.v3D, ctx=ctx)
    _ensure_z(ctx, gridmetrics)          # thkcello and edge lengths resident
    return ctx


def velocity2fluxes(u, u_lon, u_lat, v, v_lon, v_lat, gridmetrics, ρ, ctx=None):
    """velocity2fluxes, src/velocities.jl:10-39."""
    ctx = _ctx_of(gridmetrics.v3D, ctx=ctx)
    u, _, _, v, _, _ = interpolateontodefaultCgrid(u, u_lon, u_lat, v, v_lon, v_lat, gridmetrics, ctx=ctx)
    ctx = _vel_common(gridmetrics, ctx)      # last: caller-owned metrics are uploaded afresh by every step that needs them
    ua, va = _f64(_data(u)), _f64(_data(v))
    pi, pj = np.empty_like(ua), np.empty_like(va)
    rho3 = None if np.isscalar(ρ) else _f64(_data(ρ))
    ctx.check(ctx.lib.otmb_velocity2fluxes(ctx.h, _ptr(ua), _ptr(va), _ptr(rho3), float(ρ) if np.isscalar(ρ) else 0.0,
                                           _ptr(pi), _ptr(pj)))
    return pi, pj


def fluxes2velocity(ϕᵢ, ϕⱼ, gridmetrics, ρ, ctx=None):
    """fluxes2velocity, src/velocities.jl:50-74."""
    ctx = _vel_common(gridmetrics, ctx)
    pi, pj = _f64(_data(ϕᵢ)), _f64(_data(ϕⱼ))
    u, v = np.empty_like(pi), np.empty_like(pj)
    rho3 = None if np.isscalar(ρ) else _f64(_data(ρ))
    ctx.check(ctx.lib.otmb_fluxes2velocity(ctx.h, _ptr(pi), _ptr(pj), _ptr(rho3), float(ρ) if np.isscalar(ρ) else 0.0,
                                           _ptr(u), _ptr(v)))
    return u, v


def facefluxesfromvelocities(*, uo, uo_lon, uo_lat, vo, vo_lon, vo_lat, gridmetrics, indices, ρ, ctx=None) -> FaceFluxes:
    """facefluxesfromvelocities, src/velocities.jl:140-151."""
    fill = uo.properties["_FillValue"]
    fv = vo.properties["_FillValue"]
    assert (fill == fv) or (fill != fill and fv != fv), "AssertionError: isequal(FillValue, vo.properties[\"_FillValue\"])"
    umo, vmo = velocity2fluxes(uo, uo_lon, uo_lat, vo, vo_lon, vo_lat, gridmetrics, ρ, ctx=ctx)
    return facefluxes(umo, vmo, gridmetrics, indices, FillValue=fill, ctx=ctx)


# ---- lump_and_spray (src/extratools.jl:38-112) ------------------------------------------------------------
def lump_and_spray(wet3D, vol, T, mask=None, *, di=2, dj=2, dk=1, ctx=None):
    """lump_and_spray(wet3D, vol, T, mask; di, dj, dk) -> (LUMP, SPRAY, vol_c).  Coarsen with LUMP @ x and
    LUMP @ T @ SPRAY.  Only the default mask (lump everywhere) runs on the device: a custom mask makes the
    reference's sweep data dependent and is rejected."""
    wet3D = np.asarray(wet3D, dtype=bool)
    if mask is not None and not np.asarray(mask, dtype=bool).all():
        raise OTMBError(_L.ERR_BADARG, "lump_and_spray: a custom mask is not supported on the device path")
    ctx = ctx or default_context()
    v = ctx.resident.get("v3D")
    if v is None or v.shape != wet3D.shape or not np.array_equal(~np.isnan(v), wet3D):
        v = np.asfortranarray(np.where(wet3D, 1.0, np.nan))          # any array with this wet pattern will do
        _ensure_indices(ctx, v, "bipolar")                             # the topology plays no role here
    N = ctx.resident["N"]
    vol = np.ascontiguousarray(vol, dtype=np.float64)
    Tc = sp.csc_matrix(T)
    if not Tc.has_sorted_indices:                 # a SparseMatrixCSC always is; the library checks the pattern it is given
        Tc = Tc.sorted_indices()
    assert Tc.shape == (N, N) and vol.shape == (N,)
    cp, rv = Tc.indptr.astype(np.int64), Tc.indices.astype(np.int64)
    Nc = C.c_int64()
    ctx.check(ctx.lib.otmb_lump_and_spray_build(ctx.h, int(di), int(dj), int(dk), _ptr(vol), _ptr(cp), _ptr(rv), 0, 0, C.byref(Nc)))
    Nc = Nc.value
    lcp, lrv, lnz = np.empty(N + 1, np.int64), np.empty(N, np.int64), np.empty(N, np.float64)
    scp, srv, snz = np.empty(Nc + 1, np.int64), np.empty(N, np.int64), np.empty(N, np.float64)
    vol_c = np.empty(Nc, np.float64)
    ctx.check(ctx.lib.otmb_lump_and_spray_fetch(ctx.h, _ptr(lcp), _ptr(lrv), _ptr(lnz), _ptr(scp), _ptr(srv), _ptr(snz), _ptr(vol_c)))
    LUMP = sp.csc_matrix((Nc, N), dtype=np.float64)
    LUMP.data, LUMP.indices, LUMP.indptr = lnz, lrv, lcp
    SPRAY = sp.csc_matrix((N, Nc), dtype=np.float64)
    SPRAY.data, SPRAY.indices, SPRAY.indptr = snz, srv, scp
    return LUMP, SPRAY, vol_c


def coarsen(name="T", *, ctx=None):
    """T_c = LUMP * T * SPRAY (test/local_full.jl:161) on the device, with the LUMP / SPRAY of the last `lump_and_spray`
    call and the RESIDENT matrix `name` of the last `transportmatrix` call on this context — the coarse operator of the
    reference's downstream solve, without moving T.  Returns a scipy CSC (N_c x N_c)."""
    ctx = ctx or default_context()
    Nc, nnz = C.c_int64(), C.c_int64()
    ctx.check(ctx.lib.otmb_coarsen_build(ctx.h, _L.MAT[name], C.byref(Nc), C.byref(nnz)))
    cp, rv, nz = np.empty(Nc.value + 1, np.int64), np.empty(nnz.value, np.int64), np.empty(nnz.value, np.float64)
    ctx.check(ctx.lib.otmb_coarsen_fetch(ctx.h, _ptr(cp), _ptr(rv), _ptr(nz)))
    return _csc(Nc.value, cp, rv, nz)


def dump_resident(path, names=MATRICES, *, ctx=None):
    """Write the resident result matrices of the last transportmatrix call to `path` (otmb_transportmatrix_dump: device ->
    pinned staging -> file, no host matrices in between)."""
    ctx = ctx or default_context()
    mask = sum(1 << _L.MAT[n] for n in names)
    ctx.check(ctx.lib.otmb_transportmatrix_dump(ctx.h, mask, str(path).encode()))


def load_dump(path):
    """Read a file written by `dump_resident` / otmb_transportmatrix_dump: dict name -> scipy CSC (0-based)."""
    with open(path, "rb") as f:
        assert f.read(8) == b"OTMBCSC1", "not an OTMBCSC1 file"
        N, base, nmat = np.fromfile(f, "<i8", 3)
        recs = np.fromfile(f, "<i8", 2 * nmat).reshape(nmat, 2)
        out = {}
        for mid, nnz in recs:
            cp = np.fromfile(f, "<i8", N + 1) - base
            rv = np.fromfile(f, "<i8", nnz) - base
            nz = np.fromfile(f, "<f8", nnz)
            out[MATRICES[mid]] = _csc(int(N), cp, rv, nz)
    return out


# ---- products with the resident matrices (no host copy of the matrix needed) ------------------------------
def resident_matvec(name, x, *, transpose=False, ctx=None):
    """y = X @ x (or X.T @ x) with X the resident matrix `name` ("T", "Tadv", "TκH", "TκVML", "TκVdeep") of the last
    transportmatrix call on this context — e.g. the reference's conservation checks (test/online.jl:110-115)."""
    ctx = ctx or default_context()
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    ctx.check(ctx.lib.otmb_spmv(ctx.h, _L.MAT[name], int(bool(transpose)), _ptr(x), _ptr(y)))
    return y


# ---- Redi/GM helpers (experimental and non-exported in the reference) -------------------------
def _ensure_z(ctx, gridmetrics):
    _ensure_grid(ctx, gridmetrics)
    if ctx.resident.get("metrics_src") is None or not _same_frozen(ctx.resident["metrics_src"][1], gridmetrics.thkcello):
        stack = lambda d: np.asfortranarray(np.stack([_f64(d[k]) for k in DIRS], axis=-1))
        zt = np.ascontiguousarray(gridmetrics.zt, dtype=np.float64)
        ctx.check(ctx.lib.otmb_set_gridmetrics(
            ctx.h, _ptr(_f64(gridmetrics.area2D)), _ptr(_f64(gridmetrics.thkcello)), _ptr(zt),
            _ptr(stack(gridmetrics.edge_length_2D)), _ptr(stack(gridmetrics.distance_to_neighbour_2D)),
            _ptr(_f64(gridmetrics.Z3D)), _ptr(_f64(gridmetrics.lon)), _ptr(_f64(gridmetrics.lat))))
        ctx.resident["metrics_src"] = None


def globalverticalfacetriadderivative(χ, gridmetrics, indices, dir, ctx=None):
    """src/triads.jl:134-146; dir is "I" or "J" (Icoord / Jcoord)."""
    ctx = _ctx_of(gridmetrics.v3D, ctx=ctx)
    _ensure_z(ctx, gridmetrics)
    chi = _f64(χ)
    out = np.empty_like(chi)
    ctx.check(ctx.lib.otmb_triad_derivative(ctx.h, _ptr(chi), {"I": 0, "J": 1}[dir], _ptr(out)))
    return out


def globalverticaldyadderivative(χ, gridmetrics, indices, ctx=None):
    """src/dyads.jl:66-78."""
    ctx = _ctx_of(gridmetrics.v3D, ctx=ctx)
    _ensure_z(ctx, gridmetrics)
    chi = _f64(χ)
    out = np.empty_like(chi)
    ctx.check(ctx.lib.otmb_dyad_derivative(ctx.h, _ptr(chi), _ptr(out)))
    return out


def bolus_GM_velocity(ρ, gridmetrics, indices, *, κGM=600, maxslope=0.01, ctx=None):
    """src/RediGM.jl:46-79."""
    ctx = _ctx_of(gridmetrics.v3D, ctx=ctx)
    _ensure_z(ctx, gridmetrics)
    rho = _f64(ρ)
    u, v = np.empty_like(rho), np.empty_like(rho)
    ctx.check(ctx.lib.otmb_bolus_gm_velocity(ctx.h, _ptr(rho), float(κGM), float(maxslope), _ptr(u), _ptr(v)))
    return u, v


def facefluxes_GM(*, umo, vmo, gridmetrics, indices, ρ, κGM=600, maxslope=0.01, ρ_flux=None, return_bolus_fluxes=False, ctx=None):
    """BASELINE configs[2] ("C3") — an EXTENSION, the reference has no such function (parity unpinned): face fluxes of
    the mass transport plus the Gent-McWilliams bolus transport,
        ϕ = facefluxes(umo + ϕᵢ*, vmo + ϕⱼ*),  (ϕᵢ*, ϕⱼ*) = velocity2fluxes(bolus_GM_velocity(ρ; κGM, maxslope)..., ρ_flux)
    (src/RediGM.jl:46-79, src/velocities.jl:10-39, :190-255), chained on the device.  `ρ` is the 3-D density the
    slopes are taken of; `ρ_flux` the density of velocity2fluxes (None: the same 3-D field; or a scalar).  The result
    goes into transportmatrix like any ϕ; T keeps its 7-point pattern.  κGM = 0 gives facefluxesfrommasstransport."""
    fill = umo.properties["_FillValue"]
    fv = vmo.properties["_FillValue"]
    assert (fill == fv) or (fill != fill and fv != fv), "AssertionError: isequal(FillValue, vmo.properties[\"_FillValue\"])"
    ctx = _ctx_of(gridmetrics.v3D, ctx=ctx)
    _ensure_z(ctx, gridmetrics)
    u, v, rho = _f64(_data(umo)), _f64(_data(vmo)), _f64(_data(ρ))
    shape = gridmetrics.v3D.shape
    out = [ctx.pinned_empty(gridmetrics.v3D.size, np.float64, shape, "F") for _ in range(6)]
    gi = np.empty(shape, order="F") if return_bolus_fluxes else None
    gj = np.empty(shape, order="F") if return_bolus_fluxes else None
    ctx.check(ctx.lib.otmb_facefluxes_gm(ctx.h, _ptr(u), _ptr(v), float(fill), _ptr(rho), float(κGM), float(maxslope),
                                         0.0 if ρ_flux is None else float(ρ_flux), int(ρ_flux is None),
                                         *[_ptr(o) for o in out], _ptr(gi), _ptr(gj)))
    _freeze(*out)
    phi = FaceFluxes(*out)
    ctx.resident["phi_arrays"] = tuple(out)
    return (phi, gi, gj) if return_bolus_fluxes else phi


# ---- bare sparse helpers (SparseArrays.sparse / +), exposed for parity tests ------------------
def sparse(I, J, V, n, ctx=None):
    ctx = ctx or default_context()
    I, J = np.ascontiguousarray(I, np.int64), np.ascontiguousarray(J, np.int64)
    V = np.ascontiguousarray(V, np.float64)
    nnz = C.c_int64()
    ctx.check(ctx.lib.otmb_sparse_build(ctx.h, len(I), _ptr(I), _ptr(J), _ptr(V), n, C.byref(nnz)))
    cp, rv, nz = np.empty(n + 1, np.int64), np.empty(nnz.value, np.int64), np.empty(nnz.value, np.float64)
    ctx.check(ctx.lib.otmb_sparse_fetch(ctx.h, _ptr(cp), _ptr(rv), _ptr(nz)))
    return cp, rv, nz          # 1-based, like the SparseMatrixCSC fields


def spadd(A, B, n, ctx=None):
    ctx = ctx or default_context()
    a = [np.ascontiguousarray(A[0], np.int64), np.ascontiguousarray(A[1], np.int64), np.ascontiguousarray(A[2], np.float64)]
    b = [np.ascontiguousarray(B[0], np.int64), np.ascontiguousarray(B[1], np.int64), np.ascontiguousarray(B[2], np.float64)]
    nnz = C.c_int64()
    ctx.check(ctx.lib.otmb_spadd_build(ctx.h, n, *map(_ptr, a), *map(_ptr, b), C.byref(nnz)))
    cp, rv, nz = np.empty(n + 1, np.int64), np.empty(nnz.value, np.int64), np.empty(nnz.value, np.float64)
    ctx.check(ctx.lib.otmb_spadd_fetch(ctx.h, _ptr(cp), _ptr(rv), _ptr(nz)))
    return cp, rv, nz
