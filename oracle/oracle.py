"""ctypes front-end to the CPU oracle (oracle/otmb_oracle.cpp).

TEST INFRASTRUCTURE ONLY — imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never by the product package.  PARITY UNPINNED
(no Julia in the image, no golden vectors in the reference): see the header of
otmb_oracle.cpp and DESIGN.md.

All arrays are Fortran-ordered float64 / int64, shaped like the reference's Julia
arrays, (nx, ny, nz) / (nx, ny) / (4, nx, ny).
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import scipy.sparse as sp

_HERE = Path(__file__).resolve().parent
_LIB = None

TOPO = {"bipolar": 0, "tripolar": 1, "unknown": 2}
DIRS = ("south", "east", "north", "west")            # src/gridcellgeometry.jl:304
FACES = ("east", "west", "north", "south", "top", "bottom")   # src/velocities.jl:245-252
MATS = ("T", "Tadv", "TkH", "TkVML", "TkVdeep")
ERRORS = {
    1: "Tadv contains NaNs.", 2: "TκH contains NaNs.", 3: "TκVML contains NaNs.",
    4: "TκVdeep contains NaNs.", 5: "ρ contains NaNs", 6: "Unknown grid type",
    7: "AssertionError: all umo/vmo values are NaN or FillValue",
    8: "flux from a dry or absent neighbour (reference: MethodError)", 9: "bad argument",
}


class OracleError(RuntimeError):
    def __init__(self, code):
        super().__init__(ERRORS.get(code, f"oracle error {code}"))
        self.code = code


def build(force=False) -> Path:
    so = _HERE / "libotmb_oracle.so"
    src = _HERE / "otmb_oracle.cpp"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["make", "-C", str(_HERE), "-B", "libotmb_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(str(build()))
        dp, ip, i64, dbl = C.c_void_p, C.c_void_p, C.c_int64, C.c_double
        L.orc_makeindices.argtypes = [dp, i64, i64, i64, ip, ip, ip, C.POINTER(i64)]
        L.orc_gridmetrics.argtypes = [dp] * 6 + [i64] * 3 + [C.c_int] + [dp] * 5
        L.orc_facefluxes.argtypes = [dp, dp, dp, i64, i64, i64, C.c_int, dbl] + [dp] * 6
        L.orc_tm_build.argtypes = [dp] * 13 + [i64] * 3 + [C.c_int, dp, dbl, dbl, dbl, dbl, C.c_int, C.c_int]
        L.orc_tm_build.restype = C.c_void_p
        L.orc_tm_build_columns.argtypes = [dp] * 13 + [i64] * 3 + [C.c_int, dp, dbl, dbl, dbl, dbl, C.c_int, C.c_int, i64, i64]
        L.orc_tm_build_columns.restype = C.c_void_p
        for f in (L.orc_tm_status,):
            f.argtypes = [C.c_void_p]
        L.orc_tm_seconds.argtypes = [C.c_void_p]
        L.orc_tm_seconds.restype = dbl
        L.orc_tm_n.argtypes = [C.c_void_p, C.c_int]
        L.orc_tm_n.restype = i64
        L.orc_tm_nnz.argtypes = [C.c_void_p, C.c_int]
        L.orc_tm_nnz.restype = i64
        L.orc_tm_fetch.argtypes = [C.c_void_p, C.c_int, ip, ip, dp]
        L.orc_tm_ntriplets.argtypes = [C.c_void_p, C.c_int]
        L.orc_tm_ntriplets.restype = i64
        L.orc_tm_triplets.argtypes = [C.c_void_p, C.c_int, ip, ip, dp]
        L.orc_tm_free.argtypes = [C.c_void_p]
        L.orc_sparse.argtypes = [ip, ip, dp, i64, i64]
        L.orc_sparse.restype = C.c_void_p
        L.orc_spadd.argtypes = [i64, ip, ip, dp, ip, ip, dp]
        L.orc_spadd.restype = C.c_void_p
        L.orc_haversine.argtypes = [dbl] * 4
        L.orc_haversine.restype = dbl
        L.orc_sind.argtypes = [dbl]
        L.orc_sind.restype = dbl
        L.orc_cosd.argtypes = [dbl]
        L.orc_cosd.restype = dbl
        L.orc_triad.argtypes = [dp] * 5 + [i64] * 3 + [C.c_int, C.c_int, dp]
        L.orc_dyad.argtypes = [dp] * 3 + [i64] * 3 + [C.c_int, dp]
        L.orc_bolus_gm.argtypes = [dp] * 5 + [i64] * 3 + [C.c_int, dbl, dbl, dp, dp]
        _LIB = L
    return _LIB


def _f(a):
    return np.asfortranarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class CSC:
    """Julia SparseMatrixCSC fields, 1-based colptr/rowval."""

    def __init__(self, n, colptr, rowval, nzval):
        self.n, self.colptr, self.rowval, self.nzval = n, colptr, rowval, nzval

    @property
    def nnz(self):
        return len(self.rowval)

    def scipy(self):
        return sp.csc_matrix((self.nzval, self.rowval - 1, self.colptr - 1), shape=(self.n, self.n))


def _fetch(L, h, which):
    n, nnz = L.orc_tm_n(h, which), L.orc_tm_nnz(h, which)
    cp = np.empty(n + 1, np.int64)
    rv = np.empty(nnz, np.int64)
    nz = np.empty(nnz, np.float64)
    L.orc_tm_fetch(h, which, _p(cp), _p(rv), _p(nz))
    return CSC(n, cp, rv, nz)


def makeindices(v3D):
    L = lib()
    v3D = _f(v3D)
    nx, ny, nz = v3D.shape
    M = v3D.size
    chunks = np.zeros((M + 63) // 64, np.uint64)
    Lwet3D = np.zeros(v3D.shape, np.int64, order="F")
    Lwet = np.zeros(M, np.int64)
    N = C.c_int64(0)
    L.orc_makeindices(_p(v3D), nx, ny, nz, _p(chunks), _p(Lwet3D), _p(Lwet), C.byref(N))
    return dict(wet_chunks=chunks, Lwet3D=Lwet3D, Lwet=Lwet[: N.value].copy(), N=N.value,
                wet3D=~np.isnan(v3D))


def gridmetrics(area2D, v3D, lon, lat, lonv, latv, topology):
    L = lib()
    area2D, v3D, lon, lat, lonv, latv = map(_f, (area2D, v3D, lon, lat, lonv, latv))
    nx, ny, nz = v3D.shape
    P = nx * ny
    thk = np.empty_like(v3D)
    Z3D = np.empty_like(v3D)
    edge = np.empty((nx, ny, 4), order="F")
    dedge = np.empty((nx, ny, 4), order="F")
    dnbr = np.empty((nx, ny, 4), order="F")
    st = L.orc_gridmetrics(_p(area2D), _p(v3D), _p(lon), _p(lat), _p(lonv), _p(latv), nx, ny, nz,
                           TOPO[topology], _p(thk), _p(Z3D), _p(edge), _p(dedge), _p(dnbr))
    if st:
        raise OracleError(st)
    return dict(thkcello=thk, Z3D=Z3D, edge=edge, dedge=dedge, dnbr=dnbr)


def facefluxes(umo, vmo, v3D, topology, fill):
    L = lib()
    umo, vmo = _f(umo).copy(order="F"), _f(vmo).copy(order="F")
    v3D = _f(v3D)
    nx, ny, nz = v3D.shape
    out = [np.empty_like(v3D) for _ in range(6)]
    st = L.orc_facefluxes(_p(umo), _p(vmo), _p(v3D), nx, ny, nz, TOPO[topology], float(fill), *map(_p, out))
    if st:
        raise OracleError(st)
    return dict(zip(FACES, out))


def transportmatrix(phi, mlotst, v3D, thk, area2D, zt, edge, dnbr, topology, rho, kH=500.0, kVML=0.1,
                    kVdeep=1.0e-5, upwind=True, keep_triplets=False):
    """Returns dict T/Tadv/TkH/TkVML/TkVdeep -> CSC, plus 'seconds' (the oracle's own wall
    time for emit + sparse + adds) and optionally 'triplets'."""
    L = lib()
    ph = [_f(phi[k]) for k in FACES]
    mlotst, v3D, thk, area2D, edge, dnbr = map(_f, (mlotst, v3D, thk, area2D, edge, dnbr))
    zt = np.ascontiguousarray(zt, dtype=np.float64)
    nx, ny, nz = v3D.shape
    rho3 = None if np.isscalar(rho) else _f(rho)
    rs = float(rho) if np.isscalar(rho) else 0.0
    h = L.orc_tm_build(*map(_p, ph), _p(mlotst), _p(v3D), _p(thk), _p(area2D), _p(zt), _p(edge), _p(dnbr),
                       nx, ny, nz, TOPO[topology], _p(rho3), rs, float(kH), float(kVML), float(kVdeep),
                       int(bool(upwind)), int(keep_triplets))
    try:
        st = L.orc_tm_status(h)
        if st:
            raise OracleError(st)
        out = {name: _fetch(L, h, w) for w, name in enumerate(MATS)}
        out["seconds"] = L.orc_tm_seconds(h)
        if keep_triplets:
            tr = {}
            for op, name in enumerate(MATS[1:]):
                n = L.orc_tm_ntriplets(h, op)
                I, J, V = np.empty(n, np.int64), np.empty(n, np.int64), np.empty(n, np.float64)
                L.orc_tm_triplets(h, op, _p(I), _p(J), _p(V))
                tr[name] = (I, J, V)
            out["triplets"] = tr
        return out
    finally:
        L.orc_tm_free(h)


def transportmatrix_columns(phi, mlotst, v3D, thk, area2D, zt, edge, dnbr, topology, rho, col_lo, col_hi, kH=500.0, kVML=0.1,
                            kVdeep=1.0e-5, upwind=True):
    """The columns [col_lo, col_hi) (0-based) of the five matrices of `transportmatrix`, built by the same emitters,
    `sparse` and `+` restricted to the triplets of those columns (orc_tm_build_columns): for grids whose full matrices
    are too large to hold beside the inputs.  Returns dict name -> CSC with n = col_hi - col_lo columns, colptr local
    (1-based), rowval global (1-based)."""
    L = lib()
    ph = [_f(phi[k]) for k in FACES]
    mlotst, v3D, thk, area2D, edge, dnbr = map(_f, (mlotst, v3D, thk, area2D, edge, dnbr))
    zt = np.ascontiguousarray(zt, dtype=np.float64)
    nx, ny, nz = v3D.shape
    rho3 = None if np.isscalar(rho) else _f(rho)
    rs = float(rho) if np.isscalar(rho) else 0.0
    h = L.orc_tm_build_columns(*map(_p, ph), _p(mlotst), _p(v3D), _p(thk), _p(area2D), _p(zt), _p(edge), _p(dnbr),
                               nx, ny, nz, TOPO[topology], _p(rho3), rs, float(kH), float(kVML), float(kVdeep),
                               int(bool(upwind)), 0, int(col_lo), int(col_hi))
    try:
        st = L.orc_tm_status(h)
        if st:
            raise OracleError(st)
        return {name: _fetch(L, h, w) for w, name in enumerate(MATS)}
    finally:
        L.orc_tm_free(h)


def sparse(I, J, V, n):
    L = lib()
    I = np.ascontiguousarray(I, np.int64)
    J = np.ascontiguousarray(J, np.int64)
    V = np.ascontiguousarray(V, np.float64)
    h = L.orc_sparse(_p(I), _p(J), _p(V), len(I), n)
    try:
        return _fetch(L, h, 0)
    finally:
        L.orc_tm_free(h)


def spadd(A: CSC, B: CSC):
    L = lib()
    h = L.orc_spadd(A.n, _p(A.colptr), _p(A.rowval), _p(A.nzval), _p(B.colptr), _p(B.rowval), _p(B.nzval))
    try:
        return _fetch(L, h, 0)
    finally:
        L.orc_tm_free(h)


def haversine(p, q):
    return lib().orc_haversine(float(p[0]), float(p[1]), float(q[0]), float(q[1]))


def sind(x):
    return lib().orc_sind(float(x))


def cosd(x):
    return lib().orc_cosd(float(x))


def triad(chi, lon, lat, Z3D, v3D, topology, direction):
    L = lib()
    chi, lon, lat, Z3D, v3D = map(_f, (chi, lon, lat, Z3D, v3D))
    nx, ny, nz = v3D.shape
    out = np.empty_like(v3D)
    st = L.orc_triad(_p(chi), _p(lon), _p(lat), _p(Z3D), _p(v3D), nx, ny, nz, TOPO[topology],
                     {"I": 0, "J": 1}[direction], _p(out))
    if st:
        raise OracleError(st)
    return out


def dyad(chi, Z3D, v3D, topology):
    L = lib()
    chi, Z3D, v3D = map(_f, (chi, Z3D, v3D))
    nx, ny, nz = v3D.shape
    out = np.empty_like(v3D)
    st = L.orc_dyad(_p(chi), _p(Z3D), _p(v3D), nx, ny, nz, TOPO[topology], _p(out))
    if st:
        raise OracleError(st)
    return out


def bolus_gm(rho, lon, lat, Z3D, v3D, topology, kGM=600.0, maxslope=0.01):
    L = lib()
    rho, lon, lat, Z3D, v3D = map(_f, (rho, lon, lat, Z3D, v3D))
    nx, ny, nz = v3D.shape
    u, v = np.empty_like(v3D), np.empty_like(v3D)
    st = L.orc_bolus_gm(_p(rho), _p(lon), _p(lat), _p(Z3D), _p(v3D), nx, ny, nz, TOPO[topology],
                        float(kGM), float(maxslope), _p(u), _p(v))
    if st:
        raise OracleError(st)
    return u, v


# ---- host-side pieces of the reference that stay on the host -------
def clean_missing(a, fills=()):
    """makegridmetrics :269-280: missing/nothing/0/_FillValue -> NaN.  `replace` matches with
    isequal, so -0.0 is NOT replaced by the `0 => NaN` pair."""
    a = np.array(a, dtype=np.float64, order="F")
    bad = (a == 0.0) & ~np.signbit(a)
    for fv in fills:
        bad |= a == np.float64(fv)
    a[bad] = np.nan
    return a


def vertexpermutation(lonv, latv):
    """src/gridcellgeometry.jl:158-178 (0-based result)."""
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


def getgridtopology(lonv, latv):
    """src/gridtopology.jl:33-53."""
    NPlon = lonv[2:4, :, -1]
    NPlat = latv[2:4, :, -1]
    if np.all(NPlat == 90):
        return "bipolar"
    rot = lambda a: a[::-1, ::-1]
    d = np.mod(NPlon - rot(NPlon) + 180, 360) - 180
    lon_ok = np.linalg.norm(d) <= np.spacing(180.0)       # isapprox(Δ, zeros; atol=eps(180.0)), rtol = 0
    a, b = NPlat, rot(NPlat)
    lat_ok = np.linalg.norm(a - b) <= np.sqrt(np.finfo(float).eps) * max(np.linalg.norm(a), np.linalg.norm(b))
    return "tripolar" if (lon_ok and lat_ok) else "unknown"
