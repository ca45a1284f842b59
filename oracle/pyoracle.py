"""Independent pure-Python restatement of the reference's emitters, `sparse` and sparse `+`.

TEST INFRASTRUCTURE ONLY (same rules as oracle.py).  Written with different data
structures from otmb_oracle.cpp on purpose (dicts keyed by (row, col) that accumulate in
insertion order instead of the stdlib's counting sorts) so that the two restatements
check each other.  Pure-Python loops: small grids only.  PARITY UNPINNED, see DESIGN.md.

File:line citations are relative to /root/reference.
"""
from __future__ import annotations

import math

import numpy as np


def _nbr(kind, nx, ny, nz):
    """src/gridtopology.jl:57-68, 94.  1-based (i, j, k) tuples or None."""
    def ip1(c): return (c[0] + 1, c[1], c[2]) if c[0] < nx else (1, c[1], c[2])
    def im1(c): return (c[0] - 1, c[1], c[2]) if c[0] > 1 else (nx, c[1], c[2])
    def jm1(c): return (c[0], c[1] - 1, c[2]) if c[1] > 1 else None
    def jp1(c):
        if c[1] < ny:
            return (c[0], c[1] + 1, c[2])
        return (nx - c[0] + 1, ny, c[2]) if kind == "tripolar" else None
    def kp1(c): return (c[0], c[1], c[2] + 1) if c[2] < nz else None
    def km1(c): return (c[0], c[1], c[2] - 1) if c[2] > 1 else None
    return ip1, im1, jp1, jm1, kp1, km1


def jl_min(x, y):
    if math.isnan(x) or math.isnan(y):
        return x + y
    return x if math.copysign(1.0, x - y) < 0 else y


def jl_max(x, y):
    if math.isnan(x) or math.isnan(y):
        return x + y
    return y if math.copysign(1.0, x - y) < 0 else x


def _colptr(keys, n):
    counts = np.bincount(np.array([k[0] for k in keys], dtype=np.int64), minlength=n + 1)[1:]
    return np.concatenate([[1], 1 + np.cumsum(counts)]).astype(np.int64)


def sparse_py(I, J, V, n):
    """sparse(I,J,V,n,n): duplicates summed left to right in input order, rows ascending in
    each column, explicit zeros kept.  Returns 1-based (colptr, rowval, nzval)."""
    acc = {}
    for i, j, v in zip(I, J, V):
        key = (j, i)
        acc[key] = acc[key] + v if key in acc else v
    keys = sorted(acc)
    colptr = _colptr(keys, n)
    rowval = np.array([k[1] for k in keys], np.int64)
    nzval = np.array([acc[k] for k in keys], np.float64)
    return colptr, rowval, nzval


def spadd_py(A, B, n):
    """A + B on 1-based CSC triples; results equal to zero are not stored."""
    def todict(M):
        cp, rv, nz = M
        d = {}
        for j in range(1, n + 1):
            for p in range(cp[j - 1] - 1, cp[j] - 1):
                d[(j, rv[p])] = nz[p]
        return d
    a, b = todict(A), todict(B)
    out = {}
    for key in sorted(set(a) | set(b)):
        x = a.get(key, 0.0) + b.get(key, 0.0)
        if x != 0.0:
            out[key] = x
    keys = sorted(out)
    return _colptr(keys, n), np.array([k[1] for k in keys], np.int64), np.array([out[k] for k in keys], np.float64)


def transportmatrix_py(phi, mlotst, v3D, thk, area2D, zt, edge, dnbr, topology, rho, kH, kVML, kVdeep, upwind=True):
    """transportmatrix, src/matrixbuilding.jl:128-150.  edge/dnbr: (nx, ny, 4) in the order
    south, east, north, west.  Returns dict name -> (colptr, rowval, nzval), 1-based."""
    np.seterr(divide="ignore", invalid="ignore")
    nx, ny, nz = v3D.shape
    ip1, im1, jp1, jm1, kp1, km1 = _nbr(topology, nx, ny, nz)
    S, E, Nn, W = 0, 1, 2, 3
    wet = ~np.isnan(v3D)
    Lwet3D = {}
    cells = []
    for k in range(1, nz + 1):
        for j in range(1, ny + 1):
            for i in range(1, nx + 1):
                if wet[i - 1, j - 1, k - 1]:
                    cells.append((i, j, k))
                    Lwet3D[(i, j, k)] = len(cells)
    N = len(cells)
    at = lambda a, c: float(a[c[0] - 1, c[1] - 1, c[2] - 1])
    rho_at = (lambda c: float(rho)) if np.isscalar(rho) else (lambda c: at(rho, c))

    # ---- Tadv :221-299
    I, J, V = [], [], []
    def push_adv(wi, cj, p, ci):
        wj = Lwet3D[cj]                       # KeyError <-> the reference's MethodError
        r = (rho_at(ci) + rho_at(cj)) / 2
        mi, mj = r * at(v3D, ci), r * at(v3D, cj)
        I.append(wi); J.append(wj); V.append(float(np.float64(-p) / np.float64(mi)))
        I.append(wj); J.append(wj); V.append(float(np.float64(p) / np.float64(mj)))
    for wi, c in enumerate(cells, start=1):
        def flux(name, up):
            x = at(phi[name], c)
            return up(x, 0.0) if upwind else x / 2
        f = flux("west", jl_max)
        if f > 0 or f < 0: push_adv(wi, im1(c), f, c)
        f = flux("east", jl_min)
        if f > 0 or f < 0: push_adv(wi, ip1(c), -f, c)
        f = flux("south", jl_max)
        if f > 0 or f < 0: push_adv(wi, jm1(c), f, c)
        f = flux("north", jl_min)
        if f > 0 or f < 0: push_adv(wi, jp1(c), -f, c)
        f = flux("bottom", jl_max)
        if f > 0 or f < 0: push_adv(wi, kp1(c), f, c)
        f = flux("top", jl_min)
        if c[2] > 1 and (f > 0 or f < 0): push_adv(wi, km1(c), -f, c)
    Tadv = sparse_py(I, J, V, N)

    # ---- TκH :337-418
    I, J, V = [], [], []
    def mix(wi, wj, kappa, a, d, Vol):
        t = float(np.float64(kappa * a) / np.float64(d * Vol))      # IEEE division (inf/nan, no exception)
        I.append(wi); J.append(wi); V.append(t)
        I.append(wi); J.append(wj); V.append(-t)
    for wi, c in enumerate(cells, start=1):
        Vol = at(v3D, c)
        for cj, d_, opp in ((im1(c), W, E), (ip1(c), E, W), (jm1(c), S, Nn),
                            (jp1(c), Nn, Nn if c[1] == ny else S)):
            if cj is None or cj not in Lwet3D:
                continue
            aij = at(thk, c) * float(edge[c[0] - 1, c[1] - 1, d_])
            aji = at(thk, cj) * float(edge[cj[0] - 1, cj[1] - 1, opp])
            mix(wi, Lwet3D[cj], kH, jl_min(aij, aji), float(dnbr[c[0] - 1, c[1] - 1, d_]), Vol)
    TkH = sparse_py(I, J, V, N)

    # ---- TκVML / TκVdeep :438-479, mask :85
    def vdiff(kV, Omega):
        I.clear(); J.clear(); V.clear()
        for wi, c in enumerate(cells, start=1):
            if not Omega[wi - 1]:
                continue
            Vol, a = at(v3D, c), float(area2D[c[0] - 1, c[1] - 1])
            for cj in (kp1(c), km1(c)):
                if cj is None or cj not in Lwet3D or not Omega[Lwet3D[cj] - 1]:
                    continue
                mix(wi, Lwet3D[cj], kV, a, abs(float(zt[c[2] - 1]) - float(zt[cj[2] - 1])), Vol)
        return sparse_py(I, J, V, N)
    OmegaML = [bool(float(zt[c[2] - 1]) < float(mlotst[c[0] - 1, c[1] - 1])) for c in cells]
    TkVML = vdiff(kVML, OmegaML)
    TkVdeep = vdiff(kVdeep, [True] * N)

    T = spadd_py(spadd_py(spadd_py(Tadv, TkH, N), TkVML, N), TkVdeep, N)
    return dict(T=T, Tadv=Tadv, TkH=TkH, TkVML=TkVML, TkVdeep=TkVdeep, N=N)
