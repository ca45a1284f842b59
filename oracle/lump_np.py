"""Python restatement of lump_and_spray (/root/reference/src/extratools.jl:38-112), SURVEY.md §8f rank 3.

TEST INFRASTRUCTURE ONLY (same rules as oracle.py); PARITY UNPINNED like the rest of the oracle: the
reference only checks that the function runs (test/online.jl:129) and that a coarsened ideal-age solve is
plausible (test/local_full.jl:151-188).  The third-party piece is Graphs.jl `connected_components`
(Project.toml compat "1"): components are labelled by their smallest vertex and returned in order of that
label, each component's vertices ascending — restated here.  Loops in pure Python: small grids only.
Everything is 0-based here; the comments give the reference's 1-based line.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def _connected_components(nv, edges):
    """Graphs.connected_components on an undirected simple graph: list of vertex lists, ordered by the
    smallest vertex of each component, vertices ascending."""
    parent = list(range(nv))

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    for a, b in edges:
        ra, rb = find(a), find(b)
        if ra != rb:
            parent[max(ra, rb)] = min(ra, rb)
    comps = {}
    for v in range(nv):
        comps.setdefault(find(v), []).append(v)
    return [comps[k] for k in sorted(comps)]


def lump_and_spray(wet3D, vol, T, mask=None, di=2, dj=2, dk=1):
    """Returns (LUMP csc (Nc x N), SPRAY csc (N x Nc), vol_c).  T: scipy sparse (its stored pattern is the
    connectivity, `findnz(T)` :47)."""
    wet3D = np.asarray(wet3D, dtype=bool)
    nx, ny, nz = wet3D.shape
    mask = np.ones(wet3D.shape, bool) if mask is None else np.asarray(mask, bool)
    ex, ey, ez = nx + di - 1, ny + dj - 1, nz + dk - 1                      # extended grid, :43-45
    wet_ext = np.zeros((ex, ey, ez), bool)
    wet_ext[:nx, :ny, :nz] = wet3D
    lump = np.zeros((ex, ey, ez), np.int64)                                 # LUMPidx, 0 = unassigned
    lin_ext = lambda i, j, k: i + ex * (j + ey * k)
    # wet rank of every wet cell (column-major), and T's stored pattern as a set of (row, col) wet ranks
    rank = -np.ones(wet3D.shape, np.int64)
    order = np.argwhere(wet3D.transpose(2, 1, 0))[:, ::-1]                  # (i, j, k) in column-major order
    rank[tuple(order.T)] = np.arange(len(order))
    Tc = sp.csc_matrix(T)
    pattern = set(zip(Tc.indices.tolist(), np.repeat(np.arange(Tc.shape[1]), np.diff(Tc.indptr)).tolist()))
    c = 2                                                                     # 1 is reserved for dry cells, :53
    for k in range(nz):                                                       # eachindex(C): column-major, :55
        for j in range(ny):
            for i in range(nx):
                if lump[i, j, k] > 0 and mask[i, j, k]:
                    continue
                if mask[i, j, k]:
                    cells = [(i + a, j + b, k + d) for d in range(dk) for b in range(dj) for a in range(di)]   # vec(L[C𝑖 .+ neighbours])
                    wetcells = []
                    for cell in cells:
                        if wet_ext[cell]:
                            wetcells.append(cell)
                        else:
                            lump[cell] = 1                                    # :64-66
                    edges = []
                    for a in range(len(wetcells)):
                        for b in range(len(wetcells)):
                            if a != b and (int(rank[wetcells[a]]), int(rank[wetcells[b]])) in pattern:
                                edges.append((a, b))
                    # SimpleGraph(adjacency) needs a symmetric matrix (Graphs.jl throws otherwise)
                    es = set(edges)
                    if any((b, a) not in es for a, b in es):
                        raise ValueError("connectivity of a lumping box is not symmetric (SimpleGraph would throw)")
                    for comp in _connected_components(len(wetcells), edges):
                        for v in comp:
                            lump[wetcells[v]] = c                             # :72-75
                        c += 1
                else:
                    lump[i, j, k] = c                                         # :77-79
                    c += 1
    ids = lump[:nx, :ny, :nz].ravel(order="F")                                # LUMPidx[C][:], :84
    M = nx * ny * nz
    LUMP = sp.csc_matrix((np.ones(M), (ids - 1, np.arange(M))), shape=(int(ids.max()), M))
    wet = wet3D.ravel(order="F")
    wet_c = (LUMP @ wet.astype(np.float64)) > 0                               # :87
    LUMP = sp.csc_matrix(LUMP[wet_c][:, wet])                                 # :90
    vol = np.asarray(vol, np.float64)
    # vol_c = LUMP * vol: y[r] += 1 * vol[j] for j ascending (CSC SpMV), :95
    vol_c = np.zeros(LUMP.shape[0])
    rows = LUMP.indices
    for jcol in range(LUMP.shape[1]):
        for p in range(LUMP.indptr[jcol], LUMP.indptr[jcol + 1]):
            vol_c[rows[p]] += 1.0 * vol[jcol]
    # Diagonal(1 ./ vol_c) * LUMP * Diagonal(vol): ((1/vol_c[r]) * 1) * vol[j], :96
    L2 = LUMP.copy()
    L2.data = ((1.0 / vol_c)[rows] * 1.0) * np.repeat(vol, np.diff(LUMP.indptr))
    SPRAY = sp.csc_matrix(L2.T)                                               # copy(LUMP'), :100-101
    SPRAY.sort_indices()
    SPRAY.data[:] = 1.0
    return L2, SPRAY, vol_c


def spmatmul(A, B):
    """C = A * B as SparseArrays computes it (Gustavson, stdlib `spmatmul`; restated from its published algorithm, parity
    unpinned): for each column j of B, for each stored B[k, j] in ascending k, for each stored A[i, k] in ascending i:
    the first product for row i is stored, later ones are added, in that order; rows of the result column ascending;
    no zero dropping.  A, B: scipy CSC with sorted indices.  Pure-Python loops: small cases only."""
    import scipy.sparse as sp
    A, B = sp.csc_matrix(A), sp.csc_matrix(B)
    indptr, indices, data = [0], [], []
    for j in range(B.shape[1]):
        acc = {}
        for t in range(B.indptr[j], B.indptr[j + 1]):
            k, b = B.indices[t], B.data[t]
            for s in range(A.indptr[k], A.indptr[k + 1]):
                i, p = A.indices[s], A.data[s] * b
                acc[i] = p if i not in acc else acc[i] + p
        for i in sorted(acc):
            indices.append(i)
            data.append(acc[i])
        indptr.append(len(indices))
    C = sp.csc_matrix((A.shape[0], B.shape[1]), dtype=np.float64)
    C.data, C.indices, C.indptr = np.array(data, np.float64), np.array(indices, np.int64), np.array(indptr, np.int64)
    return C


def coarsen(LUMP, T, SPRAY):
    """T_c = LUMP * T * SPRAY, left to right like Julia parses it (/root/reference/test/local_full.jl:161)."""
    return spmatmul(spmatmul(LUMP, T), SPRAY)
