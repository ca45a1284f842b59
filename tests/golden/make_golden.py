#!/usr/bin/env python
"""Generates the fixtures under tests/golden/.

    python tests/golden/make_golden.py            # rewrite known_answers.json and the *.npz fixtures

Two kinds of fixture, kept apart on purpose:

1. known_answers.json — HAND-DERIVED expected results (no oracle involved in making them).  The
   reference holds no golden vectors for this path (its only CI test downloads CMIP6 data,
   /root/reference/test/online.jl:19-65) and Julia is not available, so these small cases are
   what pins the oracle: every expected matrix below is written out from the reference's
   formulas (file:line quoted next to each case) by hand, with unit metrics so that the
   arithmetic is exact.

2. case_*.npz — frozen inputs + the oracle's outputs on small seeded oceans (this script imports
   oracle/, which tests/ may do).  They do not pin the oracle against the reference (nothing
   here can); they freeze today's oracle so that (a) the CUDA path is checked against committed
   bytes, not only against a library built in the same run, (b) any later change of the oracle
   is visible as a diff, and (c) a future Julia run can be compared with the same inputs.
"""
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


# ------------------------------------------------------------------------------------------
# 1. hand-derived known answers
# ------------------------------------------------------------------------------------------
def known_answers():
    R = 6371000.0
    ka = {"_comment": "hand-derived; see tests/golden/make_golden.py for the derivations"}

    # Distances.haversine (lon, lat in degrees, R = 6371000): great-circle identities
    ka["haversine"] = [
        {"p": [0, 0], "q": [0, 90], "d": R * np.pi / 2},        # equator -> pole: quarter circle
        {"p": [0, 0], "q": [1, 0], "d": R * np.pi / 180},       # one degree of longitude on the equator
        {"p": [0, 0], "q": [180, 0], "d": R * np.pi},           # antipodes
        {"p": [10, 20], "q": [10, 20], "d": 0.0},
        {"p": [80, 90], "q": [170, 90], "d": 0.0},              # both at the pole: cosd(90) == 0 exactly
        {"p": [0, 60], "q": [180, 60], "d": R * np.pi / 3},     # over the pole: 30 + 30 degrees
    ]

    # ---- TκVdeep on a 4x2x2 all-wet box, unit metrics ---------------------------------------
    # src/matrixbuilding.jl:450-477: per wet cell, Bottom (k+1) then Top (k-1):
    #   t = κ*area/(|zt[k]-zt[k']|*V);  emit (𝑖,𝑖,+t), (𝑖,𝑗,-t).
    # area = V = 1, zt = [0.5, 1.5] -> d = 1, t = κ.  Cell 𝑖 (level 1) pairs with 𝑖+8 (level 2):
    #   T = κ [[ I, -I], [-I, I]]   (8x8 blocks)
    kap = 0.25
    I8 = np.eye(8)
    ka["kvdeep_4x2x2"] = {"shape": [4, 2, 2], "topology": "tripolar", "kVdeep": kap,
                          "dense": (kap * np.block([[I8, -I8], [-I8, I8]])).tolist()}

    # ---- TκH on the same box (tripolar: periodic in i, fold on the top row) ------------------
    # src/matrixbuilding.jl:348-415, :426-435: per wet cell W,E,S,N: a = min(thk*edge, thk'*edge') = 1,
    #   d = 1, V = 1 -> t = κH; emit (𝑖,𝑖,+t), (𝑖,𝑗,-t).  Row j=1 cells (i=1..4 -> index i): W=i-1, E=i+1
    #   (periodic), N=(i,2) -> index 4+i, no S.  Row j=2 cells (index 4+i): W, E periodic in the row,
    #   S=(i,1), N = fold (nx-i+1, 2) = (5-i, 2): 1<->4 and 2<->3 (src/gridtopology.jl:94-95).
    #   (1,2)<->(4,2) are ALSO periodic W/E neighbours and (2,2)<->(3,2) are E/W neighbours, so those
    #   pairs get two triplets each, which sparse() sums: -κH + -κH = -2κH.
    #   M[row 𝑖, col 𝑗] below; diagonal = (number of emitted slots)*κH.
    kH = 2.0
    A = np.zeros((8, 8))
    def idx(i, j):          # 1-based (i,j) -> 0-based index in a level
        return (i - 1) + 4 * (j - 1)
    for i in range(1, 5):
        w, e = (i - 2) % 4 + 1, i % 4 + 1
        # row j = 1: W, E, N
        for nb in (idx(w, 1), idx(e, 1), idx(i, 2)):
            A[idx(i, 1), idx(i, 1)] += kH
            A[idx(i, 1), nb] -= kH
        # row j = 2: W, E, S, N(fold)
        for nb in (idx(w, 2), idx(e, 2), idx(i, 1), idx(5 - i, 2)):
            A[idx(i, 2), idx(i, 2)] += kH
            A[idx(i, 2), nb] -= kH
    # spot checks of the hand derivation itself
    assert A[idx(1, 2), idx(4, 2)] == -2 * kH and A[idx(2, 2), idx(3, 2)] == -2 * kH
    assert A[idx(1, 1), idx(1, 1)] == 3 * kH and A[idx(1, 2), idx(1, 2)] == 4 * kH
    assert (A.sum(axis=1) == 0).all()
    Z = np.zeros((8, 8))
    ka["kh_4x2x2"] = {"shape": [4, 2, 2], "topology": "tripolar", "kH": kH,
                      "dense": np.block([[A, Z], [Z, A]]).tolist()}

    # ---- Tadv on a periodic ring of 4 cells (4x1x1, bipolar) with a uniform eastward flux ------
    # src/matrixbuilding.jl:237-297 + :193-204, upwind: only the West slot is active
    # (max(ϕwest,0) = F > 0; min(ϕeast,0) = 0): cell 𝑖 with west neighbour 𝑗 = 𝑖-1 emits
    #   (𝑖, 𝑗, -F/m𝑖), (𝑗, 𝑗, +F/m𝑗),  m = ((ρ+ρ)/2)*V = 1*2 = 2.
    # => T = (F/2) (I - S),  S[𝑖, 𝑖-1] = 1 (periodic).   Centred (upwind=false): ϕ/2 through both faces:
    #   West slot: p = F/2 -> (𝑖, 𝑖-1, -F/4), (𝑖-1, 𝑖-1, +F/4); East slot: f = F/2, pushed ϕ = -f ->
    #   (𝑖, 𝑖+1, +F/4), (𝑖+1, 𝑖+1, -F/4).  Diagonal: +F/4 - F/4 = 0 (kept by sparse, dropped from T by +).
    F = 8.0
    S = np.roll(np.eye(4), 1, axis=0)        # S[i, i-1] = 1
    ka["adv_ring4"] = {"shape": [4, 1, 1], "topology": "bipolar", "F": F, "rho": 1.0, "V": 2.0,
                       "upwind_dense": ((F / 2) * (np.eye(4) - S)).tolist(),
                       "centred_dense": ((F / 4) * (S.T - S)).tolist()}

    # ---- SparseArrays.sparse / + semantics (stdlib; restated in SURVEY.md Appendix A.4/A.5) ------
    ka["sparse"] = {"I": [2, 1, 2, 3, 2, 1], "J": [1, 1, 1, 2, 1, 3], "V": [1e16, 1.0, -1e16, 0.0, 1.0, 5.0], "n": 3,
                    "colptr": [1, 3, 4, 5], "rowval": [1, 2, 3, 1], "nzval": [1.0, 1.0, 0.0, 5.0]}
    ka["spadd"] = {"A": {"I": [1, 2, 3], "J": [1, 1, 2], "V": [1.0, 2.0, 0.0]},
                   "B": {"I": [1, 2, 3], "J": [1, 1, 3], "V": [-1.0, 5.0, 4.0]}, "n": 3,
                   "colptr": [1, 2, 2, 3], "rowval": [2, 3], "nzval": [7.0, 4.0]}
    return ka


# ------------------------------------------------------------------------------------------
# 2. frozen oracle outputs on small seeded oceans
# ------------------------------------------------------------------------------------------
CASES = {
    # name: (nx, ny, nz, topology, seed, make_ocean kwargs, transportmatrix kwargs)
    "tripolar_12x10x6": (12, 10, 6, "tripolar", 0, {"land_frac": 0.25}, {}),
    "bipolar_10x8x4": (10, 8, 4, "bipolar", 3, {"land_frac": 0.25}, {}),
    "tripolar_13x9x5_oddnx": (13, 9, 5, "tripolar", 1, {"land_frac": 0.25}, {}),
    "tripolar_2x4x3_nx2": (2, 4, 3, "tripolar", 2, {"land_frac": 0.25}, {}),
    "tripolar_16x12x5_centred_rho3d": (16, 12, 5, "tripolar", 9, {"land_frac": 0.2}, {"upwind": False, "rho3d": True}),
    "bipolar_24x12x6_dirty_f32": (24, 12, 6, "bipolar", 7, {"land_frac": 0.25, "dirty": True, "float32_roundtrip": True}, {}),
}


def frozen_case(name):
    import otmb_b200  # noqa: F401  (package import only for the synthetic generator)
    from otmb_b200 import synthetic
    from oracle import oracle as O
    nx, ny, nz, topo, seed, okw, tkw = CASES[name]
    oc = synthetic.make_ocean(nx, ny, nz, topo, seed=seed, **okw)
    v3D, area = O.clean_missing(oc.volcello), O.clean_missing(oc.areacello)
    ix = O.makeindices(v3D)
    gm = O.gridmetrics(area, v3D, oc.lon, oc.lat, oc.lon_vertices, oc.lat_vertices, oc.topology)
    phi = O.facefluxes(oc.umo, oc.vmo, v3D, oc.topology, oc.fill)
    rho = oc.rho3d if tkw.get("rho3d") else 1035.0
    tm = O.transportmatrix(phi, oc.mlotst, v3D, gm["thkcello"], area, oc.lev, gm["edge"], gm["dnbr"], oc.topology, rho,
                           upwind=tkw.get("upwind", True))
    out = dict(
        topology=np.array(oc.topology), fill=np.float64(oc.fill), upwind=np.bool_(tkw.get("upwind", True)),
        rho_scalar=np.float64(1035.0), use_rho3d=np.bool_(bool(tkw.get("rho3d"))),
        # inputs
        volcello=oc.volcello, areacello=oc.areacello, lon=oc.lon, lat=oc.lat, lev=oc.lev, lon_vertices=oc.lon_vertices,
        lat_vertices=oc.lat_vertices, umo=oc.umo, vmo=oc.vmo, mlotst=oc.mlotst, rho3d=oc.rho3d,
        # oracle outputs
        N=np.int64(ix["N"]), Lwet=ix["Lwet"], wet_chunks=ix["wet_chunks"],
        thkcello=gm["thkcello"], Z3D=gm["Z3D"], edge=gm["edge"], dedge=gm["dedge"], dnbr=gm["dnbr"],
    )
    for k, a in phi.items():
        out["phi_" + k] = a
    for m in O.MATS:
        out[m + "_colptr"], out[m + "_rowval"], out[m + "_nzval"] = tm[m].colptr, tm[m].rowval, tm[m].nzval
    return out


def main():
    (HERE / "known_answers.json").write_text(json.dumps(known_answers(), indent=1))
    for name in CASES:
        np.savez_compressed(HERE / f"case_{name}.npz", **frozen_case(name))
        print("wrote", name)


if __name__ == "__main__":
    main()
