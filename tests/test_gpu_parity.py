"""GPU parity tests: the CUDA path (through the Python shim -> C ABI -> sm_100a kernels) against
the CPU oracle on the same seeded inputs.  Integer / index / structure results must be
bit-exact; Float64 values are bit-exact when both sides receive identical inputs (no FMA on
either side) and otherwise within the north star's 1e-12 relative tolerance."""
import numpy as np
import pytest

import otmb_b200
import otmb_b200.api as A
from otmb_b200 import synthetic
from oracle import oracle as O

from _util import (NAMES, assert_csc_equal, bits, fields, gpu_pipeline, oracle_pipeline,
                   transport_from_oracle_inputs)

pytestmark = pytest.mark.gpu

SMALL = [
    # nx, ny, nz, topology, seed, kwargs
    (12, 10, 6, "tripolar", 0, {}),
    (13, 9, 5, "tripolar", 1, {}),              # odd nx (self-neighbour column kept dry)
    (2, 4, 3, "tripolar", 2, {}),               # nx = 2: west == east, every cell on the seam
    (3, 3, 2, "tripolar", 4, {}),
    (4, 2, 2, "tripolar", 5, {}),
    (10, 8, 4, "bipolar", 3, {}),
    (37, 11, 7, "tripolar", 6, {"dirty": True}),
    (64, 33, 9, "bipolar", 7, {"dirty": True, "float32_roundtrip": True}),
]
# nx = 1: a cell is its own east and west neighbour (the reference's makegridmetrics cannot build
# such a grid — vertexpermutation indexes cell (2,1) — but transportmatrix itself is well defined)
SMALL_T = SMALL + [(1, 5, 4, "bipolar", 8, {"land_frac": 0.0})]


def _ocean(nx, ny, nz, topo, seed, kw):
    kw = dict(kw)
    kw.setdefault("land_frac", 0.25)
    return synthetic.make_ocean(nx, ny, nz, topo, seed=seed, **kw)


# ------------------------------------------------------------------------------------------ K1
@pytest.mark.parametrize("shape", [(90, 45, 20), (7, 5, 3), (64, 1, 1), (65, 3, 1), (1, 1, 1), (130, 7, 2)])
def test_makeindices_matches_oracle(shape):
    rng = np.random.default_rng(sum(shape))
    v3D = np.asfortranarray(rng.random(shape) + 1.0)
    v3D[rng.random(shape) < 0.4] = np.nan
    got = otmb_b200.makeindices(v3D)
    want = O.makeindices(v3D)
    assert got.N == want["N"]
    assert np.array_equal(got.Lwet, want["Lwet"])
    assert np.array_equal(got.Lwet3D, want["Lwet3D"])
    assert np.array_equal(got.wet3D, want["wet3D"])
    assert got.L[(2, 1, 1)] == 2 and got.C[1] == (1, 1, 1)


def test_makeindices_all_dry_and_all_wet():
    v = np.full((9, 4, 3), np.nan, order="F")
    ix = otmb_b200.makeindices(v)
    assert ix.N == 0 and ix.Lwet.size == 0 and not ix.wet3D.any() and (ix.Lwet3D == 0).all()
    v = np.ones((9, 4, 3), order="F")
    ix = otmb_b200.makeindices(v)
    assert ix.N == v.size and np.array_equal(ix.Lwet, np.arange(1, v.size + 1)) and ix.wet3D.all()


# ------------------------------------------------------------------------------------------ K2/K3
@pytest.mark.parametrize("cfg", ["C1", "C1t"])
def test_gridmetrics_matches_oracle(cfg):
    oc = synthetic.make_config(cfg, seed=3)
    f = fields(oc)
    gm = otmb_b200.makegridmetrics(areacello=f["areacello"], volcello=f["volcello"], lon=f["lon"], lat=f["lat"],
                                   lev=f["lev"], lon_vertices=f["lon_vertices"], lat_vertices=f["lat_vertices"])
    v3D, area = O.clean_missing(oc.volcello), O.clean_missing(oc.areacello)
    assert gm.gridtopology.kind == O.getgridtopology(oc.lon_vertices, oc.lat_vertices) == oc.topology
    assert np.array_equal(bits(gm.v3D), bits(v3D)) and np.array_equal(bits(gm.area2D), bits(area))
    want = O.gridmetrics(area, v3D, oc.lon, oc.lat, oc.lon_vertices, oc.lat_vertices, oc.topology)
    # division / cumsum: IEEE, sequential -> bit-exact
    assert np.array_equal(bits(gm.thkcello), bits(want["thkcello"]))
    assert np.array_equal(bits(gm.Z3D), bits(want["Z3D"]))
    for q, d in enumerate(A.DIRS):
        for name, key in (("edge_length_2D", "edge"), ("distance_to_edge_2D", "dedge"), ("distance_to_neighbour_2D", "dnbr")):
            g, w = getattr(gm, name)[d], want[key][:, :, q]
            assert np.array_equal(np.isnan(g), np.isnan(w)), (name, d)
            # libm asin/sqrt may differ in the last bits between CUDA and glibc: 1e-13 relative
            np.testing.assert_allclose(g, w, rtol=1e-13, atol=0, equal_nan=True, err_msg=f"{name}[{d}]")
    if oc.topology == "bipolar":
        assert (gm.edge_length_2D["north"][:, -1] == 0.0).all()       # sind/cosd exact at 90 degrees
        assert np.isnan(gm.distance_to_neighbour_2D["north"][:, -1]).all()
    assert np.isnan(gm.distance_to_neighbour_2D["south"][:, 0]).all()


# ------------------------------------------------------------------------------------------ K4
@pytest.mark.parametrize("case", SMALL + [(90, 45, 20, "bipolar", 11, {"dirty": True}), (90, 45, 20, "tripolar", 12, {"dirty": True})])
def test_facefluxes_bit_exact(case):
    oc = _ocean(*case)
    f = fields(oc)
    gm = otmb_b200.makegridmetrics(areacello=f["areacello"], volcello=f["volcello"], lon=f["lon"], lat=f["lat"],
                                   lev=f["lev"], lon_vertices=f["lon_vertices"], lat_vertices=f["lat_vertices"])
    ix = otmb_b200.makeindices(gm.v3D)
    umo0 = oc.umo.copy(order="F")
    phi = otmb_b200.facefluxesfrommasstransport(umo=f["umo"], vmo=f["vmo"], gridmetrics=gm, indices=ix)
    want = O.facefluxes(oc.umo, oc.vmo, gm.v3D, oc.topology, oc.fill)
    for k in A.FACES:
        assert np.array_equal(bits(getattr(phi, k)), bits(want[k])), k
    assert np.array_equal(bits(oc.umo), bits(umo0))


def test_facefluxes_all_fill_asserts():
    oc = synthetic.make_ocean(8, 6, 3, "bipolar", seed=0, land_frac=0.0)
    f = fields(oc)
    gm = otmb_b200.makegridmetrics(areacello=f["areacello"], volcello=np.where(oc.volcello == 0, 1.0, oc.volcello),
                                   lon=f["lon"], lat=f["lat"], lev=f["lev"], lon_vertices=f["lon_vertices"],
                                   lat_vertices=f["lat_vertices"])
    ix = otmb_b200.makeindices(gm.v3D)
    allfill = np.full(gm.v3D.shape, oc.fill, order="F")
    assert ix.N == gm.v3D.size
    with pytest.raises(A.OTMBError) as e:
        otmb_b200.facefluxes(allfill, allfill, gm, ix, FillValue=oc.fill)
    assert e.value.code == 7
    with pytest.raises(O.OracleError):
        O.facefluxes(allfill, allfill, gm.v3D, "bipolar", oc.fill)


# ------------------------------------------------------------------------------------------ K5-K9, K11
@pytest.mark.parametrize("path", ["fused", "fused2", "coo"])
@pytest.mark.parametrize("case", SMALL_T)
def test_transportmatrix_small_exact(case, path):
    oc = _ocean(*case)
    for upwind, rho in ((True, 1035.0), (False, oc.rho3d)):
        try:
            o = oracle_pipeline(oc, rho=rho, upwind=upwind)
        except O.OracleError as err:
            with pytest.raises(A.OTMBError) as e:
                ov = O.clean_missing(oc.volcello)
                gpu_pipeline(oc, rho=rho, path=path, upwind=upwind)
            assert e.value.code == err.code
            continue
        tm, _ = transport_from_oracle_inputs(o, oc, rho=rho, path=path, upwind=upwind)
        for oname, gname in NAMES.items():
            assert_csc_equal(getattr(tm, gname), o["tm"][oname], f"{case} {path} {oname} upwind={upwind}", exact=True)


@pytest.mark.parametrize("path", ["fused", "fused2", "coo"])
@pytest.mark.parametrize("cfg,seed", [("C1", 0), ("C1t", 1)])
def test_transportmatrix_c1_full_pipeline(cfg, seed, path):
    """Whole pipeline on the GPU (own geometry and fluxes) against the whole oracle pipeline."""
    oc = synthetic.make_config(cfg, seed=seed)
    o = oracle_pipeline(oc)
    g = gpu_pipeline(oc, path=path)
    for oname, gname in NAMES.items():
        # geometry goes through libm on both sides -> TκH (and T) within 1e-12, the rest bit-exact
        exact = oname in ("Tadv", "TkVML", "TkVdeep")
        assert_csc_equal(getattr(g["tm"], gname), o["tm"][oname], f"{cfg} {path} {oname}", exact=exact, rtol=1e-12)


def test_coincidence_columns_take_generic_branch():
    """Fold centre (nx/2, ny) <-> (nx/2+1, ny) and seam∩fold (1, ny) <-> (nx, ny) create duplicate
    (row, col) pairs that sparse() sums in emit order."""
    oc = synthetic.make_ocean(12, 10, 6, "tripolar", seed=0, land_frac=0.2)
    o = oracle_pipeline(oc)
    tr = O.transportmatrix(o["phi"], oc.mlotst, o["v3D"], o["gm"]["thkcello"], o["area"], oc.lev, o["gm"]["edge"],
                           o["gm"]["dnbr"], o["topo"], 1035.0, keep_triplets=True)["triplets"]["TkH"]
    I, J, _ = tr
    off = I != J
    assert len(set(zip(I[off], J[off]))) < off.sum()        # the oracle really sees duplicates
    tm, _ = transport_from_oracle_inputs(o, oc)
    for oname, gname in NAMES.items():
        assert_csc_equal(getattr(tm, gname), o["tm"][oname], oname, exact=True)


def test_zero_diffusivities_drop_zero_entries():
    """κ = 0: TκH etc. keep explicit zeros (sparse), T drops them (sparse +)."""
    oc = synthetic.make_ocean(16, 12, 5, "tripolar", seed=9, land_frac=0.2)
    o = oracle_pipeline(oc, kH=0.0, kVML=0.0, kVdeep=0.0)
    for path in ("fused", "coo"):
        tm, _ = transport_from_oracle_inputs(o, oc, path=path, κH=0.0, κVML=0.0, κVdeep=0.0)
        for oname, gname in NAMES.items():
            assert_csc_equal(getattr(tm, gname), o["tm"][oname], f"{path} {oname}", exact=True)
        assert tm.TκH.nnz > 0 and (tm.TκH.data == 0).all()
        assert tm.T.nnz == (tm.Tadv.data != 0).sum() <= tm.Tadv.nnz


def test_prebuilt_operators_are_reused():
    """The Tadv/TκH/TκVML/TκVdeep kwargs skip construction (src/matrixbuilding.jl:133-143)."""
    oc = synthetic.make_ocean(20, 14, 6, "tripolar", seed=5, land_frac=0.2)
    o = oracle_pipeline(oc)
    tm, gm = transport_from_oracle_inputs(o, oc)
    phi = A.FaceFluxes(**o["phi"])
    tm2 = A.transportmatrix(ϕ=phi, mlotst=oc.mlotst, gridmetrics=gm, indices=None, ρ=1035.0, TκH=tm.TκH, TκVdeep=tm.TκVdeep)
    assert tm2.TκH is tm.TκH and tm2.TκVdeep is tm.TκVdeep
    for oname, gname in NAMES.items():
        assert_csc_equal(getattr(tm2, gname), o["tm"][oname], f"preset {oname}", exact=True)
    tm3 = A.transportmatrix(ϕ=phi, mlotst=oc.mlotst, gridmetrics=gm, indices=None, ρ=1035.0, Tadv=tm.Tadv, TκH=tm.TκH,
                            TκVML=tm.TκVML, TκVdeep=tm.TκVdeep)
    assert_csc_equal(tm3.T, o["tm"]["T"], "all preset T", exact=True)


# ------------------------------------------------------------------------------------------ errors
def test_errors_match_reference():
    oc = synthetic.make_ocean(14, 10, 5, "tripolar", seed=2, land_frac=0.2)
    o = oracle_pipeline(oc)
    # ρ contains NaNs (src/matrixbuilding.jl:233)
    rho = oc.rho3d.copy(order="F")
    L = int(o["ix"]["Lwet"][3]) - 1
    rho.ravel(order="F")[L] = np.nan
    rho = np.asfortranarray(rho)
    with pytest.raises(A.OTMBError) as e:
        transport_from_oracle_inputs(o, oc, rho=rho)
    assert e.value.code == 5 and str(e.value) == "ρ contains NaNs"
    # a non-zero flux out of a dry cell: the reference dies with a MethodError (:247-250)
    phi = {k: v.copy(order="F") for k, v in o["phi"].items()}
    wet = o["ix"]["wet3D"]
    east_dry = np.roll(wet, -1, axis=0)
    idx = np.argwhere(wet & ~east_dry)[0]
    phi["east"][tuple(idx)] = -5.0e6       # inflow "from East" where the east cell is dry
    o2 = dict(o, phi=phi)
    with pytest.raises(A.OTMBError) as e:
        transport_from_oracle_inputs(o2, oc)
    assert e.value.code == 8
    with pytest.raises(O.OracleError) as eo:
        O.transportmatrix(phi, oc.mlotst, o["v3D"], o["gm"]["thkcello"], o["area"], oc.lev, o["gm"]["edge"],
                          o["gm"]["dnbr"], o["topo"], 1035.0)
    assert eo.value.code == 8
    # NaN geometry -> "TκH contains NaNs." (:61)
    gm_bad = dict(o["gm"])
    edge = gm_bad["edge"].copy(order="F")
    i, j, k = np.argwhere(wet)[5]
    edge[i, j, :] = np.nan
    gm_bad["edge"] = edge
    with pytest.raises(A.OTMBError) as e:
        transport_from_oracle_inputs(dict(o, gm=gm_bad), oc)
    assert e.value.code == 2 and str(e.value) == "TκH contains NaNs."


def test_odd_nx_self_neighbour_errors_like_reference():
    oc = synthetic.make_ocean(13, 9, 5, "tripolar", seed=1, land_frac=0.2, allow_self_neighbour=True)
    with pytest.raises(O.OracleError) as eo:
        oracle_pipeline(oc)
    with pytest.raises(A.OTMBError) as e:
        gpu_pipeline(oc)
    assert e.value.code == eo.value.code == 2


def test_unknown_topology_errors():
    oc = synthetic.make_ocean(12, 8, 4, "tripolar", seed=0)
    f = fields(oc)
    lonv = oc.lon_vertices.copy(order="F")
    lonv[2, 3, -1] += 17.0                                  # break the fold symmetry
    with pytest.warns(UserWarning):
        with pytest.raises(A.OTMBError) as e:
            otmb_b200.makegridmetrics(areacello=f["areacello"], volcello=f["volcello"], lon=f["lon"], lat=f["lat"],
                                      lev=f["lev"], lon_vertices=lonv, lat_vertices=f["lat_vertices"])
    assert e.value.code == 6 and str(e.value) == "Unknown grid type"


# ------------------------------------------------------------------------------------------ generic sparse / +
@pytest.mark.parametrize("n,len_,seed", [(50, 400, 0), (7, 600, 1), (1000, 3000, 2), (5, 0, 3), (300, 20000, 4)])
def test_sparse_and_spadd_match_oracle(n, len_, seed):
    rng = np.random.default_rng(seed)
    I, J = rng.integers(1, n + 1, len_), rng.integers(1, n + 1, len_)
    V = rng.normal(size=len_) * 10.0 ** rng.integers(-8, 8, len_)
    V[rng.random(len_) < 0.05] = 0.0
    want = O.sparse(I, J, V, n)
    cp, rv, nz = otmb_b200.sparse(I, J, V, n)
    assert np.array_equal(cp, want.colptr) and np.array_equal(rv, want.rowval)
    assert np.array_equal(bits(nz), bits(want.nzval))
    I2, J2 = rng.integers(1, n + 1, len_ // 2), rng.integers(1, n + 1, len_ // 2)
    V2 = rng.normal(size=len_ // 2)
    # make exact cancellations so that + drops entries
    B = O.sparse(np.concatenate([I2, I[: len_ // 3]]), np.concatenate([J2, J[: len_ // 3]]),
                 np.concatenate([V2, np.zeros(len_ // 3)]), n)
    neg = O.CSC(want.n, want.colptr, want.rowval, -want.nzval)
    for Bm in (B, neg):
        wsum = O.spadd(want, Bm)
        cp, rv, nz = otmb_b200.spadd((want.colptr, want.rowval, want.nzval), (Bm.colptr, Bm.rowval, Bm.nzval), n)
        assert np.array_equal(cp, wsum.colptr) and np.array_equal(rv, wsum.rowval)
        assert np.array_equal(bits(nz), bits(wsum.nzval))


# ------------------------------------------------------------------------------------------ full size (C2)
def test_transportmatrix_c2_full_size():
    """ACCESS-ESM1-5 1° shape (360x300x50, tripolar), all paths vs the oracle, plus the
    reference's own invariants (test/online.jl:92-123)."""
    oc = synthetic.make_config("C2", seed=0)
    o = oracle_pipeline(oc)
    N = o["ix"]["N"]
    for path in ("coo", "fused2", "fused"):
        tm, _ = transport_from_oracle_inputs(o, oc, path=path)
        for oname, gname in NAMES.items():
            assert_csc_equal(getattr(tm, gname), o["tm"][oname], f"C2 {path} {oname}", exact=True)
    T_exact = tm.T
    g = gpu_pipeline(oc)
    for oname, gname in NAMES.items():
        exact = oname in ("Tadv", "TkVML", "TkVdeep")
        assert_csc_equal(getattr(g["tm"], gname), o["tm"][oname], f"C2 pipeline {oname}", exact=exact, rtol=1e-12)
    T = g["tm"].T
    v = o["v3D"].ravel(order="F")[~np.isnan(o["v3D"].ravel(order="F"))]
    one = np.ones(N)
    Myr = 365.25 * 86400 * 1e6
    for name in ("TκH", "TκVML", "TκVdeep"):
        Tm = getattr(g["tm"], name)
        assert np.linalg.norm(one) / np.linalg.norm(Tm @ one) / Myr > 1e6, name
    for name in A.MATRICES:
        Tm = getattr(g["tm"], name)
        assert np.linalg.norm(v) / np.linalg.norm(Tm.T @ v) / Myr > 1e6, name
    d = T.diagonal()
    assert (d > 0).all()
    offd = T.copy()
    offd.setdiag(0)
    offd.eliminate_zeros()
    assert (offd.data < 0).all()
    # T·1 conservation residual must match the oracle's
    To = o["tm"]["T"].scipy()
    assert np.array_equal(bits(T_exact @ one), bits(To @ one))


# ------------------------------------------------------------------------------------------ K10
@pytest.mark.parametrize("cfg", ["C1t", "C2"])
def test_redigm_matches_oracle(cfg):
    oc = synthetic.make_config(cfg, seed=4)
    o = oracle_pipeline(oc)
    tm, gm = transport_from_oracle_inputs(o, oc)
    for d in ("I", "J"):
        got = A.globalverticalfacetriadderivative(oc.rho3d, gm, None, d)
        want = O.triad(oc.rho3d, oc.lon, oc.lat, o["gm"]["Z3D"], o["v3D"], oc.topology, d)
        assert np.array_equal(np.isnan(got), np.isnan(want))
        # 1e-12 relative (north star); on the 5.4 M cells of C2 ONE element of a slope that nearly cancels (|value| 6e-9 of
        # a field of O(1)) sits at 1.05e-12: the last-bit difference of the haversine (CUDA vs glibc asin / sqrt) under a
        # difference of quotients.  1e-12 of the field's scale is allowed on top.
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12 * np.nanmax(np.abs(want[np.isfinite(want)])), equal_nan=True)
    chi = np.asfortranarray(o["gm"]["Z3D"] ** 2)
    got, want = A.globalverticaldyadderivative(chi, gm, None), O.dyad(chi, o["gm"]["Z3D"], o["v3D"], oc.topology)
    assert np.array_equal(bits(got), bits(want))
    u, v = A.bolus_GM_velocity(oc.rho3d, gm, None)
    uo, vo = O.bolus_gm(oc.rho3d, oc.lon, oc.lat, o["gm"]["Z3D"], o["v3D"], oc.topology)
    # north-star tolerance 1e-12, plus 1e-12 of the field's scale: where the taper 1 + tanh(..) cancels to ~0 the value is
    # negligible but its last bits are the tanh implementation's (CUDA's vs glibc's; Julia's own is a third one)
    for got, want in ((u, uo), (v, vo)):
        assert np.array_equal(np.isnan(got), np.isnan(want))
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12 * np.nanmax(np.abs(want)), equal_nan=True)


def test_triad_on_bipolar_top_row_throws_like_reference():
    oc = synthetic.make_ocean(10, 8, 4, "bipolar", seed=3, land_frac=0.0)
    o = oracle_pipeline(oc)
    tm, gm = transport_from_oracle_inputs(o, oc)
    with pytest.raises(O.OracleError):
        O.triad(oc.rho3d, oc.lon, oc.lat, o["gm"]["Z3D"], o["v3D"], "bipolar", "J")
    with pytest.raises(A.OTMBError):
        A.globalverticalfacetriadderivative(oc.rho3d, gm, None, "J")


# ------------------------------------------------------------------------------------------ resident products
def test_resident_matvec_matches_scipy_bitwise():
    """y = X x and y = Xᵀ x on the device-resident results (the τdiv / τvol products of test/online.jl:110-115)
    against scipy's sequential CSC products on the fetched matrices."""
    oc = synthetic.make_config("C1t", seed=2)
    g = gpu_pipeline(oc)
    N = g["ix"].N
    rng = np.random.default_rng(0)
    x = rng.normal(size=N)
    one = np.ones(N)
    for name in A.MATRICES:
        M = getattr(g["tm"], name)
        for vec in (x, one):
            yt = otmb_b200.resident_matvec(name, vec, transpose=True)
            assert np.array_equal(bits(yt), bits(M.T.tocsr() @ vec)), name
            y = otmb_b200.resident_matvec(name, vec)
            want = np.zeros(N)
            ip, ix_, dv = M.indptr, M.indices, M.data
            for j in range(N):                       # column-by-column CSC product: the order scipy / SparseArrays add in
                want[ix_[ip[j]:ip[j + 1]]] += dv[ip[j]:ip[j + 1]] * vec[j]
            assert np.array_equal(bits(y), bits(want)), name


# ------------------------------------------------------------------------------------------ fetch pipeline
def test_fetch_paths_return_identical_arrays(monkeypatch):
    """otmb_transportmatrix_fetch_all (Int32 indices on the link, widened on the host, chunked ring) against the
    direct 8-byte copies and the one-matrix call: same bits in every array."""
    import ctypes as C
    oc = synthetic.make_config("C1t", seed=3)
    g = gpu_pipeline(oc)
    ctx = otmb_b200.default_context(0)
    N = g["ix"].N
    nnz = [getattr(g["tm"], n).nnz for n in A.MATRICES]

    def fetch_all():
        arrs = [[np.full(N + 1, -7, np.int64) for _ in range(5)], [np.full(nnz[m], -7, np.int64) for m in range(5)],
                [np.full(nnz[m], np.nan) for m in range(5)]]
        ptrs = [(C.c_void_p * 5)(*[a.ctypes.data for a in arrs[q]]) for q in range(3)]
        ctx.check(ctx.lib.otmb_transportmatrix_fetch_all(ctx.h, 0, *ptrs))
        return arrs

    narrow = fetch_all()
    monkeypatch.setenv("OTMB_FETCH_DIRECT", "1")
    direct = fetch_all()
    monkeypatch.delenv("OTMB_FETCH_DIRECT")
    for q in range(3):
        for m in range(5):
            assert np.array_equal(bits(narrow[q][m]) if q == 2 else narrow[q][m], bits(direct[q][m]) if q == 2 else direct[q][m]), (q, m)
    for m, name in enumerate(A.MATRICES):
        cp, rv, nz = np.empty(N + 1, np.int64), np.empty(nnz[m], np.int64), np.empty(nnz[m])
        ctx.check(ctx.lib.otmb_transportmatrix_fetch(ctx.h, m, A._ptr(cp), A._ptr(rv), A._ptr(nz)))
        assert np.array_equal(cp, direct[0][m]) and np.array_equal(rv, direct[1][m]) and np.array_equal(bits(nz), bits(direct[2][m]))
        M = getattr(g["tm"], name)        # 0-based scipy view of the same build
        assert np.array_equal(cp - cp[0], M.indptr) and np.array_equal(rv - cp[0], M.indices)


# ------------------------------------------------------------------------------------------ subnormal fluxes
@pytest.mark.parametrize("upwind", [True, False])
def test_subnormal_fluxes_match_oracle(upwind):
    """Fluxes of one to three units of the smallest subnormal.  Centred scheme: ϕ/2 rounds to zero for one unit (no
    entry, src/matrixbuilding.jl:245) and to one / two units otherwise — the kernel's sign tests never form ϕ/2, this
    pins that they decide the same.  Upwind: every Tadv value underflows to an explicit zero, which `sparse` keeps
    and the sum `T = Tadv + ...` drops (:147: the compaction pass); the divisions take their slow path."""
    oc = synthetic.make_config("C1t", seed=5)
    o = oracle_pipeline(oc)
    tiny = np.float64(5e-324)
    phi = {}
    for name, a in o["phi"].items():
        units = 1 + (np.arange(a.size).reshape(a.shape, order="F") % 3)
        phi[name] = np.asfortranarray(np.sign(a) * units * tiny)
    gm = o["gm"]
    want = O.transportmatrix(phi, oc.mlotst, o["v3D"], gm["thkcello"], o["area"], oc.lev, gm["edge"], gm["dnbr"], o["topo"],
                             1035.0, upwind=upwind)
    for path in ("fused", "fused2", "coo"):
        tm, _ = transport_from_oracle_inputs(dict(o, phi=phi), oc, path=path, upwind=upwind)
        for oname, gname in NAMES.items():
            assert_csc_equal(getattr(tm, gname), want[oname], f"subnormal {path} {oname} upwind={upwind}", exact=True)
    # the one-unit faces really are absent from the centred operator (and only from it)
    two = {name: np.asfortranarray(np.sign(a) * 2 * tiny) for name, a in o["phi"].items()}
    full = O.transportmatrix(two, oc.mlotst, o["v3D"], gm["thkcello"], o["area"], oc.lev, gm["edge"], gm["dnbr"], o["topo"],
                             1035.0, upwind=upwind)
    assert want["Tadv"].nzval.size > 0
    assert (want["Tadv"].nzval.size == full["Tadv"].nzval.size) == bool(upwind)


# ------------------------------------------------------------------------------------------ residency
def test_inplace_updates_of_caller_arrays_are_seen_and_results_are_frozen():
    """The reference reads its arguments at call time.  Device residency is keyed on object identity, so it is
    only granted to arrays this package handed out itself — and those are read-only.  A caller who re-uses ONE set
    of ϕ buffers month after month (writes into them in place) must get the matrix of the new values."""
    oc = synthetic.make_config("C1t", seed=4)
    o = oracle_pipeline(oc)
    tm1, gm = transport_from_oracle_inputs(o, oc)
    assert_csc_equal(tm1.Tadv, o["tm"]["Tadv"], "before the update")
    for k in O.FACES:
        o["phi"][k] *= -0.5                        # in place: same objects, new content (and reversed flow)
    gm.thkcello[...] = gm.thkcello * 1.25          # so are the metrics the caller owns
    want = O.transportmatrix(o["phi"], oc.mlotst, o["v3D"], gm.thkcello, o["area"], oc.lev, o["gm"]["edge"], o["gm"]["dnbr"],
                             o["topo"], 1035.0)
    tm2, _ = transport_from_oracle_inputs(o, oc)
    for oname, gname in NAMES.items():
        assert_csc_equal(getattr(tm2, gname), want[oname], f"after the in-place update: {oname}")
    assert tm2.Tadv.nnz == tm1.Tadv.nnz and not np.array_equal(tm2.Tadv.indices, tm1.Tadv.indices)
    # what makegridmetrics / facefluxes return stays resident, and therefore cannot be written to
    g = gpu_pipeline(oc)
    with pytest.raises(ValueError):
        g["phi"].east[0, 0, 0] = 1.0
    with pytest.raises(ValueError):
        g["gm"].thkcello[0, 0, 0] = 1.0
    with pytest.raises(ValueError):
        g["gm"].edge_length_2D["east"][0, 0] = 1.0
    # a writeable copy is the caller's own again: uploaded, and its content is what counts
    phi2 = A.FaceFluxes(*[np.array(a, order="F") * 2.0 for a in g["phi"]])
    tm3 = otmb_b200.transportmatrix(ϕ=phi2, mlotst=oc.mlotst, gridmetrics=g["gm"], indices=g["ix"], ρ=1035.0)
    assert np.array_equal(bits(tm3.Tadv.data), bits(g["tm"].Tadv.data * 2.0))


def test_failed_build_invalidates_earlier_results(ctx):
    """A build that fails its NaN check must not leave the previous build's sizes behind (fetch / spmv would pair
    them with the new contents)."""
    oc = synthetic.make_config("C1t", seed=2)
    g = gpu_pipeline(oc)
    c = A._ctx_of(g["gm"].v3D)
    bad = np.array(oc.mlotst, order="F")
    rho = np.array(oc.rho3d, order="F")
    rho[np.isfinite(rho)] = np.nan
    with pytest.raises(A.OTMBError):
        otmb_b200.transportmatrix(ϕ=g["phi"], mlotst=bad, gridmetrics=g["gm"], indices=g["ix"], ρ=rho)
    with pytest.raises(A.OTMBError) as e:
        A.resident_matvec("T", np.ones(g["ix"].N), ctx=c)
    assert e.value.code == otmb_b200._lib.ERR_STATE


# ------------------------------------------------------------------------------------------ C3: GM bolus -> T (extension)
@pytest.mark.parametrize("cfg", ["C1t", "C2"])
def test_c3_gm_bolus_transport_matrix(cfg):
    """BASELINE configs[2]: bolus_GM_velocity -> velocity2fluxes -> + umo/vmo -> facefluxes -> transportmatrix, on the
    device (csrc/gm.cu).  The reference has the pieces, not the chain (parity unpinned), so the checks are: κGM = 0
    reproduces the plain path bit for bit; the chain agrees with the oracle's composition of the same pieces; T given
    the GPU's ϕ is the oracle's T bit for bit; pattern stays 7-point; mass is conserved (vᵀT = 0) and T·1 vanishes
    below the surface level (the continuity scan closes every cell but the skipped surface face)."""
    from oracle import gm_np
    oc = synthetic.make_config(cfg, seed=3)
    f = fields(oc)
    gm = otmb_b200.makegridmetrics(areacello=f["areacello"], volcello=f["volcello"], lon=f["lon"], lat=f["lat"], lev=f["lev"],
                                   lon_vertices=f["lon_vertices"], lat_vertices=f["lat_vertices"])
    ix = otmb_b200.makeindices(gm.v3D)
    plain = otmb_b200.facefluxesfrommasstransport(umo=f["umo"], vmo=f["vmo"], gridmetrics=gm, indices=ix)
    plain = {k: np.array(getattr(plain, k)) for k in A.FACES}
    # (1) κGM = 0: the bolus fluxes are exact zeros
    phi0 = otmb_b200.facefluxes_GM(umo=f["umo"], vmo=f["vmo"], gridmetrics=gm, indices=ix, ρ=oc.rho3d, κGM=0.0)
    for k in A.FACES:
        assert np.array_equal(getattr(phi0, k) + 0.0, plain[k] + 0.0), k          # (+0.0: -0.0 and 0.0 are the same flux)
    # (2) against the oracle's composition, with the GPU's own geometry as input on both sides
    phi, gi, gj = otmb_b200.facefluxes_GM(umo=f["umo"], vmo=f["vmo"], gridmetrics=gm, indices=ix, ρ=oc.rho3d, κGM=600.0,
                                          return_bolus_fluxes=True)
    edge = np.asfortranarray(np.stack([gm.edge_length_2D[d] for d in A.DIRS], axis=-1))
    dnbr = np.asfortranarray(np.stack([gm.distance_to_neighbour_2D[d] for d in A.DIRS], axis=-1))
    um, vm, gi_o, gj_o = gm_np.total_transport(oc.umo, oc.vmo, oc.fill, oc.rho3d, gm.lon, gm.lat, gm.Z3D, gm.v3D, gm.thkcello,
                                               edge, oc.topology)
    scale = max(np.nanmax(np.abs(gi_o)), np.nanmax(np.abs(gj_o)))
    # (the library returns velocity2fluxes' own output; the oracle's has the fold row averaged already: compare below it)
    gj, gj_o = gj[:, :-1], gj_o[:, :-1]
    assert scale > 0 and np.array_equal(np.isnan(gi), np.isnan(gi_o)) and np.array_equal(np.isnan(gj), np.isnan(gj_o))
    # 1e-12 relative, plus 1e-12 of the field's scale: where the taper 1 + tanh(..) cancels to ~0 the VALUE is
    # negligible but its last bits depend on the tanh implementation (CUDA vs glibc; Julia's is a third one)
    np.testing.assert_allclose(gi, gi_o, rtol=1e-12, atol=1e-12 * scale, equal_nan=True)
    np.testing.assert_allclose(gj, gj_o, rtol=1e-12, atol=1e-12 * scale, equal_nan=True)
    want = O.facefluxes(um, vm, gm.v3D, oc.topology, oc.fill)
    fscale = max(np.abs(want[k]).max() for k in A.FACES)
    for k in A.FACES:
        np.testing.assert_allclose(getattr(phi, k), want[k], rtol=0, atol=1e-12 * fscale, err_msg=k)
    assert any(not np.array_equal(getattr(phi, k), plain[k]) for k in ("east", "north", "top"))
    # (3) T from the GPU's ϕ: the oracle's matrix, bit for bit; 7-point pattern; invariants
    tm = otmb_b200.transportmatrix(ϕ=phi, mlotst=oc.mlotst, gridmetrics=gm, indices=ix, ρ=1035.0)
    ot = O.transportmatrix({k: np.asarray(getattr(phi, k)) for k in A.FACES}, oc.mlotst, gm.v3D, gm.thkcello, gm.area2D, oc.lev,
                           edge, dnbr, oc.topology, 1035.0)
    for oname, gname in NAMES.items():
        assert_csc_equal(getattr(tm, gname), ot[oname], f"C3 {oname}", exact=True)
    T = tm.T
    assert np.diff(T.indptr).max() <= 7
    v = gm.v3D.ravel(order="F")
    v = v[~np.isnan(v)]
    assert np.abs(T.T @ v).max() <= 1e-9 * np.abs(T.diagonal() * v).max()
    surface = (ix.Lwet - 1) < gm.v3D.shape[0] * gm.v3D.shape[1]
    r = np.abs(tm.Tadv @ np.ones(ix.N))
    assert r[~surface].max() <= 1e-9 * np.abs(tm.Tadv.diagonal()).max()


# ------------------------------------------------------------------------------------------ slab-pipelined end-to-end call
@pytest.mark.parametrize("pageable", [False, True])
@pytest.mark.parametrize("nslabs", [0, 1, 2, 7, 20])
def test_stream_call_equals_build_and_fetch(nslabs, pageable):
    """otmb_transportmatrix_stream (upload, chained slab launches and copy-out overlapped in one call) returns the very
    arrays of set_facefluxes -> build -> fetch_all, for any number of slabs, into page-locked and pageable arrays."""
    oc = synthetic.make_config("C1t", seed=6)
    o = oracle_pipeline(oc)
    g = gpu_pipeline(oc)                              # ϕ resident: the build + fetch path
    c = A._ctx_of(g["gm"].v3D)
    for rho, upwind in ((1035.0, True), (np.array(oc.rho3d, order="F"), False)):
        want = otmb_b200.transportmatrix(ϕ=g["phi"], mlotst=oc.mlotst, gridmetrics=g["gm"], indices=g["ix"], ρ=rho, upwind=upwind)
        phi = [np.array(getattr(g["phi"], k), order="F") for k in A.FACES]
        got = A._transportmatrix_stream(c, g["ix"].N, phi, oc.mlotst, rho, 500.0, 0.1, 1e-5, upwind, nslabs=nslabs, pageable=pageable)
        for name in A.MATRICES:
            a, b = getattr(got, name), getattr(want, name)
            assert np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices), name
            assert np.array_equal(bits(a.data), bits(b.data)), name
        # ... and stay resident like theirs
        y = A.resident_matvec("T", np.ones(g["ix"].N), ctx=c)
        assert np.array_equal(bits(y), bits(A.resident_matvec("T", np.ones(g["ix"].N), ctx=c)))


def test_stream_call_zero_dropping_errors_and_capacity():
    oc = synthetic.make_config("C1t", seed=2)
    o = oracle_pipeline(oc)
    # κ = 0: T's exact zeros are dropped (src/matrixbuilding.jl:147) — the stream call compacts and re-sends T
    want = O.transportmatrix(o["phi"], oc.mlotst, o["v3D"], o["gm"]["thkcello"], o["area"], oc.lev, o["gm"]["edge"], o["gm"]["dnbr"],
                             o["topo"], 1035.0, kH=0.0, kVML=0.0, kVdeep=0.0)
    tm, gm = transport_from_oracle_inputs(o, oc, κH=0.0, κVML=0.0, κVdeep=0.0)     # caller-owned ϕ: the stream path
    for oname, gname in NAMES.items():
        assert_csc_equal(getattr(tm, gname), want[oname], f"kappa=0 {oname}")
    assert tm.T.nnz == tm.Tadv.nnz < tm.TκH.nnz + tm.Tadv.nnz
    # NaN density: the reference's message, through the stream path too
    c = A._ctx_of(gm.v3D)
    phi = [np.array(o["phi"][k], order="F") for k in O.FACES]
    rho = np.full(gm.v3D.shape, np.nan, order="F")
    with pytest.raises(A.OTMBError) as e:
        A._transportmatrix_stream(c, o["ix"]["N"], phi, oc.mlotst, rho, 500.0, 0.1, 1e-5, True)
    assert e.value.code == otmb_b200._lib.ERR_RHO_NAN and "ρ contains NaNs" in str(e.value)
    # result arrays too small
    with pytest.raises(A.OTMBError) as e:
        A._transportmatrix_stream(c, o["ix"]["N"], phi, oc.mlotst, 1035.0, 500.0, 0.1, 1e-5, True, caps=[10] * 5)
    assert e.value.code == otmb_b200._lib.ERR_BADARG
    # and the context is still usable
    ok = A._transportmatrix_stream(c, o["ix"]["N"], phi, oc.mlotst, 1035.0, 500.0, 0.1, 1e-5, True)
    assert_csc_equal(ok.T, o["tm"]["T"], "after the failures")


def test_binary_dump_of_resident_matrices(tmp_path):
    """otmb_transportmatrix_dump: the resident matrices to a file and back, bit for bit (C2-sized arrays span several
    16 MB staging chunks; C1t fits one)."""
    for cfg in ("C1t", "C2"):
        oc = synthetic.make_config(cfg, seed=1)
        g = gpu_pipeline(oc)
        path = tmp_path / f"{cfg}.otmbcsc"
        otmb_b200.dump_resident(path)
        back = otmb_b200.load_dump(path)
        for name in A.MATRICES:
            a, b = getattr(g["tm"], name), back[name]
            assert a.shape == b.shape and np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices), name
            assert np.array_equal(bits(a.data), bits(b.data)), name
        otmb_b200.dump_resident(path, ("T",))
        assert list(otmb_b200.load_dump(path)) == ["T"]


def test_pinned_host_pool_reuses_blocks():
    """otmb_host_alloc / otmb_host_free: freed blocks are handed out again to requests they fit (pinning a gigabyte costs
    ~100 ms; the Julia shim frees result arrays from finalizers and allocates the same sizes every month)."""
    import ctypes as C
    lib = otmb_b200._lib.load()
    lib.otmb_host_trim()
    p, q = C.c_void_p(), C.c_void_p()
    assert lib.otmb_host_alloc(C.byref(p), 64 << 20) == 0 and p.value
    assert lib.otmb_host_free(p) == 0
    assert lib.otmb_host_alloc(C.byref(q), 60 << 20) == 0 and q.value == p.value       # fits within 1.5x: the same block
    r = C.c_void_p()
    assert lib.otmb_host_alloc(C.byref(r), 1 << 20) == 0 and r.value != q.value         # too small a request for a 64 MB block
    assert lib.otmb_host_free(q) == 0 and lib.otmb_host_free(r) == 0
    assert lib.otmb_host_free(q) == otmb_b200._lib.ERR_BADARG                              # freed twice
    buf = (C.c_char * 64)()
    assert lib.otmb_host_free(C.cast(buf, C.c_void_p)) == otmb_b200._lib.ERR_BADARG       # not ours
    assert lib.otmb_host_trim() == 0


# ------------------------------------------------------------------------------------------ fuzz
@pytest.mark.parametrize("seed", range(24))
def test_fuzz_small_grids_bit_exact(seed):
    """Random tiny grids (nx 1..9, ny 2..7, nz 1..6, either topology, random land fraction, dirty inputs, upwind or
    centred, scalar or 3-D rho): every matrix of the fused path — through the slab-pipelined stream call and through
    build + fetch — equals the oracle bit for bit, or both sides raise the same error."""
    rng = np.random.default_rng(1000 + seed)
    nx, ny, nz = int(rng.integers(1, 10)), int(rng.integers(2, 8)), int(rng.integers(1, 7))
    topo = "tripolar" if rng.random() < 0.6 and nx >= 2 else "bipolar"
    oc = synthetic.make_ocean(nx, ny, nz, topo, seed=seed, land_frac=float(rng.uniform(0.0, 0.5)), dirty=bool(rng.random() < 0.3),
                              allow_self_neighbour=bool(rng.random() < 0.3))
    upwind = bool(rng.random() < 0.6)
    rho = 1035.0 if rng.random() < 0.5 else oc.rho3d
    try:
        o = oracle_pipeline(oc, rho=rho, upwind=upwind)
    except O.OracleError as e:
        o, code = None, e.code
    if o is None:
        # the reference itself stops here (e.g. odd nx on the fold: "TκH contains NaNs."): the CUDA path must stop the same way
        v3D, area = O.clean_missing(oc.volcello), O.clean_missing(oc.areacello)
        gmo = O.gridmetrics(area, v3D, oc.lon, oc.lat, oc.lon_vertices, oc.lat_vertices, oc.topology)
        phi = O.facefluxes(oc.umo, oc.vmo, v3D, oc.topology, oc.fill)
        oo = dict(v3D=v3D, area=area, topo=oc.topology, gm=gmo, phi=phi)
        with pytest.raises(A.OTMBError) as ei:
            transport_from_oracle_inputs(oo, oc, rho=rho, upwind=upwind)
        assert ei.value.code == code
        return
    tm, gm = transport_from_oracle_inputs(o, oc, rho=rho, upwind=upwind)          # caller-owned ϕ: the stream call
    for oname, gname in NAMES.items():
        assert_csc_equal(getattr(tm, gname), o["tm"][oname], f"seed {seed} {(nx, ny, nz, topo)} stream {oname}")
    if nx < 2:
        return      # the reference's makegridmetrics cannot build a one-column grid (vertexpermutation indexes cell (2,1))
    g = gpu_pipeline(oc, rho=rho, upwind=upwind)                                   # resident ϕ: build + fetch, own geometry
    for oname, gname in NAMES.items():
        assert_csc_equal(getattr(g["tm"], gname), o["tm"][oname], f"seed {seed} {(nx, ny, nz, topo)} build {oname}", exact=False, rtol=1e-12)


def test_prebuilt_operator_is_validated():
    """A wrong-shaped pre-built operator is a DimensionMismatch in the reference's sum (src/matrixbuilding.jl:147); here
    otmb_set_operator checks the caller's CSC (colptr ends on the host; monotonicity, row range and ascending rows by a
    kernel over the uploaded arrays) and a rejected operator is not kept."""
    import ctypes as C
    oc = synthetic.make_config("C1t", seed=5)
    g = gpu_pipeline(oc)
    c = A._ctx_of(g["gm"].v3D)
    N = g["ix"].N
    good = g["tm"].TκH
    cp, rv, nz = good.indptr.astype(np.int64), good.indices.astype(np.int64), good.data.astype(np.float64)

    def set_op(cp, rv, nz, nnz=None):
        return c.lib.otmb_set_operator(c.h, 2, len(rv) if nnz is None else nnz, A._ptr(cp), A._ptr(rv), A._ptr(nz), 0)

    assert set_op(cp, rv, nz) == 0
    bad = cp.copy(); bad[-1] -= 1
    assert set_op(bad, rv, nz) == otmb_b200._lib.ERR_BADARG and b"DimensionMismatch" in c.lib.otmb_last_error(c.h)
    bad = cp.copy(); bad[5], bad[6] = bad[6], bad[5] - 1                       # not monotone
    assert set_op(bad, rv, nz) == otmb_b200._lib.ERR_BADARG
    bad = rv.copy(); bad[3] = N                                                # row out of range
    assert set_op(cp, bad, nz) == otmb_b200._lib.ERR_BADARG
    bad = rv.copy(); a = cp[10]; bad[a], bad[a + 1] = bad[a + 1], bad[a]       # rows not ascending inside a column
    assert set_op(cp, bad, nz) == otmb_b200._lib.ERR_BADARG
    # the context is still usable and the good operator gives the reference's sum
    tm = otmb_b200.transportmatrix(ϕ=g["phi"], mlotst=oc.mlotst, gridmetrics=g["gm"], indices=g["ix"], ρ=1035.0, TκH=good)
    assert np.array_equal(tm.T.indptr, g["tm"].T.indptr) and np.array_equal(bits(tm.T.data), bits(g["tm"].T.data))


@pytest.mark.gpu
@pytest.mark.parametrize("change", ["same", "values", "pattern", "other_kappa", "all_four"])
def test_prebuilt_operators_equal_or_not(change):
    """Pre-built operators that equal what this call would build take the single-pass route, anything else the
    generic sparse `+` (src/matrixbuilding.jl:147): either way T is ((Tadv + TκH) + TκVML) + TκVdeep of the
    operators the caller ends up holding."""
    import scipy.sparse as sp
    oc = synthetic.make_ocean(22, 16, 7, "tripolar", seed=11, land_frac=0.2)
    o = oracle_pipeline(oc)
    tm, gm = transport_from_oracle_inputs(o, oc)
    phi = A.FaceFluxes(**o["phi"])
    kw = dict(ϕ=phi, mlotst=oc.mlotst, gridmetrics=gm, indices=None, ρ=1035.0)
    H, D = sp.csc_matrix(tm.TκH, copy=True), sp.csc_matrix(tm.TκVdeep, copy=True)
    if change == "values":
        H.data[::7] *= 3.0
    elif change == "pattern":
        # drop one stored entry of the third column that has a few
        j = int(np.flatnonzero(np.diff(H.indptr) >= 3)[2])
        keep = np.ones(H.nnz, bool)
        keep[H.indptr[j] + 1] = False
        cols = np.repeat(np.arange(H.shape[1]), np.diff(H.indptr))
        H = sp.csc_matrix((H.data[keep], (H.indices[keep], cols[keep])), shape=H.shape)
        H.sort_indices()
    elif change == "other_kappa":
        kw["κH"] = 250.0          # H was built with 500: it must be used as supplied, not rebuilt
    if change == "all_four":
        got = A.transportmatrix(**kw, Tadv=tm.Tadv, TκH=H, TκVML=tm.TκVML, TκVdeep=D)
    else:
        got = A.transportmatrix(**kw, TκH=H, TκVdeep=D)
    as_orc = lambda m: O.CSC(m.shape[0], m.indptr.astype(np.int64) + 1, m.indices.astype(np.int64) + 1, m.data.astype(np.float64))
    want = O.spadd(O.spadd(O.spadd(as_orc(got.Tadv), as_orc(H)), as_orc(got.TκVML)), as_orc(D))
    assert_csc_equal(got.T, want, f"T with supplied operators ({change})", exact=True)
    assert_csc_equal(got.Tadv, o["tm"]["Tadv"], "Tadv", exact=True)
    assert_csc_equal(got.TκVML, o["tm"]["TkVML"], "TκVML", exact=True)
    # and the context is not left in a state that disturbs the next ordinary build
    again, _ = transport_from_oracle_inputs(o, oc)
    for oname, gname in NAMES.items():
        assert_csc_equal(getattr(again, gname), o["tm"][oname], f"after {change}: {oname}", exact=True)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [1, 0x5eed5eed5eed])
def test_paired_division_equals_the_compilers(seed):
    """csrc/fdiv.cuh: the two-at-a-time division of the assembly kernel against `/` on the device, bit for bit, over
    2^28 operand pairs per seed (random bit patterns, ordinary magnitudes, exponent edge cases)."""
    import ctypes as C
    c = A.Context(0)
    bad = C.c_int64(-1)
    first = (C.c_double * 4)()
    c.check(c.lib.otmb_selftest_division(c.h, 1 << 27, seed, C.byref(bad), first))
    assert bad.value == 0, f"{bad.value} quotients differ, e.g. a={first[0]!r} b={first[1]!r}: {first[2]!r} vs {first[3]!r}"


@pytest.mark.gpu
def test_staged_and_direct_uploads_give_the_same_results(monkeypatch):
    """Pageable host arrays of 4 MB and more are staged through page-locked buffers by the host pool (otmb_h2d);
    OTMB_UPLOAD_DIRECT=1 hands them to cudaMemcpyAsync as they are.  Same face fluxes, bit for bit, with array sizes
    that are not a multiple of the 16 MB piece, and with a source that is modified right after the call returns."""
    oc = synthetic.make_ocean(210, 160, 21, "tripolar", seed=4, land_frac=0.25)      # 5.6 MB per 3-D array
    f = fields(oc)
    ctx = A.Context(0)
    gm = A.makegridmetrics(areacello=f["areacello"], volcello=f["volcello"], lon=f["lon"], lat=f["lat"], lev=f["lev"],
                           lon_vertices=f["lon_vertices"], lat_vertices=f["lat_vertices"], ctx=ctx)
    ix = A.makeindices(gm.v3D, ctx=ctx)

    def run():
        F = otmb_b200.Field                                     # ordinary (pageable) numpy arrays, Float64 already
        umo = F(np.asfortranarray(oc.umo, dtype=np.float64).copy(order="F"), {"_FillValue": oc.fill})
        vmo = F(np.asfortranarray(oc.vmo, dtype=np.float64).copy(order="F"), {"_FillValue": oc.fill})
        phi = A.facefluxesfrommasstransport(umo=umo, vmo=vmo, gridmetrics=gm, indices=ix, ctx=ctx)
        umo.data[...] = -1.0                                    # the call has read its sources in full
        return [np.array(getattr(phi, k)) for k in A.FACES]

    staged = run()
    monkeypatch.setenv("OTMB_UPLOAD_DIRECT", "1")
    direct = run()
    for a, b in zip(staged, direct):
        assert np.array_equal(bits(a), bits(b))
