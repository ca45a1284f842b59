"""Golden fixtures (tests/golden/, made by tests/golden/make_golden.py).

known_answers.json holds HAND-DERIVED expected results (no oracle involved): they pin the
oracle (CPU tests) and the CUDA path (gpu tests) to the reference's formulas on cases small
enough to write out.  case_*.npz freeze inputs + oracle outputs on small seeded oceans: the
oracle is checked against the committed bytes (a changed oracle shows up here), and the CUDA
path is checked against the same committed bytes on the GPU box."""
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle as O
from oracle import pyoracle as PO

from _util import bits

GOLD = Path(__file__).resolve().parent / "golden"
KA = json.loads((GOLD / "known_answers.json").read_text())
CASES = sorted(p.stem[len("case_"):] for p in GOLD.glob("case_*.npz"))
MATS = ("T", "Tadv", "TkH", "TkVML", "TkVdeep")
GPU_NAMES = {"T": "T", "Tadv": "Tadv", "TkH": "TκH", "TkVML": "TκVML", "TkVdeep": "TκVdeep"}


def unit_box(shape, zt=None):
    """All-wet box with unit metrics: V = thk = area = edge = distance = 1."""
    nx, ny, nz = shape
    one3, one2 = np.ones(shape, order="F"), np.ones((nx, ny), order="F")
    four = np.ones((nx, ny, 4), order="F")
    zt = np.arange(nz) + 0.5 if zt is None else np.asarray(zt, float)
    zero = {k: np.zeros(shape, order="F") for k in O.FACES}
    return dict(v3D=one3, thk=one3.copy(order="F"), area=one2, edge=four, dnbr=four.copy(order="F"), zt=zt, phi=zero,
                mlotst=np.full((nx, ny), np.nan, order="F"))


def ring_fluxes(shape, F):
    phi = {k: np.zeros(shape, order="F") for k in O.FACES}
    phi["east"][...] = F
    phi["west"][...] = F
    return phi


def oracle_tm(b, topo, rho=1035.0, **kw):
    return O.transportmatrix(b["phi"], b["mlotst"], b["v3D"], b["thk"], b["area"], b["zt"], b["edge"], b["dnbr"], topo, rho, **kw)


def gpu_tm(b, topo, rho=1035.0, path="fused", **kw):
    import otmb_b200.api as A
    shape = b["v3D"].shape
    as_dict = lambda a: {d: np.asfortranarray(a[:, :, q]) for q, d in enumerate(A.DIRS)}
    nan2 = np.full(shape[:2], np.nan, order="F")
    nanv = np.full((4,) + shape[:2], np.nan, order="F")
    gm = A.GridMetrics(b["area"], b["v3D"], b["thk"], nanv, nanv, nan2, nan2, np.zeros(shape, order="F"), b["zt"],
                       as_dict(b["edge"]), as_dict(b["edge"]), as_dict(b["dnbr"]), A.GridTopology(topo, *shape))
    names = {"kH": "κH", "kVML": "κVML", "kVdeep": "κVdeep", "upwind": "upwind"}
    return A.transportmatrix(ϕ=A.FaceFluxes(**b["phi"]), mlotst=b["mlotst"], gridmetrics=gm, indices=None, ρ=rho, path=path,
                             **{names[k]: v for k, v in kw.items()})


# ------------------------------------------------------------------------------------------ CPU: oracle vs hand-derived
def test_known_haversine():
    for c in KA["haversine"]:
        assert O.haversine(c["p"], c["q"]) == pytest.approx(c["d"], rel=2e-15, abs=0.0 if c["d"] else 1e-300)


def test_known_kvdeep_and_kh_unit_box():
    for key, op, kw in (("kvdeep_4x2x2", "TkVdeep", "kVdeep"), ("kh_4x2x2", "TkH", "kH")):
        c = KA[key]
        b = unit_box(c["shape"])
        tm = oracle_tm(b, c["topology"], **{kw: c[kw]})
        assert np.array_equal(tm[op].scipy().toarray(), np.array(c["dense"])), key
        # rows ascending and unique inside every column (SparseMatrixCSC invariant)
        m = tm[op]
        for j in range(m.n):
            r = m.rowval[m.colptr[j] - 1:m.colptr[j + 1] - 1]
            assert (np.diff(r) > 0).all()


def test_known_adv_ring():
    c = KA["adv_ring4"]
    b = unit_box(c["shape"])
    b["v3D"] = np.full(c["shape"], c["V"], order="F")
    b["phi"] = ring_fluxes(c["shape"], c["F"])
    tm = oracle_tm(b, c["topology"], rho=c["rho"], kH=0.0, kVML=0.0, kVdeep=0.0, upwind=True)
    assert np.array_equal(tm["Tadv"].scipy().toarray(), np.array(c["upwind_dense"]))
    assert np.array_equal(tm["T"].scipy().toarray(), np.array(c["upwind_dense"]))
    tc = oracle_tm(b, c["topology"], rho=c["rho"], kH=0.0, kVML=0.0, kVdeep=0.0, upwind=False)
    assert np.array_equal(tc["Tadv"].scipy().toarray(), np.array(c["centred_dense"]))
    # sparse keeps the cancelled diagonal as an explicit zero, + drops it from T
    assert tc["Tadv"].nnz == 12 and tc["T"].nnz == 8


def test_known_sparse_and_spadd():
    c = KA["sparse"]
    m = O.sparse(c["I"], c["J"], c["V"], c["n"])
    assert m.colptr.tolist() == c["colptr"] and m.rowval.tolist() == c["rowval"] and m.nzval.tolist() == c["nzval"]
    c = KA["spadd"]
    s = O.spadd(O.sparse(c["A"]["I"], c["A"]["J"], c["A"]["V"], c["n"]), O.sparse(c["B"]["I"], c["B"]["J"], c["B"]["V"], c["n"]))
    assert s.colptr.tolist() == c["colptr"] and s.rowval.tolist() == c["rowval"] and s.nzval.tolist() == c["nzval"]


def test_known_answers_hold_for_the_independent_restatement():
    c = KA["kh_4x2x2"]
    b = unit_box(c["shape"])
    got = PO.transportmatrix_py(b["phi"], b["mlotst"], b["v3D"], b["thk"], b["area"], b["zt"], b["edge"], b["dnbr"],
                                c["topology"], 1035.0, c["kH"], 0.1, 1e-5)
    cp, rv, nz = got["TkH"]
    assert np.array_equal(O.CSC(got["N"], np.asarray(cp), np.asarray(rv), np.asarray(nz, float)).scipy().toarray(),
                          np.array(c["dense"]))


# ------------------------------------------------------------------------------------------ CPU: oracle vs frozen bytes
def _load(name):
    return np.load(GOLD / f"case_{name}.npz")


def _oracle_on_case(z):
    v3D, area = O.clean_missing(z["volcello"]), O.clean_missing(z["areacello"])
    topo = str(z["topology"])
    ix = O.makeindices(v3D)
    gm = O.gridmetrics(area, v3D, z["lon"], z["lat"], z["lon_vertices"], z["lat_vertices"], topo)
    phi = O.facefluxes(z["umo"], z["vmo"], v3D, topo, float(z["fill"]))
    rho = z["rho3d"] if bool(z["use_rho3d"]) else float(z["rho_scalar"])
    tm = O.transportmatrix(phi, z["mlotst"], v3D, gm["thkcello"], area, z["lev"], gm["edge"], gm["dnbr"], topo, rho,
                           upwind=bool(z["upwind"]))
    return v3D, area, topo, ix, gm, phi, tm


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_frozen_fixture(name):
    z = _load(name)
    v3D, area, topo, ix, gm, phi, tm = _oracle_on_case(z)
    assert O.getgridtopology(z["lon_vertices"], z["lat_vertices"]) == topo
    assert ix["N"] == int(z["N"]) and np.array_equal(ix["Lwet"], z["Lwet"]) and np.array_equal(ix["wet_chunks"], z["wet_chunks"])
    for k in ("thkcello", "Z3D", "edge", "dedge", "dnbr"):
        assert np.array_equal(bits(gm[k]), bits(z[k])), k
    for k in O.FACES:
        assert np.array_equal(bits(phi[k]), bits(z["phi_" + k])), k
    for m in MATS:
        assert np.array_equal(tm[m].colptr, z[m + "_colptr"]) and np.array_equal(tm[m].rowval, z[m + "_rowval"]), m
        assert np.array_equal(bits(tm[m].nzval), bits(z[m + "_nzval"])), m


# ------------------------------------------------------------------------------------------ GPU: CUDA path vs the same fixtures
@pytest.mark.gpu
@pytest.mark.parametrize("path", ["fused", "fused2", "coo"])
def test_gpu_known_answers(path):
    c = KA["kvdeep_4x2x2"]
    tm = gpu_tm(unit_box(c["shape"]), c["topology"], path=path, kVdeep=c["kVdeep"])
    assert np.array_equal(tm.TκVdeep.toarray(), np.array(c["dense"]))
    c = KA["kh_4x2x2"]
    tm = gpu_tm(unit_box(c["shape"]), c["topology"], path=path, kH=c["kH"])
    assert np.array_equal(tm.TκH.toarray(), np.array(c["dense"]))
    c = KA["adv_ring4"]
    b = unit_box(c["shape"])
    b["v3D"] = np.full(c["shape"], c["V"], order="F")
    b["phi"] = ring_fluxes(c["shape"], c["F"])
    tm = gpu_tm(b, c["topology"], rho=c["rho"], path=path, kH=0.0, kVML=0.0, kVdeep=0.0, upwind=True)
    assert np.array_equal(tm.Tadv.toarray(), np.array(c["upwind_dense"])) and np.array_equal(tm.T.toarray(), np.array(c["upwind_dense"]))
    tm = gpu_tm(b, c["topology"], rho=c["rho"], path=path, kH=0.0, kVML=0.0, kVdeep=0.0, upwind=False)
    assert np.array_equal(tm.Tadv.toarray(), np.array(c["centred_dense"]))
    assert tm.Tadv.nnz == 12 and tm.T.nnz == 8


@pytest.mark.gpu
def test_gpu_known_sparse():
    import otmb_b200
    c = KA["sparse"]
    cp, rv, nz = otmb_b200.sparse(c["I"], c["J"], c["V"], c["n"])
    assert cp.tolist() == c["colptr"] and rv.tolist() == c["rowval"] and nz.tolist() == c["nzval"]
    c = KA["spadd"]
    A_ = otmb_b200.sparse(c["A"]["I"], c["A"]["J"], c["A"]["V"], c["n"])
    B_ = otmb_b200.sparse(c["B"]["I"], c["B"]["J"], c["B"]["V"], c["n"])
    cp, rv, nz = otmb_b200.spadd(A_, B_, c["n"])
    assert cp.tolist() == c["colptr"] and rv.tolist() == c["rowval"] and nz.tolist() == c["nzval"]


@pytest.mark.gpu
@pytest.mark.parametrize("path", ["fused", "coo"])
@pytest.mark.parametrize("name", CASES)
def test_gpu_reproduces_frozen_fixture(name, path):
    """Whole GPU pipeline from the committed raw inputs against the committed outputs: index maps,
    face fluxes, thkcello/Z3D and the matrices' structure bit-exact; haversine-derived values and
    therefore TκH / T within 1e-12 (CUDA vs glibc libm), everything else bit-exact."""
    import otmb_b200
    import otmb_b200.api as A
    z = _load(name)
    F = otmb_b200.Field
    fill = float(z["fill"])
    gm = otmb_b200.makegridmetrics(areacello=F(z["areacello"]), volcello=F(z["volcello"]), lon=z["lon"], lat=z["lat"],
                                   lev=z["lev"], lon_vertices=z["lon_vertices"], lat_vertices=z["lat_vertices"])
    assert gm.gridtopology.kind == str(z["topology"])
    ix = otmb_b200.makeindices(gm.v3D)
    assert ix.N == int(z["N"]) and np.array_equal(ix.Lwet, z["Lwet"])
    assert np.array_equal(bits(gm.thkcello), bits(z["thkcello"])) and np.array_equal(bits(gm.Z3D), bits(z["Z3D"]))
    for q, d in enumerate(A.DIRS):
        np.testing.assert_allclose(gm.edge_length_2D[d], z["edge"][:, :, q], rtol=1e-13, atol=0, equal_nan=True)
        np.testing.assert_allclose(gm.distance_to_neighbour_2D[d], z["dnbr"][:, :, q], rtol=1e-13, atol=0, equal_nan=True)
    phi = otmb_b200.facefluxesfrommasstransport(umo=F(z["umo"], {"_FillValue": fill}), vmo=F(z["vmo"], {"_FillValue": fill}),
                                                gridmetrics=gm, indices=ix)
    for k in A.FACES:
        assert np.array_equal(bits(getattr(phi, k)), bits(z["phi_" + k])), k
    rho = z["rho3d"] if bool(z["use_rho3d"]) else float(z["rho_scalar"])
    tm = otmb_b200.transportmatrix(ϕ=phi, mlotst=z["mlotst"], gridmetrics=gm, indices=ix, ρ=rho, upwind=bool(z["upwind"]), path=path)
    for m in MATS:
        g = getattr(tm, GPU_NAMES[m])
        assert np.array_equal(g.indptr + 1, z[m + "_colptr"]) and np.array_equal(g.indices + 1, z[m + "_rowval"]), m
        if m in ("Tadv", "TkVML", "TkVdeep"):
            assert np.array_equal(bits(g.data), bits(z[m + "_nzval"])), m
        else:
            np.testing.assert_allclose(g.data, z[m + "_nzval"], rtol=1e-12, atol=0, err_msg=m)
