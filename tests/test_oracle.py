"""CPU tests of the oracle itself (no GPU): known-answer cases, analytic identities, the
independent pure-Python restatement and scipy as structural cross-check.  The reference has
no golden vectors for this path (PARITY UNPINNED) — these tests are what pins
the oracle instead."""
import math

import numpy as np
import pytest
import scipy.sparse as sp

from otmb_b200 import synthetic
from oracle import oracle as O
from oracle import pyoracle as PO

from _util import oracle_pipeline

R = 6371000.0


def test_haversine_known_answers():
    assert O.haversine((0, 0), (0, 90)) == pytest.approx(R * math.pi / 2, rel=1e-15)
    assert O.haversine((0, 0), (1, 0)) == pytest.approx(R * math.pi / 180, rel=1e-14)
    assert O.haversine((10, 20), (10, 20)) == 0.0
    assert O.haversine((80, 90), (170, 90)) == 0.0            # both at the pole: cosd(90) == 0 exactly
    assert O.haversine((0, 0), (180, 0)) == pytest.approx(R * math.pi, rel=1e-15)
    assert O.haversine((350, 10), (370, 10)) == pytest.approx(O.haversine((-10, 10), (10, 10)), rel=1e-14)


def test_sind_cosd_exact_at_multiples_of_90():
    for k in range(-8, 9):
        x = 90.0 * k
        assert O.sind(x) == [0.0, 1.0, 0.0, -1.0][k % 4]
        assert O.cosd(x) == [1.0, 0.0, -1.0, 0.0][k % 4]
    assert O.sind(30.0) == pytest.approx(0.5, rel=1e-15) and O.cosd(60.0) == pytest.approx(0.5, rel=1e-15)
    xs = np.linspace(-720, 720, 2001)
    np.testing.assert_allclose([O.sind(x) for x in xs], np.sin(np.deg2rad(xs)), rtol=0, atol=2e-15)
    np.testing.assert_allclose([O.cosd(x) for x in xs], np.cos(np.deg2rad(xs)), rtol=0, atol=2e-15)


def test_sparse_known_answer():
    # duplicates are summed left to right in input order; explicit zeros are kept; rows ascend
    I = [2, 1, 2, 3, 2, 1]
    J = [1, 1, 1, 2, 1, 3]
    V = [1e16, 1.0, -1e16, 0.0, 1.0, 5.0]
    m = O.sparse(I, J, V, 3)
    assert m.colptr.tolist() == [1, 3, 4, 5]
    assert m.rowval.tolist() == [1, 2, 3, 1]
    assert m.nzval.tolist() == [1.0, (1e16 + -1e16) + 1.0, 0.0, 5.0]
    # the other association would give a different bit pattern
    m2 = O.sparse([2, 2, 2], [1, 1, 1], [1.0, 1e16, -1e16], 3)
    assert m2.nzval.tolist() == [(1.0 + 1e16) + -1e16] == [0.0]


def test_spadd_drops_zeros_only_in_sum():
    A = O.sparse([1, 2, 3], [1, 1, 2], [1.0, 2.0, 0.0], 3)
    B = O.sparse([1, 2, 3], [1, 1, 3], [-1.0, 5.0, 4.0], 3)
    C = O.spadd(A, B)
    assert C.colptr.tolist() == [1, 2, 2, 3] and C.rowval.tolist() == [2, 3] and C.nzval.tolist() == [7.0, 4.0]


def test_hand_computed_3x1x2_column():
    """Two stacked wet cells in a 3x1x2 bipolar box with hand-set metrics: every entry of every
    operator written out by hand from src/matrixbuilding.jl:193-204, 426-435."""
    nx, ny, nz = 3, 1, 2
    v3D = np.full((nx, ny, nz), np.nan, order="F")
    v3D[1, 0, 0], v3D[1, 0, 1] = 100.0, 50.0                # wet indices 1 (top) and 2 (bottom)
    thk = np.asfortranarray(v3D / 10.0)
    area = np.full((nx, ny), 10.0, order="F")
    zt = np.array([5.0, 17.5])
    edge = np.ones((nx, ny, 4), order="F")
    dnbr = np.ones((nx, ny, 4), order="F")
    z = lambda: np.zeros((nx, ny, nz), order="F")
    phi = dict(east=z(), west=z(), north=z(), south=z(), top=z(), bottom=z())
    phi["bottom"][1, 0, 0] = 8.0                              # upwelling: bottom cell -> top cell
    phi["top"][1, 0, 1] = 8.0
    ml = np.full((nx, ny), 20.0, order="F")                   # both cells in the mixed layer
    rho, kV, kML = 2.0, 1e-5, 0.1
    tm = O.transportmatrix(phi, ml, v3D, thk, area, zt, edge, dnbr, "bipolar", rho, kH=500.0, kVML=kML, kVdeep=kV)
    # Tadv: cell 1 receives from cell 2 (From Bottom): (1,2,-8/(2*100)), (2,2,+8/(2*50))
    assert tm["Tadv"].scipy().toarray().tolist() == [[0.0, -8.0 / 200.0], [0.0, 8.0 / 100.0]]
    d = abs(5.0 - 17.5)
    t1, t2 = kV * 10.0 / (d * 100.0), kV * 10.0 / (d * 50.0)
    assert tm["TkVdeep"].scipy().toarray().tolist() == [[t1, -t1], [-t2, t2]]
    m1, m2 = kML * 10.0 / (d * 100.0), kML * 10.0 / (d * 50.0)
    assert tm["TkVML"].scipy().toarray().tolist() == [[m1, -m1], [-m2, m2]]
    assert tm["TkH"].nnz == 0
    T = tm["T"].scipy().toarray()
    assert T[0, 1] == ((-8.0 / 200.0 + 0.0) + -m1) + -t1 and T[1, 1] == ((8.0 / 100.0 + 0.0) + m2) + t2
    assert T[0, 0] == ((0.0 + 0.0) + m1) + t1


@pytest.mark.parametrize("case", [(12, 10, 6, "tripolar", 0, True, False), (13, 9, 5, "tripolar", 1, False, True),
                                  (2, 4, 3, "tripolar", 2, True, False), (10, 8, 4, "bipolar", 3, True, True),
                                  (3, 3, 2, "tripolar", 4, True, False), (1, 5, 4, "bipolar", 8, True, False)])
def test_cpp_oracle_equals_python_restatement(case):
    nx, ny, nz, topo, seed, upwind, rho3 = case
    oc = synthetic.make_ocean(nx, ny, nz, topo, seed=seed, land_frac=0.2, dirty=True)
    rho = oc.rho3d if rho3 else 1035.0
    o = oracle_pipeline(oc, rho=rho, upwind=upwind)
    pm = PO.transportmatrix_py(o["phi"], oc.mlotst, o["v3D"], o["gm"]["thkcello"], o["area"], oc.lev, o["gm"]["edge"],
                               o["gm"]["dnbr"], o["topo"], rho, 500.0, 0.1, 1e-5, upwind=upwind)
    for k in O.MATS:
        a = o["tm"][k]
        cp, rv, nz_ = pm[k]
        assert np.array_equal(a.colptr, cp) and np.array_equal(a.rowval, rv), k
        assert np.array_equal(a.nzval.view(np.int64), nz_.view(np.int64)), k


def test_sparse_structure_matches_scipy():
    oc = synthetic.make_config("C1t", seed=2)
    o = oracle_pipeline(oc)
    tm = O.transportmatrix(o["phi"], oc.mlotst, o["v3D"], o["gm"]["thkcello"], o["area"], oc.lev, o["gm"]["edge"],
                           o["gm"]["dnbr"], o["topo"], 1035.0, keep_triplets=True)
    N = o["ix"]["N"]
    for k in O.MATS[1:]:
        I, J, V = tm["triplets"][k]
        S = sp.coo_matrix((V, (I - 1, J - 1)), shape=(N, N)).tocsc()
        S.sort_indices()
        assert np.array_equal(S.indptr + 1, tm[k].colptr) and np.array_equal(S.indices + 1, tm[k].rowval)
        np.testing.assert_allclose(S.data, tm[k].nzval, rtol=1e-12, atol=0)
    Ts = (tm["Tadv"].scipy() + tm["TkH"].scipy() + tm["TkVML"].scipy() + tm["TkVdeep"].scipy()).tocsc()
    Ts.sort_indices()
    np.testing.assert_allclose(Ts.toarray() if N < 3000 else Ts.data, tm["T"].scipy().toarray() if N < 3000 else tm["T"].nzval,
                               rtol=1e-12, atol=0)


@pytest.mark.parametrize("cfg", ["C1", "C1t"])
def test_reference_invariants_hold_on_oracle(cfg):
    """The reference's own property tests (test/online.jl:92-123) on the synthetic case."""
    oc = synthetic.make_config(cfg, seed=0)
    o = oracle_pipeline(oc)
    N = o["ix"]["N"]
    vflat = o["v3D"].ravel(order="F")
    v = vflat[~np.isnan(vflat)]
    one = np.ones(N)
    Myr = 365.25 * 86400 * 1e6
    for k in O.MATS:
        M = o["tm"][k].scipy()
        if k not in ("T", "Tadv"):
            assert np.linalg.norm(one) / np.linalg.norm(M @ one) / Myr > 1e6, k
        assert np.linalg.norm(v) / np.linalg.norm(M.T @ v) / Myr > 1e6, k
    T = o["tm"]["T"].scipy()
    d = T.diagonal()
    assert (d > 0).all()
    off = T - sp.diags(d)
    off.eliminate_zeros()
    assert (off.data < 0).all()
    # surface residual of the synthetic transports is at rounding level
    assert np.abs(o["phi"]["top"][:, :, 0]).max() < 1e-5 * np.abs(o["phi"]["east"]).max()


def test_makeindices_and_bitarray_layout():
    v = np.full((5, 3, 2), np.nan, order="F")
    v[0, 0, 0] = v[4, 2, 1] = v[1, 0, 0] = 1.0
    ix = O.makeindices(v)
    assert ix["N"] == 3 and ix["Lwet"].tolist() == [1, 2, 30]
    assert ix["wet_chunks"].tolist() == [(1 << 0) | (1 << 1) | (1 << 29)]
    assert ix["Lwet3D"][4, 2, 1] == 3 and ix["Lwet3D"][2, 0, 0] == 0


def test_facefluxes_known_answer():
    # 2x2x2 all wet bipolar: west/south shifts and the bottom-up continuity sum
    v = np.ones((2, 2, 2), order="F")
    umo = np.arange(1.0, 9.0).reshape((2, 2, 2), order="F")
    vmo = 10 * umo
    phi = O.facefluxes(umo, vmo, v, "bipolar", 1e20)
    assert phi["east"].ravel(order="F").tolist() == list(range(1, 9))
    assert phi["west"][:, :, 0].ravel(order="F").tolist() == [2.0, 1.0, 4.0, 3.0]       # periodic in i
    assert phi["north"][:, 1, :].tolist() == [[0.0, 0.0], [0.0, 0.0]]                     # no north neighbour: zeroed
    assert phi["south"][:, 0, :].tolist() == [[0.0, 0.0], [0.0, 0.0]]
    assert phi["south"][:, 1, 0].tolist() == [10.0, 20.0]
    assert (phi["bottom"][:, :, 1] == 0).all()
    top1 = phi["west"][:, :, 1] + phi["south"][:, :, 1] - phi["east"][:, :, 1] - phi["north"][:, :, 1]
    assert np.array_equal(phi["top"][:, :, 1], top1) and np.array_equal(phi["bottom"][:, :, 0], top1)


def test_host_helpers():
    a = O.clean_missing(np.array([[0.0, -0.0], [1e20, 3.0]]), fills=[1e20])
    assert np.isnan(a[0, 0]) and a[0, 1] == 0 and np.signbit(a[0, 1]) and np.isnan(a[1, 0]) and a[1, 1] == 3
    for topo in ("bipolar", "tripolar"):
        oc = synthetic.make_ocean(12, 8, 3, topo, seed=0)
        assert O.getgridtopology(oc.lon_vertices, oc.lat_vertices) == topo
        assert O.vertexpermutation(oc.lon_vertices, oc.lat_vertices) == [0, 1, 2, 3]
        perm = [2, 0, 3, 1]
        assert O.vertexpermutation(oc.lon_vertices[perm], oc.lat_vertices[perm]) == [perm.index(q) for q in range(4)]


@pytest.mark.parametrize("case", [(12, 10, 6, "tripolar", 0), (13, 9, 5, "tripolar", 1), (10, 8, 4, "bipolar", 3), (37, 11, 7, "tripolar", 6)])
def test_column_streamed_oracle_equals_full_oracle(case):
    """orc_tm_build_columns (the streamed form used to check grids too large for full matrices on the host) gives exactly
    the column blocks of the full build: same emitters, same `sparse`, same `+`, restricted to the triplets of a window."""
    nx, ny, nz, topo, seed = case
    oc = synthetic.make_ocean(nx, ny, nz, topo, seed=seed, land_frac=0.25)
    o = oracle_pipeline(oc)
    N = o["ix"]["N"]
    rng = np.random.default_rng(seed)
    windows = [(0, N), (0, 1), (N - 1, N)] + [tuple(sorted(rng.choice(N + 1, 2, replace=False))) for _ in range(6)]
    for upwind, rho in ((True, 1035.0), (False, oc.rho3d)):
        full = O.transportmatrix(o["phi"], oc.mlotst, o["v3D"], o["gm"]["thkcello"], o["area"], oc.lev, o["gm"]["edge"],
                                 o["gm"]["dnbr"], o["topo"], rho, upwind=upwind)
        for lo, hi in windows:
            seg = O.transportmatrix_columns(o["phi"], oc.mlotst, o["v3D"], o["gm"]["thkcello"], o["area"], oc.lev, o["gm"]["edge"],
                                            o["gm"]["dnbr"], o["topo"], rho, int(lo), int(hi), upwind=upwind)
            for name in O.MATS:
                f, s = full[name], seg[name]
                a, b = f.colptr[lo] - 1, f.colptr[hi] - 1
                assert s.n == hi - lo and np.array_equal(s.colptr, f.colptr[lo:hi + 1] - f.colptr[lo] + 1), (name, lo, hi)
                assert np.array_equal(s.rowval, f.rowval[a:b]), (name, lo, hi)
                assert np.array_equal(s.nzval.view(np.int64), f.nzval[a:b].view(np.int64)), (name, lo, hi)


def test_raw_twins_hold_every_file_the_julia_pin_script_reads(tmp_path):
    """tests/golden/pin_with_julia.jl cannot run here (no Julia).  What can be checked: `npz_to_raw.py` writes, for every
    golden case, every file the script opens, with the element count the script reshapes it to."""
    import re
    import subprocess
    import sys
    from pathlib import Path
    golden = Path(__file__).resolve().parent / "golden"
    subprocess.run([sys.executable, str(golden / "npz_to_raw.py"), str(tmp_path)], check=True, capture_output=True)
    script = (golden / "pin_with_julia.jl").read_text()
    shaped = {"f3": 3, "f2": 2, "fv": -1, "f4": -2}                       # reader -> kind of shape
    wanted = {(m.group(2) + ".f64", shaped[m.group(1)]) for m in re.finditer(r"\b(f3|f2|fv|f4)\(\"([A-Za-z0-9_]+)\"\)", script)}
    wanted |= {(m.group(1), 0) for m in re.finditer(r"joinpath\(dir, \"([A-Za-z0-9_]+\.(?:f64|i64|u64))\"\)", script)}
    for mat in re.findall(r":\w+ => \"(\w+)\"", script):                  # the five matrices' CSC fields
        wanted |= {(f"{mat}_colptr.i64", 0), (f"{mat}_rowval.i64", 0), (f"{mat}_nzval.f64", 0)}
    for name in ("name", "n"):                                            # loop variables of the script, not files
        wanted = {w for w in wanted if w[0] != name + ".f64"}
    assert len(wanted) > 30
    cases = sorted(p for p in tmp_path.iterdir() if p.is_dir())
    assert len(cases) == len(list(golden.glob("case_*.npz"))) > 0
    for case in cases:
        meta = (case / "meta.txt").read_text().split()
        assert len(meta) == 8
        nx, ny, nz = map(int, meta[:3])
        use3d = meta[7] == "1"
        for fname, kind in sorted(wanted):
            if fname == "rho3d.f64" and not use3d:
                continue
            f = case / fname
            assert f.exists(), f"{case.name}: the Julia script reads {fname}, npz_to_raw.py did not write it"
            n = f.stat().st_size // 8
            expect = {3: nx * ny * nz, 2: nx * ny, -1: 4 * nx * ny, -2: nx * ny * 4}.get(kind)
            if expect is not None:
                assert n == expect, f"{case.name}/{fname}: {n} elements, the script reshapes to {expect}"
        N = (case / "Lwet.i64").stat().st_size // 8
        for mat in ("T", "Tadv", "TkH", "TkVML", "TkVdeep"):
            assert (case / f"{mat}_colptr.i64").stat().st_size // 8 == N + 1
            assert (case / f"{mat}_rowval.i64").stat().st_size == (case / f"{mat}_nzval.f64").stat().st_size
