"""k-slab sharding of one matrix: host logic on CPU (plan, offsets, carry chain,
concatenation — threads and a world_size-2 gloo process group, compute by the oracle test double)
and the CUDA slab path on one GPU (ranks emulated as threads, one context each)."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import otmb_b200  # noqa: F401
import otmb_b200.api as A
from otmb_b200 import _lib as _L
from otmb_b200 import sharded, synthetic
from oracle import oracle as O

from _sharded_double import NAMES, OracleSlab, oracle_gridmetrics
from _util import assert_csc_equal, oracle_pipeline

ROOT = Path(__file__).resolve().parent.parent


def test_plan_slabs_properties():
    rng = np.random.default_rng(0)
    for nx, ny, nz in ((5, 4, 1), (7, 3, 2), (6, 5, 5), (9, 4, 20)):
        v = np.asfortranarray(np.where(rng.random((nx, ny, nz)) < 0.6, 1.0, np.nan))
        wet_rows = (~np.isnan(v)).sum(axis=0).ravel(order="F")            # wet cells per grid row R = j + ny*k
        for levels in (False, True):
            units = nz if levels else ny * nz
            for R in range(1, min(units, 9) + 1):
                slabs, wet = sharded.plan_slabs(v, R, level_cuts_only=levels)
                assert len(slabs) == R and slabs[0][0] == 0 and slabs[-1][1] == ny * nz
                assert all(a < b for a, b in slabs) and all(slabs[r][1] == slabs[r + 1][0] for r in range(R - 1))
                assert wet == [int(wet_rows[a:b].sum()) for a, b in slabs] and sum(wet) == int(wet_rows.sum())
                if levels:
                    assert all(a % ny == 0 and b % ny == 0 for a, b in slabs)
            with pytest.raises(ValueError):
                sharded.plan_slabs(v, units + 1, level_cuts_only=levels)
    # a realistic mask: row cuts balance to within one grid row, level cuts are much coarser
    oc = synthetic.make_ocean(90, 45, 20, "tripolar", seed=1)
    v = O.clean_missing(oc.volcello)
    nx = v.shape[0]
    for R in (2, 4, 8):
        _, wet = sharded.plan_slabs(v, R)
        assert max(abs(w - sum(wet) / R) for w in wet) <= nx
        _, wet_lv = sharded.plan_slabs(v, R, level_cuts_only=True)
        assert max(wet_lv) >= max(wet)


def test_column_levels_tile_the_grid():
    ny, nz = 7, 5
    cuts = [0, 3, 9, 10, 24, 35]
    seen = np.zeros((ny, nz), int)
    for a, b in zip(cuts[:-1], cuts[1:]):
        kb, ke = sharded.column_levels((a, b), ny)
        for j in range(ny):
            for k in range(kb[j], ke[j]):
                assert a <= j + ny * k < b
                seen[j, k] += 1
    assert (seen == 1).all()


def _run_double(oc, R, **kw):
    gm = oracle_gridmetrics(oc)
    fn = lambda ex: sharded.transportmatrix_sharded(exchange=ex, gridmetrics=gm, mlotst=oc.mlotst, ρ=1035.0, umo=oc.umo,
                                                    vmo=oc.vmo, FillValue=oc.fill, slab_factory=OracleSlab, **kw)
    return sharded.run_threaded(R, fn)


@pytest.mark.parametrize("levels", [False, True])
@pytest.mark.parametrize("R", [1, 2, 3, 5])
def test_sharded_host_logic_threads(R, levels):
    oc = synthetic.make_ocean(12, 10, 6, "tripolar", seed=0, land_frac=0.25)
    o = oracle_pipeline(oc)
    res = _run_double(oc, R, level_cuts_only=levels)
    full, segs, info = res[0]
    assert all(r[0] is None for r in res[1:])
    for oname, gname in {v: k for k, v in NAMES.items()}.items():
        assert_csc_equal(getattr(full, gname), o["tm"][oname], f"R={R} {oname}", exact=True)
    if not levels and R > 1:
        assert any(a % 10 for a, _ in info["slabs"]), "row cuts should fall inside a level on this mask"
    # segments tile the column range and carry global entry offsets
    col = 0
    for r in range(R):
        s = res[r][1]["T"]
        assert s.col0 == col and s.N == o["ix"]["N"]
        col += s.ncols
        if r + 1 < R:
            assert s.colptr[-1] == res[r + 1][1]["T"].colptr[0]
    assert col == o["ix"]["N"]


def test_sharded_gloo_world_size_2():
    """The same driver over torch.distributed (gloo, 2 processes on 127.0.0.1)."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), str(ROOT / "tests" / "_sharded_worker.py")],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "SHARDED-GLOO-OK ranks=2" in r.stdout


# ------------------------------------------------------------------------------------------ CUDA slab path
@pytest.mark.gpu
@pytest.mark.parametrize("levels", [False, True])
@pytest.mark.parametrize("R", [2, 3, 7])
@pytest.mark.parametrize("case", [(12, 10, 6, "tripolar", 0), (10, 8, 4, "bipolar", 3), (90, 45, 20, "tripolar", 5)])
def test_sharded_cuda_slabs_match_oracle(case, R, levels):
    nx, ny, nz, topo, seed = case
    if levels and R > nz:
        pytest.skip("more ranks than levels")
    oc = synthetic.make_ocean(nx, ny, nz, topo, seed=seed, land_frac=0.25)
    o = oracle_pipeline(oc)
    gm = oracle_gridmetrics(oc)
    for upwind, rho in ((True, 1035.0), (False, oc.rho3d)):
        want = O.transportmatrix(o["phi"], oc.mlotst, o["v3D"], o["gm"]["thkcello"], o["area"], oc.lev, o["gm"]["edge"],
                                 o["gm"]["dnbr"], o["topo"], rho, upwind=upwind)
        fn = lambda ex: sharded.transportmatrix_sharded(exchange=ex, gridmetrics=gm, mlotst=oc.mlotst, ρ=rho, umo=oc.umo,
                                                        vmo=oc.vmo, FillValue=oc.fill, upwind=upwind, level_cuts_only=levels)
        full, segs, info = sharded.run_threaded(R, fn)[0]
        assert info["N"] == o["ix"]["N"]
        for gname, oname in NAMES.items():
            assert_csc_equal(getattr(full, gname), want[oname], f"{case} R={R} {oname} upwind={upwind}", exact=True)


@pytest.mark.gpu
def test_sharded_cuda_with_precomputed_phi():
    oc = synthetic.make_ocean(20, 14, 8, "tripolar", seed=7, land_frac=0.2)
    o = oracle_pipeline(oc)
    gm = oracle_gridmetrics(oc)
    fn = lambda ex: sharded.transportmatrix_sharded(exchange=ex, gridmetrics=gm, mlotst=oc.mlotst, ρ=1035.0, ϕ=o["phi"])
    full, _, _ = sharded.run_threaded(4, fn)[0]
    for gname, oname in NAMES.items():
        assert_csc_equal(getattr(full, gname), o["tm"][oname], oname, exact=True)


@pytest.mark.gpu
def test_large_grid_sharded_equals_unsharded_and_invariants():
    """A grid too large for the oracle to be the checker in reasonable time (720x540x25, ~5 M wet cells): the
    size-independent properties — CSC structure (monotone colptr, strictly ascending rows per column), the
    reference's sign pattern and conservation invariants (test/online.jl:110-123), and bit-identity of the
    k-slab sharded assembly (4 slabs, one context each) with the single-context assembly."""
    import otmb_b200.api as A
    from _util import fields
    oc = synthetic.make_ocean(720, 540, 25, "tripolar", seed=11)
    f = fields(oc)
    ctx = A.Context(0)
    gm = A.makegridmetrics(areacello=f["areacello"], volcello=f["volcello"], lon=f["lon"], lat=f["lat"], lev=f["lev"],
                           lon_vertices=f["lon_vertices"], lat_vertices=f["lat_vertices"], ctx=ctx)
    ix = A.makeindices(gm.v3D, ctx=ctx)
    phi = A.facefluxesfrommasstransport(umo=f["umo"], vmo=f["vmo"], gridmetrics=gm, indices=ix, ctx=ctx)
    tm = A.transportmatrix(ϕ=phi, mlotst=oc.mlotst, gridmetrics=gm, indices=ix, ρ=1035.0, ctx=ctx)
    N = ix.N
    assert N > 3_000_000
    for name in A.MATRICES:
        m = getattr(tm, name)
        assert m.shape == (N, N) and m.indptr[0] == 0 and m.indptr[-1] == m.nnz
        assert (np.diff(m.indptr) >= 0).all()
        d = np.diff(m.indices)
        starts = m.indptr[1:-1][np.diff(m.indptr)[:-1] > 0]           # positions where a new column begins
        inner = np.ones(m.nnz - 1, bool)
        inner[starts[(starts > 0) & (starts < m.nnz)] - 1] = False
        assert (d[inner] > 0).all(), f"{name}: rows not strictly ascending inside a column"
        assert m.indices.min() >= 0 and m.indices.max() < N and np.isfinite(m.data).all()
    T = tm.T
    one, v = np.ones(N), gm.v3D.ravel(order="F")[~np.isnan(gm.v3D.ravel(order="F"))]
    Myr = 365.25 * 86400 * 1e6
    for name in ("TκH", "TκVML", "TκVdeep"):
        assert np.linalg.norm(one) / np.linalg.norm(getattr(tm, name) @ one) / Myr > 1e6, name
    for name in A.MATRICES:
        assert np.linalg.norm(v) / np.linalg.norm(getattr(tm, name).T @ v) / Myr > 1e6, name
    assert (T.diagonal() > 0).all()
    off = T.copy()
    off.setdiag(0)
    off.eliminate_zeros()
    assert (off.data < 0).all()
    ctx.close()
    # the same matrix from four slabs (threads, one context each, all on this GPU)
    fn = lambda ex: sharded.transportmatrix_sharded(exchange=ex, gridmetrics=gm, mlotst=oc.mlotst, ρ=1035.0, umo=oc.umo,
                                                    vmo=oc.vmo, FillValue=oc.fill)
    full, segs, info = sharded.run_threaded(4, fn)[0]
    assert info["N"] == N and len(info["slabs"]) == 4
    for name in A.MATRICES:
        a, b = getattr(tm, name), getattr(full, name)
        assert np.array_equal(a.indptr, b.indptr) and np.array_equal(a.indices, b.indices), name
        assert np.array_equal(a.data.view(np.int64), b.data.view(np.int64)), name


# ------------------------------------------------------------------------------------------ errors are collective
class _FailingSlab(OracleSlab):
    """Test double whose rank-local step raises on ONE rank (the slab that owns grid row `bad_row`)."""
    bad_row, where = 0, "build"

    def _mine(self):
        return self.rows[0] <= self.bad_row < self.rows[1]

    def facefluxes(self, *a, **kw):
        if self.where == "fluxes" and self._mine():
            raise A.OTMBError(_L.ERR_CUDA, "injected failure in facefluxes")
        if self.where == "fluxes":           # the plane from the failed rank below is a dummy: do not check it
            return True, True
        return super().facefluxes(*a, **kw)

    def transportmatrix(self, *a, **kw):
        if self.where == "build" and self._mine():
            raise A.OTMBError(_L.ERR_TADV_NAN, "Tadv contains NaNs.")
        return super().transportmatrix(*a, **kw)


@pytest.mark.parametrize("where,bad_row,code", [("build", 0, 1), ("build", 59, 1), ("fluxes", 59, 100), ("fluxes", 30, 100)])
def test_rank_local_errors_are_raised_on_every_rank(where, bad_row, code):
    """One slab fails (NaN check, CUDA error): no rank may be left blocking in the next collective or in the
    carry chain's recv — every rank raises, the failing one its own error, the others one that names it."""
    oc = synthetic.make_ocean(12, 10, 6, "tripolar", seed=0, land_frac=0.25)
    gm = oracle_gridmetrics(oc)
    factory = type("F", (_FailingSlab,), dict(bad_row=bad_row, where=where))
    raised = []

    def fn(ex):
        try:
            sharded.transportmatrix_sharded(exchange=ex, gridmetrics=gm, mlotst=oc.mlotst, ρ=1035.0, umo=oc.umo, vmo=oc.vmo,
                                            FillValue=oc.fill, slab_factory=factory)
        except A.OTMBError as e:
            raised.append((ex.rank, e.code, str(e)))
            return "raised"
        return "completed"

    assert sharded.run_threaded(3, fn) == ["raised"] * 3
    assert sorted(r for r, _, _ in raised) == [0, 1, 2] and {c for _, c, _ in raised} == {code}
    assert sum("rank " in m for _, _, m in raised) == 2


@pytest.mark.gpu
@pytest.mark.parametrize("nchunks", [1, 3, 8])
def test_native_driver_single_rank_and_chunked_columns(nchunks):
    """The in-library sharded driver (csrc/comm.cu) with one rank (no NCCL is loaded): makeindices, the face-flux
    kernel launched over column chunks of the plane — the unit the carry chain pipelines — and the assembly must
    reproduce the oracle bit for bit, whatever the number of chunks."""
    oc = synthetic.make_ocean(37, 11, 7, "tripolar", seed=6, land_frac=0.25, dirty=True)
    o = oracle_pipeline(oc)
    gm = oracle_gridmetrics(oc)
    ns = sharded.NativeSharded(gridmetrics=gm, rank=0, nranks=1, id_bytes=None)
    assert (ns.N, ns.w0) == (o["ix"]["N"], 0)
    ns.set_masstransport(oc.umo, oc.vmo, oc.fill)
    outs = [np.full(gm.v3D.shape, np.nan, order="F") for _ in range(6)]
    ns.facefluxes(nchunks, outs)
    for k, a in zip(A.FACES, outs):
        assert np.array_equal(a.view(np.int64), o["phi"][k].view(np.int64)), k
    ns.build(oc.mlotst, 1035.0)
    assert ns.nnz_before == [0] * 5 and ns.nnz_total == ns.slab.nnz
    segs = ns.segments()
    for gname, oname in NAMES.items():
        sgm = segs[gname]
        assert_csc_equal(A._csc(ns.N, sgm.colptr, sgm.rowval, sgm.nzval), o["tm"][oname], oname, exact=True)
    # the device-resident form of the chain (what a per-month pipeline enqueues) gives the same matrix
    ns.facefluxes_enqueue(nchunks)
    assert ns.build(oc.mlotst, 1035.0, upload=False) == ns.nnz_total
    ns.close()


@pytest.mark.gpu
def test_c4_quarter_degree_sharded_matches_streamed_oracle():
    """BASELINE configs[3] at FULL size (1440x1080x50, 40.4 M wet cells, 280 M entries in T): the row-slab sharded
    assembly (8 slabs, cuts inside levels, one context each on this GPU) against the oracle, bit for bit.  The oracle
    cannot hold the five full matrices beside the inputs, so it is streamed (orc_tm_build_columns): three windows of
    columns — the first million (surface, mixed layer), a million straddling the cut between ranks 3 and 4 (a cut
    inside a level: halo rows, carry plane, rank offsets), and the last million (sea floor).  Both sides get the same
    geometry (the GPU's) and the same umo / vmo; the oracle computes its own face fluxes."""
    import otmb_b200.api as A
    from _util import fields
    R = 8
    oc = synthetic.make_config("C4", seed=0)
    f = fields(oc)
    ctx = A.Context(0)
    gm = A.makegridmetrics(areacello=f["areacello"], volcello=f["volcello"], lon=f["lon"], lat=f["lat"], lev=f["lev"],
                           lon_vertices=f["lon_vertices"], lat_vertices=f["lat_vertices"], ctx=ctx)
    ctx.close()
    slabs, wet = sharded.plan_slabs(gm.v3D, R)
    N = sum(wet)
    col0 = np.concatenate([[0], np.cumsum(wet)])
    assert max(abs(w - N / R) for w in wet) <= gm.v3D.shape[0] and any(a % gm.v3D.shape[1] for a, _ in slabs[1:])
    windows = [(0, 1_000_000), (int(col0[4]) - 500_000, int(col0[4]) + 500_000), (N - 1_000_000, N)]
    need = {r for lo, hi in windows for r in range(R) if col0[r] < hi and col0[r + 1] > lo}

    def fn(ex):
        slab, w0, n, info = sharded.prepare_sharded(exchange=ex, gridmetrics=gm, umo=oc.umo, vmo=oc.vmo, FillValue=oc.fill)
        nnz = slab.build(oc.mlotst, 1035.0, 500.0, 0.1, 1.0e-5, True)
        nnz_all = ex.allgather_ints(nnz)
        local = slab.fetch() if ex.rank in need else None
        slab.ctx.close()
        return w0, n, nnz_all, local

    res = sharded.run_threaded(R, fn)
    assert [r[0] for r in res] == [int(c) for c in col0[:-1]] and all(r[1] == N for r in res)
    nnz_all = res[0][2]
    phi = O.facefluxes(oc.umo, oc.vmo, gm.v3D, oc.topology, oc.fill)
    edge = np.asfortranarray(np.stack([gm.edge_length_2D[d] for d in A.DIRS], axis=-1))
    dnbr = np.asfortranarray(np.stack([gm.distance_to_neighbour_2D[d] for d in A.DIRS], axis=-1))
    for lo, hi in windows:
        want = O.transportmatrix_columns(phi, oc.mlotst, gm.v3D, gm.thkcello, gm.area2D, oc.lev, edge, dnbr, oc.topology, 1035.0, lo, hi)
        for m, (gname, oname) in enumerate(NAMES.items()):
            w = want[oname]
            # the GPU's columns lo..hi from the rank segments that hold them (local colptr + entries of the lower ranks)
            cps, rvs, nzs = [], [], []
            for r in sorted(need):
                a, b = max(lo, int(col0[r])) - int(col0[r]), min(hi, int(col0[r + 1])) - int(col0[r])
                if a >= b:
                    continue
                cp, rv, nz = res[r][3][gname]
                off = sum(nnz_all[q][m] for q in range(r))
                cps.append(cp[a:b] + off)
                rvs.append(rv[cp[a]:cp[b]])
                nzs.append(nz[cp[a]:cp[b]])
                last = cp[b] + off
            cp = np.concatenate(cps + [np.array([last])])
            assert np.array_equal(cp - cp[0] + 1, w.colptr), f"{gname} colptr, columns {lo}:{hi}"
            assert np.array_equal(np.concatenate(rvs) + 1, w.rowval), f"{gname} rowval, columns {lo}:{hi}"
            assert np.array_equal(np.concatenate(nzs).view(np.int64), w.nzval.view(np.int64)), f"{gname} nzval, columns {lo}:{hi}"


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(12))
def test_fuzz_row_slabs_bit_exact(seed):
    """Random small grids cut into a random number of row slabs (cuts inside levels; slabs thinner than a level; windows
    that reach the first / last level): the concatenated matrices equal the oracle's bit for bit."""
    rng = np.random.default_rng(2000 + seed)
    nx, ny, nz = int(rng.integers(2, 13)), int(rng.integers(2, 9)), int(rng.integers(1, 7))
    topo = "tripolar" if rng.random() < 0.6 else "bipolar"
    oc = synthetic.make_ocean(nx, ny, nz, topo, seed=100 + seed, land_frac=float(rng.uniform(0.0, 0.4)), dirty=bool(rng.random() < 0.3))
    R = int(rng.integers(2, min(9, ny * nz) + 1))
    upwind = bool(rng.random() < 0.6)
    rho = 1035.0 if rng.random() < 0.5 else oc.rho3d
    o = oracle_pipeline(oc, rho=rho, upwind=upwind)
    gm = oracle_gridmetrics(oc)
    fn = lambda ex: sharded.transportmatrix_sharded(exchange=ex, gridmetrics=gm, mlotst=oc.mlotst, ρ=rho, umo=oc.umo, vmo=oc.vmo,
                                                    FillValue=oc.fill, upwind=upwind)
    full, segs, info = sharded.run_threaded(R, fn)[0]
    for gname, oname in NAMES.items():
        assert_csc_equal(getattr(full, gname), o["tm"][oname], f"seed {seed} {(nx, ny, nz, topo)} R={R} {oname}", exact=True)
