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
from otmb_b200 import sharded, synthetic
from oracle import oracle as O

from _sharded_double import NAMES, OracleSlab, oracle_gridmetrics
from _util import assert_csc_equal, oracle_pipeline

ROOT = Path(__file__).resolve().parent.parent


def test_plan_slabs_properties():
    rng = np.random.default_rng(0)
    for nz in (1, 2, 5, 20, 50):
        wet = rng.integers(0, 1000, nz)
        for R in range(1, min(nz, 9) + 1):
            slabs = sharded.plan_slabs(wet, R)
            assert len(slabs) == R and slabs[0][0] == 0 and slabs[-1][1] == nz
            assert all(a < b for a, b in slabs) and all(slabs[r][1] == slabs[r + 1][0] for r in range(R - 1))
    # balanced on a surface-heavy profile (upper levels are wetter, like a real ocean)
    wet = np.linspace(1000, 100, 50).astype(int)
    slabs = sharded.plan_slabs(wet, 8)
    per = [wet[a:b].sum() for a, b in slabs]
    assert max(per) / (sum(per) / 8) < 1.25
    with pytest.raises(ValueError):
        sharded.plan_slabs([1, 2], 3)


def _run_double(oc, R, **kw):
    gm = oracle_gridmetrics(oc)
    fn = lambda ex: sharded.transportmatrix_sharded(exchange=ex, gridmetrics=gm, mlotst=oc.mlotst, ρ=1035.0, umo=oc.umo,
                                                    vmo=oc.vmo, FillValue=oc.fill, slab_factory=OracleSlab, **kw)
    return sharded.run_threaded(R, fn)


@pytest.mark.parametrize("R", [1, 2, 3, 5])
def test_sharded_host_logic_threads(R):
    oc = synthetic.make_ocean(12, 10, 6, "tripolar", seed=0, land_frac=0.25)
    o = oracle_pipeline(oc)
    res = _run_double(oc, R)
    full, segs, info = res[0]
    assert all(r[0] is None for r in res[1:])
    for oname, gname in {v: k for k, v in NAMES.items()}.items():
        assert_csc_equal(getattr(full, gname), o["tm"][oname], f"R={R} {oname}", exact=True)
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
@pytest.mark.parametrize("R", [2, 3])
@pytest.mark.parametrize("case", [(12, 10, 6, "tripolar", 0), (10, 8, 4, "bipolar", 3), (90, 45, 20, "tripolar", 5)])
def test_sharded_cuda_slabs_match_oracle(case, R):
    nx, ny, nz, topo, seed = case
    oc = synthetic.make_ocean(nx, ny, nz, topo, seed=seed, land_frac=0.25)
    o = oracle_pipeline(oc)
    gm = oracle_gridmetrics(oc)
    for upwind, rho in ((True, 1035.0), (False, oc.rho3d)):
        want = O.transportmatrix(o["phi"], oc.mlotst, o["v3D"], o["gm"]["thkcello"], o["area"], oc.lev, o["gm"]["edge"],
                                 o["gm"]["dnbr"], o["topo"], rho, upwind=upwind)
        fn = lambda ex: sharded.transportmatrix_sharded(exchange=ex, gridmetrics=gm, mlotst=oc.mlotst, ρ=rho, umo=oc.umo,
                                                        vmo=oc.vmo, FillValue=oc.fill, upwind=upwind)
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
