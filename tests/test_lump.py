"""lump_and_spray (/root/reference/src/extratools.jl:38-112): the Python restatement
against hand-derived values and operator identities on CPU; the CUDA path bit-exact against it on the GPU."""
import numpy as np
import pytest
import scipy.sparse as sp

import otmb_b200
import otmb_b200.api as A
from otmb_b200 import synthetic
from oracle import lump_np

from _util import bits, oracle_pipeline


def _case(shape=(12, 10, 6), topo="tripolar", seed=0):
    oc = synthetic.make_ocean(*shape, topo, seed=seed, land_frac=0.3)
    o = oracle_pipeline(oc)
    wet = ~np.isnan(o["v3D"])
    vol = o["v3D"].ravel(order="F")[wet.ravel(order="F")]
    return wet, vol, o["tm"]["T"].scipy()


def test_hand_derived_2x2_box():
    # 2x2x1 all wet; T connects 0-1 and 2-3 only (two components in the one lumping box)
    wet = np.ones((2, 2, 1), bool)
    vol = np.array([1.0, 3.0, 2.0, 6.0])
    T = sp.csc_matrix(np.array([[1, 1, 0, 0], [1, 1, 0, 0], [0, 0, 1, 1], [0, 0, 1, 1]], float))
    L, S, vc = lump_np.lump_and_spray(wet, vol, T)
    assert vc.tolist() == [4.0, 8.0]
    assert np.array_equal(L.toarray(), np.array([[(1 / 4.0) * 1.0, (1 / 4.0) * 3.0, 0, 0], [0, 0, (1 / 8.0) * 2.0, (1 / 8.0) * 6.0]]))
    assert np.array_equal(S.toarray(), np.array([[1, 0], [1, 0], [0, 1], [0, 1.0]]))
    # fully connected: one coarse cell; an isolated dry cell does not count
    T = sp.csc_matrix(np.ones((4, 4)))
    L, S, vc = lump_np.lump_and_spray(wet, vol, T)
    assert vc.tolist() == [12.0] and L.shape == (1, 4)
    wet[1, 1, 0] = False
    L, S, vc = lump_np.lump_and_spray(wet, vol[:3], sp.csc_matrix(np.ones((3, 3))))
    assert L.shape == (1, 3) and vc.tolist() == [6.0]


@pytest.mark.parametrize("d", [(2, 2, 1), (3, 2, 2), (1, 1, 1), (5, 3, 2)])
def test_operator_identities(d):
    wet, vol, T = _case()
    L, S, vc = lump_np.lump_and_spray(wet, vol, T, di=d[0], dj=d[1], dk=d[2])
    N = len(vol)
    assert L.shape == (len(vc), N) and S.shape == (N, len(vc)) and L.nnz == S.nnz == N
    np.testing.assert_allclose(L @ np.ones(N), 1.0, rtol=1e-14)                 # a volume-weighted average
    np.testing.assert_allclose((L @ S).toarray(), np.eye(len(vc)), atol=1e-14)  # SPRAY then LUMP is the identity
    np.testing.assert_allclose(vc @ L.toarray(), vol, rtol=1e-14)               # volume conserving
    if d == (1, 1, 1):
        assert len(vc) == N and np.array_equal(vc, vol)
    Tc = (L @ T @ S).toarray()                                                  # the coarsened operator keeps mass conservation:
    assert (np.abs(vc @ Tc) <= 1e-9 * (vc @ np.abs(Tc))).all()                  # v_c' T_c = v' T SPRAY ~ 0 (cancellation to rounding)


def test_custom_mask_semantics():
    wet, vol, T = _case((8, 6, 3))
    mask = np.zeros(wet.shape, bool)
    mask[4:] = True                                   # lump only the eastern half
    L, S, vc = lump_np.lump_and_spray(wet, vol, T, mask)
    flat = -np.ones(wet.size, int)
    flat[wet.ravel(order="F")] = np.arange(len(vol))
    rank = flat.reshape(wet.shape, order="F")
    west = rank[:4][wet[:4]]
    Ld = L.toarray() != 0
    assert (Ld.sum(axis=0) == 1).all()                              # every fine cell belongs to exactly one coarse cell
    rows_of_west = Ld[:, west].argmax(axis=0)
    assert (Ld[rows_of_west].sum(axis=1) == 1).all()                # cells outside the mask stay on their own
    assert len(set(rows_of_west)) == len(west) and L.shape[0] < len(vol)


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("d", [(2, 2, 1), (3, 2, 2), (1, 1, 1), (4, 4, 2)])
@pytest.mark.parametrize("case", [((12, 10, 6), "tripolar", 0), ((37, 11, 7), "tripolar", 6), ((90, 45, 20), "bipolar", 2)])
def test_gpu_lump_and_spray_bit_exact(case, d):
    wet, vol, T = _case(*case)
    LUMP, SPRAY, vol_c = otmb_b200.lump_and_spray(wet, vol, T, di=d[0], dj=d[1], dk=d[2])
    L, S, vc = lump_np.lump_and_spray(wet, vol, T, di=d[0], dj=d[1], dk=d[2])
    assert np.array_equal(bits(vol_c), bits(vc))
    for g, w in ((LUMP, L), (SPRAY, S)):
        assert g.shape == w.shape and np.array_equal(g.indptr, w.indptr) and np.array_equal(g.indices, w.indices)
        assert np.array_equal(bits(g.data), bits(w.data))


@pytest.mark.gpu
def test_gpu_lump_errors():
    wet, vol, T = _case()
    with pytest.raises(A.OTMBError):
        otmb_b200.lump_and_spray(wet, vol, T, np.zeros(wet.shape, bool))       # custom mask
    with pytest.raises(A.OTMBError):
        otmb_b200.lump_and_spray(wet, vol, T, di=4, dj=4, dk=4)                # box > 32 cells
    Tasym = sp.csc_matrix(sp.triu(T))                                            # SimpleGraph would throw
    with pytest.raises(A.OTMBError):
        otmb_b200.lump_and_spray(wet, vol, Tasym)
    # a caller-supplied pattern that is not an N x N CSC is refused by the library, like a pre-built operator
    import ctypes as C
    ctx = otmb_b200.default_context()
    otmb_b200.lump_and_spray(wet, vol, T)                                        # leaves this grid's indices resident
    Tc = sp.csc_matrix(T); Tc.sort_indices()
    N = Tc.shape[0]
    cp, rv = Tc.indptr.astype(np.int64), Tc.indices.astype(np.int64)
    volc = np.ascontiguousarray(vol, dtype=np.float64)

    def build(cp, rv):
        Nc = C.c_int64()
        return ctx.lib.otmb_lump_and_spray_build(ctx.h, 2, 2, 1, A._ptr(volc), A._ptr(cp), A._ptr(rv), 0, 0, C.byref(Nc))

    assert build(cp, rv) == 0
    bad = cp.copy(); bad[-1] = -3
    assert build(bad, rv) == otmb_b200._lib.ERR_BADARG
    bad = cp.copy(); bad[4], bad[5] = bad[5], bad[4] - 1                          # not monotone
    assert build(bad, rv) == otmb_b200._lib.ERR_BADARG
    bad = rv.copy(); bad[2] = N + 5                                               # row outside the matrix
    assert build(cp, bad) == otmb_b200._lib.ERR_BADARG
    assert build(cp, rv) == 0                                                     # and the context still works


# ------------------------------------------------------------------------------------------ T_c = LUMP * T * SPRAY
def test_spmatmul_restatement_against_scipy_and_by_hand():
    """The Gustavson restatement (accumulation order of SparseArrays' product) agrees with scipy's product to rounding
    and keeps structural zeros; a 2x2 example by hand."""
    A_ = sp.csc_matrix(np.array([[1.0, 2.0], [0.0, 3.0]]))
    B_ = sp.csc_matrix(np.array([[4.0, 0.0], [5.0, 6.0]]))
    Cm = lump_np.spmatmul(A_, B_)
    assert np.array_equal(Cm.toarray(), np.array([[14.0, 12.0], [15.0, 18.0]]))
    Z = lump_np.spmatmul(sp.csc_matrix(np.array([[1.0, -1.0]])), sp.csc_matrix(np.array([[2.0], [2.0]])))
    assert Z.nnz == 1 and Z.data[0] == 0.0                                     # 2 - 2: the zero is stored
    wet, vol, T = _case((12, 10, 6), "tripolar", 0)
    L, S, _ = lump_np.lump_and_spray(wet, vol, T)
    Tc = lump_np.coarsen(L, T, S)
    ref = sp.csc_matrix(L @ T @ S)
    ref.sort_indices()
    assert Tc.shape == ref.shape and np.array_equal(Tc.indptr, ref.indptr) and np.array_equal(Tc.indices, ref.indices)
    np.testing.assert_allclose(Tc.data, ref.data, rtol=1e-13, atol=1e-25)
    # volume conservation carries over to the coarse operator: v_cᵀ T_c = (vᵀ T) SPRAY
    _, _, vc = lump_np.lump_and_spray(wet, vol, T)
    np.testing.assert_allclose(Tc.T @ vc, S.T @ (T.T @ vol), rtol=0, atol=1e-9 * np.abs(Tc.diagonal() * vc).max())


@pytest.mark.gpu
@pytest.mark.parametrize("d", [(2, 2, 1), (3, 2, 2), (4, 4, 2)])
@pytest.mark.parametrize("name", ["T", "TκH"])
def test_gpu_coarse_operator_bit_exact(d, name):
    """T_c = LUMP * T * SPRAY on the device (resident T, resident LUMP / SPRAY) against the restated SparseArrays product."""
    from _util import gpu_pipeline
    oc = synthetic.make_ocean(37, 11, 7, "tripolar", seed=6, land_frac=0.25)
    g = gpu_pipeline(oc)
    wet = g["ix"].wet3D
    vol = g["gm"].v3D.ravel(order="F")[wet.ravel(order="F")]
    X = getattr(g["tm"], name)
    LUMP, SPRAY, vol_c = otmb_b200.lump_and_spray(wet, vol, g["tm"].T, di=d[0], dj=d[1], dk=d[2])
    got = otmb_b200.coarsen(name)
    want = lump_np.coarsen(LUMP, sp.csc_matrix(X), SPRAY)
    assert got.shape == want.shape and np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
    assert np.array_equal(bits(got.data), bits(want.data))
