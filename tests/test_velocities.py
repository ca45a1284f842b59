"""velocity2fluxes / fluxes2velocity / facefluxesfromvelocities / B-grid interpolation (SURVEY.md §8f
ranks 1-2; /root/reference/src/velocities.jl:10-108, 140-151, src/gridcellgeometry.jl:50-140).
CPU: the numpy restatement against hand-derived values and the reference's own round-trip check
(test/local_full.jl:301-304); host-side Arakawa detection.  GPU: the CUDA kernels bit-exact against it."""
import numpy as np
import pytest

import otmb_b200
import otmb_b200.api as A
from otmb_b200 import synthetic
from oracle import oracle as O
from oracle import velocities_np as V

from _sharded_double import oracle_gridmetrics
from _util import bits, oracle_pipeline


def same(a, b):
    """Bit-identical where finite/inf; NaN matches NaN (sign and payload of a NaN carry no meaning)."""
    na, nb = np.isnan(a), np.isnan(b)
    return np.array_equal(na, nb) and np.array_equal(bits(a)[~na], bits(b)[~nb])


def test_nanmean2_nanmin2_known_answers():
    nan = np.nan
    a = np.array([1.0, nan, 3.0, nan]); b = np.array([3.0, 5.0, nan, nan])
    m = V.nanmean2(a, b)
    assert m[0] == 2.0 and m[1] == 5.0 and m[2] == 3.0 and np.isnan(m[3])
    n = V.nanmin2(a, b)
    assert n[0] == 1.0 and n[1] == 5.0 and n[2] == 3.0 and np.isnan(n[3])


def _ocean():
    oc = synthetic.make_ocean(12, 10, 6, "tripolar", seed=4, land_frac=0.25)
    o = oracle_pipeline(oc)
    return oc, o


def test_velocity_flux_round_trip_and_hand_value():
    oc, o = _ocean()
    thk, edge = o["gm"]["thkcello"], o["gm"]["edge"]
    rng = np.random.default_rng(0)
    u = np.asfortranarray(rng.normal(size=thk.shape)); v = np.asfortranarray(rng.normal(size=thk.shape))
    for rho in (1035.0, oc.rho3d):
        pi, pj = V.velocity2fluxes(u, v, thk, edge, rho, "tripolar")
        u2, v2 = V.fluxes2velocity(pi, pj, thk, edge, rho, "tripolar")
        ok = ~np.isnan(pi)
        np.testing.assert_allclose(u2[ok], u[ok], rtol=1e-14)        # the reference's own test (≈)
        ok = ~np.isnan(pj)
        np.testing.assert_allclose(v2[ok], v[ok], rtol=1e-14)
    # one value by hand: ((u*ρ)*min(thk_i, thk_east))*edge_east[i,j]
    wet = ~np.isnan(thk)
    i, j, k = np.argwhere(wet & np.roll(wet, -1, axis=0))[3]
    pi, _ = V.velocity2fluxes(u, v, thk, edge, 1035.0, "tripolar")
    want = ((u[i, j, k] * 1035.0) * min(thk[i, j, k], thk[(i + 1) % 12, j, k])) * edge[i, j, 1]
    assert pi[i, j, k] == want
    # fold: the north neighbour of (i, ny) is (nx-i+1, ny)
    _, pj = V.velocity2fluxes(u, v, thk, edge, 1035.0, "tripolar")
    ii = np.argwhere(wet[:, -1, 0] & wet[::-1, -1, 0])
    if len(ii):
        i = int(ii[0][0])
        want = ((v[i, -1, 0] * 1035.0) * min(thk[i, -1, 0], thk[11 - i, -1, 0])) * edge[i, -1, 2]
        assert pj[i, -1, 0] == want
    with pytest.raises(ValueError):
        V.velocity2fluxes(u, v, thk, edge, 1035.0, "bipolar")


def test_bgrid_to_cgrid_hand_values():
    u = np.asfortranarray(np.arange(24, dtype=float).reshape(2, 3, 4, order="F"))
    v = u + 100.0
    u[1, 1, 2] = 1e20
    u2, v2 = V.bgrid_to_cgrid(u, v, 1e20)
    assert u2[0, 0, 0] == 0.5 * u[0, 0, 0] and u2[0, 1, 0] == 0.5 * (u[0, 1, 0] + u[0, 0, 0])
    assert u2[1, 1, 2] == 0.5 * u[1, 0, 2] and u2[1, 2, 2] == 0.5 * u[1, 2, 2]          # fill -> 0
    assert v2[0, 1, 1] == 0.5 * v[0, 1, 1] and v2[1, 1, 1] == 0.5 * (v[1, 1, 1] + v[0, 1, 1])


def test_getarakawagrid_host_detection():
    oc = synthetic.make_ocean(12, 10, 4, "tripolar", seed=1)
    gm = oracle_gridmetrics(oc)
    lonv, latv = oc.lon_vertices, oc.lat_vertices
    east = (0.5 * (lonv[1] + lonv[2]), 0.5 * (latv[1] + latv[2]))
    north = (0.5 * (lonv[2] + lonv[3]), 0.5 * (latv[2] + latv[3]))
    g = A.getarakawagrid(east[0], east[1], north[0], north[1], gm)
    assert (g.kind, g.u_pos, g.v_pos) == ("C", "E", "N")
    g = A.getarakawagrid(lonv[2], latv[2], lonv[2], latv[2], gm)
    assert (g.kind, g.u_pos, g.v_pos) == ("B", "NE", "NE")
    g = A.getarakawagrid(oc.lon, oc.lat, oc.lon, oc.lat, gm)
    assert g.kind == "A"
    with pytest.raises(A.OTMBError):
        A.getarakawagrid(east[0], east[1], east[0], east[1], gm)      # u and v both on the east face


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_gpu_velocity2fluxes_fluxes2velocity_bit_exact():
    oc, o = _ocean()
    gm = oracle_gridmetrics(oc)
    thk, edge = o["gm"]["thkcello"], o["gm"]["edge"]
    rng = np.random.default_rng(1)
    u = np.asfortranarray(rng.normal(size=thk.shape)); v = np.asfortranarray(rng.normal(size=thk.shape))
    u[rng.random(thk.shape) < 0.1] = np.nan
    lonv, latv = oc.lon_vertices, oc.lat_vertices
    ulon, ulat = 0.5 * (lonv[1] + lonv[2]), 0.5 * (latv[1] + latv[2])
    vlon, vlat = 0.5 * (lonv[2] + lonv[3]), 0.5 * (latv[2] + latv[3])
    for rho in (1035.0, oc.rho3d):
        pi, pj = otmb_b200.velocity2fluxes(A.Field(u, {"_FillValue": 1e20}), ulon, ulat, A.Field(v, {"_FillValue": 1e20}),
                                           vlon, vlat, gm, rho)
        wi, wj = V.velocity2fluxes(u, v, thk, edge, rho, "tripolar")
        assert same(pi, wi) and same(pj, wj)
        u2, v2 = otmb_b200.fluxes2velocity(pi, pj, gm, rho)
        wu, wv = V.fluxes2velocity(wi, wj, thk, edge, rho, "tripolar")
        assert same(u2, wu) and same(v2, wv)


@pytest.mark.gpu
def test_gpu_bgrid_and_facefluxesfromvelocities():
    oc, o = _ocean()
    gm = oracle_gridmetrics(oc)
    thk, edge = o["gm"]["thkcello"], o["gm"]["edge"]
    rng = np.random.default_rng(2)
    fill = 1e20
    ub = np.asfortranarray(rng.normal(size=thk.shape) * 0.1); vb = np.asfortranarray(rng.normal(size=thk.shape) * 0.1)
    dry = np.isnan(o["v3D"])
    ub[dry] = fill; vb[dry] = fill
    lonv, latv = oc.lon_vertices, oc.lat_vertices
    # B-grid: velocities at the NE corner
    u2, u2lon, u2lat, v2, v2lon, v2lat = otmb_b200.interpolateontodefaultCgrid(
        A.Field(ub, {"_FillValue": fill}), lonv[2], latv[2], A.Field(vb, {"_FillValue": fill}), lonv[2], latv[2], gm)
    wu, wv = V.bgrid_to_cgrid(ub, vb, fill)
    assert np.array_equal(bits(u2), bits(wu)) and np.array_equal(bits(v2), bits(wv))
    assert u2lon.shape == oc.lon.shape and np.isfinite(u2lat).all()
    # whole chain from B-grid velocities to the six face fluxes
    ix = otmb_b200.makeindices(gm.v3D)
    phi = otmb_b200.facefluxesfromvelocities(uo=A.Field(ub, {"_FillValue": fill}), uo_lon=lonv[2], uo_lat=latv[2],
                                             vo=A.Field(vb, {"_FillValue": fill}), vo_lon=lonv[2], vo_lat=latv[2],
                                             gridmetrics=gm, indices=ix, ρ=1035.0)
    umo, vmo = V.velocity2fluxes(wu, wv, thk, edge, 1035.0, "tripolar")
    want = O.facefluxes(umo, vmo, o["v3D"], "tripolar", fill)
    for k in A.FACES:
        assert np.array_equal(bits(getattr(phi, k)), bits(want[k])), k


@pytest.mark.gpu
def test_gpu_velocity2fluxes_bipolar_errors_like_reference():
    oc = synthetic.make_ocean(10, 8, 4, "bipolar", seed=3, land_frac=0.2)
    gm = oracle_gridmetrics(oc)
    lonv, latv = oc.lon_vertices, oc.lat_vertices
    z = np.zeros(gm.v3D.shape, order="F")
    with pytest.raises(A.OTMBError):
        otmb_b200.velocity2fluxes(A.Field(z, {"_FillValue": 1e20}), 0.5 * (lonv[1] + lonv[2]), 0.5 * (latv[1] + latv[2]),
                                  A.Field(z, {"_FillValue": 1e20}), 0.5 * (lonv[2] + lonv[3]), 0.5 * (latv[2] + latv[3]), gm, 1035.0)
