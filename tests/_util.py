"""Shared helpers for the parity tests: synthetic inputs prepared the way the reference's
host code prepares them, and exact / tolerance comparisons of CSC matrices."""
import numpy as np

import otmb_b200
from otmb_b200 import synthetic
from oracle import oracle as O

NAMES = {"T": "T", "Tadv": "Tadv", "TkH": "TκH", "TkVML": "TκVML", "TkVdeep": "TκVdeep"}


def fields(oc):
    """Wrap a SyntheticOcean like the YAXArrays the reference receives."""
    F = otmb_b200.Field
    return dict(
        areacello=F(oc.areacello), volcello=F(oc.volcello), lon=oc.lon, lat=oc.lat, lev=oc.lev,
        lon_vertices=oc.lon_vertices, lat_vertices=oc.lat_vertices,
        umo=F(oc.umo, {"_FillValue": oc.fill}), vmo=F(oc.vmo, {"_FillValue": oc.fill}),
    )


def oracle_pipeline(oc, rho=1035.0, kH=500.0, kVML=0.1, kVdeep=1e-5, upwind=True):
    """Oracle end to end on a SyntheticOcean: cleaned inputs -> metrics -> fluxes -> matrices."""
    v3D = O.clean_missing(oc.volcello)
    area = O.clean_missing(oc.areacello)
    topo = O.getgridtopology(oc.lon_vertices, oc.lat_vertices)
    ix = O.makeindices(v3D)
    gm = O.gridmetrics(area, v3D, oc.lon, oc.lat, oc.lon_vertices, oc.lat_vertices, topo)
    phi = O.facefluxes(oc.umo, oc.vmo, v3D, topo, oc.fill)
    tm = O.transportmatrix(phi, oc.mlotst, v3D, gm["thkcello"], area, oc.lev, gm["edge"], gm["dnbr"], topo, rho,
                           kH=kH, kVML=kVML, kVdeep=kVdeep, upwind=upwind)
    return dict(v3D=v3D, area=area, topo=topo, ix=ix, gm=gm, phi=phi, tm=tm)


def gpu_pipeline(oc, rho=1035.0, path="fused", **kw):
    f = fields(oc)
    gm = otmb_b200.makegridmetrics(areacello=f["areacello"], volcello=f["volcello"], lon=f["lon"], lat=f["lat"],
                                   lev=f["lev"], lon_vertices=f["lon_vertices"], lat_vertices=f["lat_vertices"])
    ix = otmb_b200.makeindices(gm.v3D)
    phi = otmb_b200.facefluxesfrommasstransport(umo=f["umo"], vmo=f["vmo"], gridmetrics=gm, indices=ix)
    tm = otmb_b200.transportmatrix(ϕ=phi, mlotst=oc.mlotst, gridmetrics=gm, indices=ix, ρ=rho, path=path, **kw)
    return dict(gm=gm, ix=ix, phi=phi, tm=tm)


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.int64)


def assert_csc_equal(got, want, name="", exact=True, rtol=1e-12):
    """got: scipy csc (0-based, int64); want: oracle CSC (1-based).  The sparsity pattern must
    be bit-exact; values bit-exact (same inputs, no FMA) or within rtol (1e-12, north star)."""
    assert got.shape == (want.n, want.n), name
    assert got.indptr.dtype == np.int64 and got.indices.dtype == np.int64 and got.data.dtype == np.float64, name
    assert np.array_equal(got.indptr + 1, want.colptr), f"{name}: colptr differs"
    assert np.array_equal(got.indices + 1, want.rowval), f"{name}: rowval differs"
    if exact:
        bad = np.flatnonzero(bits(got.data) != bits(want.nzval))
        assert bad.size == 0, f"{name}: {bad.size} nzval differ bitwise, first at {bad[:5]}: " \
                              f"{got.data[bad[:5]]} vs {want.nzval[bad[:5]]}"
    else:
        np.testing.assert_allclose(got.data, want.nzval, rtol=rtol, atol=0, err_msg=name)


def transport_from_oracle_inputs(o, oc, rho=1035.0, path="fused", ctx=None, **kw):
    """GPU transportmatrix fed with the ORACLE's metrics and fluxes (identical inputs on both
    sides -> values must agree bit for bit)."""
    import otmb_b200.api as A
    gm_o = o["gm"]
    topo = A.GridTopology(o["topo"], *o["v3D"].shape)
    as_dict = lambda a: {d: np.asfortranarray(a[:, :, q]) for q, d in enumerate(A.DIRS)}
    gm = A.GridMetrics(o["area"], o["v3D"], gm_o["thkcello"], oc.lon_vertices, oc.lat_vertices, oc.lon, oc.lat,
                       gm_o["Z3D"], oc.lev, as_dict(gm_o["edge"]), as_dict(gm_o["dedge"]), as_dict(gm_o["dnbr"]), topo)
    ix = None
    phi = A.FaceFluxes(**o["phi"])
    return A.transportmatrix(ϕ=phi, mlotst=oc.mlotst, gridmetrics=gm, indices=ix, ρ=rho, path=path, ctx=ctx, **kw), gm
