"""numpy composition of the reference's own pieces for BASELINE configs[2] ("C3"): the GM bolus transport added to
the mass transport before facefluxes.  TEST INFRASTRUCTURE ONLY (same rules as oracle.py).

EXTENSION, PARITY UNPINNED: the reference (v0.8.3) has `bolus_GM_velocity` (/root/reference/src/RediGM.jl:46-79),
`velocity2fluxes` (/root/reference/src/velocities.jl:10-39) and `facefluxes` (:190-255) but never chains them; the
chain and the rule for the sum (valid transport + non-NaN bolus flux; fill / NaN transports untouched) are defined
by this repository (csrc/gm.cu) and restated here from the oracle's restatements of those three functions."""
from __future__ import annotations

import numpy as np

from . import oracle as O
from . import velocities_np as V


def total_transport(umo, vmo, fill, rho3d, lon, lat, Z3D, v3D, thk, edge, topology, kGM=600.0, maxslope=0.01, rho_flux=None):
    """(umo', vmo', ϕᵢ*, ϕⱼ*): the transports facefluxes is given, and the bolus fluxes on their own."""
    u, v = O.bolus_gm(rho3d, lon, lat, Z3D, v3D, topology, kGM=kGM, maxslope=maxslope)
    gi, gj = V.velocity2fluxes(u, v, thk, edge, rho3d if rho_flux is None else rho_flux, topology)
    # the fold row: one face, two estimates of opposite sign -> antisymmetric NaN-aware mean (csrc/gm.cu, k_add_gm)
    a, b = gj[:, -1, :], gj[::-1, -1, :]
    wa, wb = ~np.isnan(a), ~np.isnan(b)
    with np.errstate(invalid="ignore", divide="ignore"):
        fold = (np.where(wa, a, 0.0) - np.where(wb, b, 0.0)) / (wa.astype(np.float64) + wb.astype(np.float64))
    gj = np.array(gj, order="F")
    gj[:, -1, :] = fold
    ok_u = ~(np.isnan(umo) | (umo == fill)) & ~np.isnan(gi)
    ok_v = ~(np.isnan(vmo) | (vmo == fill)) & ~np.isnan(gj)
    with np.errstate(invalid="ignore"):
        um = np.where(ok_u, umo + gi, umo)
        vm = np.where(ok_v, vmo + gj, vmo)
    return np.asfortranarray(um), np.asfortranarray(vm), gi, gj
