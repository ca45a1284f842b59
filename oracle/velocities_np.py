"""numpy restatement of the reference's velocity <-> mass-flux helpers.

TEST INFRASTRUCTURE ONLY (same rules as oracle.py); PARITY UNPINNED like the rest of the oracle —
the reference's own check of these functions is a round trip (`/root/reference/test/local_full.jl:301-304`,
`test/test_fluxes2velocity.jl:52-53`), which tests/test_velocities.py repeats.
File:line citations are relative to /root/reference.  Arrays are Fortran-ordered (nx, ny, nz).
numpy evaluates each binary operation in IEEE double without contraction, so the products below
round exactly like the reference's left-to-right `a * b * c * d`.
"""
from __future__ import annotations

import numpy as np


def _east(a):
    """x[i₊₁(𝑖)]: periodic in i (src/gridtopology.jl:57)."""
    return np.roll(a, -1, axis=0)


def _north(a, topology):
    """x[j₊₁(𝑖)]: (i, j+1), and on the tripolar fold (nx-i+1, ny) (src/gridtopology.jl:59, 94-95).
    On bipolar grids the reference gets `nothing` at j = ny and throws (src/velocities.jl:32-33)."""
    if topology != "tripolar":
        raise ValueError("the reference throws on bipolar grids (thkcello[nothing])")
    out = np.empty_like(a)
    out[:, :-1] = a[:, 1:]
    out[:, -1] = a[::-1, -1]
    return out


def nanmean2(a, b):
    """src/velocities.jl:89-93: (wa*a + wb*b)/(wa + wb) with Bool weights (false*NaN == 0.0 in Julia)."""
    wa, wb = ~np.isnan(a), ~np.isnan(b)
    with np.errstate(invalid="ignore", divide="ignore"):
        return (np.where(wa, a, 0.0) + np.where(wb, b, 0.0)) / (wa.astype(np.float64) + wb.astype(np.float64))


def nanmin2(a, b):
    """src/velocities.jl:108."""
    return np.where(np.isnan(a), b, np.where(np.isnan(b), a, np.minimum(a, b)))


def _face_factors(thk, edge, rho, topology):
    e_east, e_north = edge[:, :, 1][:, :, None], edge[:, :, 2][:, :, None]       # dirs south, east, north, west
    if np.isscalar(rho):
        r_e = r_n = np.float64(rho)                                               # twocellnanmean(x::Number) = x, :86
    else:
        r_e, r_n = nanmean2(rho, _east(rho)), nanmean2(rho, _north(rho, topology))
    t_e, t_n = nanmin2(thk, _east(thk)), nanmin2(thk, _north(thk, topology))
    return r_e, t_e, e_east, r_n, t_n, e_north


def velocity2fluxes(u, v, thk, edge, rho, topology):
    """src/velocities.jl:10-39 on a C-grid: ϕᵢ = ((u·ρ̄)·thk)·edge_east, ϕⱼ = ((v·ρ̄)·thk)·edge_north, all cells."""
    r_e, t_e, e_e, r_n, t_n, e_n = _face_factors(thk, edge, rho, topology)
    with np.errstate(invalid="ignore"):
        return np.asfortranarray(((u * r_e) * t_e) * e_e), np.asfortranarray(((v * r_n) * t_n) * e_n)


def fluxes2velocity(phi_i, phi_j, thk, edge, rho, topology):
    """src/velocities.jl:50-74: u = ϕᵢ / ((ρ̄·thk)·edge_east), v = ϕⱼ / ((ρ̄·thk)·edge_north)."""
    r_e, t_e, e_e, r_n, t_n, e_n = _face_factors(thk, edge, rho, topology)
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.asfortranarray(phi_i / ((r_e * t_e) * e_e)), np.asfortranarray(phi_j / ((r_n * t_n) * e_n))


def bgrid_to_cgrid(u, v, fill):
    """B-grid (NE) branch of interpolateontodefaultCgrid, src/gridcellgeometry.jl:123-128."""
    u2 = np.where(u == fill, 0.0, u)
    v2 = np.where(v == fill, 0.0, v)
    us = np.zeros_like(u2)
    us[:, 1:] = u2[:, :-1]
    vw = np.zeros_like(v2)
    vw[1:] = v2[:-1]
    return np.asfortranarray(0.5 * (u2 + us)), np.asfortranarray(0.5 * (v2 + vw))
