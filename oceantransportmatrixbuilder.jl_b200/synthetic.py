"""Seeded synthetic ocean grids, bathymetry masks and divergence-free mass transports.

The reference ships no fixtures (its only CI test downloads CMIP6 data,
/root/reference/test/online.jl:19-65), so every parity case is generated here
(SURVEY.md §8d).  The generator is plain numpy, deterministic in `seed`, and
produces arrays in the reference's layout: Julia column-major `(nx, ny, nz)`,
i.e. Fortran-ordered numpy arrays (i fastest, k slowest), and `(4, nx, ny)`
vertex arrays in the reference's default vertex order SW, SE, NE, NW
(/root/reference/src/gridcellgeometry.jl:146-157) so that `vertexpermutation`
returns the identity.

Grids
  * bipolar  : regular lon/lat; the top vertex row sits at lat == 90 exactly so
               `getgridtopology` (/root/reference/src/gridtopology.jl:41) says Bipolar.
  * tripolar : regular up to a join latitude, then a cap whose top vertex row is
               folded, `P[i, ny] == P[nx - i, ny]` bit for bit
               (/root/reference/src/gridtopology.jl:44, 87-95).

Fluxes: `umo = d_j psi + d_k chi_u`, `vmo = -d_i psi + d_k chi_v` with psi on
corner points (zero on corners touching land, mirrored on the fold) and chi on
level interfaces (zero at the surface and at/below the face's sea floor), so the
transports are non-divergent in exact arithmetic, never cross a land face, and
carry zero net transport through every face column.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

R_EARTH = 6371000.0
FILL = 1.0e20


@dataclass
class SyntheticOcean:
    nx: int
    ny: int
    nz: int
    topology: str
    seed: int
    lon: np.ndarray            # (nx, ny)  F
    lat: np.ndarray            # (nx, ny)  F
    lon_vertices: np.ndarray   # (4, nx, ny) F
    lat_vertices: np.ndarray   # (4, nx, ny) F
    lev: np.ndarray            # (nz,)
    areacello: np.ndarray      # (nx, ny)  F, 0 on land columns
    volcello: np.ndarray       # (nx, ny, nz) F, 0 on dry cells
    umo: np.ndarray            # (nx, ny, nz) F, FILL on dry cells
    vmo: np.ndarray            # (nx, ny, nz) F, FILL on dry cells
    mlotst: np.ndarray         # (nx, ny)  F, NaN on land
    rho3d: np.ndarray          # (nx, ny, nz) F, NaN on dry cells
    kbot: np.ndarray           # (nx, ny) int, number of wet levels per column
    fill: float = FILL
    meta: dict = field(default_factory=dict)

    def dump(self, directory) -> None:
        """Raw little-endian Float64 dumps so any implementation (incl. a future
        Julia run of the reference) can read identical bytes."""
        d = Path(directory)
        d.mkdir(parents=True, exist_ok=True)
        for name in ("lon", "lat", "lon_vertices", "lat_vertices", "lev", "areacello",
                     "volcello", "umo", "vmo", "mlotst", "rho3d"):
            a = np.asarray(getattr(self, name), dtype="<f8")
            a.ravel(order="F").tofile(d / f"{name}.f64")
        (d / "shape.txt").write_text(f"{self.nx} {self.ny} {self.nz} {self.topology} {self.seed}\n")


def _xyz(lon_deg, lat_deg):
    lo = np.deg2rad(lon_deg)
    la = np.deg2rad(lat_deg)
    return np.stack([np.cos(la) * np.cos(lo), np.cos(la) * np.sin(lo), np.sin(la)], axis=-1)


def _lonlat(xyz, lon_near):
    n = np.linalg.norm(xyz, axis=-1, keepdims=True)
    u = xyz / n
    lat = np.rad2deg(np.arcsin(np.clip(u[..., 2], -1.0, 1.0)))
    lon = np.rad2deg(np.arctan2(u[..., 1], u[..., 0]))
    lon = lon + 360.0 * np.round((lon_near - lon) / 360.0)
    return lon, lat


def _corner_points(nx, ny, topology):
    """Corner lon/lat, shape (ny+1, nx+1), index [j, i]."""
    i = np.arange(nx + 1)
    lonc1 = 80.0 + 360.0 * i / nx
    if topology == "bipolar":
        latc1 = np.linspace(-78.0, 90.0, ny + 1)
        latc1[-1] = 90.0
        lonc = np.broadcast_to(lonc1, (ny + 1, nx + 1)).copy()
        latc = np.broadcast_to(latc1[:, None], (ny + 1, nx + 1)).copy()
        return lonc, latc
    if topology != "tripolar":
        raise ValueError(topology)
    jcap = min(max(1, int(round(0.78 * ny))), ny - 1)
    lat_join, latp = 60.0, 75.0
    lonc = np.empty((ny + 1, nx + 1))
    latc = np.empty((ny + 1, nx + 1))
    latreg = np.linspace(-78.0, lat_join, jcap + 1)
    lonc[: jcap + 1] = lonc1
    latc[: jcap + 1] = latreg[:, None]
    # fold line: a segment through the geographic pole between two displaced poles
    s = np.minimum(i, nx - i)
    t = s / (nx / 2.0)
    theta = (1.0 - 2.0 * t) * (90.0 - latp)
    flon = np.where(theta >= 0, 80.0, 260.0)
    flat = 90.0 - np.abs(theta)
    ring = _xyz(lonc1, np.full(nx + 1, lat_join))
    fold = _xyz(flon, flat)
    for j in range(jcap + 1, ny):
        w = (j - jcap) / (ny - jcap)
        p = (1.0 - w) * ring + w * fold
        lo, la = _lonlat(p, lonc1)
        lo[nx] = lo[0] + 360.0
        la[nx] = la[0]
        lonc[j], latc[j] = lo, la
    lonc[ny], latc[ny] = flon, flat          # bit-identical at i and nx - i
    return lonc, latc


def _levels(nz, total=5500.0, dz0=10.0):
    lo, hi = 1.0, 3.0
    for _ in range(200):
        r = 0.5 * (lo + hi)
        tot = dz0 * nz if abs(r - 1) < 1e-14 else dz0 * (r ** nz - 1) / (r - 1)
        if tot > total:
            hi = r
        else:
            lo = r
    dz = dz0 * r ** np.arange(nz)
    zbot = np.cumsum(dz)
    return dz, zbot - 0.5 * dz, zbot


def _smooth(rng, nx, ny, nterms=12):
    x = 2 * np.pi * (np.arange(nx) + 0.5) / nx
    y = (np.arange(ny) + 0.5) / ny
    f = np.zeros((ny, nx))
    for _ in range(nterms):
        kx = rng.integers(0, 5)
        ky = rng.integers(0, 4)
        a = rng.normal() / (1.0 + kx + ky)
        f += a * np.cos(kx * x[None, :] + rng.uniform(0, 2 * np.pi)) * np.cos(
            np.pi * ky * y[:, None] + rng.uniform(0, 2 * np.pi))
    f -= f.min()
    m = f.max()
    return f / m if m > 0 else f


def make_ocean(nx, ny, nz, topology="tripolar", seed=0, land_frac=0.32, float32_roundtrip=False,
               dirty=False, force_fold_wet=True, flux_scale=1.0e8,
               allow_self_neighbour=False) -> SyntheticOcean:
    rng = np.random.Generator(np.random.PCG64(seed))
    lonc, latc = _corner_points(nx, ny, topology)

    # ---- cells: vertices SW, SE, NE, NW; centres = normalised mean of the corners
    vlon = np.stack([lonc[:-1, :-1], lonc[:-1, 1:], lonc[1:, 1:], lonc[1:, :-1]], axis=-1)  # (ny,nx,4)
    vlat = np.stack([latc[:-1, :-1], latc[:-1, 1:], latc[1:, 1:], latc[1:, :-1]], axis=-1)
    cxyz = _xyz(vlon, vlat)                                   # (ny,nx,4,3)
    lon_near = 80.0 + 360.0 * (np.arange(nx) + 0.5) / nx
    lon_c, lat_c = _lonlat(cxyz.mean(axis=2), lon_near[None, :])
    t1 = np.cross(cxyz[:, :, 1] - cxyz[:, :, 0], cxyz[:, :, 2] - cxyz[:, :, 0])
    t2 = np.cross(cxyz[:, :, 2] - cxyz[:, :, 0], cxyz[:, :, 3] - cxyz[:, :, 0])
    area = 0.5 * (np.linalg.norm(t1, axis=-1) + np.linalg.norm(t2, axis=-1)) * R_EARTH ** 2

    # ---- vertical grid and bathymetry
    dz, zt, zbot = _levels(nz)
    ztop = zbot - dz
    H = _smooth(rng, nx, ny)
    thr = np.quantile(H, land_frac)
    d = np.clip((H - thr) / max(1e-12, 1.0 - thr), 0.0, 1.0)
    depth = 150.0 + d ** 0.7 * (zbot[-1] - 150.0)
    kbot = (ztop[None, None, :] < depth[:, :, None]).sum(axis=-1)
    kbot = np.clip(kbot, min(3, nz), nz)
    kbot[H < thr] = 0
    kbot[0, :] = 0                                            # southern row: land
    if topology == "tripolar" and force_fold_wet:
        # make the fold row and both coincidence points wet (seam i=1/nx; centre nx/2, nx/2+1)
        cols = {0, 1 % nx, nx - 1, (nx - 2) % nx, nx // 2 - 1, nx // 2, (nx // 2 + 1) % nx, max(nx // 2 - 2, 0)}
        for c in cols:
            kbot[ny - 1, c] = max(kbot[ny - 1, c], min(3, nz))
            if ny >= 2:
                kbot[ny - 2, c] = max(kbot[ny - 2, c], min(3, nz))
    if topology == "tripolar" and nx % 2 == 1 and not allow_self_neighbour:
        # odd nx: cell ((nx+1)/2, ny) is its own north neighbour (src/gridtopology.jl:94) with a
        # zero-length north edge and zero neighbour distance -> the reference itself stops with
        # "TκH contains NaNs."; keep that column dry unless a test asks for it
        kbot[ny - 1, (nx - 1) // 2] = 0
    wet = np.arange(nz)[:, None, None] < kbot[None, :, :]     # (nz,ny,nx)

    pf = rng.uniform(0.25, 1.0, size=(ny, nx))
    vol = area[None] * dz[:, None, None] * np.ones((nz, 1, 1))
    bottom = np.arange(nz)[:, None, None] == (kbot[None] - 1)
    vol = np.where(bottom, vol * pf[None], vol)
    vol = np.where(wet, vol, 0.0)
    area_out = np.where(kbot > 0, area, 0.0)

    # ---- stream function on corners (nz, ny+1, nx+1)
    wl = np.roll(wet, 1, axis=2)                              # cell to the west (periodic)
    cw = np.zeros((nz, ny + 1, nx), dtype=bool)               # corner (j, i) for i = 0..nx-1; cells (i-1, i)
    cw[:, 1:ny, :] = wet[:, :-1, :] & wl[:, :-1, :] & wet[:, 1:, :] & wl[:, 1:, :]
    if topology == "tripolar":
        top = wet[:, ny - 1, :] & wl[:, ny - 1, :]            # cells (i-1, i) at the top row, corner i
        # corner nx - i touches cells nx-i-1, nx-i
        mir = top[:, (nx - np.arange(nx)) % nx]
        cw[:, ny, :] = top & mir
    psi = np.zeros((nz, ny + 1, nx + 1))
    base = rng.normal(size=(nz, ny + 1, nx)) * flux_scale
    base += flux_scale * 3.0 * _smooth(rng, nx, ny + 1)[None] * np.linspace(1.0, 0.1, nz)[:, None, None]
    if topology == "tripolar":
        s = np.minimum(np.arange(nx), (nx - np.arange(nx)) % nx)
        base[:, ny, :] = base[:, ny, s]
    psi[:, :, :nx] = np.where(cw, base, 0.0)
    psi[:, :, nx] = psi[:, :, 0]
    umo = psi[:, 1:, 1:] - psi[:, :-1, 1:]
    vmo = -(psi[:, 1:, 1:] - psi[:, 1:, :-1])

    # ---- vertical-shear part: chi on interfaces (nz+1, ny, nx)
    kb_e = np.minimum(kbot, np.roll(kbot, -1, axis=1))
    kb_n = np.zeros_like(kbot)
    kb_n[:-1] = np.minimum(kbot[:-1], kbot[1:])
    if topology == "tripolar":
        kb_n[ny - 1] = np.minimum(kbot[ny - 1], kbot[ny - 1, ::-1])
    kk = np.arange(nz + 1)[:, None, None]
    chiu = rng.normal(size=(nz + 1, ny, nx)) * 0.3 * flux_scale
    chiv = rng.normal(size=(nz + 1, ny, nx)) * 0.3 * flux_scale
    if topology == "tripolar":
        chiv[:, ny - 1, :] = 0.5 * (chiv[:, ny - 1, :] - chiv[:, ny - 1, ::-1])   # antisymmetric on the fold
    chiu = np.where((kk == 0) | (kk >= kb_e[None]), 0.0, chiu)
    chiv = np.where((kk == 0) | (kk >= kb_n[None]), 0.0, chiv)
    umo = umo + (chiu[:-1] - chiu[1:])
    vmo = vmo + (chiv[:-1] - chiv[1:])

    if dirty:
        # garbage the reference's nofluxboundaries! (/root/reference/src/velocities.jl:154-179) must remove
        g = rng.random(size=umo.shape) < 0.02
        umo = np.where(g & ~np.roll(wet, -1, axis=2), rng.normal(size=umo.shape) * flux_scale, umo)
        g = rng.random(size=vmo.shape) < 0.02
        wn = np.zeros_like(wet)
        wn[:, :-1] = wet[:, 1:]
        vmo = np.where(g & ~wn, rng.normal(size=vmo.shape) * flux_scale, vmo)
    if float32_roundtrip:
        umo = umo.astype(np.float32).astype(np.float64)
        vmo = vmo.astype(np.float32).astype(np.float64)
    fillv = float(np.float32(FILL)) if float32_roundtrip else FILL
    umo = np.where(wet, umo, fillv)
    vmo = np.where(wet, vmo, fillv)
    if dirty:
        g = (rng.random(size=umo.shape) < 0.05) & ~wet
        umo = np.where(g, np.nan, umo)
        vmo = np.where(g, np.nan, vmo)

    ml = 10.0 + 790.0 * _smooth(rng, nx, ny) ** 2
    ml = np.where(kbot > 0, ml, np.nan)
    rho = 1025.0 + 3.0 * _smooth(rng, nx, ny)[None] + 0.004 * zt[:, None, None] \
        + 0.05 * rng.normal(size=(nz, ny, nx))
    rho = np.where(wet, rho, np.nan)

    F = np.asfortranarray
    return SyntheticOcean(
        nx=nx, ny=ny, nz=nz, topology=topology, seed=seed,
        lon=F(lon_c.T), lat=F(lat_c.T),
        lon_vertices=F(vlon.transpose(2, 1, 0)), lat_vertices=F(vlat.transpose(2, 1, 0)),
        lev=zt.copy(), areacello=F(area_out.T), volcello=F(vol.transpose(2, 1, 0)),
        umo=F(umo.transpose(2, 1, 0)), vmo=F(vmo.transpose(2, 1, 0)),
        mlotst=F(ml.T), rho3d=F(rho.transpose(2, 1, 0)), kbot=F(kbot.T), fill=fillv,
        meta=dict(land_frac=land_frac, dirty=dirty, float32_roundtrip=float32_roundtrip),
    )


# the named configurations of BASELINE.json
CONFIGS = {
    "C1": dict(nx=90, ny=45, nz=20, topology="bipolar"),
    "C1t": dict(nx=90, ny=45, nz=20, topology="tripolar"),
    "C2": dict(nx=360, ny=300, nz=50, topology="tripolar"),
    "C4": dict(nx=1440, ny=1080, nz=50, topology="tripolar"),
}


def make_config(name, seed=0, **kw) -> SyntheticOcean:
    cfg = dict(CONFIGS[name])
    cfg.update(kw)
    return make_ocean(seed=seed, **cfg)
