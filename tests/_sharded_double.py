"""Test double for the slab compute (tests only): the CPU oracle stands in for libotmb.so so that
the HOST logic of the sharded driver — slab plan, wet-rank offsets, the bottom-up carry chain,
nnz offsets, concatenation — can run without a GPU (threads here, gloo processes in
tests/_sharded_worker.py).  It runs the full oracle on every rank and hands out the rank's slice."""
import numpy as np

from oracle import oracle as O

NAMES = {"T": "T", "Tadv": "Tadv", "TκH": "TkH", "TκVML": "TkVML", "TκVdeep": "TkVdeep"}


class OracleSlab:
    def __init__(self, shape, topology, row0, row1, device=0):
        self.shape, self.topology, self.rows = shape, topology, (row0, row1)
        nx, ny, nz = shape
        self.L0, self.L1 = row0 * nx, row1 * nx          # owned linear cells
        self.phi = None
        self.log = []

    def makeindices(self, v3D):
        self.v3D = np.asfortranarray(v3D)
        wet = ~np.isnan(self.v3D.ravel(order="F"))
        P = self.shape[0] * self.shape[1]
        self.n_owned = int(wet[self.L0:self.L1].sum())
        self.h_up = int(wet[max(0, self.L0 - P):self.L0].sum())
        self.below = int(wet[:self.L0].sum())
        return self.n_owned, self.h_up

    def set_rank_offset(self, w0):
        assert w0 == self.below, "global wet-rank offset must equal the wet cells before the slab"
        self.w0 = w0

    def set_metrics(self, gm):
        self.gm = gm

    def facefluxes(self, umo, vmo, fill, carry_in, carry_out, outputs=None):
        from otmb_b200.sharded import column_levels
        nx, ny, nz = self.shape
        full = O.facefluxes(umo, vmo, self.v3D, self.topology, fill)
        kb, ke = column_levels(self.rows, ny)                    # owned levels per j
        top = np.concatenate([full["top"], np.zeros((nx, ny, 1))], axis=2)      # level nz: 0 under the sea floor
        jj = np.arange(ny)
        if self.L1 < nx * ny * nz:
            assert carry_in is not None
            got = np.asarray(carry_in[3]).reshape(nx, ny, order="F")
            want = top[:, jj, ke]                                 # ϕtop of the cell below the last owned one
            assert np.array_equal(got.view(np.int64), want.view(np.int64)), "carry from the slab below"
        else:
            assert carry_in is None
        if self.L0 > 0:
            carry_out[3][...] = top[:, jj, kb].ravel(order="F")   # ϕtop of the first owned level (passed on if none)
        else:
            assert carry_out is None
        self.phi = full
        own = np.zeros(nx * ny * nz, bool)
        own[self.L0:self.L1] = True
        own = own.reshape(self.shape, order="F")
        vu = bool((~(np.isnan(umo) | (umo == fill)) & own).any())
        vv = bool((~(np.isnan(vmo) | (vmo == fill)) & own).any())
        return vu, vv

    def set_facefluxes(self, phi):
        self.phi = {k: np.asfortranarray(getattr(phi, k) if not isinstance(phi, dict) else phi[k]) for k in O.FACES}

    def transportmatrix(self, mlotst, rho, kH, kVML, kVdeep, upwind):
        gm = self.gm
        stack = lambda d: np.asfortranarray(np.stack([d[k] for k in O.DIRS], axis=-1))
        tm = O.transportmatrix(self.phi, mlotst, self.v3D, gm.thkcello, gm.area2D, gm.zt, stack(gm.edge_length_2D),
                               stack(gm.distance_to_neighbour_2D), self.topology, rho, kH=kH, kVML=kVML, kVdeep=kVdeep,
                               upwind=upwind)
        out = {}
        a, b = self.w0, self.w0 + self.n_owned
        for name, oname in NAMES.items():
            m = tm[oname]
            cp = m.colptr[a:b + 1] - 1
            out[name] = ((cp - cp[0]).astype(np.int64), (m.rowval[cp[0]:cp[-1]] - 1).astype(np.int64),
                         m.nzval[cp[0]:cp[-1]].copy())
        return out


def oracle_gridmetrics(oc):
    """GridMetrics namedtuple filled by the oracle (host arrays, no GPU)."""
    import otmb_b200.api as A
    v3D, area = O.clean_missing(oc.volcello), O.clean_missing(oc.areacello)
    topo = O.getgridtopology(oc.lon_vertices, oc.lat_vertices)
    g = O.gridmetrics(area, v3D, oc.lon, oc.lat, oc.lon_vertices, oc.lat_vertices, topo)
    as_dict = lambda a: {d: np.asfortranarray(a[:, :, q]) for q, d in enumerate(A.DIRS)}
    return A.GridMetrics(area, v3D, g["thkcello"], oc.lon_vertices, oc.lat_vertices, oc.lon, oc.lat, g["Z3D"], oc.lev,
                         as_dict(g["edge"]), as_dict(g["dedge"]), as_dict(g["dnbr"]), A.GridTopology(topo, *v3D.shape))
