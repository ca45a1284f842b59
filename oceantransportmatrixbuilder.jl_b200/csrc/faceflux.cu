// faceflux.cu — K4: nofluxboundaries! + facefluxes in one pass.
// Replaces /root/reference/src/velocities.jl:154-179 (masking), :199-200 (asserts), :203-224
// (NaN/fill -> 0, west/south shifts) and :234-243 (bottom-up continuity scan).
//
// One thread per (i,j) column, walking k from the sea floor up: the only sequential
// dependency on the whole path is phi_top[k] = ((((phi_bottom + west) + south) - east) - north),
// evaluated left to right exactly as the reference's broadcast does (adds only => bit-exact).
// Threads are consecutive in i, so every load/store of a level is a coalesced row segment;
// the west/south operands are the east/north values of the i-1 / j-1 columns re-read through
// L1 (same or adjacent cache lines).  The wet tests use the packed bit mask.
#include "common.cuh"

namespace {

__device__ __forceinline__ double clean(double x, double fill) { return (isnan(x) || x == fill) ? 0.0 : x; }

template <int UNROLL>
__global__ void __launch_bounds__(128) k_faceflux(const double* __restrict__ umo, const double* __restrict__ vmo,
                                                  const u64* __restrict__ mask, GridDims g, double fill,
                                                  double* __restrict__ east, double* __restrict__ west,
                                                  double* __restrict__ north, double* __restrict__ south,
                                                  double* __restrict__ top, double* __restrict__ bottom,
                                                  DevFlags* __restrict__ flags, int k_begin, int k_end,
                                                  const double* __restrict__ carry_in, double* __restrict__ carry_out) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    bool valid_u = false, valid_v = false;
    if (p < g.P) {
        const int i = p % g.nx, j = p / g.nx;
        const int pE = i < g.nx - 1 ? p + 1 : p - (g.nx - 1);
        const int pW = i > 0 ? p - 1 : p + (g.nx - 1);
        const int pS = j > 0 ? p - g.nx : -1;
        // north neighbour of this column, and of the column to the south (always regular: j-1 < ny-1)
        const int pN = j < g.ny - 1 ? p + g.nx : (g.topo == OTMB_TOPO_TRIPOLAR ? (g.nx - 1 - i) + g.nx * (g.ny - 1) : -1);
        // phi_top of the level below: 0 under the sea floor, or the plane handed up by the slab below
        double carry = carry_in ? carry_in[p] : 0.0;
        if (carry_in && k_end < g.nz) top[(size_t)k_end * g.P + p] = carry;   // halo level: the top flux of the cell below
        for (int k0 = k_end - 1; k0 >= k_begin; k0 -= UNROLL) {
            double e[UNROLL], w[UNROLL], n[UNROLL], s[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int k = k0 - u;
                e[u] = w[u] = n[u] = s[u] = 0.0;
                if (k < k_begin) continue;
                const int off = k * g.P;
                const int L = off + p;
                const bool wc = wet_at(mask, L);
                const bool wE = wet_at(mask, off + pE);
                const bool wW = wet_at(mask, off + pW);
                const bool wN = pN >= 0 && wet_at(mask, off + pN);
                const bool wS = pS >= 0 && wet_at(mask, off + pS);
                // nofluxboundaries!: zero at dry cells and towards dry/absent east / north neighbours
                const double ue = (wc && wE) ? __ldg(umo + L) : 0.0;
                const double vn = (wc && wN) ? __ldg(vmo + L) : 0.0;
                const double uw = (wW && wc) ? __ldg(umo + off + pW) : 0.0;
                const double vs = (wS && wc) ? __ldg(vmo + off + pS) : 0.0;   // north nbr of (i,j-1) is (i,j)
                valid_u |= !(isnan(ue) || ue == fill);
                valid_v |= !(isnan(vn) || vn == fill);
                e[u] = clean(ue, fill);
                n[u] = clean(vn, fill);
                w[u] = clean(uw, fill);
                s[u] = clean(vs, fill);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int k = k0 - u;
                if (k < k_begin) continue;
                const size_t L = (size_t)k * g.P + p;
                const double b = carry;
                const double t = (((b + w[u]) + s[u]) - e[u]) - n[u];
                east[L] = e[u];
                west[L] = w[u];
                north[L] = n[u];
                south[L] = s[u];
                bottom[L] = b;
                top[L] = t;
                carry = t;
            }
        }
        if (carry_out) carry_out[p] = carry;                                  // phi_top of the slab's first level
        if (k_begin > 0) bottom[(size_t)(k_begin - 1) * g.P + p] = carry;     // halo level: the bottom flux of the cell above
    }
    const unsigned bu = __ballot_sync(0xffffffffu, valid_u), bv = __ballot_sync(0xffffffffu, valid_v);
    if ((threadIdx.x & 31) == 0) {
        if (bu) atomicOr(&flags->any_valid_u, 1);
        if (bv) atomicOr(&flags->any_valid_v, 1);
    }
}

}  // namespace

static int facefluxes_impl(otmb_ctx* c, const double* umo, const double* vmo, double fill, const double* carry_in,
                           double* carry_out, int carry_on_device, int32_t* valid_uv, double* const outs[6]) {
    if (!c || !umo || !vmo) return OTMB_ERR_BADARG;
    OT_TRY(otmb_need(c, c->have_indices, "otmb_makeindices"));
    if (c->topo == OTMB_TOPO_UNKNOWN) return otmb_fail(c, OTMB_ERR_UNKNOWN_GRID, otmb_status_string(OTMB_ERR_UNKNOWN_GRID));
    CU_TRY(c, cudaSetDevice(c->device));
    const size_t M8 = (size_t)c->M * 8, P8 = (size_t)c->P * 8;
    OT_TRY(otmb_upload3d(c, c->stage_a, umo));
    OT_TRY(otmb_upload3d(c, c->stage_b, vmo));
    for (int q = 0; q < 6; ++q) CU_TRY(c, c->phi[q].ensure(M8));
    // carry planes: device pointers are used in place (e.g. buffers an NCCL send/recv works on)
    const double* d_in = nullptr;
    double* d_out = nullptr;
    if (carry_in) {
        if (carry_on_device)
            d_in = carry_in;
        else {
            CU_TRY(c, c->carry[0].ensure(P8));
            CU_TRY(c, cudaMemcpyAsync(c->carry[0].p, carry_in, P8, cudaMemcpyHostToDevice, c->stream));
            d_in = c->carry[0].as<double>();
        }
    }
    if (carry_out) {
        if (carry_on_device)
            d_out = carry_out;
        else {
            CU_TRY(c, c->carry[1].ensure(P8));
            d_out = c->carry[1].as<double>();
        }
    }
    OT_TRY(otmb_reset_flags(c));
    GridDims g{(int)c->nx, (int)c->ny, (int)c->nz, c->topo, (int)c->P, (int)c->M};
    k_faceflux<5><<<grid_for(c->P, 128), 128, 0, c->stream>>>(
        c->stage_a.as<double>(), c->stage_b.as<double>(), c->mask.as<u64>(), g, fill,
        c->phi[OTMB_FACE_EAST].as<double>(), c->phi[OTMB_FACE_WEST].as<double>(), c->phi[OTMB_FACE_NORTH].as<double>(),
        c->phi[OTMB_FACE_SOUTH].as<double>(), c->phi[OTMB_FACE_TOP].as<double>(), c->phi[OTMB_FACE_BOTTOM].as<double>(),
        c->flags.as<DevFlags>(), (int)c->k_own0, (int)c->k_own1, d_in, d_out);
    LAUNCHED(c);
    CU_TRY(c, cudaGetLastError());
    OT_TRY(otmb_fetch_flags(c));
    if (valid_uv) {
        // a slab only sees its own levels: the caller combines the flags of all ranks (src/velocities.jl:199-200)
        valid_uv[0] = c->h_flags->any_valid_u;
        valid_uv[1] = c->h_flags->any_valid_v;
    } else if (!c->h_flags->any_valid_u || !c->h_flags->any_valid_v) {
        c->have_phi = false;
        return otmb_fail(c, OTMB_ERR_ALL_FILL, otmb_status_string(OTMB_ERR_ALL_FILL));
    }
    if (carry_out && !carry_on_device) CU_TRY(c, cudaMemcpyAsync(carry_out, d_out, P8, cudaMemcpyDeviceToHost, c->stream));
    // results: the owned levels, written into the caller's full-size arrays
    const size_t a = (size_t)c->k_own0 * c->P, b = (size_t)c->k_own1 * c->P;
    for (int q = 0; q < 6; ++q)
        if (outs[q])
            CU_TRY(c, cudaMemcpyAsync(outs[q] + a, c->phi[q].as<double>() + a, (b - a) * 8, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    c->have_phi = true;
    return OTMB_OK;
}

extern "C" int otmb_facefluxes(otmb_ctx* c, const double* umo, const double* vmo, double fill, double* east,
                               double* west, double* north, double* south, double* top, double* bottom) {
    if (c && c->sharded) return otmb_fail(c, OTMB_ERR_STATE, "slab context: use otmb_facefluxes_slab");
    double* const outs[6] = {east, west, north, south, top, bottom};
    return facefluxes_impl(c, umo, vmo, fill, nullptr, nullptr, 0, nullptr, outs);
}

extern "C" int otmb_facefluxes_slab(otmb_ctx* c, const double* umo, const double* vmo, double fill, const double* carry_in,
                                    double* carry_out, int32_t carry_on_device, int32_t valid_uv[2], double* east,
                                    double* west, double* north, double* south, double* top, double* bottom) {
    if (!valid_uv) return OTMB_ERR_BADARG;
    double* const outs[6] = {east, west, north, south, top, bottom};
    return facefluxes_impl(c, umo, vmo, fill, carry_in, carry_out, carry_on_device, valid_uv, outs);
}
