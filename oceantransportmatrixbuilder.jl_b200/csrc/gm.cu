// gm.cu — BASELINE configs[2] ("C3"): the Gent-McWilliams bolus transport folded into the advective mass fluxes.
//
// EXTENSION, PARITY UNPINNED.  The reference (v0.8.3) has the pieces but no such wiring: `bolus_GM_velocity`
// (/root/reference/src/RediGM.jl:46-79) returns velocities, `velocity2fluxes` (/root/reference/src/velocities.jl:10-39)
// turns velocities into face mass fluxes, and `transportmatrix` has no Redi/GM keyword.  SURVEY.md §8a (scope note
// (i)) names the principled route, which this file implements entirely on the device:
//     (u*, v*)   = bolus_GM_velocity(ρ; κGM, maxslope)                     k_triad x2, k_gm_taper, k_dyad x2
//     (ϕᵢ*, ϕⱼ*) = velocity2fluxes(u*, v*, gridmetrics, ρ)                   k_velflux<0>
//     umo' = umo + ϕᵢ*,  vmo' = vmo + ϕⱼ*                                   k_add_gm (below)
//     ϕ = facefluxes(umo', vmo')  ->  transportmatrix(ϕ, ...)               k_faceflux, k_fused_v4
// GM stays an ADVECTION by the bolus velocity, so T keeps its 7-point pattern; with κGM = 0 the bolus fluxes are
// exact zeros and the result is bit-identical to the plain path.  The sum is defined here (no reference arithmetic
// exists for it): a valid umo / vmo value gets the bolus flux added when that is not NaN (u* is NaN where a triad or
// dyad has no valid neighbour); fill / NaN values of umo / vmo stay as they are, so nofluxboundaries! and the
// all-fill assertion see the same cells as without GM; on the tripolar fold row the two cells that share a north
// face average their estimates of its flux antisymmetrically (k_add_gm), which keeps the transport non-divergent.
#include "common.cuh"

int otmb_bolus_gm_dev(otmb_ctx* c, const double* d_rho, double kGM, double maxslope, double* d_si, double* d_sj, double* d_u,
                      double* d_v);
int otmb_redigm_prereq(otmb_ctx* c);
int otmb_velocity2fluxes_dev(otmb_ctx* c, const double* d_u, const double* d_v, const double* d_rho3d, double rho, double* d_phi_i,
                             double* d_phi_j);

namespace {

__global__ void __launch_bounds__(256) k_add_gm(double* __restrict__ umo, double* __restrict__ vmo, const double* __restrict__ gi,
                                                const double* __restrict__ gj, double fill, GridDims g) {
    const int L = blockIdx.x * blockDim.x + threadIdx.x;
    if (L >= g.M) return;
    const double u = umo[L], v = vmo[L], a = gi[L];
    double b = gj[L];
    // The north face of (i, ny) IS the north face of its fold partner (nx-i+1, ny) (src/gridtopology.jl:94-95), seen with
    // the opposite sign, and facefluxes carries no fold term (src/velocities.jl:219-221): the two cells' estimates of
    // that one flux are averaged antisymmetrically, (b - b')/2 (the one that is not NaN when the other is), so that
    // what leaves one cell enters the other.
    const int p = L % g.P, j = p / g.nx;
    if (j == g.ny - 1) {
        const int i = p - j * g.nx;
        const double bm = gj[L + (g.nx - 1 - 2 * i)];
        const bool wa = !isnan(b), wb = !isnan(bm);
        b = ((wa ? b : 0.0) - (wb ? bm : 0.0)) / (double)((int)wa + (int)wb);
    }
    if (!(isnan(u) || u == fill) && !isnan(a)) umo[L] = u + a;
    if (!(isnan(v) || v == fill) && !isnan(b)) vmo[L] = v + b;
}

}  // namespace

extern "C" int otmb_facefluxes_gm(otmb_ctx* c, const double* umo, const double* vmo, double fill, const double* rho3d, double kGM,
                                  double maxslope, double rho_scalar, int32_t flux_rho_is_3d, double* east, double* west,
                                  double* north, double* south, double* top, double* bottom, double* gm_phi_i, double* gm_phi_j) {
    if (!c || !umo || !vmo || !rho3d) return OTMB_ERR_BADARG;
    OT_TRY(otmb_redigm_prereq(c));
    OT_TRY(otmb_need(c, c->have_metrics, "otmb_gridmetrics / otmb_set_gridmetrics"));
    if (c->sharded) return otmb_fail(c, OTMB_ERR_STATE, "the GM extension is not available on a slab context");
    if (c->topo == OTMB_TOPO_BIPOLAR)
        return otmb_fail(c, OTMB_ERR_BADARG, "the GM extension needs a tripolar grid: velocity2fluxes and the J-triads index the "
                                             "missing north neighbour of the last row on bipolar grids (the reference throws)");
    CU_TRY(c, cudaSetDevice(c->device));
    const size_t M8 = (size_t)c->M * 8;
    DevBuf *b = c->coo;   // scratch of the COO path, idle here: ρ, Sᵢ, Sⱼ, u*, v*, ϕᵢ*, ϕⱼ*
    for (int q = 0; q < 7; ++q) CU_TRY(c, b[q].ensure(M8));
    OT_TRY(otmb_h2d(c, b[0].p, rho3d, M8, c->stream));
    OT_TRY(otmb_upload_uv(c, umo, vmo, fill));
    OT_TRY(otmb_reset_flags(c));
    OT_TRY(otmb_bolus_gm_dev(c, b[0].as<double>(), kGM, maxslope, b[1].as<double>(), b[2].as<double>(), b[3].as<double>(),
                             b[4].as<double>()));
    OT_TRY(otmb_velocity2fluxes_dev(c, b[3].as<double>(), b[4].as<double>(), flux_rho_is_3d ? b[0].as<double>() : nullptr, rho_scalar,
                                    b[5].as<double>(), b[6].as<double>()));
    GridDims g{(int)c->nx, (int)c->ny, (int)c->nz, c->topo, (int)c->P, (int)c->M};
    k_add_gm<<<grid_for(c->M, 256), 256, 0, c->stream>>>(c->stage_a.as<double>(), c->stage_b.as<double>(), b[5].as<double>(),
                                                         b[6].as<double>(), fill, g);
    LAUNCHED(c);
    CU_TRY(c, cudaGetLastError());
    OT_TRY(otmb_fetch_flags(c));
    if (c->h_flags->err_dry_neighbour)
        return otmb_fail(c, OTMB_ERR_DRY_NEIGHBOUR, "triad group needs a neighbour that does not exist (reference throws)");
    if (gm_phi_i) CU_TRY(c, cudaMemcpyAsync(gm_phi_i, b[5].p, M8, cudaMemcpyDeviceToHost, c->stream));
    if (gm_phi_j) CU_TRY(c, cudaMemcpyAsync(gm_phi_j, b[6].p, M8, cudaMemcpyDeviceToHost, c->stream));
    // facefluxes of the summed transports (nofluxboundaries! + continuity, faceflux.cu)
    OT_TRY(otmb_faceflux_begin(c, fill));
    OT_TRY(otmb_faceflux_columns(c, fill, 0, c->P, nullptr, nullptr));
    OT_TRY(otmb_fetch_flags(c));
    if (!c->h_flags->any_valid_u || !c->h_flags->any_valid_v) {
        c->have_phi = false;
        return otmb_fail(c, OTMB_ERR_ALL_FILL, otmb_status_string(OTMB_ERR_ALL_FILL));
    }
    double* const outs[6] = {east, west, north, south, top, bottom};
    OT_TRY(otmb_faceflux_copy_out(c, outs));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    c->have_phi = true;
    return OTMB_OK;
}
