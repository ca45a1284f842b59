// velocity.cu — velocity <-> mass-flux conversion and the B-grid -> C-grid interpolation, the step
// in front of facefluxes for models that publish uo/vo instead of umo/vmo.
// Replaces /root/reference/src/velocities.jl:10-39 (velocity2fluxes), :50-74 (fluxes2velocity),
// :81-108 (twocellnanmean / nanmean2 / twocellnanmin / nanmin2) and the B-grid branch of
// interpolateontodefaultCgrid, /root/reference/src/gridcellgeometry.jl:103-140.
// One thread per grid cell (the reference loops over ALL cells, `for 𝑖 in indices.C`), i fastest:
// every load is a coalesced row segment; the east / north neighbour values are the same or the
// next cache line.  Multiplications are evaluated left to right as written in the reference and
// never contracted (-fmad=false), so results are bit-identical to a plain IEEE evaluation.
#include "common.cuh"

namespace {

// nanmean2, src/velocities.jl:89-93: Bool * NaN is 0.0 in Julia, and 0/0 = NaN when both are NaN
__device__ __forceinline__ double nanmean2(double a, double b) {
    const bool wa = !isnan(a), wb = !isnan(b);
    const double s = (wa ? a : 0.0) + (wb ? b : 0.0);
    return s / (double)((int)wa + (int)wb);
}
// nanmin2, src/velocities.jl:108
__device__ __forceinline__ double nanmin2(double a, double b) { return isnan(a) ? b : isnan(b) ? a : jl_min(a, b); }

// MODE 0: ϕ = ((u * ρ̄) * thk) * edge   (:31-33);  MODE 1: u = ϕ / ((ρ̄ * thk) * edge)   (:66-68)
template <int MODE>
__global__ void __launch_bounds__(256) k_velflux(const double* __restrict__ a_i, const double* __restrict__ a_j,
                                                 const double* __restrict__ rho3d, double rho, const double* __restrict__ thk,
                                                 const double* __restrict__ edge, GridDims g, double* __restrict__ o_i,
                                                 double* __restrict__ o_j) {
    const int L = blockIdx.x * blockDim.x + threadIdx.x;
    if (L >= g.M) return;
    const int k = L / g.P, p = L - k * g.P, j = p / g.nx, i = p - j * g.nx;
    const int LE = i < g.nx - 1 ? L + 1 : L - (g.nx - 1);                                 // i₊₁, periodic
    const int LN = j < g.ny - 1 ? L + g.nx : k * g.P + (g.ny - 1) * g.nx + (g.nx - 1 - i);   // j₊₁, tripolar fold at j = ny
    const double tC = __ldg(thk + L);
    const double rC = rho3d ? __ldg(rho3d + L) : rho;
    const double rE = rho3d ? nanmean2(rC, __ldg(rho3d + LE)) : rho;
    const double rN = rho3d ? nanmean2(rC, __ldg(rho3d + LN)) : rho;
    const double tE = nanmin2(tC, __ldg(thk + LE)), tN = nanmin2(tC, __ldg(thk + LN));
    const double eE = __ldg(edge + OTMB_DIR_EAST * g.P + p), eN = __ldg(edge + OTMB_DIR_NORTH * g.P + p);
    if (MODE == 0) {
        o_i[L] = ((__ldg(a_i + L) * rE) * tE) * eE;
        o_j[L] = ((__ldg(a_j + L) * rN) * tN) * eN;
    } else {
        o_i[L] = __ldg(a_i + L) / ((rE * tE) * eE);
        o_j[L] = __ldg(a_j + L) / ((rN * tN) * eN);
    }
}

// B-grid (u, v at the NE corner) -> C-grid, src/gridcellgeometry.jl:123-128: fill -> 0, then
// u2[i,j] = 0.5 (u[i,j] + u[i,j-1]) with 0 at j = 1;  v2[i,j] = 0.5 (v[i,j] + v[i-1,j]) with 0 at i = 1
// (no periodic wrap, like the reference).
__global__ void __launch_bounds__(256) k_bgrid2cgrid(const double* __restrict__ u, const double* __restrict__ v, double fill,
                                                     GridDims g, double* __restrict__ u2, double* __restrict__ v2) {
    const int L = blockIdx.x * blockDim.x + threadIdx.x;
    if (L >= g.M) return;
    const int p = L % g.P, j = p / g.nx, i = p - j * g.nx;
    auto clean = [fill](double x) { return x == fill ? 0.0 : x; };      // replace(u, _FillValue => 0.0): isequal match
    const double uS = j > 0 ? clean(__ldg(u + L - g.nx)) : 0.0;
    const double vW = i > 0 ? clean(__ldg(v + L - 1)) : 0.0;
    u2[L] = 0.5 * (clean(__ldg(u + L)) + uS);
    v2[L] = 0.5 * (clean(__ldg(v + L)) + vW);
}

int run_velflux(otmb_ctx* c, int mode, const double* a_i, const double* a_j, const double* rho3d, double rho, double* o_i,
                double* o_j) {
    if (!c || !a_i || !a_j || !o_i || !o_j) return OTMB_ERR_BADARG;
    OT_TRY(otmb_need(c, c->have_metrics, "otmb_gridmetrics / otmb_set_gridmetrics"));
    if (c->sharded) return otmb_fail(c, OTMB_ERR_STATE, "velocity conversion is not available on a slab context");
    if (c->topo == OTMB_TOPO_UNKNOWN) return otmb_fail(c, OTMB_ERR_UNKNOWN_GRID, otmb_status_string(OTMB_ERR_UNKNOWN_GRID));
    // bipolar grids: the reference indexes thkcello[nothing] at j = ny (src/velocities.jl:32-33) and throws
    if (c->topo == OTMB_TOPO_BIPOLAR)
        return otmb_fail(c, OTMB_ERR_BADARG, "velocity2fluxes / fluxes2velocity need a tripolar grid: the reference indexes "
                                             "the missing north neighbour of the last row on bipolar grids and throws");
    CU_TRY(c, cudaSetDevice(c->device));
    const size_t M8 = (size_t)c->M * 8;
    DevBuf* b = c->add_tmp;   // scratch: inputs 0,1 (+ rho 2), outputs 3,4
    for (int q = 0; q < 5; ++q) CU_TRY(c, b[q].ensure(M8));
    OT_TRY(otmb_h2d(c, b[0].p, a_i, M8, c->stream));
    OT_TRY(otmb_h2d(c, b[1].p, a_j, M8, c->stream));
    if (rho3d) OT_TRY(otmb_h2d(c, b[2].p, rho3d, M8, c->stream));
    GridDims g{(int)c->nx, (int)c->ny, (int)c->nz, c->topo, (int)c->P, (int)c->M};
    const double* dr = rho3d ? b[2].as<double>() : nullptr;
    if (mode == 0)
        k_velflux<0><<<grid_for(c->M, 256), 256, 0, c->stream>>>(b[0].as<double>(), b[1].as<double>(), dr, rho, c->thk.as<double>(),
                                                                  c->edge.as<double>(), g, b[3].as<double>(), b[4].as<double>());
    else
        k_velflux<1><<<grid_for(c->M, 256), 256, 0, c->stream>>>(b[0].as<double>(), b[1].as<double>(), dr, rho, c->thk.as<double>(),
                                                                  c->edge.as<double>(), g, b[3].as<double>(), b[4].as<double>());
    LAUNCHED(c);
    CU_TRY(c, cudaGetLastError());
    CU_TRY(c, cudaMemcpyAsync(o_i, b[3].p, M8, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaMemcpyAsync(o_j, b[4].p, M8, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return OTMB_OK;
}

}  // namespace

// velocity2fluxes on DEVICE buffers (src/velocities.jl:10-39); d_rho3d may be null (scalar rho)
int otmb_velocity2fluxes_dev(otmb_ctx* c, const double* d_u, const double* d_v, const double* d_rho3d, double rho, double* d_phi_i,
                             double* d_phi_j) {
    GridDims g{(int)c->nx, (int)c->ny, (int)c->nz, c->topo, (int)c->P, (int)c->M};
    k_velflux<0><<<grid_for(c->M, 256), 256, 0, c->stream>>>(d_u, d_v, d_rho3d, rho, c->thk.as<double>(), c->edge.as<double>(), g,
                                                              d_phi_i, d_phi_j);
    LAUNCHED(c);
    CU_TRY(c, cudaGetLastError());
    return OTMB_OK;
}

extern "C" {

int otmb_velocity2fluxes(otmb_ctx* c, const double* u, const double* v, const double* rho3d, double rho, double* phi_i,
                         double* phi_j) {
    return run_velflux(c, 0, u, v, rho3d, rho, phi_i, phi_j);
}

int otmb_fluxes2velocity(otmb_ctx* c, const double* phi_i, const double* phi_j, const double* rho3d, double rho, double* u,
                         double* v) {
    return run_velflux(c, 1, phi_i, phi_j, rho3d, rho, u, v);
}

int otmb_bgrid_to_cgrid(otmb_ctx* c, const double* u, const double* v, double fill_value, double* u2, double* v2) {
    if (!c || !u || !v || !u2 || !v2) return OTMB_ERR_BADARG;
    OT_TRY(otmb_need(c, c->have_grid, "otmb_set_grid"));
    CU_TRY(c, cudaSetDevice(c->device));
    const size_t M8 = (size_t)c->M * 8;
    DevBuf* b = c->add_tmp;
    for (int q = 0; q < 4; ++q) CU_TRY(c, b[q].ensure(M8));
    OT_TRY(otmb_h2d(c, b[0].p, u, M8, c->stream));
    OT_TRY(otmb_h2d(c, b[1].p, v, M8, c->stream));
    GridDims g{(int)c->nx, (int)c->ny, (int)c->nz, c->topo, (int)c->P, (int)c->M};
    k_bgrid2cgrid<<<grid_for(c->M, 256), 256, 0, c->stream>>>(b[0].as<double>(), b[1].as<double>(), fill_value, g,
                                                               b[2].as<double>(), b[3].as<double>());
    LAUNCHED(c);
    CU_TRY(c, cudaGetLastError());
    CU_TRY(c, cudaMemcpyAsync(u2, b[2].p, M8, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaMemcpyAsync(v2, b[3].p, M8, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return OTMB_OK;
}

}  // extern "C"
