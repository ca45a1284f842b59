// redigm.cu — K10: triad slopes, vertical dyad derivative, GM bolus velocity.
// Replaces globalverticalfacetriadderivative (/root/reference/src/triads.jl:84-146),
// globalverticaldyadderivative (/root/reference/src/dyads.jl:38-78) and bolus_GM_velocity
// (/root/reference/src/RediGM.jl:46-79).  These are experimental, non-exported helpers in the
// reference that return 3-D fields; nothing there turns them into matrix entries (SURVEY.md §8a
// rows A18-A20), so there is no Redi/GM matrix operator to be drop-in for.
//
// One thread per linear cell, NaN at dry cells.  The nan-mean idiom multiplies by Bool weights
// (false * NaN == 0.0 in Julia) and sums left to right in declaration order.
#include "common.cuh"
#include "sphere.cuh"

namespace {

__device__ __forceinline__ double qnan() { return __longlong_as_double(0x7ff8000000000000ll); }
__device__ __forceinline__ double getornan(const double* __restrict__ a, int L) { return L >= 0 ? __ldg(a + L) : qnan(); }
__device__ __forceinline__ double bmul(bool w, double v) { return w ? v : 0.0; }

// dir 0 = Icoord (east neighbour), 1 = Jcoord (north neighbour incl. fold)
__global__ void __launch_bounds__(256) k_triad(const double* __restrict__ chi, const double* __restrict__ lon,
                                               const double* __restrict__ lat, const double* __restrict__ Z,
                                               const u64* __restrict__ mask, GridDims g, int dir,
                                               double* __restrict__ out, int* __restrict__ err) {
    const int L = blockIdx.x * blockDim.x + threadIdx.x;
    if (L >= g.M) return;
    if (!wet_at(mask, L)) {
        out[L] = qnan();
        return;
    }
    const int k = L / g.P, p2 = L - k * g.P, j = p2 / g.nx, i = p2 - j * g.nx;
    const int LN = k > 0 ? L - g.P : -1, LS = k < g.nz - 1 ? L + g.P : -1;
    int LE;
    if (dir == 0)
        LE = i < g.nx - 1 ? L + 1 : L - (g.nx - 1);
    else
        LE = j < g.ny - 1 ? L + g.nx : (g.topo == OTMB_TOPO_TRIPOLAR ? k * g.P + (g.ny - 1) * g.nx + (g.nx - 1 - i) : -1);
    if (LE < 0) {  // k₋₁(nothing, ...) throws in the reference (triads.jl:87-89)
        atomicOr(err, 1);
        out[L] = qnan();
        return;
    }
    const int LNE = k > 0 ? LE - g.P : -1, LSE = k < g.nz - 1 ? LE + g.P : -1;
    const double vC = __ldg(chi + L), vN = getornan(chi, LN), vS = getornan(chi, LS), vE = __ldg(chi + LE),
                 vNE = getornan(chi, LNE), vSE = getornan(chi, LSE);
    const double zC = __ldg(Z + L), zE = __ldg(Z + LE);
    const double dCN = fabs(getornan(Z, LN) - zC), dCS = fabs(getornan(Z, LS) - zC);
    const int pE = LE - k * g.P;
    const double dCE = haversine_dev(__ldg(lon + p2), __ldg(lat + p2), __ldg(lon + pE), __ldg(lat + pE));
    const double dENE = fabs(getornan(Z, LNE) - zE), dESE = fabs(getornan(Z, LSE) - zE);
    const double CN = (vN - vC) / dCN, CS = (vC - vS) / dCS, CE = (vE - vC) / dCE, ENE = (vNE - vE) / dENE,
                 ESE = (vE - vSE) / dESE;
    const double r0 = CE / CN, r1 = CE / CS, r2 = CE / ENE, r3 = CE / ESE;
    const bool w0 = !isnan(r0), w1 = !isnan(r1), w2 = !isnan(r2), w3 = !isnan(r3);
    double s = bmul(w0, r0);
    s = s + bmul(w1, r1);
    s = s + bmul(w2, r2);
    s = s + bmul(w3, r3);
    out[L] = s / (double)((int)w0 + (int)w1 + (int)w2 + (int)w3);
}

__global__ void __launch_bounds__(256) k_dyad(const double* __restrict__ chi, const double* __restrict__ Z,
                                              const u64* __restrict__ mask, GridDims g, double* __restrict__ out) {
    const int L = blockIdx.x * blockDim.x + threadIdx.x;
    if (L >= g.M) return;
    if (!wet_at(mask, L)) {
        out[L] = qnan();
        return;
    }
    const int k = L / g.P;
    const int LN = k > 0 ? L - g.P : -1, LS = k < g.nz - 1 ? L + g.P : -1;
    const double vC = __ldg(chi + L), zC = __ldg(Z + L);
    const double d0 = (getornan(chi, LN) - vC) / fabs(getornan(Z, LN) - zC);
    const double d1 = (vC - getornan(chi, LS)) / fabs(getornan(Z, LS) - zC);
    const bool w0 = !isnan(d0), w1 = !isnan(d1);
    out[L] = (bmul(w0, d0) + bmul(w1, d1)) / (double)((int)w0 + (int)w1);
}

// clamp, taper and scale by κGM (RediGM.jl:56-76), in place on the two slope fields
__global__ void __launch_bounds__(256) k_gm_taper(double* __restrict__ Si, double* __restrict__ Sj, int M, double kGM,
                                                  double maxslope) {
    const int L = blockIdx.x * blockDim.x + threadIdx.x;
    if (L >= M) return;
    double a = Si[L], b = Sj[L];
    a = a > maxslope ? maxslope : (a < -maxslope ? -maxslope : a);
    b = b > maxslope ? maxslope : (b < -maxslope ? -maxslope : b);
    const double Sc = 0.004, Sd = 0.001;
    const double taper = 0.5 * (1 + tanh((Sc - sqrt(a * a + b * b)) / Sd));
    Si[L] = kGM * (taper * a);
    Sj[L] = kGM * (taper * b);
}

int prereq(otmb_ctx* c) {
    OT_TRY(otmb_need(c, c->have_indices, "otmb_makeindices"));
    OT_TRY(otmb_need(c, c->have_z3d && c->have_lonlat, "otmb_gridmetrics (Z3D, lon, lat)"));
    if (c->topo == OTMB_TOPO_UNKNOWN) return otmb_fail(c, OTMB_ERR_UNKNOWN_GRID, otmb_status_string(OTMB_ERR_UNKNOWN_GRID));
    return OTMB_OK;
}

int triad_dev(otmb_ctx* c, const double* dchi, int dir, double* dout) {
    GridDims g{(int)c->nx, (int)c->ny, (int)c->nz, c->topo, (int)c->P, (int)c->M};
    k_triad<<<grid_for(c->M, 256), 256, 0, c->stream>>>(dchi, c->lon.as<double>(), c->lat.as<double>(), c->Z3D.as<double>(),
                                                        c->mask.as<u64>(), g, dir, dout,
                                                        &c->flags.as<DevFlags>()->err_dry_neighbour);
    LAUNCHED(c);
    CU_TRY(c, cudaGetLastError());
    return OTMB_OK;
}
int dyad_dev(otmb_ctx* c, const double* dchi, double* dout) {
    GridDims g{(int)c->nx, (int)c->ny, (int)c->nz, c->topo, (int)c->P, (int)c->M};
    k_dyad<<<grid_for(c->M, 256), 256, 0, c->stream>>>(dchi, c->Z3D.as<double>(), c->mask.as<u64>(), g, dout);
    LAUNCHED(c);
    CU_TRY(c, cudaGetLastError());
    return OTMB_OK;
}

}  // namespace

// bolus_GM_velocity on DEVICE buffers (all of M doubles): d_si / d_sj are scratch for the two slope fields, d_u / d_v
// the results.  The caller resets / reads the flag block (err_dry_neighbour: a triad group needs a neighbour that
// does not exist, where the reference throws).
int otmb_bolus_gm_dev(otmb_ctx* c, const double* d_rho, double kGM, double maxslope, double* d_si, double* d_sj, double* d_u,
                      double* d_v) {
    OT_TRY(triad_dev(c, d_rho, 0, d_si));
    OT_TRY(triad_dev(c, d_rho, 1, d_sj));
    k_gm_taper<<<grid_for(c->M, 256), 256, 0, c->stream>>>(d_si, d_sj, (int)c->M, kGM, maxslope);
    LAUNCHED(c);
    CU_TRY(c, cudaGetLastError());
    OT_TRY(dyad_dev(c, d_si, d_u));
    OT_TRY(dyad_dev(c, d_sj, d_v));
    return OTMB_OK;
}
int otmb_redigm_prereq(otmb_ctx* c) { return prereq(c); }

extern "C" {

int otmb_triad_derivative(otmb_ctx* c, const double* chi, int dir, double* out) {
    if (!c || !chi || !out || dir < 0 || dir > 1) return OTMB_ERR_BADARG;
    OT_TRY(prereq(c));
    CU_TRY(c, cudaSetDevice(c->device));
    const size_t M8 = (size_t)c->M * 8;
    CU_TRY(c, c->stage_a.ensure(M8));
    CU_TRY(c, c->stage_b.ensure(M8));
    OT_TRY(otmb_h2d(c, c->stage_a.p, chi, M8, c->stream));
    OT_TRY(otmb_reset_flags(c));
    OT_TRY(triad_dev(c, c->stage_a.as<double>(), dir, c->stage_b.as<double>()));
    OT_TRY(otmb_fetch_flags(c));
    if (c->h_flags->err_dry_neighbour)
        return otmb_fail(c, OTMB_ERR_DRY_NEIGHBOUR, "triad group needs a neighbour that does not exist (reference throws)");
    CU_TRY(c, cudaMemcpyAsync(out, c->stage_b.p, M8, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return OTMB_OK;
}

int otmb_dyad_derivative(otmb_ctx* c, const double* chi, double* out) {
    if (!c || !chi || !out) return OTMB_ERR_BADARG;
    OT_TRY(prereq(c));
    CU_TRY(c, cudaSetDevice(c->device));
    const size_t M8 = (size_t)c->M * 8;
    CU_TRY(c, c->stage_a.ensure(M8));
    CU_TRY(c, c->stage_b.ensure(M8));
    OT_TRY(otmb_h2d(c, c->stage_a.p, chi, M8, c->stream));
    OT_TRY(dyad_dev(c, c->stage_a.as<double>(), c->stage_b.as<double>()));
    CU_TRY(c, cudaMemcpyAsync(out, c->stage_b.p, M8, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return OTMB_OK;
}

int otmb_bolus_gm_velocity(otmb_ctx* c, const double* rho, double kGM, double maxslope, double* u, double* v) {
    if (!c || !rho || !u || !v) return OTMB_ERR_BADARG;
    OT_TRY(prereq(c));
    CU_TRY(c, cudaSetDevice(c->device));
    const size_t M8 = (size_t)c->M * 8;
    DevBuf &dr = c->stage_a, &si = c->stage_b, &sj = c->add_tmp[0], &du = c->add_tmp[1], &dv = c->add_tmp[2];
    CU_TRY(c, dr.ensure(M8));
    CU_TRY(c, si.ensure(M8));
    CU_TRY(c, sj.ensure(M8));
    CU_TRY(c, du.ensure(M8));
    CU_TRY(c, dv.ensure(M8));
    OT_TRY(otmb_h2d(c, dr.p, rho, M8, c->stream));
    OT_TRY(otmb_reset_flags(c));
    OT_TRY(triad_dev(c, dr.as<double>(), 0, si.as<double>()));
    OT_TRY(triad_dev(c, dr.as<double>(), 1, sj.as<double>()));
    k_gm_taper<<<grid_for(c->M, 256), 256, 0, c->stream>>>(si.as<double>(), sj.as<double>(), (int)c->M, kGM, maxslope);
    LAUNCHED(c);
    CU_TRY(c, cudaGetLastError());
    OT_TRY(dyad_dev(c, si.as<double>(), du.as<double>()));
    OT_TRY(dyad_dev(c, sj.as<double>(), dv.as<double>()));
    OT_TRY(otmb_fetch_flags(c));
    if (c->h_flags->err_dry_neighbour)
        return otmb_fail(c, OTMB_ERR_DRY_NEIGHBOUR, "triad group needs a neighbour that does not exist (reference throws)");
    CU_TRY(c, cudaMemcpyAsync(u, du.p, M8, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaMemcpyAsync(v, dv.p, M8, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return OTMB_OK;
}

}  // extern "C"
