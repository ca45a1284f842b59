// geometry.cu — K2 (thkcello, Z3D) and K3 (haversine edge lengths / distances).
// Replaces the numerics of makegridmetrics, /root/reference/src/gridcellgeometry.jl:283-285, 304-308.
//
// haversine restates Distances.jl v0.10 (not under /root/reference; compat "0.10",
// /root/reference/Project.toml:14): a = sind(dlat/2)^2 + cosd(lat1)*cosd(lat2)*sind(dlon/2)^2,
// d = 2*(6371000*asin(min(sqrt(a),1))), with Julia-Base-style sind/cosd (exact quadrant
// reduction in degrees, double-double deg->rad, fdlibm kernels) so that multiples of 90
// degrees are exact (a bipolar grid's top edge at lat = 90 must come out exactly 0).
#include "common.cuh"
#include "sphere.cuh"

namespace {

// K2: one thread per (i,j) column, sequential in k like cumsum(dims=3).  There are only nx*ny threads (108 000 at 1
// degree: a third of the GPU's thread slots), so each keeps UNROLL levels in flight: the loads of a batch are issued
// before the first quotient is formed, the UNROLL divisions are independent, and only the running sum is serial.
template <int UNROLL>
__global__ void __launch_bounds__(128) k_metrics3d(const double* __restrict__ v3D, const double* __restrict__ area2D, int P,
                                                   int nz, double* __restrict__ thk, double* __restrict__ Z3D) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const double a = __ldg(area2D + p);
    double zbot = 0.0;
    for (int k0 = 0; k0 < nz; k0 += UNROLL) {
        double v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) v[u] = k0 + u < nz ? __ldg(v3D + (size_t)p + (size_t)P * (k0 + u)) : 0.0;
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) v[u] = v[u] / a;                       // thkcello = v3D ./ area2D  (:283)
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (k0 + u >= nz) break;
            const size_t L = (size_t)p + (size_t)P * (k0 + u);
            const double t = v[u];
            __stcs(thk + L, t);
            zbot = (k0 + u == 0) ? t : zbot + t;                                // cumsum(thkcello, dims = 3)  (:284)
            __stcs(Z3D + L, zbot - 0.5 * t);                                    // - 0.5 thkcello             (:285)
        }
    }
}

// K3: one thread per (i,j): 4 edge lengths, 4 centre->edge-midpoint, 4 centre->neighbour
__global__ void __launch_bounds__(128) k_geom2d(const double* __restrict__ lon, const double* __restrict__ lat,
                                                const double* __restrict__ lonv, const double* __restrict__ latv,
                                                GridDims g, double* __restrict__ edge, double* __restrict__ dedge,
                                                double* __restrict__ dnbr) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= g.P) return;
    const int i = p % g.nx, j = p / g.nx;
    const double2* lv2 = reinterpret_cast<const double2*>(lonv + 4 * (size_t)p);
    const double2* tv2 = reinterpret_cast<const double2*>(latv + 4 * (size_t)p);
    const double2 la = __ldg(lv2), lb = __ldg(lv2 + 1), ta = __ldg(tv2), tb = __ldg(tv2 + 1);
    const double vl[4] = {la.x, la.y, lb.x, lb.y};
    const double vt[4] = {ta.x, ta.y, tb.x, tb.y};
    const double clon = __ldg(lon + p), clat = __ldg(lat + p);
    // dirs south, east, north, west -> vertex pairs (1,2) (2,3) (3,4) (1,4), gridcellgeometry.jl:209-215
    const int va[4] = {0, 1, 2, 0}, vb[4] = {1, 2, 3, 3};
#pragma unroll
    for (int d = 0; d < 4; ++d) {
        const double alon = vl[va[d]], alat = vt[va[d]], blon = vl[vb[d]], blat = vt[vb[d]];
        edge[(size_t)d * g.P + p] = haversine_dev(alon, alat, blon, blat);
        double mlon, mlat;  // midpointonsphere, gridcellgeometry.jl:249-255
        if (fabs(alon - blon) < 180) {
            mlon = (alon + blon) / 2;
            mlat = (alat + blat) / 2;
        } else {
            mlon = (alon + blon) / 2 + 180;
            mlat = (alat + blat) / 2 + 0;
        }
        dedge[(size_t)d * g.P + p] = haversine_dev(clon, clat, mlon, mlat);
    }
    // neighbours j-1, i+1, j+1, i-1 (gridcellgeometry.jl:305) under the topology (gridtopology.jl:59-65,95)
    int q[4];
    q[0] = j > 0 ? p - g.nx : -1;
    q[1] = i < g.nx - 1 ? p + 1 : p - (g.nx - 1);
    q[2] = j < g.ny - 1 ? p + g.nx : (g.topo == OTMB_TOPO_TRIPOLAR ? (g.nx - 1 - i) + g.nx * (g.ny - 1) : -1);
    q[3] = i > 0 ? p - 1 : p + (g.nx - 1);
#pragma unroll
    for (int d = 0; d < 4; ++d) {
        double v = __longlong_as_double(0x7ff8000000000000ll);  // horizontaldistance(..., ::Nothing) = NaN, :189
        if (q[d] >= 0) v = haversine_dev(clon, clat, __ldg(lon + q[d]), __ldg(lat + q[d]));
        dnbr[(size_t)d * g.P + p] = v;
    }
}

int upload(otmb_ctx* ctx, DevBuf& buf, const void* host, size_t bytes) {
    CU_TRY(ctx, buf.ensure(bytes));
    OT_TRY(otmb_h2d(ctx, buf.p, host, bytes, ctx->stream));
    return OTMB_OK;
}

}  // namespace

extern "C" int otmb_gridmetrics(otmb_ctx* c, const double* area2D, const double* lon, const double* lat,
                                const double* lonv, const double* latv, const double* zt, double* thk, double* Z3D,
                                double* edge, double* dedge, double* dnbr) {
    if (!c || !area2D || !lon || !lat || !lonv || !latv || !zt) return OTMB_ERR_BADARG;
    OT_TRY(otmb_need(c, c->have_indices, "otmb_makeindices"));
    if (c->sharded)
        return otmb_fail(c, OTMB_ERR_STATE, "otmb_gridmetrics needs the whole grid (Z3D is a top-down cumsum): compute the metrics on an "
                                            "unsharded context and give the slab contexts otmb_set_gridmetrics");
    CU_TRY(c, cudaSetDevice(c->device));
    const size_t P8 = (size_t)c->P * 8, M8 = (size_t)c->M * 8;
    OT_TRY(upload(c, c->area2D, area2D, P8));
    OT_TRY(upload(c, c->lon, lon, P8));
    OT_TRY(upload(c, c->lat, lat, P8));
    OT_TRY(upload(c, c->lonv, lonv, 4 * P8));
    OT_TRY(upload(c, c->latv, latv, 4 * P8));
    OT_TRY(upload(c, c->zt, zt, (size_t)c->nz * 8));
    CU_TRY(c, c->thk.ensure(M8));
    CU_TRY(c, c->Z3D.ensure(M8));
    CU_TRY(c, c->edge.ensure(4 * P8));
    CU_TRY(c, c->dedge.ensure(4 * P8));
    CU_TRY(c, c->dnbr.ensure(4 * P8));
    GridDims g{(int)c->nx, (int)c->ny, (int)c->nz, c->topo, (int)c->P, (int)c->M};
    k_metrics3d<10><<<grid_for(c->P, 128), 128, 0, c->stream>>>(c->v3D.as<double>(), c->area2D.as<double>(), g.P, g.nz,
                                                                c->thk.as<double>(), c->Z3D.as<double>());
    LAUNCHED(c);
    CU_TRY(c, cudaGetLastError());
    c->have_z3d = true;
    c->have_lonlat = true;
    if (thk) CU_TRY(c, cudaMemcpyAsync(thk, c->thk.p, M8, cudaMemcpyDeviceToHost, c->stream));
    if (Z3D) CU_TRY(c, cudaMemcpyAsync(Z3D, c->Z3D.p, M8, cudaMemcpyDeviceToHost, c->stream));
    if (c->topo == OTMB_TOPO_UNKNOWN) {
        // the comprehension at gridcellgeometry.jl:308 calls j₋₁ etc., which error for Unknown grids
        CU_TRY(c, cudaStreamSynchronize(c->stream));
        return otmb_fail(c, OTMB_ERR_UNKNOWN_GRID, otmb_status_string(OTMB_ERR_UNKNOWN_GRID));
    }
    k_geom2d<<<grid_for(c->P, 128), 128, 0, c->stream>>>(c->lon.as<double>(), c->lat.as<double>(), c->lonv.as<double>(),
                                                         c->latv.as<double>(), g, c->edge.as<double>(),
                                                         c->dedge.as<double>(), c->dnbr.as<double>());
    LAUNCHED(c);
    CU_TRY(c, cudaGetLastError());
    if (edge) CU_TRY(c, cudaMemcpyAsync(edge, c->edge.p, 4 * P8, cudaMemcpyDeviceToHost, c->stream));
    if (dedge) CU_TRY(c, cudaMemcpyAsync(dedge, c->dedge.p, 4 * P8, cudaMemcpyDeviceToHost, c->stream));
    if (dnbr) CU_TRY(c, cudaMemcpyAsync(dnbr, c->dnbr.p, 4 * P8, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    c->have_metrics = true;
    return OTMB_OK;
}
