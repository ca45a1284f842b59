// lump.cu — lump_and_spray: the coarsening operators applied to T in every downstream use
// (LUMP * T * SPRAY), /root/reference/src/extratools.jl:38-112.
//
// With the default mask (lump everywhere) the reference's sequential sweep assigns boxes of
// di x dj x dk cells anchored on the regular lattice, in lattice order (i fastest): inside a box the dry
// (and ghost) cells get the dry index, the wet cells are split into connected components of T's stored
// pattern (Graphs.connected_components: labelled by the smallest vertex, returned in that order) and every
// component gets the next coarse index.  One thread per box does the component search (a box holds at
// most 32 cells and T's columns at most 7 rows); an exclusive scan of the per-box component counts gives
// the coarse indices; a second pass sums the coarse volumes in ascending fine index (the order of the
// reference's `LUMP * vol` SpMV) and fills LUMP (N_c x N, one entry per column, (1/vol_c) * vol) and
// SPRAY = LUMP' with ones.  A custom `mask` makes the anchors data dependent (greedy sweep): that case
// stays with the host implementation and is rejected here.
#include "common.cuh"

namespace {

constexpr int BOXMAX = 32;

struct LumpParams {
    GridDims g;
    int di, dj, dk, nbx, nby, nbz;
    const int* rank3d;
    const i64* t_colptr;
    const i64* t_rowval;
    int t_base;
};

__device__ __forceinline__ bool has_entry(const LumpParams& P, int row, int col) {
    const i64 a = P.t_colptr[col] - P.t_base, b = P.t_colptr[col + 1] - P.t_base;
    for (i64 p = a; p < b; ++p)
        if (P.t_rowval[p] - P.t_base == row) return true;
    return false;
}

// wet cells of box `box` in the reference's order vec(L[C𝑖 .+ neighbours]) (i fastest); returns their number
__device__ __forceinline__ int box_cells(const LumpParams& P, int box, int* ranks) {
    const int bi = box % P.nbx, bj = (box / P.nbx) % P.nby, bk = box / (P.nbx * P.nby);
    int n = 0;
    for (int d = 0; d < P.dk; ++d)
        for (int b = 0; b < P.dj; ++b)
            for (int a = 0; a < P.di; ++a) {
                const int i = bi * P.di + a, j = bj * P.dj + b, k = bk * P.dk + d;
                if (i >= P.g.nx || j >= P.g.ny || k >= P.g.nz) continue;      // ghost cells of the extended grid are dry
                const int r = __ldg(P.rank3d + i + P.g.nx * (j + P.g.ny * k));
                if (r >= 0) ranks[n++] = r;
            }
    return n;
}

__global__ void __launch_bounds__(128) k_lump_cc(LumpParams P, int nboxes, uint32_t* __restrict__ ncomp,
                                                 int* __restrict__ cellcomp, int* __restrict__ asym) {
    const int box = blockIdx.x * blockDim.x + threadIdx.x;
    if (box >= nboxes) return;
    int ranks[BOXMAX], parent[BOXMAX];
    const int n = box_cells(P, box, ranks);
    for (int v = 0; v < n; ++v) parent[v] = v;
    auto find = [&](int a) {
        while (parent[a] != a) a = parent[a];
        return a;
    };
    for (int a = 0; a < n; ++a)
        for (int b = a + 1; b < n; ++b) {
            const bool ab = has_entry(P, ranks[a], ranks[b]), ba = has_entry(P, ranks[b], ranks[a]);
            if (ab != ba) *asym = 1;     // SimpleGraph(adjacency matrix) throws unless it is symmetric
            if (ab || ba) {
                const int ra = find(a), rb = find(b);
                if (ra != rb) parent[ra > rb ? ra : rb] = ra > rb ? rb : ra;   // the smaller vertex stays the root
            }
        }
    int nc = 0;
    for (int v = 0; v < n; ++v) {
        const int root = find(v);
        int before = 0;                   // components are numbered in the order of their smallest vertex
        for (int u = 0; u < root; ++u) before += find(u) == u;
        cellcomp[ranks[v]] = before;
        nc += root == v;
    }
    ncomp[box] = (uint32_t)nc;
}

__global__ void __launch_bounds__(128) k_lump_sizes(LumpParams P, int nboxes, const uint32_t* __restrict__ boxbase,
                                                    const int* __restrict__ cellcomp, const double* __restrict__ vol,
                                                    uint32_t* __restrict__ size, double* __restrict__ vol_c) {
    const int box = blockIdx.x * blockDim.x + threadIdx.x;
    if (box >= nboxes) return;
    int ranks[BOXMAX];
    const int n = box_cells(P, box, ranks);
    const uint32_t base = boxbase[box];
    // ascending fine index = the order in which `LUMP * vol` accumulates (extratools.jl:95)
    for (int v = 0; v < n; ++v) {
        const uint32_t cidx = base + (uint32_t)cellcomp[ranks[v]];
        size[cidx] += 1u;                                   // the box's coarse cells belong to this thread only
        vol_c[cidx] = vol_c[cidx] + 1.0 * vol[ranks[v]];
    }
}

__global__ void __launch_bounds__(128) k_lump_fill(LumpParams P, int nboxes, const uint32_t* __restrict__ boxbase,
                                                   const int* __restrict__ cellcomp, const double* __restrict__ vol,
                                                   const double* __restrict__ vol_c, const i64* __restrict__ spray_colptr,
                                                   int base, i64* __restrict__ l_colptr, i64* __restrict__ l_rowval,
                                                   double* __restrict__ l_nzval, i64* __restrict__ s_rowval,
                                                   double* __restrict__ s_nzval) {
    const int box = blockIdx.x * blockDim.x + threadIdx.x;
    if (box >= nboxes) return;
    int ranks[BOXMAX], filled[BOXMAX];
    const int n = box_cells(P, box, ranks);
    for (int v = 0; v < BOXMAX; ++v) filled[v] = 0;
    const uint32_t b0 = boxbase[box];
    for (int v = 0; v < n; ++v) {
        const int r = ranks[v], comp = cellcomp[r];
        const uint32_t cidx = b0 + (uint32_t)comp;
        l_colptr[r] = (i64)r + base;
        l_rowval[r] = (i64)cidx + base;
        l_nzval[r] = ((1.0 / vol_c[cidx]) * 1.0) * vol[r];   // Diagonal(1 ./ vol_c) * LUMP * Diagonal(vol), :96
        const i64 p = spray_colptr[cidx] - base + filled[comp]++;
        s_rowval[p] = (i64)r + base;                          // SPRAY = LUMP' with ones, :100-101
        s_nzval[p] = 1.0;
    }
}

__global__ void k_set_last(i64* p, i64 idx, i64 v) { p[idx] = v; }

// ---- T_c = LUMP * T * SPRAY (/root/reference/test/local_full.jl:161: the step right behind lump_and_spray in every
// downstream use).  Not a general SpGEMM: LUMP has ONE entry per column (the cell's box component I(k), weight w_k =
// (1/vol_c) * 1 * vol) and SPRAY column J holds ones at the members of component J, so
//     T_c[I, J] = sum over members j of J, ascending ( X[I, j] * 1.0 ),   X[I, j] = sum over rows k of T[:, j] with I(k) = I,
//                                                                                  ascending ( w_k * T[k, j] )
// which is exactly the order in which SparseArrays' Gustavson product (LUMP * T first, then * SPRAY) accumulates: the
// first product of an entry is stored, later ones are added left to right.  One thread per coarse column J gathers
// its entries into a small sorted list (a box of <= 32 cells with <= 7 entries per column touches few components);
// PASS 0 counts, PASS 1 fills behind an exclusive scan of the counts.  Structural zeros are kept (the product does
// not drop them).
constexpr int TC_MAX = 96;   // distinct coarse rows one coarse column can hold before the kernel gives up
struct TripleParams {
    const i64 *t_colptr, *t_rowval;
    const double* t_nzval;
    int t_base;
    const i64* lump_rowval;      // I(k), 0-based
    const double* lump_nzval;    // w_k
    const i64 *spray_colptr, *spray_rowval;   // members of component J, ascending, 0-based
    i64 Nc;
};
template <int PASS>
__global__ void __launch_bounds__(128) k_triple(TripleParams P, uint32_t* __restrict__ count, const i64* __restrict__ start, int base,
                                                i64* __restrict__ out_colptr, i64* __restrict__ out_rowval,
                                                double* __restrict__ out_nzval, int* __restrict__ overflow) {
    const i64 J = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (J > P.Nc) return;
    if (J == P.Nc) {
        if (PASS == 0) count[J] = 0; else out_colptr[J] = start[J] + base;
        return;
    }
    int rows[TC_MAX];
    double vals[TC_MAX];
    int n = 0;
    bool over = false;
    for (i64 e = P.spray_colptr[J]; e < P.spray_colptr[J + 1]; ++e) {
        const i64 j = P.spray_rowval[e];
        // X[:, j]: the coarse rows this fine column touches, each accumulated over T's rows in ascending order
        int xr[8];
        double xv[8];
        int xn = 0;
        for (i64 t = P.t_colptr[j] - P.t_base; t < P.t_colptr[j + 1] - P.t_base; ++t) {
            const i64 k = P.t_rowval[t] - P.t_base;
            const int I = (int)P.lump_rowval[k];
            const double prod = P.lump_nzval[k] * P.t_nzval[t];
            int q = 0;
            while (q < xn && xr[q] != I) ++q;
            if (q == xn) {
                if (xn < 8) xr[xn] = I, xv[xn] = prod, ++xn;      // a column of T holds at most 7 entries
            } else {
                xv[q] = xv[q] + prod;
            }
        }
        // T_c[:, J] += X[:, j] * 1.0
        for (int q = 0; q < xn; ++q) {
            const double term = xv[q] * 1.0;
            int w = 0;
            while (w < n && rows[w] != xr[q]) ++w;
            if (w == n) {
                if (n < TC_MAX) rows[n] = xr[q], vals[n] = term, ++n; else over = true;
            } else {
                vals[w] = vals[w] + term;
            }
        }
    }
    if (over) atomicOr(overflow, 1);
    if constexpr (PASS == 0) {
        count[J] = (uint32_t)n;
        return;
    } else {
    // rows ascending inside the column, like every SparseMatrixCSC
    for (int a = 1; a < n; ++a) {
        const int r = rows[a];
        const double v = vals[a];
        int b = a - 1;
        while (b >= 0 && rows[b] > r) rows[b + 1] = rows[b], vals[b + 1] = vals[b], --b;
        rows[b + 1] = r, vals[b + 1] = v;
    }
    const i64 o = start[J];
    out_colptr[J] = o + base;
    for (int q = 0; q < n; ++q) out_rowval[o + q] = rows[q] + base, out_nzval[o + q] = vals[q];
    }
}

}  // namespace

extern "C" {

int otmb_lump_and_spray_build(otmb_ctx* c, int64_t di, int64_t dj, int64_t dk, const double* vol, const int64_t* t_colptr,
                              const int64_t* t_rowval, int32_t t_index_base, int32_t index_base, int64_t* n_coarse) {
    if (!c || !vol || di < 1 || dj < 1 || dk < 1) return OTMB_ERR_BADARG;
    if (di * dj * dk > BOXMAX) return otmb_fail(c, OTMB_ERR_BADARG, "lump_and_spray: a lumping box holds at most 32 cells");
    if ((index_base != 0 && index_base != 1) || (t_colptr && !t_rowval)) return OTMB_ERR_BADARG;
    OT_TRY(otmb_need(c, c->have_indices, "otmb_makeindices"));
    if (c->sharded) return otmb_fail(c, OTMB_ERR_STATE, "lump_and_spray is not available on a slab context");
    CU_TRY(c, cudaSetDevice(c->device));
    const i64 N = c->N;
    LumpParams P;
    P.g = GridDims{(int)c->nx, (int)c->ny, (int)c->nz, c->topo, (int)c->P, (int)c->M};
    P.di = (int)di; P.dj = (int)dj; P.dk = (int)dk;
    P.nbx = (int)((c->nx + di - 1) / di); P.nby = (int)((c->ny + dj - 1) / dj); P.nbz = (int)((c->nz + dk - 1) / dk);
    P.rank3d = c->rank3d.as<int>();
    const i64 nboxes = (i64)P.nbx * P.nby * P.nbz;
    // T's stored pattern: the caller's, or the T of the last transportmatrix build on this context
    if (t_colptr) {
        // the caller's pattern is checked like a pre-built operator (otmb_set_operator): the two ends of colptr here,
        // monotonicity, row range and order by a kernel over the uploaded arrays
        const char* const bad_t = "lump_and_spray: T's colptr / rowval are not the pattern of an N x N SparseMatrixCSC "
                                  "(N+1 non-decreasing colptr entries from t_index_base, rows ascending inside [0, N))";
        if (t_index_base != 0 && t_index_base != 1) return otmb_fail(c, OTMB_ERR_BADARG, "t_index_base must be 0 or 1");
        const i64 tn = t_colptr[N] - t_index_base;
        if (t_colptr[0] != t_index_base || tn < 0) return otmb_fail(c, OTMB_ERR_BADARG, bad_t);
        CU_TRY(c, c->add_tmp[0].ensure((size_t)(N + 1) * 8));
        CU_TRY(c, c->add_tmp[1].ensure((size_t)(tn + 1) * 8));
        OT_TRY(otmb_h2d(c, c->add_tmp[0].p, t_colptr, (size_t)(N + 1) * 8, c->stream));
        if (tn > 0) OT_TRY(otmb_h2d(c, c->add_tmp[1].p, t_rowval, (size_t)tn * 8, c->stream));
        int verdict = 0;
        OT_TRY(otmb_check_csc_dev(c, c->add_tmp[0].as<i64>(), c->add_tmp[1].as<i64>(), N, tn, t_index_base, &verdict));
        if (verdict) return otmb_fail(c, OTMB_ERR_BADARG, bad_t);
        P.t_colptr = c->add_tmp[0].as<i64>();
        P.t_rowval = c->add_tmp[1].as<i64>();
        P.t_base = t_index_base;
    } else {
        OT_TRY(otmb_need(c, c->have_mat[OTMB_MAT_T], "otmb_transportmatrix_build (or pass T's colptr / rowval)"));
        P.t_colptr = c->colptr[OTMB_MAT_T].as<i64>();
        P.t_rowval = c->rowval[OTMB_MAT_T].as<i64>();
        P.t_base = c->out_base;
    }
    DevBuf* b = c->coo;   // scratch: 0 vol, 1 ncomp, 2 boxbase, 3 cellcomp, 4 size, 5 asym flag
    CU_TRY(c, b[0].ensure((size_t)(N + 1) * 8));
    CU_TRY(c, b[1].ensure((size_t)(nboxes + 1) * 4));
    CU_TRY(c, b[2].ensure((size_t)(nboxes + 1) * 4));
    CU_TRY(c, b[3].ensure((size_t)(N + 1) * 4));
    CU_TRY(c, b[5].ensure(8));
    OT_TRY(otmb_h2d(c, b[0].p, vol, (size_t)N * 8, c->stream));
    CU_TRY(c, cudaMemsetAsync(b[5].p, 0, 8, c->stream));
    OT_TRY(otmb_reset_flags(c));
    const unsigned grid = grid_for(nboxes, 128);
    k_lump_cc<<<grid, 128, 0, c->stream>>>(P, (int)nboxes, b[1].as<uint32_t>(), b[3].as<int>(), b[5].as<int>());
    LAUNCHED(c);
    OT_TRY(otmb_scan_u32(c, b[1].as<uint32_t>(), b[2].as<uint32_t>(), nboxes, &c->flags.as<DevFlags>()->nnz[0]));
    OT_TRY(otmb_fetch_flags(c));
    int asym = 0;
    CU_TRY(c, cudaMemcpy(&asym, b[5].p, 4, cudaMemcpyDeviceToHost));
    if (asym) return otmb_fail(c, OTMB_ERR_BADARG, "lump_and_spray: T's pattern is not symmetric inside a lumping box "
                                                  "(the reference's SimpleGraph(adjacency) throws)");
    const i64 Nc = (i64)c->h_flags->nnz[0];
    c->lump_nc = Nc;
    c->lump_base = index_base;
    CU_TRY(c, b[4].ensure((size_t)(Nc + 1) * 4));
    CU_TRY(c, c->lump[0].ensure((size_t)(N + 1) * 8));      // LUMP colptr
    CU_TRY(c, c->lump[1].ensure((size_t)(N + 1) * 8));      // LUMP rowval
    CU_TRY(c, c->lump[2].ensure((size_t)(N + 1) * 8));      // LUMP nzval
    CU_TRY(c, c->lump[3].ensure((size_t)(Nc + 1) * 8));     // SPRAY colptr
    CU_TRY(c, c->lump[4].ensure((size_t)(N + 1) * 8));      // SPRAY rowval
    CU_TRY(c, c->lump[5].ensure((size_t)(N + 1) * 8));      // SPRAY nzval
    CU_TRY(c, c->lump[6].ensure((size_t)(Nc + 1) * 8));     // vol_c
    CU_TRY(c, cudaMemsetAsync(b[4].p, 0, (size_t)(Nc + 1) * 4, c->stream));
    CU_TRY(c, cudaMemsetAsync(c->lump[6].p, 0, (size_t)(Nc + 1) * 8, c->stream));
    k_lump_sizes<<<grid, 128, 0, c->stream>>>(P, (int)nboxes, b[2].as<uint32_t>(), b[3].as<int>(), b[0].as<double>(),
                                               b[4].as<uint32_t>(), c->lump[6].as<double>());
    LAUNCHED(c);
    OT_TRY(otmb_scan_u32_to_i64(c, b[4].as<uint32_t>(), c->lump[3].as<i64>(), Nc + 1, nullptr));
    k_lump_fill<<<grid, 128, 0, c->stream>>>(P, (int)nboxes, b[2].as<uint32_t>(), b[3].as<int>(), b[0].as<double>(),
                                              c->lump[6].as<double>(), c->lump[3].as<i64>(), 0, c->lump[0].as<i64>(),
                                              c->lump[1].as<i64>(), c->lump[2].as<double>(), c->lump[4].as<i64>(),
                                              c->lump[5].as<double>());
    LAUNCHED(c);
    k_set_last<<<1, 1, 0, c->stream>>>(c->lump[0].as<i64>(), N, N);
    LAUNCHED(c);
    CU_TRY(c, cudaGetLastError());
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    c->have_lump = true;
    if (n_coarse) *n_coarse = Nc;
    return OTMB_OK;
}

int otmb_lump_and_spray_fetch(otmb_ctx* c, int64_t* lump_colptr, int64_t* lump_rowval, double* lump_nzval,
                              int64_t* spray_colptr, int64_t* spray_rowval, double* spray_nzval, double* vol_c) {
    if (!c) return OTMB_ERR_BADARG;
    OT_TRY(otmb_need(c, c->have_lump, "otmb_lump_and_spray_build"));
    CU_TRY(c, cudaSetDevice(c->device));
    const i64 N = c->N, Nc = c->lump_nc;
    auto get = [&](void* dst, const DevBuf& src, size_t bytes) -> cudaError_t {
        return dst && bytes ? cudaMemcpyAsync(dst, src.p, bytes, cudaMemcpyDeviceToHost, c->stream) : cudaSuccess;
    };
    CU_TRY(c, get(lump_colptr, c->lump[0], (size_t)(N + 1) * 8));
    CU_TRY(c, get(lump_rowval, c->lump[1], (size_t)N * 8));
    CU_TRY(c, get(lump_nzval, c->lump[2], (size_t)N * 8));
    CU_TRY(c, get(spray_colptr, c->lump[3], (size_t)(Nc + 1) * 8));
    CU_TRY(c, get(spray_rowval, c->lump[4], (size_t)N * 8));
    CU_TRY(c, get(spray_nzval, c->lump[5], (size_t)N * 8));
    CU_TRY(c, get(vol_c, c->lump[6], (size_t)Nc * 8));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    // device arrays are 0-based; shift the index arrays to the requested base on the way out
    if (c->lump_base) {
        if (lump_colptr) for (i64 q = 0; q <= N; ++q) lump_colptr[q] += 1;
        if (lump_rowval) for (i64 q = 0; q < N; ++q) lump_rowval[q] += 1;
        if (spray_colptr) for (i64 q = 0; q <= Nc; ++q) spray_colptr[q] += 1;
        if (spray_rowval) for (i64 q = 0; q < N; ++q) spray_rowval[q] += 1;
    }
    return OTMB_OK;
}

// T_c = LUMP * T * SPRAY with the LUMP / SPRAY of the last otmb_lump_and_spray_build and the RESIDENT matrix `which`
// (OTMB_MAT_*) of the last transportmatrix build: the coarse operator of the reference's downstream solve
// (test/local_full.jl:161), built on the device, bit-identical to SparseArrays' product.  Two-phase like the other
// sparse results: *n_coarse, *nnz, then otmb_coarsen_fetch (colptr N_c+1, rowval, nzval; index base of the lump build).
int otmb_coarsen_build(otmb_ctx* c, int which, int64_t* n_coarse, int64_t* nnz) {
    if (!c || which < 0 || which > 4) return OTMB_ERR_BADARG;
    OT_TRY(otmb_need(c, c->have_lump, "otmb_lump_and_spray_build"));
    OT_TRY(otmb_need(c, c->have_mat[which], "otmb_transportmatrix_build"));
    CU_TRY(c, cudaSetDevice(c->device));
    const i64 Nc = c->lump_nc;
    TripleParams P;
    P.t_colptr = c->colptr[which].as<i64>();
    P.t_rowval = c->rowval[which].as<i64>();
    P.t_nzval = c->nzval[which].as<double>();
    P.t_base = c->out_base;
    P.lump_rowval = c->lump[1].as<i64>();
    P.lump_nzval = c->lump[2].as<double>();
    P.spray_colptr = c->lump[3].as<i64>();
    P.spray_rowval = c->lump[4].as<i64>();
    P.Nc = Nc;
    DevBuf* b = c->coo;   // scratch: 7 counts, 8 starts, 9 overflow flag
    CU_TRY(c, b[7].ensure((size_t)(Nc + 2) * 4));
    CU_TRY(c, b[8].ensure((size_t)(Nc + 2) * 8));
    CU_TRY(c, b[9].ensure(8));
    CU_TRY(c, cudaMemsetAsync(b[9].p, 0, 8, c->stream));
    OT_TRY(otmb_reset_flags(c));
    const unsigned grid = grid_for(Nc + 1, 128);
    k_triple<0><<<grid, 128, 0, c->stream>>>(P, b[7].as<uint32_t>(), nullptr, 0, nullptr, nullptr, nullptr, b[9].as<int>());
    LAUNCHED(c);
    OT_TRY(otmb_scan_u32_to_i64(c, b[7].as<uint32_t>(), b[8].as<i64>(), Nc + 1, &c->flags.as<DevFlags>()->nnz[0]));
    OT_TRY(otmb_fetch_flags(c));
    int over = 0;
    CU_TRY(c, cudaMemcpy(&over, b[9].p, 4, cudaMemcpyDeviceToHost));
    if (over) return otmb_fail(c, OTMB_ERR_TOO_LARGE, "LUMP*T*SPRAY: a coarse column touches more than 96 coarse rows");
    const i64 total = (i64)c->h_flags->nnz[0];
    CU_TRY(c, c->sp_colptr.ensure((size_t)(Nc + 1) * 8));
    CU_TRY(c, c->sp_rowval.ensure((size_t)(total + 1) * 8));
    CU_TRY(c, c->sp_nzval.ensure((size_t)(total + 1) * 8));
    k_triple<1><<<grid, 128, 0, c->stream>>>(P, nullptr, b[8].as<i64>(), c->lump_base, c->sp_colptr.as<i64>(), c->sp_rowval.as<i64>(),
                                             c->sp_nzval.as<double>(), b[9].as<int>());
    LAUNCHED(c);
    CU_TRY(c, cudaGetLastError());
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    c->sp_n = Nc;
    c->sp_nnz = total;
    if (n_coarse) *n_coarse = Nc;
    if (nnz) *nnz = total;
    return OTMB_OK;
}

int otmb_coarsen_fetch(otmb_ctx* c, int64_t* colptr, int64_t* rowval, double* nzval) { return otmb_sparse_fetch(c, colptr, rowval, nzval); }

}  // extern "C"
