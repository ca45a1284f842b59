// spmv.cu — y = X x and y = Xᵀ x on a RESIDENT result matrix (consumers of T that
// do not need the 0.99 GB device-to-host copy first).  These are the products the reference's own checks
// use: τdiv = ‖1‖ / ‖T 1‖ and τvol = ‖v‖ / ‖Tᵀ v‖ (/root/reference/test/online.jl:110-115).
//
// Both are gather kernels, one thread per output element, adding in a fixed order, so the result is
// deterministic and bit-identical to a sequential CSC product: Xᵀ x walks a column of X (rows ascending);
// X x walks a column of Xᵀ, which is built once per matrix with the device `sparse` (a stable transpose:
// the entries of a row come out in ascending column order — the order in which a column-by-column CSC
// product adds them into y[i]).
#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) k_expand_cols(const i64* __restrict__ colptr, i64 n, int base, i64* __restrict__ J1,
                                                     const i64* __restrict__ rowval, i64* __restrict__ I1) {
    const i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    for (i64 p = colptr[j] - base; p < colptr[j + 1] - base; ++p) {
        J1[p] = j + 1;                    // 1-based column of X  -> row of Xᵀ
        I1[p] = rowval[p] - base + 1;     // 1-based row of X     -> column of Xᵀ
    }
}

__global__ void __launch_bounds__(256) k_csc_dot(const i64* __restrict__ colptr, const i64* __restrict__ rowval,
                                                 const double* __restrict__ nzval, const double* __restrict__ x, i64 n,
                                                 int base, double* __restrict__ y) {
    const i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double s = 0.0;
    for (i64 p = colptr[j] - base; p < colptr[j + 1] - base; ++p) s = s + nzval[p] * x[rowval[p] - base];
    y[j] = s;
}

}  // namespace

extern "C" int otmb_spmv(otmb_ctx* c, int which, int transpose, const double* x, double* y) {
    if (!c || which < 0 || which > 4 || !x || !y) return OTMB_ERR_BADARG;
    OT_TRY(otmb_need(c, c->have_mat[which], "otmb_transportmatrix_build"));
    if (c->sharded) return otmb_fail(c, OTMB_ERR_STATE, "otmb_spmv is not available on a slab context");
    CU_TRY(c, cudaSetDevice(c->device));
    const i64 n = c->N, nnz = c->nnz[which];
    const int base = c->out_base;
    CU_TRY(c, c->spmv_x.ensure((size_t)(n + 1) * 8));
    CU_TRY(c, c->spmv_y.ensure((size_t)(n + 1) * 8));
    OT_TRY(otmb_h2d(c, c->spmv_x.p, x, (size_t)n * 8, c->stream));
    if (transpose) {
        k_csc_dot<<<grid_for(n, 256), 256, 0, c->stream>>>(c->colptr[which].as<i64>(), c->rowval[which].as<i64>(),
                                                            c->nzval[which].as<double>(), c->spmv_x.as<double>(), n, base,
                                                            c->spmv_y.as<double>());
        LAUNCHED(c);
    } else {
        if (c->tp_serial[which] != c->build_serial) {
            // Xᵀ as CSC = the device `sparse` of X's triplets with rows and columns exchanged (stable, no duplicates)
            CU_TRY(c, c->coo[0].ensure((size_t)(nnz + 1) * 8));
            CU_TRY(c, c->coo[1].ensure((size_t)(nnz + 1) * 8));
            k_expand_cols<<<grid_for(n, 256), 256, 0, c->stream>>>(c->colptr[which].as<i64>(), n, base, c->coo[0].as<i64>(),
                                                                    c->rowval[which].as<i64>(), c->coo[1].as<i64>());
            LAUNCHED(c);
            OT_TRY(otmb_reset_flags(c));
            i64 total = 0;
            OT_TRY(otmb_dev_sparse(c, nnz, c->coo[0].as<i64>(), c->coo[1].as<i64>(), c->nzval[which].as<double>(), nullptr, n, 0,
                                   c->tp[which][0], c->tp[which][1], c->tp[which][2], &total));
            if (total != nnz) return otmb_fail(c, OTMB_ERR_STATE, "transpose lost entries");
            c->tp_serial[which] = c->build_serial;
        }
        k_csc_dot<<<grid_for(n, 256), 256, 0, c->stream>>>(c->tp[which][0].as<i64>(), c->tp[which][1].as<i64>(),
                                                            c->tp[which][2].as<double>(), c->spmv_x.as<double>(), n, 0,
                                                            c->spmv_y.as<double>());
        LAUNCHED(c);
    }
    CU_TRY(c, cudaGetLastError());
    CU_TRY(c, cudaMemcpyAsync(y, c->spmv_y.p, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return OTMB_OK;
}
