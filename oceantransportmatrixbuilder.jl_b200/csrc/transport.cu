// transport.cu — C-ABI entry points for transportmatrix and the generic sparse helpers.
// transportmatrix: /root/reference/src/matrixbuilding.jl:128-150.
#include <cstdlib>

#include "common.cuh"

namespace {

int check_flags(otmb_ctx* c, int build) {
    const DevFlags& f = *c->h_flags;
    // same order as the reference raises them: ρ check and Tadv first (:233, :39), then TκH (:61),
    // TκVML (:90), TκVdeep (:114)
    if (build & 2) {
        if (f.nan_rho) return otmb_fail(c, OTMB_ERR_RHO_NAN, otmb_status_string(OTMB_ERR_RHO_NAN));
        if (f.err_dry_neighbour) return otmb_fail(c, OTMB_ERR_DRY_NEIGHBOUR, otmb_status_string(OTMB_ERR_DRY_NEIGHBOUR));
        if (f.nan_adv) return otmb_fail(c, OTMB_ERR_TADV_NAN, otmb_status_string(OTMB_ERR_TADV_NAN));
    }
    if ((build & 4) && f.nan_kh) return otmb_fail(c, OTMB_ERR_TKH_NAN, otmb_status_string(OTMB_ERR_TKH_NAN));
    if ((build & 8) && f.nan_kvml) return otmb_fail(c, OTMB_ERR_TKVML_NAN, otmb_status_string(OTMB_ERR_TKVML_NAN));
    if ((build & 16) && f.nan_kvdeep) return otmb_fail(c, OTMB_ERR_TKVDEEP_NAN, otmb_status_string(OTMB_ERR_TKVDEEP_NAN));
    return OTMB_OK;
}

}  // namespace

extern "C" {

int otmb_transportmatrix_build(otmb_ctx* c, const otmb_tm_params* prm, int64_t nnz_out[5]) {
    if (!c || !prm) return OTMB_ERR_BADARG;
    OT_TRY(otmb_need(c, c->have_indices, "otmb_makeindices"));
    OT_TRY(otmb_need(c, c->have_metrics, "otmb_gridmetrics / otmb_set_gridmetrics"));
    if (prm->index_base != 0 && prm->index_base != 1) return otmb_fail(c, OTMB_ERR_BADARG, "index_base must be 0 or 1");
    int ops = prm->build_mask & 30;
    if (prm->build_mask == 0) ops = 30;
    for (int m = 1; m <= 4; ++m)
        if (!(ops >> m & 1) && !c->preset[m])
            return otmb_fail(c, OTMB_ERR_STATE, "operator excluded from build_mask but not supplied with otmb_set_operator");
    if (ops & 2) OT_TRY(otmb_need(c, c->have_phi, "otmb_facefluxes / otmb_set_facefluxes"));
    if (ops & 8) OT_TRY(otmb_need(c, c->have_mlotst, "otmb_set_mlotst"));
    if (c->topo == OTMB_TOPO_UNKNOWN) return otmb_fail(c, OTMB_ERR_UNKNOWN_GRID, otmb_status_string(OTMB_ERR_UNKNOWN_GRID));
    if (c->sharded) {
        OT_TRY(otmb_need(c, c->have_rank_offset, "otmb_set_rank_offset"));
        if (prm->path != OTMB_PATH_FUSED || ops != 30)
            return otmb_fail(c, OTMB_ERR_BADARG, "a slab context builds all four operators with OTMB_PATH_FUSED");
    }
    if ((ops & 2) && !c->have_rho3d && prm->rho != prm->rho)
        return otmb_fail(c, OTMB_ERR_RHO_NAN, otmb_status_string(OTMB_ERR_RHO_NAN));
    CU_TRY(c, cudaSetDevice(c->device));
    // the fused kernel dereferences every input; give unused ones a valid dummy
    if (!c->have_mlotst) CU_TRY(c, c->mlotst.ensure((size_t)c->P * 8));
    if (!c->have_phi)
        for (int q = 0; q < 6; ++q) CU_TRY(c, c->phi[q].ensure((size_t)c->M * 8));
    for (int m = 1; m <= 4; ++m)
        if ((ops >> m & 1)) c->preset[m] = false;
    if (c->preset[1] || c->preset[2] || c->preset[3] || c->preset[4]) {
        if (c->out_base != prm->index_base)
            return otmb_fail(c, OTMB_ERR_BADARG, "pre-built operators were supplied with a different index_base");
    }
    c->out_base = prm->index_base;
    c->build_serial++;
    OT_TRY(otmb_reset_flags(c));
    CU_TRY(c, cudaEventRecord(c->ev_b0, c->stream));
    const bool all4 = ops == 30;
    int st = OTMB_OK;
    if (c->ncols == 0) {
        // empty ocean: five empty matrices
        for (int m = 0; m < 5; ++m) {
            CU_TRY(c, c->colptr[m].ensure(8));
            i64 b = prm->index_base;
            CU_TRY(c, cudaMemcpyAsync(c->colptr[m].p, &b, 8, cudaMemcpyHostToDevice, c->stream));
            c->nnz[m] = 0;
            c->have_mat[m] = true;
        }
        CU_TRY(c, cudaStreamSynchronize(c->stream));
    } else if (ops == 0) {
        // all four operators were supplied by the caller (:140-143 all skipped): only the sum remains
        OT_TRY(otmb_sum_operators(c, prm->index_base));
    } else if (prm->path == OTMB_PATH_COO) {
        st = otmb_coo_build(c, prm, ops);
        if (st != OTMB_OK) return st;
        OT_TRY(otmb_fetch_flags(c));
        OT_TRY(check_flags(c, ops));
        for (int m = 1; m <= 4; ++m) c->have_mat[m] = true;
        OT_TRY(otmb_sum_operators(c, prm->index_base));
    } else {
        const int build = ops | (all4 ? 1 : 0);
        // OTMB_FUSED_IMPL=v2 selects the earlier unrolled schedule of the same column kernel (A/B measurements)
        static const char* impl = getenv("OTMB_FUSED_IMPL");
        const int ver = impl && impl[0] == 'v' ? atoi(impl + 1) : 4;
        st = prm->path == OTMB_PATH_FUSED2 ? otmb_fused_build(c, prm, build, true)
             : ver == 2                    ? otmb_fused_v2_build(c, prm, build)
                                           : otmb_fused_v4_build(c, prm, build);
        if (st != OTMB_OK) return st;
        CU_TRY(c, cudaEventRecord(c->ev_b1, c->stream));
        OT_TRY(otmb_fetch_flags(c));
        OT_TRY(check_flags(c, ops));
        for (int m = 0; m < 5; ++m)
            if (build >> m & 1) {
                c->nnz[m] = (i64)c->h_flags->nnz[m];
                c->have_mat[m] = true;
            }
        if ((build & 1) && c->h_flags->zero_dropped) {
            // sparse + does not store results equal to zero (:147): rare, handled by a compaction pass
            OT_TRY(otmb_drop_zeros(c, 0, prm->index_base));
            CU_TRY(c, cudaEventRecord(c->ev_b1, c->stream));
        }
        if (!all4) {
            for (int m = 1; m <= 4; ++m) c->have_mat[m] = true;
            OT_TRY(otmb_sum_operators(c, prm->index_base));
        }
    }
    if (!(c->ncols != 0 && prm->path != OTMB_PATH_COO && all4)) CU_TRY(c, cudaEventRecord(c->ev_b1, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    CU_TRY(c, cudaEventElapsedTime(&c->last_build_ms, c->ev_b0, c->ev_b1));
    if (nnz_out)
        for (int m = 0; m < 5; ++m) nnz_out[m] = c->nnz[m];
    return OTMB_OK;
}

// otmb_transportmatrix_fetch / otmb_transportmatrix_fetch_all: fetch.cu

int otmb_set_operator(otmb_ctx* c, int which, int64_t nnz, const int64_t* colptr, const int64_t* rowval,
                      const double* nzval, int32_t index_base) {
    if (!c || which < 1 || which > 4 || nnz < 0 || !colptr || (nnz > 0 && (!rowval || !nzval))) return OTMB_ERR_BADARG;
    OT_TRY(otmb_need(c, c->have_indices, "otmb_makeindices"));
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, c->colptr[which].ensure((size_t)(c->ncols + 1) * 8));
    CU_TRY(c, c->rowval[which].ensure((size_t)(nnz + 1) * 8));
    CU_TRY(c, c->nzval[which].ensure((size_t)(nnz + 1) * 8));
    CU_TRY(c, cudaMemcpyAsync(c->colptr[which].p, colptr, (size_t)(c->ncols + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    if (nnz > 0) {
        CU_TRY(c, cudaMemcpyAsync(c->rowval[which].p, rowval, (size_t)nnz * 8, cudaMemcpyHostToDevice, c->stream));
        CU_TRY(c, cudaMemcpyAsync(c->nzval[which].p, nzval, (size_t)nnz * 8, cudaMemcpyHostToDevice, c->stream));
    }
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    c->nnz[which] = nnz;
    c->build_serial++;
    c->preset[which] = true;
    c->have_mat[which] = true;
    c->out_base = index_base;
    return OTMB_OK;
}

int otmb_sparse_build(otmb_ctx* c, int64_t len, const int64_t* I, const int64_t* J, const double* V, int64_t n,
                      int64_t* nnz) {
    if (!c || len < 0 || n < 0 || (len > 0 && (!I || !J || !V))) return OTMB_ERR_BADARG;
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, c->coo[0].ensure((size_t)(len + 1) * 8));
    CU_TRY(c, c->coo[1].ensure((size_t)(len + 1) * 8));
    CU_TRY(c, c->coo[2].ensure((size_t)(len + 1) * 8));
    if (len > 0) {
        CU_TRY(c, cudaMemcpyAsync(c->coo[0].p, I, (size_t)len * 8, cudaMemcpyHostToDevice, c->stream));
        CU_TRY(c, cudaMemcpyAsync(c->coo[1].p, J, (size_t)len * 8, cudaMemcpyHostToDevice, c->stream));
        CU_TRY(c, cudaMemcpyAsync(c->coo[2].p, V, (size_t)len * 8, cudaMemcpyHostToDevice, c->stream));
    }
    OT_TRY(otmb_reset_flags(c));
    i64 total = 0;
    OT_TRY(otmb_dev_sparse(c, len, c->coo[0].as<i64>(), c->coo[1].as<i64>(), c->coo[2].as<double>(), nullptr, n, 1,
                           c->sp_colptr, c->sp_rowval, c->sp_nzval, &total));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    c->sp_n = n;
    c->sp_nnz = total;
    if (nnz) *nnz = total;
    return OTMB_OK;
}

int otmb_sparse_fetch(otmb_ctx* c, int64_t* colptr, int64_t* rowval, double* nzval) {
    if (!c) return OTMB_ERR_BADARG;
    CU_TRY(c, cudaSetDevice(c->device));
    if (colptr) CU_TRY(c, cudaMemcpyAsync(colptr, c->sp_colptr.p, (size_t)(c->sp_n + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
    if (rowval && c->sp_nnz > 0)
        CU_TRY(c, cudaMemcpyAsync(rowval, c->sp_rowval.p, (size_t)c->sp_nnz * 8, cudaMemcpyDeviceToHost, c->stream));
    if (nzval && c->sp_nnz > 0)
        CU_TRY(c, cudaMemcpyAsync(nzval, c->sp_nzval.p, (size_t)c->sp_nnz * 8, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return OTMB_OK;
}

int otmb_spadd_build(otmb_ctx* c, int64_t n, const int64_t* acp, const int64_t* arv, const double* anz,
                     const int64_t* bcp, const int64_t* brv, const double* bnz, int64_t* nnz) {
    if (!c || n < 0 || !acp || !bcp) return OTMB_ERR_BADARG;
    CU_TRY(c, cudaSetDevice(c->device));
    const i64 na = acp[n] - 1, nb = bcp[n] - 1;
    DevBuf* b = c->add_tmp;
    CU_TRY(c, b[0].ensure((size_t)(n + 1) * 8));
    CU_TRY(c, b[1].ensure((size_t)(na + 1) * 8));
    CU_TRY(c, b[2].ensure((size_t)(na + 1) * 8));
    CU_TRY(c, b[3].ensure((size_t)(n + 1) * 8));
    CU_TRY(c, b[4].ensure((size_t)(nb + 1) * 8));
    CU_TRY(c, b[5].ensure((size_t)(nb + 1) * 8));
    CU_TRY(c, cudaMemcpyAsync(b[0].p, acp, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(c, cudaMemcpyAsync(b[3].p, bcp, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    if (na > 0) {
        CU_TRY(c, cudaMemcpyAsync(b[1].p, arv, (size_t)na * 8, cudaMemcpyHostToDevice, c->stream));
        CU_TRY(c, cudaMemcpyAsync(b[2].p, anz, (size_t)na * 8, cudaMemcpyHostToDevice, c->stream));
    }
    if (nb > 0) {
        CU_TRY(c, cudaMemcpyAsync(b[4].p, brv, (size_t)nb * 8, cudaMemcpyHostToDevice, c->stream));
        CU_TRY(c, cudaMemcpyAsync(b[5].p, bnz, (size_t)nb * 8, cudaMemcpyHostToDevice, c->stream));
    }
    OT_TRY(otmb_reset_flags(c));
    i64 total = 0;
    OT_TRY(otmb_dev_spadd(c, n, 1, b[0].as<i64>(), b[1].as<i64>(), b[2].as<double>(), b[3].as<i64>(), b[4].as<i64>(),
                          b[5].as<double>(), c->sp_colptr, c->sp_rowval, c->sp_nzval, &total));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    c->sp_n = n;
    c->sp_nnz = total;
    if (nnz) *nnz = total;
    return OTMB_OK;
}

int otmb_spadd_fetch(otmb_ctx* c, int64_t* colptr, int64_t* rowval, double* nzval) {
    return otmb_sparse_fetch(c, colptr, rowval, nzval);
}

}  // extern "C"
