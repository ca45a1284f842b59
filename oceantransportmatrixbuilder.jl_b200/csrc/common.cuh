// common.cuh — context, error handling, device-side grid helpers shared by all kernels.
// sm_100a only.  Compiled with -fmad=false: the reference never contracts a*b+c into an FMA,
// and every value that can be bit-exact should be.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "otmb.h"

typedef long long i64;
typedef unsigned long long u64;

#define OTMB_SM_COUNT_HINT 148

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap && p) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        if (bytes == 0) bytes = 8;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T* as() const { return reinterpret_cast<T*>(p); }
};

// device-side status block written by kernels, read back once per API call
struct DevFlags {
    int err_dry_neighbour;
    int nan_adv, nan_kh, nan_kvml, nan_kvdeep, nan_rho;
    int any_valid_u, any_valid_v;
    int generic_columns;   // columns that took the coincidence (generic) branch
    int zero_dropped;      // an exactly-zero T entry was stored and must be compacted away
    unsigned spare;
    int lookback_timeout;  // k_fused_v4: a look-back spun past its bound (diagnostic instead of a hang)
    int pad[4];
    u64 nnz[5];
    u64 ticket;            // dynamic tile id for the look-back kernel
    u64 pad2[2];
};

struct otmb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr, ev_b0 = nullptr, ev_b1 = nullptr;
    std::string err;
    int sm_count = OTMB_SM_COUNT_HINT;

    i64 nx = 0, ny = 0, nz = 0, P = 0, M = 0, N = 0, nwords = 0;
    // Row-slab sharding of one matrix (otmb_set_slab_rows / otmb_set_slab).  A grid row is R = j + ny*k, so a range of
    // rows [row0, row1) is a contiguous range of linear cells [L_own0, L_own1) AND — wet ranks being ordered by linear
    // index — a contiguous block of rows/columns of every matrix.  The context keeps one level (P cells) of halo on
    // either side resident: the window [L_win0, L_win1).  3-D arrays are allocated for the window only and handed
    // to the kernels through pointers biased by -L_win0 (win<T>()), so every kernel keeps using GLOBAL linear indices.
    // Unsharded: own = window = [0, M).
    // N = wet cells in the window (local ranks 0..N), ncols = owned wet cells = columns this context assembles,
    // h_up = wet cells of the halo above the owned range, w0 = global wet rank of the first owned cell.
    bool sharded = false, have_rank_offset = false;
    i64 L_own0 = 0, L_own1 = 0, L_win0 = 0, L_win1 = 0, ncols = 0, h_up = 0, w0 = 0;
    template <typename T>
    T* win(const DevBuf& b) const { return b.as<T>() - L_win0; }                      // index with a global linear cell index
    u64* mask_win() const { return mask.as<u64>() - (L_win0 >> 6); }                  // index with (L >> 6)
    uint32_t* wpre_win() const { return wpre.as<uint32_t>() - (L_win0 >> 6); }
    size_t win_cells() const { return (size_t)(L_win1 - L_win0); }
    // NCCL communicator of a sharded run (comm.cu; the library is dlopen'ed, ncclComm_t kept opaque here)
    void* comm = nullptr;
    int comm_rank = 0, comm_size = 1;
    DevBuf comm_buf;          // small device staging for the integer all-gathers
    // peer-memory carry chain (comm.cu): this rank's inbox (carry plane + per-block flags, exported with CUDA IPC) and
    // the mapped inbox of the rank above; peer_state: 0 = not tried, 1 = ready, -1 = unavailable (NCCL chunks instead)
    DevBuf peer_inbox;
    void* peer_above = nullptr;
    void* peer_below = nullptr;
    int peer_state = 0;
    i64 peer_P = 0;
    unsigned peer_epoch = 0;
    bool have_uv = false;     // umo / vmo are resident in stage_a / stage_b (otmb_set_masstransport)
    double uv_fill = 0.0;
    int topo = OTMB_TOPO_UNKNOWN;
    bool have_grid = false, have_indices = false, have_metrics = false, have_phi = false, have_mlotst = false,
         have_rho3d = false, have_z3d = false, have_lonlat = false;

    DevBuf v3D, mask, wcount, wpre, lwet, rank3d, area2D, thk, Z3D, zt, edge, dnbr, dedge, lon, lat, lonv, latv, mlotst, rho3d;
    DevBuf phi[6];
    DevBuf carry[2];          // (nx,ny) planes handed between k-slabs by the continuity scan
    DevBuf stage_a, stage_b;  // generic staging (uploads for facefluxes / Redi-GM inputs)

    // results
    DevBuf colptr[5], rowval[5], nzval[5];
    i64 nnz[5] = {0, 0, 0, 0, 0};
    bool have_mat[5] = {false, false, false, false, false};
    bool preset[5] = {false, false, false, false, false};
    int out_base = 1;

    // scratch
    DevBuf flags;        // DevFlags
    DevFlags* h_flags = nullptr;  // pinned
    // completion record of k_fused_v4 in mapped pinned memory: the last tile stores the flags + nnz there and
    // then the launch's serial number, the host polls it (no memcpy, no stream synchronise per build)
    struct HostDone {
        volatile u64 seq;
        u64 pad[7];
        DevFlags snap;
    };
    static constexpr int DONE_RING = 64;   // launches whose records may be unread at once (slab-pipelined builds)
    HostDone* h_done = nullptr;   // DONE_RING records; launch `serial` uses record serial % DONE_RING
    HostDone* d_done = nullptr;   // device alias of h_done
    DevBuf run_nnz;               // 5 running entry totals on the device: chained slab launches of one build
    std::vector<long long> level_cum;   // wet cells above level k (nz+1 entries), cached per makeindices (slab plans)
    bool flags_clean = false;     // the device flag block is known to be all zero (k_fused_v4 re-zeroes it itself)
    u64 v4_serial = 0;            // launches of k_fused_v4 on this context (epoch of the look-back descriptors)
    size_t ts_zeroed = 0;         // bytes of tile_state known to hold no descriptor of a conflicting epoch
    DevBuf tile_state;   // look-back descriptors / block totals
    DevBuf scan_tmp;     // block sums for the generic scan
    DevBuf coo[12];      // COO path scratch
    DevBuf sp_colptr, sp_rowval, sp_nzval;  // results of otmb_sparse_build / otmb_spadd_build
    i64 sp_n = 0, sp_nnz = 0;
    DevBuf add_tmp[6];
    DevBuf held[5][3];   // caller-supplied operators set aside while a build checks them against a full rebuild
    DevBuf held_diff;    // one int: the comparison's verdict
    DevBuf tp[5][3];     // transposes of the result matrices (CSC of Xᵀ, 0-based), built on demand by otmb_spmv
    i64 tp_serial[5] = {-1, -1, -1, -1, -1};
    i64 build_serial = 0; // bumped by every transportmatrix build / set_operator
    DevBuf spmv_x, spmv_y;
    DevBuf lump[7];      // lump_and_spray results: LUMP colptr/rowval/nzval, SPRAY colptr/rowval/nzval, vol_c
    i64 lump_nc = 0;
    int lump_base = 0;
    bool have_lump = false;
    DevBuf l2;

    // host-side fetch pipeline (fetch.cu): aux stream, pinned ring, widening threads; created on first use
    void* fetch_state = nullptr;
    void (*fetch_state_free)(void*) = nullptr;

    i64 launches = 0;
    float last_build_ms = 0.f;
    bool build_ms_valid = false;  // ev_b0 / ev_b1 bracket a finished build whose duration has not been read yet
    bool time_builds = true;      // record that event pair around every build (otmb_set_build_timing)
};

inline int otmb_fail(otmb_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg;
    return code;
}

#define CU_TRY(ctx, expr)                                                                            \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess) {                                                                     \
            char _b[512];                                                                            \
            snprintf(_b, sizeof(_b), "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__,   \
                     __LINE__, cudaGetErrorString(_e));                                              \
            return otmb_fail((ctx), OTMB_ERR_CUDA, _b);                                              \
        }                                                                                            \
    } while (0)

#define OT_TRY(expr)                  \
    do {                              \
        int _s = (expr);              \
        if (_s != OTMB_OK) return _s; \
    } while (0)

#define LAUNCHED(ctx) ((ctx)->launches++)

inline unsigned grid_for(i64 n, int block) { return (unsigned)((n + block - 1) / block); }

// ----------------------------------------------------------------------------------------
// device-side grid helpers
// ----------------------------------------------------------------------------------------
struct GridDims {
    int nx, ny, nz, topo;
    int P;   // nx*ny
    int M;   // nx*ny*nz  (< 2^31, checked in otmb_set_grid)
};

__device__ __forceinline__ bool wet_at(const u64* __restrict__ mask, int L) {
    return (__ldg(mask + (L >> 6)) >> (L & 63)) & 1ull;
}
// 0-based wet index of linear cell L (valid when wet_at(L))
__device__ __forceinline__ int rank_at(const u64* __restrict__ mask, const uint32_t* __restrict__ wpre, int L) {
    u64 w = __ldg(mask + (L >> 6));
    return (int)__ldg(wpre + (L >> 6)) + __popcll(w & ((1ull << (L & 63)) - 1ull));
}

// Julia's min/max for Float64: NaN-propagating, -0.0 < +0.0
__device__ __forceinline__ double jl_min(double x, double y) {
    double diff = x - y;
    double r = signbit(diff) ? x : y;
    return (isnan(x) || isnan(y)) ? diff : r;
}
__device__ __forceinline__ double jl_max(double x, double y) {
    double diff = x - y;
    double r = signbit(diff) ? y : x;
    return (isnan(x) || isnan(y)) ? diff : r;
}

// scans (scan.cu)
int otmb_scan_u32(otmb_ctx* ctx, const uint32_t* in, uint32_t* out, i64 n, u64* total_dev /* may be null */);
int otmb_scan_i64(otmb_ctx* ctx, const i64* in, i64* out, i64 n, u64* total_dev);
int otmb_scan_u32_to_i64(otmb_ctx* ctx, const uint32_t* in, i64* out, i64 n, u64* total_dev);

// internal entry points across translation units
int otmb_need(otmb_ctx* ctx, bool cond, const char* what);
int otmb_check_csc_dev(otmb_ctx* ctx, const i64* colptr, const i64* rowval, i64 n, i64 nnz, i64 base, int* verdict);   // transport.cu
int otmb_h2d(otmb_ctx* ctx, void* dst, const void* src, size_t bytes, cudaStream_t st);   // fetch.cu: pageable sources are staged by the host pool
int otmb_upload3d(otmb_ctx* ctx, DevBuf& buf, const double* host);   // whole (nx,ny,nz) array, or only the slab window
int otmb_fused_build(otmb_ctx* ctx, const otmb_tm_params* prm, int mask, bool two_pass);
// columns [col0, col0 + ncols) of the context's owned columns (ncols < 0: all of them).  chain: 0 = a build of its
// own; 1 = first launch of a chained build (running entry totals start at zero), 2 = continues the chain: the
// launch's entries follow those of the launches before it (slab-pipelined builds, stream.cu)
int otmb_fused_v4_build(otmb_ctx* ctx, const otmb_tm_params* prm, int mask, i64 col0 = 0, i64 ncols = -1, int chain = 0);
// enqueue the completion record of the last k_fused_v4 launch (totals_out: device copy of the five running nnz, or null)
int otmb_v4_publish(otmb_ctx* ctx, u64* totals_out = nullptr);
int otmb_wait_v4(otmb_ctx* ctx, u64 serial, bool block);   // 0 = done (flags in h_flags), -1 = not yet (block == false)
int otmb_check_build_flags(otmb_ctx* ctx, int ops);
int otmb_drop_zeros(otmb_ctx* ctx, int m, int base);
int otmb_coo_build(otmb_ctx* ctx, const otmb_tm_params* prm, int mask);
int otmb_sum_operators(otmb_ctx* ctx, int base);
int otmb_dev_sparse(otmb_ctx* ctx, i64 len, const i64* dI, const i64* dJ, const double* dV, const int* dValid, i64 n,
                    int base, DevBuf& colptr, DevBuf& rowval, DevBuf& nzval, i64* nnz);
int otmb_dev_spadd(otmb_ctx* ctx, i64 n, int base, const i64* acp, const i64* arv, const double* anz, const i64* bcp,
                   const i64* brv, const double* bnz, DevBuf& colptr, DevBuf& rowval, DevBuf& nzval, i64* nnz);
int otmb_faceflux_begin(otmb_ctx* ctx, double fill);
// peer-memory link of the face-flux carry chain: per-block flags (one per 128 columns) in this GPU's memory (written by
// the rank below) and in the memory of the rank above (mapped with CUDA IPC), and the value that means "this launch"
// ack_in (this GPU, written by the rank above) / ack_out (the rank below): "your plane of epoch e has been read" — the
// sender may not overwrite a block of the inbox before the receiver has taken the previous epoch's values out of it
struct PeerLink {
    const unsigned* flag_in = nullptr;
    unsigned* flag_out = nullptr;
    const unsigned* ack_in = nullptr;
    unsigned* ack_out = nullptr;
    unsigned epoch = 0;
};
int otmb_faceflux_columns(otmb_ctx* ctx, double fill, i64 p_begin, i64 p_end, const double* d_in, double* d_out,
                          PeerLink link = PeerLink());
int otmb_faceflux_copy_out(otmb_ctx* ctx, double* const outs[6]);
int otmb_upload_uv(otmb_ctx* ctx, const double* umo, const double* vmo, double fill);
void otmb_comm_release(otmb_ctx* ctx);
int otmb_fetch_flags(otmb_ctx* ctx);
int otmb_reset_flags(otmb_ctx* ctx);
