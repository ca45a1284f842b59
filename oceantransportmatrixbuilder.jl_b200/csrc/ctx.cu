// ctx.cu — context life cycle, uploads, measurement helpers, and makeindices (K1).
//
// makeindices replaces /root/reference/src/matrixbuilding.jl:10-24.  Device representation of
// the 3D<->1D wet-cell mapping: a packed bit mask (one UInt64 per 64 linear cells — the same
// chunk layout as Julia's BitArray, so wet3D is a straight copy) plus an exclusive prefix of
// the per-word popcounts.  wet index of cell L = wpre[L/64] + popc(mask[L/64] & lowbits(L%64)):
// 12 bytes per 64 cells instead of the reference's 9 bytes per cell Lwet3D array, small
// enough to live in L1/L2 for every neighbour query of the assembly kernels.
#include <cooperative_groups.h>

#include <algorithm>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) k_fill_indices(const u64* __restrict__ mask, const uint32_t* __restrict__ wpre,
                                                      i64 M, i64* __restrict__ Lwet, i64* __restrict__ Lwet3D) {
    const i64 L = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (L >= M) return;
    const bool wet = wet_at(mask, (int)L);
    const int r = rank_at(mask, wpre, (int)L);
    if (Lwet3D) Lwet3D[L] = wet ? (i64)r + 1 : 0;
    if (wet && Lwet) Lwet[r] = L + 1;
}

__global__ void __launch_bounds__(256) k_fill_lwet32(const u64* __restrict__ mask, const uint32_t* __restrict__ wpre,
                                                     i64 L0, i64 L1, int rank_offset, int* __restrict__ lwet,
                                                     int* __restrict__ rank3d) {
    const i64 L = L0 + (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (L >= L1) return;
    const bool wet = wet_at(mask, (int)L);
    const int r = rank_at(mask, wpre, (int)L);
    // the reference's Lwet3D (0-based, -1 = missing), src/matrixbuilding.jl:18-20; global rank when sharded
    rank3d[L] = wet ? r + rank_offset : -1;
    if (wet) lwet[r] = (int)L;
}

// makeindices in ONE launch (src/matrixbuilding.jl:10-24), two phases around a grid-wide barrier (cooperative launch,
// every block resident).  A block owns a contiguous range of 64-cell chunks.
//   phase 1: the one read of v3D — `isnan` -> ballots = the BitArray chunks, written out; the block's wet count
//   grid.sync()
//   phase 2: exclusive prefix of the block counts (a few thousand values, summed by every block for itself), then the
//            block walks its range again FROM THE MASK (1 bit per cell, L2-resident), every warp its own contiguous run
//            of chunks (no block-wide barrier inside the loops): per-chunk prefix, and — when the wet-rank offset is already known (FILL: unsharded contexts) — Lwet3D (`rank3d`,
//            -1 = dry) and the compacted wet list.
// No serial dependency between blocks (a decoupled look-back over 38 000 small blocks ran at 43 % of the HBM roofline on
// the 0.25-degree grid: every block waits for the prefix to travel down the chain), and no second read of v3D.
// Algorithmic bytes: read 8 per cell; write 1/8 + 1/16 per cell, and with FILL 4 per cell + 4 per wet cell.
constexpr int MI_WPW = 4, MI_WARPS = 8, MI_WORDS = MI_WPW * MI_WARPS;
template <bool FILL>
__global__ void __launch_bounds__(32 * MI_WARPS) k_makeindices(const double* __restrict__ v3D, i64 L0, i64 L1, i64 word0, i64 word1,
                                                               u64* __restrict__ mask, uint32_t* __restrict__ wpre,
                                                               int* __restrict__ rank3d, int* __restrict__ lwet, int rank_offset,
                                                               unsigned* __restrict__ block_tot, u64* __restrict__ total) {
    __shared__ unsigned s_wtot[MI_WARPS];   // the warps' wet counts (phase 1), read again in phase 2
    __shared__ unsigned s_part[MI_WARPS];   // partial sums of the lower blocks' counts
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // this block's chunks: `rounds` x MI_WORDS of them, the same split in both phases; inside the block every WARP owns a
    // contiguous run of `rounds` x MI_WPW chunks, so that no loop below needs a block-wide barrier
    const i64 nwords = word1 - word0;
    const i64 rounds_all = (nwords + MI_WORDS - 1) / MI_WORDS;
    const i64 r0 = rounds_all * blockIdx.x / gridDim.x, r1 = rounds_all * (blockIdx.x + 1) / gridDim.x;
    const i64 rounds = r1 - r0;
    const i64 wbegin = word0 + r0 * MI_WORDS + (i64)wid * rounds * MI_WPW;   // first chunk of this warp
    const double dry = __longlong_as_double(0x7ff8000000000000ll);
    // ---------------- phase 1: mask + counts
    unsigned mine = 0;
    for (i64 r = 0; r < rounds; ++r) {
        const i64 wfirst = wbegin + r * MI_WPW;
        // A warp has its eight loads in flight only while it waits for them, not while it ballots and stores: the lines
        // of its next step are requested into L2 now (4 chunks x 512 B = 16 lines, one per lane), so that those loads
        // find them there (0.25-degree grid: 301 -> 241 us; two steps ahead, or the mask words of phase 2 as well: 263 us).
        if (r + 1 < rounds && lane < 4 * MI_WPW) {
            const i64 nxt = (wfirst + MI_WPW) * 64 + (i64)lane * 16;
            if (nxt >= L0 && nxt < L1) asm volatile("prefetch.global.L2 [%0];" ::"l"(v3D + nxt));
        }
        double va[MI_WPW], vb[MI_WPW];
#pragma unroll
        for (int q = 0; q < MI_WPW; ++q) {   // all eight loads of a lane in flight at once
            const i64 a = (wfirst + q) * 64 + lane, b = a + 32;
            va[q] = (wfirst + q < word1 && a >= L0 && a < L1) ? __ldcs(v3D + a) : dry;
            vb[q] = (wfirst + q < word1 && b >= L0 && b < L1) ? __ldcs(v3D + b) : dry;
        }
#pragma unroll
        for (int q = 0; q < MI_WPW; ++q) {
            const unsigned lo = __ballot_sync(0xffffffffu, !isnan(va[q])), hi = __ballot_sync(0xffffffffu, !isnan(vb[q]));
            if (wfirst + q < word1) {
                if (lane == 0) mask[wfirst + q] = (u64)lo | ((u64)hi << 32);
                mine += __popc(lo) + __popc(hi);
            }
        }
    }
    if (lane == 0) s_wtot[wid] = mine;   // (every lane of a warp holds the same count)
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
#pragma unroll
        for (int q = 0; q < MI_WARPS; ++q) t += s_wtot[q];
        block_tot[blockIdx.x] = t;
    }
    __threadfence();
    cooperative_groups::this_grid().sync();
    // ---------------- phase 2: prefix of the block counts, then every warp walks its run again from the mask
    unsigned part = 0;
    for (int b = threadIdx.x; b < (int)blockIdx.x; b += blockDim.x) part += __ldcg(block_tot + b);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) s_part[wid] = part;
    __syncthreads();
    unsigned pre = 0;   // wet cells before this warp's run
#pragma unroll
    for (int q = 0; q < MI_WARPS; ++q) pre += s_part[q] + (q < wid ? s_wtot[q] : 0u);
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
        unsigned t = pre;   // warp 0: the lower blocks' sum
#pragma unroll
        for (int q = 0; q < MI_WARPS; ++q) t += s_wtot[q];
        *total = (u64)t;
    }
    for (i64 r = 0; r < rounds; ++r) {
        const i64 wfirst = wbegin + r * MI_WPW;
        u64 word[MI_WPW];
#pragma unroll
        for (int q = 0; q < MI_WPW; ++q) word[q] = wfirst + q < word1 ? __ldcg(mask + wfirst + q) : 0ull;
#pragma unroll
        for (int q = 0; q < MI_WPW; ++q) {
            const i64 w = wfirst + q;
            if (w < word1) {
                if (lane == 0) wpre[w] = pre;
                if (FILL) {
                    const unsigned lo = (unsigned)word[q], hi = (unsigned)(word[q] >> 32);
                    const i64 a = w * 64 + lane, b = a + 32;
                    const unsigned below = (1u << lane) - 1u;
                    const int ra = (int)(pre + __popc(lo & below)), rb = (int)(pre + __popc(lo) + __popc(hi & below));
                    const bool wa = lo >> lane & 1u, wb = hi >> lane & 1u;
                    if (a >= L0 && a < L1) rank3d[a] = wa ? ra + rank_offset : -1;
                    if (b >= L0 && b < L1) rank3d[b] = wb ? rb + rank_offset : -1;
                    if (wa) lwet[ra] = (int)a;
                    if (wb) lwet[rb] = (int)b;
                }
            }
            pre += __popcll(word[q]);
        }
    }
}

__global__ void k_l2_flush(uint4* __restrict__ buf, i64 n) {
    i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (; i < n; i += stride) buf[i] = make_uint4((unsigned)i, 1u, 2u, 3u);
}

int upload(otmb_ctx* ctx, DevBuf& buf, const void* host, size_t bytes) {
    CU_TRY(ctx, buf.ensure(bytes));
    OT_TRY(otmb_h2d(ctx, buf.p, host, bytes, ctx->stream));
    return OTMB_OK;
}

}  // namespace

// `host` is the caller's FULL (nx,ny,nz) array; the device buffer holds the context's window only (win<T>())
int otmb_upload3d(otmb_ctx* c, DevBuf& buf, const double* host) {
    CU_TRY(c, buf.ensure(c->win_cells() * 8));
    OT_TRY(otmb_h2d(c, buf.p, host + c->L_win0, c->win_cells() * 8, c->stream));
    return OTMB_OK;
}

// number of wet cells with linear index < X (two small reads of the resident mask / prefix)
// (counted from the window's first cell)
static int wet_below(otmb_ctx* c, i64 X, i64* out) {
    if (X >= c->L_win1) {
        *out = c->N;
        return OTMB_OK;
    }
    if (X <= c->L_win0) {
        *out = 0;
        return OTMB_OK;
    }
    u64 word = 0;
    uint32_t pre = 0;
    CU_TRY(c, cudaMemcpyAsync(&word, c->mask_win() + (X >> 6), 8, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaMemcpyAsync(&pre, c->wpre_win() + (X >> 6), 4, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    *out = (i64)pre + __builtin_popcountll(word & ((1ull << (X & 63)) - 1ull));
    return OTMB_OK;
}

static int fill_ranks(otmb_ctx* c) {
    CU_TRY(c, c->lwet.ensure((size_t)(c->N + 1) * 4));
    CU_TRY(c, c->rank3d.ensure((c->win_cells() + 1) * 4));
    k_fill_lwet32<<<grid_for((i64)c->win_cells(), 256), 256, 0, c->stream>>>(c->mask_win(), c->wpre_win(), c->L_win0, c->L_win1,
                                                                             (int)(c->w0 - c->h_up), c->lwet.as<int>(),
                                                                             c->win<int>(c->rank3d));
    LAUNCHED(c);
    CU_TRY(c, cudaGetLastError());
    c->have_rank_offset = true;
    return OTMB_OK;
}

int otmb_need(otmb_ctx* ctx, bool cond, const char* what) {
    if (cond) return OTMB_OK;
    return otmb_fail(ctx, OTMB_ERR_STATE, std::string("missing prerequisite: ") + what);
}

int otmb_reset_flags(otmb_ctx* ctx) {
    ctx->flags_clean = false;   // whoever asked for the reset is about to write them
    CU_TRY(ctx, cudaMemsetAsync(ctx->flags.p, 0, sizeof(DevFlags), ctx->stream));
    return OTMB_OK;
}
int otmb_fetch_flags(otmb_ctx* ctx) {
    CU_TRY(ctx, cudaMemcpyAsync(ctx->h_flags, ctx->flags.p, sizeof(DevFlags), cudaMemcpyDeviceToHost, ctx->stream));
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return OTMB_OK;
}

extern "C" {

int otmb_version(void) { return 100; }

int otmb_device_count(int* count) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    if (count) *count = n;
    return OTMB_OK;
}

const char* otmb_status_string(int s) {
    switch (s) {
        case OTMB_OK: return "ok";
        case OTMB_ERR_TADV_NAN: return "Tadv contains NaNs.";
        case OTMB_ERR_TKH_NAN: return "T\xce\xbaH contains NaNs.";
        case OTMB_ERR_TKVML_NAN: return "T\xce\xbaVML contains NaNs.";
        case OTMB_ERR_TKVDEEP_NAN: return "T\xce\xbaVdeep contains NaNs.";
        case OTMB_ERR_RHO_NAN: return "\xcf\x81 contains NaNs";
        case OTMB_ERR_UNKNOWN_GRID: return "Unknown grid type";
        case OTMB_ERR_ALL_FILL: return "AssertionError: all umo/vmo values are NaN or FillValue";
        case OTMB_ERR_DRY_NEIGHBOUR: return "non-zero flux from a dry or absent neighbour";
        case OTMB_ERR_BADARG: return "bad argument";
        case OTMB_ERR_STATE: return "missing prerequisite call";
        case OTMB_ERR_COMM: return "communicator (NCCL) error";
        case OTMB_ERR_CUDA: return "CUDA error";
        case OTMB_ERR_NO_GPU: return "no usable sm_100 GPU (there is no CPU fallback)";
        case OTMB_ERR_TOO_LARGE: return "grid too large";
        default: return "unknown status";
    }
}

int otmb_create(otmb_ctx** out, int device) {
    if (!out) return OTMB_ERR_BADARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return OTMB_ERR_NO_GPU;
    }
    if (device < 0 || device >= n) return OTMB_ERR_BADARG;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return OTMB_ERR_NO_GPU;
    if (prop.major != 10) return OTMB_ERR_NO_GPU;  // the kernels are sm_100a SASS only
    otmb_ctx* c = new otmb_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&c->ev_t0) != cudaSuccess || cudaEventCreate(&c->ev_t1) != cudaSuccess ||
        cudaEventCreate(&c->ev_b0) != cudaSuccess || cudaEventCreate(&c->ev_b1) != cudaSuccess ||
        c->flags.ensure(sizeof(DevFlags)) != cudaSuccess ||
        cudaMallocHost((void**)&c->h_flags, sizeof(DevFlags)) != cudaSuccess ||
        cudaHostAlloc((void**)&c->h_done, sizeof(otmb_ctx::HostDone) * otmb_ctx::DONE_RING, cudaHostAllocMapped) != cudaSuccess ||
        c->run_nnz.ensure(64) != cudaSuccess ||
        cudaHostGetDevicePointer((void**)&c->d_done, c->h_done, 0) != cudaSuccess) {
        cudaGetLastError();
        delete c;
        return OTMB_ERR_CUDA;
    }
    cudaMemset(c->flags.p, 0, sizeof(DevFlags));
    memset(c->h_done, 0, sizeof(otmb_ctx::HostDone) * otmb_ctx::DONE_RING);
    c->flags_clean = true;
    *out = c;
    return OTMB_OK;
}

int otmb_destroy(otmb_ctx* c) {
    if (!c) return OTMB_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    otmb_comm_release(c);
    c->peer_inbox.release();
    c->comm_buf.release();
    c->run_nnz.release();
    DevBuf* bufs[] = {&c->v3D, &c->mask, &c->wcount, &c->wpre, &c->lwet, &c->rank3d, &c->area2D, &c->thk, &c->Z3D, &c->zt, &c->edge, &c->dnbr,
                      &c->dedge, &c->lon, &c->lat, &c->lonv, &c->latv, &c->mlotst, &c->rho3d, &c->stage_a, &c->stage_b,
                      &c->flags, &c->tile_state, &c->scan_tmp, &c->sp_colptr, &c->sp_rowval, &c->sp_nzval, &c->l2};
    for (DevBuf* b : bufs) b->release();
    for (int q = 0; q < 6; ++q) c->phi[q].release();
    for (int q = 0; q < 2; ++q) c->carry[q].release();
    for (int q = 0; q < 7; ++q) c->lump[q].release();
    for (int q = 0; q < 5; ++q)
        for (int r = 0; r < 3; ++r) c->tp[q][r].release();
    c->spmv_x.release();
    c->spmv_y.release();
    for (int q = 0; q < 6; ++q) c->add_tmp[q].release();
    for (int m = 0; m < 5; ++m)
        for (int q = 0; q < 3; ++q) c->held[m][q].release();
    c->held_diff.release();
    for (int q = 0; q < 12; ++q) c->coo[q].release();
    for (int q = 0; q < 5; ++q) {
        c->colptr[q].release();
        c->rowval[q].release();
        c->nzval[q].release();
    }
    if (c->fetch_state && c->fetch_state_free) c->fetch_state_free(c->fetch_state);
    if (c->h_flags) cudaFreeHost(c->h_flags);
    if (c->h_done) cudaFreeHost(c->h_done);
    cudaEventDestroy(c->ev_t0);
    cudaEventDestroy(c->ev_t1);
    cudaEventDestroy(c->ev_b0);
    cudaEventDestroy(c->ev_b1);
    cudaStreamDestroy(c->stream);
    delete c;
    return OTMB_OK;
}

const char* otmb_last_error(const otmb_ctx* c) { return c ? c->err.c_str() : "null context"; }

// Page-locked host memory with a process-wide pool.  Pinning a gigabyte costs on the order of 100 ms, and a host shim
// that wraps results in garbage-collected arrays (the Julia shim: unsafe_wrap + finalizer) allocates the same sizes month
// after month: freed blocks are kept (up to POOL_MAX bytes) and handed out again to requests they fit within 1.5x.
namespace {
struct HostPoolState {
    std::mutex mu;
    std::vector<std::pair<size_t, void*>> free_blocks;     // (capacity, pointer)
    std::unordered_map<void*, size_t> live;                // capacity of every block handed out
    size_t pooled = 0;
    static constexpr size_t POOL_MAX = (size_t)8 << 30;
};
HostPoolState& host_pool() {
    static HostPoolState* p = new HostPoolState();   // never destroyed: finalizers may run after static destructors
    return *p;
}
}  // namespace

int otmb_host_alloc(void** ptr, int64_t bytes) {
    if (!ptr || bytes < 0) return OTMB_ERR_BADARG;
    const size_t need = (size_t)(bytes > 0 ? bytes : 8);
    HostPoolState& hp = host_pool();
    {
        std::lock_guard<std::mutex> lk(hp.mu);
        int best = -1;
        for (int q = 0; q < (int)hp.free_blocks.size(); ++q) {
            const size_t cap = hp.free_blocks[q].first;
            if (cap >= need && cap <= need + need / 2 + ((size_t)1 << 20) && (best < 0 || cap < hp.free_blocks[best].first)) best = q;
        }
        if (best >= 0) {
            *ptr = hp.free_blocks[best].second;
            hp.live[*ptr] = hp.free_blocks[best].first;
            hp.pooled -= hp.free_blocks[best].first;
            hp.free_blocks.erase(hp.free_blocks.begin() + best);
            return OTMB_OK;
        }
    }
    if (cudaMallocHost(ptr, need) != cudaSuccess) {
        cudaGetLastError();
        otmb_host_trim();   // give the pooled blocks back and try once more
        if (cudaMallocHost(ptr, need) != cudaSuccess) {
            cudaGetLastError();
            *ptr = nullptr;
            return OTMB_ERR_CUDA;
        }
    }
    std::lock_guard<std::mutex> lk(hp.mu);
    hp.live[*ptr] = need;
    return OTMB_OK;
}
int otmb_host_free(void* ptr) {
    if (!ptr) return OTMB_OK;
    HostPoolState& hp = host_pool();
    {
        std::lock_guard<std::mutex> lk(hp.mu);
        auto it = hp.live.find(ptr);
        if (it == hp.live.end()) return OTMB_ERR_BADARG;   // not a block of otmb_host_alloc (or freed twice)
        const size_t cap = it->second;
        hp.live.erase(it);
        if (hp.pooled + cap <= HostPoolState::POOL_MAX) {
            hp.free_blocks.emplace_back(cap, ptr);
            hp.pooled += cap;
            return OTMB_OK;
        }
    }
    if (cudaFreeHost(ptr) != cudaSuccess) {
        cudaGetLastError();
        return OTMB_ERR_CUDA;
    }
    return OTMB_OK;
}
int otmb_host_trim(void) {
    HostPoolState& hp = host_pool();
    std::vector<std::pair<size_t, void*>> blocks;
    {
        std::lock_guard<std::mutex> lk(hp.mu);
        blocks.swap(hp.free_blocks);
        hp.pooled = 0;
    }
    for (auto& b : blocks)
        if (cudaFreeHost(b.second) != cudaSuccess) cudaGetLastError();
    return OTMB_OK;
}

int otmb_set_grid(otmb_ctx* c, int64_t nx, int64_t ny, int64_t nz, int topology) {
    if (!c) return OTMB_ERR_BADARG;
    if (nx < 1 || ny < 1 || nz < 1) return otmb_fail(c, OTMB_ERR_BADARG, "grid dimensions must be positive");
    if (topology < 0 || topology > 2) return otmb_fail(c, OTMB_ERR_BADARG, "bad topology tag");
    const long double m = (long double)nx * ny * nz;
    if (m >= 2147483000.0L) return otmb_fail(c, OTMB_ERR_TOO_LARGE, "nx*ny*nz must be below 2^31");
    CU_TRY(c, cudaSetDevice(c->device));
    c->nx = nx;
    c->ny = ny;
    c->nz = nz;
    c->P = nx * ny;
    c->M = nx * ny * nz;
    c->nwords = (c->M + 63) / 64;
    c->topo = topology;
    c->have_grid = true;
    c->have_indices = c->have_metrics = c->have_phi = c->have_mlotst = c->have_rho3d = c->have_z3d = c->have_lonlat = false;
    for (int q = 0; q < 5; ++q) c->have_mat[q] = c->preset[q] = false;
    c->N = 0;
    c->sharded = c->have_rank_offset = c->have_uv = false;
    c->L_own0 = c->L_win0 = 0;
    c->L_own1 = c->L_win1 = c->M;
    c->ncols = c->h_up = c->w0 = 0;
    return OTMB_OK;
}

int otmb_set_slab_rows(otmb_ctx* c, int64_t row_begin, int64_t row_end) {
    if (!c) return OTMB_ERR_BADARG;
    OT_TRY(otmb_need(c, c->have_grid, "otmb_set_grid"));
    if (row_begin < 0 || row_end > c->ny * c->nz || row_begin >= row_end)
        return otmb_fail(c, OTMB_ERR_BADARG, "slab must be a non-empty range of grid rows (row = j + ny*k)");
    c->L_own0 = row_begin * c->nx;
    c->L_own1 = row_end * c->nx;
    c->L_win0 = std::max<i64>(0, c->L_own0 - c->P);
    c->L_win1 = std::min<i64>(c->M, c->L_own1 + c->P);
    c->sharded = !(c->L_own0 == 0 && c->L_own1 == c->M);
    c->have_indices = c->have_metrics = c->have_phi = c->have_mlotst = c->have_rho3d = c->have_z3d = c->have_lonlat = false;
    c->have_rank_offset = c->have_uv = false;
    for (int q = 0; q < 5; ++q) c->have_mat[q] = c->preset[q] = false;
    c->N = c->ncols = c->h_up = c->w0 = 0;
    return OTMB_OK;
}

int otmb_set_slab(otmb_ctx* c, int64_t k_begin, int64_t k_end) {
    if (!c) return OTMB_ERR_BADARG;
    OT_TRY(otmb_need(c, c->have_grid, "otmb_set_grid"));
    if (k_begin < 0 || k_end > c->nz || k_begin >= k_end) return otmb_fail(c, OTMB_ERR_BADARG, "slab must be a non-empty level range");
    return otmb_set_slab_rows(c, k_begin * c->ny, k_end * c->ny);
}

int otmb_slab_counts(otmb_ctx* c, int64_t* n_owned, int64_t* n_halo_above) {
    if (!c) return OTMB_ERR_BADARG;
    OT_TRY(otmb_need(c, c->have_indices, "otmb_makeindices"));
    if (n_owned) *n_owned = c->ncols;
    if (n_halo_above) *n_halo_above = c->h_up;
    return OTMB_OK;
}

int otmb_set_rank_offset(otmb_ctx* c, int64_t w0) {
    if (!c || w0 < 0) return OTMB_ERR_BADARG;
    OT_TRY(otmb_need(c, c->have_indices, "otmb_makeindices"));
    if (w0 < c->h_up) return otmb_fail(c, OTMB_ERR_BADARG, "rank offset smaller than the wet count of the halo level above");
    CU_TRY(c, cudaSetDevice(c->device));
    c->w0 = w0;
    OT_TRY(fill_ranks(c));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return OTMB_OK;
}

int otmb_makeindices(otmb_ctx* c, const double* v3D, int64_t* N) {
    if (!c || !v3D) return OTMB_ERR_BADARG;
    OT_TRY(otmb_need(c, c->have_grid, "otmb_set_grid"));
    CU_TRY(c, cudaSetDevice(c->device));
    // Only the window's cells reach this GPU (and only they are scanned); cells outside it read as dry.
    OT_TRY(otmb_upload3d(c, c->v3D, v3D));
    const i64 word0 = c->L_win0 >> 6, word1 = (c->L_win1 + 63) >> 6;
    c->nwords = word1 - word0;
    CU_TRY(c, c->mask.ensure((size_t)(c->nwords + 1) * 8));
    CU_TRY(c, c->wpre.ensure((size_t)(c->nwords + 1) * 4));
    // one cooperative launch: mask, per-chunk prefix, and (unsharded: the rank offset is 0) Lwet3D + the wet list
    OT_TRY(otmb_reset_flags(c));
    c->h_up = 0;
    c->w0 = 0;
    if (!c->sharded) {
        // the wet list cannot be longer than the window; N is only known afterwards
        CU_TRY(c, c->lwet.ensure((c->win_cells() + 1) * 4));
        CU_TRY(c, c->rank3d.ensure((c->win_cells() + 1) * 4));
    }
    {
        void* kern = c->sharded ? (void*)k_makeindices<false> : (void*)k_makeindices<true>;
        int per_sm = 0;
        CU_TRY(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * MI_WARPS, 0));
        const i64 rounds = (c->nwords + MI_WORDS - 1) / MI_WORDS;
        const int blocks = (int)std::max<i64>(1, std::min<i64>(rounds, (i64)std::max(per_sm, 1) * c->sm_count));
        CU_TRY(c, c->scan_tmp.ensure((size_t)(blocks + 1) * 4));
        const double* v = c->win<double>(c->v3D);
        i64 L0 = c->L_win0, L1 = c->L_win1, w0 = word0, w1 = word1;
        u64* m = c->mask_win();
        uint32_t* wp = c->wpre_win();
        int* r3 = c->sharded ? nullptr : c->win<int>(c->rank3d);
        int* lw = c->sharded ? nullptr : c->lwet.as<int>();
        int off = 0;
        unsigned* bt = c->scan_tmp.as<unsigned>();
        u64* tot = &c->flags.as<DevFlags>()->nnz[0];
        void* args[] = {&v, &L0, &L1, &w0, &w1, &m, &wp, &r3, &lw, &off, &bt, &tot};
        CU_TRY(c, cudaLaunchCooperativeKernel(kern, dim3((unsigned)blocks), dim3(32 * MI_WARPS), args, 0, c->stream));
    }
    LAUNCHED(c);
    CU_TRY(c, cudaGetLastError());
    OT_TRY(otmb_fetch_flags(c));
    c->N = (i64)c->h_flags->nnz[0];
    c->ncols = c->N;
    if (c->sharded) {
        i64 below_own = 0, below_end = 0;
        OT_TRY(wet_below(c, c->L_own0, &below_own));
        OT_TRY(wet_below(c, c->L_own1, &below_end));
        c->h_up = below_own;
        c->ncols = below_end - below_own;
        c->have_rank_offset = false;   // global ranks need otmb_set_rank_offset (sum of the lower ranks' counts)
    } else {
        c->have_rank_offset = true;   // Lwet3D and the wet list were filled by the same pass
    }
    c->have_indices = true;
    c->level_cum.clear();
    for (int q = 0; q < 5; ++q) c->have_mat[q] = c->preset[q] = false;
    if (N) *N = c->N;
    return OTMB_OK;
}

int otmb_get_indices(otmb_ctx* c, uint64_t* wet_chunks, int64_t* Lwet, int64_t* Lwet3D) {
    if (!c) return OTMB_ERR_BADARG;
    OT_TRY(otmb_need(c, c->have_indices, "otmb_makeindices"));
    if (c->sharded) return otmb_fail(c, OTMB_ERR_STATE, "otmb_get_indices is not available on a slab context");
    CU_TRY(c, cudaSetDevice(c->device));
    if (wet_chunks)
        CU_TRY(c, cudaMemcpyAsync(wet_chunks, c->mask.p, (size_t)c->nwords * 8, cudaMemcpyDeviceToHost, c->stream));
    if (Lwet || Lwet3D) {
        i64* dLwet = nullptr;
        i64* dL3 = nullptr;
        if (Lwet) {
            CU_TRY(c, c->stage_a.ensure((size_t)(c->N + 1) * 8));
            dLwet = c->stage_a.as<i64>();
        }
        if (Lwet3D) {
            CU_TRY(c, c->stage_b.ensure((size_t)c->M * 8));
            dL3 = c->stage_b.as<i64>();
        }
        k_fill_indices<<<grid_for(c->M, 256), 256, 0, c->stream>>>(c->mask.as<u64>(), c->wpre.as<uint32_t>(), c->M, dLwet,
                                                                    dL3);
        LAUNCHED(c);
        CU_TRY(c, cudaGetLastError());
        if (Lwet) CU_TRY(c, cudaMemcpyAsync(Lwet, dLwet, (size_t)c->N * 8, cudaMemcpyDeviceToHost, c->stream));
        if (Lwet3D) CU_TRY(c, cudaMemcpyAsync(Lwet3D, dL3, (size_t)c->M * 8, cudaMemcpyDeviceToHost, c->stream));
    }
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return OTMB_OK;
}

int otmb_set_facefluxes(otmb_ctx* c, const double* const phi[6]) {
    if (!c || !phi) return OTMB_ERR_BADARG;
    OT_TRY(otmb_need(c, c->have_grid, "otmb_set_grid"));
    CU_TRY(c, cudaSetDevice(c->device));
    for (int q = 0; q < 6; ++q) {
        if (!phi[q]) return otmb_fail(c, OTMB_ERR_BADARG, "null face-flux array");
        OT_TRY(otmb_upload3d(c, c->phi[q], phi[q]));
    }
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    c->have_phi = true;
    return OTMB_OK;
}

int otmb_set_mlotst(otmb_ctx* c, const double* mlotst) {
    if (!c || !mlotst) return OTMB_ERR_BADARG;
    OT_TRY(otmb_need(c, c->have_grid, "otmb_set_grid"));
    CU_TRY(c, cudaSetDevice(c->device));
    OT_TRY(upload(c, c->mlotst, mlotst, (size_t)c->P * 8));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    c->have_mlotst = true;
    return OTMB_OK;
}

int otmb_set_rho3d(otmb_ctx* c, const double* rho3d) {
    if (!c) return OTMB_ERR_BADARG;
    OT_TRY(otmb_need(c, c->have_grid, "otmb_set_grid"));
    CU_TRY(c, cudaSetDevice(c->device));
    if (!rho3d) {
        c->have_rho3d = false;
        return OTMB_OK;
    }
    OT_TRY(otmb_upload3d(c, c->rho3d, rho3d));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    c->have_rho3d = true;
    return OTMB_OK;
}

int otmb_set_gridmetrics(otmb_ctx* c, const double* area2D, const double* thk, const double* zt, const double* edge,
                         const double* dnbr, const double* Z3D, const double* lon, const double* lat) {
    if (!c || !area2D || !thk || !zt || !edge || !dnbr) return OTMB_ERR_BADARG;
    OT_TRY(otmb_need(c, c->have_grid, "otmb_set_grid"));
    CU_TRY(c, cudaSetDevice(c->device));
    OT_TRY(upload(c, c->area2D, area2D, (size_t)c->P * 8));
    OT_TRY(otmb_upload3d(c, c->thk, thk));
    OT_TRY(upload(c, c->zt, zt, (size_t)c->nz * 8));
    OT_TRY(upload(c, c->edge, edge, (size_t)c->P * 32));
    OT_TRY(upload(c, c->dnbr, dnbr, (size_t)c->P * 32));
    if (Z3D) {
        OT_TRY(otmb_upload3d(c, c->Z3D, Z3D));
        c->have_z3d = true;
    }
    if (lon && lat) {
        OT_TRY(upload(c, c->lon, lon, (size_t)c->P * 8));
        OT_TRY(upload(c, c->lat, lat, (size_t)c->P * 8));
        c->have_lonlat = true;
    }
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    c->have_metrics = true;
    return OTMB_OK;
}

int otmb_timer_start(otmb_ctx* c) {
    if (!c) return OTMB_ERR_BADARG;
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, cudaEventRecord(c->ev_t0, c->stream));
    return OTMB_OK;
}
int otmb_timer_stop(otmb_ctx* c, float* ms) {
    if (!c || !ms) return OTMB_ERR_BADARG;
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, cudaEventRecord(c->ev_t1, c->stream));
    CU_TRY(c, cudaEventSynchronize(c->ev_t1));
    CU_TRY(c, cudaEventElapsedTime(ms, c->ev_t0, c->ev_t1));
    return OTMB_OK;
}
int otmb_l2_flush(otmb_ctx* c) {
    if (!c) return OTMB_ERR_BADARG;
    CU_TRY(c, cudaSetDevice(c->device));
    const size_t bytes = (size_t)256 << 20;  // 2x the 126 MB L2
    CU_TRY(c, c->l2.ensure(bytes));
    k_l2_flush<<<c->sm_count * 8, 256, 0, c->stream>>>(c->l2.as<uint4>(), (i64)(bytes / 16));
    CU_TRY(c, cudaGetLastError());
    return OTMB_OK;
}
int otmb_launch_count(otmb_ctx* c, int64_t* launches) {
    if (!c || !launches) return OTMB_ERR_BADARG;
    *launches = c->launches;
    return OTMB_OK;
}
int otmb_set_build_timing(otmb_ctx* c, int32_t on) {
    if (!c) return OTMB_ERR_BADARG;
    c->time_builds = on != 0;
    if (!c->time_builds) c->build_ms_valid = false;
    return OTMB_OK;
}
int otmb_last_build_ms(otmb_ctx* c, float* ms) {
    if (!c || !ms) return OTMB_ERR_BADARG;
    if (c->build_ms_valid) {
        CU_TRY(c, cudaSetDevice(c->device));
        CU_TRY(c, cudaEventSynchronize(c->ev_b1));
        CU_TRY(c, cudaEventElapsedTime(&c->last_build_ms, c->ev_b0, c->ev_b1));
        c->build_ms_valid = false;
    }
    *ms = c->last_build_ms;
    return OTMB_OK;
}
int otmb_synchronize(otmb_ctx* c) {
    if (!c) return OTMB_ERR_BADARG;
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return OTMB_OK;
}

}  // extern "C"
