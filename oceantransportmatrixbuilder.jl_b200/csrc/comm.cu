// comm.cu — one matrix sharded over the GPUs of a box: slab plan, NCCL communicator, and the three collective
// steps of the sharded pipeline, all behind the C ABI so that a Julia (or any) host only has to
// hand every rank the same 128-byte id.
//
// The reference is a single process; what is sharded is its loop over wet cells (`for 𝑖 in eachindex(Lwet)`,
// /root/reference/src/matrixbuilding.jl:237, 348, 450).  Exchanges between the ranks:
//   1. all-gather of the owned wet counts  -> global wet-rank offsets            (otmb_sharded_makeindices)
//   2. the bottom-up continuity scan of facefluxes (/root/reference/src/velocities.jl:234-243) is a floating-point
//      recurrence, ϕtop[k] = ((((ϕtop[k+1] + w) + s) - e) - n), which cannot be re-associated bit-exactly: the slabs
//      form a chain in which the rank below hands the plane of its topmost ϕtop values (nx*ny doubles) to the rank
//      above over NVLink (ncclSend / ncclRecv between device buffers, on the context's stream).  The plane is cut
//      into column chunks and PIPELINED: rank r starts on chunk c as soon as rank r+1 has sent it, so the chain
//      costs (R + C - 1) chunk steps instead of R * C                           (otmb_sharded_facefluxes)
//   3. all-gather of status + the five nnz -> CSC entry offsets, and every rank learns of a failure on any rank
//      (otmb_sharded_transportmatrix_build) — no rank is left waiting in a collective for a peer that has raised.
//
// NCCL is loaded with dlopen at the first communicator call (libnccl.so.2: the one already in the process when the
// host program is PyTorch, else the system library), so libotmb.so itself has no link-time dependency on it and
// single-GPU users never load it.
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace {

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    std::string error;
};

NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (!api.handle) {
            api.error = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "not found");
            return;
        }
        auto sym = [&](const char* n) -> void* {
            void* p = dlsym(api.handle, n);
            if (!p && api.error.empty()) api.error = std::string("libnccl lacks ") + n;
            return p;
        };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.Send = (decltype(api.Send))sym("ncclSend");
        api.Recv = (decltype(api.Recv))sym("ncclRecv");
        api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    });
    return &api;
}

#define NCCL_TRY(ctx, api, expr)                                                                          \
    do {                                                                                                  \
        ncclResult_t _r = (expr);                                                                         \
        if (_r != ncclSuccess) {                                                                          \
            char _b[512];                                                                                 \
            snprintf(_b, sizeof(_b), "NCCL error at %s:%d: %s", __FILE__, __LINE__, (api)->GetErrorString(_r)); \
            return otmb_fail((ctx), OTMB_ERR_COMM, _b);                                                   \
        }                                                                                                 \
    } while (0)

int need_api(otmb_ctx* c, NcclApi** out) {
    NcclApi* a = nccl_api();
    if (!a->error.empty()) return otmb_fail(c, OTMB_ERR_COMM, a->error);
    *out = a;
    return OTMB_OK;
}

// all-gather of `count` Int64 per rank through a small device buffer (host values in, host values out)
int allgather_i64(otmb_ctx* c, const int64_t* mine, int count, int64_t* all) {
    if (c->comm_size == 1 || !c->comm) {
        std::copy(mine, mine + count, all);
        return OTMB_OK;
    }
    NcclApi* a = nullptr;
    OT_TRY(need_api(c, &a));
    const size_t n = (size_t)count, R = (size_t)c->comm_size;
    CU_TRY(c, c->comm_buf.ensure((n + n * R) * 8));
    int64_t* d_mine = c->comm_buf.as<int64_t>();
    int64_t* d_all = d_mine + n;
    CU_TRY(c, cudaMemcpyAsync(d_mine, mine, n * 8, cudaMemcpyHostToDevice, c->stream));
    NCCL_TRY(c, a, a->AllGather(d_mine, d_all, n, ncclInt64, (ncclComm_t)c->comm, c->stream));
    CU_TRY(c, cudaMemcpyAsync(all, d_all, n * R * 8, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return OTMB_OK;
}

// position-dependent checksum of a CSC segment: sum over entries of hash(global position, value) mod 2^64.  A sum, so the
// ranks' partial checksums add up to the checksum of the concatenated matrix whatever the number of ranks.
__device__ __forceinline__ u64 mix64(u64 pos, u64 v) {
    u64 z = pos * 0x9E3779B97F4A7C15ull + v;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__global__ void __launch_bounds__(256) k_checksum(const i64* __restrict__ colptr, const i64* __restrict__ rowval,
                                                  const double* __restrict__ nzval, i64 ncols, i64 nnz, int base, i64 col_offset,
                                                  i64 entry_offset, u64* __restrict__ out) {
    u64 a = 0, b = 0, d = 0;
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 q = (i64)blockIdx.x * blockDim.x + threadIdx.x; q < ncols; q += stride)
        a += mix64((u64)(q + col_offset), (u64)(colptr[q] - base + entry_offset));
    for (i64 q = (i64)blockIdx.x * blockDim.x + threadIdx.x; q < nnz; q += stride) {
        b += mix64((u64)(q + entry_offset), (u64)(rowval[q] - base));
        d += mix64((u64)(q + entry_offset), (u64)__double_as_longlong(nzval[q]));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
        d += __shfl_xor_sync(0xffffffffu, d, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(out + 0, a);
        atomicAdd(out + 1, b);
        atomicAdd(out + 2, d);
    }
}

}  // namespace

// inbox layout: P doubles (the carry plane), then per block of 128 columns one flag ("epoch e has landed", written by
// the rank below) and one ack ("epoch e has been read", written by the rank above into the inbox of the rank it acks)
static i64 inbox_blocks(i64 P) { return (P + 127) / 128; }
static size_t inbox_bytes(i64 P) { return (size_t)P * 8 + (size_t)inbox_blocks(P) * 8 + 64; }

// Peer-memory transport of the carry chain: every rank allocates an inbox and exports it (cudaIpcGetMemHandle), the 64-byte
// handles are all-gathered, and every rank maps the inbox of the rank above it.  Any failure (no P2P, IPC refused)
// leaves peer_state = -1 on EVERY rank (the outcome is all-gathered too) and the chain uses NCCL send / recv instead.
static int peer_setup(otmb_ctx* c) {
    if (c->peer_state != 0 && c->peer_P == c->P) return OTMB_OK;
    const int R = c->comm_size, r = c->comm_rank;
    for (void** q : {&c->peer_above, &c->peer_below})
        if (*q) {
            cudaIpcCloseMemHandle(*q);
            *q = nullptr;
        }
    c->peer_state = -1;
    c->peer_P = c->P;
    int64_t mine[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    cudaIpcMemHandle_t h;
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    bool ok = c->peer_inbox.ensure(inbox_bytes(c->P)) == cudaSuccess &&
              cudaMemsetAsync(c->peer_inbox.p, 0, inbox_bytes(c->P), c->stream) == cudaSuccess &&
              cudaStreamSynchronize(c->stream) == cudaSuccess && cudaIpcGetMemHandle(&h, c->peer_inbox.p) == cudaSuccess;
    if (ok) memcpy(mine, &h, 64);
    mine[8] = ok ? 1 : 0;
    cudaGetLastError();
    std::vector<int64_t> all((size_t)R * 9);
    OT_TRY(allgather_i64(c, mine, 9, all.data()));
    bool everyone = true;
    for (int q = 0; q < R; ++q) everyone = everyone && all[(size_t)q * 9 + 8] == 1;
    int64_t opened = 1;
    auto open_inbox = [&](int rank, void** out) {
        cudaIpcMemHandle_t hh;
        memcpy(&hh, &all[(size_t)rank * 9], 64);
        if (cudaIpcOpenMemHandle(out, hh, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            *out = nullptr;
            opened = 0;
        }
    };
    if (everyone && r > 0) open_inbox(r - 1, &c->peer_above);        // carry + flags go up
    if (everyone && r < R - 1) open_inbox(r + 1, &c->peer_below);    // acks go down
    std::vector<int64_t> res((size_t)R);
    OT_TRY(allgather_i64(c, &opened, 1, res.data()));
    for (int q = 0; q < R; ++q) everyone = everyone && res[(size_t)q] == 1;
    c->peer_state = everyone ? 1 : -1;
    c->peer_epoch = 0;
    return OTMB_OK;
}

// The continuity chain on the device, rank R-1 (sea floor) -> rank 0 (surface); everything is enqueued on the context's
// stream, nothing blocks.  nchunks == 0: the peer-memory form when it is available — ONE k_faceflux launch per rank whose
// blocks wait for / raise per-block flags in peer memory (the send is a plain store over NVLink from inside the kernel) —
// else NCCL with 8 chunks; nchunks > 0: NCCL send / recv of that many column chunks, pipelined.
static int enqueue_chain(otmb_ctx* c, int32_t nchunks) {
    const int R = c->comm_size, r = c->comm_rank;
    NcclApi* a = nullptr;
    if (R > 1) {
        OT_TRY(need_api(c, &a));
        OT_TRY(otmb_need(c, c->comm != nullptr, "otmb_comm_init"));
    }
    const bool recv = r < R - 1, send = r > 0;   // rank R-1 owns the sea floor, rank 0 the surface
    OT_TRY(otmb_faceflux_begin(c, c->uv_fill));
    if (R > 1 && nchunks == 0) OT_TRY(peer_setup(c));
    if (R > 1 && nchunks == 0 && c->peer_state == 1) {
        PeerLink link;
        link.epoch = ++c->peer_epoch;
        double* d_in = nullptr;
        double* d_out = nullptr;
        const i64 nb = inbox_blocks(c->P);
        unsigned* my_words = reinterpret_cast<unsigned*>(c->peer_inbox.as<double>() + c->P);
        if (recv) {
            d_in = c->peer_inbox.as<double>();
            link.flag_in = my_words;                                                                     // raised by the rank below
            link.ack_out = reinterpret_cast<unsigned*>(reinterpret_cast<double*>(c->peer_below) + c->P) + nb;   // its ack array
        }
        if (send) {
            d_out = reinterpret_cast<double*>(c->peer_above);
            link.flag_out = reinterpret_cast<unsigned*>(d_out + c->P);
            link.ack_in = my_words + nb;                                                                 // written by the rank above
        }
        return otmb_faceflux_columns(c, c->uv_fill, 0, c->P, d_in, d_out, link);
    }
    if (recv) CU_TRY(c, c->carry[0].ensure((size_t)c->P * 8));
    if (send) CU_TRY(c, c->carry[1].ensure((size_t)c->P * 8));
    double* d_in = recv ? c->carry[0].as<double>() : nullptr;
    double* d_out = send ? c->carry[1].as<double>() : nullptr;
    // column chunks in units of whole thread blocks
    if (nchunks < 1) nchunks = R > 1 ? 8 : 1;
    const i64 blocks = (c->P + 127) / 128;
    nchunks = (int)std::min<i64>(nchunks, blocks);
    for (int q = 0; q < nchunks; ++q) {
        const i64 p0 = std::min<i64>(c->P, blocks * q / nchunks * 128), p1 = std::min<i64>(c->P, blocks * (q + 1) / nchunks * 128);
        if (recv) NCCL_TRY(c, a, a->Recv(d_in + p0, (size_t)(p1 - p0), ncclFloat64, r + 1, (ncclComm_t)c->comm, c->stream));
        OT_TRY(otmb_faceflux_columns(c, c->uv_fill, p0, p1, d_in, d_out));
        if (send) NCCL_TRY(c, a, a->Send(d_out + p0, (size_t)(p1 - p0), ncclFloat64, r - 1, (ncclComm_t)c->comm, c->stream));
    }
    return OTMB_OK;
}

void otmb_comm_release(otmb_ctx* c) {
    for (void** q : {&c->peer_above, &c->peer_below})
        if (*q) {
            cudaIpcCloseMemHandle(*q);
            *q = nullptr;
        }
    c->peer_state = 0;
    if (c->comm) {
        NcclApi* a = nccl_api();
        if (a->CommDestroy) a->CommDestroy((ncclComm_t)c->comm);
        c->comm = nullptr;
    }
    c->comm_rank = 0;
    c->comm_size = 1;
}

extern "C" {

// ---- slab plan (host only, no GPU needed) ----------------------------------------------------------------------
int otmb_plan_slabs(const double* v3D, int64_t nx, int64_t ny, int64_t nz, int32_t nranks, int32_t level_cuts_only,
                    int64_t* row_cuts, int64_t* wet_per_rank) {
    if (!v3D || !row_cuts || nx < 1 || ny < 1 || nz < 1 || nranks < 1) return OTMB_ERR_BADARG;
    const int64_t rows = ny * nz;
    const int64_t units = level_cuts_only ? nz : rows;   // cut candidates
    if (nranks > units) return OTMB_ERR_BADARG;
    // wet cells per grid row (wet <=> !isnan, src/matrixbuilding.jl:14), counted by a few host threads
    std::vector<int64_t> cum((size_t)rows + 1, 0);
    {
        const unsigned nt = std::max(1u, std::min(8u, std::thread::hardware_concurrency()));
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < nt; ++t)
            pool.emplace_back([&, t] {
                for (int64_t r = rows * t / nt; r < rows * (t + 1) / nt; ++r) {
                    const double* row = v3D + r * nx;
                    int64_t wet = 0;
                    for (int64_t i = 0; i < nx; ++i) wet += row[i] == row[i] ? 1 : 0;
                    cum[(size_t)r + 1] = wet;
                }
            });
        for (auto& th : pool) th.join();
    }
    for (int64_t r = 0; r < rows; ++r) cum[(size_t)r + 1] += cum[(size_t)r];
    const int64_t total = cum[(size_t)rows];
    const int64_t step = level_cuts_only ? ny : 1;       // rows per cut candidate
    // cut r at the candidate whose cumulative wet count is closest to r*N/R, every rank keeping at least one unit
    row_cuts[0] = 0;
    for (int32_t r = 1; r < nranks; ++r) {
        const int64_t lo = row_cuts[r - 1] / step + 1, hi = units - (nranks - r);
        const long double target = (long double)total * r / nranks;
        // cum is non-decreasing: binary search for the first candidate at or above the target, then compare with its predecessor
        int64_t a = lo, b = hi;
        while (a < b) {
            const int64_t mid = (a + b) / 2;
            if ((long double)cum[(size_t)(mid * step)] < target) a = mid + 1; else b = mid;
        }
        int64_t best = a;
        if (a > lo) {
            const long double da = fabsl((long double)cum[(size_t)(a * step)] - target), dp = fabsl((long double)cum[(size_t)((a - 1) * step)] - target);
            if (dp <= da) best = a - 1;
        }
        row_cuts[r] = best * step;
    }
    row_cuts[nranks] = rows;
    if (wet_per_rank)
        for (int32_t r = 0; r < nranks; ++r) wet_per_rank[r] = cum[(size_t)row_cuts[r + 1]] - cum[(size_t)row_cuts[r]];
    return OTMB_OK;
}

// ---- communicator -------------------------------------------------------------------------------------------------
int otmb_comm_unique_id(uint8_t id[OTMB_COMM_ID_BYTES]) {
    if (!id) return OTMB_ERR_BADARG;
    static_assert(sizeof(ncclUniqueId) == OTMB_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    NcclApi* a = nccl_api();
    if (!a->error.empty()) return OTMB_ERR_COMM;
    ncclUniqueId u;
    if (a->GetUniqueId(&u) != ncclSuccess) return OTMB_ERR_COMM;
    memcpy(id, &u, sizeof(u));
    return OTMB_OK;
}

int otmb_comm_init(otmb_ctx* c, int32_t nranks, int32_t rank, const uint8_t id[OTMB_COMM_ID_BYTES]) {
    if (!c || nranks < 1 || rank < 0 || rank >= nranks) return OTMB_ERR_BADARG;
    CU_TRY(c, cudaSetDevice(c->device));
    otmb_comm_release(c);
    c->comm_rank = rank;
    c->comm_size = nranks;
    if (nranks == 1) return OTMB_OK;   // nothing to exchange: no NCCL needed
    if (!id) return OTMB_ERR_BADARG;
    NcclApi* a = nullptr;
    OT_TRY(need_api(c, &a));
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    ncclComm_t comm = nullptr;
    NCCL_TRY(c, a, a->CommInitRank(&comm, nranks, u, rank));
    c->comm = comm;
    return OTMB_OK;
}

int otmb_comm_free(otmb_ctx* c) {
    if (!c) return OTMB_ERR_BADARG;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    otmb_comm_release(c);
    return OTMB_OK;
}

// which transport the default (nchunks == 0) carry chain uses on this context: 0 = not decided yet (no chain has run),
// 1 = peer memory (CUDA IPC inboxes, fused kernel), -1 = NCCL send / recv of column chunks
int otmb_comm_chain_transport(otmb_ctx* c, int32_t* transport) {
    if (!c || !transport) return OTMB_ERR_BADARG;
    *transport = c->comm_size > 1 ? c->peer_state : 0;
    return OTMB_OK;
}

int otmb_comm_allgather_i64(otmb_ctx* c, const int64_t* mine, int32_t count, int64_t* all) {
    if (!c || !mine || !all || count < 1) return OTMB_ERR_BADARG;
    CU_TRY(c, cudaSetDevice(c->device));
    return allgather_i64(c, mine, count, all);
}

// ---- 1. indices: local makeindices on the slab, global wet-rank offsets from the ranks' owned counts ----------------
int otmb_sharded_makeindices(otmb_ctx* c, const double* v3D, int64_t* N_global, int64_t* w0, int64_t* n_owned) {
    if (!c || !v3D) return OTMB_ERR_BADARG;
    int64_t nloc = 0;
    int st = otmb_makeindices(c, v3D, &nloc);
    // every rank takes part in the all-gather even if its own step failed, and learns of the failure
    std::vector<int64_t> all((size_t)c->comm_size * 2);
    const int64_t mine[2] = {st, c->ncols};
    OT_TRY(allgather_i64(c, mine, 2, all.data()));
    int64_t before = 0, total = 0;
    for (int r = 0; r < c->comm_size; ++r) {
        if (all[2 * r] != OTMB_OK)
            return r == c->comm_rank ? st : otmb_fail(c, (int)all[2 * r], "rank " + std::to_string(r) + ": makeindices failed: " + otmb_status_string((int)all[2 * r]));
        if (r < c->comm_rank) before += all[2 * r + 1];
        total += all[2 * r + 1];
    }
    if (before + c->ncols + 2 >= ((int64_t)1 << 31)) return otmb_fail(c, OTMB_ERR_TOO_LARGE, "more than 2^31 wet cells");
    if (c->sharded)
        OT_TRY(otmb_set_rank_offset(c, before));
    if (N_global) *N_global = total;
    if (w0) *w0 = before;
    if (n_owned) *n_owned = c->ncols;
    return OTMB_OK;
}

// ---- 2. face fluxes: chunk-pipelined carry chain, rank R-1 (sea floor) -> rank 0 (surface) ------------------------------
int otmb_set_masstransport(otmb_ctx* c, const double* umo, const double* vmo, double fill) {
    if (!c || !umo || !vmo) return OTMB_ERR_BADARG;
    OT_TRY(otmb_need(c, c->have_indices, "otmb_makeindices"));
    CU_TRY(c, cudaSetDevice(c->device));
    OT_TRY(otmb_upload_uv(c, umo, vmo, fill));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return OTMB_OK;
}

int otmb_sharded_facefluxes(otmb_ctx* c, int32_t nchunks, double* east, double* west, double* north, double* south,
                            double* top, double* bottom) {
    if (!c) return OTMB_ERR_BADARG;
    OT_TRY(otmb_need(c, c->have_indices, "otmb_makeindices"));
    OT_TRY(otmb_need(c, c->have_uv, "otmb_set_masstransport"));
    if (c->topo == OTMB_TOPO_UNKNOWN) return otmb_fail(c, OTMB_ERR_UNKNOWN_GRID, otmb_status_string(OTMB_ERR_UNKNOWN_GRID));
    CU_TRY(c, cudaSetDevice(c->device));
    const int R = c->comm_size;
    OT_TRY(enqueue_chain(c, nchunks));
    OT_TRY(otmb_fetch_flags(c));
    // the reference asserts on the WHOLE arrays (src/velocities.jl:199-200): combine the ranks' flags
    std::vector<int64_t> all((size_t)R * 3);
    const int64_t mine[3] = {c->h_flags->any_valid_u, c->h_flags->any_valid_v, c->h_flags->lookback_timeout};
    OT_TRY(allgather_i64(c, mine, 3, all.data()));
    bool any_u = false, any_v = false, timeout = false;
    for (int q = 0; q < R; ++q) any_u |= all[3 * q] != 0, any_v |= all[3 * q + 1] != 0, timeout |= all[3 * q + 2] != 0;
    if (timeout) {
        c->have_phi = false;
        return otmb_fail(c, OTMB_ERR_COMM, "face-flux carry chain: a rank never received the plane of the rank below it");
    }
    if (!any_u || !any_v) {
        c->have_phi = false;
        return otmb_fail(c, OTMB_ERR_ALL_FILL, otmb_status_string(OTMB_ERR_ALL_FILL));
    }
    double* const outs[6] = {east, west, north, south, top, bottom};
    OT_TRY(otmb_faceflux_copy_out(c, outs));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    c->have_phi = true;
    return OTMB_OK;
}

// the device-resident chain alone (no flag read-back, no copies): what a per-month pipeline enqueues between the upload
// of umo / vmo and the assembly.  The all-fill assertion is not evaluated here.
int otmb_sharded_facefluxes_enqueue(otmb_ctx* c, int32_t nchunks) {
    if (!c) return OTMB_ERR_BADARG;
    OT_TRY(otmb_need(c, c->have_indices, "otmb_makeindices"));
    OT_TRY(otmb_need(c, c->have_uv, "otmb_set_masstransport"));
    CU_TRY(c, cudaSetDevice(c->device));
    OT_TRY(enqueue_chain(c, nchunks));
    c->have_phi = true;
    return OTMB_OK;
}

// ---- 3. assembly of this rank's columns + entry offsets; a failure on any rank is reported on every rank ---------------
int otmb_sharded_transportmatrix_build(otmb_ctx* c, const otmb_tm_params* prm, int64_t nnz_local[5], int64_t nnz_before[5],
                                       int64_t nnz_total[5]) {
    if (!c || !prm) return OTMB_ERR_BADARG;
    int64_t mine[6] = {0, 0, 0, 0, 0, 0};
    const int st = otmb_transportmatrix_build(c, prm, mine + 1);
    mine[0] = st;
    std::vector<int64_t> all((size_t)c->comm_size * 6);
    const std::string own_msg = c->err;
    OT_TRY(allgather_i64(c, mine, 6, all.data()));
    // Which error to report when several ranks fail: the one the reference would have raised first — it checks ρ over
    // the whole array before anything else (src/matrixbuilding.jl:233), then Tadv (:39), TκH (:61), TκVML (:90), TκVdeep (:114)
    auto order = [](int code) {
        static const int first[] = {OTMB_ERR_RHO_NAN, OTMB_ERR_DRY_NEIGHBOUR, OTMB_ERR_TADV_NAN, OTMB_ERR_TKH_NAN, OTMB_ERR_TKVML_NAN, OTMB_ERR_TKVDEEP_NAN};
        for (int q = 0; q < 6; ++q)
            if (code == first[q]) return q;
        return 6;
    };
    int worst = -1;
    for (int r = 0; r < c->comm_size; ++r)
        if (all[6 * r] != OTMB_OK && (worst < 0 || order((int)all[6 * r]) < order((int)all[6 * worst]))) worst = r;
    if (worst >= 0) {
        const int code = (int)all[6 * worst];
        if (worst == c->comm_rank) return otmb_fail(c, code, own_msg);
        return otmb_fail(c, code, "rank " + std::to_string(worst) + ": " + otmb_status_string(code));
    }
    for (int m = 0; m < 5; ++m) {
        int64_t before = 0, total = 0;
        for (int r = 0; r < c->comm_size; ++r) {
            if (r < c->comm_rank) before += all[6 * r + 1 + m];
            total += all[6 * r + 1 + m];
        }
        if (nnz_local) nnz_local[m] = mine[1 + m];
        if (nnz_before) nnz_before[m] = before;
        if (nnz_total) nnz_total[m] = total;
    }
    return OTMB_OK;
}

// checksum of the resident result `which` as a segment of the whole matrix (see k_checksum): out[0] over colptr[0..ncols)
// + entry_offset at column positions + col_offset, out[1] over rowval, out[2] over the bit patterns of nzval, positions
// + entry_offset.  Indices are taken relative to the build's index base.
int otmb_result_checksum(otmb_ctx* c, int which, int64_t col_offset, int64_t entry_offset, uint64_t out[3]) {
    if (!c || which < 0 || which > 4 || !out) return OTMB_ERR_BADARG;
    OT_TRY(otmb_need(c, c->have_mat[which], "otmb_transportmatrix_build"));
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, c->comm_buf.ensure(64));
    CU_TRY(c, cudaMemsetAsync(c->comm_buf.p, 0, 24, c->stream));
    k_checksum<<<c->sm_count * 4, 256, 0, c->stream>>>(c->colptr[which].as<i64>(), c->rowval[which].as<i64>(),
                                                       c->nzval[which].as<double>(), c->ncols, c->nnz[which], c->out_base,
                                                       col_offset, entry_offset, c->comm_buf.as<u64>());
    LAUNCHED(c);
    CU_TRY(c, cudaGetLastError());
    CU_TRY(c, cudaMemcpyAsync(out, c->comm_buf.p, 24, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return OTMB_OK;
}

}  // extern "C"
