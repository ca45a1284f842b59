// fetch.cu — results out to the host: otmb_transportmatrix_fetch / otmb_transportmatrix_fetch_all.
//
// The API's colptr / rowval are Int64 (SparseMatrixCSC{Float64,Int64}, /root/reference/src/matrixbuilding.jl:41),
// but every index of a matrix with fewer than 2^31 rows and entries fits 32 bits, and the PCIe link is the
// end-to-end floor (0.99 GB of CSC arrays per ACCESS-ESM1-5 matrix set against 0.4 ms of assembly).  So the indices
// cross the link as Int32 and are widened on the host while the values are still in flight:
//   device : k_narrow  Int64 -> Int32 (colptr and rowval of every requested matrix, one contiguous scratch array)
//   link   : the Int32 array in 16 MB chunks into pinned staging, then nzval straight into the caller's arrays
//   host   : a small pool of threads widens each chunk into the caller's Int64 arrays as soon as its event fires
//            (AVX2 sign extension, non-temporal stores), i.e. while the values are still crossing the link.
// 28 % fewer bytes on the link (16 -> 12 B per entry, 8 -> 4 B per column); the arrays the caller receives are the
// same bits as a direct 8-byte copy.  Matrices whose indices do not fit 32 bits take the direct copy.
#include <algorithm>
#include <atomic>
#include <chrono>
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#include <condition_variable>
#include <cstdlib>
#include <mutex>
#include <thread>
#include <vector>

#include "common.cuh"

namespace {

__global__ void __launch_bounds__(256) k_narrow(const i64* __restrict__ in, int* __restrict__ out, i64 n) {
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (int)in[i];
}

// The slab-pipelined call narrows on the COMPUTE stream, right behind the assembly of a slab, so that the copy-out stream
// carries nothing but copies: the entry ranges are read from the device (the running totals before / after the slab),
// the narrowed arrays keep the global positions.  blockIdx.y: 0-4 colptr of matrix m, 5-9 rowval of matrix m.
struct NarrowSlab {
    const i64* colptr[5];
    const i64* rowval[5];
    int* nar_colptr[5];
    int* nar_rowval[5];
    i64 col0, ncp;          // colptr entries [col0, col0 + ncp)
    const u64* before;      // 5 totals before the slab (device)
    const u64* after;       // 5 totals after it
};
__global__ void __launch_bounds__(256) k_narrow_slab(const __grid_constant__ NarrowSlab S) {
    const int m = blockIdx.y % 5;
    const bool rows = blockIdx.y >= 5;
    const i64 lo = rows ? (i64)S.before[m] : S.col0, hi = rows ? (i64)S.after[m] : S.col0 + S.ncp;
    const i64* __restrict__ in = rows ? S.rowval[m] : S.colptr[m];
    int* __restrict__ out = rows ? S.nar_rowval[m] : S.nar_colptr[m];
    const i64 stride = (i64)gridDim.x * blockDim.x;
    for (i64 i = lo + (i64)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += stride) out[i] = (int)in[i];
}

// host threads that widen Int32 -> Int64 or copy bytes; run() returns when the whole range is done
struct HostPool {
    std::vector<std::thread> th;
    std::mutex mu;
    std::condition_variable cv, cv_done;
    const void* src = nullptr;
    void* dst = nullptr;
    size_t n = 0;
    int kind = 0;   // 0: widen n Int32 -> Int64;  1: copy n bytes
    int gen = 0, pending = 0;
    bool stop = false;

    explicit HostPool(int nthreads) {
        for (int t = 0; t < nthreads; ++t) th.emplace_back([this, t] { work(t); });
    }
    ~HostPool() {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv.notify_all();
        for (auto& t : th) t.join();
    }
    static void widen(const int* s, int64_t* d, size_t lo, size_t hi) {
#if defined(__x86_64__)
        static const bool avx2 = __builtin_cpu_supports("avx2");
        if (avx2) {
            widen_avx2(s, d, lo, hi);
            return;
        }
#endif
        for (size_t i = lo; i < hi; ++i) d[i] = (int64_t)s[i];
    }
#if defined(__x86_64__)
    // sign-extend 8 indices per step; non-temporal stores: the destination is written once and not read here
    __attribute__((target("avx2"))) static void widen_avx2(const int* s, int64_t* d, size_t lo, size_t hi) {
        size_t i = lo;
        for (; i < hi && ((uintptr_t)(d + i) & 31); ++i) d[i] = (int64_t)s[i];
        for (; i + 8 <= hi; i += 8) {
            const __m256i v = _mm256_loadu_si256((const __m256i*)(s + i));
            _mm256_stream_si256((__m256i*)(d + i), _mm256_cvtepi32_epi64(_mm256_castsi256_si128(v)));
            _mm256_stream_si256((__m256i*)(d + i + 4), _mm256_cvtepi32_epi64(_mm256_extracti128_si256(v, 1)));
        }
        for (; i < hi; ++i) d[i] = (int64_t)s[i];
        _mm_sfence();
    }
#endif
    static void slice(int kind, const void* s, void* d, size_t lo, size_t hi) {
        if (kind == 0)
            widen((const int*)s, (int64_t*)d, lo, hi);
        else
            memcpy((char*)d + lo, (const char*)s + lo, hi - lo);
    }
    void work(int id) {
        int seen = 0;
        for (;;) {
            const void* s;
            void* d;
            size_t cnt;
            int k;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return stop || gen != seen; });
                if (stop) return;
                seen = gen;
                s = src, d = dst, cnt = n, k = kind;
            }
            const size_t T = th.size();
            // byte copies are split on 64-byte boundaries
            const size_t lo = k == 0 ? cnt * id / T : (cnt * id / T) & ~(size_t)63;
            const size_t hi = k == 0 ? cnt * (id + 1) / T : ((size_t)id + 1 == T ? cnt : (cnt * (id + 1) / T) & ~(size_t)63);
            slice(k, s, d, lo, hi);
            {
                std::lock_guard<std::mutex> lk(mu);
                if (--pending == 0) cv_done.notify_one();
            }
        }
    }
    void run(int k, const void* s, void* d, size_t cnt) {
        if (cnt < (size_t)1 << 16 || th.empty()) {
            slice(k, s, d, 0, cnt);
            return;
        }
        {
            std::lock_guard<std::mutex> lk(mu);
            src = s, dst = d, n = cnt, kind = k;
            pending = (int)th.size();
            ++gen;
        }
        cv.notify_all();
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return pending == 0; });
    }
};

#ifndef OTMB_CHUNK_MB
#define OTMB_CHUNK_MB 16
#endif
constexpr size_t CHUNK_BYTES = (size_t)OTMB_CHUNK_MB << 20;   // bytes per staged D2H copy / host step
constexpr size_t STAGE_CAP = (size_t)3 << 30;          // bytes of pinned staging at most; beyond: direct copies

struct FetchState {
    std::vector<cudaEvent_t> ev;
    char* stage = nullptr;       // pinned: narrowed indices (and values bound for pageable arrays) of one fetch
    size_t stage_cap = 0;
    DevBuf narrow;
    HostPool* pool = nullptr;
    bool few_threads = false;
    cudaStream_t s_up = nullptr, s_dn = nullptr;   // transfer streams of the slab-pipelined call (stream.cu)
    char* up_stage = nullptr;                      // pinned: UP_BUFS pieces of a pageable upload in flight (otmb_h2d)
    cudaEvent_t up_ev[3] = {};
    cudaEvent_t ev_slab[2 * otmb_ctx::DONE_RING + 1] = {};   // [s] slab uploaded, [RING + s] slab assembled, [2 RING] call start
};

int fetch_state(otmb_ctx* c, FetchState** out) {
    if (!c->fetch_state) {
        FetchState* f = new FetchState();
        c->fetch_state = f;
        c->fetch_state_free = [](void* p) {
            FetchState* f = static_cast<FetchState*>(p);
            delete f->pool;
            for (cudaEvent_t e : f->ev) cudaEventDestroy(e);
            for (cudaEvent_t e : f->ev_slab)
                if (e) cudaEventDestroy(e);
            if (f->s_up) cudaStreamDestroy(f->s_up);
            if (f->s_dn) cudaStreamDestroy(f->s_dn);
            if (f->stage) cudaFreeHost(f->stage);
            if (f->up_stage) cudaFreeHost(f->up_stage);
            for (cudaEvent_t e : f->up_ev)
                if (e) cudaEventDestroy(e);
            f->narrow.release();
            delete f;
        };
        // half the cores, shared with the other ranks of this box (torchrun exports LOCAL_WORLD_SIZE)
        int lws = 1;
        if (const char* e = getenv("LOCAL_WORLD_SIZE")) lws = std::max(1, atoi(e));
        int nt = (int)std::min<unsigned>(8u, std::max(1u, std::thread::hardware_concurrency() / (2u * (unsigned)lws)));
        // with fewer than four threads to spare the widening does not keep up with the link (measured: 4 ranks on a
        // 16-core host, 2 threads each: 63 ms per fetch against 61 ms with plain copies) -> the caller copies directly
        f->few_threads = nt < 4;
        if (const char* e = getenv("OTMB_HOST_THREADS")) nt = std::max(0, atoi(e)), f->few_threads = false;
        f->pool = new HostPool(nt);
    }
    *out = static_cast<FetchState*>(c->fetch_state);
    return OTMB_OK;
}

bool is_pageable(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return at.type == cudaMemoryTypeUnregistered;
}

}  // namespace

// Host array -> device on stream `st`.  A page-locked source goes straight to the copy engine.  A pageable one the driver
// would stage through its own bounce buffer with one thread (measured: 290 MB of pre-built operators in 64 ms, 4.5 GB/s);
// here the pool's threads copy 16 MB pieces into three page-locked buffers while the earlier pieces are on the link.
// As with any copy from pageable memory the source has been read in full when this returns.
int otmb_h2d(otmb_ctx* c, void* dst, const void* src, size_t bytes, cudaStream_t st) {
    constexpr int UP_BUFS = 3;
    FetchState* f = nullptr;
    if (bytes >= ((size_t)4 << 20) && is_pageable(src) && !getenv("OTMB_UPLOAD_DIRECT")) OT_TRY(fetch_state(c, &f));
    if (f && !f->few_threads && !f->up_stage) {
        if (cudaMallocHost((void**)&f->up_stage, UP_BUFS * CHUNK_BYTES) != cudaSuccess) {
            cudaGetLastError();
            f->up_stage = nullptr;
        } else
            for (auto& e : f->up_ev) CU_TRY(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    if (!f || f->few_threads || !f->up_stage) {
        CU_TRY(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
        return OTMB_OK;
    }
    int piece = 0;
    for (size_t lo = 0; lo < bytes; lo += CHUNK_BYTES, ++piece) {
        const size_t cnt = std::min(CHUNK_BYTES, bytes - lo);
        const int b = piece % UP_BUFS;
        char* stage = f->up_stage + (size_t)b * CHUNK_BYTES;
        if (piece >= UP_BUFS) CU_TRY(c, cudaEventSynchronize(f->up_ev[b]));
        f->pool->run(1, (const char*)src + lo, stage, cnt);
        CU_TRY(c, cudaMemcpyAsync((char*)dst + lo, stage, cnt, cudaMemcpyHostToDevice, st));
        CU_TRY(c, cudaEventRecord(f->up_ev[b], st));
    }
    // the buffers are reused by the next call: wait until the link has taken the last pieces
    for (int b = 0; b < std::min(piece, UP_BUFS); ++b) CU_TRY(c, cudaEventSynchronize(f->up_ev[b]));
    return OTMB_OK;
}

namespace {

// One copy-out in flight: device arrays -> caller arrays, on stream `st`.
//   indices : narrowed to Int32 on the device, copied into pinned staging in 16 MB chunks (an event per chunk), widened
//             into the caller's Int64 array by the pool as each chunk lands;
//   values  : straight into the caller's array when that is page-locked; otherwise staged the same way and copied out by
//             the pool (a D2H copy into pageable memory runs at a fraction of the link rate and blocks the host).
// Everything is enqueued by indices() / values(); consume() / drain() do the host half in order.
struct CopyOut {
    struct Chunk {
        cudaEvent_t ev;
        int kind;
        const void* src;
        void* dst;
        size_t n;
    };
    otmb_ctx* c;
    FetchState* f;
    cudaStream_t st;
    bool narrow_ok = false;
    size_t stage_off = 0, narrow_off = 0, ev_used = 0, head = 0;
    std::vector<Chunk> q;

    CopyOut(otmb_ctx* c_, FetchState* f_, cudaStream_t st_) : c(c_), f(f_), st(st_) {}

    // staging for `stage_bytes` of host staging and `narrow_ints` narrowed indices; on failure everything goes direct
    int prepare(bool fits32, size_t narrow_ints, size_t stage_bytes) {
        narrow_ok = false;
        if (!fits32 || f->few_threads || getenv("OTMB_FETCH_DIRECT") || stage_bytes > STAGE_CAP) return OTMB_OK;
        if (stage_bytes > f->stage_cap) {
            if (f->stage) cudaFreeHost(f->stage);
            f->stage = nullptr, f->stage_cap = 0;
            const size_t cap = stage_bytes + stage_bytes / 8 + 4096;
            if (cudaMallocHost((void**)&f->stage, cap) != cudaSuccess) {   // no pinnable memory left: direct copies
                cudaGetLastError();
                f->stage = nullptr;
                return OTMB_OK;
            }
            f->stage_cap = cap;
        }
        CU_TRY(c, f->narrow.ensure(std::max<size_t>(narrow_ints, 1) * sizeof(int)));
        narrow_ok = true;
        return OTMB_OK;
    }
    int next_event(cudaEvent_t* e) {
        if (ev_used == f->ev.size()) {
            cudaEvent_t n;
            CU_TRY(c, cudaEventCreateWithFlags(&n, cudaEventDisableTiming));
            f->ev.push_back(n);
        }
        *e = f->ev[ev_used++];
        return OTMB_OK;
    }
    // dev -> staging in chunks, one host job per chunk
    int staged(const void* dev, void* host, size_t bytes, int kind, size_t elem_out) {
        const char* d = (const char*)dev;
        char* stage = f->stage + stage_off;
        for (size_t lo = 0; lo < bytes; lo += CHUNK_BYTES) {
            const size_t cnt = std::min(CHUNK_BYTES, bytes - lo);
            CU_TRY(c, cudaMemcpyAsync(stage + lo, d + lo, cnt, cudaMemcpyDeviceToHost, st));
            cudaEvent_t e;
            OT_TRY(next_event(&e));
            CU_TRY(c, cudaEventRecord(e, st));
            // widen: cnt/4 Int32 -> Int64 at element offset lo/4;  copy: cnt bytes at byte offset lo
            if (kind == 0)
                q.push_back({e, 0, stage + lo, (char*)host + (lo / 4) * elem_out, cnt / 4});
            else
                q.push_back({e, 1, stage + lo, (char*)host + lo, cnt});
        }
        stage_off += (bytes + 63) & ~(size_t)63;
        return OTMB_OK;
    }
    int indices(const i64* dev, int64_t* host, size_t n) {
        if (n == 0 || !host) return OTMB_OK;
        if (!narrow_ok || stage_off + n * 4 + 64 > f->stage_cap || (narrow_off + n) * sizeof(int) > f->narrow.cap) {
            CU_TRY(c, cudaMemcpyAsync(host, dev, n * 8, cudaMemcpyDeviceToHost, st));
            return OTMB_OK;
        }
        int* nar = f->narrow.as<int>() + narrow_off;
        k_narrow<<<std::min<unsigned>(grid_for((i64)n, 256), (unsigned)c->sm_count * 8), 256, 0, st>>>(dev, nar, (i64)n);
        LAUNCHED(c);
        CU_TRY(c, cudaGetLastError());
        narrow_off += n;
        return staged(nar, host, n * 4, 0, 8);
    }
    // indices that were narrowed already (k_narrow_slab): `nar` is the Int32 device array, same positions as `host`
    int narrowed(const int* nar, int64_t* host, size_t n) {
        if (n == 0 || !host) return OTMB_OK;
        return staged(nar, host, n * 4, 0, 8);
    }
    bool has_stage_room(size_t bytes) const { return narrow_ok && stage_off + bytes + 64 <= f->stage_cap; }
    int values(const double* dev, double* host, size_t n, bool pageable) {
        if (n == 0 || !host) return OTMB_OK;
        if (!pageable || !narrow_ok || stage_off + n * 8 + 64 > f->stage_cap) {
            CU_TRY(c, cudaMemcpyAsync(host, dev, n * 8, cudaMemcpyDeviceToHost, st));
            return OTMB_OK;
        }
        return staged(dev, host, n * 8, 1, 1);
    }
    bool front_ready() {
        if (head >= q.size()) return false;
        const cudaError_t e = cudaEventQuery(q[head].ev);
        if (e == cudaErrorNotReady) return false;
        return true;   // success, or an error that consume() will report
    }
    int consume() {
        const Chunk& k = q[head];
        CU_TRY(c, cudaEventSynchronize(k.ev));
        f->pool->run(k.kind, k.src, k.dst, k.n);
        ++head;
        return OTMB_OK;
    }
    bool empty() const { return head >= q.size(); }
    int drain() {
        while (!empty()) OT_TRY(consume());
        return OTMB_OK;
    }
};

int fetch_impl(otmb_ctx* c, int mask, int64_t* const colptr[5], int64_t* const rowval[5], double* const nzval[5]) {
    CU_TRY(c, cudaSetDevice(c->device));
    const size_t ncp = (size_t)c->ncols + 1;
    bool fits = (c->nx * c->ny * c->nz + 2) < ((i64)1 << 31) && (c->w0 + c->N + 2) < ((i64)1 << 31);
    size_t nidx = 0, stage = 0;
    bool pageable[5] = {false, false, false, false, false};
    for (int m = 0; m < 5; ++m) {
        if (!(mask >> m & 1)) continue;
        fits = fits && (c->nnz[m] + 2) < ((i64)1 << 31);
        if (colptr && colptr[m]) nidx += ncp, stage += ncp * 4 + 64;
        if (rowval && rowval[m]) nidx += (size_t)c->nnz[m], stage += (size_t)c->nnz[m] * 4 + 64;
        if (nzval && nzval[m] && c->nnz[m] > 0 && (pageable[m] = is_pageable(nzval[m]))) stage += (size_t)c->nnz[m] * 8 + 64;
    }
    FetchState* f = nullptr;
    OT_TRY(fetch_state(c, &f));
    CopyOut co(c, f, c->stream);
    OT_TRY(co.prepare(fits, nidx, stage));
    // indices first: they are widened on the host while the values are still crossing the link
    for (int m = 0; m < 5; ++m) {
        if (!(mask >> m & 1)) continue;
        if (colptr && colptr[m]) OT_TRY(co.indices(c->colptr[m].as<i64>(), colptr[m], ncp));
        if (rowval && rowval[m]) OT_TRY(co.indices(c->rowval[m].as<i64>(), rowval[m], (size_t)c->nnz[m]));
    }
    for (int m = 0; m < 5; ++m)
        if ((mask >> m & 1) && nzval && nzval[m])
            OT_TRY(co.values(c->nzval[m].as<double>(), nzval[m], (size_t)c->nnz[m], pageable[m]));
    OT_TRY(co.drain());
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    return OTMB_OK;
}

}  // namespace

int otmb_transportmatrix_fetch(otmb_ctx* c, int which, int64_t* colptr, int64_t* rowval, double* nzval) {
    if (!c || which < 0 || which > 4) return OTMB_ERR_BADARG;
    OT_TRY(otmb_need(c, c->have_mat[which], "otmb_transportmatrix_build"));
    int64_t* cp[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    int64_t* rv[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    double* nv[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    cp[which] = colptr, rv[which] = rowval, nv[which] = nzval;
    return fetch_impl(c, 1 << which, cp, rv, nv);
}

int otmb_transportmatrix_fetch_all(otmb_ctx* c, int mask, int64_t* const colptr[5], int64_t* const rowval[5],
                                   double* const nzval[5]) {
    if (!c || mask < 0 || mask > 31) return OTMB_ERR_BADARG;
    if (mask == 0) mask = 31;
    for (int m = 0; m < 5; ++m)
        if (mask >> m & 1) OT_TRY(otmb_need(c, c->have_mat[m], "otmb_transportmatrix_build"));
    return fetch_impl(c, mask, colptr, rowval, nzval);
}

// the host half of the pipeline on its own (no device involved): widen n Int32 indices into Int64 with `threads`
// pool threads (0 = the calling thread only) — what the CPU test suite exercises
int otmb_host_widen(const int32_t* src, int64_t* dst, int64_t n, int32_t threads) {
    if (n < 0 || threads < 0 || threads > 64 || (n > 0 && (!src || !dst))) return OTMB_ERR_BADARG;
    HostPool pool(threads);
    const size_t step = (size_t)3 << 20;   // several jobs through the same pool, like the chunks of a fetch
    for (size_t lo = 0; lo < (size_t)n; lo += step) pool.run(0, src + lo, dst + lo, std::min(step, (size_t)n - lo));
    return OTMB_OK;
}


// =======================================================================================================================
// otmb_transportmatrix_stream — transportmatrix end to end in ONE call, host arrays in, host CSC arrays out, pipelined by
// level slabs so that both directions of the PCIe link are busy at once:
//     upload stream : ϕ (six arrays, and ρ when it is 3-D) level slab by level slab
//     compute stream: k_fused_v4 on the columns of slab s as soon as the levels it reads (its own + one halo level either
//                     side) have landed; the launches are CHAINED — each continues the entry totals of the one before
//                     (V4Params.start / run_out), so the device arrays are exactly those of a single launch
//     copy-out stream + host pool: as soon as a slab's completion record arrives, its colptr / rowval / nzval segments
//                     leave (indices as Int32, widened on the host) while the next slabs are uploaded and assembled.
// A plain set_facefluxes -> build -> fetch_all runs the two transfers back to back (4.7 ms + 12.5 ms on the 1-degree
// grid); here the upload hides behind the copy-out, which is the floor (it moves 2.7 times the bytes).
// nnz is data dependent, so the caller passes arrays with `capacity[m]` entries (an upper bound: N x {7,7,5,3,3}) and
// receives the five nnz; a SparseMatrixCSC wraps the first nnz entries (Julia: resize!; numpy: a view).
// =======================================================================================================================
namespace {

__global__ void k_level_prefix(const u64* __restrict__ mask, const uint32_t* __restrict__ wpre, i64 P, i64 M, i64 N, int nz,
                               i64* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > nz) return;
    const i64 X = (i64)k * P;
    out[k] = X >= M ? N : (i64)wpre[X >> 6] + __popcll(mask[X >> 6] & ((1ull << (X & 63)) - 1ull));
}

struct Uploader {
    std::thread th;
    std::atomic<int> recorded{0};   // slabs whose upload has been enqueued and its event recorded
    std::atomic<int> allowed{0};    // slabs the uploader may enqueue (pacing, see below)
    std::atomic<int> status{OTMB_OK};
    ~Uploader() {
        allowed.store(1 << 30);   // never leave the thread waiting for its turn
        if (th.joinable()) th.join();
    }
};

}  // namespace

extern "C" int otmb_transportmatrix_stream(otmb_ctx* c, const otmb_tm_params* prm, const double* const phi[6], const double* mlotst,
                                           const double* rho3d, int32_t nslabs, const int64_t capacity[5], int64_t* const colptr[5],
                                           int64_t* const rowval[5], double* const nzval[5], int64_t nnz_out[5]) {
    if (!c || !prm || !phi || !mlotst || !capacity || !colptr || !rowval || !nzval) return OTMB_ERR_BADARG;
    for (int q = 0; q < 6; ++q)
        if (!phi[q]) return otmb_fail(c, OTMB_ERR_BADARG, "null face-flux array");
    OT_TRY(otmb_need(c, c->have_indices, "otmb_makeindices"));
    OT_TRY(otmb_need(c, c->have_metrics, "otmb_gridmetrics / otmb_set_gridmetrics"));
    if (c->sharded) return otmb_fail(c, OTMB_ERR_STATE, "otmb_transportmatrix_stream is not available on a slab context");
    if (prm->index_base != 0 && prm->index_base != 1) return otmb_fail(c, OTMB_ERR_BADARG, "index_base must be 0 or 1");
    if (prm->build_mask != 0 && (prm->build_mask & 30) != 30)
        return otmb_fail(c, OTMB_ERR_BADARG, "otmb_transportmatrix_stream builds all four operators");
    if (c->topo == OTMB_TOPO_UNKNOWN) return otmb_fail(c, OTMB_ERR_UNKNOWN_GRID, otmb_status_string(OTMB_ERR_UNKNOWN_GRID));
    if (!rho3d && prm->rho != prm->rho) return otmb_fail(c, OTMB_ERR_RHO_NAN, otmb_status_string(OTMB_ERR_RHO_NAN));
    const i64 N = c->ncols;
    const i64 cap_per_col[5] = {7, 7, 5, 3, 3};
    for (int m = 0; m < 5; ++m) {
        if (!colptr[m] || !rowval[m] || !nzval[m]) return otmb_fail(c, OTMB_ERR_BADARG, "null result array");
        if (capacity[m] < 0) return OTMB_ERR_BADARG;
    }
    CU_TRY(c, cudaSetDevice(c->device));
    FetchState* f = nullptr;
    OT_TRY(fetch_state(c, &f));
    if (!f->s_up) {
        CU_TRY(c, cudaStreamCreateWithFlags(&f->s_up, cudaStreamNonBlocking));
        CU_TRY(c, cudaStreamCreateWithFlags(&f->s_dn, cudaStreamNonBlocking));
        for (auto& e : f->ev_slab) CU_TRY(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    for (int m = 0; m < 5; ++m) {
        c->have_mat[m] = false;
        c->nnz[m] = 0;
        c->preset[m] = false;
    }
    c->out_base = prm->index_base;
    c->build_serial++;
    c->build_ms_valid = false;
    if (N == 0) {
        for (int m = 0; m < 5; ++m) colptr[m][0] = prm->index_base, c->have_mat[m] = false;
        if (nnz_out)
            for (int m = 0; m < 5; ++m) nnz_out[m] = 0;
        return OTMB_OK;
    }
    // ---- slab plan: level boundaries closest to s*N/S in wet count (cached per makeindices)
    if ((i64)c->level_cum.size() != c->nz + 1) {
        CU_TRY(c, c->comm_buf.ensure((size_t)(c->nz + 1) * 8));
        k_level_prefix<<<grid_for(c->nz + 1, 128), 128, 0, c->stream>>>(c->mask.as<u64>(), c->wpre.as<uint32_t>(), c->P, c->M, c->N,
                                                                       (int)c->nz, c->comm_buf.as<i64>());
        LAUNCHED(c);
        c->level_cum.resize((size_t)c->nz + 1);
        CU_TRY(c, cudaMemcpyAsync(c->level_cum.data(), c->comm_buf.p, (size_t)(c->nz + 1) * 8, cudaMemcpyDeviceToHost, c->stream));
        CU_TRY(c, cudaStreamSynchronize(c->stream));
    }
    const std::vector<long long>& cum = c->level_cum;
    int S = nslabs > 0 ? nslabs : 8;
    S = (int)std::min<i64>(std::min<i64>(S, c->nz), otmb_ctx::DONE_RING / 2);
    // default plan: the first two slabs are small, so that the copy-out — the long pole — starts after 0.5 ms of upload
    // instead of 1.2 ms (a slab is assembled once it and the slab behind it have been uploaded)
    static const double ramp[8] = {0.04, 0.12, 0.25, 0.40, 0.55, 0.70, 0.85, 1.0};
    const bool ramped = nslabs <= 0 && S == 8;
    std::vector<int> cut(1, 0);
    for (int s = 1; s < S; ++s) {
        const long double target = ramped ? (long double)N * ramp[s - 1] : (long double)N * s / S;
        int best = cut.back() + 1;
        for (int k = best; k <= (int)c->nz - (S - s); ++k)
            if (fabsl((long double)cum[k] - target) < fabsl((long double)cum[best] - target)) best = k;
        if (best >= (int)c->nz) break;
        cut.push_back(best);
    }
    cut.push_back((int)c->nz);
    S = (int)cut.size() - 1;

    // ---- staging for the copy-out, by upper bound (the nnz are not known yet)
    bool fits = (c->M + 2) < ((i64)1 << 31) && (N * 7 + 2) < ((i64)1 << 31);
    bool pageable[5];
    size_t nidx = 0, stage = 0;
    for (int m = 0; m < 5; ++m) {
        pageable[m] = is_pageable(nzval[m]);
        const size_t bound = (size_t)std::min<i64>(capacity[m], N * cap_per_col[m]);
        nidx += (size_t)N + 1 + bound;
        stage += ((size_t)N + 1 + bound) * 4 + 64 * (size_t)(2 * S);
        if (pageable[m]) stage += bound * 8 + 64 * (size_t)S;
    }
    CopyOut co(c, f, f->s_dn);
    OT_TRY(co.prepare(fits, nidx, stage));
    // narrowed index arrays at global positions: per matrix N+1 colptr + `bound` rowval entries; slab totals [S+1][5]
    int* nar_cp[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    int* nar_rv[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    if (co.narrow_ok) {
        size_t off = 0;
        for (int m = 0; m < 5; ++m) {
            nar_cp[m] = f->narrow.as<int>() + off, off += (size_t)N + 1;
            nar_rv[m] = f->narrow.as<int>() + off, off += (size_t)std::min<i64>(capacity[m], N * cap_per_col[m]);
        }
    }
    CU_TRY(c, c->comm_buf.ensure((size_t)(S + 1) * 5 * 8));
    u64* slab_tot = c->comm_buf.as<u64>();
    CU_TRY(c, cudaMemsetAsync(slab_tot, 0, 40, c->stream));

    // ---- device buffers of the whole matrix set (as for a single launch)
    for (int q = 0; q < 6; ++q) CU_TRY(c, c->phi[q].ensure((size_t)c->M * 8));
    if (rho3d) CU_TRY(c, c->rho3d.ensure((size_t)c->M * 8));
    CU_TRY(c, c->mlotst.ensure((size_t)c->P * 8));
    c->have_rho3d = rho3d != nullptr;
    if (!c->flags_clean) {
        OT_TRY(otmb_reset_flags(c));
        c->flags_clean = true;
    }
    // uploads start once everything enqueued earlier on the compute stream is done with the old ϕ
    cudaEvent_t ev_start = f->ev_slab[2 * otmb_ctx::DONE_RING];
    CU_TRY(c, cudaEventRecord(ev_start, c->stream));
    CU_TRY(c, cudaStreamWaitEvent(f->s_up, ev_start, 0));
    CU_TRY(c, cudaStreamWaitEvent(f->s_dn, ev_start, 0));
    cudaEvent_t* ev_up = f->ev_slab;                          // [s]: slab s uploaded
    cudaEvent_t* ev_k = f->ev_slab + otmb_ctx::DONE_RING;     // [s]: slab s assembled

    Uploader up;
    up.th = std::thread([&] {
        auto fail = [&](cudaError_t e) {
            if (e != cudaSuccess) up.status.store(OTMB_ERR_CUDA);
            return e != cudaSuccess;
        };
        if (fail(cudaSetDevice(c->device))) return;
        if (fail(cudaMemcpyAsync(c->mlotst.p, mlotst, (size_t)c->P * 8, cudaMemcpyHostToDevice, f->s_up))) return;
        for (int s = 0; s < S; ++s) {
            while (up.allowed.load(std::memory_order_acquire) <= s && up.status.load() == OTMB_OK) _mm_pause();
            if (up.status.load() != OTMB_OK) return;
            const size_t a = (size_t)cut[s] * c->P, n = (size_t)(cut[s + 1] - cut[s]) * c->P;
            for (int q = 0; q < 6; ++q)
                if (fail(cudaMemcpyAsync(c->phi[q].as<double>() + a, phi[q] + a, n * 8, cudaMemcpyHostToDevice, f->s_up))) return;
            if (rho3d && fail(cudaMemcpyAsync(c->rho3d.as<double>() + a, rho3d + a, n * 8, cudaMemcpyHostToDevice, f->s_up))) return;
            if (fail(cudaEventRecord(ev_up[s], f->s_up))) return;
            up.recorded.store(s + 1, std::memory_order_release);
        }
    });

#ifdef OTMB_AB
    std::vector<std::pair<const char*, double>> trace;
    const auto tr0 = std::chrono::steady_clock::now();
    auto mark = [&](const char* what) {
        trace.emplace_back(what, std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - tr0).count());
    };
#else
    auto mark = [](const char*) {};
#endif
    // Pacing of the uploads.  With both directions busy the link moves about 72 GB/s in total (57 GB/s one way), and the
    // copy-out is 2.7 times the upload: uploading everything as fast as possible finishes the upload early and leaves the
    // copy-out alone on a half-used link for the rest of the call.  So the uploader stays only `lead` slabs ahead of the
    // copy-out that has been issued, which spreads the upload over the whole call.
    int lead = 3;
#ifdef OTMB_AB
    if (const char* e = getenv("OTMB_STREAM_LEAD")) lead = atoi(e);
#endif
    lead = std::max(lead, 2);   // slab s is launched once slab s+1 has been uploaded
    up.allowed.store(std::min(S, lead), std::memory_order_release);
    // ---- event loop: launch slabs as their inputs are enqueued, send results out as slabs complete, do the host half
    int launched = 0, issued = 0;
    u64 serial[otmb_ctx::DONE_RING];
    i64 prev[5] = {0, 0, 0, 0, 0}, tot[5] = {0, 0, 0, 0, 0};
    DevFlags agg;
    memset(&agg, 0, sizeof(agg));
    int rc = OTMB_OK;
    auto guard = [&](int st) {
        if (st != OTMB_OK && rc == OTMB_OK) rc = st;
        return st == OTMB_OK;
    };
    CU_TRY(c, cudaEventRecord(c->ev_b0, c->stream));
    while (rc == OTMB_OK && (issued < S || !co.empty())) {
        bool progress = false;
        if (up.status.load() != OTMB_OK) {
            rc = otmb_fail(c, OTMB_ERR_CUDA, "upload of the face fluxes failed");
            break;
        }
        if (launched < S && up.recorded.load(std::memory_order_acquire) >= std::min(launched + 2, S)) {
            const int s = launched;
            cudaStreamWaitEvent(c->stream, ev_up[std::min(s + 1, S - 1)], 0);
            const i64 col0 = cum[cut[s]], ncols = cum[cut[s + 1]] - col0;
            if (ncols > 0) {
                if (!guard(otmb_fused_v4_build(c, prm, 31, col0, ncols, s == 0 ? 1 : 2))) break;
            } else if (s == 0) {
                cudaMemsetAsync(c->run_nnz.p, 0, 40, c->stream);
            }
            if (ncols > 0) {
                if (!guard(otmb_v4_publish(c, slab_tot + 5 * (s + 1)))) break;
                serial[s] = c->v4_serial;
            } else {
                serial[s] = 0;   // nothing launched: nothing to wait for
                cudaMemcpyAsync(slab_tot + 5 * (s + 1), slab_tot + 5 * s, 40, cudaMemcpyDeviceToDevice, c->stream);
            }
            if (co.narrow_ok) {   // narrow this slab's indices here, on the compute stream (ranges read from the device)
                NarrowSlab ns;
                for (int m = 0; m < 5; ++m) {
                    ns.colptr[m] = c->colptr[m].as<i64>(), ns.rowval[m] = c->rowval[m].as<i64>();
                    ns.nar_colptr[m] = nar_cp[m], ns.nar_rowval[m] = nar_rv[m];
                }
                ns.col0 = col0, ns.ncp = ncols + (s == S - 1 ? 1 : 0);
                ns.before = slab_tot + 5 * s, ns.after = slab_tot + 5 * (s + 1);
                k_narrow_slab<<<dim3((unsigned)c->sm_count, 10), 256, 0, c->stream>>>(ns);
                LAUNCHED(c);
            }
            cudaEventRecord(ev_k[s], c->stream);
            ++launched;
            progress = true;
            mark("launched");
        }
        if (issued < launched) {
            const int s = issued;
            int done = 0;
            if (serial[s] != 0) {
                done = otmb_wait_v4(c, serial[s], co.empty() && launched == S);   // block only when nothing else is left to do
                if (done > 0) {
                    rc = done;
                    break;
                }
            }
            if (done == 0) {
                mark("slab done seen");
                if (serial[s] != 0) {
                    const DevFlags& h = *c->h_flags;
                    agg.err_dry_neighbour |= h.err_dry_neighbour, agg.nan_adv |= h.nan_adv, agg.nan_kh |= h.nan_kh;
                    agg.nan_kvml |= h.nan_kvml, agg.nan_kvdeep |= h.nan_kvdeep, agg.nan_rho |= h.nan_rho;
                    agg.zero_dropped |= h.zero_dropped, agg.generic_columns += h.generic_columns;
                    for (int m = 0; m < 5; ++m) tot[m] = (i64)h.nnz[m];
                }
                const i64 col0 = cum[cut[s]], ncols = cum[cut[s + 1]] - col0;
                for (int m = 0; m < 5 && rc == OTMB_OK; ++m)
                    if (tot[m] > capacity[m])
                        rc = otmb_fail(c, OTMB_ERR_BADARG, "result arrays too small: capacity[m] must hold nnz (at most N x {7,7,5,3,3})");
                if (rc != OTMB_OK) break;
                cudaStreamWaitEvent(f->s_dn, ev_k[s], 0);
                for (int m = 0; m < 5; ++m) {
                    const size_t ncp = (size_t)ncols + (s == S - 1 ? 1 : 0), nrv = (size_t)(tot[m] - prev[m]);
                    if (co.has_stage_room((ncp + nrv) * 4 + 64)) {   // narrowed on the compute stream already
                        if (!guard(co.narrowed(nar_cp[m] + col0, colptr[m] + col0, ncp))) break;
                        if (!guard(co.narrowed(nar_rv[m] + prev[m], rowval[m] + prev[m], nrv))) break;
                    } else {                                            // 8-byte copies as they are
                        co.narrow_ok = false;
                        if (!guard(co.indices(c->colptr[m].as<i64>() + col0, colptr[m] + col0, ncp))) break;
                        if (!guard(co.indices(c->rowval[m].as<i64>() + prev[m], rowval[m] + prev[m], nrv))) break;
                    }
                }
                for (int m = 0; m < 5 && rc == OTMB_OK; ++m)
                    guard(co.values(c->nzval[m].as<double>() + prev[m], nzval[m] + prev[m], (size_t)(tot[m] - prev[m]), pageable[m]));
                for (int m = 0; m < 5; ++m) prev[m] = tot[m];
                ++issued;
                up.allowed.store(std::min(S, issued + lead), std::memory_order_release);
                progress = true;
                mark("copy-out issued");
            }
        }
        if (rc == OTMB_OK && co.front_ready()) {
            mark("chunk ready");
            guard(co.consume());
            progress = true;
            mark("chunk consumed");
        }
        if (!progress) _mm_pause();
    }
    if (rc != OTMB_OK) up.status.store(rc);   // releases an uploader that is waiting for its turn
    up.allowed.store(S, std::memory_order_release);
    up.th.join();
    mark("loop end");
    if (rc == OTMB_OK && up.status.load() != OTMB_OK) rc = otmb_fail(c, OTMB_ERR_CUDA, "upload of the face fluxes failed");
    cudaEventRecord(c->ev_b1, c->stream);
    // everything enqueued must have left the streams before the caller's arrays (or an error) are handed back
    cudaStreamSynchronize(f->s_up);
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(f->s_dn);
    mark("streams idle");
#ifdef OTMB_AB
    if (getenv("OTMB_STREAM_TRACE")) {
        for (auto& t : trace) fprintf(stderr, "stream-trace %10.1f us  %s\n", t.second, t.first);
        fprintf(stderr, "stream-trace ----\n");
    }
#endif
    if (rc != OTMB_OK) {
        otmb_reset_flags(c);
        return rc;
    }
    CU_TRY(c, cudaGetLastError());
    c->have_phi = c->have_mlotst = true;
    c->build_ms_valid = true;
    *c->h_flags = agg;
    OT_TRY(otmb_check_build_flags(c, 30));
    for (int m = 0; m < 5; ++m) {
        c->nnz[m] = tot[m];
        c->have_mat[m] = true;
    }
    if (agg.zero_dropped) {
        // sparse + drops results equal to zero (src/matrixbuilding.jl:147): rare (e.g. κ = 0); compact T and send it again
        OT_TRY(otmb_drop_zeros(c, 0, prm->index_base));
        CU_TRY(c, cudaStreamSynchronize(c->stream));
        int64_t* cp[5] = {colptr[0], nullptr, nullptr, nullptr, nullptr};
        int64_t* rv[5] = {rowval[0], nullptr, nullptr, nullptr, nullptr};
        double* nv[5] = {nzval[0], nullptr, nullptr, nullptr, nullptr};
        OT_TRY(fetch_impl(c, 1, cp, rv, nv));
    }
    if (nnz_out)
        for (int m = 0; m < 5; ++m) nnz_out[m] = c->nnz[m];
    return OTMB_OK;
}

// =======================================================================================================================
// otmb_transportmatrix_dump — the resident result matrices straight to a binary file (the 12-month
// batch keeps T on the device and writes it out without ever building host SparseMatrixCSC objects).  The arrays go
// through two pinned 16 MB slots: while one is being written to the file the next chunk crosses the link.
// File layout (little endian): 8 bytes magic "OTMBCSC1"; Int64 N, index_base, nmat; then per matrix Int64 id (OTMB_MAT_*),
// Int64 nnz; then per matrix colptr (N+1 Int64), rowval (nnz Int64), nzval (nnz Float64) — the SparseMatrixCSC fields.
// =======================================================================================================================
extern "C" int otmb_transportmatrix_dump(otmb_ctx* c, int mask, const char* path) {
    if (!c || !path || mask < 0 || mask > 31) return OTMB_ERR_BADARG;
    if (mask == 0) mask = 31;
    for (int m = 0; m < 5; ++m)
        if (mask >> m & 1) OT_TRY(otmb_need(c, c->have_mat[m], "otmb_transportmatrix_build"));
    CU_TRY(c, cudaSetDevice(c->device));
    FetchState* f = nullptr;
    OT_TRY(fetch_state(c, &f));
    const size_t SLOT = CHUNK_BYTES;
    if (f->stage_cap < 2 * SLOT) {
        if (f->stage) cudaFreeHost(f->stage);
        f->stage = nullptr, f->stage_cap = 0;
        CU_TRY(c, cudaMallocHost((void**)&f->stage, 2 * SLOT));
        f->stage_cap = 2 * SLOT;
    }
    while (f->ev.size() < 2) {
        cudaEvent_t e;
        CU_TRY(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        f->ev.push_back(e);
    }
    FILE* fp = fopen(path, "wb");
    if (!fp) return otmb_fail(c, OTMB_ERR_BADARG, std::string("cannot open ") + path);
    bool io_ok = true;
    auto put = [&](const void* p, size_t n) { io_ok = io_ok && fwrite(p, 1, n, fp) == n; };
    int64_t nmat = 0;
    for (int m = 0; m < 5; ++m) nmat += mask >> m & 1;
    const int64_t head[3] = {c->ncols, c->out_base, nmat};
    put("OTMBCSC1", 8);
    put(head, sizeof(head));
    for (int m = 0; m < 5; ++m)
        if (mask >> m & 1) {
            const int64_t rec[2] = {m, c->nnz[m]};
            put(rec, sizeof(rec));
        }
    // every array as a sequence of chunks through the two slots
    struct Piece {
        const char* dev;
        size_t bytes;
    };
    std::vector<Piece> pieces;
    for (int m = 0; m < 5; ++m)
        if (mask >> m & 1) {
            pieces.push_back({(const char*)c->colptr[m].p, (size_t)(c->ncols + 1) * 8});
            pieces.push_back({(const char*)c->rowval[m].p, (size_t)c->nnz[m] * 8});
            pieces.push_back({(const char*)c->nzval[m].p, (size_t)c->nnz[m] * 8});
        }
    std::vector<Piece> chunks;
    for (const Piece& p : pieces)
        for (size_t lo = 0; lo < p.bytes; lo += SLOT) chunks.push_back({p.dev + lo, std::min(SLOT, p.bytes - lo)});
    int rc = OTMB_OK;
    for (size_t q = 0; q <= chunks.size() && rc == OTMB_OK; ++q) {
        if (q < chunks.size()) {   // chunk q into slot q & 1 (its previous content was written out two steps ago)
            if (cudaMemcpyAsync(f->stage + (q & 1) * SLOT, chunks[q].dev, chunks[q].bytes, cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
                cudaEventRecord(f->ev[q & 1], c->stream) != cudaSuccess)
                rc = otmb_fail(c, OTMB_ERR_CUDA, "copy-out for the dump failed");
        }
        if (q > 0 && rc == OTMB_OK) {   // ... while chunk q-1 goes to the file
            if (cudaEventSynchronize(f->ev[(q - 1) & 1]) != cudaSuccess)
                rc = otmb_fail(c, OTMB_ERR_CUDA, "copy-out for the dump failed");
            else
                put(f->stage + ((q - 1) & 1) * SLOT, chunks[q - 1].bytes);
        }
    }
    cudaStreamSynchronize(c->stream);
    io_ok = (fclose(fp) == 0) && io_ok;
    if (rc != OTMB_OK) return rc;
    if (!io_ok) return otmb_fail(c, OTMB_ERR_BADARG, std::string("write error on ") + path);
    return OTMB_OK;
}
