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

// host threads that widen Int32 -> Int64; run() returns when the whole range is done
struct WidenPool {
    std::vector<std::thread> th;
    std::mutex mu;
    std::condition_variable cv, cv_done;
    const int* src = nullptr;
    int64_t* dst = nullptr;
    size_t n = 0;
    int gen = 0, pending = 0;
    bool stop = false;

    explicit WidenPool(int nthreads) {
        for (int t = 0; t < nthreads; ++t) th.emplace_back([this, t] { work(t); });
    }
    ~WidenPool() {
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
    void work(int id) {
        int seen = 0;
        for (;;) {
            const int* s;
            int64_t* d;
            size_t cnt;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return stop || gen != seen; });
                if (stop) return;
                seen = gen;
                s = src, d = dst, cnt = n;
            }
            const size_t T = th.size();
            widen(s, d, cnt * id / T, cnt * (id + 1) / T);
            {
                std::lock_guard<std::mutex> lk(mu);
                if (--pending == 0) cv_done.notify_one();
            }
        }
    }
    void run(const int* s, int64_t* d, size_t cnt) {
        if (cnt < (size_t)1 << 16 || th.empty()) {
            widen(s, d, 0, cnt);
            return;
        }
        {
            std::lock_guard<std::mutex> lk(mu);
            src = s, dst = d, n = cnt;
            pending = (int)th.size();
            ++gen;
        }
        cv.notify_all();
        std::unique_lock<std::mutex> lk(mu);
        cv_done.wait(lk, [&] { return pending == 0; });
    }
};

constexpr size_t CHUNK = (size_t)4 << 20;              // Int32 entries per D2H copy / widening step (16 MB)
constexpr size_t STAGE_CAP = (size_t)512 << 20;        // Int32 entries of pinned staging at most (2 GB); beyond: direct copy
constexpr int MAX_EVENTS = 1024;

struct FetchState {
    cudaEvent_t ev[MAX_EVENTS] = {};
    int nev = 0;
    int* stage = nullptr;        // pinned, holds every narrowed index of one fetch
    size_t stage_cap = 0;
    DevBuf narrow;
    WidenPool* pool = nullptr;
    bool few_threads = false;
};

struct Segment {
    const i64* dev;
    int64_t* host;
    size_t start, len;   // position in the narrowed array
};

int fetch_state(otmb_ctx* c, FetchState** out, size_t entries) {
    if (!c->fetch_state) {
        FetchState* f = new FetchState();
        c->fetch_state = f;
        c->fetch_state_free = [](void* p) {
            FetchState* f = static_cast<FetchState*>(p);
            delete f->pool;
            for (int s = 0; s < f->nev; ++s) cudaEventDestroy(f->ev[s]);
            if (f->stage) cudaFreeHost(f->stage);
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
        f->pool = new WidenPool(nt);
    }
    FetchState* f = static_cast<FetchState*>(c->fetch_state);
    if (f->few_threads) {
        *out = nullptr;
        return OTMB_OK;
    }
    if (entries > f->stage_cap) {
        if (f->stage) cudaFreeHost(f->stage);
        f->stage = nullptr, f->stage_cap = 0;
        const size_t cap = entries + entries / 8 + 1024;
        if (cudaMallocHost((void**)&f->stage, cap * sizeof(int)) != cudaSuccess) {   // no pinnable memory left:
            cudaGetLastError();                                                      // the caller copies directly
            f->stage = nullptr;
            *out = nullptr;
            return OTMB_OK;
        }
        f->stage_cap = cap;
    }
    const int need = (int)((entries + CHUNK - 1) / CHUNK);
    while (f->nev < need) {
        CU_TRY(c, cudaEventCreateWithFlags(&f->ev[f->nev], cudaEventDisableTiming));
        ++f->nev;
    }
    *out = f;
    return OTMB_OK;
}

int fetch_impl(otmb_ctx* c, int mask, int64_t* const colptr[5], int64_t* const rowval[5], double* const nzval[5]) {
    CU_TRY(c, cudaSetDevice(c->device));
    const size_t ncp = (size_t)c->ncols + 1;
    bool fits = (c->nx * c->ny * c->nz + 2) < ((i64)1 << 31) && (c->w0 + c->N + 2) < ((i64)1 << 31) && !getenv("OTMB_FETCH_DIRECT");
    std::vector<Segment> segs;
    size_t total = 0;
    for (int m = 0; m < 5; ++m) {
        if (!(mask >> m & 1)) continue;
        fits = fits && (c->nnz[m] + 2) < ((i64)1 << 31);
        if (colptr && colptr[m]) segs.push_back({c->colptr[m].as<i64>(), colptr[m], total, ncp}), total += ncp;
        if (rowval && rowval[m] && c->nnz[m] > 0)
            segs.push_back({c->rowval[m].as<i64>(), rowval[m], total, (size_t)c->nnz[m]}), total += (size_t)c->nnz[m];
    }
    auto direct = [&]() -> int {   // the arrays cross the link as they are
        for (const Segment& s : segs) CU_TRY(c, cudaMemcpyAsync(s.host, s.dev, s.len * 8, cudaMemcpyDeviceToHost, c->stream));
        for (int m = 0; m < 5; ++m)
            if ((mask >> m & 1) && nzval && nzval[m] && c->nnz[m] > 0)
                CU_TRY(c, cudaMemcpyAsync(nzval[m], c->nzval[m].p, (size_t)c->nnz[m] * 8, cudaMemcpyDeviceToHost, c->stream));
        CU_TRY(c, cudaStreamSynchronize(c->stream));
        return OTMB_OK;
    };
    // indices beyond 32 bits, or more of them than the staging cap
    if (!fits || total > STAGE_CAP || (total + CHUNK - 1) / CHUNK > (size_t)MAX_EVENTS) return direct();
    FetchState* f = nullptr;
    OT_TRY(fetch_state(c, &f, total));
    if (!f) return direct();
    // indices: narrow on the device, cross the link first, chunk by chunk into pinned staging ...
    CU_TRY(c, f->narrow.ensure(std::max<size_t>(total, 1) * sizeof(int)));
    int* const nar = f->narrow.as<int>();
    for (const Segment& s : segs) {
        k_narrow<<<std::min<unsigned>(grid_for((i64)s.len, 256), (unsigned)c->sm_count * 8), 256, 0, c->stream>>>(s.dev, nar + s.start, (i64)s.len);
        LAUNCHED(c);
    }
    CU_TRY(c, cudaGetLastError());
    const size_t nch = (total + CHUNK - 1) / CHUNK;
    for (size_t i = 0; i < nch; ++i) {
        const size_t lo = i * CHUNK, cnt = std::min(CHUNK, total - lo);
        CU_TRY(c, cudaMemcpyAsync(f->stage + lo, nar + lo, cnt * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        CU_TRY(c, cudaEventRecord(f->ev[i], c->stream));
    }
    // ... the values follow on the same stream, straight into the caller's arrays ...
    for (int m = 0; m < 5; ++m)
        if ((mask >> m & 1) && nzval && nzval[m] && c->nnz[m] > 0)
            CU_TRY(c, cudaMemcpyAsync(nzval[m], c->nzval[m].p, (size_t)c->nnz[m] * 8, cudaMemcpyDeviceToHost, c->stream));
    // ... and while they are in flight the host widens each chunk as soon as it has landed
    size_t si = 0;
    for (size_t i = 0; i < nch; ++i) {
        CU_TRY(c, cudaEventSynchronize(f->ev[i]));
        const size_t lo = i * CHUNK, hi = std::min(total, lo + CHUNK);
        while (si < segs.size() && segs[si].start + segs[si].len <= lo) ++si;
        for (size_t s = si; s < segs.size() && segs[s].start < hi; ++s) {
            const size_t a = std::max(lo, segs[s].start), b = std::min(hi, segs[s].start + segs[s].len);
            f->pool->run(f->stage + a, segs[s].host + (a - segs[s].start), b - a);
        }
    }
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
    WidenPool pool(threads);
    const size_t step = (size_t)3 << 20;   // several jobs through the same pool, like the chunks of a fetch
    for (size_t lo = 0; lo < (size_t)n; lo += step) pool.run(src + lo, dst + lo, std::min(step, (size_t)n - lo));
    return OTMB_OK;
}

