// selftest.cu — device self-tests of arithmetic helpers whose results must equal the compiler's bit for bit.
#include "common.cuh"
#include "fdiv.cuh"

namespace {

__device__ __forceinline__ u64 mix(u64 x) {   // splitmix64
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}

// operand bit patterns by mode: 0 any 64 bits; 1 ordinary magnitudes (exponents within ±40 of 1.0), the assembly's
// range; 2 exponents drawn from the edges (zero / subnormal / just above the chain's lower bounds / huge / Inf / NaN)
__device__ double operand(u64 r, int mode) {
    const u64 man = r & 0x000fffffffffffffull, sign = r & 0x8000000000000000ull;
    if (mode == 0) return __longlong_as_double((long long)r);
    if (mode == 1) return __longlong_as_double((long long)(sign | ((u64)(1023 - 40 + (r >> 52) % 81) << 52) | man));
    const unsigned edges[16] = {0, 0, 1, 2, 0x035, 0x036, 0x037, 0x038, 0x3ff, 0x400, 0x7f7, 0x7f8, 0x7f9, 0x7fd, 0x7fe, 0x7ff};
    const u64 ex = edges[(r >> 52) & 15];
    const u64 m2 = (r >> 56 & 3) == 0 ? 0 : (r >> 56 & 3) == 1 ? 0x000fffffffffffffull : man;
    return __longlong_as_double((long long)(sign | (ex << 52) | m2));
}

__global__ void k_selftest_div(i64 n, u64 seed, unsigned long long* bad, double* first_bad) {
    for (i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const u64 r0 = mix(seed + 4 * (u64)i);
        const int mode = (int)(r0 % 3), mode_b = (r0 >> 8) % 5 == 0 ? (int)((r0 >> 16) % 3) : mode;
        const double a1 = operand(mix(r0 + 1), mode), b1 = operand(mix(r0 + 2), mode_b);
        const double a2 = operand(mix(r0 + 3), mode), b2 = operand(mix(r0 + 4), mode_b);
        double q1, q2;
        otmb_fdiv::div2(a1, b1, a2, b2, q1, q2);
        // the reference quotients, kept away from div2's own fall-back code by the volatile copies
        volatile double va1 = a1, vb1 = b1, va2 = a2, vb2 = b2;
        const double w1 = va1 / vb1, w2 = va2 / vb2;
        const bool d1 = __double_as_longlong(q1) != __double_as_longlong(w1), d2 = __double_as_longlong(q2) != __double_as_longlong(w2);
        if (d1 || d2) {
            if (atomicAdd(bad, 1ull) == 0) {
                first_bad[0] = d1 ? a1 : a2;
                first_bad[1] = d1 ? b1 : b2;
                first_bad[2] = d1 ? q1 : q2;
                first_bad[3] = d1 ? w1 : w2;
            }
        }
    }
}

}  // namespace

extern "C" int otmb_selftest_division(otmb_ctx* c, int64_t n, uint64_t seed, int64_t* mismatches, double first_bad[4]) {
    if (!c || n < 0 || !mismatches) return OTMB_ERR_BADARG;
    CU_TRY(c, cudaSetDevice(c->device));
    CU_TRY(c, c->held_diff.ensure(8));
    DevBuf fb;
    CU_TRY(c, fb.ensure(32));
    CU_TRY(c, cudaMemsetAsync(c->held_diff.p, 0, 8, c->stream));
    CU_TRY(c, cudaMemsetAsync(fb.p, 0, 32, c->stream));
    if (n > 0) k_selftest_div<<<c->sm_count * 8, 256, 0, c->stream>>>(n, seed, c->held_diff.as<unsigned long long>(), fb.as<double>());
    CU_TRY(c, cudaGetLastError());
    unsigned long long h = 0;
    double hb[4] = {0, 0, 0, 0};
    CU_TRY(c, cudaMemcpyAsync(&h, c->held_diff.p, 8, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaMemcpyAsync(hb, fb.p, 32, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    fb.release();
    *mismatches = (int64_t)h;
    if (first_bad)
        for (int q = 0; q < 4; ++q) first_bad[q] = hb[q];
    return OTMB_OK;
}
