// scan.cu — hand-written device-wide exclusive scan (block scan -> scan of block sums -> add),
// used by makeindices (word popcounts -> wet-rank prefix) and by the COO->CSC / sparse-add
// passes (per-column counts -> colptr).  The fused assembly kernel carries its own
// single-pass decoupled look-back scan (fused.cu).
#include "common.cuh"

namespace {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <typename T>
__device__ __forceinline__ T warp_incl_scan(T v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += o;
    }
    return v;
}

template <typename Tin, typename Tout>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_block(const Tin* __restrict__ in, Tout* __restrict__ out, i64 n,
                                                             Tout* __restrict__ blocksums) {
    __shared__ Tout warp_sums[SCAN_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const i64 base = (i64)blockIdx.x * SCAN_TILE + (i64)tid * SCAN_ITEMS;
    Tout v[SCAN_ITEMS];
    Tout tsum = 0;
#pragma unroll
    for (int it = 0; it < SCAN_ITEMS; ++it) {
        i64 idx = base + it;
        v[it] = idx < n ? (Tout)in[idx] : (Tout)0;
        tsum += v[it];
    }
    Tout incl = warp_incl_scan<Tout>(tsum, lane);
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        Tout w = lane < SCAN_THREADS / 32 ? warp_sums[lane] : (Tout)0;
        Tout wi = warp_incl_scan<Tout>(w, lane);
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = wi - w;  // exclusive
        if (lane == SCAN_THREADS / 32 - 1) blocksums[blockIdx.x] = wi;
    }
    __syncthreads();
    Tout run = warp_sums[wid] + (incl - tsum);
#pragma unroll
    for (int it = 0; it < SCAN_ITEMS; ++it) {
        i64 idx = base + it;
        if (idx < n) out[idx] = run;
        run += v[it];
    }
}

// single block: exclusive scan of the block sums in place, total to *total
template <typename Tout>
__global__ void __launch_bounds__(1024) k_scan_sums(Tout* __restrict__ sums, i64 nb, u64* __restrict__ total) {
    __shared__ Tout warp_sums[32];
    __shared__ Tout carry_s;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (i64 base = 0; base < nb; base += 1024) {
        i64 idx = base + tid;
        Tout v = idx < nb ? sums[idx] : (Tout)0;
        Tout incl = warp_incl_scan<Tout>(v, lane);
        if (lane == 31) warp_sums[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            Tout w = warp_sums[lane];
            Tout wi = warp_incl_scan<Tout>(w, lane);
            warp_sums[lane] = wi - w;
        }
        __syncthreads();
        Tout carry = carry_s;
        Tout excl = carry + warp_sums[wid] + (incl - v);
        if (idx < nb) sums[idx] = excl;
        __syncthreads();
        if (tid == 1023) carry_s = excl + v;
        __syncthreads();
    }
    if (tid == 0 && total) *total = (u64)carry_s;
}

template <typename Tout>
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_add(Tout* __restrict__ out, i64 n,
                                                           const Tout* __restrict__ blocksums) {
    const Tout add = blocksums[blockIdx.x];
    const i64 base = (i64)blockIdx.x * SCAN_TILE + (i64)threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int it = 0; it < SCAN_ITEMS; ++it) {
        i64 idx = base + it;
        if (idx < n) out[idx] += add;
    }
}

template <typename Tin, typename Tout>
int scan_impl(otmb_ctx* ctx, const Tin* in, Tout* out, i64 n, u64* total_dev) {
    if (n <= 0) {
        if (total_dev) CU_TRY(ctx, cudaMemsetAsync(total_dev, 0, sizeof(u64), ctx->stream));
        return OTMB_OK;
    }
    const i64 nb = (n + SCAN_TILE - 1) / SCAN_TILE;
    CU_TRY(ctx, ctx->scan_tmp.ensure((size_t)nb * sizeof(Tout)));
    Tout* sums = ctx->scan_tmp.as<Tout>();
    k_scan_block<Tin, Tout><<<(unsigned)nb, SCAN_THREADS, 0, ctx->stream>>>(in, out, n, sums);
    LAUNCHED(ctx);
    k_scan_sums<Tout><<<1, 1024, 0, ctx->stream>>>(sums, nb, total_dev);
    LAUNCHED(ctx);
    if (nb > 1) {
        k_scan_add<Tout><<<(unsigned)nb, SCAN_THREADS, 0, ctx->stream>>>(out, n, sums);
        LAUNCHED(ctx);
    }
    CU_TRY(ctx, cudaGetLastError());
    return OTMB_OK;
}

}  // namespace

int otmb_scan_u32(otmb_ctx* ctx, const uint32_t* in, uint32_t* out, i64 n, u64* total_dev) {
    return scan_impl<uint32_t, uint32_t>(ctx, in, out, n, total_dev);
}
int otmb_scan_i64(otmb_ctx* ctx, const i64* in, i64* out, i64 n, u64* total_dev) {
    return scan_impl<i64, i64>(ctx, in, out, n, total_dev);
}
int otmb_scan_u32_to_i64(otmb_ctx* ctx, const uint32_t* in, i64* out, i64 n, u64* total_dev) {
    return scan_impl<uint32_t, i64>(ctx, in, out, n, total_dev);
}
