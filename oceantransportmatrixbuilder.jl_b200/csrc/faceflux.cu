// faceflux.cu — K4: nofluxboundaries! + facefluxes in one pass.
// Replaces /root/reference/src/velocities.jl:154-179 (masking), :199-200 (asserts), :203-224
// (NaN/fill -> 0, west/south shifts) and :234-243 (bottom-up continuity scan).
//
// One thread per (i,j) column, walking k from the sea floor up: the only sequential
// dependency on the whole path is phi_top[k] = ((((phi_bottom + west) + south) - east) - north),
// evaluated left to right exactly as the reference's broadcast does (adds only => bit-exact).
// Threads are consecutive in i, so every load/store of a level is a coalesced row segment;
// the west/south operands are the east/north values of the i-1 / j-1 columns re-read through
// L1 (same or adjacent cache lines).  The wet tests use the packed bit mask.
#include <cstddef>
#include <cstdlib>

#include "common.cuh"

namespace {

__device__ __forceinline__ double clean(double x, double fill) { return (isnan(x) || x == fill) ? 0.0 : x; }

template <int UNROLL>
__global__ void __launch_bounds__(128) k_faceflux(const double* __restrict__ umo, const double* __restrict__ vmo,
                                                  const u64* __restrict__ mask, GridDims g, double fill,
                                                  double* __restrict__ east, double* __restrict__ west,
                                                  double* __restrict__ north, double* __restrict__ south,
                                                  double* __restrict__ top, double* __restrict__ bottom,
                                                  DevFlags* __restrict__ flags, int row0, int row1, int p_begin, int p_end,
                                                  const double* __restrict__ carry_in, double* __restrict__ carry_out,
                                                  PeerLink link) {
    // (all 3-D pointers are window-biased and indexed by the global linear cell index; the kernel covers the columns
    // [p_begin, p_end) of the plane — one chunk of the pipelined carry chain — and, per column, the owned levels)
    const int p = p_begin + blockIdx.x * blockDim.x + threadIdx.x;
    bool valid_u = false, valid_v = false;
    // Peer-memory form of the carry chain (comm.cu): carry_in lives in THIS GPU's memory and is written by the rank below
    // over NVLink, block by block; a block waits for the flag of its own 128 columns, so the ranks' kernels overlap
    // column block by column block instead of exchanging whole planes.  The spin is bounded: a peer that never
    // arrives becomes an error code (flags->lookback_timeout), not a hang.
    double carry0 = 0.0, carry_last = 0.0;
    if (link.flag_in) {
        if (threadIdx.x == 0) {
            unsigned v, spins = 0;
            do {
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(link.flag_in + blockIdx.x) : "memory");
            } while (v != link.epoch && ++spins < (1u << 25));
            if (v != link.epoch) atomicOr(&flags->lookback_timeout, 1);
        }
        __syncthreads();
        if (p < p_end) carry0 = __ldcv(carry_in + p);   // (the peer wrote it behind this SM's back: read around L1)
        __syncthreads();                                // every value of the block is in a register: the inbox may be reused
        if (threadIdx.x == 0)
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(link.ack_out + blockIdx.x), "r"(link.epoch) : "memory");
    }
    if (p < p_end) {
        const int i = p % g.nx, j = p / g.nx;
        // owned levels of this column: the grid rows R = j + ny*k inside [row0, row1)
        const int k_begin = (row0 - j + g.ny - 1) / g.ny, k_end = (row1 - j + g.ny - 1) / g.ny;
        const int pE = i < g.nx - 1 ? p + 1 : p - (g.nx - 1);
        const int pW = i > 0 ? p - 1 : p + (g.nx - 1);
        const int pS = j > 0 ? p - g.nx : -1;
        // north neighbour of this column, and of the column to the south (always regular: j-1 < ny-1)
        const int pN = j < g.ny - 1 ? p + g.nx : (g.topo == OTMB_TOPO_TRIPOLAR ? (g.nx - 1 - i) + g.nx * (g.ny - 1) : -1);
        // phi_top of the level below: 0 under the sea floor, or the plane handed up by the slab below
        double carry = carry_in ? (link.flag_in ? carry0 : carry_in[p]) : 0.0;
        const bool owns = k_begin < k_end;
        if (owns && carry_in && k_end < g.nz) top[(size_t)k_end * g.P + p] = carry;   // halo cell below: its top flux
        for (int k0 = k_end - 1; k0 >= k_begin; k0 -= UNROLL) {
            double e[UNROLL], w[UNROLL], n[UNROLL], s[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int k = k0 - u;
                e[u] = w[u] = n[u] = s[u] = 0.0;
                if (k < k_begin) continue;
                const int off = k * g.P;
                const int L = off + p;
                const bool wc = wet_at(mask, L);
                const bool wE = wet_at(mask, off + pE);
                const bool wW = wet_at(mask, off + pW);
                const bool wN = pN >= 0 && wet_at(mask, off + pN);
                const bool wS = pS >= 0 && wet_at(mask, off + pS);
                // nofluxboundaries!: zero at dry cells and towards dry/absent east / north neighbours
                const double ue = (wc && wE) ? __ldg(umo + L) : 0.0;
                const double vn = (wc && wN) ? __ldg(vmo + L) : 0.0;
                const double uw = (wW && wc) ? __ldg(umo + off + pW) : 0.0;
                const double vs = (wS && wc) ? __ldg(vmo + off + pS) : 0.0;   // north nbr of (i,j-1) is (i,j)
                valid_u |= !(isnan(ue) || ue == fill);
                valid_v |= !(isnan(vn) || vn == fill);
                e[u] = clean(ue, fill);
                n[u] = clean(vn, fill);
                w[u] = clean(uw, fill);
                s[u] = clean(vs, fill);
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const int k = k0 - u;
                if (k < k_begin) continue;
                const size_t L = (size_t)k * g.P + p;
                const double b = carry;
                const double t = (((b + w[u]) + s[u]) - e[u]) - n[u];
                east[L] = e[u];
                west[L] = w[u];
                north[L] = n[u];
                south[L] = s[u];
                bottom[L] = b;
                top[L] = t;
                carry = t;
            }
        }
        carry_last = carry;
        if (carry_out && !link.flag_out) carry_out[p] = carry;   // phi_top of the column's first owned level (a column that owns nothing passes it on)
        if (owns && k_begin > 0) bottom[(size_t)(k_begin - 1) * g.P + p] = carry;   // halo cell above: its bottom flux
    }
    const unsigned bu = __ballot_sync(0xffffffffu, valid_u), bv = __ballot_sync(0xffffffffu, valid_v);
    if ((threadIdx.x & 31) == 0) {
        if (bu) atomicOr(&flags->any_valid_u, 1);
        if (bv) atomicOr(&flags->any_valid_v, 1);
    }
    if (link.flag_out) {   // carry_out is the inbox of the rank above (mapped peer memory)
        if (threadIdx.x == 0) {   // ... which must have taken the previous epoch's values of this block out of it
            unsigned v, spins = 0;
            do {
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(link.ack_in + blockIdx.x) : "memory");
            } while ((int)(v - (link.epoch - 1u)) < 0 && ++spins < (1u << 25));
            if ((int)(v - (link.epoch - 1u)) < 0) atomicOr(&flags->lookback_timeout, 1);
        }
        __syncthreads();
        if (p < p_end) carry_out[p] = carry_last;   // a plain store over NVLink
        __threadfence_system();                    // visible before the flag
        __syncthreads();
        if (threadIdx.x == 0)
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(link.flag_out + blockIdx.x), "r"(link.epoch) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// The same computation with the inputs staged by the bulk-copy engine (TMA, `cp.async.bulk`).  There are only nx*ny threads (108 000 at 1 degree, a third of the GPU's thread slots) and
// the register-only way to keep more loads in flight — a deeper unroll — measured slower (profiles/README.md).  Here one
// elected thread per block keeps FT_STAGES levels of the block's input rows in flight into shared memory — per level three
// contiguous segments: umo[p0-2 .. p0+128) (own east flux + the west neighbour's), vmo[p0 .. p0+128), vmo[p0-nx ..
// p0-nx+128) (the south neighbour's north flux) — each completing on the slot's mbarrier; the 128 threads take their
// operands from the slot, hand it back for the level FT_STAGES further up, and compute / store as before.  The queue to
// DRAM stays full while the threads compute, independent of the thread count.  Needs nx and nx*ny even (16-byte source
// alignment of every segment); other grids take the plain kernel.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int FT_TI = 128, FT_STAGES = 8;
struct FtSlot {
    double u[FT_TI + 2];
    double vn[FT_TI];
    double vs[FT_TI];
};
static_assert(sizeof(FtSlot) % 16 == 0 && offsetof(FtSlot, vn) % 16 == 0 && offsetof(FtSlot, vs) % 16 == 0, "16-byte aligned segments");

__device__ __forceinline__ unsigned smem_addr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)), "l"(src),
                 "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void bar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(
            smem_addr(bar)),
        "r"(parity)
        : "memory");
}

__global__ void __launch_bounds__(FT_TI) k_faceflux_tile(const double* __restrict__ umo, const double* __restrict__ vmo,
                                                         const u64* __restrict__ mask, GridDims g, double fill,
                                                         double* __restrict__ east, double* __restrict__ west,
                                                         double* __restrict__ north, double* __restrict__ south,
                                                         double* __restrict__ top, double* __restrict__ bottom,
                                                         DevFlags* __restrict__ flags, int row0, int row1, int p_begin, int p_end,
                                                         const double* __restrict__ carry_in, double* __restrict__ carry_out,
                                                         PeerLink link) {
    // (same arguments and semantics as k_faceflux: window-biased 3-D pointers, columns [p_begin, p_end), owned rows
    // [row0, row1), carry planes / peer link of a slab chain; p_begin is a multiple of FT_TI)
    __shared__ __align__(128) FtSlot slot[FT_STAGES];
    __shared__ __align__(8) unsigned long long full[FT_STAGES];
    const int tid = threadIdx.x, p0 = p_begin + blockIdx.x * FT_TI, p = p0 + tid;
    const int pend = min(p0 + FT_TI, p_end);
    // the three segments of a level, in columns of the plane (all even: p0, nx and P are)
    const int ua = max(p0 - 2, 0), un = pend - ua;
    const int vn_n = pend - p0;
    const int vs_lo = max(p0 - g.nx, 0), vs_n = max(pend - g.nx - vs_lo, 0);
    const unsigned bytes = (unsigned)(un + vn_n + vs_n) * 8u;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < FT_STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(&full[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // levels the block visits: from the deepest owned level of its first column (smallest j) up to the first owned level
    // of its last column — the owned ranges of the columns of a block differ by at most the rows it spans
    const int kmax = (row1 - p0 / g.nx + g.ny - 1) / g.ny, kmin = (row0 - (pend - 1) / g.nx + g.ny - 1) / g.ny;
    auto issue = [&](const int k) {   // one thread: arm the slot's barrier with the byte count, then the three copies
        const int s = (kmax - 1 - k) % FT_STAGES;
        FtSlot& S = slot[s];
        unsigned long long* bar = &full[s];
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
        const size_t off = (size_t)k * g.P;
        bulk_load(S.u, umo + off + ua, (unsigned)un * 8u, bar);
        bulk_load(S.vn, vmo + off + p0, (unsigned)vn_n * 8u, bar);
        if (vs_n > 0) bulk_load(S.vs, vmo + off + vs_lo, (unsigned)vs_n * 8u, bar);
    };
    if (tid == 0)
        for (int q = 0; q < FT_STAGES; ++q)
            if (kmax - 1 - q >= kmin) issue(kmax - 1 - q);
    const bool in = p < pend;
    const int pc = in ? p : pend - 1;   // threads past the end follow the last column and store nothing
    const int i = pc % g.nx, j = pc / g.nx;
    const int k_begin = (row0 - j + g.ny - 1) / g.ny, k_end = (row1 - j + g.ny - 1) / g.ny;   // this column's owned levels
    const bool owns = in && k_begin < k_end;
    const int pE = i < g.nx - 1 ? pc + 1 : pc - (g.nx - 1);
    const int pW = i > 0 ? pc - 1 : pc + (g.nx - 1);
    const int pS = j > 0 ? pc - g.nx : -1;
    const int pN = j < g.ny - 1 ? pc + g.nx : (g.topo == OTMB_TOPO_TRIPOLAR ? (g.nx - 1 - i) + g.nx * (g.ny - 1) : -1);
    bool valid_u = false, valid_v = false;
    // phi_top of the level below: 0 under the sea floor, or the plane handed up by the slab below (see k_faceflux)
    double carry = 0.0;
    if (link.flag_in) {
        if (tid == 0) {
            unsigned v, spins = 0;
            do {
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(link.flag_in + blockIdx.x) : "memory");
            } while (v != link.epoch && ++spins < (1u << 25));
            if (v != link.epoch) atomicOr(&flags->lookback_timeout, 1);
        }
        __syncthreads();
        if (in) carry = __ldcv(carry_in + p);
        __syncthreads();
        if (tid == 0)
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(link.ack_out + blockIdx.x), "r"(link.epoch) : "memory");
    } else if (carry_in && in) {
        carry = carry_in[p];
    }
    if (owns && carry_in && k_end < g.nz) top[(size_t)k_end * g.P + p] = carry;   // halo cell below: its top flux
    for (int k = kmax - 1; k >= kmin; --k) {
        const int q = kmax - 1 - k, s = q % FT_STAGES;
        bar_wait(&full[s], (unsigned)((q / FT_STAGES) & 1));
        const FtSlot& S = slot[s];
        const int off = k * g.P;
        // operands out of the slot (the west neighbour of i = 0 sits at the other end of the row: a plain load)
        const bool act = owns && k >= k_begin && k < k_end;
        const double ue_raw = S.u[pc - ua];
        const double uw_raw = i > 0 ? S.u[pc - 1 - ua] : (act ? __ldg(umo + off + pW) : 0.0);
        const double vn_raw = S.vn[pc - p0];
        const double vs_raw = pS >= 0 ? S.vs[pS - vs_lo] : 0.0;
        __syncthreads();                                   // every thread holds its operands: the slot is free
        if (tid == 0 && k - FT_STAGES >= kmin) issue(k - FT_STAGES);
        if (!act) continue;
        const int L = off + pc;
        const bool wc = wet_at(mask, L);
        const bool wE = wet_at(mask, off + pE);
        const bool wW = wet_at(mask, off + pW);
        const bool wN = pN >= 0 && wet_at(mask, off + pN);
        const bool wS = pS >= 0 && wet_at(mask, off + pS);
        // nofluxboundaries!: zero at dry cells and towards dry/absent east / north neighbours
        const double ue = (wc && wE) ? ue_raw : 0.0;
        const double vn = (wc && wN) ? vn_raw : 0.0;
        const double uw = (wW && wc) ? uw_raw : 0.0;
        const double vs = (wS && wc) ? vs_raw : 0.0;   // north nbr of (i,j-1) is (i,j)
        valid_u |= !(isnan(ue) || ue == fill);
        valid_v |= !(isnan(vn) || vn == fill);
        const double e = clean(ue, fill), n = clean(vn, fill), w = clean(uw, fill), so = clean(vs, fill);
        const double b = carry;
        const double t = (((b + w) + so) - e) - n;
        __stcs(east + L, e);
        __stcs(west + L, w);
        __stcs(north + L, n);
        __stcs(south + L, so);
        __stcs(bottom + L, b);
        __stcs(top + L, t);
        carry = t;
    }
    if (in && carry_out && !link.flag_out) carry_out[p] = carry;   // phi_top of the column's first owned level (passed on if it owns nothing)
    if (owns && k_begin > 0) bottom[(size_t)(k_begin - 1) * g.P + p] = carry;   // halo cell above: its bottom flux
    const unsigned bu = __ballot_sync(0xffffffffu, valid_u), bv = __ballot_sync(0xffffffffu, valid_v);
    if ((tid & 31) == 0) {
        if (bu) atomicOr(&flags->any_valid_u, 1);
        if (bv) atomicOr(&flags->any_valid_v, 1);
    }
    if (link.flag_out) {   // carry_out is the inbox of the rank above (mapped peer memory), see k_faceflux
        if (tid == 0) {
            unsigned v, spins = 0;
            do {
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(link.ack_in + blockIdx.x) : "memory");
            } while ((int)(v - (link.epoch - 1u)) < 0 && ++spins < (1u << 25));
            if ((int)(v - (link.epoch - 1u)) < 0) atomicOr(&flags->lookback_timeout, 1);
        }
        __syncthreads();
        if (in) carry_out[p] = carry;
        __threadfence_system();
        __syncthreads();
        if (tid == 0)
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(link.flag_out + blockIdx.x), "r"(link.epoch) : "memory");
    }
}

// A slab cut inside a level: the assembly of the first / last owned grid row reads the north-face flux of the row
// before it and the south-face flux of the row after it (its south / north neighbours, same level, owned by the
// adjacent rank).  Both are pure functions of the inputs (vmo + mask, src/velocities.jl:169-174, :213-221), which the
// window holds, so they are recomputed here instead of exchanged.
__global__ void __launch_bounds__(128) k_faceflux_halo_rows(const double* __restrict__ vmo, const u64* __restrict__ mask,
                                                            GridDims g, double fill, double* __restrict__ north,
                                                            double* __restrict__ south, int row_before, int row_after) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.nx) return;
    if (row_before >= 0) {   // north[L] of the cells of that row; its north neighbour (j+1 < ny) is owned
        const size_t L = (size_t)row_before * g.nx + i;
        const double vn = (wet_at(mask, (int)L) && wet_at(mask, (int)L + g.nx)) ? __ldg(vmo + L) : 0.0;
        north[L] = clean(vn, fill);
    }
    if (row_after >= 0) {    // south[L] = north flux of the cell to the south (owned), masked by both cells
        const size_t L = (size_t)row_after * g.nx + i;
        const double vs = (wet_at(mask, (int)L - g.nx) && wet_at(mask, (int)L)) ? __ldg(vmo + L - g.nx) : 0.0;
        south[L] = clean(vs, fill);
    }
}

}  // namespace

// ---- internal pieces shared with the NCCL-chained driver (comm.cu) ---------------------------------------------
// window-sized ϕ buffers, flags reset, halo rows of a mid-level cut
int otmb_faceflux_begin(otmb_ctx* c, double fill) {
    for (int q = 0; q < 6; ++q) CU_TRY(c, c->phi[q].ensure(c->win_cells() * 8));
    OT_TRY(otmb_reset_flags(c));
    const i64 rows = c->ny * c->nz, row0 = c->L_own0 / c->nx, row1 = c->L_own1 / c->nx;
    const int before = (row0 > 0 && row0 % c->ny != 0) ? (int)(row0 - 1) : -1;
    const int after = (row1 < rows && row1 % c->ny != 0) ? (int)row1 : -1;
    if (before >= 0 || after >= 0) {
        GridDims g{(int)c->nx, (int)c->ny, (int)c->nz, c->topo, (int)c->P, (int)c->M};
        k_faceflux_halo_rows<<<grid_for(c->nx, 128), 128, 0, c->stream>>>(c->win<double>(c->stage_b), c->mask_win(), g, fill,
                                                                          c->win<double>(c->phi[OTMB_FACE_NORTH]),
                                                                          c->win<double>(c->phi[OTMB_FACE_SOUTH]), before, after);
        LAUNCHED(c);
        CU_TRY(c, cudaGetLastError());
    }
    return OTMB_OK;
}
// the columns [p_begin, p_end) of the plane; d_in / d_out are device planes of nx*ny doubles (or null)
int otmb_faceflux_columns(otmb_ctx* c, double fill, i64 p_begin, i64 p_end, const double* d_in, double* d_out, PeerLink link) {
    if (p_end <= p_begin) return OTMB_OK;
    GridDims g{(int)c->nx, (int)c->ny, (int)c->nz, c->topo, (int)c->P, (int)c->M};
    // segments 16-byte aligned (nx and nx*ny even): inputs staged by the bulk-copy engine (k_faceflux_tile)
    bool tile = c->nx % 2 == 0 && c->P % 2 == 0 && c->nx >= 2 && p_begin % FT_TI == 0;
#ifdef OTMB_AB
    if (getenv("OTMB_FACEFLUX_PLAIN")) tile = false;
#endif
    if (tile) {
        k_faceflux_tile<<<grid_for(p_end - p_begin, FT_TI), FT_TI, 0, c->stream>>>(
            c->win<double>(c->stage_a), c->win<double>(c->stage_b), c->mask_win(), g, fill, c->win<double>(c->phi[OTMB_FACE_EAST]),
            c->win<double>(c->phi[OTMB_FACE_WEST]), c->win<double>(c->phi[OTMB_FACE_NORTH]), c->win<double>(c->phi[OTMB_FACE_SOUTH]),
            c->win<double>(c->phi[OTMB_FACE_TOP]), c->win<double>(c->phi[OTMB_FACE_BOTTOM]), c->flags.as<DevFlags>(),
            (int)(c->L_own0 / c->nx), (int)(c->L_own1 / c->nx), (int)p_begin, (int)p_end, d_in, d_out, link);
        LAUNCHED(c);
        CU_TRY(c, cudaGetLastError());
        return OTMB_OK;
    }
    // Five levels of loads in flight per column.  Ten were measured slower even on the 1-degree grid, where there are
    // only 108 000 columns (74 us vs 121 us, profiles/README.md): the batch then no longer fits the registers.
    int unroll = 5;
#ifdef OTMB_AB
    if (const char* e = getenv("OTMB_FACEFLUX_UNROLL")) unroll = atoi(e);
#endif
    auto kern = unroll == 10 ? k_faceflux<10> : k_faceflux<5>;
    kern<<<grid_for(p_end - p_begin, 128), 128, 0, c->stream>>>(
        c->win<double>(c->stage_a), c->win<double>(c->stage_b), c->mask_win(), g, fill,
        c->win<double>(c->phi[OTMB_FACE_EAST]), c->win<double>(c->phi[OTMB_FACE_WEST]), c->win<double>(c->phi[OTMB_FACE_NORTH]),
        c->win<double>(c->phi[OTMB_FACE_SOUTH]), c->win<double>(c->phi[OTMB_FACE_TOP]), c->win<double>(c->phi[OTMB_FACE_BOTTOM]),
        c->flags.as<DevFlags>(), (int)(c->L_own0 / c->nx), (int)(c->L_own1 / c->nx), (int)p_begin, (int)p_end, d_in, d_out, link);
    LAUNCHED(c);
    CU_TRY(c, cudaGetLastError());
    return OTMB_OK;
}
// copy the owned part of the six results into the caller's full-size arrays (any may be null)
int otmb_faceflux_copy_out(otmb_ctx* c, double* const outs[6]) {
    const size_t n = (size_t)(c->L_own1 - c->L_own0);
    for (int q = 0; q < 6; ++q)
        if (outs[q])
            CU_TRY(c, cudaMemcpyAsync(outs[q] + c->L_own0, c->win<double>(c->phi[q]) + c->L_own0, n * 8, cudaMemcpyDeviceToHost, c->stream));
    return OTMB_OK;
}
int otmb_upload_uv(otmb_ctx* c, const double* umo, const double* vmo, double fill) {
    OT_TRY(otmb_upload3d(c, c->stage_a, umo));
    OT_TRY(otmb_upload3d(c, c->stage_b, vmo));
    c->have_uv = true;
    c->uv_fill = fill;
    return OTMB_OK;
}

static int facefluxes_impl(otmb_ctx* c, const double* umo, const double* vmo, double fill, const double* carry_in,
                           double* carry_out, int carry_on_device, int32_t* valid_uv, double* const outs[6]) {
    if (!c || !umo || !vmo) return OTMB_ERR_BADARG;
    OT_TRY(otmb_need(c, c->have_indices, "otmb_makeindices"));
    if (c->topo == OTMB_TOPO_UNKNOWN) return otmb_fail(c, OTMB_ERR_UNKNOWN_GRID, otmb_status_string(OTMB_ERR_UNKNOWN_GRID));
    CU_TRY(c, cudaSetDevice(c->device));
    const size_t P8 = (size_t)c->P * 8;
    OT_TRY(otmb_upload_uv(c, umo, vmo, fill));
    // carry planes: device pointers are used in place (e.g. buffers an NCCL send/recv works on)
    const double* d_in = nullptr;
    double* d_out = nullptr;
    if (carry_in) {
        if (carry_on_device)
            d_in = carry_in;
        else {
            CU_TRY(c, c->carry[0].ensure(P8));
            CU_TRY(c, cudaMemcpyAsync(c->carry[0].p, carry_in, P8, cudaMemcpyHostToDevice, c->stream));
            d_in = c->carry[0].as<double>();
        }
    }
    if (carry_out) {
        if (carry_on_device)
            d_out = carry_out;
        else {
            CU_TRY(c, c->carry[1].ensure(P8));
            d_out = c->carry[1].as<double>();
        }
    }
    OT_TRY(otmb_faceflux_begin(c, fill));
    OT_TRY(otmb_faceflux_columns(c, fill, 0, c->P, d_in, d_out));
    OT_TRY(otmb_fetch_flags(c));
    if (valid_uv) {
        // a slab only sees its own levels: the caller combines the flags of all ranks (src/velocities.jl:199-200)
        valid_uv[0] = c->h_flags->any_valid_u;
        valid_uv[1] = c->h_flags->any_valid_v;
    } else if (!c->h_flags->any_valid_u || !c->h_flags->any_valid_v) {
        c->have_phi = false;
        return otmb_fail(c, OTMB_ERR_ALL_FILL, otmb_status_string(OTMB_ERR_ALL_FILL));
    }
    if (carry_out && !carry_on_device) CU_TRY(c, cudaMemcpyAsync(carry_out, d_out, P8, cudaMemcpyDeviceToHost, c->stream));
    OT_TRY(otmb_faceflux_copy_out(c, outs));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    c->have_phi = true;
    return OTMB_OK;
}

extern "C" int otmb_facefluxes(otmb_ctx* c, const double* umo, const double* vmo, double fill, double* east,
                               double* west, double* north, double* south, double* top, double* bottom) {
    if (c && c->sharded) return otmb_fail(c, OTMB_ERR_STATE, "slab context: use otmb_facefluxes_slab");
    double* const outs[6] = {east, west, north, south, top, bottom};
    return facefluxes_impl(c, umo, vmo, fill, nullptr, nullptr, 0, nullptr, outs);
}

extern "C" int otmb_facefluxes_slab(otmb_ctx* c, const double* umo, const double* vmo, double fill, const double* carry_in,
                                    double* carry_out, int32_t carry_on_device, int32_t valid_uv[2], double* east,
                                    double* west, double* north, double* south, double* top, double* bottom) {
    if (!valid_uv) return OTMB_ERR_BADARG;
    double* const outs[6] = {east, west, north, south, top, bottom};
    return facefluxes_impl(c, umo, vmo, fill, carry_in, carry_out, carry_on_device, valid_uv, outs);
}
