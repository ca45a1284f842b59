// fused_v4.cu — OTMB_PATH_FUSED: single-pass direct CSC assembly, one thread per WET cell.
//
// Same mathematics and ordering rules as fused.cu (read its header first: gather form, emit order, generic
// branch for coincident neighbours).  The schedule is the result of the ncu profiles under profiles/
// (history in profiles/README.md).  The fully unrolled predecessor (fused_v2.cu) was 8 000 SASS instructions
// of straight-line code at 128 registers; a version with rolled loops over the candidates fixed the
// instruction-cache misses but paid for the warps that held a seam / fold cell.  What is here:
//   * per-thread candidate arrays (row index, running T value) live in shared memory, [candidate][thread],
//     conflict-free, so that only scalars stay in registers (80, two blocks per SM); the candidates' linear indices
//     are recomputed where needed — shared memory is paid for in L1 here: 64.75 KB per block lets two blocks use the
//     132 KB carve-out (L1 124 KB), with the index array it was the 164 KB one and 3.5 % slower;
//   * the row order inside a column is a function of the cell's CLASS only (regular T S W C E N B; west seam
//     T S C E W N B; east seam T S E W C N B; right half of the tripolar fold row T S N W C E B).  Each class
//     is four bit masks "rows that precede candidate c"; the staging position of an entry is a popcount of
//     the matrix's pattern under that mask, and the diagonal of Tadv adds the emitters' contributions in the
//     class's order (ascending wet rank = the reference's emit order).  One unrolled code path with
//     compile-time candidates serves every class; absent entries are skipped by a branch;
//   * neighbours are resolved through `rank3d` (Int32 wet rank per grid cell, -1 = dry — the reference's own
//     Lwet3D, /root/reference/src/matrixbuilding.jl:18-20);
//   * a dedicated scan warp per block runs the decoupled look-back for all five counters and releases each
//     column warp through its own named barrier, so no warp waits for a sibling or for the look-back while it
//     still has values to compute; the TκH values are computed ahead of that barrier;
//   * staging is warp-private and reused matrix by matrix; each warp flushes its own contiguous slice of
//     rowval / nzval with coalesced stores as soon as a matrix is staged;
//   * the lines of the next level are prefetched into L2 while a level is being assembled;
//   * the upwind switch is a template parameter (flux-sign tests are single comparisons); the TκH inputs are
//     requested in two batches of two directions, the vertical operators' inputs and the cell's own volume ahead of
//     the barrier / with the phase-0 batch; results leave with streaming stores; the scan warp sums a tile's counts
//     and a look-back window's aggregates with redux.sync, the per-warp offsets inside a tile are derived by the
//     column warps themselves (they are not on the path to the publication of the tile's aggregate);
//   * OTMB_V4_TIMELINE=<file> runs a debug instantiation that stamps every tile's phases (profiles/timeline.py).
//
// Launch geometry: one tile of TILE consecutive wet cells per block (+ the scan warp), tiles in block-index order
// (the decoupled look-back only waits on lower-numbered tiles, which are resident or finished).
// A launch covers the wet ranks [w0, w0 + ncols): the whole matrix on one GPU, or the columns of
// one k-slab when a matrix is sharded across GPUs (rows are global wet ranks either way).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <atomic>
#include <type_traits>

#include "fused_generic.cuh"
#include "fdiv.cuh"

using namespace fusedg;

namespace {

// look-back descriptor: [flag:2][epoch:20][count:42].  flag 1 = the tile's own aggregate, 2 = inclusive prefix; the
// epoch is the launch's serial number, so descriptors left behind by earlier launches read as "not published yet"
// and the array needs no memset per launch (it is zeroed when it is allocated and when the epoch wraps).
constexpr int ST_SHIFT = 42;
constexpr u64 ST_MASK = (1ull << ST_SHIFT) - 1;
constexpr unsigned ST_EPOCHS = 1u << 20;
constexpr int WDATA = 7 * 32;       // staged entries per warp: 7 per column
constexpr int WCAP = WDATA + 32;    // + one dump slot per lane for absent entries (branch-free staging)
// candidates in ascending row order, one nibble each, per class
constexpr unsigned ORD0 = 0x6543210u;   // T S W C E N B
constexpr unsigned ORD1 = 0x6524310u;   // T S C E W N B   (west seam: W wraps to the end of the row)
constexpr unsigned ORD2 = 0x6532410u;   // T S E W C N B   (east seam: E wraps to the start of the row)
constexpr unsigned ORD3 = 0x6432510u;   // T S N W C E B   (fold row, right half: N mirrors to the left of W)

struct FastDiv {
    u64 mul;
    unsigned shift;
};
__device__ __forceinline__ unsigned fdiv(unsigned n, FastDiv f) { return (unsigned)((n * f.mul) >> f.shift); }

struct V4Params {
    GridDims g;
    FastDiv divP, divNx;
    const double *v3D, *thk, *area2D, *zt, *edge, *dnbr, *mlotst, *rho3d;
    const double *pe, *pw, *pn, *ps, *pt, *pb;
    // flux array read at candidate c (the slot of the neighbour that points back at this cell):
    // T -> its bottom, S -> its north, W -> its east, (C unused), E -> its west, N -> its south, B -> its top,
    // [7] = north again, used for N on the tripolar fold
    const double* phi_nb[8];
    const int* rank3d;   // (M) global wet rank, -1 = dry
    const int* lwet;     // (ncols) linear index of the launch's wet cells
    double kH, kVML, kVdeep, rho;
    int upwind, base, build, prefetch;
    int ntiles;
    int w0;              // global wet rank of the launch's first column
    int ncols;
    // a slab context holds a WINDOW of the 3-D arrays (common.cuh): idle threads of the last tile read the cell Lsafe
    // (the window's first owned cell, i = 0) instead of cell 0, and the two-levels-ahead prefetch stops at Lmax
    int Lsafe, ksafe, psafe, Lmax;
    int lwet_ahead;      // columns ahead whose wet-list entries a tile prefetches into L2 (0 = off)
    i64* colptr[5];
    i64* rowval[5];
    double* nzval[5];
    DevFlags* flags;
    u64* tile_state;
    const u64* start;                // 5 entry totals of the launches before this one (chained slab launches), or null
    u64* run_out;                    // where the last tile leaves the totals including this launch, or null
    unsigned epoch;                  // look-back epoch of this launch (serial mod 2^20)
    long long* timeline;   // TLINE instantiation only (OTMB_V4_TIMELINE): 64 stamps per tile, see profiles/timeline.py
};

// look-back descriptors: device-scope relaxed accesses (a plain `volatile` access is system scope)
__device__ __forceinline__ u64 ld_vol(const u64* p) {
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_vol(u64* p, u64 v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
// upwind flux max(ϕ, 0) / min(ϕ, 0), or ϕ/2 when centred (:244-295).  Julia's max/min against the literal 0.0 reduce
// to one comparison: NaN propagates (both comparisons are false for NaN), max(-0.0, 0.0) = 0.0, min(-0.0, 0.0) = -0.0.
__device__ __forceinline__ double upflux(double x, bool take_max, bool up) {
    return up ? (take_max ? (!(x <= 0.0) ? x : 0.0) : (!(x > 0.0) ? x : 0.0)) : x / 2;
}
// "is that flux non-zero?" (the reference's `f > 0 || f < 0`, :245), without forming it: upwind -> the sign test
// itself; centred -> ϕ/2 rounds to zero only for |ϕ| <= the smallest subnormal.  NaN: no.
template <bool UP>
__device__ __forceinline__ bool inflow(double x, bool take_max) {
    if (UP) return take_max ? x > 0.0 : x < 0.0;
    return fabs(x) > __longlong_as_double(1ll);
}
// q1 = a1 / b1 and q2 = a2 / b2: PAIRED with overlapping chains (fdiv.cuh), otherwise two plain divisions
// (-DOTMB_NO_DIV2: plain everywhere)
template <bool PAIRED = true>
__device__ __forceinline__ void div_pair(const double a1, const double b1, const double a2, const double b2, double& q1, double& q2) {
#ifdef OTMB_NO_DIV2
    q1 = a1 / b1;
    q2 = a2 / b2;
#else
    if (PAIRED) {
        otmb_fdiv::div2(a1, b1, a2, b2, q1, q2);
    } else {
        q1 = a1 / b1;
        q2 = a2 / b2;
    }
#endif
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int TILE>
struct Smem {
    double val[TILE / 32][WCAP];   // warp-private staging: values
    double Tv[7][TILE];            // running T = ((Tadv + TκH) + TκVML) + TκVdeep, by candidate
    int row[TILE / 32][WCAP];      // warp-private staging: row indices (+ index base)
    int rk[7][TILE];               // row index (+ index base) of candidate c
    u64 lexcl[TILE];               // in-warp exclusive offsets of the column, five 12-bit fields
    u64 warp[TILE / 32];           // entries of each warp, five 12-bit fields
    u64 excl[5];
};

// ---------------------------------------------------------------------------------------
template <bool RHO3D, bool UP, int TILE, int MINB, int KHB = 2, bool VAH = true, bool VC0 = true, bool TLINE = false, bool XSM = true>
// __grid_constant__: P is indexed dynamically and its address is taken by the generic branch
__global__ void __launch_bounds__(TILE + 32, MINB) k_fused_v4(const __grid_constant__ V4Params P) {
    constexpr int NW = TILE / 32;   // column warps; warp NW is the scan warp
    static_assert(TILE * 7 < 4096, "per-warp / per-tile entry counts are kept in 12-bit fields");
    constexpr int BG = (NW + 13) / 14;          // column warps per release barrier (ids 2 .. 15)
    constexpr int NBG = (NW + BG - 1) / BG;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem<TILE>& S = *reinterpret_cast<Smem<TILE>*>(smem_raw);

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int tile = blockIdx.x;
    // the completion record (k_publish, launched with programmatic stream serialisation) may be scheduled as soon as
    // every tile has started; it waits for this grid to complete before it reads anything
    asm volatile("griddepcontrol.launch_dependents;");
    // timeline instrumentation (debug instantiation): SM clock stamps of one tile's phases
    long long* const tl = TLINE ? P.timeline + (size_t)tile * 64 : nullptr;
    // (a clock read right behind BAR.SYNC.DEFER_BLOCKING issues before the barrier resolves: stamps behind a barrier
    // take a value loaded from shared memory AFTER it as an input, which orders them)
    auto stamp = [&](const int slot, const u64 dep = 0) {
        if (TLINE) {
            long long t;
            asm volatile("mov.u64 %0, %%clock64;" : "=l"(t) : "l"(dep) : "memory");
            if (lane == 0) tl[slot] = t;
        }
    };
    if (TLINE && tid == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
        tl[0] = (long long)gt;
        tl[1] = clock64();
        unsigned smid;
        asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
        tl[6] = smid;
    }

    // ================= the scan warp: decoupled look-back for all five counters =================
    // It owns no columns, so the column warps never wait for a warp that still has values to compute:
    // by the time they reach the second barrier the offsets have long been resolved.
    if (wid == NW) {
        // producer / consumer named barriers: the column warps only ARRIVE at barrier 1 (no wait) once their
        // counts are in S.warp; the scan warp only arrives at barrier 2 once the offsets are in S.excl
        asm volatile("bar.sync 1, %0;" ::"r"(TILE + 32) : "memory");
        stamp(2, TLINE ? *reinterpret_cast<volatile u64*>(&S.warp[0]) : 0);
        // the tile's five totals: lane q < NW holds the packed counts of column warp q, one redux.sync per counter; lane
        // m < 5 publishes counter m.  (The in-tile offsets of the column warps are NOT computed here: every column warp
        // derives its own from S.warp after its release barrier, which keeps them off the path to the publication.)
        const u64 pk = lane < NW ? *reinterpret_cast<volatile u64*>(&S.warp[lane]) : 0ull;
        unsigned agg_m = 0;
#pragma unroll
        for (int m = 0; m < 5; ++m) {
            const unsigned tot = __reduce_add_sync(0xffffffffu, (unsigned)((pk >> (12 * m)) & 0xfffull));
            if (lane == m) agg_m = tot;
        }
        const unsigned tagA = (1u << 20) | P.epoch, tagP = (2u << 20) | P.epoch;   // the descriptor's top 22 bits
        // entries of earlier launches of the same build (slab-pipelined builds): the first tile's exclusive prefix
        const u64 first = (tile == 0 && P.start != nullptr && lane < 5) ? P.start[lane] : 0ull;
        if (lane < 5) st_vol(P.tile_state + (size_t)tile * 8 + lane, ((u64)(tile == 0 ? tagP : tagA) << ST_SHIFT) | (first + (u64)agg_m));
        stamp(3);
        int rounds = 0;
        u64 excl[5] = {0, 0, 0, 0, 0};
        unsigned pending = tile > 0 ? 31u : 0u;   // counters still looking back
        int look = tile - 1;
        // Each step covers LB_SUB windows of 32 tiles with one round trip (all loads are issued before any is tested).
        // Measured: 1 window 0.466 ms, 2 windows 0.494 ms, 4 windows 0.506 ms — the wider the window, the more scan
        // warps poll the same hot descriptor lines and the later a publisher's store becomes visible.
        constexpr int LB_SUB = 1;
        while (pending) {
            u64 wv[LB_SUB][5];
#pragma unroll
            for (int s = 0; s < LB_SUB; ++s)
#pragma unroll
                for (int m = 0; m < 5; ++m) wv[s][m] = (look - 32 * s - lane) >= 0 ? 0ull : ((u64)tagP << ST_SHIFT);   // before the first tile: prefix 0
            bool again;
            unsigned spins = 0;
            // published by THIS launch: the epoch matches and the flag is set
            auto ready = [&](const u64 v) {
                const unsigned hi = (unsigned)(v >> ST_SHIFT);
                return hi == tagA || hi == tagP;
            };
            do {   // the five descriptors of a tile share a 64-byte line
#pragma unroll
                for (int s = 0; s < LB_SUB; ++s)
#pragma unroll
                    for (int m = 0; m < 5; ++m)
                        if ((pending >> m & 1) && !ready(wv[s][m]))
                            wv[s][m] = ld_vol(P.tile_state + (size_t)(look - 32 * s - lane) * 8 + m);
                again = false;
#pragma unroll
                for (int s = 0; s < LB_SUB; ++s)
#pragma unroll
                    for (int m = 0; m < 5; ++m) again |= (pending >> m & 1) && !ready(wv[s][m]);
                // Lower tiles are resident or finished (blocks start in index order), so this terminates; the bound
                // (seconds) turns a broken assumption into an error code instead of a hang.
                if (++spins > (1u << 24)) break;
            } while (__any_sync(0xffffffffu, again));
            if (__any_sync(0xffffffffu, spins > (1u << 24))) {
                // give up: report it, and zero the warps' entry counts so that no column warp flushes anything
                // (the offsets are meaningless); the host turns the flag into an error code
                if (lane == 0) atomicOr(&P.flags->lookback_timeout, 1);
                if (lane < NW) S.warp[lane] = 0ull;
                break;
            }
#pragma unroll
            for (int s = 0; s < LB_SUB; ++s) {   // nearest window first; a counter stops at its first inclusive prefix
#pragma unroll
                for (int m = 0; m < 5; ++m) {
                    if (!(pending >> m & 1)) continue;
                    // a tile's aggregate fits 13 bits, 32 of them one redux.sync; only an inclusive prefix needs 62 bits
                    const u64 val = wv[s][m] & ST_MASK;
                    const unsigned pm = __ballot_sync(0xffffffffu, (unsigned)(wv[s][m] >> ST_SHIFT) == tagP);
                    if (pm) {
                        const int first = __ffs(pm) - 1;
                        excl[m] += (u64)__reduce_add_sync(0xffffffffu, lane < first ? (unsigned)val : 0u) + __shfl_sync(0xffffffffu, val, first);
                        pending &= ~(1u << m);
                    } else {
                        excl[m] += (u64)__reduce_add_sync(0xffffffffu, (unsigned)val);
                    }
                }
            }
            look -= 32 * LB_SUB;
            ++rounds;
        }
        stamp(4);
        if (TLINE && lane == 0) tl[5] = rounds;
        const u64 mine = first + (lane == 0 ? excl[0] : lane == 1 ? excl[1] : lane == 2 ? excl[2] : lane == 3 ? excl[3] : excl[4]);
        if (lane < 5) {
            const u64 agg = agg_m;
            if (tile > 0) st_vol(P.tile_state + (size_t)tile * 8 + lane, ((u64)tagP << ST_SHIFT) | (mine + agg));
            S.excl[lane] = mine;
            if (tile == P.ntiles - 1) {
                P.flags->nnz[lane] = mine + agg;
                if (P.run_out) P.run_out[lane] = mine + agg;
                if (P.build >> lane & 1) P.colptr[lane][P.ncols] = (i64)(mine + agg) + P.base;
            }
        }
        __threadfence_block();
        // S.excl published: release the column warps through their own barriers (ids 2 .. 15, one per
        // warp when the tile has at most 14 of them; participants: the warps of the group + this one), so that no
        // column warp waits for a sibling's values
#pragma unroll
        for (int gq = 0; gq < NBG; ++gq) {
            const int members = (gq + 1) * BG <= NW ? BG : NW - gq * BG;
            asm volatile("bar.arrive %0, %1;" ::"r"(2 + gq), "r"(32 * members + 32) : "memory");
        }
        return;
    }
    const GridDims g = P.g;
    const int PP = g.P;
    const int w = tile * TILE + tid;       // column of this launch
    const bool valid = w < P.ncols;
    const int rC = P.w0 + w;               // global wet rank = row/column index
    constexpr bool up = UP;

    // ================= phase 0: pattern =================
    int L = P.Lsafe, k = P.ksafe, p2 = P.psafe;
    unsigned wetm = 0, act = 0, mlm = 0;
    unsigned ord = ORD0;
    bool fold = false, generic = false;
    unsigned m_T = 0, m_adv = 0, m_kh = 0, m_ml = 0, m_dp = 0;
    unsigned errbits = 0;  // 1 dry nbr, 2 nan adv, 4 nan kh, 8 nan ml, 16 nan deep, 32 zero dropped, 64 nan rho
    u64 packed = 0;        // 5 counts, 12 bits each
    double vC0 = 0.0;      // own volume, requested with the phase-0 batch
    if (valid) {
        int Lc[7], r[7];
        L = __ldg(P.lwet + w);
        // the wet list of the tile that will take this slot a few waves from now: one L2 prefetch per thread, so that
        // the tile's very first load (everything else depends on it) is an L2 hit instead of a DRAM round trip
        if (P.lwet_ahead > 0 && w + P.lwet_ahead < P.ncols) prefetch_l2(P.lwet + w + P.lwet_ahead);
        k = (int)fdiv((unsigned)L, P.divP);
        p2 = L - k * PP;
        const int j = (int)fdiv((unsigned)p2, P.divNx);
        const int i = p2 - j * g.nx;
        fold = (j == g.ny - 1) && (g.topo == OTMB_TOPO_TRIPOLAR);
        const bool hasT = k > 0, hasB = k < g.nz - 1, hasS = j > 0, hasN = (j < g.ny - 1) || fold;
        const bool seamW = i == 0, seamE = i == g.nx - 1;
        ord = seamW ? ORD1 : seamE ? ORD2 : (fold && (g.nx - 1 - i < i)) ? ORD3 : ORD0;
        Lc[cC] = L;
        Lc[cT] = hasT ? L - PP : L;
        Lc[cB] = hasB ? L + PP : L;
        Lc[cS] = hasS ? L - g.nx : L;
        Lc[cW] = seamW ? L + (g.nx - 1) : L - 1;
        Lc[cE] = seamE ? L - (g.nx - 1) : L + 1;
        Lc[cN] = (j < g.ny - 1) ? L + g.nx : (fold ? L + (g.nx - 1 - 2 * i) : L);
        // ---- L2 prefetch one level ahead.  Tiles sweep the wet cells level by level, so the lines of level
        // k+1 (k+2 for what is read at the bottom neighbour) are what the tiles one sweep-window later will
        // demand; asking for them now turns those DRAM-latency loads into L2 hits.  Ocean columns are wet
        // from the surface down, so a wet cell at k+1 always has this wet cell above it to do the asking.
        if (P.prefetch && hasB) {
            const int L1 = L + PP;
            prefetch_l2(P.pe + L1);
            prefetch_l2(P.pw + L1);
            prefetch_l2(P.pn + L1);
            prefetch_l2(P.ps + L1);
            prefetch_l2(P.pb + L1);
            prefetch_l2(P.thk + L1);
            if (RHO3D) prefetch_l2(P.rho3d + L1);
            const int L2 = min(k < g.nz - 2 ? L1 + PP : L1, P.Lmax);
            prefetch_l2(P.pt + L2);
            prefetch_l2(P.v3D + L2);
            prefetch_l2(P.rank3d + L2);
        }
        // ---- loads: neighbour ranks, the six face fluxes the neighbours carry, mixed-layer inputs
        const int qT = __ldg(P.rank3d + Lc[cT]), qS = __ldg(P.rank3d + Lc[cS]), qW = __ldg(P.rank3d + Lc[cW]),
                  qE = __ldg(P.rank3d + Lc[cE]), qN = __ldg(P.rank3d + Lc[cN]), qB = __ldg(P.rank3d + Lc[cB]);
        const double xT = __ldg(P.pb + Lc[cT]);                       // emitter above: its Bottom slot, max
        const double xS = __ldg(P.pn + Lc[cS]);                       // its North slot, min
        const double xW = __ldg(P.pe + Lc[cW]);                       // its East slot, min
        const double xE = __ldg(P.pw + Lc[cE]);                       // its West slot, max
        const double xN = __ldg((fold ? P.pn : P.ps) + Lc[cN]);       // its South slot (max), or North on the fold (min)
        const double xB = __ldg(P.pt + Lc[cB]);                       // emitter below: its Top slot, min
        const double oW = __ldg(P.pw + L), oE = __ldg(P.pe + L), oS = __ldg(P.ps + L), oN = __ldg(P.pn + L),
                     oB = __ldg(P.pb + L), oT = __ldg(P.pt + L);   // own faces, for the dry-neighbour check below
        if (XSM) {   // the neighbours' fluxes wait for Tadv in the (still unused) running-T slots instead of being re-loaded
            S.Tv[cT][tid] = xT, S.Tv[cS][tid] = xS, S.Tv[cW][tid] = xW, S.Tv[cE][tid] = xE, S.Tv[cN][tid] = xN, S.Tv[cB][tid] = xB;
        }
        const double ml = __ldg(P.mlotst + p2);
        if (VC0) vC0 = __ldg(P.v3D + L);
        const double z0 = __ldg(P.zt + k), zT = __ldg(P.zt + (hasT ? k - 1 : k)), zB = __ldg(P.zt + (hasB ? k + 1 : k));
        r[cC] = rC;
        r[cT] = hasT ? qT : -1;
        r[cS] = hasS ? qS : -1;
        r[cW] = qW;
        r[cE] = qE;
        r[cN] = hasN ? qN : -1;
        r[cB] = hasB ? qB : -1;
#pragma unroll
        for (int c = 0; c < 7; ++c) {
            if (c != cC && r[c] >= 0) wetm |= 1u << c;
            S.rk[c][tid] = r[c] + P.base;
        }
        // face flux each neighbour carries through the face it shares with this cell (the value the
        // reference reads at the neighbour, :244-295): only the sign pattern is needed here
        if (P.build & 2) {
            if ((wetm & bT) && inflow<UP>(xT, true)) act |= bT;
            if ((wetm & bS) && inflow<UP>(xS, false)) act |= bS;
            if ((wetm & bW) && inflow<UP>(xW, false)) act |= bW;
            if ((wetm & bE) && inflow<UP>(xE, true)) act |= bE;
            if ((wetm & bN) && inflow<UP>(xN, !fold)) act |= bN;
            if ((wetm & bB) && inflow<UP>(xB, false)) act |= bB;
            // own faces that point at a dry or absent cell: the reference would push `missing` (:247-250)
            // (the six own faces were loaded with the first batch: a coastal warp does not pay a second round trip)
            const unsigned dry = ~wetm;
            bool bad = false;
            if (dry & bW) bad |= inflow<UP>(oW, true);
            if (dry & bE) bad |= inflow<UP>(oE, false);
            if (dry & bS) bad |= inflow<UP>(oS, true);
            if (dry & bN) bad |= inflow<UP>(oN, false);
            if (dry & bB) bad |= inflow<UP>(oB, true);
            if ((dry & bT) && hasT) bad |= inflow<UP>(oT, false);
            if (bad) errbits |= 1u;
        }
        // mixed-layer mask Ω = zt[k] < mlotst[i,j] (false for NaN / missing), :85
        if ((P.build & 8) && z0 < ml) {
            if ((wetm & bT) && zT < ml) mlm |= bT;
            if ((wetm & bB) && zB < ml) mlm |= bB;
        }
        // patterns (bit cC = diagonal)
        if (P.build & 2) m_adv = act ? (act | bC) : 0u;
        if (P.build & 4) m_kh = (wetm & HMASK) ? ((wetm & HMASK) | bC) : 0u;
        if (P.build & 8) m_ml = mlm ? (mlm | bC) : 0u;
        if (P.build & 16) m_dp = (wetm & VMASK) ? ((wetm & VMASK) | bC) : 0u;
        if (P.build & 1) m_T = m_adv | m_kh | m_ml | m_dp;
        // coincident neighbours inside the grid row (seam, fold, nx <= 2): generic branch
        if (ord != ORD0 || fold) {
            const bool pW = wetm & bW, pE = wetm & bE, pNf = fold && (wetm & bN);
            generic = (pW && r[cW] == rC) || (pE && r[cE] == rC) || (pW && pE && r[cW] == r[cE]) ||
                      (pNf && (r[cN] == rC || (pW && r[cN] == r[cW]) || (pE && r[cN] == r[cE])));
        }
        int c0, c1, c2, c3, c4;
        if (!generic) {
            c0 = __popc(m_T), c1 = __popc(m_adv), c2 = __popc(m_kh), c3 = __popc(m_ml), c4 = __popc(m_dp);
        } else {
            int g_r[7];
#pragma unroll
            for (int c = 0; c < 7; ++c) g_r[c] = r[c];
            c0 = distinct_rows(m_T, g_r), c1 = distinct_rows(m_adv, g_r), c2 = distinct_rows(m_kh, g_r),
            c3 = distinct_rows(m_ml, g_r), c4 = distinct_rows(m_dp, g_r);
        }
        packed = (u64)c0 | ((u64)c1 << 12) | ((u64)c2 << 24) | ((u64)c3 << 36) | ((u64)c4 << 48);
    } else {
#pragma unroll
        for (int c = 0; c < 7; ++c) {
            S.rk[c][tid] = 0;
        }
    }

    // Linear index of candidate c (clamped to a valid cell), recomputed from L, k, p2 and the class where it is needed:
    // two or three integer instructions each.  It used to be kept in shared memory; without that array two blocks fit
    // the 132 KB shared-memory carve-out instead of 164 KB, which leaves L1 124 KB instead of 92 KB.
    if (!valid) ord = ORD1;   // L = Lsafe (i = 0): keeps every neighbour index of an idle thread inside the window
    // first / last cell of a grid row, from the class: ORD1 = west seam; ORD2 = east seam unless the row has one cell
    const bool atW = ord == ORD1, atE = ord == ORD2 || g.nx == 1;
    auto Lc_of = [&](const int c) -> int {
        return c == cT   ? (k > 0 ? L - PP : L)
               : c == cB ? (k < g.nz - 1 ? L + PP : L)
               : c == cS ? (p2 >= g.nx ? L - g.nx : L)
               : c == cW ? (atW ? L + (g.nx - 1) : L - 1)
               : c == cE ? (atE ? L - (g.nx - 1) : L + 1)
               : c == cN ? (p2 < PP - g.nx ? L + g.nx : (fold ? L + (g.nx - 1 - 2 * (p2 - (g.ny - 1) * g.nx)) : L))
                         : L;
    };

    // ================= tile scan + decoupled look-back =================
    u64 incl = packed;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u64 o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) S.warp[wid] = incl;
    __threadfence_block();
    asm volatile("bar.arrive 1, %0;" ::"r"(TILE + 32) : "memory");   // counts published; nobody waits here
    stamp(8 + 4 * wid);
    // kept in shared memory, not in registers: they are needed again only at the five flushes, and as
    // registers they were spilled to local memory (ncu: 25 % of the long-scoreboard stalls were their reloads)
    S.lexcl[tid] = incl - packed;                              // in-warp exclusive offsets of this column
    // ================= phase 1: operator by operator: walk -> warp staging -> flush =================
    int* const srow = S.row[wid];
    double* const sval = S.val[wid];
    const int dump = WDATA + lane;
    const double vC = VC0 ? vC0 : __ldg(P.v3D + L);
    const int rkC = rC + P.base;
    GenOut gen;      // local memory, only touched by generic columns

    if (valid && generic) {
        double g_p[7];
        int g_Lc[7], g_r[7];
        unsigned g_err = 0;
        for (int c = 0; c < 7; ++c) {
            g_Lc[c] = c == 0 ? Lc_of(0) : c == 1 ? Lc_of(1) : c == 2 ? Lc_of(2) : c == 3 ? Lc_of(3) : c == 4 ? Lc_of(4) : c == 5 ? Lc_of(5) : Lc_of(6);
            g_r[c] = S.rk[c][tid] - P.base;
            const bool mx = c == cT || c == cE || (c == cN && !fold);
            const double f = c == cC ? 0.0 : upflux(__ldg(P.phi_nb[(c == cN && fold) ? 7 : c] + g_Lc[c]), mx, up);
            g_p[c] = mx ? f : -f;
        }
        generic_full(P, L, k, g_Lc, g_r, wetm, fold, act, g_p, mlm, &gen, &g_err);
        errbits |= g_err;
        atomicAdd(&P.flags->generic_columns, 1);
        m_T = m_adv = m_kh = m_ml = m_dp = 0;   // the walks stage nothing for this column
        wetm = act = mlm = 0;
    }

    // Rows that precede candidate c inside a column, per class (see ORD0..ORD3): the staging position of an entry
    // is the popcount of the matrix's mask under this mask.  T, S and B are the same in every class.
    unsigned lowW, lowC, lowE, lowN;
    {
        const unsigned TS = bT | bS;
        if (ord == ORD0) {            // T S W C E N B
            lowW = TS; lowC = TS | bW; lowE = TS | bW | bC; lowN = TS | bW | bC | bE;
        } else if (ord == ORD1) {     // T S C E W N B
            lowC = TS; lowE = TS | bC; lowW = TS | bC | bE; lowN = TS | bW | bC | bE;
        } else if (ord == ORD2) {     // T S E W C N B
            lowE = TS; lowW = TS | bE; lowC = TS | bE | bW; lowN = TS | bW | bC | bE;
        } else {                      // T S N W C E B
            lowN = TS; lowW = TS | bN; lowC = TS | bN | bW; lowE = TS | bN | bW | bC;
        }
    }
    auto low_of = [&](const int c) -> unsigned {
        return c == cT ? 0u : c == cS ? bT : c == cW ? lowW : c == cC ? lowC : c == cE ? lowE : c == cN ? lowN : 0x3fu;
    };

    // copies a generic column's entries of matrix q behind the regular ones
    auto stage_generic = [&](const int q, const int off) {
        if (generic)
            for (int a = 0; a < gen.cnt[q]; ++a) {
                srow[off + a] = gen.rows[q][a] + P.base;
                sval[off + a] = gen.vals[q][a];
            }
    };
    // flush of the staged slice: colptr of the 32 columns, then rowval / nzval, coalesced
    // (iters = compile-time bound on ceil(entries of a warp / 32): the matrix's maximum entries per column)
    auto flush = [&](auto iters, const int q, const int off) {
        constexpr int IT = decltype(iters)::value;
        // offset of this warp inside the tile: the counts of the column warps before it
        const u64 pkq = lane < wid ? *reinterpret_cast<volatile u64*>(&S.warp[lane]) : 0ull;
        const u64 g0 = *reinterpret_cast<volatile u64*>(&S.excl[q]) + __reduce_add_sync(0xffffffffu, (unsigned)((pkq >> (12 * q)) & 0xfffull));
        const int n = (int)((S.warp[wid] >> (12 * q)) & 0xfffull);
        if (valid) P.colptr[q][w] = (i64)(g0 + (u64)off) + P.base;
        __syncwarp();
        i64* __restrict__ rv = P.rowval[q] + g0 + lane;
        double* __restrict__ nv = P.nzval[q] + g0 + lane;
        const int* sr = srow + lane;
        const double* sv = sval + lane;
#pragma unroll
        for (int it = 0; it < IT; ++it) {
            if (lane + 32 * it < n) {
                // streaming stores: the results are never read again by this kernel; keep L2 for the inputs
                __stcs(rv + 32 * it, (i64)(unsigned)sr[32 * it]);
                __stcs(nv + 32 * it, sv[32 * it]);
            }
        }
        __syncwarp();   // the staging buffer may be overwritten by the next matrix
    };
    auto offset_of = [&](const int q) { return (int)((S.lexcl[tid] >> (12 * q)) & 0xfffull); };

    // ---- Tadv (:193-204, :237-297).
    // Off-diagonal (𝑖, 𝑗) = -p/m𝑖 as the emitter 𝑖 computes it; the diagonal adds p/m𝑗 per emitter,
    // sparse! keeping the first value and adding the later ones in that order.
    const int off1 = offset_of(1);
    if (P.build & 2) {
        const double rhoC = RHO3D ? __ldg(P.rho3d + L) : P.rho;
        if (RHO3D && valid && isnan(rhoC)) errbits |= 64u;
        double dsum = 0.0;
        bool first = true, bad = false;
        int posC = dump;
        {
            // unrolled with compile-time candidates: the loads are batched, absent faces are skipped, the staging
            // position is a popcount under the class's row-order mask
            double xs[7], vn[7], rn[7];
#pragma unroll
            for (int c = 0; c < 7; ++c) {
                if (c == cC) continue;
                const int Lc = Lc_of(c);
                const double* ph = c == cT ? P.pb : c == cS ? P.pn : c == cW ? P.pe : c == cE ? P.pw : c == cB ? P.pt
                                                                                               : (fold ? P.pn : P.ps);
                xs[c] = XSM ? S.Tv[c][tid] : __ldg(ph + Lc);
                vn[c] = __ldg(P.v3D + Lc);
                rn[c] = RHO3D ? __ldg(P.rho3d + Lc) : P.rho;
            }
            auto add = [&](const bool on, const double d) {
                if (on) {
                    dsum = first ? d : dsum + d;
                    first = false;
                }
            };
            double dd[7];
#pragma unroll
            for (int c = 0; c < 7; ++c) {
                dd[c] = 0.0;
                if (c == cC) continue;
                if ((m_adv >> c) & 1) {
                    const bool mx = c == cT || c == cE || (c == cN && !fold);
                    const double f = UP ? xs[c] : xs[c] / 2;   // active: max / min picked the flux itself
                    const double p = mx ? f : -f;
                    const double rb = (rn[c] + rhoC) / 2;
                    double a;
                    // not paired: this loop holds the six neighbours' fluxes and volumes, and the second chain's
                    // registers would spill (measured: 14 bytes of spills, +2.5 %)
                    div_pair<false>(-p, rb * vn[c], p, rb * vC, a, dd[c]);
                    bad |= isnan(a) || isnan(dd[c]);
                    const int pp = off1 + __popc(m_adv & low_of(c));
                    srow[pp] = S.rk[c][tid];
                    sval[pp] = a;
                    S.Tv[c][tid] = a;
                } else {
                    S.Tv[c][tid] = 0.0;
                }
            }
            // the diagonal adds the emitters' contributions in ascending wet rank = row order of the class
            add(m_adv & bT, dd[cT]);
            add(m_adv & bS, dd[cS]);
            if (ord == ORD0) {
                add(m_adv & bW, dd[cW]);
                add(m_adv & bE, dd[cE]);
                add(m_adv & bN, dd[cN]);
            } else if (ord == ORD3) {
                add(m_adv & bN, dd[cN]);
                add(m_adv & bW, dd[cW]);
                add(m_adv & bE, dd[cE]);
            } else {
                add(m_adv & bE, dd[cE]);
                add(m_adv & bW, dd[cW]);
                add(m_adv & bN, dd[cN]);
            }
            add(m_adv & bB, dd[cB]);
            posC = (m_adv & bC) ? off1 + __popc(m_adv & lowC) : dump;
        }
        if (bad) errbits |= 2u;
        srow[posC] = rkC;
        sval[posC] = dsum;
        S.Tv[cC][tid] = (m_adv & bC) ? dsum : 0.0;
        stage_generic(1, off1);
    } else {
#pragma unroll
        for (int c = 0; c < 7; ++c) S.Tv[c][tid] = 0.0;
    }

    // ---- TκH values of regular rows, computed ahead of the barrier so that the look-back has time to finish
    // (:348-415, :426-435): own slots in emit order W,E,S,N with compile-time directions
    double khv[4] = {0.0, 0.0, 0.0, 0.0}, kh_dsum = 0.0;
    if (P.build & 4) {
        const double thC = __ldg(P.thk + L);
        bool first = true, bad = false;
        // loads of KHB directions are issued as one batch (one round trip), then their values are computed;
        // the S / N neighbours live on other cache lines than the cell itself, W / E mostly on its own
#pragma unroll
        for (int q0 = 0; q0 < 4; q0 += KHB) {
            double e_own[KHB], d_own[KHB], e_opp[KHB], d_opp[KHB], th[KHB], vnb[KHB];
#pragma unroll
            for (int u = 0; u < KHB; ++u) {
                const int q = q0 + u;
                const int c = q == 0 ? cW : q == 1 ? cE : q == 2 ? cS : cN;
                const int own = c == cW ? OTMB_DIR_WEST : c == cE ? OTMB_DIR_EAST : c == cS ? OTMB_DIR_SOUTH : OTMB_DIR_NORTH;
                const int Lc = Lc_of(c);
                const int q2 = Lc - k * PP;
                // unconditional loads (clamped indices)
                e_own[u] = __ldg(P.edge + own * PP + p2);
                d_own[u] = __ldg(P.dnbr + own * PP + p2);
                const double* e_oppp = c == cW ? P.edge + OTMB_DIR_EAST * PP : c == cE ? P.edge + OTMB_DIR_WEST * PP
                                       : c == cS ? P.edge + OTMB_DIR_NORTH * PP
                                                 : P.edge + (fold ? OTMB_DIR_NORTH : OTMB_DIR_SOUTH) * PP;
                e_opp[u] = __ldg(e_oppp + q2);
                d_opp[u] = __ldg(e_oppp + (P.dnbr - P.edge) + q2);
                th[u] = __ldg(P.thk + Lc);
                vnb[u] = __ldg(P.v3D + Lc);
            }
#pragma unroll
            for (int u = 0; u < KHB; ++u) {
                const int q = q0 + u;
                const int c = q == 0 ? cW : q == 1 ? cE : q == 2 ? cS : cN;
                if ((m_kh >> c) & 1) {
                    const double ka = P.kH * jl_min(thC * e_own[u], th[u] * e_opp[u]);
                    double ts, tn;   // row 𝑗 seen from 𝑗, row 𝑖 seen from 𝑖
                    div_pair(ka, d_own[u] * vC, ka, d_opp[u] * vnb[u], ts, tn);
                    bad |= isnan(ts) || isnan(tn);
                    kh_dsum = first ? ts : kh_dsum + ts;
                    first = false;
                    khv[q] = -tn;
                }
            }
        }
        if (bad) errbits |= 4u;
    }

    // ---- inputs of the vertical operators, requested ahead of the barrier: they arrive while the warp waits for
    // its offsets and flushes Tadv / TκH (clamped indices, unconditional)
    double vt_area = 0.0, vt_dB = 0.0, vt_dT = 0.0, vt_vB = 0.0, vt_vT = 0.0;
    if (VAH && (P.build & 24)) {
        vt_area = __ldg(P.area2D + p2);
        const double ztC = __ldg(P.zt + k);
        vt_dB = fabs(ztC - __ldg(P.zt + (k < g.nz - 1 ? k + 1 : k)));
        vt_dT = fabs(ztC - __ldg(P.zt + (k > 0 ? k - 1 : k)));
        vt_vB = __ldg(P.v3D + Lc_of(cB));
        vt_vT = __ldg(P.v3D + Lc_of(cT));
    }

    // meet the scan warp's offsets (S.excl) at this warp's own barrier; it has normally arrived long ago
    stamp(9 + 4 * wid);
    {
        const int gq = wid / BG;
        const int members = (gq + 1) * BG <= NW ? BG : NW - gq * BG;
        asm volatile("bar.sync %0, %1;" ::"r"(2 + gq), "r"(32 * members + 32) : "memory");
    }
    stamp(10 + 4 * wid, TLINE ? *reinterpret_cast<volatile u64*>(&S.excl[0]) : 0);

    if (tile * TILE + wid * 32 < P.ncols) {
        if (P.build & 2) flush(std::integral_constant<int, 7>{}, 1, off1);

        // ---- TκH (:348-415, :426-435): stage what was computed ahead of the barrier.  Off-diagonal (𝑖, 𝑗) = -t as 𝑖
        // computes it; the diagonal sums 𝑗's own slots in emit order W,E,S,N.
        if (P.build & 4) {
            const int off2 = offset_of(2);
            int posC = dump;
            double dsum = 0.0;
            // the values were computed ahead of the look-back barrier (khv, kh_dsum)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int c = q == 0 ? cW : q == 1 ? cE : q == 2 ? cS : cN;
                if ((m_kh >> c) & 1) {
                    const int pp = off2 + __popc(m_kh & low_of(c));
                    srow[pp] = S.rk[c][tid];
                    sval[pp] = khv[q];
                    S.Tv[c][tid] = S.Tv[c][tid] + khv[q];
                }
            }
            dsum = kh_dsum;
            posC = (m_kh & bC) ? off2 + __popc(m_kh & lowC) : dump;
            srow[posC] = rkC;
            sval[posC] = dsum;
            if (m_kh & bC) S.Tv[cC][tid] = S.Tv[cC][tid] + dsum;
            stage_generic(2, off2);
            flush(std::integral_constant<int, 5>{}, 2, off2);
        }

        // ---- TκVML and TκVdeep (:450-477); own slots in emit order B, T; T folds TκVML before TκVdeep
        if (P.build & 24) {
            double dpT = 0.0, dpB = 0.0, dps = 0.0, mlT = 0.0, mlB = 0.0, mls = 0.0;
            if (m_dp | m_ml) {
                const double area = VAH ? vt_area : __ldg(P.area2D + p2), ztC = VAH ? 0.0 : __ldg(P.zt + k);
                bool firstm = true, firstd = true, badm = false, badd = false;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int c = q == 0 ? cB : cT;
                    if (!(wetm >> c & 1)) continue;
                    const double d = VAH ? (c == cT ? vt_dT : vt_dB) : fabs(ztC - __ldg(P.zt + (c == cT ? k - 1 : k + 1)));
                    const double qs = d * vC, qn = d * (VAH ? (c == cT ? vt_vT : vt_vB) : __ldg(P.v3D + Lc_of(c)));
                    if (m_dp) {
                        const double ka = P.kVdeep * area;
                        double ts, tn;
                        div_pair(ka, qs, ka, qn, ts, tn);
                        badd |= isnan(ts) || isnan(tn);
                        dps = firstd ? ts : dps + ts;
                        firstd = false;
                        if (c == cT) dpT = -tn; else dpB = -tn;
                    }
                    if (mlm >> c & 1) {
                        const double ka = P.kVML * area;
                        double ts, tn;
                        div_pair(ka, qs, ka, qn, ts, tn);
                        badm |= isnan(ts) || isnan(tn);
                        mls = firstm ? ts : mls + ts;
                        firstm = false;
                        if (c == cT) mlT = -tn; else mlB = -tn;
                    }
                }
                if (badm) errbits |= 8u;
                if (badd) errbits |= 16u;
            }
            // rows of a vertical operator: T, C, B in every class
            auto emit_vertical = [&](const int q, const unsigned m, const double vT, const double vCd, const double vB) {
                const int off = offset_of(q);
                int pos = off;
                int pp = (m & bT) ? pos : dump;
                srow[pp] = S.rk[cT][tid];
                sval[pp] = vT;
                pos += (m & bT) ? 1 : 0;
                pp = (m & bC) ? pos : dump;
                srow[pp] = rkC;
                sval[pp] = vCd;
                pos += (m & bC) ? 1 : 0;
                pp = (m & bB) ? pos : dump;
                srow[pp] = S.rk[cB][tid];
                sval[pp] = vB;
                if (m & bT) S.Tv[cT][tid] = S.Tv[cT][tid] + vT;
                if (m & bC) S.Tv[cC][tid] = S.Tv[cC][tid] + vCd;
                if (m & bB) S.Tv[cB][tid] = S.Tv[cB][tid] + vB;
                stage_generic(q, off);
                flush(std::integral_constant<int, 3>{}, q, off);
            };
            if (P.build & 8) emit_vertical(3, m_ml, mlT, mls, mlB);
            if (P.build & 16) emit_vertical(4, m_dp, dpT, dps, dpB);
        }

        // ---- T: union pattern; exact zeros are flagged and removed by the compaction pass
        if (P.build & 1) {
            const int off0 = offset_of(0);
            bool zero = false;
#pragma unroll
            for (int c = 0; c < 7; ++c) {
                const bool on = (m_T >> c) & 1;
                const double v = S.Tv[c][tid];
                zero |= on && (v == 0.0);
                const int pp = on ? off0 + __popc(m_T & low_of(c)) : dump;
                srow[pp] = S.rk[c][tid];
                sval[pp] = v;
            }
            if (zero) errbits |= 32u;
            stage_generic(0, off0);
            flush(std::integral_constant<int, 7>{}, 0, off0);
        }
    }

    stamp(11 + 4 * wid);
    // ---- flags: one atomic per warp and kind
    if (__any_sync(0xffffffffu, errbits != 0)) {
#pragma unroll
        for (int b = 0; b < 7; ++b) {
            const unsigned any = __ballot_sync(0xffffffffu, (errbits >> b) & 1u);
            if (lane == 0 && any) {
                int* dst = b == 0 ? &P.flags->err_dry_neighbour : b == 1 ? &P.flags->nan_adv : b == 2 ? &P.flags->nan_kh
                         : b == 3 ? &P.flags->nan_kvml : b == 4 ? &P.flags->nan_kvdeep : b == 5 ? &P.flags->zero_dropped
                                                                                                : &P.flags->nan_rho;
                atomicOr(dst, 1);
            }
        }
    }
}

// Completion record.  Launched behind k_fused_v4 on the same stream, so every tile has finished and every flag is in
// place: copy the flag block and the five nnz into the host-mapped record of this launch, re-zero the device block
// for the next launch, then store the launch's serial number — the word the host polls (transport.cu, otmb_wait_v4)
// instead of issuing a copy and a stream synchronise per build.
__global__ void __launch_bounds__(32) k_publish(DevFlags* __restrict__ flags, otmb_ctx::HostDone* __restrict__ rec, u64 serial,
                                                u64* __restrict__ totals_out) {
    constexpr int NI = (int)(sizeof(DevFlags) / sizeof(int));
    static_assert(NI <= 32, "one warp copies the flag block");
    // launched with programmatic stream serialisation: it may be scheduled while the assembly kernel still runs (its
    // launch latency is hidden) and waits HERE until that kernel has completed and its writes are visible
    cudaGridDependencySynchronize();
    int* src = reinterpret_cast<int*>(flags);
    volatile int* dst = reinterpret_cast<volatile int*>(&rec->snap);
    if (totals_out && threadIdx.x < 5) totals_out[threadIdx.x] = flags->nnz[threadIdx.x];   // (read before the block is zeroed:
    __syncwarp();                                                                          //  the ints below alias nnz)
    if (threadIdx.x < NI) {
        dst[threadIdx.x] = src[threadIdx.x];
        src[threadIdx.x] = 0;
    }
    __threadfence_system();
    __syncwarp();
    if (threadIdx.x == 0) *reinterpret_cast<volatile u64*>(&rec->seq) = serial;
}

FastDiv make_fastdiv(unsigned d) {
    unsigned s = 0;
    while ((1ull << s) < d) ++s;
    FastDiv f;
    f.shift = 32 + s;
    f.mul = ((1ull << f.shift) + d - 1) / d;
    return f;
}

template <bool RHO3D, bool UP, int TILE, int MINB, int KHB = 2, bool VAH = true, bool VC0 = true, bool TLINE = false, bool XSM = true>
int launch_v4(otmb_ctx* c, V4Params& P) {
    const int ntiles = (int)(((i64)P.ncols + TILE - 1) / TILE);
    P.ntiles = ntiles;
    // look-back descriptors: zeroed when the array grows and when the 20-bit epoch wraps, otherwise left as they are
    // (a descriptor of an earlier launch carries another epoch and reads as "not published")
    const size_t ts_bytes = (size_t)ntiles * 8 * sizeof(u64);
    c->v4_serial++;
    P.epoch = (unsigned)(c->v4_serial % ST_EPOCHS);
    if (ts_bytes > c->tile_state.cap || !c->tile_state.p) c->ts_zeroed = 0;
    CU_TRY(c, c->tile_state.ensure(ts_bytes));
    if (P.epoch == 0) c->ts_zeroed = 0;
    if (c->ts_zeroed < ts_bytes) {
        CU_TRY(c, cudaMemsetAsync(c->tile_state.p, 0, c->tile_state.cap, c->stream));
        c->ts_zeroed = c->tile_state.cap;
    }
    P.tile_state = c->tile_state.as<u64>();
    DevBuf tline;
    if (TLINE) {
        CU_TRY(c, tline.ensure((size_t)ntiles * 64 * 8));
        CU_TRY(c, cudaMemsetAsync(tline.p, 0, (size_t)ntiles * 64 * 8, c->stream));
    }
    P.timeline = tline.as<long long>();
    const size_t smem = sizeof(Smem<TILE>);
    auto kern = k_fused_v4<RHO3D, UP, TILE, MINB, KHB, VAH, VC0, TLINE, XSM>;
    // function attributes once per instantiation and device
    static std::atomic<unsigned long long> configured{0};
    if (!(configured.load(std::memory_order_acquire) >> (c->device & 63) & 1ull)) {
        CU_TRY(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        // Shared memory is paid for in L1 on this kernel (3 - 4 % per 32 KB, profiles/README.md): ask for the smallest
        // carve-out that still holds MINB blocks (sm_100: 0/8/16/32/64/100/132/164/196/228 KB; the driver rounds a
        // percentage UP to the next of these, so the request is rounded down).
        static const int kb[] = {0, 8, 16, 32, 64, 100, 132, 164, 196, 228};
        const size_t need = (size_t)MINB * (smem + 1024);
        int pct = 100;
        for (int q = 0; q < 10; ++q)
            if ((size_t)kb[q] * 1024 >= need) {
                pct = kb[q] * 100 / 228;
                break;
            }
#ifdef OTMB_AB
        if (const char* e = getenv("OTMB_V4_CARVEOUT")) pct = atoi(e);
#endif
        CU_TRY(c, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
        configured.fetch_or(1ull << (c->device & 63), std::memory_order_release);
    }
    kern<<<ntiles, TILE + 32, smem, c->stream>>>(P);
    LAUNCHED(c);
    CU_TRY(c, cudaGetLastError());
#ifdef OTMB_AB
    if (TLINE) {   // raw dump: ntiles x 64 int64 (profiles/timeline.py)
        std::vector<long long> h((size_t)ntiles * 64);
        CU_TRY(c, cudaMemcpyAsync(h.data(), tline.p, h.size() * 8, cudaMemcpyDeviceToHost, c->stream));
        CU_TRY(c, cudaStreamSynchronize(c->stream));
        if (FILE* f = fopen(getenv("OTMB_V4_TIMELINE"), "wb")) {
            fwrite(h.data(), 8, h.size(), f);
            fclose(f);
        }
        tline.release();
    }
#endif
    return OTMB_OK;
}

}  // namespace

int otmb_v4_publish(otmb_ctx* c, u64* totals_out) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1);
    cfg.blockDim = dim3(32);
    cfg.stream = c->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CU_TRY(c, cudaLaunchKernelEx(&cfg, k_publish, c->flags.as<DevFlags>(), c->d_done + (c->v4_serial % otmb_ctx::DONE_RING),
                                 (u64)c->v4_serial, totals_out));
    LAUNCHED(c);
    CU_TRY(c, cudaGetLastError());
    return OTMB_OK;
}

int otmb_fused_v4_build(otmb_ctx* c, const otmb_tm_params* prm, int build, i64 col0, i64 ncols, int chain) {
    if (ncols < 0) ncols = c->ncols - col0;
    V4Params P;
    P.start = chain == 2 ? c->run_nnz.as<u64>() : nullptr;
    P.run_out = chain ? c->run_nnz.as<u64>() : nullptr;
    P.g = GridDims{(int)c->nx, (int)c->ny, (int)c->nz, c->topo, (int)c->P, (int)c->M};
    P.divP = make_fastdiv((unsigned)c->P);
    P.divNx = make_fastdiv((unsigned)c->nx);
    // 3-D arrays: window-biased pointers, indexed by the global linear cell index (common.cuh)
    P.v3D = c->win<double>(c->v3D);
    P.thk = c->win<double>(c->thk);
    P.area2D = c->area2D.as<double>();
    P.zt = c->zt.as<double>();
    P.edge = c->edge.as<double>();
    P.dnbr = c->dnbr.as<double>();
    P.mlotst = c->mlotst.as<double>();
    P.rho3d = c->have_rho3d ? c->win<double>(c->rho3d) : nullptr;
    P.pe = c->win<double>(c->phi[OTMB_FACE_EAST]);
    P.pw = c->win<double>(c->phi[OTMB_FACE_WEST]);
    P.pn = c->win<double>(c->phi[OTMB_FACE_NORTH]);
    P.ps = c->win<double>(c->phi[OTMB_FACE_SOUTH]);
    P.pt = c->win<double>(c->phi[OTMB_FACE_TOP]);
    P.pb = c->win<double>(c->phi[OTMB_FACE_BOTTOM]);
    P.phi_nb[cT] = P.pb;
    P.phi_nb[cS] = P.pn;
    P.phi_nb[cW] = P.pe;
    P.phi_nb[cC] = P.pe;
    P.phi_nb[cE] = P.pw;
    P.phi_nb[cN] = P.ps;
    P.phi_nb[cB] = P.pt;
    P.phi_nb[7] = P.pn;
    P.rank3d = c->win<int>(c->rank3d);
    P.lwet = c->lwet.as<int>() + c->h_up + col0;   // the owned cells (all wet cells when unsharded), from column col0
    P.kH = prm->kH;
    P.kVML = prm->kVML;
    P.kVdeep = prm->kVdeep;
    P.rho = prm->rho;
    P.upwind = prm->upwind;
    P.base = prm->index_base;
    P.build = build;
    P.prefetch = 1;
#ifdef OTMB_AB
    if (getenv("OTMB_V4_NOPREFETCH")) P.prefetch = 0;
#endif
    P.Lsafe = (int)c->L_own0;
    P.ksafe = (int)(c->L_own0 / c->P);
    P.psafe = (int)(c->L_own0 % c->P);
    P.Lmax = (int)(c->L_win1 - 1);
    P.w0 = (int)(c->w0 + col0);
    P.ncols = (int)ncols;
    // one wave of tiles ahead (two blocks per SM): measured 0.3765 vs 0.3815 ms per launch on C2, 148 / 296 tiles alike,
    // 600 and more no gain (profiles/README.md)
    P.lwet_ahead = c->sm_count * 2 * 352;
#ifdef OTMB_AB
    if (const char* e = getenv("OTMB_V4_LWET_AHEAD")) P.lwet_ahead = atoi(e) * 352;   // in tiles
#endif
    P.flags = c->flags.as<DevFlags>();
    const int cap_per_col[5] = {7, 7, 5, 3, 3};
    for (int m = 0; m < 5; ++m) {
        P.colptr[m] = nullptr;
        P.rowval[m] = nullptr;
        P.nzval[m] = nullptr;
        if (!(build >> m & 1)) continue;
        const size_t cap = (size_t)c->ncols * cap_per_col[m] + 8;
        CU_TRY(c, c->colptr[m].ensure((size_t)(c->ncols + 1) * 8));
        CU_TRY(c, c->rowval[m].ensure(cap * 8));
        CU_TRY(c, c->nzval[m].ensure(cap * 8));
        P.colptr[m] = c->colptr[m].as<i64>() + col0;
        P.rowval[m] = c->rowval[m].as<i64>();
        P.nzval[m] = c->nzval[m].as<double>();
    }
    // TILE column threads + one scan warp per block
    const bool up = prm->upwind != 0;
    if (c->have_rho3d) return up ? launch_v4<true, true, 352, 2>(c, P) : launch_v4<true, false, 352, 2>(c, P);
    if (!up) return launch_v4<false, false, 352, 2>(c, P);
#ifdef OTMB_AB   // measurement builds only (nvcc -DOTMB_AB): launch geometries / schedules for A/B runs, per-tile phase stamps
    if (getenv("OTMB_V4_TIMELINE")) return launch_v4<false, true, 352, 2, 2, true, true, true>(c, P);
    switch (getenv("OTMB_V4_VARIANT") ? atoi(getenv("OTMB_V4_VARIANT")) : 0) {
        case 1: return launch_v4<false, true, 352, 2, 2, true, true, false, false>(c, P);   // Tadv re-loads the neighbours' fluxes
        case 2: return launch_v4<false, true, 416, 2>(c, P);   // 72 registers
        case 3: return launch_v4<false, true, 352, 2, 1>(c, P);   // TκH loads direction by direction
        case 4: return launch_v4<false, true, 352, 2, 2, false, false>(c, P);   // vertical inputs and own volume loaded where they are used
        case 5: return launch_v4<false, true, 320, 2>(c, P);   // 88 registers, 20 + 2 warps per SM
        case 6: return launch_v4<false, true, 288, 2>(c, P);   // 96 registers, 18 + 2 warps per SM
        default: break;
    }
#endif
    return launch_v4<false, true, 352, 2>(c, P);   // measured best on C2 (profiles/README.md)
}
