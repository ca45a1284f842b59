// fused_v3.cu — OTMB_PATH_FUSED: single-pass direct CSC assembly, one thread per WET cell,
// scheduled for occupancy.
//
// Same mathematics and ordering rules as fused.cu (read its header first: gather form, emit
// order, generic branch for coincident neighbours).  What changed against fused_v2.cu, driven by
// the ncu profiles under profiles/ (v2f: 128 registers + 77 KB of staging per 256-cell tile ->
// 16 warps per SM, IPC 0.32, 2 500 instructions per cell, DRAM at 24 % with traffic equal to the
// algorithmic bytes — a latency/issue-bound kernel, not a bandwidth-bound one):
//   * neighbours are resolved through `rank3d` (Int32 wet rank per grid cell, -1 = dry — the
//     reference's own Lwet3D, /root/reference/src/matrixbuilding.jl:18-20) with one 4-byte load
//     per neighbour instead of mask word + prefix + popcount;
//   * the row order inside a column is a function of the cell's CLASS only (regular, west seam,
//     east seam, right half of the fold row), not of rank comparisons: four fixed emission
//     sequences, entries are appended in order with a running position (no per-entry popcount);
//   * staging is warp-private and reused matrix by matrix (7 entries x 32 columns x 12 B per
//     warp instead of 25 x 256 x 12 B per tile): 21 KB per tile, no block barrier between the
//     operators, each warp flushes its own contiguous slice with 16-byte stores;
//   * values are computed one operator at a time directly before they are staged, so only the T
//     accumulators live across operators.
//
// Launch geometry: one tile of TILE consecutive wet cells per block, tiles in block-index order
// (the decoupled look-back only waits on lower-numbered tiles, which are resident or finished).
// A launch covers the wet ranks [w0, w0 + ncols): the whole matrix on one GPU, or the columns of
// one k-slab when a matrix is sharded across GPUs (rows are global wet ranks either way).
#include <cstdlib>

#include "common.cuh"

namespace {

constexpr u64 ST_AGG = 1ull << 62, ST_PRE = 2ull << 62, ST_MASK = (1ull << 62) - 1;

enum { cT = 0, cS = 1, cW = 2, cC = 3, cE = 4, cN = 5, cB = 6 };
enum { sW = 0, sE = 1, sS = 2, sN = 3, sB = 4, sT = 5 };  // emit-order slots W,E,S,N,B,T
constexpr unsigned bT = 1u << cT, bS = 1u << cS, bW = 1u << cW, bC = 1u << cC, bE = 1u << cE, bN = 1u << cN, bB = 1u << cB;
constexpr unsigned HMASK = bS | bW | bE | bN;
constexpr unsigned VMASK = bT | bB;
constexpr int WDATA = 7 * 32;       // staged entries per warp: 7 per column
constexpr int WCAP = WDATA + 32;    // + one dump slot per lane for absent entries (branch-free staging)

struct FastDiv {
    u64 mul;
    unsigned shift;
};
__device__ __forceinline__ unsigned fdiv(unsigned n, FastDiv f) { return (unsigned)((n * f.mul) >> f.shift); }

struct V3Params {
    GridDims g;
    FastDiv divP, divNx;
    const double *v3D, *thk, *area2D, *zt, *edge, *dnbr, *mlotst, *rho3d;
    const double *pe, *pw, *pn, *ps, *pt, *pb;
    const int* rank3d;   // (M) global wet rank, -1 = dry
    const int* lwet;     // (ncols) linear index of the launch's wet cells
    double kH, kVML, kVdeep, rho;
    int upwind, base, build;
    int ntiles;
    int w0;              // global wet rank of the launch's first column
    int ncols;
    i64* colptr[5];
    i64* rowval[5];
    double* nzval[5];
    DevFlags* flags;
    u64* tile_state;
};

__device__ __forceinline__ u64 ld_vol(const u64* p) { return *reinterpret_cast<const volatile u64*>(p); }
__device__ __forceinline__ void st_vol(u64* p, u64 v) { *reinterpret_cast<volatile u64*>(p) = v; }
__device__ __forceinline__ u64 warp_sum64(u64 v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}
__device__ __forceinline__ double upflux(double x, bool take_max, bool up) {
    return up ? (take_max ? jl_max(x, 0.0) : jl_min(x, 0.0)) : x / 2;
}
__device__ __forceinline__ bool nz(double f) { return f > 0 || f < 0; }

// ---------------------------------------------------------------------------------------
// generic branch (coincident neighbours): everything recomputed out of line, sparse()'s
// sort + in-order combine reproduced literally.  Rare (a handful of columns per level).
// ---------------------------------------------------------------------------------------
struct Ent {
    int row;
    i64 key;
    double val;
};
struct GenOut {
    int cnt[5];
    int rows[5][8];
    double vals[5][8];
};
__device__ void ent_sort(Ent* e, int n) {
    for (int a = 1; a < n; ++a) {
        Ent x = e[a];
        int b = a - 1;
        while (b >= 0 && (e[b].row > x.row || (e[b].row == x.row && e[b].key > x.key))) {
            e[b + 1] = e[b];
            --b;
        }
        e[b + 1] = x;
    }
}
__device__ int ent_combine(const Ent* e, int n, int* rows, double* vals) {
    int m = 0;
    for (int a = 0; a < n; ++a) {
        if (m > 0 && rows[m - 1] == e[a].row)
            vals[m - 1] = vals[m - 1] + e[a].val;
        else {
            rows[m] = e[a].row;
            vals[m] = e[a].val;
            ++m;
        }
    }
    return m;
}
__device__ __noinline__ int distinct_rows(unsigned m, const int* r) {
    int n = 0;
    for (int c = 0; c < 7; ++c) {
        if (!(m >> c & 1)) continue;
        bool dup = false;
        for (int d = 0; d < c; ++d)
            if ((m >> d & 1) && r[d] == r[c]) dup = true;
        n += !dup;
    }
    return n;
}

__device__ __noinline__ void generic_full(const V3Params& P, int L, int k, const int* Lc, const int* r, unsigned wetm,
                                          bool fold, unsigned act_adv, const double* pmag, unsigned act_ml, GenOut* out,
                                          unsigned* errbits) {
    const GridDims g = P.g;
    const int PP = g.P, p2 = L - k * PP;
    const int emit_slot[7] = {sB, sN, sE, -1, sW, fold ? sN : sS, sT};
    const int own_slot[7] = {sT, sS, sW, -1, sE, sN, sB};
    const int rC = r[cC];
    const double vC = __ldg(P.v3D + L);
    const double rhoC = P.rho3d ? __ldg(P.rho3d + L) : P.rho;
    Ent e[16];
    int n = 0;
    for (int q = 0; q < 5; ++q) out->cnt[q] = 0;
    if (P.build & 2) {
        for (int c = 0; c < 7; ++c)
            if (c != cC && (act_adv >> c & 1)) {
                const double rhoi = P.rho3d ? __ldg(P.rho3d + Lc[c]) : P.rho;
                const double rb = (rhoi + rhoC) / 2;
                const double mi = rb * __ldg(P.v3D + Lc[c]), mj = rb * vC;
                const double a = -pmag[c] / mi, d = pmag[c] / mj;
                if (isnan(a) || isnan(d)) *errbits |= 2u;
                const i64 kb = (i64)r[c] * 16 + emit_slot[c] * 2;
                e[n++] = Ent{r[c], kb, a};
                e[n++] = Ent{rC, kb + 1, d};
            }
        ent_sort(e, n);
        out->cnt[1] = ent_combine(e, n, out->rows[1], out->vals[1]);
    }
    if (P.build & 4) {
        n = 0;
        const double thC = __ldg(P.thk + L);
        for (int c = 0; c < 7; ++c)
            if ((HMASK >> c & 1) && (wetm >> c & 1)) {
                const int own = c == cW ? OTMB_DIR_WEST : c == cE ? OTMB_DIR_EAST : c == cS ? OTMB_DIR_SOUTH : OTMB_DIR_NORTH;
                const int opp = c == cW ? OTMB_DIR_EAST : c == cE ? OTMB_DIR_WEST : c == cS ? OTMB_DIR_NORTH
                                                                                  : (fold ? OTMB_DIR_NORTH : OTMB_DIR_SOUTH);
                const int q2 = Lc[c] - k * PP;
                const double a = jl_min(thC * __ldg(P.edge + own * PP + p2), __ldg(P.thk + Lc[c]) * __ldg(P.edge + opp * PP + q2));
                const double ka = P.kH * a;
                const double ts = ka / (__ldg(P.dnbr + own * PP + p2) * vC);
                const double tn = ka / (__ldg(P.dnbr + opp * PP + q2) * __ldg(P.v3D + Lc[c]));
                if (isnan(ts) || isnan(tn)) *errbits |= 4u;
                e[n++] = Ent{rC, (i64)rC * 16 + own_slot[c] * 2, ts};
                e[n++] = Ent{r[c], (i64)r[c] * 16 + emit_slot[c] * 2 + 1, -tn};
            }
        ent_sort(e, n);
        out->cnt[2] = ent_combine(e, n, out->rows[2], out->vals[2]);
    }
    for (int op = 3; op <= 4; ++op) {
        if (!(P.build >> op & 1)) continue;
        n = 0;
        const double area = __ldg(P.area2D + p2), ztC = __ldg(P.zt + k);
        const double kap = op == 3 ? P.kVML : P.kVdeep;
        for (int c = 0; c < 7; ++c) {
            const bool on = op == 3 ? (act_ml >> c & 1) : ((VMASK >> c & 1) && (wetm >> c & 1));
            if (!on) continue;
            const int kc = c == cT ? k - 1 : k + 1;
            const double d = fabs(ztC - __ldg(P.zt + kc));
            const double ka = kap * area;
            const double ts = ka / (d * vC), tn = ka / (d * __ldg(P.v3D + Lc[c]));
            if (isnan(ts) || isnan(tn)) *errbits |= (op == 3 ? 8u : 16u);
            e[n++] = Ent{rC, (i64)rC * 16 + own_slot[c] * 2, ts};
            e[n++] = Ent{r[c], (i64)r[c] * 16 + emit_slot[c] * 2 + 1, -tn};
        }
        ent_sort(e, n);
        out->cnt[op] = ent_combine(e, n, out->rows[op], out->vals[op]);
    }
    if (P.build & 1) {  // union merge; exact zeros are KEPT here and flagged (the compaction pass drops them)
        int idx[4] = {0, 0, 0, 0};
        int m = 0;
        while (true) {
            int row = 0x7fffffff;
            for (int q = 0; q < 4; ++q)
                if (idx[q] < out->cnt[q + 1] && out->rows[q + 1][idx[q]] < row) row = out->rows[q + 1][idx[q]];
            if (row == 0x7fffffff) break;
            double x = 0.0;
            for (int q = 0; q < 4; ++q) {
                double v = 0.0;
                if (idx[q] < out->cnt[q + 1] && out->rows[q + 1][idx[q]] == row) {
                    v = out->vals[q + 1][idx[q]];
                    ++idx[q];
                }
                x = x + v;
            }
            if (x == 0.0) *errbits |= 32u;
            out->rows[0][m] = row;
            out->vals[0][m] = x;
            ++m;
        }
        out->cnt[0] = m;
    }
}

// ---------------------------------------------------------------------------------------
// warp-private staging
// ---------------------------------------------------------------------------------------
// Branch-free append: an absent entry is written to the lane's private dump slot instead.
struct Stager {
    int* srow;
    double* sval;
    int pos, dump;
    __device__ __forceinline__ void put(bool on, int row, double v) {
        const int p = on ? pos : dump;
        srow[p] = row;
        sval[p] = v;
        pos += on ? 1 : 0;
    }
};

// Appends the present entries (mask m) of one column in ascending-row order.  cls: 0 regular
// (T S W C E N B; the fold's left half has N after E as well), 1 west seam (W wraps to the end of
// the row: T S C E W N B), 2 east seam (E wraps to the start: T S E W C N B), 3 right half of
// the fold row (N mirrors to the left of W: T S N W C E B).  r holds row + index_base.
__device__ __forceinline__ void stage_regular(Stager& st, unsigned m, const int (&r)[7], const double (&v)[7]) {
    st.put(m & bT, r[cT], v[cT]);
    st.put(m & bS, r[cS], v[cS]);
    st.put(m & bW, r[cW], v[cW]);
    st.put(m & bC, r[cC], v[cC]);
    st.put(m & bE, r[cE], v[cE]);
    st.put(m & bN, r[cN], v[cN]);     // regular north, or fold north with a larger i than east
    st.put(m & bB, r[cB], v[cB]);
}
__device__ __noinline__ int stage_irregular(int* srow, double* sval, int pos, unsigned m, int cls, bool fold, const int* r,
                                            const double* v) {
    auto put = [&](bool on, int c) {
        if (on) {
            srow[pos] = r[c];
            sval[pos] = v[c];
            ++pos;
        }
    };
    put(m & bT, cT);
    put(m & bS, cS);
    const bool nf = fold && (m & bN);
    if (cls == 1) {
        put(m & bC, cC);
        put(m & bE, cE);
        put(m & bW, cW);
        put(nf, cN);
    } else if (cls == 2) {
        put(m & bE, cE);
        put(m & bW, cW);
        put(m & bC, cC);
        put(nf, cN);
    } else {
        put(nf, cN);
        put(m & bW, cW);
        put(m & bC, cC);
        put(m & bE, cE);
    }
    put(!fold && (m & bN), cN);
    put(m & bB, cB);
    return pos;
}

// ---------------------------------------------------------------------------------------
template <bool RHO3D, int TILE, int MINB>
// __grid_constant__: the generic branch takes the address of P; without it every thread would copy
// the whole parameter block to local memory at kernel entry
__global__ void __launch_bounds__(TILE, MINB) k_fused_v3(const __grid_constant__ V3Params P) {
    constexpr int NW = TILE / 32;
    __shared__ __align__(16) double s_val[NW][WCAP];
    __shared__ __align__(16) int s_row[NW][WCAP];
    __shared__ u64 s_warp[NW];
    __shared__ u64 s_excl[5];

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int tile = blockIdx.x;
    const GridDims g = P.g;
    const int PP = g.P;
    const int w = tile * TILE + tid;       // column of this launch
    const bool valid = w < P.ncols;
    const int rC = P.w0 + w;               // global wet rank = row/column index
    const bool up = P.upwind != 0;

    // ================= phase 0: pattern =================
    int L = 0, k = 0, p2 = 0;
    int Lc[7], r[7];
    unsigned wetm = 0, act = 0, mlm = 0;
    int cls = 0;
    bool fold = false, generic = false;
    unsigned m_T = 0, m_adv = 0, m_kh = 0, m_ml = 0, m_dp = 0;
    unsigned errbits = 0;  // 1 dry nbr, 2 nan adv, 4 nan kh, 8 nan ml, 16 nan deep, 32 zero dropped, 64 nan rho
    u64 packed = 0;        // 5 counts, 12 bits each
#pragma unroll
    for (int c = 0; c < 7; ++c) {
        Lc[c] = 0;
        r[c] = -1;
    }
    if (valid) {
        L = __ldg(P.lwet + w);
        k = (int)fdiv((unsigned)L, P.divP);
        p2 = L - k * PP;
        const int j = (int)fdiv((unsigned)p2, P.divNx);
        const int i = p2 - j * g.nx;
        fold = (j == g.ny - 1) && (g.topo == OTMB_TOPO_TRIPOLAR);
        const bool hasT = k > 0, hasB = k < g.nz - 1, hasS = j > 0, hasN = (j < g.ny - 1) || fold;
        const bool seamW = i == 0, seamE = i == g.nx - 1;
        cls = seamW ? 1 : seamE ? 2 : (fold && (g.nx - 1 - i < i)) ? 3 : 0;
        Lc[cC] = L;
        Lc[cT] = hasT ? L - PP : L;
        Lc[cB] = hasB ? L + PP : L;
        Lc[cS] = hasS ? L - g.nx : L;
        Lc[cW] = seamW ? L + (g.nx - 1) : L - 1;
        Lc[cE] = seamE ? L - (g.nx - 1) : L + 1;
        Lc[cN] = (j < g.ny - 1) ? L + g.nx : (fold ? L + (g.nx - 1 - 2 * i) : L);
        // ---- loads: neighbour ranks, the six face fluxes the neighbours carry, mixed-layer inputs
        const int qT = __ldg(P.rank3d + Lc[cT]), qS = __ldg(P.rank3d + Lc[cS]), qW = __ldg(P.rank3d + Lc[cW]),
                  qE = __ldg(P.rank3d + Lc[cE]), qN = __ldg(P.rank3d + Lc[cN]), qB = __ldg(P.rank3d + Lc[cB]);
        const double xT = __ldg(P.pb + Lc[cT]);                       // emitter above: its Bottom slot, max
        const double xS = __ldg(P.pn + Lc[cS]);                       // its North slot, min
        const double xW = __ldg(P.pe + Lc[cW]);                       // its East slot, min
        const double xE = __ldg(P.pw + Lc[cE]);                       // its West slot, max
        const double xN = __ldg((fold ? P.pn : P.ps) + Lc[cN]);       // its South slot (max), or North on the fold (min)
        const double xB = __ldg(P.pt + Lc[cB]);                       // emitter below: its Top slot, min
        const double ml = __ldg(P.mlotst + p2);
        const double z0 = __ldg(P.zt + k), zT = __ldg(P.zt + (hasT ? k - 1 : k)), zB = __ldg(P.zt + (hasB ? k + 1 : k));
        r[cC] = rC;
        r[cT] = hasT ? qT : -1;
        r[cS] = hasS ? qS : -1;
        r[cW] = qW;
        r[cE] = qE;
        r[cN] = hasN ? qN : -1;
        r[cB] = hasB ? qB : -1;
#pragma unroll
        for (int c = 0; c < 7; ++c)
            if (c != cC && r[c] >= 0) wetm |= 1u << c;
        // face flux each neighbour carries through the face it shares with this cell (the value the
        // reference reads at the neighbour, :244-295): only the sign pattern is needed here
        if (P.build & 2) {
            if ((wetm & bT) && nz(upflux(xT, true, up))) act |= bT;
            if ((wetm & bS) && nz(upflux(xS, false, up))) act |= bS;
            if ((wetm & bW) && nz(upflux(xW, false, up))) act |= bW;
            if ((wetm & bE) && nz(upflux(xE, true, up))) act |= bE;
            if ((wetm & bN) && nz(upflux(xN, !fold, up))) act |= bN;
            if ((wetm & bB) && nz(upflux(xB, false, up))) act |= bB;
            // own faces that point at a dry or absent cell: the reference would push `missing` (:247-250)
            const unsigned dry = ~wetm;
            bool bad = false;
            if (dry & bW) bad |= nz(upflux(__ldg(P.pw + L), true, up));
            if (dry & bE) bad |= nz(upflux(__ldg(P.pe + L), false, up));
            if (dry & bS) bad |= nz(upflux(__ldg(P.ps + L), true, up));
            if (dry & bN) bad |= nz(upflux(__ldg(P.pn + L), false, up));
            if (dry & bB) bad |= nz(upflux(__ldg(P.pb + L), true, up));
            if ((dry & bT) && hasT) bad |= nz(upflux(__ldg(P.pt + L), false, up));
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
        if (cls != 0 || fold) {
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
    }

    // ================= tile scan + decoupled look-back =================
    u64 incl = packed;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u64 o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) s_warp[wid] = incl;
    const u64 lane_excl = incl - packed;                       // in-warp exclusive offsets of this column
    const u64 warp_tot = __shfl_sync(0xffffffffu, incl, 31);   // entries of this warp, per matrix
    __syncthreads();
    u64 wbase = 0, total = 0;                                  // in-tile offset of this warp, tile totals
#pragma unroll
    for (int q = 0; q < NW; ++q) {
        const u64 sw = s_warp[q];
        if (q < wid) wbase += sw;
        total += sw;
    }
    if (wid < (NW < 5 ? NW : 5)) {   // one warp per counter; the other warps go straight to the values
        for (int m = wid; m < 5; m += NW) {
            const u64 agg = (total >> (12 * m)) & 0xfffull;
            if (lane == 0) st_vol(P.tile_state + (size_t)tile * 8 + m, (tile == 0 ? ST_PRE : ST_AGG) | agg);
            u64 excl = 0;
            if (tile > 0) {
                int look = tile - 1;
                while (true) {
                    const int t = look - lane;
                    u64 wv = ST_PRE;
                    if (t >= 0) {
                        do {
                            wv = ld_vol(P.tile_state + (size_t)t * 8 + m);
                        } while ((wv >> 62) == 0);
                    }
                    const u64 val = wv & ST_MASK;
                    const unsigned pm = __ballot_sync(0xffffffffu, (wv >> 62) == 2);
                    if (pm) {
                        const int first = __ffs(pm) - 1;
                        excl += warp_sum64(lane <= first ? val : 0ull);
                        break;
                    }
                    excl += warp_sum64(val);
                    look -= 32;
                }
                if (lane == 0) st_vol(P.tile_state + (size_t)tile * 8 + m, ST_PRE | (excl + agg));
            }
            if (lane == 0) {
                s_excl[m] = excl;
                if (tile == P.ntiles - 1) {
                    const u64 nnz = excl + agg;
                    P.flags->nnz[m] = nnz;
                    if (P.build >> m & 1) P.colptr[m][P.ncols] = (i64)nnz + P.base;
                }
            }
        }
    }

    // ================= phase 1a: Tadv values (no dependence on the look-back) =================
    double val[7];   // values of the operator being staged, by candidate
    double Tv[7];    // running T = ((Tadv + TκH) + TκVML) + TκVdeep, by candidate
#pragma unroll
    for (int c = 0; c < 7; ++c) val[c] = Tv[c] = 0.0;
    GenOut gen;      // local memory, only touched by generic columns
    double vn[7];
    if (valid) {
#pragma unroll
        for (int c = 0; c < 7; ++c) vn[c] = __ldg(P.v3D + Lc[c]);
    } else {
#pragma unroll
        for (int c = 0; c < 7; ++c) vn[c] = 0.0;
    }
    const double vC = vn[cC];

    if (valid && generic) {
        double g_p[7];
        int g_Lc[7], g_r[7];
        unsigned g_err = 0;
        const double xs[7] = {__ldg(P.pb + Lc[cT]), __ldg(P.pn + Lc[cS]), __ldg(P.pe + Lc[cW]), 0.0, __ldg(P.pw + Lc[cE]),
                              __ldg((fold ? P.pn : P.ps) + Lc[cN]), __ldg(P.pt + Lc[cB])};
        const bool mx[7] = {true, false, false, false, true, !fold, false};
#pragma unroll
        for (int c = 0; c < 7; ++c) {
            const double f = upflux(xs[c], mx[c], up);
            g_p[c] = (c == cC) ? 0.0 : (mx[c] ? f : -f);
            g_Lc[c] = Lc[c];
            g_r[c] = r[c];
        }
        generic_full(P, L, k, g_Lc, g_r, wetm, fold, act, g_p, mlm, &gen, &g_err);
        errbits |= g_err;
        atomicAdd(&P.flags->generic_columns, 1);
    } else if (valid) {
        if (m_adv) {
            // ---- Tadv (:193-204): off-diagonal (𝑖, 𝑗) = -p/m𝑖 seen from the emitter 𝑖; diagonal = Σ p/m𝑗
            const double rhoC = RHO3D ? __ldg(P.rho3d + L) : P.rho;
            if (RHO3D && isnan(rhoC)) errbits |= 64u;
            const double xs[7] = {__ldg(P.pb + Lc[cT]), __ldg(P.pn + Lc[cS]), __ldg(P.pe + Lc[cW]), 0.0, __ldg(P.pw + Lc[cE]),
                                  __ldg((fold ? P.pn : P.ps) + Lc[cN]), __ldg(P.pt + Lc[cB])};
            const bool mx[7] = {true, false, false, false, true, !fold, false};
            double dd[7];
            bool bad = false;
#pragma unroll
            for (int c = 0; c < 7; ++c) {
                dd[c] = 0.0;
                if (c == cC || !(act >> c & 1)) continue;
                const double f = upflux(xs[c], mx[c], up);
                const double p = mx[c] ? f : -f;     // pushed magnitude: ϕ for W,S,B slots of the emitter, -ϕ for E,N,T
                const double rb = ((RHO3D ? __ldg(P.rho3d + Lc[c]) : P.rho) + rhoC) / 2;
                const double a = -p / (rb * vn[c]);
                dd[c] = p / (rb * vC);
                bad |= isnan(a) || isnan(dd[c]);
                val[c] = a;
            }
            if (bad) errbits |= 2u;
            // diagonal: emitter contributions in ascending wet rank; sparse! keeps the first value and
            // adds the later ones.  Order: T, S, the row group {W, E, N-on-fold} by class, N-regular, B.
            double dsum = 0.0;
            bool first = true;
            auto add = [&](bool on, double d) {
                if (on) {
                    dsum = first ? d : dsum + d;
                    first = false;
                }
            };
            const bool nf = fold && (act & bN);
            add(act & bT, dd[cT]);
            add(act & bS, dd[cS]);
            if (cls == 0) {
                add(act & bW, dd[cW]);
                add(act & bE, dd[cE]);
                add(nf, dd[cN]);
            } else if (cls == 3) {
                add(nf, dd[cN]);
                add(act & bW, dd[cW]);
                add(act & bE, dd[cE]);
            } else {
                add(act & bE, dd[cE]);
                add(act & bW, dd[cW]);
            }
            add(!fold && (act & bN), dd[cN]);
            add(act & bB, dd[cB]);
            val[cC] = dsum;
        } else if (RHO3D && (P.build & 2)) {
            if (isnan(__ldg(P.rho3d + L))) errbits |= 64u;
        }
    }

    __syncthreads();   // look-back results (s_excl) published

    // ================= phase 1b/2: operator by operator: values -> warp staging -> flush =================
    int* const srow = s_row[wid];
    double* const sval = s_val[wid];
    const int wcol0 = tile * TILE + wid * 32;   // first column of this warp
    const bool warp_live = wcol0 < P.ncols;

    int rb[7];       // row index as stored: global wet rank + index base
#pragma unroll
    for (int c = 0; c < 7; ++c) rb[c] = r[c] + P.base;

    auto emit_matrix = [&](const int q, const unsigned m) {
        // global offset of the warp's slice, number of entries, this column's offset in the slice
        const u64 g0 = s_excl[q] + ((wbase >> (12 * q)) & 0xfffull);
        const int n = (int)((warp_tot >> (12 * q)) & 0xfffull);
        const int off = (int)((lane_excl >> (12 * q)) & 0xfffull);
        if (valid) P.colptr[q][w] = (i64)(g0 + (u64)off) + P.base;
        __syncwarp();   // the previous matrix has been flushed out of the staging buffer
        if (valid) {
            if (cls == 0 && !generic) {
                Stager st{srow, sval, off, WDATA + lane};
                stage_regular(st, m, rb, val);
            } else if (!generic) {
                int l_r[7];
                double l_v[7];
#pragma unroll
                for (int c = 0; c < 7; ++c) {
                    l_r[c] = rb[c];
                    l_v[c] = val[c];
                }
                stage_irregular(srow, sval, off, m, cls, fold, l_r, l_v);
            } else {
                for (int a = 0; a < gen.cnt[q]; ++a) {
                    srow[off + a] = gen.rows[q][a] + P.base;
                    sval[off + a] = gen.vals[q][a];
                }
            }
        }
        __syncwarp();
        i64* __restrict__ rv = P.rowval[q] + g0;
        double* __restrict__ nv = P.nzval[q] + g0;
#pragma unroll 1
        for (int e = lane; e < n; e += 32) {
            rv[e] = (i64)(unsigned)srow[e];
            nv[e] = sval[e];
        }
    };

    if (warp_live) {
        // ---- Tadv
        if (P.build & 2) {
            emit_matrix(1, m_adv);
#pragma unroll
            for (int c = 0; c < 7; ++c) Tv[c] = val[c];
        }
        // ---- TκH (:348-415, :426-435); own slots in emit order W,E,S,N
        if (P.build & 4) {
#pragma unroll
            for (int c = 0; c < 7; ++c) val[c] = 0.0;
            if (valid && !generic && m_kh) {
                const double thC = __ldg(P.thk + L);
                double dsum = 0.0;
                bool first = true, bad = false;
                const int ord[4] = {cW, cE, cS, cN};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int c = ord[q];
                    if (!(wetm >> c & 1)) continue;
                    const int own = c == cW ? OTMB_DIR_WEST : c == cE ? OTMB_DIR_EAST : c == cS ? OTMB_DIR_SOUTH : OTMB_DIR_NORTH;
                    const int opp = c == cW ? OTMB_DIR_EAST : c == cE ? OTMB_DIR_WEST : c == cS ? OTMB_DIR_NORTH
                                                                                      : (fold ? OTMB_DIR_NORTH : OTMB_DIR_SOUTH);
                    const int q2 = Lc[c] - k * PP;
                    const double a_own = thC * __ldg(P.edge + own * PP + p2);
                    const double a_nbr = __ldg(P.thk + Lc[c]) * __ldg(P.edge + opp * PP + q2);
                    const double ka = P.kH * jl_min(a_own, a_nbr);
                    const double ts = ka / (__ldg(P.dnbr + own * PP + p2) * vC);        // row 𝑗 seen from 𝑗
                    const double tn = ka / (__ldg(P.dnbr + opp * PP + q2) * vn[c]);     // row 𝑖 seen from 𝑖
                    bad |= isnan(ts) || isnan(tn);
                    dsum = first ? ts : dsum + ts;
                    first = false;
                    val[c] = -tn;
                }
                if (bad) errbits |= 4u;
                val[cC] = dsum;
            }
            emit_matrix(2, m_kh);
#pragma unroll
            for (int c = 0; c < 7; ++c) Tv[c] = Tv[c] + val[c];
        }
        // ---- TκVML and TκVdeep (:450-477); own slots in emit order B, T; T folds TκVML before TκVdeep
        double dpT = 0.0, dpB = 0.0, dps = 0.0;
        {
#pragma unroll
            for (int c = 0; c < 7; ++c) val[c] = 0.0;
            if (valid && !generic && (m_dp | m_ml)) {
                const double area = __ldg(P.area2D + p2), ztC = __ldg(P.zt + k);
                const double ztT = __ldg(P.zt + (k > 0 ? k - 1 : k)), ztB = __ldg(P.zt + (k < g.nz - 1 ? k + 1 : k));
                double mls = 0.0;
                bool firstm = true, firstd = true, badm = false, badd = false;
                const int ord[2] = {cB, cT};
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int c = ord[q];
                    if (!(wetm >> c & 1)) continue;
                    const double d = fabs(ztC - (c == cT ? ztT : ztB));
                    const double qs = d * vC, qn = d * vn[c];
                    if (m_dp) {
                        const double ka = P.kVdeep * area;
                        const double ts = ka / qs, tn = ka / qn;
                        badd |= isnan(ts) || isnan(tn);
                        dps = firstd ? ts : dps + ts;
                        firstd = false;
                        if (c == cT) dpT = -tn; else dpB = -tn;
                    }
                    if (mlm >> c & 1) {
                        const double ka = P.kVML * area;
                        const double ts = ka / qs, tn = ka / qn;
                        badm |= isnan(ts) || isnan(tn);
                        mls = firstm ? ts : mls + ts;
                        firstm = false;
                        val[c] = -tn;
                    }
                }
                if (badm) errbits |= 8u;
                if (badd) errbits |= 16u;
                val[cC] = mls;
            }
        }
        if (P.build & 8) {
            emit_matrix(3, m_ml);
#pragma unroll
            for (int c = 0; c < 7; ++c) Tv[c] = Tv[c] + val[c];
        }
        if (P.build & 16) {
#pragma unroll
            for (int c = 0; c < 7; ++c) val[c] = 0.0;
            val[cT] = dpT;
            val[cB] = dpB;
            val[cC] = dps;
            emit_matrix(4, m_dp);
#pragma unroll
            for (int c = 0; c < 7; ++c) Tv[c] = Tv[c] + val[c];
        }
        // ---- T: union pattern; exact zeros are flagged and removed by the compaction pass
        if (P.build & 1) {
            if (valid && !generic && m_T) {
                bool zero = false;
#pragma unroll
                for (int c = 0; c < 7; ++c) zero |= (m_T >> c & 1) && (Tv[c] == 0.0);
                if (zero) errbits |= 32u;
            }
#pragma unroll
            for (int c = 0; c < 7; ++c) val[c] = Tv[c];
            emit_matrix(0, m_T);
        }
    }

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

FastDiv make_fastdiv(unsigned d) {
    unsigned s = 0;
    while ((1ull << s) < d) ++s;
    FastDiv f;
    f.shift = 32 + s;
    f.mul = ((1ull << f.shift) + d - 1) / d;
    return f;
}

template <bool RHO3D, int TILE, int MINB>
int launch_v3(otmb_ctx* c, V3Params& P) {
    const int ntiles = (int)(((i64)P.ncols + TILE - 1) / TILE);
    P.ntiles = ntiles;
    CU_TRY(c, c->tile_state.ensure((size_t)ntiles * 8 * sizeof(u64)));
    P.tile_state = c->tile_state.as<u64>();
    CU_TRY(c, cudaMemsetAsync(P.tile_state, 0, (size_t)ntiles * 8 * sizeof(u64), c->stream));
    CU_TRY(c, cudaFuncSetAttribute(k_fused_v3<RHO3D, TILE, MINB>, cudaFuncAttributePreferredSharedMemoryCarveout, 50));
    k_fused_v3<RHO3D, TILE, MINB><<<ntiles, TILE, 0, c->stream>>>(P);
    LAUNCHED(c);
    CU_TRY(c, cudaGetLastError());
    return OTMB_OK;
}

}  // namespace

int otmb_fused_v3_build(otmb_ctx* c, const otmb_tm_params* prm, int build) {
    V3Params P;
    P.g = GridDims{(int)c->nx, (int)c->ny, (int)c->nz, c->topo, (int)c->P, (int)c->M};
    P.divP = make_fastdiv((unsigned)c->P);
    P.divNx = make_fastdiv((unsigned)c->nx);
    P.v3D = c->v3D.as<double>();
    P.thk = c->thk.as<double>();
    P.area2D = c->area2D.as<double>();
    P.zt = c->zt.as<double>();
    P.edge = c->edge.as<double>();
    P.dnbr = c->dnbr.as<double>();
    P.mlotst = c->mlotst.as<double>();
    P.rho3d = c->have_rho3d ? c->rho3d.as<double>() : nullptr;
    P.pe = c->phi[OTMB_FACE_EAST].as<double>();
    P.pw = c->phi[OTMB_FACE_WEST].as<double>();
    P.pn = c->phi[OTMB_FACE_NORTH].as<double>();
    P.ps = c->phi[OTMB_FACE_SOUTH].as<double>();
    P.pt = c->phi[OTMB_FACE_TOP].as<double>();
    P.pb = c->phi[OTMB_FACE_BOTTOM].as<double>();
    P.rank3d = c->rank3d.as<int>();
    P.lwet = c->lwet.as<int>();
    P.kH = prm->kH;
    P.kVML = prm->kVML;
    P.kVdeep = prm->kVdeep;
    P.rho = prm->rho;
    P.upwind = prm->upwind;
    P.base = prm->index_base;
    P.build = build;
    P.w0 = 0;
    P.ncols = (int)c->N;
    P.flags = c->flags.as<DevFlags>();
    const int cap_per_col[5] = {7, 7, 5, 3, 3};
    for (int m = 0; m < 5; ++m) {
        P.colptr[m] = nullptr;
        P.rowval[m] = nullptr;
        P.nzval[m] = nullptr;
        if (!(build >> m & 1)) continue;
        const size_t cap = (size_t)c->N * cap_per_col[m] + 8;
        CU_TRY(c, c->colptr[m].ensure((size_t)(c->N + 1) * 8));
        CU_TRY(c, c->rowval[m].ensure(cap * 8));
        CU_TRY(c, c->nzval[m].ensure(cap * 8));
        P.colptr[m] = c->colptr[m].as<i64>();
        P.rowval[m] = c->rowval[m].as<i64>();
        P.nzval[m] = c->nzval[m].as<double>();
    }
    if (c->have_rho3d) return launch_v3<true, 256, 3>(c, P);
    return launch_v3<false, 256, 2>(c, P);
}
