// fused_v2.cu — OTMB_PATH_FUSED: single-pass direct CSC assembly, one thread per WET cell.
//
// Same mathematics and ordering rules as fused.cu (read its header first: gather form, emit
// order, generic branch for coincident neighbours); what changes is the schedule, driven by the
// ncu profiles under profiles/ (v1: 19 % of executed instructions in the look-back spin, IPC 0.2,
// 128 registers; v2a/b: L1/LSU-bound on scattered 8-byte stores and spill traffic; v2d: latency
// chain per tile, 85 % integer/control instructions):
//   * threads map to wet cells through the compacted Lwet list (no idle dry lanes; a tile of
//     TILE consecutive wet cells owns a contiguous slice of every output array);
//   * phase 0 derives the sparsity PATTERN of all five matrices from cheap data only (wet bits,
//     sign of the six face fluxes, mixed-layer test), so the tile aggregate is published a few
//     hundred cycles after the tile starts and the decoupled look-back never waits on anybody's
//     floating-point work;
//   * every load of a column is unconditional (indices clamped to a valid cell) and issued in two
//     batches (pattern inputs, value inputs): two memory round trips per column, not one per
//     direction;
//   * phase 1 computes operator by operator (Tadv, TκH, TκVML/TκVdeep), folds the entries into the
//     running T = ((Tadv + TκH) + TκVML) + TκVdeep accumulators and stages (row, value) pairs of
//     the whole tile in shared memory at their final in-tile position;
//   * phase 2 flushes the staged tile to the five CSC arrays with coalesced 16-byte stores.
//
// Row order inside a column.  Wet rank is monotone in the linear index, so the candidates always
// sort as  [top] [south] {west, self, east, north-if-on-the-fold} [north-if-regular] [bottom]:
// only the members of the braces (all in grid row j) can permute (periodic seam, tripolar fold),
// and only they can coincide.  Six integer comparisons settle the order; a coincidence sends the
// column to the generic sort-and-combine branch.
//
// The pattern of T is the union of the four patterns; sparse `+` additionally drops results that
// are exactly zero (/root/reference/src/matrixbuilding.jl:147).  Those are counted by a flag and,
// only when any occurred (e.g. κ = 0), a compaction pass (k_count_nonzero / k_copy_nonzero)
// removes them.
#include <cstdlib>

#include "common.cuh"

namespace {

constexpr u64 ST_AGG = 1ull << 62, ST_PRE = 2ull << 62, ST_MASK = (1ull << 62) - 1;

enum { cT = 0, cS = 1, cW = 2, cC = 3, cE = 4, cN = 5, cB = 6 };
enum { sW = 0, sE = 1, sS = 2, sN = 3, sB = 4, sT = 5 };  // emit-order slots W,E,S,N,B,T
constexpr unsigned bT = 1u << cT, bS = 1u << cS, bW = 1u << cW, bC = 1u << cC, bE = 1u << cE, bN = 1u << cN, bB = 1u << cB;
constexpr unsigned HMASK = bS | bW | bE | bN;
constexpr unsigned VMASK = bT | bB;
// staging capacity per column and matrix (T, Tadv, TκH, TκVML, TκVdeep) and their prefix
constexpr int CAP0 = 7, CAP1 = 7, CAP2 = 5, CAP3 = 3, CAP4 = 3, CAPSUM = 25;
__host__ __device__ constexpr int ebase(int q) {
    return q == 0 ? 0 : q == 1 ? CAP0 : q == 2 ? CAP0 + CAP1 : q == 3 ? CAP0 + CAP1 + CAP2 : CAP0 + CAP1 + CAP2 + CAP3;
}

struct FastDiv {
    u64 mul;
    unsigned shift;
};
__device__ __forceinline__ unsigned fdiv(unsigned n, FastDiv f) { return (unsigned)((n * f.mul) >> f.shift); }

struct V2Params {
    GridDims g;
    FastDiv divP, divNx;
    const double *v3D, *thk, *area2D, *zt, *edge, *dnbr, *mlotst, *rho3d;
    const double *pe, *pw, *pn, *ps, *pt, *pb;
    const u64* mask;
    const uint32_t* wpre;
    const int* lwet;
    double kH, kVML, kVdeep, rho;
    int upwind, base, build;
    int ntiles;
    i64 N;
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
// generic branch (coincident neighbours): everything recomputed in local memory, sparse()'s
// sort + in-order combine reproduced literally.  Rare (a handful of columns per level).
// ---------------------------------------------------------------------------------------
struct Ent {
    int row;
    i64 key;
    double val;
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

// stages the column's entries of all five matrices at srow/sval[sidx[q] + ...]
__device__ __noinline__ void generic_full(const V2Params& P, int L, int k, const int* Lc, const int* r, unsigned wetm,
                                          bool fold, unsigned act_adv, const double* pmag, unsigned act_ml,
                                          const int* sidx /*[5] smem index of the column's first entry*/, int* srow,
                                          double* sval, unsigned* errbits) {
    const GridDims g = P.g;
    const int PP = g.P, p2 = L - k * PP;
    const int emit_slot[7] = {sB, sN, sE, -1, sW, fold ? sN : sS, sT};
    const int own_slot[7] = {sT, sS, sW, -1, sE, sN, sB};
    const int rC = r[cC];
    const double vC = __ldg(P.v3D + L);
    const double rhoC = P.rho3d ? __ldg(P.rho3d + L) : P.rho;
    int rows[5][8];
    double vals[5][8];
    int cnt[5] = {0, 0, 0, 0, 0};
    Ent e[16];
    int n = 0;
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
        cnt[1] = ent_combine(e, n, rows[1], vals[1]);
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
        cnt[2] = ent_combine(e, n, rows[2], vals[2]);
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
        cnt[op] = ent_combine(e, n, rows[op], vals[op]);
    }
    if (P.build & 1) {  // union merge; exact zeros are KEPT here and flagged (the compaction pass drops them)
        int idx[4] = {0, 0, 0, 0};
        int m = 0;
        while (true) {
            int row = 0x7fffffff;
            for (int q = 0; q < 4; ++q)
                if (idx[q] < cnt[q + 1] && rows[q + 1][idx[q]] < row) row = rows[q + 1][idx[q]];
            if (row == 0x7fffffff) break;
            double x = 0.0;
            for (int q = 0; q < 4; ++q) {
                double v = 0.0;
                if (idx[q] < cnt[q + 1] && rows[q + 1][idx[q]] == row) {
                    v = vals[q + 1][idx[q]];
                    ++idx[q];
                }
                x = x + v;
            }
            if (x == 0.0) *errbits |= 32u;
            rows[0][m] = row;
            vals[0][m] = x;
            ++m;
        }
        cnt[0] = m;
    }
    for (int q = 0; q < 5; ++q) {
        if (!(P.build >> q & 1)) continue;
        for (int a = 0; a < cnt[q]; ++a) {
            srow[sidx[q] + a] = rows[q][a];
            sval[sidx[q] + a] = vals[q][a];
        }
    }
}

// compare-exchange of (rank, value) pairs, ascending rank
__device__ __forceinline__ void cex(int& ka, double& va, int& kb, double& vb) {
    const bool sw = kb < ka;
    const int k0 = sw ? kb : ka, k1 = sw ? ka : kb;
    const double v0 = sw ? vb : va, v1 = sw ? va : vb;
    ka = k0; kb = k1; va = v0; vb = v1;
}

// ---------------------------------------------------------------------------------------
template <bool RHO3D, int TILE, int MINB>
// __grid_constant__: the generic branch takes the address of P; without it every thread would copy
// the whole parameter block to local memory at kernel entry (measured: 1.2 GB of local stores)
__global__ void __launch_bounds__(TILE, MINB) k_fused_v2(const __grid_constant__ V2Params P) {
    constexpr int NW = TILE / 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double* sval = reinterpret_cast<double*>(smem_raw);            // CAPSUM*TILE doubles
    int* srow = reinterpret_cast<int*>(sval + CAPSUM * TILE);      // CAPSUM*TILE ints
    __shared__ u64 s_warp[NW];
    __shared__ u64 s_excl[5];
    __shared__ u64 s_agg[5];

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    // tile id = block id: blocks are dispatched in index order, so every tile this one waits on in
    // the look-back is already resident or finished (the same assumption CUB's scan makes)
    const int tile = blockIdx.x;
    const GridDims g = P.g;
    const i64 w64 = (i64)tile * TILE + tid;
    const bool valid = w64 < P.N;
    const int rC = (int)w64;
    const bool up = P.upwind != 0;

    // ================= phase 0: pattern =================
    int L = 0, k = 0, p2 = 0;
    int Lc[7];
    int r[7];
    unsigned low[7];   // low[c]: present candidates with a smaller wet rank than c
    unsigned wetm = 0, act = 0, mlm = 0;   // wet neighbours; neighbours whose flux enters this column; ML pairs
    double pmag[7];
    bool fold = false, generic = false, irregular = false;
    unsigned m_T = 0, m_adv = 0, m_kh = 0, m_ml = 0, m_dp = 0;
    unsigned errbits = 0;  // 1 dry nbr, 2 nan adv, 4 nan kh, 8 nan ml, 16 nan deep, 32 zero dropped, 64 nan rho
    int cnt[5] = {0, 0, 0, 0, 0};
#pragma unroll
    for (int c = 0; c < 7; ++c) {
        Lc[c] = 0;
        r[c] = 0x7fffffff;
        pmag[c] = 0.0;
        low[c] = 0;
    }
    if (valid) {
        L = __ldg(P.lwet + rC);
        k = (int)fdiv((unsigned)L, P.divP);
        p2 = L - k * g.P;
        const int j = (int)fdiv((unsigned)p2, P.divNx);
        const int i = p2 - j * g.nx;
        fold = (j == g.ny - 1) && (g.topo == OTMB_TOPO_TRIPOLAR);
        const bool hasT = k > 0, hasB = k < g.nz - 1, hasS = j > 0, hasN = (j < g.ny - 1) || fold;
        const bool seamW = i == 0, seamE = i == g.nx - 1;
        irregular = seamW || seamE || fold;
        Lc[cC] = L;
        Lc[cT] = hasT ? L - g.P : L;
        Lc[cB] = hasB ? L + g.P : L;
        Lc[cS] = hasS ? L - g.nx : L;
        Lc[cW] = seamW ? L + (g.nx - 1) : L - 1;
        Lc[cE] = seamE ? L - (g.nx - 1) : L + 1;
        Lc[cN] = (j < g.ny - 1) ? L + g.nx : (fold ? k * g.P + (g.ny - 1) * g.nx + (g.nx - 1 - i) : L);
        r[cC] = rC;
        const bool ex[7] = {hasT, hasS, true, true, true, hasN, hasB};
        // ---- batch 1 of loads (all unconditional): mask word + word prefix of every candidate
        // (L1/L2 resident, 12 bytes per 64 cells), the six face fluxes, the mixed-layer inputs
        u64 word[7];
        int pre[7];
#pragma unroll
        for (int c = 0; c < 7; ++c) {
            if (c == cC) continue;
            word[c] = __ldg(P.mask + (Lc[c] >> 6));
            pre[c] = (int)__ldg(P.wpre + (Lc[c] >> 6));
        }
        const double xT = __ldg(P.pb + Lc[cT]);                       // emitter above: its Bottom slot, max
        const double xS = __ldg(P.pn + Lc[cS]);                       // its North slot, min
        const double xW = __ldg(P.pe + Lc[cW]);                       // its East slot, min
        const double xE = __ldg(P.pw + Lc[cE]);                       // its West slot, max
        const double xN = __ldg((fold ? P.pn : P.ps) + Lc[cN]);       // its South slot (max), or North on the fold (min)
        const double xB = __ldg(P.pt + Lc[cB]);                       // emitter below: its Top slot, min
        const double ml = __ldg(P.mlotst + p2);
        const double z0 = __ldg(P.zt + k), zT = __ldg(P.zt + (hasT ? k - 1 : k)), zB = __ldg(P.zt + (hasB ? k + 1 : k));
        // wet bits and ranks; W/E ranks follow from linear adjacency off the seam
#pragma unroll
        for (int c = 0; c < 7; ++c) {
            if (c == cC) continue;
            const bool wet = ex[c] && ((word[c] >> (Lc[c] & 63)) & 1ull);
            int rk = pre[c] + __popcll(word[c] & ((1ull << (Lc[c] & 63)) - 1ull));
            if (c == cW && !seamW) rk = rC - 1;
            if (c == cE && !seamE) rk = rC + 1;
            if (wet) {
                wetm |= 1u << c;
                r[c] = rk;
            }
        }
        // face flux each neighbour carries through the face it shares with this cell (the value the
        // reference reads at the neighbour, :244-295)
        {
            const double fT = upflux(xT, true, up), fS = upflux(xS, false, up), fW = upflux(xW, false, up),
                         fE = upflux(xE, true, up), fN = upflux(xN, !fold, up), fB = upflux(xB, false, up);
            if ((wetm & bT) && nz(fT)) { act |= bT; pmag[cT] = fT; }
            if ((wetm & bS) && nz(fS)) { act |= bS; pmag[cS] = -fS; }
            if ((wetm & bW) && nz(fW)) { act |= bW; pmag[cW] = -fW; }
            if ((wetm & bE) && nz(fE)) { act |= bE; pmag[cE] = fE; }
            if ((wetm & bN) && nz(fN)) { act |= bN; pmag[cN] = fold ? -fN : fN; }
            if ((wetm & bB) && nz(fB)) { act |= bB; pmag[cB] = -fB; }
        }
        // own faces that point at a dry or absent cell: the reference would push `missing` (:247-250)
        if (P.build & 2) {
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
        // ---- row order: [T] [S] {W, C, E, N-on-fold} [N-regular] [B]; only the braces can permute
        {
            const unsigned present = wetm | bC;
            const unsigned before = present & (bT | bS);
            low[cT] = 0;
            low[cS] = present & bT;
            low[cB] = present & ~bB;
            if (!irregular) {
                low[cW] = before;
                low[cC] = before | (present & bW);
                low[cE] = before | (present & (bW | bC));
                low[cN] = before | (present & (bW | bC | bE));
            } else {
                const bool pW = present & bW, pE = present & bE, pNf = fold && (present & bN);
                unsigned lW = 0, lC = 0, lE = 0, lN = 0;
                if (pW) { if (r[cW] < rC) lC |= bW; else lW |= bC; generic |= r[cW] == rC; }
                if (pE) { if (r[cE] < rC) lC |= bE; else lE |= bC; generic |= r[cE] == rC; }
                if (pW && pE) { if (r[cW] < r[cE]) lE |= bW; else lW |= bE; generic |= r[cW] == r[cE]; }
                if (pNf) {
                    if (r[cN] < rC) lC |= bN; else lN |= bC;
                    generic |= r[cN] == rC;
                    if (pW) { if (r[cN] < r[cW]) lW |= bN; else lN |= bW; generic |= r[cN] == r[cW]; }
                    if (pE) { if (r[cN] < r[cE]) lE |= bN; else lN |= bE; generic |= r[cN] == r[cE]; }
                }
                low[cW] = before | lW;
                low[cC] = before | lC;
                low[cE] = before | lE;
                low[cN] = fold ? (before | lN) : (before | (present & (bW | bC | bE)));
            }
        }
        if (!generic) {
            cnt[0] = __popc(m_T);
            cnt[1] = __popc(m_adv);
            cnt[2] = __popc(m_kh);
            cnt[3] = __popc(m_ml);
            cnt[4] = __popc(m_dp);
        } else {
            int g_r[7];   // local copy: distinct_rows indexes dynamically
#pragma unroll
            for (int c = 0; c < 7; ++c) g_r[c] = r[c];
            cnt[0] = distinct_rows(m_T, g_r);
            cnt[1] = distinct_rows(m_adv, g_r);
            cnt[2] = distinct_rows(m_kh, g_r);
            cnt[3] = distinct_rows(m_ml, g_r);
            cnt[4] = distinct_rows(m_dp, g_r);
        }
    }

    // ================= tile scan + decoupled look-back =================
    u64 packed = 0;
#pragma unroll
    for (int q = 0; q < 5; ++q) packed |= (u64)cnt[q] << (12 * q);
    u64 incl = packed;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const u64 o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    u64 wbase = 0, total = 0;
#pragma unroll
    for (int q = 0; q < NW; ++q) {
        const u64 sw = s_warp[q];
        if (q < wid) wbase += sw;
        total += sw;
    }
    const u64 excl_packed = wbase + incl - packed;   // in-tile exclusive offsets of this column, 12 bits each
    if (wid < (NW < 5 ? NW : 5)) {   // one warp per counter; the other warps go straight to phase 1
        for (int m = wid; m < 5; m += NW) {
            const u64 agg = (total >> (12 * m)) & 0xfffull;
            if (lane == 0) {
                s_agg[m] = agg;
                st_vol(P.tile_state + (size_t)tile * 8 + m, (tile == 0 ? ST_PRE : ST_AGG) | agg);
            }
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
            if (lane == 0) s_excl[m] = excl;
        }
    }

    // ================= phase 1: values, staged in shared memory =================
    if (valid) {
        int sidx[5];   // smem index of this column's first entry, per matrix
#pragma unroll
        for (int q = 0; q < 5; ++q) sidx[q] = ebase(q) * TILE + (int)((excl_packed >> (12 * q)) & 0xfffull);
        if (generic) {
            // copies: only these escape to the out-of-line routine, so Lc/r/pmag stay in registers
            int g_Lc[7], g_r[7], g_s[5];
            double g_p[7];
            unsigned g_err = 0;
#pragma unroll
            for (int c = 0; c < 7; ++c) {
                g_Lc[c] = Lc[c];
                g_r[c] = r[c];
                g_p[c] = pmag[c];
            }
#pragma unroll
            for (int q = 0; q < 5; ++q) g_s[q] = sidx[q];
            generic_full(P, L, k, g_Lc, g_r, wetm, fold, act, g_p, mlm, g_s, srow, sval, &g_err);
            errbits |= g_err;
            atomicAdd(&P.flags->generic_columns, 1);
        } else {
            const int PP = g.P;
            // ---- batch 2 of loads: every value input of the column, unconditional, one round trip
            double vn[7], rh[7];
#pragma unroll
            for (int c = 0; c < 7; ++c) {
                vn[c] = __ldg(P.v3D + Lc[c]);
                rh[c] = RHO3D ? __ldg(P.rho3d + Lc[c]) : P.rho;
            }
            const double vC = vn[cC];
            // horizontal neighbours S,W,E,N: thickness, own/opposite edge length, own/opposite centre distance
            double thn[7], e_own[7], e_opp[7], d_own[7], d_opp[7];
#pragma unroll
            for (int c = 0; c < 7; ++c) {
                thn[c] = e_own[c] = e_opp[c] = d_own[c] = d_opp[c] = 0.0;
                if (!(HMASK >> c & 1)) continue;
                const int own = c == cW ? OTMB_DIR_WEST : c == cE ? OTMB_DIR_EAST : c == cS ? OTMB_DIR_SOUTH : OTMB_DIR_NORTH;
                const int opp = c == cW ? OTMB_DIR_EAST : c == cE ? OTMB_DIR_WEST : c == cS ? OTMB_DIR_NORTH
                                                                                  : (fold ? OTMB_DIR_NORTH : OTMB_DIR_SOUTH);
                const int q2 = Lc[c] - k * PP;
                thn[c] = __ldg(P.thk + Lc[c]);
                e_own[c] = __ldg(P.edge + own * PP + p2);
                e_opp[c] = __ldg(P.edge + opp * PP + q2);
                d_own[c] = __ldg(P.dnbr + own * PP + p2);
                d_opp[c] = __ldg(P.dnbr + opp * PP + q2);
            }
            const double thC = __ldg(P.thk + L);
            const double area = __ldg(P.area2D + p2), ztC = __ldg(P.zt + k);
            const double ztT = __ldg(P.zt + (k > 0 ? k - 1 : k)), ztB = __ldg(P.zt + (k < g.nz - 1 ? k + 1 : k));
            double Tv[7];
#pragma unroll
            for (int c = 0; c < 7; ++c) Tv[c] = 0.0;

            // ---- Tadv (:193-204)
            if (m_adv) {
                const double rhoC = rh[cC];
                if (RHO3D && isnan(rhoC)) errbits |= 64u;
                double dd[7];
                bool bad = false;
#pragma unroll
                for (int c = 0; c < 7; ++c) {
                    dd[c] = 0.0;
                    if (c == cC || !(act >> c & 1)) continue;
                    const double rb = (rh[c] + rhoC) / 2;
                    const double a = -pmag[c] / (rb * vn[c]);
                    dd[c] = pmag[c] / (rb * vC);
                    bad |= isnan(a) || isnan(dd[c]);
                    const int s = sidx[1] + __popc(m_adv & low[c]);
                    srow[s] = r[c];
                    sval[s] = a;
                    Tv[c] = a;
                }
                if (bad) errbits |= 2u;
                // diagonal: emitter contributions in ascending wet rank; sparse! keeps the first value and
                // adds the later ones.  Order: T, S, the row group {W, E, N-on-fold} by rank, N-regular, B.
                double dsum = 0.0;
                bool first = true;
                auto add = [&](bool on, double d) {
                    if (on) {
                        dsum = first ? d : dsum + d;
                        first = false;
                    }
                };
                add(act & bT, dd[cT]);
                add(act & bS, dd[cS]);
                if (!irregular) {
                    add(act & bW, dd[cW]);
                    add(act & bE, dd[cE]);
                } else {
                    int k0 = (act & bW) ? r[cW] : 0x7fffffff, k1 = (act & bE) ? r[cE] : 0x7fffffff,
                        k2 = (fold && (act & bN)) ? r[cN] : 0x7fffffff;
                    double v0 = dd[cW], v1 = dd[cE], v2 = dd[cN];
                    cex(k0, v0, k1, v1);
                    cex(k1, v1, k2, v2);
                    cex(k0, v0, k1, v1);
                    add(k0 != 0x7fffffff, v0);
                    add(k1 != 0x7fffffff, v1);
                    add(k2 != 0x7fffffff, v2);
                }
                add(!fold && (act & bN), dd[cN]);
                add(act & bB, dd[cB]);
                const int s = sidx[1] + __popc(m_adv & low[cC]);
                srow[s] = rC;
                sval[s] = dsum;
                Tv[cC] = dsum;
            } else if (RHO3D && (P.build & 2)) {
                if (isnan(rh[cC])) errbits |= 64u;
            }

            // ---- TκH (:348-415, :426-435); own slots in emit order W,E,S,N
            if (m_kh) {
                double dsum = 0.0;
                bool first = true, bad = false;
                const int ord[4] = {cW, cE, cS, cN};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int c = ord[q];
                    if (!(wetm >> c & 1)) continue;
                    const double a_own = thC * e_own[c];
                    const double a_nbr = thn[c] * e_opp[c];
                    const double ka = P.kH * jl_min(a_own, a_nbr);
                    const double ts = ka / (d_own[c] * vC);       // row 𝑗 seen from 𝑗
                    const double tn = ka / (d_opp[c] * vn[c]);    // row 𝑖 seen from 𝑖
                    bad |= isnan(ts) || isnan(tn);
                    dsum = first ? ts : dsum + ts;
                    first = false;
                    const int s = sidx[2] + __popc(m_kh & low[c]);
                    srow[s] = r[c];
                    sval[s] = -tn;
                    Tv[c] = Tv[c] + (-tn);
                }
                if (bad) errbits |= 4u;
                const int s = sidx[2] + __popc(m_kh & low[cC]);
                srow[s] = rC;
                sval[s] = dsum;
                Tv[cC] = Tv[cC] + dsum;
            }

            // ---- TκVML and TκVdeep (:450-477); own slots in emit order B, T
            if (m_dp | m_ml) {
                double mls = 0.0, dps = 0.0, mlT = 0.0, mlB = 0.0, dpT = 0.0, dpB = 0.0;
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
                        if (c == cT) mlT = -tn; else mlB = -tn;
                    }
                }
                if (badm) errbits |= 8u;
                if (badd) errbits |= 16u;
                // T folds TκVML before TκVdeep
                if (m_ml) {
                    if (mlm & bT) { const int s = sidx[3] + __popc(m_ml & low[cT]); srow[s] = r[cT]; sval[s] = mlT; Tv[cT] = Tv[cT] + mlT; }
                    if (mlm & bB) { const int s = sidx[3] + __popc(m_ml & low[cB]); srow[s] = r[cB]; sval[s] = mlB; Tv[cB] = Tv[cB] + mlB; }
                    const int s = sidx[3] + __popc(m_ml & low[cC]);
                    srow[s] = rC;
                    sval[s] = mls;
                    Tv[cC] = Tv[cC] + mls;
                }
                if (m_dp) {
                    if (wetm & bT) { const int s = sidx[4] + __popc(m_dp & low[cT]); srow[s] = r[cT]; sval[s] = dpT; Tv[cT] = Tv[cT] + dpT; }
                    if (wetm & bB) { const int s = sidx[4] + __popc(m_dp & low[cB]); srow[s] = r[cB]; sval[s] = dpB; Tv[cB] = Tv[cB] + dpB; }
                    const int s = sidx[4] + __popc(m_dp & low[cC]);
                    srow[s] = rC;
                    sval[s] = dps;
                    Tv[cC] = Tv[cC] + dps;
                }
            }

            // ---- T: union pattern; exact zeros are flagged and removed by the compaction pass
            if (m_T) {
                bool zero = false;
#pragma unroll
                for (int c = 0; c < 7; ++c) {
                    if (!(m_T >> c & 1)) continue;
                    const int s = sidx[0] + __popc(m_T & low[c]);
                    srow[s] = r[c];
                    sval[s] = Tv[c];
                    zero |= (Tv[c] == 0.0);
                }
                if (zero) errbits |= 32u;
            }
        }
    }
    __syncthreads();   // staged tile complete, s_excl / s_agg published

    // ================= phase 2: coalesced flush =================
    if (valid) {
#pragma unroll
        for (int q = 0; q < 5; ++q)
            if (P.build >> q & 1)
                P.colptr[q][rC] = (i64)(s_excl[q] + ((excl_packed >> (12 * q)) & 0xfffull)) + P.base;
    }
    if (tile == P.ntiles - 1 && tid < 5) {
        const u64 nnz = s_excl[tid] + s_agg[tid];
        P.flags->nnz[tid] = nnz;
        if (P.build >> tid & 1) P.colptr[tid][P.N] = (i64)nnz + P.base;
    }
#pragma unroll
    for (int q = 0; q < 5; ++q) {
        if (!(P.build >> q & 1)) continue;
        const int n = (int)s_agg[q];
        const u64 g0 = s_excl[q];
        i64* __restrict__ rv = P.rowval[q] + g0;
        double* __restrict__ nv = P.nzval[q] + g0;
        const int* sr = srow + ebase(q) * TILE;
        const double* sv = sval + ebase(q) * TILE;
        const int head = (int)(g0 & 1ull) < n ? (int)(g0 & 1ull) : n;   // make the global side 16-byte aligned
        if (tid == 0 && head) {
            rv[0] = (i64)sr[0] + P.base;
            nv[0] = sv[0];
        }
        const int npair = (n - head) >> 1;
        for (int pidx = tid; pidx < npair; pidx += TILE) {
            const int idx = head + 2 * pidx;
            longlong2 rr;
            rr.x = (i64)sr[idx] + P.base;
            rr.y = (i64)sr[idx + 1] + P.base;
            double2 vv;
            vv.x = sv[idx];
            vv.y = sv[idx + 1];
            *reinterpret_cast<longlong2*>(rv + idx) = rr;
            *reinterpret_cast<double2*>(nv + idx) = vv;
        }
        if (tid == 0 && ((n - head) & 1)) {
            rv[n - 1] = (i64)sr[n - 1] + P.base;
            nv[n - 1] = sv[n - 1];
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

// ---------------------------------------------------------------------------------------
// zero-dropping compaction of one CSC matrix (only runs when the flag says a zero was stored)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_count_nonzero(const i64* __restrict__ colptr, const double* __restrict__ nzv, i64 n,
                                                       int base, uint32_t* __restrict__ cnt) {
    const i64 col = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (col > n) return;
    uint32_t m = 0;
    if (col < n)
        for (i64 p = colptr[col] - base; p < colptr[col + 1] - base; ++p) m += (nzv[p] != 0.0);
    cnt[col] = m;
}
__global__ void __launch_bounds__(256) k_copy_nonzero(const i64* __restrict__ colptr, const i64* __restrict__ rv,
                                                      const double* __restrict__ nzv, i64 n, int base,
                                                      const i64* __restrict__ ncp0, i64* __restrict__ ncp,
                                                      i64* __restrict__ nrv, double* __restrict__ nnz_) {
    const i64 col = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (col > n) return;
    i64 o = ncp0[col];
    ncp[col] = o + base;
    if (col == n) return;
    for (i64 p = colptr[col] - base; p < colptr[col + 1] - base; ++p)
        if (nzv[p] != 0.0) {
            nrv[o] = rv[p];
            nnz_[o] = nzv[p];
            ++o;
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
int launch_v2(otmb_ctx* c, V2Params& P) {
    const int ntiles = (int)((c->N + TILE - 1) / TILE);
    P.ntiles = ntiles;
    CU_TRY(c, c->tile_state.ensure((size_t)ntiles * 8 * sizeof(u64)));
    P.tile_state = c->tile_state.as<u64>();
    CU_TRY(c, cudaMemsetAsync(P.tile_state, 0, (size_t)ntiles * 8 * sizeof(u64), c->stream));
    const size_t smem = (size_t)CAPSUM * TILE * (sizeof(double) + sizeof(int));
    // per device (a process may hold one context per GPU); the call is cheap
    CU_TRY(c, cudaFuncSetAttribute(k_fused_v2<RHO3D, TILE, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_fused_v2<RHO3D, TILE, MINB><<<ntiles, TILE, smem, c->stream>>>(P);
    LAUNCHED(c);
    CU_TRY(c, cudaGetLastError());
    return OTMB_OK;
}

}  // namespace

int otmb_drop_zeros(otmb_ctx* c, int m, int base) {
    const i64 n = c->ncols;
    DevBuf &cnt = c->coo[3], &cp0 = c->coo[11];
    CU_TRY(c, cnt.ensure((size_t)(n + 1) * 4));
    CU_TRY(c, cp0.ensure((size_t)(n + 1) * 8));
    k_count_nonzero<<<grid_for(n + 1, 256), 256, 0, c->stream>>>(c->colptr[m].as<i64>(), c->nzval[m].as<double>(), n, base,
                                                                  cnt.as<uint32_t>());
    LAUNCHED(c);
    OT_TRY(otmb_scan_u32_to_i64(c, cnt.as<uint32_t>(), cp0.as<i64>(), n + 1, &c->flags.as<DevFlags>()->nnz[m]));
    CU_TRY(c, c->add_tmp[0].ensure(c->colptr[m].cap));
    CU_TRY(c, c->add_tmp[1].ensure(c->rowval[m].cap));
    CU_TRY(c, c->add_tmp[2].ensure(c->nzval[m].cap));
    k_copy_nonzero<<<grid_for(n + 1, 256), 256, 0, c->stream>>>(c->colptr[m].as<i64>(), c->rowval[m].as<i64>(),
                                                                 c->nzval[m].as<double>(), n, base, cp0.as<i64>(),
                                                                 c->add_tmp[0].as<i64>(), c->add_tmp[1].as<i64>(),
                                                                 c->add_tmp[2].as<double>());
    LAUNCHED(c);
    CU_TRY(c, cudaGetLastError());
    OT_TRY(otmb_fetch_flags(c));
    std::swap(c->colptr[m], c->add_tmp[0]);
    std::swap(c->rowval[m], c->add_tmp[1]);
    std::swap(c->nzval[m], c->add_tmp[2]);
    c->nnz[m] = (i64)c->h_flags->nnz[m];
    return OTMB_OK;
}

int otmb_fused_v2_build(otmb_ctx* c, const otmb_tm_params* prm, int build) {
    V2Params P;
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
    P.mask = c->mask.as<u64>();
    P.wpre = c->wpre.as<uint32_t>();
    P.lwet = c->lwet.as<int>();
    P.kH = prm->kH;
    P.kVML = prm->kVML;
    P.kVdeep = prm->kVdeep;
    P.rho = prm->rho;
    P.upwind = prm->upwind;
    P.base = prm->index_base;
    P.build = build;
    P.N = c->N;
    P.flags = c->flags.as<DevFlags>();
    const int cap_per_col[5] = {CAP0, CAP1, CAP2, CAP3, CAP4};
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
    if (c->have_rho3d) return launch_v2<true, 256, 2>(c, P);
    return launch_v2<false, 256, 2>(c, P);
}
