// fused_generic.cuh — the coincident-neighbour ("generic") column branch shared by the fused
// assembly kernels, plus the candidate / slot numbering they use.
#pragma once
#include "common.cuh"

namespace fusedg {

enum { cT = 0, cS = 1, cW = 2, cC = 3, cE = 4, cN = 5, cB = 6 };   // candidates in regular row order
enum { sW = 0, sE = 1, sS = 2, sN = 3, sB = 4, sT = 5 };             // emit-order slots W,E,S,N,B,T
constexpr unsigned bT = 1u << cT, bS = 1u << cS, bW = 1u << cW, bC = 1u << cC, bE = 1u << cE, bN = 1u << cN, bB = 1u << cB;
constexpr unsigned HMASK = bS | bW | bE | bN;
constexpr unsigned VMASK = bT | bB;

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

template <class Params>
__device__ __noinline__ void generic_full(const Params& P, int L, int k, const int* Lc, const int* r, unsigned wetm,
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


}  // namespace fusedg
