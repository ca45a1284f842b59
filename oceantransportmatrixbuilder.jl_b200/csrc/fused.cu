// fused.cu — direct CSC assembly of Tadv, TκH, TκVML, TκVdeep and T, one thread per grid cell
// (= one matrix column per wet cell), without materialising COO triplets.
//
// Replaces, for the whole of transportmatrix (/root/reference/src/matrixbuilding.jl:128-150):
//   advection_operator_sparse_entries   :221-299  + pushTadvectionvalues! :193-204
//   horizontal_diffusion_..._entries    :337-418  + pushTmixingvalues!    :426-435
//   vertical_diffusion_..._entries      :438-479  (Ω mask :85 for TκVML, trues for TκVdeep :109)
//   the NaN checks :39,61,90,114,233, the four sparse(...) calls :41,63,92,116 and the sum :147.
//
// Why gather form is exact.  The reference loops over emitting cells 𝑖 and pushes triplets;
// column 𝑗 of every operator only ever receives triplets from 𝑗 itself and from its ≤ 6 grid
// neighbours, because the neighbour relation is symmetric: (𝑖, slot s) points at 𝑗 iff
// (𝑗, slot s') points at 𝑖 with s = opposite(s') (W<->E, S<->N, B<->T; on the tripolar fold
// N<->N).  So a thread that owns column 𝑗 re-evaluates, for each neighbour 𝑖, exactly the
// arithmetic the reference performs when it visits 𝑖 (𝑖's volume, 𝑖's distance, the face flux
// stored at 𝑖), and SparseArrays.sparse's duplicate summation order — input order, i.e.
// ascending emitter index, then slot order W,E,S,N,B,T — is reproduced by adding diagonal
// contributions in ascending wet-rank order.  Rows inside a column are ordered by wet rank
// (natural order top, south, west, self, east, north, bottom except at the periodic seam and
// the fold, where the ranks are compared explicitly).  Neighbour coincidences (fold centre,
// seam∩fold, nx ≤ 2, odd-nx self neighbour) create duplicate (row, col) pairs: those columns
// take a generic in-thread sort-and-combine branch.
//
// Output offsets: each 256-cell tile counts its entries per matrix, and a single-pass
// decoupled look-back scan (tickets in launch order, 5 counters, one warp per counter)
// turns the counts into global offsets inside the same kernel — inputs are read once and
// outputs written once.  OTMB_PATH_FUSED2 runs the same column code as count pass + scan +
// fill pass and serves as the cross-check of the look-back.
#include "common.cuh"

namespace {

constexpr int TILE = 256;
constexpr int NW = TILE / 32;
constexpr u64 ST_AGG = 1ull << 62, ST_PRE = 2ull << 62, ST_MASK = (1ull << 62) - 1;

enum { cT = 0, cS = 1, cW = 2, cC = 3, cE = 4, cN = 5, cB = 6 };
// emit-order slot numbers of the reference loop: W,E,S,N,B,T
enum { sW = 0, sE = 1, sS = 2, sN = 3, sB = 4, sT = 5 };

struct FusedParams {
    GridDims g;
    const double *v3D, *thk, *area2D, *zt, *edge, *dnbr, *mlotst, *rho3d;
    const double *pe, *pw, *pn, *ps, *pt, *pb;
    const u64* mask;
    const uint32_t* wpre;
    double kH, kVML, kVdeep, rho;
    int upwind, base, build;  // build: bit m set -> matrix m is produced (bit 0 = T)
    int ntiles;
    i64 N;
    i64* colptr[5];
    i64* rowval[5];
    double* nzval[5];
    DevFlags* flags;
    u64* tile_state;  // look-back: [tile*8 + m]; two-pass: totals / offsets [m*ntiles + tile]
};

struct GenOut {
    int cnt[5];
    int rows[5][8];
    double vals[5][8];
};

struct Ent {
    int row;
    i64 key;
    double val;
};

__device__ __forceinline__ void ent_sort(Ent* e, int n) {
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
// sparse!'s in-order combine: first occurrence kept, later ones added left to right
__device__ __forceinline__ int ent_combine(const Ent* e, int n, int* rows, double* vals) {
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

// Generic branch for columns whose candidate neighbours coincide.  Inputs are the per-slot
// values already computed by the caller, indexed by candidate c (T,S,W,self,E,N,B).
__device__ __noinline__ void generic_column(const int* r, unsigned wetm, bool fold, const unsigned* act /*adv,kh,ml,dp*/,
                                            const double* adv, const double* dcon, const double* khn,
                                            const double* khs, const double* mln, const double* mls,
                                            const double* dpn, const double* dps, int build, GenOut& out) {
    const int emit_slot[7] = {sB, sN, sE, -1, sW, fold ? sN : sS, sT};  // slot of the neighbour that points at us
    const int own_slot[7] = {sT, sS, sW, -1, sE, sN, sB};               // our slot that points at the neighbour
    const int rC = r[cC];
    Ent e[16];
    int n;
    // Tadv: emitter = neighbour c: (row 𝑖, -p/m𝑖) then (row 𝑗, +p/m𝑗)
    n = 0;
    for (int c = 0; c < 7; ++c)
        if (c != cC && (act[0] >> c & 1)) {
            i64 kb = (i64)r[c] * 16 + emit_slot[c] * 2;
            e[n++] = Ent{r[c], kb, adv[c]};
            e[n++] = Ent{rC, kb + 1, dcon[c]};
        }
    ent_sort(e, n);
    out.cnt[1] = ent_combine(e, n, out.rows[1], out.vals[1]);
    // mixing operators: emitter 𝑗 (self): (row 𝑗, +t_self); emitter 𝑖: (row 𝑖, -t_nbr)
    for (int op = 1; op <= 3; ++op) {
        const double* nb = op == 1 ? khn : (op == 2 ? mln : dpn);
        const double* sf = op == 1 ? khs : (op == 2 ? mls : dps);
        n = 0;
        for (int c = 0; c < 7; ++c)
            if (c != cC && (act[op] >> c & 1)) {
                e[n++] = Ent{rC, (i64)rC * 16 + own_slot[c] * 2, sf[c]};
                e[n++] = Ent{r[c], (i64)r[c] * 16 + emit_slot[c] * 2 + 1, nb[c]};
            }
        ent_sort(e, n);
        out.cnt[op + 1] = ent_combine(e, n, out.rows[op + 1], out.vals[op + 1]);
    }
    // T = ((Tadv + TκH) + TκVML) + TκVdeep on the union of the four row lists, zeros dropped
    int idx[4] = {0, 0, 0, 0};
    int m = 0;
    while (true) {
        int row = 0x7fffffff;
        for (int q = 0; q < 4; ++q)
            if (idx[q] < out.cnt[q + 1] && out.rows[q + 1][idx[q]] < row) row = out.rows[q + 1][idx[q]];
        if (row == 0x7fffffff) break;
        double x = 0.0;
        for (int q = 0; q < 4; ++q) {
            double v = 0.0;
            if (idx[q] < out.cnt[q + 1] && out.rows[q + 1][idx[q]] == row) {
                v = out.vals[q + 1][idx[q]];
                ++idx[q];
            }
            x = (q == 0) ? v : x + v;
        }
        // q == 0 contributes "adv or 0.0"; the first addition is adv + kH exactly as in the fast path
        if (x != 0.0) {
            out.rows[0][m] = row;
            out.vals[0][m] = x;
            ++m;
        }
    }
    out.cnt[0] = m;
    for (int q = 0; q < 5; ++q)
        if (!(build >> q & 1)) out.cnt[q] = 0;
    (void)wetm;
}

__device__ __forceinline__ u64 ld_vol(const u64* p) { return *reinterpret_cast<const volatile u64*>(p); }
__device__ __forceinline__ void st_vol(u64* p, u64 v) { *reinterpret_cast<volatile u64*>(p) = v; }

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

// MODE 0: count only (tile totals); MODE 1: fill with precomputed tile offsets; MODE 2: single pass, look-back
template <int MODE>
__global__ void __launch_bounds__(TILE) k_fused(const FusedParams P) {
    __shared__ u64 s_warp[NW];
    __shared__ u64 s_excl[5];
    __shared__ u64 s_agg[5];
    __shared__ int s_tile;

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int tile;
    if (MODE == 2) {
        if (tid == 0) s_tile = (int)atomicAdd(&P.flags->ticket, 1ull);
        __syncthreads();
        tile = s_tile;
    } else {
        tile = blockIdx.x;
    }
    const GridDims g = P.g;
    const int L = tile * TILE + tid;
    const bool inb = L < g.M;
    const bool wetC = inb && wet_at(P.mask, L);

    // ---- per-column state (all statically indexed -> registers)
    int r[7];
    unsigned wetm = 0;            // candidate exists and is wet (bit cC = self)
    unsigned m_adv = 0, m_kh = 0, m_ml = 0, m_dp = 0, m_T = 0;  // presence per matrix
    double adv[7], khv[7], mlv[7], dpv[7], Tv[7];
    // per-slot pieces kept for the generic branch
    double dcon[7], khs[7], mls[7], dps[7];
    unsigned a_adv = 0, a_kh = 0, a_ml = 0, a_dp = 0;  // active neighbour slots per operator
    unsigned lower[7];
    bool fold = false, generic = false;
    int cnt[5] = {0, 0, 0, 0, 0};
#pragma unroll
    for (int c = 0; c < 7; ++c) {
        r[c] = 0x7fffffff;
        adv[c] = khv[c] = mlv[c] = dpv[c] = Tv[c] = 0.0;
        dcon[c] = khs[c] = mls[c] = dps[c] = 0.0;
        lower[c] = 0;
    }
    bool e_dry = false, e_nadv = false, e_nkh = false, e_nml = false, e_ndp = false, e_nrho = false;

    if (inb) {
        const int k = L / g.P;
        const int p2 = L - k * g.P;
        const int j = p2 / g.nx;
        const int i = p2 - j * g.nx;
        fold = (j == g.ny - 1) && (g.topo == OTMB_TOPO_TRIPOLAR);
        int Lc[7];
        Lc[cT] = k > 0 ? L - g.P : -1;
        Lc[cB] = k < g.nz - 1 ? L + g.P : -1;
        Lc[cS] = j > 0 ? L - g.nx : -1;
        Lc[cW] = i > 0 ? L - 1 : L + (g.nx - 1);
        Lc[cE] = i < g.nx - 1 ? L + 1 : L - (g.nx - 1);
        Lc[cN] = j < g.ny - 1 ? L + g.nx : (fold ? k * g.P + (g.ny - 1) * g.nx + (g.nx - 1 - i) : -1);
        Lc[cC] = L;
#pragma unroll
        for (int c = 0; c < 7; ++c) {
            if (c == cC) {
                if (wetC) { wetm |= 1u << c; r[c] = rank_at(P.mask, P.wpre, L); }
            } else if (Lc[c] >= 0 && wet_at(P.mask, Lc[c])) {
                wetm |= 1u << c;
                r[c] = rank_at(P.mask, P.wpre, Lc[c]);
            }
        }
        const bool up = P.upwind != 0;
        // ---- face flux of each wet neighbour through the face it shares with this cell:
        // the flux the reference reads at the neighbour when it visits it (:244-295)
        double pmag[7];
        unsigned fact = 0;
#pragma unroll
        for (int c = 0; c < 7; ++c) {
            pmag[c] = 0.0;
            if (c == cC || !(wetm >> c & 1)) continue;
            const double* src;
            bool take_max;  // upwind: max(ϕ,0) for W,S,B slots of the emitter, min(ϕ,0) for E,N,T
            if (c == cT) { src = P.pb; take_max = true; }          // emitter above us, its Bottom slot
            else if (c == cB) { src = P.pt; take_max = false; }    // emitter below us, its Top slot
            else if (c == cS) { src = P.pn; take_max = false; }    // its North slot
            else if (c == cW) { src = P.pe; take_max = false; }    // its East slot
            else if (c == cE) { src = P.pw; take_max = true; }     // its West slot
            else { src = fold ? P.pn : P.ps; take_max = !fold; }   // c == cN: South slot, or North on the fold
            const double x = __ldg(src + Lc[c]);
            const double f = up ? (take_max ? jl_max(x, 0.0) : jl_min(x, 0.0)) : x / 2;
            if (f > 0 || f < 0) {
                fact |= 1u << c;
                pmag[c] = take_max ? f : -f;
            }
        }
        if (!wetC) {
            // a wet neighbour with a non-zero flux through a face shared with a dry cell:
            // the reference would push `missing` into 𝑗s (MethodError)
            if (fact) e_dry = true;
        } else {
            // own faces that point at nothing (j₋₁ at j=1, k₊₁ at k=nz, j₊₁ on a bipolar top row)
            if (j == 0) {
                const double x = __ldg(P.ps + L);
                const double f = up ? jl_max(x, 0.0) : x / 2;
                if (f > 0 || f < 0) e_dry = true;
            }
            if (k == g.nz - 1) {
                const double x = __ldg(P.pb + L);
                const double f = up ? jl_max(x, 0.0) : x / 2;
                if (f > 0 || f < 0) e_dry = true;
            }
            if (j == g.ny - 1 && !fold) {
                const double x = __ldg(P.pn + L);
                const double f = up ? jl_min(x, 0.0) : x / 2;
                if (f > 0 || f < 0) e_dry = true;
            }
            const double vC = __ldg(P.v3D + L);
            const double rhoC = P.rho3d ? __ldg(P.rho3d + L) : P.rho;
            if (isnan(rhoC)) e_nrho = true;
            double vn[7];
#pragma unroll
            for (int c = 0; c < 7; ++c) vn[c] = (c != cC && (wetm >> c & 1)) ? __ldg(P.v3D + Lc[c]) : 0.0;

            // ---- Tadv (:193-204): emitter 𝑖 = neighbour c, column 𝑗 = this cell
            if (P.build & 2) {
                a_adv = fact;
#pragma unroll
                for (int c = 0; c < 7; ++c) {
                    if (c == cC || !(fact >> c & 1)) continue;
                    const double rhoi = P.rho3d ? __ldg(P.rho3d + Lc[c]) : P.rho;
                    const double rb = (rhoi + rhoC) / 2;
                    const double mi = rb * vn[c];
                    const double mj = rb * vC;
                    adv[c] = -pmag[c] / mi;
                    dcon[c] = pmag[c] / mj;
                    if (isnan(adv[c]) || isnan(dcon[c])) e_nadv = true;
                }
                m_adv = fact;
            }
            // ---- TκH (:348-415, :426-435)
            if (P.build & 4) {
                const double thC = __ldg(P.thk + L);
                const int PP = g.P;
#pragma unroll
                for (int c = 0; c < 7; ++c) {
                    if (!(c == cW || c == cE || c == cS || c == cN) || !(wetm >> c & 1)) continue;
                    const int own = c == cW ? OTMB_DIR_WEST : c == cE ? OTMB_DIR_EAST : c == cS ? OTMB_DIR_SOUTH : OTMB_DIR_NORTH;
                    const int opp = c == cW ? OTMB_DIR_EAST : c == cE ? OTMB_DIR_WEST : c == cS ? OTMB_DIR_NORTH
                                                                                      : (fold ? OTMB_DIR_NORTH : OTMB_DIR_SOUTH);
                    const int q2 = Lc[c] - k * PP;
                    const double a_own = thC * __ldg(P.edge + own * PP + p2);
                    const double a_nbr = __ldg(P.thk + Lc[c]) * __ldg(P.edge + opp * PP + q2);
                    const double a = jl_min(a_own, a_nbr);
                    const double ka = P.kH * a;
                    const double ts = ka / (__ldg(P.dnbr + own * PP + p2) * vC);      // row 𝑗, seen from 𝑗
                    const double tn = ka / (__ldg(P.dnbr + opp * PP + q2) * vn[c]);   // row 𝑖, seen from 𝑖
                    khs[c] = ts;
                    khv[c] = -tn;
                    if (isnan(ts) || isnan(tn)) e_nkh = true;
                    a_kh |= 1u << c;
                }
                m_kh = a_kh;
            }
            // ---- TκVML / TκVdeep (:450-477)
            if (P.build & (8 | 16)) {
                const double area = __ldg(P.area2D + p2);
                const double ztC = __ldg(P.zt + k);
                const double ml = __ldg(P.mlotst + p2);
                const bool omC = ztC < ml;  // false for NaN (missing), :85
#pragma unroll
                for (int c = 0; c < 7; ++c) {
                    if (!(c == cT || c == cB) || !(wetm >> c & 1)) continue;
                    const int kc = c == cT ? k - 1 : k + 1;
                    const double ztn = __ldg(P.zt + kc);
                    const double d = fabs(ztC - ztn);
                    const double qs = d * vC, qn = d * vn[c];
                    if (P.build & 16) {
                        const double ka = P.kVdeep * area;
                        dps[c] = ka / qs;
                        dpv[c] = -(ka / qn);
                        if (isnan(dps[c]) || isnan(dpv[c])) e_ndp = true;
                        a_dp |= 1u << c;
                    }
                    if ((P.build & 8) && omC && (ztn < ml)) {
                        const double ka = P.kVML * area;
                        mls[c] = ka / qs;
                        mlv[c] = -(ka / qn);
                        if (isnan(mls[c]) || isnan(mlv[c])) e_nml = true;
                        a_ml |= 1u << c;
                    }
                }
                m_ml = a_ml;
                m_dp = a_dp;
            }

            // ---- rank order of the candidates
            bool natural = true;
            {
                int prev = -1;
#pragma unroll
                for (int c = 0; c < 7; ++c)
                    if (wetm >> c & 1) {
                        if (r[c] <= prev) natural = false;
                        prev = r[c];
                    }
            }
            if (natural) {
#pragma unroll
                for (int c = 0; c < 7; ++c) lower[c] = wetm & ((1u << c) - 1u);
            } else {
#pragma unroll
                for (int c = 0; c < 7; ++c) {
                    unsigned lm = 0;
#pragma unroll
                    for (int d = 0; d < 7; ++d)
                        if (d != c && (wetm >> d & 1)) {
                            if (r[d] < r[c]) lm |= 1u << d;
                            if (r[d] == r[c] && (wetm >> c & 1)) generic = true;
                        }
                    lower[c] = lm;
                }
            }

            if (!generic) {
                // ---- diagonals.  sparse! keeps the first value and adds later ones in input order.
                // Tadv[𝑗,𝑗]: contributions of the emitters in ascending wet rank
                if (a_adv) {
                    double dsum = 0.0;
                    bool first = true;
                    if (natural) {
#pragma unroll
                        for (int c = 0; c < 7; ++c)
                            if (c != cC && (a_adv >> c & 1)) {
                                dsum = first ? dcon[c] : dsum + dcon[c];
                                first = false;
                            }
                    } else {
                        for (int t = 0; t < 7; ++t) {
#pragma unroll
                            for (int c = 0; c < 7; ++c)
                                if (c != cC && (a_adv >> c & 1) && __popc(wetm & lower[c]) == t) {
                                    dsum = first ? dcon[c] : dsum + dcon[c];
                                    first = false;
                                }
                        }
                    }
                    adv[cC] = dsum;
                    m_adv |= 1u << cC;
                }
                // mixing diagonals: own slots in emit order W,E,S,N / B,T
                if (a_kh) {
                    double dsum = 0.0;
                    bool first = true;
                    const int ord[4] = {cW, cE, cS, cN};
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (a_kh >> ord[q] & 1) {
                            dsum = first ? khs[ord[q]] : dsum + khs[ord[q]];
                            first = false;
                        }
                    khv[cC] = dsum;
                    m_kh |= 1u << cC;
                }
                if (a_ml) {
                    double dsum = 0.0;
                    bool first = true;
                    if (a_ml >> cB & 1) { dsum = mls[cB]; first = false; }
                    if (a_ml >> cT & 1) { dsum = first ? mls[cT] : dsum + mls[cT]; }
                    mlv[cC] = dsum;
                    m_ml |= 1u << cC;
                }
                if (a_dp) {
                    double dsum = 0.0;
                    bool first = true;
                    if (a_dp >> cB & 1) { dsum = dps[cB]; first = false; }
                    if (a_dp >> cT & 1) { dsum = first ? dps[cT] : dsum + dps[cT]; }
                    dpv[cC] = dsum;
                    m_dp |= 1u << cC;
                }
                // ---- T = ((Tadv + TκH) + TκVML) + TκVdeep (:147): absent operands add a literal 0.0,
                // results equal to zero are not stored
                if (P.build & 1) {
                    const unsigned any = m_adv | m_kh | m_ml | m_dp;
#pragma unroll
                    for (int c = 0; c < 7; ++c) {
                        if (!(any >> c & 1)) continue;
                        const double x = ((adv[c] + khv[c]) + mlv[c]) + dpv[c];
                        Tv[c] = x;
                        if (x != 0.0) m_T |= 1u << c;
                    }
                }
                cnt[0] = __popc(m_T);
                cnt[1] = __popc(m_adv);
                cnt[2] = __popc(m_kh);
                cnt[3] = __popc(m_ml);
                cnt[4] = __popc(m_dp);
            }
        }
    }

    // ---- generic branch (coincident neighbours): sort-and-combine in local memory
    GenOut gen;
    if (generic) {
        // copies: only these escape to the out-of-line routine, the originals stay in registers
        const unsigned act[4] = {a_adv, a_kh, a_ml, a_dp};
        int g_r[7];
        double g_adv[7], g_dcon[7], g_khv[7], g_khs[7], g_mlv[7], g_mls[7], g_dpv[7], g_dps[7];
#pragma unroll
        for (int c = 0; c < 7; ++c) {
            g_r[c] = r[c];
            g_adv[c] = adv[c]; g_dcon[c] = dcon[c];
            g_khv[c] = khv[c]; g_khs[c] = khs[c];
            g_mlv[c] = mlv[c]; g_mls[c] = mls[c];
            g_dpv[c] = dpv[c]; g_dps[c] = dps[c];
        }
        generic_column(g_r, wetm, fold, act, g_adv, g_dcon, g_khv, g_khs, g_mlv, g_mls, g_dpv, g_dps, P.build, gen);
#pragma unroll
        for (int q = 0; q < 5; ++q) cnt[q] = gen.cnt[q];
        atomicAdd(&P.flags->generic_columns, 1);
    }
#pragma unroll
    for (int q = 0; q < 5; ++q)
        if (!(P.build >> q & 1)) cnt[q] = 0;

    // ---- error flags (one atomic per warp and kind).  The flux / ρ checks belong to the advection
    // emitter (:233, :247-250) and only apply when Tadv is being built.
    if (!(P.build & 2)) e_dry = e_nrho = false;
    {
        const unsigned b0 = __ballot_sync(0xffffffffu, e_dry), b1 = __ballot_sync(0xffffffffu, e_nadv),
                       b2 = __ballot_sync(0xffffffffu, e_nkh), b3 = __ballot_sync(0xffffffffu, e_nml),
                       b4 = __ballot_sync(0xffffffffu, e_ndp), b5 = __ballot_sync(0xffffffffu, e_nrho);
        if (lane == 0) {
            if (b0) atomicOr(&P.flags->err_dry_neighbour, 1);
            if (b1) atomicOr(&P.flags->nan_adv, 1);
            if (b2) atomicOr(&P.flags->nan_kh, 1);
            if (b3) atomicOr(&P.flags->nan_kvml, 1);
            if (b4) atomicOr(&P.flags->nan_kvdeep, 1);
            if (b5) atomicOr(&P.flags->nan_rho, 1);
        }
    }

    // ---- tile-level exclusive scan of the five counts, packed 12 bits each (tile total <= 256*8)
    u64 packed = 0;
#pragma unroll
    for (int q = 0; q < 5; ++q) packed |= (u64)cnt[q] << (12 * q);
    u64 incl = packed;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        u64 o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) s_warp[wid] = incl;
    __syncthreads();
    u64 wbase = 0, total = 0;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        const u64 sw = s_warp[w];
        if (w < wid) wbase += sw;
        total += sw;
    }
    const u64 excl_packed = wbase + incl - packed;

    if (MODE == 0) {
        if (tid < 5) P.tile_state[(size_t)tid * P.ntiles + tile] = (total >> (12 * tid)) & 0xfffull;
        return;
    }

    // ---- global offsets of this tile
    if (MODE == 1) {
        if (tid < 5) s_excl[tid] = P.tile_state[(size_t)tid * P.ntiles + tile];
        __syncthreads();
    } else {
        if (tid < 5) {
            const u64 agg = (total >> (12 * tid)) & 0xfffull;
            s_agg[tid] = agg;
            st_vol(P.tile_state + (size_t)tile * 8 + tid, (tile == 0 ? ST_PRE : ST_AGG) | agg);
        }
        __syncthreads();
        if (wid < 5) {
            const int m = wid;
            u64 excl = 0;
            if (tile > 0) {
                int look = tile - 1;
                while (true) {
                    const int t = look - lane;
                    u64 w = ST_PRE;  // virtual tile -1: inclusive prefix 0
                    if (t >= 0) {
                        do {
                            w = ld_vol(P.tile_state + (size_t)t * 8 + m);
                        } while ((w >> 62) == 0);
                    }
                    const u64 val = w & ST_MASK;
                    const unsigned pm = __ballot_sync(0xffffffffu, (w >> 62) == 2);
                    if (pm) {
                        const int first = __ffs(pm) - 1;
                        excl += warp_sum<u64>(lane <= first ? val : 0ull);
                        break;
                    }
                    excl += warp_sum<u64>(val);
                    look -= 32;
                }
                if (lane == 0) st_vol(P.tile_state + (size_t)tile * 8 + m, ST_PRE | (excl + s_agg[m]));
            }
            if (lane == 0) s_excl[m] = excl;
        }
        __syncthreads();
    }
    if (tile == P.ntiles - 1 && tid < 5) {
        const u64 agg = (total >> (12 * tid)) & 0xfffull;
        const u64 nnz = s_excl[tid] + agg;
        P.flags->nnz[tid] = nnz;
        if (P.build >> tid & 1) P.colptr[tid][P.N] = (i64)nnz + P.base;
    }

    // ---- write this column
    if (!wetC) return;
    const int rC = r[cC];
    if (!generic) {
        const unsigned pm[5] = {m_T, m_adv, m_kh, m_ml, m_dp};
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            if (!(P.build >> q & 1)) continue;
            const i64 off = (i64)(s_excl[q] + ((excl_packed >> (12 * q)) & 0xfffull));
            P.colptr[q][rC] = off + P.base;
            i64* rv = P.rowval[q] + off;
            double* nv = P.nzval[q] + off;
#pragma unroll
            for (int c = 0; c < 7; ++c) {
                if (!(pm[q] >> c & 1)) continue;
                const int pos = __popc(pm[q] & lower[c]);
                rv[pos] = (i64)r[c] + P.base;
                nv[pos] = q == 0 ? Tv[c] : q == 1 ? adv[c] : q == 2 ? khv[c] : q == 3 ? mlv[c] : dpv[c];
            }
        }
    } else {
        for (int q = 0; q < 5; ++q) {
            if (!(P.build >> q & 1)) continue;
            const i64 off = (i64)(s_excl[q] + ((excl_packed >> (12 * q)) & 0xfffull));
            P.colptr[q][rC] = off + P.base;
            for (int a = 0; a < gen.cnt[q]; ++a) {
                P.rowval[q][off + a] = (i64)gen.rows[q][a] + P.base;
                P.nzval[q][off + a] = gen.vals[q][a];
            }
        }
    }
}

}  // namespace

// Fused transportmatrix on the resident inputs.  mask: bit m set -> build matrix m (bit 0 = T;
// T is only built here when all four operators are).
int otmb_fused_build(otmb_ctx* c, const otmb_tm_params* prm, int build, bool two_pass) {
    FusedParams P;
    P.g = GridDims{(int)c->nx, (int)c->ny, (int)c->nz, c->topo, (int)c->P, (int)c->M};
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
    P.kH = prm->kH;
    P.kVML = prm->kVML;
    P.kVdeep = prm->kVdeep;
    P.rho = prm->rho;
    P.upwind = prm->upwind;
    P.base = prm->index_base;
    P.build = build;
    P.N = c->N;
    P.flags = c->flags.as<DevFlags>();
    const int ntiles = (int)((c->M + TILE - 1) / TILE);
    P.ntiles = ntiles;
    // worst-case capacities: ≤ 7 rows per column for T and Tadv, 5 for TκH, 3 for the vertical operators
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
    CU_TRY(c, c->tile_state.ensure((size_t)ntiles * 8 * sizeof(u64)));
    P.tile_state = c->tile_state.as<u64>();
    if (!two_pass) {
        CU_TRY(c, cudaMemsetAsync(P.tile_state, 0, (size_t)ntiles * 8 * sizeof(u64), c->stream));
        k_fused<2><<<ntiles, TILE, 0, c->stream>>>(P);
        LAUNCHED(c);
    } else {
        k_fused<0><<<ntiles, TILE, 0, c->stream>>>(P);
        LAUNCHED(c);
        for (int m = 0; m < 5; ++m) {
            i64* seg = reinterpret_cast<i64*>(P.tile_state) + (size_t)m * ntiles;
            OT_TRY(otmb_scan_i64(c, seg, seg, ntiles, nullptr));
        }
        k_fused<1><<<ntiles, TILE, 0, c->stream>>>(P);
        LAUNCHED(c);
    }
    CU_TRY(c, cudaGetLastError());
    return OTMB_OK;
}
