// oracle/otmb_oracle.cpp
//
// TEST INFRASTRUCTURE ONLY.  Single-threaded CPU restatement of the matrix-assembly hot
// path of OceanTransportMatrixBuilder.jl v0.8.3, used as the parity checker by tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.  Nothing
// in the product package may import, link or call this file.
//
// PARITY UNPINNED: the reference is pure Julia, Julia is not installed in this image, and
// the reference holds no golden vectors or known-answer tests for this path (its only CI
// test downloads CMIP6 data, /root/reference/test/online.jl:19-65).  Two pieces of its
// arithmetic live in third-party code that is not under /root/reference and are restated
// here from their published algorithms:
//   * SparseArrays (Julia stdlib, compat "1", /root/reference/Project.toml:17):
//     sparse(I,J,V,m,n) = sparse! (stable counting sort by row, in-order duplicate
//     combine with +, counting sort by column, explicit zeros kept) and sparse A+B
//     (= map(+,A,B) -> _map_zeropres!, results equal to zero are not stored, n-ary +
//     folds left).
//   * Distances.jl haversine (compat "0.10", /root/reference/Project.toml:14):
//     a = sind(dlat/2)^2 + cosd(lat1)*cosd(lat2)*sind(dlon/2)^2 ; 2*(R*asin(min(sqrt(a),1))),
//     R = 6371000, arguments (lon, lat) in degrees; sind/cosd restated from Julia Base
//     (exact quadrant reduction in degrees, x/180 then a double-double multiply by pi,
//     fdlibm kernels).
// What pins it instead: hand-computed known-answer cases and analytic identities in
// tests/test_oracle_*.py, an independent pure-Python restatement (oracle/pyoracle.py) and
// scipy.sparse as a structural cross-check.
//
// Every function cites the reference file:line it follows (paths relative to
// /root/reference).  Arrays are Julia column-major: linear index L = i + nx*(j-1) +
// nx*ny*(k-1), 1-based in comments, 0-based in the code.  Build with
// -O2 -ffp-contract=off so that no FMA contraction changes a rounding.

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

namespace {

typedef int64_t i64;

enum {
    ORC_OK = 0,
    ORC_ERR_TADV_NAN = 1,      // "Tadv contains NaNs."     src/matrixbuilding.jl:39
    ORC_ERR_TKH_NAN = 2,       // "TκH contains NaNs."      src/matrixbuilding.jl:61
    ORC_ERR_TKVML_NAN = 3,     // "TκVML contains NaNs."    src/matrixbuilding.jl:90
    ORC_ERR_TKVDEEP_NAN = 4,   // "TκVdeep contains NaNs."  src/matrixbuilding.jl:114
    ORC_ERR_RHO_NAN = 5,       // "ρ contains NaNs"         src/matrixbuilding.jl:233
    ORC_ERR_UNKNOWN_GRID = 6,  // "Unknown grid type"       src/gridtopology.jl:111-116
    ORC_ERR_ALL_FILL = 7,      // @assert                   src/velocities.jl:199-200
    ORC_ERR_DRY_NEIGHBOUR = 8, // MethodError pushing `missing`/indexing `nothing`, src/matrixbuilding.jl:247-250
    ORC_ERR_BADARG = 9,
};

enum { TOPO_BIPOLAR = 0, TOPO_TRIPOLAR = 1, TOPO_UNKNOWN = 2 };
enum { D_SOUTH = 0, D_EAST = 1, D_NORTH = 2, D_WEST = 3 };  // dirs, src/gridcellgeometry.jl:304

struct Topo {
    int kind;
    i64 nx, ny, nz;
};

// Union{CartesianIndex{3}, Nothing}
struct CI {
    bool some;
    i64 i, j, k;  // 1-based
};
inline CI none() { return CI{false, 0, 0, 0}; }
inline CI ci(i64 i, i64 j, i64 k) { return CI{true, i, j, k}; }

// src/gridtopology.jl:57-68, 94
inline CI ip1(CI c, const Topo& g) { return c.i < g.nx ? ci(c.i + 1, c.j, c.k) : ci(1, c.j, c.k); }
inline CI im1(CI c, const Topo& g) { return c.i > 1 ? ci(c.i - 1, c.j, c.k) : ci(g.nx, c.j, c.k); }
inline CI jp1(CI c, const Topo& g) {
    if (c.j < g.ny) return ci(c.i, c.j + 1, c.k);
    if (g.kind == TOPO_TRIPOLAR) return ci(g.nx - c.i + 1, g.ny, c.k);
    return none();
}
inline CI jm1(CI c, const Topo&) { return c.j > 1 ? ci(c.i, c.j - 1, c.k) : none(); }
inline CI kp1(CI c, const Topo& g) { return c.k < g.nz ? ci(c.i, c.j, c.k + 1) : none(); }
inline CI km1(CI c, const Topo&) { return c.k > 1 ? ci(c.i, c.j, c.k - 1) : none(); }

inline i64 lin(const CI& c, const Topo& g) { return (c.i - 1) + g.nx * ((c.j - 1) + g.ny * (c.k - 1)); }
inline i64 lin2(const CI& c, const Topo& g) { return (c.i - 1) + g.nx * (c.j - 1); }
inline CI cart(i64 L, const Topo& g) {  // C[L], 0-based L
    i64 i = L % g.nx, r = L / g.nx;
    return ci(i + 1, r % g.ny + 1, r / g.ny + 1);
}

// ---------------------------------------------------------------------------------------
// Julia Base sind / cosd (base/special/trig.jl), restated: exact reduction in degrees,
// deg2rad_ext(x) = mulpi_ext(x/180) in double-double, fdlibm-derived kernels.
// ---------------------------------------------------------------------------------------
struct DD {
    double hi, lo;
};
inline DD mulpi_ext(double x) {
    const double m = 3.141592653589793, m_hi = 3.1415926218032837, m_lo = 3.178650954705639e-8;
    uint64_t u;
    std::memcpy(&u, &x, 8);
    u &= 0xfffffffff8000000ull;
    double x_hi;
    std::memcpy(&x_hi, &u, 8);
    double x_lo = x - x_hi;
    double y_hi = m * x;
    double y_lo = x_hi * m_lo + (x_lo * m_hi + ((x_hi * m_hi - y_hi) + x_lo * m_lo));
    return DD{y_hi, y_lo};
}
inline DD deg2rad_ext(double x) { return mulpi_ext(x / 180.0); }
inline double sin_kernel(DD y) {
    const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
                 S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
                 S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
    double z = y.hi * y.hi, w = z * z;
    double r = (S2 + z * (S3 + z * S4)) + z * w * (S5 + z * S6);
    double v = z * y.hi;
    return y.hi - ((z * (0.5 * y.lo - v * r) - y.lo) - v * S1);
}
inline double cos_kernel(DD y) {
    const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
                 C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
                 C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
    double z = y.hi * y.hi, w = z * z;
    double r = z * (C1 + z * (C2 + z * C3)) + w * w * (C4 + z * (C5 + z * C6));
    double hz = 0.5 * z;
    double ww = 1.0 - hz;
    return ww + (((1.0 - ww) - hz) + (z * r - y.hi * y.lo));
}
double sind(double x) {
    if (std::isnan(x) || std::isinf(x)) return std::numeric_limits<double>::quiet_NaN();
    double rx = std::copysign(std::fmod(x, 360.0), x);
    double arx = std::fabs(rx);
    if (rx == 0.0) return rx;
    if (arx < 45.0) return sin_kernel(deg2rad_ext(rx));
    if (arx <= 135.0) return std::copysign(cos_kernel(deg2rad_ext(90.0 - arx)), rx);
    if (arx == 180.0) return std::copysign(0.0, rx);
    if (arx < 225.0) return sin_kernel(deg2rad_ext((180.0 - arx) * (rx > 0 ? 1.0 : -1.0)));
    if (arx <= 315.0) return -std::copysign(cos_kernel(deg2rad_ext(270.0 - arx)), rx);
    return sin_kernel(deg2rad_ext(rx - std::copysign(360.0, rx)));
}
double cosd(double x) {
    if (std::isnan(x) || std::isinf(x)) return std::numeric_limits<double>::quiet_NaN();
    double rx = std::fabs(std::fmod(x, 360.0));
    if (rx <= 45.0) return cos_kernel(deg2rad_ext(rx));
    if (rx < 135.0) return sin_kernel(deg2rad_ext(90.0 - rx));
    if (rx <= 225.0) return -cos_kernel(deg2rad_ext(180.0 - rx));
    if (rx < 315.0) return sin_kernel(deg2rad_ext(rx - 270.0));
    return cos_kernel(deg2rad_ext(360.0 - rx));
}

// Julia `min` for Float64: NaN-propagating, -0.0 < 0.0
inline double jl_min(double x, double y) {
    if (std::isnan(x) || std::isnan(y)) return x + y;
    double diff = x - y;
    return std::signbit(diff) ? x : y;
}
inline double jl_max(double x, double y) {
    if (std::isnan(x) || std::isnan(y)) return x + y;
    double diff = x - y;
    return std::signbit(diff) ? y : x;
}

// Distances.jl Haversine (v0.10), points (lon, lat) in degrees, radius 6_371_000
double haversine(double lon1, double lat1, double lon2, double lat2) {
    double dl = lon2 - lon1, dp = lat2 - lat1;
    double s1 = sind(dp / 2), s2 = sind(dl / 2);
    double a = s1 * s1 + cosd(lat1) * cosd(lat2) * (s2 * s2);
    return 2 * (6371000.0 * std::asin(jl_min(std::sqrt(a), 1.0)));
}

// src/gridcellgeometry.jl:249-255
inline void midpointonsphere(double alon, double alat, double blon, double blat, double& mlon, double& mlat) {
    if (std::fabs(alon - blon) < 180) {
        mlon = (alon + blon) / 2;
        mlat = (alat + blat) / 2;
    } else {
        mlon = (alon + blon) / 2 + 180;
        mlat = (alat + blat) / 2 + 0;
    }
}
// src/gridcellgeometry.jl:209-215 (vertex numbers 1-based in the reference)
inline void vertexindices(int dir, int& a, int& b) {
    switch (dir) {
        case D_SOUTH: a = 0; b = 1; break;
        case D_EAST: a = 1; b = 2; break;
        case D_NORTH: a = 2; b = 3; break;
        default: a = 0; b = 3; break;  // west = (1, 4)
    }
}

struct CSC {
    i64 n = 0;
    std::vector<i64> colptr, rowval;  // 1-based like Julia
    std::vector<double> nzval;
};

// ---------------------------------------------------------------------------------------
// SparseArrays.sparse(I, J, V, m, n) with combine = + (stdlib sparse!), restated.
// Called at src/matrixbuilding.jl:41,63,92,116.
// ---------------------------------------------------------------------------------------
void jl_sparse(const std::vector<i64>& I, const std::vector<i64>& J, const std::vector<double>& V, i64 m, i64 n,
               CSC& out) {
    const i64 coolen = (i64)I.size();
    std::vector<i64> csrrowptr(m + 2, 0), csrcolval(coolen), klasttouch(n + 1, 0), csccolptr(n + 2, 0);
    std::vector<double> csrnzval(coolen);
    // row counts shifted forward by one
    for (i64 k = 0; k < coolen; ++k) csrrowptr[I[k] + 1] += 1;  // csrrowptr[Ik+1] (1-based arrays kept 1-based here)
    i64 countsum = 1;
    csrrowptr[1] = 1;
    for (i64 i = 2; i <= m + 1; ++i) {
        i64 overwritten = csrrowptr[i];
        csrrowptr[i] = countsum;
        countsum += overwritten;
    }
    // stable counting sort of (J, V) by row
    for (i64 k = 0; k < coolen; ++k) {
        i64 Ik = I[k];
        i64 csrk = csrrowptr[Ik + 1];
        csrrowptr[Ik + 1] = csrk + 1;
        csrcolval[csrk - 1] = J[k];
        csrnzval[csrk - 1] = V[k];
    }
    // sweep rows, combine repeats into the first occurrence, count columns
    i64 writek = 1, newcsrrowptri = 1, origcsrrowptri = 1, origcsrrowptrip1 = csrrowptr[2];
    for (i64 i = 1; i <= m; ++i) {
        for (i64 readk = origcsrrowptri; readk <= origcsrrowptrip1 - 1; ++readk) {
            i64 j = csrcolval[readk - 1];
            if (klasttouch[j] < newcsrrowptri) {
                klasttouch[j] = writek;
                if (writek != readk) {
                    csrcolval[writek - 1] = j;
                    csrnzval[writek - 1] = csrnzval[readk - 1];
                }
                writek += 1;
                csccolptr[j + 1] += 1;
            } else {
                i64 klt = klasttouch[j];
                csrnzval[klt - 1] = csrnzval[klt - 1] + csrnzval[readk - 1];
            }
        }
        newcsrrowptri = writek;
        origcsrrowptri = origcsrrowptrip1;
        if (origcsrrowptrip1 != writek) csrrowptr[i + 1] = writek;
        if (i < m) origcsrrowptrip1 = csrrowptr[i + 2];
    }
    // column pointers shifted forward by one, then counting sort rows into columns
    countsum = 1;
    csccolptr[1] = 1;
    for (i64 j = 2; j <= n + 1; ++j) {
        i64 overwritten = csccolptr[j];
        csccolptr[j] = countsum;
        countsum += overwritten;
    }
    const i64 nnz = writek - 1;
    out.n = n;
    out.rowval.assign(nnz, 0);
    out.nzval.assign(nnz, 0.0);
    for (i64 i = 1; i <= m; ++i) {
        for (i64 csrk = csrrowptr[i]; csrk <= csrrowptr[i + 1] - 1; ++csrk) {
            i64 j = csrcolval[csrk - 1];
            double x = csrnzval[csrk - 1];
            i64 csck = csccolptr[j + 1];
            csccolptr[j + 1] = csck + 1;
            out.rowval[csck - 1] = i;
            out.nzval[csck - 1] = x;
        }
    }
    out.colptr.assign(n + 1, 0);
    for (i64 j = 1; j <= n + 1; ++j) out.colptr[j - 1] = csccolptr[j];
}

// sparse A + B = map(+, A, B) -> _map_zeropres! (stdlib higherorderfns.jl), restated.
// Used by T = Tadv + TκH + TκVML + TκVdeep, src/matrixbuilding.jl:147 (n-ary + folds left).
void jl_spadd(const CSC& A, const CSC& B, CSC& C) {
    const i64 n = A.n;
    C.n = n;
    C.colptr.assign(n + 1, 0);
    C.rowval.clear();
    C.nzval.clear();
    C.rowval.reserve(A.rowval.size() + B.rowval.size());
    C.nzval.reserve(A.rowval.size() + B.rowval.size());
    const i64 sentinel = INT64_MAX;   // larger than any row index (the operands may be N x n column blocks)
    i64 Ck = 1;
    for (i64 j = 1; j <= n; ++j) {
        C.colptr[j - 1] = Ck;
        i64 Ak = A.colptr[j - 1], stopAk = A.colptr[j];
        i64 Bk = B.colptr[j - 1], stopBk = B.colptr[j];
        i64 Ai = Ak < stopAk ? A.rowval[Ak - 1] : sentinel;
        i64 Bi = Bk < stopBk ? B.rowval[Bk - 1] : sentinel;
        while (true) {
            double Cx;
            i64 Ci;
            if (Ai == Bi) {
                if (Ai == sentinel) break;
                Cx = A.nzval[Ak - 1] + B.nzval[Bk - 1];
                Ci = Ai;
                ++Ak; Ai = Ak < stopAk ? A.rowval[Ak - 1] : sentinel;
                ++Bk; Bi = Bk < stopBk ? B.rowval[Bk - 1] : sentinel;
            } else if (Ai < Bi) {
                Cx = A.nzval[Ak - 1] + 0.0;
                Ci = Ai;
                ++Ak; Ai = Ak < stopAk ? A.rowval[Ak - 1] : sentinel;
            } else {
                Cx = 0.0 + B.nzval[Bk - 1];
                Ci = Bi;
                ++Bk; Bi = Bk < stopBk ? B.rowval[Bk - 1] : sentinel;
            }
            if (!(Cx == 0.0)) {  // !_iszero(Cx); NaN is stored
                C.rowval.push_back(Ci);
                C.nzval.push_back(Cx);
                ++Ck;
            }
        }
    }
    C.colptr[n] = Ck;
}

struct Triplets {
    std::vector<i64> I, J;
    std::vector<double> V;
    // column window of the streamed build (orc_tm_build_columns): only triplets of the columns jlo..jhi (1-based) are
    // kept, renumbered from 1.  The order of the kept triplets is the emit order, so sparse() combines them exactly as
    // it does inside the full matrix.  Default: keep everything.
    i64 jlo = 1, jhi = INT64_MAX;
    void hint(i64 n) {  // preallocate_sparse_entries, src/matrixbuilding.jl:153-161
        I.reserve(n);
        J.reserve(n);
        V.reserve(n);
    }
    void push(i64 i, i64 j, double v) {
        if (j < jlo || j > jhi) return;
        I.push_back(i);
        J.push_back(j - jlo + 1);
        V.push_back(v);
    }
    bool anynan() const {
        for (double v : V)
            if (std::isnan(v)) return true;
        return false;
    }
};

struct Grid {
    Topo g;
    const double* v3D;
    const double* thk;
    const double* area2D;
    const double* zt;
    const double* edge;  // 4 x P, order south,east,north,west
    const double* dnbr;  // 4 x P
    i64 P() const { return g.nx * g.ny; }
    i64 M() const { return g.nx * g.ny * g.nz; }
};

struct Indices {  // makeindices, src/matrixbuilding.jl:10-24
    std::vector<i64> Lwet;      // 0-based linear indices here
    std::vector<i64> Lwet3D;    // 1-based wet index, 0 = missing
    std::vector<uint8_t> wet3D;
    i64 N = 0;
    i64 w_lo = 0, w_hi = 0;     // emitters visited (all of them, or those that can reach a column window)
};

void makeindices(const double* v3D, i64 M, Indices& ix) {
    ix.Lwet.clear();
    ix.Lwet3D.assign(M, 0);
    ix.wet3D.assign(M, 0);
    for (i64 L = 0; L < M; ++L)
        if (!std::isnan(v3D[L])) ix.Lwet.push_back(L);
    ix.N = (i64)ix.Lwet.size();
    ix.w_lo = 0;
    ix.w_hi = ix.N;
    for (i64 w = 0; w < ix.N; ++w) {
        ix.wet3D[ix.Lwet[w]] = 1;
        ix.Lwet3D[ix.Lwet[w]] = w + 1;
    }
}

// pushTadvectionvalues!, src/matrixbuilding.jl:193-204
inline void pushTadvectionvalues(Triplets& t, i64 wi, i64 wj, double phi, double rhoi, double rhoj, double vi,
                                 double vj) {
    double rho = (rhoi + rhoj) / 2;
    double mi = rho * vi;
    double mj = rho * vj;
    t.push(wi, wj, -phi / mi);
    t.push(wj, wj, phi / mj);
}
// pushTmixingvalues!, src/matrixbuilding.jl:426-435
inline void pushTmixingvalues(Triplets& t, i64 wi, i64 wj, double kappa, double a, double d, double V) {
    double Tval = kappa * a / (d * V);
    t.push(wi, wi, Tval);
    t.push(wi, wj, -Tval);
}

// advection_operator_sparse_entries, src/matrixbuilding.jl:221-299
int advection_entries(const double* const phi[6] /*east,west,north,south,top,bottom*/, const Grid& G,
                      const Indices& ix, const double* rho3d, double rho_scalar, bool upwind, Triplets& t) {
    const Topo& g = G.g;
    if (g.kind == TOPO_UNKNOWN) return ORC_ERR_UNKNOWN_GRID;
    const double *pe = phi[0], *pw = phi[1], *pn = phi[2], *ps = phi[3], *pt = phi[4], *pb = phi[5];
    auto rho_at = [&](i64 L) { return rho3d ? rho3d[L] : rho_scalar; };  // scalar -> fill, :221-225
    for (i64 w = 0; w < ix.N; ++w)
        if (std::isnan(rho_at(ix.Lwet[w]))) return ORC_ERR_RHO_NAN;  // :233
    t.hint(6 * ix.N);
    int status = ORC_OK;
    auto emit = [&](i64 wi, CI Cj, double p, double rhoi, double vi) {
        if (!Cj.some) { status = ORC_ERR_DRY_NEIGHBOUR; return; }
        i64 Lj = lin(Cj, g);
        i64 wj = ix.Lwet3D[Lj];
        if (wj == 0) { status = ORC_ERR_DRY_NEIGHBOUR; return; }
        pushTadvectionvalues(t, wi, wj, p, rhoi, rho_at(Lj), vi, G.v3D[Lj]);
    };
    for (i64 w = ix.w_lo; w < ix.w_hi; ++w) {
        i64 Li = ix.Lwet[w];
        CI Ci = cart(Li, g);
        i64 wi = w + 1;
        double vi = G.v3D[Li], rhoi = rho_at(Li);
        double f;
        f = upwind ? jl_max(pw[Li], 0.0) : pw[Li] / 2;            // From West  :244
        if (f > 0 || f < 0) emit(wi, im1(Ci, g), f, rhoi, vi);
        f = upwind ? jl_min(pe[Li], 0.0) : pe[Li] / 2;            // From East  :253
        if (f > 0 || f < 0) emit(wi, ip1(Ci, g), -f, rhoi, vi);
        f = upwind ? jl_max(ps[Li], 0.0) : ps[Li] / 2;            // From South :262
        if (f > 0 || f < 0) emit(wi, jm1(Ci, g), f, rhoi, vi);
        f = upwind ? jl_min(pn[Li], 0.0) : pn[Li] / 2;            // From North :271
        if (f > 0 || f < 0) emit(wi, jp1(Ci, g), -f, rhoi, vi);
        f = upwind ? jl_max(pb[Li], 0.0) : pb[Li] / 2;            // From Bottom :280
        if (f > 0 || f < 0) emit(wi, kp1(Ci, g), f, rhoi, vi);
        f = upwind ? jl_min(pt[Li], 0.0) : pt[Li] / 2;            // From Top   :289
        if (Ci.k > 1 && (f > 0 || f < 0)) emit(wi, km1(Ci, g), -f, rhoi, vi);
        if (status != ORC_OK) return status;
    }
    return ORC_OK;
}

// verticalfacearea(edge_length_2D, thkcello, i, j, k, dir), src/gridcellgeometry.jl:230-234
inline double verticalfacearea(const Grid& G, const CI& c, int dir) {
    double height = G.thk[lin(c, G.g)];
    double width = G.edge[dir * G.P() + lin2(c, G.g)];
    return height * width;
}

// horizontal_diffusion_operator_sparse_entries, src/matrixbuilding.jl:337-418 (ΩH = trues(N), :56)
int hdiff_entries(const Grid& G, const Indices& ix, double kH, Triplets& t) {
    const Topo& g = G.g;
    if (g.kind == TOPO_UNKNOWN) return ORC_ERR_UNKNOWN_GRID;
    t.hint(8 * ix.N);
    const i64 P = G.P();
    for (i64 w = ix.w_lo; w < ix.w_hi; ++w) {
        i64 Li = ix.Lwet[w];
        CI Ci = cart(Li, g);
        i64 wi = w + 1;
        i64 srf = lin2(Ci, g);
        double V = G.v3D[Li];
        auto side = [&](CI Cj, int dir, int opp) {
            if (!Cj.some) return;
            i64 wj = ix.Lwet3D[lin(Cj, g)];
            if (wj == 0) return;
            double aij = verticalfacearea(G, Ci, dir);
            double aji = verticalfacearea(G, Cj, opp);
            double a = jl_min(aij, aji);
            double d = G.dnbr[dir * P + srf];
            pushTmixingvalues(t, wi, wj, kH, a, d, V);
        };
        side(im1(Ci, g), D_WEST, D_EAST);                              // :356-369
        side(ip1(Ci, g), D_EAST, D_WEST);                              // :371-383
        side(jm1(Ci, g), D_SOUTH, D_NORTH);                            // :385-397
        side(jp1(Ci, g), D_NORTH, Ci.j == g.ny ? D_NORTH : D_SOUTH);   // :399-414, oppdir :407
    }
    return ORC_OK;
}

// vertical_diffusion_operator_sparse_entries, src/matrixbuilding.jl:438-479
int vdiff_entries(const Grid& G, const Indices& ix, double kV, const std::vector<uint8_t>& Omega, Triplets& t) {
    const Topo& g = G.g;
    if (g.kind == TOPO_UNKNOWN) return ORC_ERR_UNKNOWN_GRID;
    t.hint(4 * ix.N);
    for (i64 w = ix.w_lo; w < ix.w_hi; ++w) {
        if (!Omega[w]) continue;
        i64 Li = ix.Lwet[w];
        CI Ci = cart(Li, g);
        i64 wi = w + 1;
        double V = G.v3D[Li];
        double a = G.area2D[lin2(Ci, g)];
        auto side = [&](CI Cj) {
            if (!Cj.some) return;
            i64 wj = ix.Lwet3D[lin(Cj, g)];
            if (wj == 0 || !Omega[wj - 1]) return;
            double d = std::fabs(G.zt[Ci.k - 1] - G.zt[Cj.k - 1]);
            pushTmixingvalues(t, wi, wj, kV, a, d, V);
        };
        side(kp1(Ci, g));  // From Bottom :458-466
        side(km1(Ci, g));  // From Top    :468-476
    }
    return ORC_OK;
}

struct TM {
    CSC mats[5];  // T, Tadv, TkH, TkVML, TkVdeep
    Triplets trip[4];
    double seconds = 0.0;
    int status = 0;
};

}  // namespace

extern "C" {

// makeindices, src/matrixbuilding.jl:10-24.  wet_chunks: Julia BitArray chunk layout
// (bit b of chunk c <-> linear index 64c+b+1); Lwet3D: 1-based wet index, 0 = missing;
// Lwet: 1-based linear indices (caller passes room for M entries).
int orc_makeindices(const double* v3D, i64 nx, i64 ny, i64 nz, uint64_t* wet_chunks, i64* Lwet3D, i64* Lwet, i64* N) {
    const i64 M = nx * ny * nz;
    Indices ix;
    makeindices(v3D, M, ix);
    *N = ix.N;
    for (i64 c = 0; c < (M + 63) / 64; ++c) wet_chunks[c] = 0;
    for (i64 L = 0; L < M; ++L) {
        Lwet3D[L] = ix.Lwet3D[L];
        if (ix.wet3D[L]) wet_chunks[L >> 6] |= (1ull << (L & 63));
    }
    for (i64 w = 0; w < ix.N; ++w) Lwet[w] = ix.Lwet[w] + 1;
    return ORC_OK;
}

// makegridmetrics numerics, src/gridcellgeometry.jl:283-285 and :304-308.
// Inputs are already NaN-cleaned and vertex-permuted (host shim).  edge/dedge/dnbr: 4 x P
// in the reference's `dirs` order south, east, north, west.
int orc_gridmetrics(const double* area2D, const double* v3D, const double* lon, const double* lat,
                    const double* lonv, const double* latv, i64 nx, i64 ny, i64 nz, int topo, double* thk,
                    double* Z3D, double* edge, double* dedge, double* dnbr) {
    Topo g{topo, nx, ny, nz};
    const i64 P = nx * ny;
    // thkcello = v3D ./ area2D ; ZBOT3D = cumsum(thkcello, dims=3) ; Z3D = ZBOT3D - 0.5*thkcello
    for (i64 p = 0; p < P; ++p) {
        double zbot = 0.0;
        for (i64 k = 0; k < nz; ++k) {
            double t = v3D[p + P * k] / area2D[p];
            thk[p + P * k] = t;
            zbot = (k == 0) ? t : zbot + t;
            Z3D[p + P * k] = zbot - 0.5 * t;
        }
    }
    if (topo == TOPO_UNKNOWN) return ORC_ERR_UNKNOWN_GRID;  // j₊₁ etc. error, src/gridtopology.jl:111-116
    for (i64 j = 1; j <= ny; ++j)
        for (i64 i = 1; i <= nx; ++i) {
            i64 p = (i - 1) + nx * (j - 1);
            const double* vl = lonv + 4 * p;
            const double* vt = latv + 4 * p;
            for (int d = 0; d < 4; ++d) {
                int a, b;
                vertexindices(d, a, b);
                edge[d * P + p] = haversine(vl[a], vt[a], vl[b], vt[b]);  // verticalfacewidth :217-222
                double mlon, mlat;
                midpointonsphere(vl[a], vt[a], vl[b], vt[b], mlon, mlat);
                dedge[d * P + p] = haversine(lon[p], lat[p], mlon, mlat);  // centroid2edgedistance :240-247
            }
            CI c = ci(i, j, 1);
            CI nb[4] = {jm1(c, g), ip1(c, g), jp1(c, g), im1(c, g)};      // 𝑗s = (j₋₁, i₊₁, j₊₁, i₋₁) :305
            for (int d = 0; d < 4; ++d) {
                if (!nb[d].some) {
                    dnbr[d * P + p] = std::numeric_limits<double>::quiet_NaN();  // :189
                } else {
                    i64 q = lin2(nb[d], g);
                    dnbr[d * P + p] = haversine(lon[p], lat[p], lon[q], lat[q]);  // horizontaldistance :182-188
                }
            }
        }
    return ORC_OK;
}

double orc_haversine(double lon1, double lat1, double lon2, double lat2) { return haversine(lon1, lat1, lon2, lat2); }
double orc_sind(double x) { return sind(x); }
double orc_cosd(double x) { return cosd(x); }

// nofluxboundaries! + facefluxes, src/velocities.jl:154-255.  umo/vmo are modified in
// place exactly as the reference does.  Outputs: six M-sized arrays.
int orc_facefluxes(double* umo, double* vmo, const double* v3D, i64 nx, i64 ny, i64 nz, int topo, double fill,
                   double* east, double* west, double* north, double* south, double* top, double* bottom) {
    Topo g{topo, nx, ny, nz};
    if (topo == TOPO_UNKNOWN) return ORC_ERR_UNKNOWN_GRID;
    const i64 M = nx * ny * nz, P = nx * ny;
    std::vector<uint8_t> wet(M);
    for (i64 L = 0; L < M; ++L) wet[L] = !std::isnan(v3D[L]);
    // nofluxboundaries! :161-175
    for (i64 L = 0; L < M; ++L) {
        CI c = cart(L, g);
        CI E = ip1(c, g), N = jp1(c, g);
        if (!wet[L]) umo[L] = vmo[L] = 0;
        if (!E.some || !wet[lin(E, g)]) umo[L] = 0;
        if (!N.some || !wet[lin(N, g)]) vmo[L] = 0;
    }
    // :199-200
    bool allu = true, allv = true;
    for (i64 L = 0; L < M; ++L) {
        if (!(std::isnan(umo[L]) || umo[L] == fill)) allu = false;
        if (!(std::isnan(vmo[L]) || vmo[L] == fill)) allv = false;
    }
    if (allu || allv) return ORC_ERR_ALL_FILL;
    // :203, :215  replace(NaN => 0.0, FillValue => 0.0)
    for (i64 L = 0; L < M; ++L) {
        east[L] = (std::isnan(umo[L]) || umo[L] == fill) ? 0.0 : umo[L];
        north[L] = (std::isnan(vmo[L]) || vmo[L] == fill) ? 0.0 : vmo[L];
    }
    // :206-211, :219-224
    for (i64 L = 0; L < M; ++L) {
        CI c = cart(L, g);
        CI W = im1(c, g);
        west[L] = W.some ? east[lin(W, g)] : 0.0;
        CI S = jm1(c, g);
        south[L] = S.some ? north[lin(S, g)] : 0.0;
    }
    // :236-243
    for (i64 k = nz - 1; k >= 0; --k)
        for (i64 p = 0; p < P; ++p) {
            i64 L = p + P * k;
            bottom[L] = (k == nz - 1) ? 0.0 : top[L + P];
            top[L] = bottom[L] + west[L] + south[L] - east[L] - north[L];
        }
    return ORC_OK;
}

// transportmatrix, src/matrixbuilding.jl:128-150 (all four operators built, :140-143).
// phi order: east, west, north, south, top, bottom (the NamedTuple of src/velocities.jl:245-252).
// rho3d may be NULL (then rho_scalar is used, :221-225).  mlotst: NaN = missing.
// Returns an opaque handle; status via orc_tm_status.
// col_lo / col_hi (0-based, [col_lo, col_hi)): the STREAMED form used for grids whose full matrices do not fit the
// checker's memory comfortably — the same emitters visit only the wet cells that can reach those columns (their
// neighbours: within one level, nx*ny linear cells, either side), only the triplets of those columns are kept, in
// emit order, and sparse / + run on an N x (col_hi - col_lo) matrix: the result is exactly the column block of the
// full matrices (colptr local and 1-based, row indices global).  col_hi < 0: everything.
void* orc_tm_build_columns(const double* pe, const double* pw, const double* pn, const double* ps, const double* pt,
                           const double* pb, const double* mlotst, const double* v3D, const double* thk,
                           const double* area2D, const double* zt, const double* edge, const double* dnbr, i64 nx, i64 ny,
                           i64 nz, int topo, const double* rho3d, double rho_scalar, double kH, double kVML, double kVdeep,
                           int upwind, int keep_triplets, i64 col_lo, i64 col_hi) {
    TM* tm = new TM();
    auto t0 = std::chrono::steady_clock::now();
    Grid G{Topo{topo, nx, ny, nz}, v3D, thk, area2D, zt, edge, dnbr};
    Indices ix;
    makeindices(v3D, G.M(), ix);
    const i64 Nrows = ix.N;
    i64 N = ix.N;               // columns of the matrices built below
    const double* phi[6] = {pe, pw, pn, ps, pt, pb};
    Triplets tr[4];
    if (col_hi >= 0) {
        if (col_lo < 0 || col_hi > ix.N || col_lo >= col_hi) { tm->status = -1; return tm; }
        const i64 L_lo = ix.Lwet[col_lo] - G.P(), L_hi = ix.Lwet[col_hi - 1] + G.P();
        ix.w_lo = std::lower_bound(ix.Lwet.begin(), ix.Lwet.end(), L_lo) - ix.Lwet.begin();
        ix.w_hi = std::upper_bound(ix.Lwet.begin(), ix.Lwet.end(), L_hi) - ix.Lwet.begin();
        for (auto& t : tr) t.jlo = col_lo + 1, t.jhi = col_hi;
        N = col_hi - col_lo;
    }
    int st;
    // buildTadv :31-44
    st = advection_entries(phi, G, ix, rho3d, rho_scalar, upwind != 0, tr[0]);
    if (st == ORC_OK && tr[0].anynan()) st = ORC_ERR_TADV_NAN;
    if (st != ORC_OK) { tm->status = st; return tm; }
    jl_sparse(tr[0].I, tr[0].J, tr[0].V, Nrows, N, tm->mats[1]);
    // buildTκH :51-66
    st = hdiff_entries(G, ix, kH, tr[1]);
    if (st == ORC_OK && tr[1].anynan()) st = ORC_ERR_TKH_NAN;
    if (st != ORC_OK) { tm->status = st; return tm; }
    jl_sparse(tr[1].I, tr[1].J, tr[1].V, Nrows, N, tm->mats[2]);
    // buildTκVML :74-95 ; Ω = (zt[k] < mlotst[i,j], missing -> false)[Lwet]  :85
    std::vector<uint8_t> Omega(Nrows);
    for (i64 w = 0; w < Nrows; ++w) {
        CI c = cart(ix.Lwet[w], G.g);
        double ml = mlotst[lin2(c, G.g)];
        Omega[w] = (zt[c.k - 1] < ml) ? 1 : 0;  // comparison with NaN is false
    }
    st = vdiff_entries(G, ix, kVML, Omega, tr[2]);
    if (st == ORC_OK && tr[2].anynan()) st = ORC_ERR_TKVML_NAN;
    if (st != ORC_OK) { tm->status = st; return tm; }
    jl_sparse(tr[2].I, tr[2].J, tr[2].V, Nrows, N, tm->mats[3]);
    // buildTκVdeep :103-120 ; Ω = trues(N)
    std::fill(Omega.begin(), Omega.end(), 1);
    st = vdiff_entries(G, ix, kVdeep, Omega, tr[3]);
    if (st == ORC_OK && tr[3].anynan()) st = ORC_ERR_TKVDEEP_NAN;
    if (st != ORC_OK) { tm->status = st; return tm; }
    jl_sparse(tr[3].I, tr[3].J, tr[3].V, Nrows, N, tm->mats[4]);
    // T = Tadv + TκH + TκVML + TκVdeep :147
    CSC t1, t2;
    jl_spadd(tm->mats[1], tm->mats[2], t1);
    jl_spadd(t1, tm->mats[3], t2);
    jl_spadd(t2, tm->mats[4], tm->mats[0]);
    tm->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (keep_triplets)
        for (int q = 0; q < 4; ++q) tm->trip[q] = std::move(tr[q]);
    return tm;
}
void* orc_tm_build(const double* pe, const double* pw, const double* pn, const double* ps, const double* pt,
                   const double* pb, const double* mlotst, const double* v3D, const double* thk,
                   const double* area2D, const double* zt, const double* edge, const double* dnbr, i64 nx, i64 ny,
                   i64 nz, int topo, const double* rho3d, double rho_scalar, double kH, double kVML, double kVdeep,
                   int upwind, int keep_triplets) {
    return orc_tm_build_columns(pe, pw, pn, ps, pt, pb, mlotst, v3D, thk, area2D, zt, edge, dnbr, nx, ny, nz, topo, rho3d,
                                rho_scalar, kH, kVML, kVdeep, upwind, keep_triplets, 0, -1);
}
int orc_tm_status(void* h) { return ((TM*)h)->status; }
double orc_tm_seconds(void* h) { return ((TM*)h)->seconds; }
i64 orc_tm_n(void* h, int which) { return ((TM*)h)->mats[which].n; }
i64 orc_tm_nnz(void* h, int which) { return (i64)((TM*)h)->mats[which].rowval.size(); }
void orc_tm_fetch(void* h, int which, i64* colptr, i64* rowval, double* nzval) {
    const CSC& m = ((TM*)h)->mats[which];
    std::copy(m.colptr.begin(), m.colptr.end(), colptr);
    std::copy(m.rowval.begin(), m.rowval.end(), rowval);
    std::copy(m.nzval.begin(), m.nzval.end(), nzval);
}
i64 orc_tm_ntriplets(void* h, int op) { return (i64)((TM*)h)->trip[op].I.size(); }
void orc_tm_triplets(void* h, int op, i64* I, i64* J, double* V) {
    const Triplets& t = ((TM*)h)->trip[op];
    std::copy(t.I.begin(), t.I.end(), I);
    std::copy(t.J.begin(), t.J.end(), J);
    std::copy(t.V.begin(), t.V.end(), V);
}
void orc_tm_free(void* h) { delete (TM*)h; }

// bare sparse(I,J,V,n,n) and A+B for the COO->CSC / add parity tests
void* orc_sparse(const i64* I, const i64* J, const double* V, i64 len, i64 n) {
    TM* tm = new TM();
    std::vector<i64> vi(I, I + len), vj(J, J + len);
    std::vector<double> vv(V, V + len);
    jl_sparse(vi, vj, vv, n, n, tm->mats[0]);
    return tm;
}
void* orc_spadd(i64 n, const i64* acp, const i64* arv, const double* anz, const i64* bcp, const i64* brv,
                const double* bnz) {
    TM* tm = new TM();
    CSC A, B;
    A.n = B.n = n;
    A.colptr.assign(acp, acp + n + 1);
    B.colptr.assign(bcp, bcp + n + 1);
    A.rowval.assign(arv, arv + (acp[n] - 1));
    A.nzval.assign(anz, anz + (acp[n] - 1));
    B.rowval.assign(brv, brv + (bcp[n] - 1));
    B.nzval.assign(bnz, bnz + (bcp[n] - 1));
    jl_spadd(A, B, tm->mats[0]);
    return tm;
}

// ---------------------------------------------------------------------------------------
// Triads, dyads, bolus_GM_velocity (experimental in the reference; fields, not matrices)
// ---------------------------------------------------------------------------------------
static inline double getornan(const double* chi, const CI& c, const Topo& g) {  // getindexornan, src/gridtopology.jl:69
    return c.some ? chi[lin(c, g)] : std::numeric_limits<double>::quiet_NaN();
}
static inline double vdist(const double* Z, const CI& a, const CI& b, const Topo& g) {  // verticaldistance :195
    return std::fabs(getornan(Z, b, g) - getornan(Z, a, g));
}
static inline double hdist(const double* lon, const double* lat, const CI& a, const CI& b, const Topo& g) {  // :182-189
    if (!b.some) return std::numeric_limits<double>::quiet_NaN();
    i64 p = lin2(a, g), q = lin2(b, g);
    return haversine(lon[p], lat[p], lon[q], lat[q]);
}
// nan-mean with Bool weights: false*NaN == 0.0 in Julia (src/triads.jl:130-132, src/dyads.jl:63-64)
static inline double boolmul(bool w, double v) { return w ? v : 0.0; }

// globalverticalfacetriadderivative, src/triads.jl:84-146.  dir: 0 = Icoord, 1 = Jcoord.
// Returns ORC_ERR_DRY_NEIGHBOUR where the reference would throw (k₋₁(nothing, ...) on a
// bipolar grid at j = ny with Jcoord, src/triads.jl:87-89).
int orc_triad(const double* chi, const double* lon, const double* lat, const double* Z3D, const double* v3D, i64 nx,
              i64 ny, i64 nz, int topo, int dir, double* out) {
    Topo g{topo, nx, ny, nz};
    if (topo == TOPO_UNKNOWN) return ORC_ERR_UNKNOWN_GRID;
    const i64 M = nx * ny * nz;
    const double NaN = std::numeric_limits<double>::quiet_NaN();
    for (i64 L = 0; L < M; ++L) out[L] = NaN;
    for (i64 L = 0; L < M; ++L) {
        if (std::isnan(v3D[L])) continue;
        CI I = cart(L, g);
        CI N = km1(I, g), S = kp1(I, g);
        CI E = dir == 0 ? ip1(I, g) : jp1(I, g);
        if (!E.some) return ORC_ERR_DRY_NEIGHBOUR;
        CI NE = km1(E, g), SE = kp1(E, g);
        double vC = getornan(chi, I, g), vN = getornan(chi, N, g), vS = getornan(chi, S, g), vE = getornan(chi, E, g),
               vNE = getornan(chi, NE, g), vSE = getornan(chi, SE, g);
        double dCN = vdist(Z3D, I, N, g), dCS = vdist(Z3D, I, S, g), dCE = hdist(lon, lat, I, E, g),
               dENE = vdist(Z3D, E, NE, g), dESE = vdist(Z3D, E, SE, g);
        double CN = (vN - vC) / dCN, CS = (vC - vS) / dCS, CE = (vE - vC) / dCE, ENE = (vNE - vE) / dENE,
               ESE = (vE - vSE) / dESE;
        double r[4] = {CE / CN, CE / CS, CE / ENE, CE / ESE};
        bool w[4];
        for (int q = 0; q < 4; ++q) w[q] = !std::isnan(r[q]);
        double s = boolmul(w[0], r[0]);
        for (int q = 1; q < 4; ++q) s = s + boolmul(w[q], r[q]);
        int cnt = w[0] + w[1] + w[2] + w[3];
        out[L] = s / (double)cnt;
    }
    return ORC_OK;
}

// globalverticaldyadderivative, src/dyads.jl:38-78
int orc_dyad(const double* chi, const double* Z3D, const double* v3D, i64 nx, i64 ny, i64 nz, int topo, double* out) {
    Topo g{topo, nx, ny, nz};
    if (topo == TOPO_UNKNOWN) return ORC_ERR_UNKNOWN_GRID;
    const i64 M = nx * ny * nz;
    const double NaN = std::numeric_limits<double>::quiet_NaN();
    for (i64 L = 0; L < M; ++L) out[L] = NaN;
    for (i64 L = 0; L < M; ++L) {
        if (std::isnan(v3D[L])) continue;
        CI I = cart(L, g);
        CI N = km1(I, g), S = kp1(I, g);
        double vC = getornan(chi, I, g), vN = getornan(chi, N, g), vS = getornan(chi, S, g);
        double d0 = (vN - vC) / vdist(Z3D, I, N, g), d1 = (vC - vS) / vdist(Z3D, I, S, g);
        bool w0 = !std::isnan(d0), w1 = !std::isnan(d1);
        out[L] = (boolmul(w0, d0) + boolmul(w1, d1)) / (double)(w0 + w1);
    }
    return ORC_OK;
}

static inline double jl_clamp(double x, double lo, double hi) {  // clamp(x, lo, hi) = ifelse(x > hi, hi, ifelse(x < lo, lo, x))
    return x > hi ? hi : (x < lo ? lo : x);
}

// bolus_GM_velocity, src/RediGM.jl:46-79
int orc_bolus_gm(const double* rho, const double* lon, const double* lat, const double* Z3D, const double* v3D,
                 i64 nx, i64 ny, i64 nz, int topo, double kGM, double maxslope, double* u, double* v) {
    const i64 M = nx * ny * nz;
    std::vector<double> Si(M), Sj(M);
    int st = orc_triad(rho, lon, lat, Z3D, v3D, nx, ny, nz, topo, 0, Si.data());
    if (st) return st;
    st = orc_triad(rho, lon, lat, Z3D, v3D, nx, ny, nz, topo, 1, Sj.data());
    if (st) return st;
    const double Sc = 0.004, Sd = 0.001;
    for (i64 L = 0; L < M; ++L) {
        double a = jl_clamp(Si[L], -maxslope, maxslope), b = jl_clamp(Sj[L], -maxslope, maxslope);
        double taper = 0.5 * (1 + std::tanh((Sc - std::sqrt(a * a + b * b)) / Sd));
        Si[L] = kGM * (taper * a);
        Sj[L] = kGM * (taper * b);
    }
    st = orc_dyad(Si.data(), Z3D, v3D, nx, ny, nz, topo, u);
    if (st) return st;
    return orc_dyad(Sj.data(), Z3D, v3D, nx, ny, nz, topo, v);
}

}  // extern "C"
