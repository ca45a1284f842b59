/* abi_driver.c — the C ABI of libotmb.so exercised from plain C (dlopen, no Python, no ctypes).
 *
 *     gcc -O1 -o abi_driver tests/abi_driver.c -ldl -lm && ./abi_driver oceantransportmatrixbuilder.jl_b200/libotmb.so
 *
 * Resolves the entry points by name, then — on a machine with a B200 — builds a 4x2x2 all-wet tripolar box with unit
 * metrics through otmb_set_grid / otmb_makeindices / otmb_set_gridmetrics / otmb_set_facefluxes / otmb_set_mlotst /
 * otmb_transportmatrix_build / otmb_transportmatrix_fetch and checks the hand-derived answer of
 * tests/golden/known_answers.json ("kvdeep_4x2x2": TκVdeep = κ [[I, -I], [-I, I]], /root/reference/src/matrixbuilding.jl:450-477),
 * the empty Tadv of a resting ocean, the CSC invariants of every matrix and T = Tadv + TκH + TκVML + TκVdeep entry by
 * entry; the same matrices must come back from otmb_transportmatrix_stream.  Without a GPU otmb_create must fail with
 * OTMB_ERR_NO_GPU (there is no CPU fallback) and the driver says so and exits 0.
 * Prints "ABI-DRIVER: OK" / "ABI-DRIVER: no GPU" on success; any failure exits 1. */
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/otmb.h"

#define NX 4
#define NY 2
#define NZ 2
#define P (NX * NY)
#define M (NX * NY * NZ)

static void* lib;
static void* sym(const char* name) {
    void* p = dlsym(lib, name);
    if (!p) {
        fprintf(stderr, "ABI-DRIVER: %s is not exported\n", name);
        exit(1);
    }
    return p;
}
#define FAIL(...)                                 \
    do {                                          \
        fprintf(stderr, "ABI-DRIVER: " __VA_ARGS__); \
        fprintf(stderr, "\n");                    \
        exit(1);                                  \
    } while (0)

typedef int (*create_t)(otmb_ctx**, int);
typedef int (*destroy_t)(otmb_ctx*);
typedef const char* (*lasterr_t)(const otmb_ctx*);
typedef int (*setgrid_t)(otmb_ctx*, int64_t, int64_t, int64_t, int);
typedef int (*makeidx_t)(otmb_ctx*, const double*, int64_t*);
typedef int (*setgm_t)(otmb_ctx*, const double*, const double*, const double*, const double*, const double*, const double*,
                       const double*, const double*);
typedef int (*setphi_t)(otmb_ctx*, const double* const[6]);
typedef int (*setml_t)(otmb_ctx*, const double*);
typedef int (*build_t)(otmb_ctx*, const otmb_tm_params*, int64_t[5]);
typedef int (*fetch_t)(otmb_ctx*, int, int64_t*, int64_t*, double*);
typedef int (*stream_t)(otmb_ctx*, const otmb_tm_params*, const double* const[6], const double*, const double*, int32_t,
                        const int64_t[5], int64_t* const[5], int64_t* const[5], double* const[5], int64_t[5]);

static lasterr_t last_error;
static void ok(otmb_ctx* c, int st, const char* what) {
    if (st != OTMB_OK) FAIL("%s failed with status %d: %s", what, st, last_error(c));
}

int main(int argc, char** argv) {
    const char* path = argc > 1 ? argv[1] : "oceantransportmatrixbuilder.jl_b200/libotmb.so";
    lib = dlopen(path, RTLD_NOW);
    if (!lib) FAIL("cannot load %s: %s", path, dlerror());
    int (*version)(void) = (int (*)(void))sym("otmb_version");
    const char* (*status_string)(int) = (const char* (*)(int))sym("otmb_status_string");
    create_t create = (create_t)sym("otmb_create");
    destroy_t destroy = (destroy_t)sym("otmb_destroy");
    last_error = (lasterr_t)sym("otmb_last_error");
    setgrid_t set_grid = (setgrid_t)sym("otmb_set_grid");
    makeidx_t makeindices = (makeidx_t)sym("otmb_makeindices");
    setgm_t set_gridmetrics = (setgm_t)sym("otmb_set_gridmetrics");
    setphi_t set_facefluxes = (setphi_t)sym("otmb_set_facefluxes");
    setml_t set_mlotst = (setml_t)sym("otmb_set_mlotst");
    build_t build = (build_t)sym("otmb_transportmatrix_build");
    fetch_t fetch = (fetch_t)sym("otmb_transportmatrix_fetch");
    stream_t stream = (stream_t)sym("otmb_transportmatrix_stream");
    if (version() < 100) FAIL("otmb_version() = %d", version());

    otmb_ctx* c = NULL;
    int st = create(&c, 0);
    if (st == OTMB_ERR_NO_GPU) {
        if (c != NULL || !strstr(status_string(st), "no CPU fallback")) FAIL("a missing GPU must be a loud error");
        printf("ABI-DRIVER: no GPU (%s)\n", status_string(st));
        return 0;
    }
    if (st != OTMB_OK) FAIL("otmb_create: %s", status_string(st));

    /* the box: all wet, volume 1, area 1, thickness 1, edges 1, neighbour distances 1, zt = 0.5, 1.5, resting ocean */
    double v3D[M], thk[M], zero[M], area[P], ml[P], edge[4 * P], dnbr[4 * P], zt[NZ] = {0.5, 1.5};
    for (int i = 0; i < M; ++i) v3D[i] = thk[i] = 1.0, zero[i] = 0.0;
    for (int i = 0; i < P; ++i) area[i] = 1.0, ml[i] = 1.0; /* mixed layer 1 m: only level 1 (zt = 0.5) is inside */
    for (int i = 0; i < 4 * P; ++i) edge[i] = dnbr[i] = 1.0;
    const double kH = 2.0, kVML = 0.5, kVdeep = 0.25;
    int64_t N = 0;
    ok(c, set_grid(c, NX, NY, NZ, OTMB_TOPO_TRIPOLAR), "otmb_set_grid");
    ok(c, makeindices(c, v3D, &N), "otmb_makeindices");
    if (N != M) FAIL("N = %lld, expected %d", (long long)N, M);
    ok(c, set_gridmetrics(c, area, thk, zt, edge, dnbr, NULL, NULL, NULL), "otmb_set_gridmetrics");
    const double* phi[6] = {zero, zero, zero, zero, zero, zero};
    ok(c, set_facefluxes(c, phi), "otmb_set_facefluxes");
    ok(c, set_mlotst(c, ml), "otmb_set_mlotst");
    otmb_tm_params prm = {kH, kVML, kVdeep, 1035.0, 1, /*index_base*/ 1, OTMB_PATH_FUSED, 0};
    int64_t nnz[5];
    ok(c, build(c, &prm, nnz), "otmb_transportmatrix_build");
    if (nnz[OTMB_MAT_TADV] != 0) FAIL("resting ocean: Tadv must be empty, nnz = %lld", (long long)nnz[1]);
    if (nnz[OTMB_MAT_TKVDEEP] != 2 * M) FAIL("TkVdeep nnz = %lld, expected %d", (long long)nnz[4], 2 * M);
    if (nnz[OTMB_MAT_TKVML] != 0) FAIL("one level inside the mixed layer: TkVML must be empty, nnz = %lld", (long long)nnz[3]);

    static int64_t colptr[5][M + 1], rowval[5][7 * M];
    static double nzval[5][7 * M], dense[5][M][M];
    for (int m = 0; m < 5; ++m) {
        ok(c, fetch(c, m, colptr[m], rowval[m], nzval[m]), "otmb_transportmatrix_fetch");
        if (colptr[m][0] != 1 || colptr[m][M] != nnz[m] + 1) FAIL("matrix %d: colptr ends", m);
        for (int j = 0; j < M; ++j) {
            if (colptr[m][j] > colptr[m][j + 1]) FAIL("matrix %d: colptr not monotone", m);
            for (int64_t e = colptr[m][j] - 1; e < colptr[m][j + 1] - 1; ++e) {
                if (rowval[m][e] < 1 || rowval[m][e] > M) FAIL("matrix %d: row index out of range", m);
                if (e > colptr[m][j] - 1 && rowval[m][e - 1] >= rowval[m][e]) FAIL("matrix %d: rows not ascending in a column", m);
                dense[m][rowval[m][e] - 1][j] = nzval[m][e];
            }
        }
    }
    /* TκVdeep = κ [[I, -I], [-I, I]] (8x8 blocks): cell i of level 1 pairs with cell i + 8 of level 2 */
    for (int i = 0; i < M; ++i)
        for (int j = 0; j < M; ++j) {
            const double want = i == j ? kVdeep : ((i + P) % M == j ? -kVdeep : 0.0);
            if (dense[OTMB_MAT_TKVDEEP][i][j] != want) FAIL("TkVdeep[%d][%d] = %g, expected %g", i, j, dense[4][i][j], want);
        }
    /* T = ((Tadv + TκH) + TκVML) + TκVdeep, entry by entry; TκH conserves (row sums of the symmetric box vanish) */
    for (int i = 0; i < M; ++i) {
        double row = 0.0;
        for (int j = 0; j < M; ++j) {
            const double want = ((dense[1][i][j] + dense[2][i][j]) + dense[3][i][j]) + dense[4][i][j];
            if (dense[0][i][j] != want) FAIL("T[%d][%d] = %g, sum of the operators %g", i, j, dense[0][i][j], want);
            row += dense[2][i][j];
        }
        if (fabs(row) > 1e-12) FAIL("TkH row %d sums to %g", i, row);
        if (dense[2][i][i] <= 0.0) FAIL("TkH diagonal must be positive");
    }
    /* the one-call, slab-pipelined form must return the same arrays */
    static int64_t cp2[5][M + 1], rv2[5][7 * M];
    static double nz2[5][7 * M];
    int64_t cap[5], nnz2[5], *pc[5], *pr[5];
    double* pv[5];
    for (int m = 0; m < 5; ++m) cap[m] = 7 * M, pc[m] = cp2[m], pr[m] = rv2[m], pv[m] = nz2[m];
    ok(c, stream(c, &prm, phi, ml, NULL, 2, cap, pc, pr, pv, nnz2), "otmb_transportmatrix_stream");
    for (int m = 0; m < 5; ++m) {
        if (nnz2[m] != nnz[m] || memcmp(cp2[m], colptr[m], sizeof(int64_t) * (M + 1)) ||
            memcmp(rv2[m], rowval[m], sizeof(int64_t) * nnz[m]) || memcmp(nz2[m], nzval[m], sizeof(double) * nnz[m]))
            FAIL("otmb_transportmatrix_stream differs from build + fetch for matrix %d", m);
    }
    /* a prerequisite error is a status code with a message, not a crash */
    ok(c, set_grid(c, NX, NY, NZ, OTMB_TOPO_TRIPOLAR), "otmb_set_grid");
    if (build(c, &prm, nnz) != OTMB_ERR_STATE || !strstr(last_error(c), "otmb_makeindices")) FAIL("missing prerequisite not reported");
    ok(c, destroy(c), "otmb_destroy");
    printf("ABI-DRIVER: OK (N = %d, nnz T/Tadv/TkH/TkVML/TkVdeep = %lld/%lld/%lld/%lld/%lld)\n", M, (long long)nnz2[0], (long long)nnz2[1],
           (long long)nnz2[2], (long long)nnz2[3], (long long)nnz2[4]);
    return 0;
}
