/* otmb.h — C ABI of libotmb.so: B200-native (sm_100a) transport-matrix assembly.
 *
 * This is the drop-in boundary for the matrix-assembly hot path of
 * OceanTransportMatrixBuilder.jl v0.8.3.  The reference has no FFI of its own: its
 * boundary is the exported Julia API (src/OceanTransportMatrixBuilder.jl:31-36), and a
 * thin host shim (Julia `ccall`; Python ctypes in this repository)
 * keeps those signatures and forwards to the entry points below.  Each entry point
 * names the reference interface it replaces (paths relative to /root/reference).
 *
 * Conventions
 *  - plain pointers and sizes only; all reals are Float64, all indices Int64.
 *  - arrays are Julia column-major: (nx,ny,nz) -> L = i + nx*(j-1) + nx*ny*(k-1);
 *    2-D fields (nx,ny); vertex arrays (4,nx,ny); "4-direction" fields are 4 consecutive
 *    (nx,ny) planes in the reference's `dirs` order south, east, north, west
 *    (src/gridcellgeometry.jl:304).
 *  - every function returns an int32 status (0 = OTMB_OK).  otmb_last_error(ctx) gives a
 *    message valid until the next call on that ctx; the messages of the reference's own
 *    errors are reproduced verbatim.
 *  - the caller owns every host array; the library never keeps a host pointer after
 *    returning.  Host pointers may be pageable, or pinned via otmb_host_alloc.
 *  - calls on one ctx must be serialised by the caller; different ctxs (one per GPU)
 *    may be driven from different threads.
 *  - there is NO CPU fallback: without a usable sm_100 GPU, otmb_create fails.
 */
#ifndef OTMB_H
#define OTMB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct otmb_ctx otmb_ctx;

enum {
    OTMB_OK = 0,
    OTMB_ERR_TADV_NAN = 1,      /* "Tadv contains NaNs."    src/matrixbuilding.jl:39  */
    OTMB_ERR_TKH_NAN = 2,       /* "TκH contains NaNs."     src/matrixbuilding.jl:61  */
    OTMB_ERR_TKVML_NAN = 3,     /* "TκVML contains NaNs."   src/matrixbuilding.jl:90  */
    OTMB_ERR_TKVDEEP_NAN = 4,   /* "TκVdeep contains NaNs." src/matrixbuilding.jl:114 */
    OTMB_ERR_RHO_NAN = 5,       /* "ρ contains NaNs"        src/matrixbuilding.jl:233 */
    OTMB_ERR_UNKNOWN_GRID = 6,  /* "Unknown grid type"      src/gridtopology.jl:111-116 */
    OTMB_ERR_ALL_FILL = 7,      /* @assert, umo/vmo all NaN or fill, src/velocities.jl:199-200 */
    OTMB_ERR_DRY_NEIGHBOUR = 8, /* non-zero flux from a dry/absent neighbour; the reference
                                   hits a MethodError there, src/matrixbuilding.jl:247-250 */
    OTMB_ERR_BADARG = 9,
    OTMB_ERR_STATE = 10,        /* call order: a prerequisite step has not run on this ctx */
    OTMB_ERR_COMM = 11,         /* NCCL could not be loaded / a collective failed (sharded runs only) */
    OTMB_ERR_CUDA = 100,
    OTMB_ERR_NO_GPU = 101,
    OTMB_ERR_TOO_LARGE = 102
};

enum { OTMB_TOPO_BIPOLAR = 0, OTMB_TOPO_TRIPOLAR = 1, OTMB_TOPO_UNKNOWN = 2 };   /* src/gridtopology.jl:1-16 */
enum { OTMB_MAT_T = 0, OTMB_MAT_TADV = 1, OTMB_MAT_TKH = 2, OTMB_MAT_TKVML = 3, OTMB_MAT_TKVDEEP = 4 };
enum { OTMB_DIR_SOUTH = 0, OTMB_DIR_EAST = 1, OTMB_DIR_NORTH = 2, OTMB_DIR_WEST = 3 };
enum { OTMB_FACE_EAST = 0, OTMB_FACE_WEST = 1, OTMB_FACE_NORTH = 2, OTMB_FACE_SOUTH = 3,
       OTMB_FACE_TOP = 4, OTMB_FACE_BOTTOM = 5 };                                 /* src/velocities.jl:245-252 */

/* assembly strategy.  Both give bit-identical results.
 *  FUSED : one pass per wet column writes CSC directly (single-pass decoupled look-back scan)
 *  FUSED2: same column kernel, count pass + block scan + fill pass (cross-check of FUSED)
 *  COO   : fixed-slot triplet emitters + generic COO->CSC (a device `sparse`) + 3 sparse adds,
 *          i.e. the reference's own pipeline step by step */
enum { OTMB_PATH_FUSED = 0, OTMB_PATH_FUSED2 = 1, OTMB_PATH_COO = 2 };

int otmb_version(void);
int otmb_device_count(int* count);
const char* otmb_status_string(int status);

/* context: owns device buffers, a stream, cached geometry; one per GPU */
int otmb_create(otmb_ctx** ctx, int device);
int otmb_destroy(otmb_ctx* ctx);
const char* otmb_last_error(const otmb_ctx* ctx);

/* Page-locked host memory: what makes the copies run at the rate of the link (a copy into pageable arrays runs at a
 * fraction of it).  Pinning costs ~100 ms per GB, so freed blocks go to a process-wide pool (up to 8 GB) and are handed
 * out again to requests they fit: a shim that wraps results in garbage-collected arrays (Julia: unsafe_wrap + a
 * finalizer calling otmb_host_free) pays for the pinning once, not every month.  otmb_host_trim releases the pool.
 * Thread-safe; otmb_host_free rejects pointers that did not come from otmb_host_alloc. */
int otmb_host_alloc(void** ptr, int64_t bytes);
int otmb_host_free(void* ptr);
int otmb_host_trim(void);

/* grid shape + topology tag.  Replaces the AbstractGridTopology structs,
 * src/gridtopology.jl:1-16; the tag comes from getgridtopology (:33-53), which stays on the host. */
int otmb_set_grid(otmb_ctx* ctx, int64_t nx, int64_t ny, int64_t nz, int topology);

/* makeindices(v3D), src/matrixbuilding.jl:10-24.  v3D (nx,ny,nz) is NaN at dry cells and
 * stays resident on the device.  *N = number of wet cells. */
int otmb_makeindices(otmb_ctx* ctx, const double* v3D, int64_t* N);
/* the NamedTuple fields of makeindices; any pointer may be NULL.
 *   wet_chunks: ceil(M/64) UInt64 in Julia BitArray chunk layout (wet3D)
 *   Lwet:       N Int64, 1-based linear indices
 *   Lwet3D:     M Int64, 1-based wet index, 0 = missing */
int otmb_get_indices(otmb_ctx* ctx, uint64_t* wet_chunks, int64_t* Lwet, int64_t* Lwet3D);

/* makegridmetrics numerics, src/gridcellgeometry.jl:283-285 (thkcello, Z3D) and :304-308
 * (edge_length_2D, distance_to_edge_2D, distance_to_neighbour_2D).  Inputs: area2D (NaN on
 * land), lon/lat (nx,ny), vertices (4,nx,ny) already permuted (vertexpermutation :158-178 is
 * host work), zt (nz).  Uses the resident v3D.  Outputs may be NULL; thkcello, Z3D, area2D,
 * zt, edge, dnbr, lon, lat stay resident. */
int otmb_gridmetrics(otmb_ctx* ctx, const double* area2D, const double* lon, const double* lat,
                     const double* lon_vertices, const double* lat_vertices, const double* zt,
                     double* thkcello, double* Z3D, double* edge_length, double* distance_to_edge,
                     double* distance_to_neighbour);
/* upload caller-held grid metrics instead (the fields transportmatrix reads from the
 * gridmetrics NamedTuple, src/matrixbuilding.jl:340,441).  Z3D/lon/lat may be NULL. */
int otmb_set_gridmetrics(otmb_ctx* ctx, const double* area2D, const double* thkcello, const double* zt,
                         const double* edge_length, const double* distance_to_neighbour,
                         const double* Z3D, const double* lon, const double* lat);

/* facefluxesfrommasstransport -> facefluxes (+ nofluxboundaries!), src/velocities.jl:118-130,
 * 154-255.  umo/vmo (nx,ny,nz) Float64, not modified.  The six outputs (may be NULL) are the
 * NamedTuple (east, west, north, south, top, bottom); they also stay resident as ϕ. */
int otmb_facefluxes(otmb_ctx* ctx, const double* umo, const double* vmo, double fill_value,
                    double* east, double* west, double* north, double* south, double* top, double* bottom);
/* ---- ONE matrix sharded across the GPUs of a box: row slabs (one context per GPU / rank) --------------------
 * A grid row is R = j + ny*k (0-based), so a range of rows [row_begin, row_end) is a contiguous range of linear
 * cells and — wet ranks being ordered by linear index (src/matrixbuilding.jl:14-16) — a contiguous block of
 * rows/columns of every matrix.  A slab context keeps its rows plus one level (nx*ny cells) of halo on either side
 * resident (device arrays are allocated for that window only, set-up kernels scan only it), assembles the CSC
 * COLUMNS of its owned cells (row indices stay global wet ranks) and returns them with local colptr offsets; the
 * matrix is the concatenation of the ranks' segments.  Cuts may fall inside a level (any j), which balances the
 * ranks to within one grid row of N/R wet cells.
 * Every function still takes the caller's FULL (nx,ny,nz) host arrays; only the window is copied.
 * The reference has no counterpart (single process); this shards its loop `for 𝑖 in eachindex(Lwet)`
 * (src/matrixbuilding.jl:237, 348, 450).
 *
 * Two ways to drive it:
 *  (a) with the library's communicator (NCCL over NVLink, loaded on demand):
 *        otmb_plan_slabs -> otmb_set_grid -> otmb_set_slab_rows -> otmb_comm_init -> otmb_sharded_makeindices ->
 *        otmb_set_gridmetrics -> otmb_set_masstransport -> otmb_sharded_facefluxes -> otmb_set_mlotst ->
 *        otmb_sharded_transportmatrix_build -> otmb_transportmatrix_fetch[_all]
 *      The host program only has to give every rank the same 128-byte id (MPI.bcast, a file, torch.distributed, ...).
 *  (b) with exchanges done by the host (tests on one GPU, gloo): otmb_makeindices / otmb_slab_counts /
 *      otmb_set_rank_offset / otmb_facefluxes_slab / otmb_transportmatrix_build.                                  */
/* host-side plan, no GPU needed: cut the ny*nz grid rows into nranks contiguous slabs of about equal wet count
 * (wet <=> !isnan(v3D)).  level_cuts_only != 0 restricts cuts to level boundaries.  row_cuts: nranks+1 entries,
 * rank r owns [row_cuts[r], row_cuts[r+1]); wet_per_rank (nranks, may be NULL).  Deterministic. */
int otmb_plan_slabs(const double* v3D, int64_t nx, int64_t ny, int64_t nz, int32_t nranks, int32_t level_cuts_only,
                    int64_t* row_cuts, int64_t* wet_per_rank);
int otmb_set_slab_rows(otmb_ctx* ctx, int64_t row_begin, int64_t row_end);   /* (0, ny*nz) = unsharded */
int otmb_set_slab(otmb_ctx* ctx, int64_t k_begin, int64_t k_end);   /* whole levels: rows [k_begin*ny, k_end*ny) */
int otmb_slab_counts(otmb_ctx* ctx, int64_t* n_owned, int64_t* n_halo_above);
int otmb_set_rank_offset(otmb_ctx* ctx, int64_t w0);   /* global wet rank (0-based) of the first owned cell */
/* facefluxes on a slab.  The continuity scan (src/velocities.jl:234-243) runs bottom-up and is not
 * associative in floating point, so the slabs form a chain: carry_in[i,j] = ϕtop of the cell below this slab's
 * deepest owned cell of column (i,j), computed by the slab below (NULL for the deepest slab), carry_out[i,j] = ϕtop of
 * the column's first owned level, for the slab above (a column of which the slab owns nothing passes the value on).
 * carry_on_device != 0: both are DEVICE pointers of nx*ny doubles.  valid_uv[0..1]: whether this slab saw any valid
 * umo / vmo value; the caller combines the ranks and raises the reference's assertion (:199-200).
 * Outputs: owned cells of the full-size arrays. */
int otmb_facefluxes_slab(otmb_ctx* ctx, const double* umo, const double* vmo, double fill_value,
                         const double* carry_in, double* carry_out, int32_t carry_on_device, int32_t valid_uv[2],
                         double* east, double* west, double* north, double* south, double* top, double* bottom);

/* the library's communicator: one rank per context / GPU.  otmb_comm_unique_id on one rank, the 128 bytes handed
 * to every rank by the host program, otmb_comm_init on all of them (collective).  nranks == 1 needs no id and
 * loads nothing. */
#define OTMB_COMM_ID_BYTES 128
int otmb_comm_unique_id(uint8_t id[OTMB_COMM_ID_BYTES]);
int otmb_comm_init(otmb_ctx* ctx, int32_t nranks, int32_t rank, const uint8_t id[OTMB_COMM_ID_BYTES]);
int otmb_comm_free(otmb_ctx* ctx);
int otmb_comm_allgather_i64(otmb_ctx* ctx, const int64_t* mine, int32_t count, int64_t* all /* nranks*count */);
/* transport of the default face-flux carry chain on this context: 1 = peer memory (CUDA IPC, fused kernel),
 * -1 = NCCL send / recv, 0 = no chain has run yet (or one rank) */
int otmb_comm_chain_transport(otmb_ctx* ctx, int32_t* transport);
/* collective steps (every rank calls them; a failure on ANY rank is returned on EVERY rank, so nobody is left
 * waiting in a collective for a peer that has raised):
 *  - makeindices on the slab + all-gather of the owned counts: N of the whole ocean, this rank's first wet rank
 *    (already applied: no otmb_set_rank_offset needed) and its number of columns;
 *  - otmb_set_masstransport uploads the window of umo / vmo (local); otmb_sharded_facefluxes runs the continuity
 *    chain from the sea-floor rank (nranks-1) up to the surface rank (0), PIPELINED: a rank starts on a block of
 *    columns as soon as the rank below has delivered it.  nchunks == 0 (default): the peer-memory form — each rank
 *    exports an inbox with CUDA IPC, ONE k_faceflux launch per rank whose thread blocks wait for / raise per-block
 *    flags in peer memory, the send being plain stores over NVLink from inside the kernel (per-block acks keep a
 *    fast rank from overwriting values that have not been read); if IPC / P2P is unavailable, or nchunks > 0: the
 *    carry plane is cut into nchunks column chunks (8 by default) sent / received with NCCL on the context's
 *    stream.  Same bits either way.  Outputs (may be NULL): owned cells of full-size arrays.  The all-fill assertion (:199-200) is evaluated over all ranks.  otmb_sharded_facefluxes_enqueue is
 *    the device-resident chain alone (no flags, no copies, does not block);
 *  - assembly of this rank's columns + all-gather of the nnz: entries held by lower ranks (add to the local
 *    colptr) and of the whole matrix. */
int otmb_sharded_makeindices(otmb_ctx* ctx, const double* v3D, int64_t* N_global, int64_t* w0, int64_t* n_owned);
int otmb_set_masstransport(otmb_ctx* ctx, const double* umo, const double* vmo, double fill_value);
int otmb_sharded_facefluxes(otmb_ctx* ctx, int32_t nchunks, double* east, double* west, double* north, double* south,
                            double* top, double* bottom);
int otmb_sharded_facefluxes_enqueue(otmb_ctx* ctx, int32_t nchunks);
/* int otmb_sharded_transportmatrix_build(...): declared below, after otmb_tm_params */

/* upload caller-held ϕ (order OTMB_FACE_*) */
int otmb_set_facefluxes(otmb_ctx* ctx, const double* const phi[6]);
int otmb_set_mlotst(otmb_ctx* ctx, const double* mlotst /* (nx,ny), NaN = missing */);
int otmb_set_rho3d(otmb_ctx* ctx, const double* rho3d /* (nx,ny,nz) or NULL to use params.rho */);

typedef struct otmb_tm_params {
    double kH;        /* κH     default 500.0  src/matrixbuilding.jl:130 */
    double kVML;      /* κVML   default 0.1    :131 */
    double kVdeep;    /* κVdeep default 1.0e-5 :132 */
    double rho;       /* scalar ρ, used when no 3-D ρ is set (:221-225) */
    int32_t upwind;   /* default 1 (:137) */
    int32_t index_base;  /* 1 for Julia, 0 for C/Python consumers of colptr/rowval */
    int32_t path;     /* OTMB_PATH_* */
    int32_t build_mask;  /* 0 = build all four operators.  Otherwise bit m (1..4) set: build operator m;
                            a clear bit means the operator was supplied with otmb_set_operator
                            (:133-143).  Set bit 5 (32) to pass an explicit mask with no operator bits. */
} otmb_tm_params;

/* transportmatrix(; ϕ, mlotst, gridmetrics, indices, ρ, κH, κVML, κVdeep, Tadv, TκH, TκVML,
 * TκVdeep, upwind), src/matrixbuilding.jl:128-150, on the resident inputs.  Blocks until the
 * five matrices are complete in device memory; nnz_out[OTMB_MAT_*]. */
int otmb_transportmatrix_build(otmb_ctx* ctx, const otmb_tm_params* params, int64_t nnz_out[5]);
/* the collective form for a sharded run (see "ONE matrix sharded across the GPUs" above) */
int otmb_sharded_transportmatrix_build(otmb_ctx* ctx, const otmb_tm_params* params, int64_t nnz_local[5],
                                       int64_t nnz_before[5], int64_t nnz_total[5]);
/* copy one result out as SparseMatrixCSC fields: colptr (N+1), rowval (nnz), nzval (nnz) */
int otmb_transportmatrix_fetch(otmb_ctx* ctx, int which, int64_t* colptr, int64_t* rowval, double* nzval);
/* the same for several results at once (bit m of mask = matrix OTMB_MAT_m, 0 = all five; NULL entries are
 * skipped): what the shim calls to fill the NamedTuple (; T, Tadv, TκH, TκVML, TκVdeep) of
 * src/matrixbuilding.jl:149.  One pipeline for all arrays: the Int64 indices cross PCIe as Int32 and are widened
 * into the caller's arrays by host threads (OTMB_HOST_THREADS, default min(8, cores / (2 LOCAL_WORLD_SIZE))) while the
 * values are in flight; matrices with 2^31 or more rows / entries, or ranks with fewer than four such threads to
 * spare, copy the 8-byte arrays as they are.  Both calls return identical arrays. */
int otmb_transportmatrix_fetch_all(otmb_ctx* ctx, int mask, int64_t* const colptr[5], int64_t* const rowval[5],
                                   double* const nzval[5]);
/* transportmatrix end to end in ONE call — host ϕ (six arrays, order OTMB_FACE_*), mlotst and optionally a 3-D ρ in,
 * the five host CSC matrices out — pipelined by level slabs: the upload of slab s+1, the assembly of slab s and the
 * copy-out of slab s-1 overlap, so the two directions of the PCIe link are busy at once (a set_facefluxes -> build ->
 * fetch_all sequence runs them back to back).  nslabs: 0 = default.  The nnz are data dependent: the caller passes
 * rowval / nzval arrays holding capacity[m] entries (N x {7,7,5,3,3} always suffices) and colptr arrays of N+1, and
 * receives nnz_out; the matrices are the first nnz entries (Julia: resize!).  Results are bit-identical to
 * otmb_transportmatrix_build + fetch, and stay resident like theirs.  Page-locked arrays (otmb_host_alloc) move at
 * link rate; pageable result arrays are filled through pinned staging by the host threads. */
int otmb_transportmatrix_stream(otmb_ctx* ctx, const otmb_tm_params* params, const double* const phi[6],
                                const double* mlotst, const double* rho3d, int32_t nslabs, const int64_t capacity[5],
                                int64_t* const colptr[5], int64_t* const rowval[5], double* const nzval[5],
                                int64_t nnz_out[5]);
/* the resident result matrices (bit m of mask = OTMB_MAT_m, 0 = all five) straight to a binary file, through pinned
 * staging, without host SparseMatrixCSC objects in between (the 12-month batch).  Layout, little
 * endian: "OTMBCSC1"; Int64 N, index_base, nmat; per matrix Int64 id, Int64 nnz; then per matrix colptr (N+1 Int64),
 * rowval (nnz Int64), nzval (nnz Float64). */
int otmb_transportmatrix_dump(otmb_ctx* ctx, int mask, const char* path);
/* host half of that pipeline alone (no GPU needed): sign-extend n Int32 indices into Int64 on `threads` pool
 * threads (0 = the calling thread). */
int otmb_host_widen(const int32_t* src, int64_t* dst, int64_t n, int32_t threads);
/* a pre-built operator passed by the caller (the Tadv/TκH/TκVML/TκVdeep kwargs, :133-143);
 * indices in params.index_base of the next build.  The CSC is checked (N+1 non-decreasing colptr entries from
 * index_base to index_base + nnz, rows strictly ascending inside [0, N)): OTMB_ERR_BADARG otherwise, and the operator
 * is not kept.  The next build sums what was supplied with what it builds, (((Tadv + TκH) + TκVML) + TκVdeep), :147;
 * when a supplied operator equals, bit for bit, what the build's own inputs give, that is done in the single-pass
 * kernel, otherwise by the generic sparse `+` — the result is the same either way. */
int otmb_set_operator(otmb_ctx* ctx, int which, int64_t nnz, const int64_t* colptr, const int64_t* rowval,
                      const double* nzval, int32_t index_base);

/* generic device `sparse(I, J, V, n, n)` (SparseArrays.sparse, called at
 * src/matrixbuilding.jl:41,63,92,116) and sparse A + B (:147), exposed for parity tests.
 * Indices 1-based in and out.  Two-phase: *_build returns nnz, *_fetch copies out. */
int otmb_sparse_build(otmb_ctx* ctx, int64_t len, const int64_t* I, const int64_t* J, const double* V,
                      int64_t n, int64_t* nnz);
int otmb_sparse_fetch(otmb_ctx* ctx, int64_t* colptr, int64_t* rowval, double* nzval);
int otmb_spadd_build(otmb_ctx* ctx, int64_t n, const int64_t* a_colptr, const int64_t* a_rowval,
                     const double* a_nzval, const int64_t* b_colptr, const int64_t* b_rowval,
                     const double* b_nzval, int64_t* nnz);
int otmb_spadd_fetch(otmb_ctx* ctx, int64_t* colptr, int64_t* rowval, double* nzval);

/* Redi/GM helpers (experimental, not exported by the reference; fields, not matrices):
 * globalverticalfacetriadderivative src/triads.jl:134-146 (dir 0 = Icoord, 1 = Jcoord),
 * globalverticaldyadderivative src/dyads.jl:66-78, bolus_GM_velocity src/RediGM.jl:46-79.
 * chi/rho/out are (nx,ny,nz) host arrays; need resident v3D, Z3D, lon, lat. */
int otmb_triad_derivative(otmb_ctx* ctx, const double* chi, int dir, double* out);
int otmb_dyad_derivative(otmb_ctx* ctx, const double* chi, double* out);
int otmb_bolus_gm_velocity(otmb_ctx* ctx, const double* rho, double kGM, double maxslope, double* u, double* v);

/* BASELINE configs[2] ("C3"): the Gent-McWilliams bolus transport folded into the advective mass fluxes — an
 * EXTENSION (parity unpinned: the reference has bolus_GM_velocity, src/RediGM.jl:46-79, and velocity2fluxes,
 * src/velocities.jl:10-39, but no wiring between them and no Redi/GM option of transportmatrix).  On the device:
 *   (u*, v*) = bolus_GM_velocity(rho3d; kGM, maxslope);  (ϕᵢ*, ϕⱼ*) = velocity2fluxes(u*, v*, gridmetrics, ρ)
 *   ϕ = facefluxes(umo + ϕᵢ*, vmo + ϕⱼ*)   (a NaN bolus flux adds nothing; fill / NaN transports stay as they are)
 * and ϕ stays resident for otmb_transportmatrix_build, whose 7-point pattern is unchanged.  ρ of velocity2fluxes:
 * rho3d itself (flux_rho_is_3d != 0) or rho_scalar.  kGM = 0 reproduces otmb_facefluxes bit for bit.  Needs the
 * resident Z3D, lon, lat, thkcello, edge lengths (otmb_gridmetrics); tripolar grids only.  The six outputs and the
 * bolus fluxes gm_phi_i / gm_phi_j (nx,ny,nz) may be NULL. */
int otmb_facefluxes_gm(otmb_ctx* ctx, const double* umo, const double* vmo, double fill_value, const double* rho3d,
                       double kGM, double maxslope, double rho_scalar, int32_t flux_rho_is_3d, double* east, double* west,
                       double* north, double* south, double* top, double* bottom, double* gm_phi_i, double* gm_phi_j);

/* velocity <-> mass flux on the C-grid, needs the resident thkcello / edge lengths:
 * velocity2fluxes src/velocities.jl:10-39 (ϕᵢ = u·ρ̄·min(thk)·edge_east, ϕⱼ = v·ρ̄·min(thk)·edge_north, NaN-aware
 * two-cell mean / min :81-108), fluxes2velocity :50-74 (the inverse).  All arrays (nx,ny,nz); rho3d may be
 * NULL (scalar rho).  Tripolar grids only: on bipolar grids the reference indexes a missing neighbour and throws.
 * otmb_bgrid_to_cgrid: the B-grid (NE corner) branch of interpolateontodefaultCgrid,
 * src/gridcellgeometry.jl:118-128; the Arakawa-grid detection (:50-95, one cell) stays on the host. */
int otmb_velocity2fluxes(otmb_ctx* ctx, const double* u, const double* v, const double* rho3d, double rho,
                         double* phi_i, double* phi_j);
int otmb_fluxes2velocity(otmb_ctx* ctx, const double* phi_i, const double* phi_j, const double* rho3d, double rho,
                         double* u, double* v);
int otmb_bgrid_to_cgrid(otmb_ctx* ctx, const double* u, const double* v, double fill_value, double* u2, double* v2);

/* lump_and_spray(wet3D, vol, T; di, dj, dk), src/extratools.jl:38-112, default mask only:
 * LUMP (N_c x N, volume-weighted average onto di x dj x dk boxes split into connected components of T's stored
 * pattern), SPRAY (N x N_c, = LUMP' with ones) and the coarse volumes.  vol: N doubles (volumes of the wet
 * cells).  T's pattern: t_colptr / t_rowval (host, base t_index_base; checked like a pre-built operator: N+1
 * non-decreasing colptr entries, rows ascending inside [0, N), OTMB_ERR_BADARG otherwise), or NULL to use the T of the
 * last otmb_transportmatrix_build on this context.  Results in index_base; two-phase like transportmatrix. */
int otmb_lump_and_spray_build(otmb_ctx* ctx, int64_t di, int64_t dj, int64_t dk, const double* vol,
                              const int64_t* t_colptr, const int64_t* t_rowval, int32_t t_index_base,
                              int32_t index_base, int64_t* n_coarse);
int otmb_lump_and_spray_fetch(otmb_ctx* ctx, int64_t* lump_colptr /* N+1 */, int64_t* lump_rowval /* N */,
                              double* lump_nzval /* N */, int64_t* spray_colptr /* N_c+1 */, int64_t* spray_rowval /* N */,
                              double* spray_nzval /* N */, double* vol_c /* N_c */);

/* T_c = LUMP * T * SPRAY (test/local_full.jl:161, the coarse operator of the reference's downstream solve) on the
 * device: LUMP / SPRAY of the last otmb_lump_and_spray_build, T = the RESIDENT result matrix `which` (OTMB_MAT_*) of the
 * last build.  LUMP has one entry per column and SPRAY column J holds ones at the members of component J, so this is a
 * gather per coarse column (one thread each), accumulated in the order of SparseArrays' product (LUMP * T first, then
 * * SPRAY): bit-identical, structural zeros kept.  Two-phase; indices in the lump build's index base. */
int otmb_coarsen_build(otmb_ctx* ctx, int which, int64_t* n_coarse, int64_t* nnz);
int otmb_coarsen_fetch(otmb_ctx* ctx, int64_t* colptr /* N_c+1 */, int64_t* rowval, double* nzval);

/* y = X x (transpose = 0) or y = Xᵀ x (transpose != 0) on the RESIDENT result matrix `which` (OTMB_MAT_*) of the
 * last build: the products behind the reference's conservation checks τdiv = ‖1‖/‖T 1‖, τvol = ‖v‖/‖Tᵀ v‖
 * (test/online.jl:110-115) without copying the matrix to the host first.  x, y: N doubles (host).  Deterministic
 * and bit-identical to a sequential CSC product. */
int otmb_spmv(otmb_ctx* ctx, int which, int transpose, const double* x, double* y);

/* Position-dependent 64-bit checksums of the resident result `which`, taken as the segment of a larger matrix that
 * starts at column col_offset / entry entry_offset: out[0] over colptr (first ncols entries), out[1] over rowval,
 * out[2] over the bit patterns of nzval; each is a SUM of hashes of (global position, value), so the checksums of the
 * ranks of a sharded run add up (mod 2^64) to those of the same matrix assembled on one GPU — how bench.py proves
 * that the N-rank result equals the 1-rank one without moving the matrix. */
int otmb_result_checksum(otmb_ctx* ctx, int which, int64_t col_offset, int64_t entry_offset, uint64_t out[3]);

/* measurement helpers: CUDA events on the ctx stream (the stream every kernel of this
 * library is launched on), an L2 flush, and per-kernel launch counting. */
int otmb_timer_start(otmb_ctx* ctx);
int otmb_timer_stop(otmb_ctx* ctx, float* milliseconds);
int otmb_l2_flush(otmb_ctx* ctx);
int otmb_launch_count(otmb_ctx* ctx, int64_t* launches);   /* kernels launched by this ctx so far */
int otmb_last_build_ms(otmb_ctx* ctx, float* milliseconds); /* device time of the last transportmatrix_build */
/* whether otmb_transportmatrix_build brackets its kernels with a CUDA event pair for otmb_last_build_ms (default on;
 * off saves two driver calls per build — a build is then exactly two launches, the kernel and its completion record) */
int otmb_set_build_timing(otmb_ctx* ctx, int32_t on);
int otmb_synchronize(otmb_ctx* ctx);
/* device self-test: the assembly kernel divides through a two-at-a-time routine (csrc/fdiv.cuh) that must give the
 * compiler's IEEE-754 quotient bit for bit; this compares the two over 2n operand pairs drawn from `seed` (random bit
 * patterns, ordinary magnitudes, and exponents at the edges: zero, subnormal, huge, Inf, NaN).  *mismatches must be 0;
 * first_bad (may be NULL) receives {a, b, routine's quotient, a / b} of one disagreement. */
int otmb_selftest_division(otmb_ctx* ctx, int64_t n, uint64_t seed, int64_t* mismatches, double first_bad[4]);

#ifdef __cplusplus
}
#endif
#endif /* OTMB_H */
