/*
 * dugks.h — C-ABI of the B200-native DUGKS discrete-velocity update.
 *
 * This is the drop-in boundary for ONE path of zhulianhua/dugksFoam: the
 * per-time-step update behind Foam::fvDVM::evolution()
 * (reference src/fvDVM/fvDVM/fvDVM.C:1086-1108) and everything it reaches in
 * Foam::discreteVelocity (src/fvDVM/discreteVelocity/discreteVelocity.C).
 * The reference has no FFI layer of its own: the path sits behind the two C++
 * classes fvDVM / discreteVelocity.  A maintainer keeps those class interfaces
 * and replaces their bodies by calls into this header (INTEGRATION.md shows
 * the adapter).  Every entry point cites the reference code it replaces.
 *
 * Conventions
 *  - plain pointers and sizes only; all arrays are caller-owned HOST memory,
 *    borrowed for the duration of the call (the library copies what it keeps);
 *  - vectors are 3 contiguous doubles, tensors 9 (OpenFOAM `vector`/`tensor`
 *    layout, relied on by the reference at fieldMPIreducer.C:21-22);
 *  - labels are int32 (OpenFOAM default `label`);
 *  - every function returns 0 on success or a negative dugks_status; nothing
 *    throws or aborts across the ABI.  dugks_last_error() gives the text the
 *    C++ wrapper turns into FatalErrorIn(...) << ... << exit(FatalError);
 *  - a handle is driven by one host thread; there is no global state.
 *  - the library REQUIRES a CUDA device: there is no CPU fallback.
 *
 * Face numbering: internal faces 0..nInternalFaces-1 exactly as OpenFOAM's
 * owner/neighbour lists, followed by the faces of all NON-EMPTY patches in
 * polyMesh/boundary order ("boundary faces", local index b = face -
 * nInternalFaces).  Faces of `empty` patches take no part in any loop of the
 * reference and must not be passed.
 */
#ifndef DUGKS_H
#define DUGKS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DUGKS_ABI_VERSION 2

typedef enum dugks_status {
    DUGKS_OK = 0,
    DUGKS_ERR_INVALID = -1,   /* bad argument / inconsistent sizes          */
    DUGKS_ERR_NO_DEVICE = -2, /* no usable CUDA device (no CPU fallback)    */
    DUGKS_ERR_CUDA = -3,      /* CUDA runtime error, see dugks_last_error   */
    DUGKS_ERR_NOMEM = -4,     /* device or host allocation failed           */
    DUGKS_ERR_COMM = -5,      /* NCCL / reducer failure                     */
    DUGKS_ERR_UNSUPPORTED = -6
} dugks_status;

/*
 * Distribution-function boundary kind of a patch.  The reference derives it
 * from the type of the `rho` patch field through the map at
 * discreteVelocity.C:274-281; the names below are the reference's
 * fvsPatchField type names (src/fvDVM/BCs/<kind>FvsPatchField).
 */
typedef enum dugks_patch_kind {
    DUGKS_PATCH_ZERO_GRADIENT = 0, /* rho zeroGradient       -> "zeroGradient" discreteVelocity.C:551-555 */
    DUGKS_PATCH_MIXED = 1,         /* rho fixedValue         -> "mixed"        :556-573, init :312-344     */
    DUGKS_PATCH_MAXWELL_WALL = 2,  /* rho calculatedMaxwell  -> "maxwellWall"  :605-627, :693-731          */
    DUGKS_PATCH_FAR_FIELD = 3,     /* rho farField           -> "farField"     :574-604                    */
    DUGKS_PATCH_DVM_SYMMETRY = 4,  /* rho symmetryMod        -> "DVMsymmetry"  :673-689, :744-790          */
    DUGKS_PATCH_SYMMETRY_PLANE = 5,/* constraint symmetryPlane                 :673-689, :791-815          */
    DUGKS_PATCH_PRESSURE_IN = 6,   /* rho pressureIn         -> "farField" + fvDVM.C:743-773               */
    DUGKS_PATCH_PRESSURE_OUT = 7   /* rho pressureOut        -> "farField" + fvDVM.C:774-804               */
} dugks_patch_kind;

/* How the boundary value of U / T on a patch follows the cells
 * (Uvol_.correctBoundaryConditions(), fvDVM.C:698-699). */
typedef enum dugks_macro_bc {
    DUGKS_BC_FIXED_VALUE = 0,   /* value supplied by the caller, kept          */
    DUGKS_BC_ZERO_GRADIENT = 1  /* value = owner-cell value after every step   */
} dugks_macro_bc;

typedef struct dugks_patch_t {
    int32_t kind;      /* dugks_patch_kind                                      */
    int32_t start;     /* first boundary-face index b of the patch              */
    int32_t size;      /* number of faces                                       */
    int32_t U_bc;      /* dugks_macro_bc of the U patch field                   */
    int32_t T_bc;      /* dugks_macro_bc of the T patch field                   */
    int32_t reserved;
    double pressure;   /* pressureIn()/pressureOut() (fvDVM.C:751,782), else 0  */
} dugks_patch_t;

/*
 * Mesh addressing + geometry exactly as OpenFOAM hands it to the reference
 * (mesh_.owner(), neighbour(), C(), V(), Cf(), Sf() at
 * discreteVelocity.C:472-476,942-943; leastSquaresVectors pVectors/nVectors
 * as used by leastSquaresGrad, in-tree twin zeroBoundaryGrad.C:82-99;
 * boundary deltaCoeffs used by fixedGradient patches, discreteVelocity.C:121).
 */
typedef struct dugks_mesh_t {
    int32_t nCells;
    int32_t nInternalFaces;
    int32_t nBoundaryFaces;      /* faces of non-empty patches                  */
    int32_t nSolutionD;          /* mesh.nSolutionD(): 1, 2 or 3                */
    const int32_t* owner;        /* [nInternalFaces + nBoundaryFaces]           */
    const int32_t* neighbour;    /* [nInternalFaces]                            */
    const double* C;             /* [nCells][3] cell centres                    */
    const double* V;             /* [nCells]    cell volumes                    */
    const double* Cf;            /* [nFaces][3] face centres                    */
    const double* Sf;            /* [nFaces][3] face area vectors               */
    const double* ownLs;         /* [nInternalFaces][3] owner LS vectors        */
    const double* neiLs;         /* [nInternalFaces][3] neighbour LS vectors    */
    const double* patchLs;       /* [nBoundaryFaces][3] boundary LS vectors     */
    const double* deltaCoeffs;   /* [nFaces] 1/|d| (internal) and 1/|delta_b|   */
} dugks_mesh_t;

/*
 * Discrete-velocity set: the 1-D abscissae/weights of constant/Xis and
 * constant/weights (fvDVM.C:56-117).  The library builds the tensor-product
 * grid, the weights and the mirror ids exactly as fvDVM::initialiseDV
 * (fvDVM.C:140-220): global id = iz*n*n + iy*n + ix, ix fastest.
 */
typedef struct dugks_dvset_t {
    int32_t nXiPerDim;           /* DVMProperties fvDVMparas.nDV                */
    int32_t reserved;
    const double* Xis;           /* [nXiPerDim]                                 */
    const double* weights;       /* [nXiPerDim]                                 */
    double xiMax;                /* fvDVMparas.xiMax (Courant number only)      */
    double xiMin;
} dugks_dvset_t;

/* constant/DVMProperties gasProperties (fvDVM.C:928-933). */
typedef struct dugks_gas_t {
    double R;
    double omega;
    double Tref;
    double muRef;
    double Pr;
    int32_t KInner;
    int32_t reserved;
} dugks_gas_t;

/*
 * Velocity-space decomposition (the reference's -dvParallel, fvDVM.C:228-260):
 * every rank holds the whole mesh and a subset of the discrete velocities; the
 * moment sums are all-reduced (fieldMPIreducer.C:48-150).  One process per GPU.
 *
 * reduce: the collective backend that replaces fieldMPIreducer::reduceField.
 *   NULL with nRanks==1 : no collective.
 *   NULL with nRanks>1  : the library's own NCCL communicator, created from
 *                         nccl_unique_id (128 bytes from dugks_nccl_unique_id
 *                         on rank 0, broadcast by the caller: MPI_Bcast in the
 *                         OpenFOAM adapter, torch.distributed in the harness).
 *   non-NULL            : called with a DEVICE buffer of n doubles that must be
 *                         sum-reduced in place over all ranks, ordered on
 *                         `stream` (a cudaStream_t); returns 0 on success.
 */
typedef int (*dugks_allreduce_fn)(void* user, double* device_buf, size_t n, void* stream);

typedef struct dugks_par_t {
    int32_t rank;                /* fieldMPIreducer::rank()                     */
    int32_t nRanks;              /* fieldMPIreducer::nproc()                    */
    int32_t device;              /* CUDA device ordinal, -1 = current device    */
    int32_t partition;           /* reserved, must be 0: contiguous blocks of the slow DV indices,
                                    see dugks_partition */
    dugks_allreduce_fn reduce;
    void* reduce_user;
    const void* nccl_unique_id;  /* 128 bytes or NULL                           */
    size_t scratch_bytes;        /* cap (bytes) on the device memory taken for kept face values between the
                                    two phases of a step; 0 = all free device memory but a reserve.  Slabs
                                    that do not fit recompute their face values (slower, same results).   */
    int32_t store_h;             /* 1: always carry h; 0: elide h when K+3-D==0 (h==0 exactly,
                                    discreteVelocity.C:1043) */
    int32_t dv_chunk;            /* reserved, must be 0: a launch slab is 32 velocity rows */
    /* system/fvSchemes gradSchemes (doc/usage.tex:169-192).  0: "leastSquares", the scheme of every shipped case.
     * > 0: "VenkatakrishnanLimited leastSquares k" AS IT IS MEANT TO WORK (VenkatakrishnanLimitedGrads.C:59-226,
     * VenkatakrishnanSlopeMulti.C:83-128: grad *= min over the cell's faces of limitFace, eps^2 = k^3 V).  The
     * reference's own implementation limits a copy of the gradient and returns the unlimited one (:76, :225) and
     * starts the limiter from 0 instead of 1 (:140-151), so IN THE REFERENCE the scheme is inert: pass 0 to
     * reproduce a reference run that names it.  The limited gradient runs through the generic kernels (slower). */
    double limiter_k;
} dugks_par_t;

typedef struct dugks_handle dugks_handle_t;

/* ---- lifetime ---------------------------------------------------------- */

int dugks_abi_version(void);

/* Fills 128 bytes with a fresh ncclUniqueId (rank 0 calls it, then broadcasts). */
int dugks_nccl_unique_id(void* out128);

/*
 * Replaces the fvDVM constructor body (fvDVM.C:886-1075): initialiseDV
 * (:119-261), discreteVelocity::initDFtoEq (discreteVelocity.C:220-249),
 * setBCtype (:251-310), initBoundaryField (:312-344),
 * setCalculatedMaxwellRhoBC (fvDVM.C:263-309), updatePressureInOutBC
 * (:730-806), updateTau (:808-817) and the first Usurf (:1074).
 *
 * rho,U,T      : cell fields [nCells], [nCells][3], [nCells]
 * rho_b,U_b,T_b: boundary values per boundary face [nBoundaryFaces](x3)
 */
int dugks_create(const dugks_mesh_t* mesh,
                 const dugks_patch_t* patches, int32_t nPatches,
                 const dugks_dvset_t* dvset,
                 const dugks_gas_t* gas,
                 const dugks_par_t* par,
                 const double* rho, const double* U, const double* T,
                 const double* rho_b, const double* U_b, const double* T_b,
                 dugks_handle_t** out);

void dugks_destroy(dugks_handle_t* h);

/* Text of the last failure on this handle (or of the last failed create when
 * h == NULL).  Never NULL. */
const char* dugks_last_error(const dugks_handle_t* h);

/* ---- the hot path ------------------------------------------------------ */

/*
 * One time step = fvDVM::evolution() (fvDVM.C:1086-1108): stages
 * updateGHbarPvol, updateGHbarSurf, updateMaxwellWallRho,
 * updateGHbarSurfMaxwellWallIn, updateGHbarSurfSymmetryIn, updateMacroSurf,
 * updateGHsurf, updateGHtildeVol, updateMacroVol, updatePressureInOutBC.
 * dt is time_.deltaTValue() and may change every call (setDeltaTvar.H:34-47).
 * Work is enqueued on the handle's stream; the call does not synchronise.
 */
int dugks_step(dugks_handle_t* h, double dt);

/* Block until all enqueued steps are complete. */
int dugks_sync(dugks_handle_t* h);

/* Caller-side update of fixedValue boundary macros (time-varying BCs).
 * Any pointer may be NULL (unchanged).  Wall in-flux constants
 * (fvDVM.C:263-309) are recomputed.  Arrays identical to the ones passed last
 * (or at create) are a no-op. */
int dugks_set_boundary_macros(dugks_handle_t* h, const double* rho_b,
                              const double* U_b, const double* T_b);

/* ---- accessors (synchronise, D2H into caller storage; NULL = skip) ------ */

/* Page-lock caller-owned host arrays (the storage of the OpenFOAM fields the
 * accessors below fill, Field<T>::data(), fvDVM.C:427-430): the macro accessors
 * then copy asynchronously from the device straight into them instead of going
 * through a staging buffer and a host loop.  Optional; register once, release
 * before the array is freed.  Already registered memory is not an error. */
int dugks_host_register(void* p, size_t bytes);
int dugks_host_unregister(void* p);

/* rhoVol(), Uvol(), Tvol(), qVol(), tauVol()  (fvDVM.H:309-322). */
int dugks_get_cell_macros(dugks_handle_t* h, double* rho, double* U, double* T,
                          double* q, double* tau);

/* rhoSurf(), Usurf(), Tsurf(), qSurf(), tauSurf() on all nFaces
 * (fvDVM.H:324-338); stressSurf() is identically zero in the reference
 * (fvDVM.C:512-515) and is not transferred. */
int dugks_get_face_macros(dugks_handle_t* h, double* rho, double* U, double* T,
                          double* q, double* tau);

/* Boundary values of rho/U/T after the step (wall rho_w of
 * calculatedMaxwellFvPatchField.C:158, zeroGradient / pressure patches). */
int dugks_get_boundary_macros(dugks_handle_t* h, double* rho_b, double* U_b, double* T_b);

/* qWall [nBoundaryFaces][3], stressWall [nBoundaryFaces][9]; non-zero on
 * maxwellWall patches only (fvDVM.C:539-581). */
int dugks_get_wall_diag(dugks_handle_t* h, double* qWall, double* stressWall);

/* fvDVM::getCoNum (fvDVM.C:1111-1119) evaluated on the device. */
int dugks_courant(dugks_handle_t* h, double dt, double* maxCo, double* meanCo);

/* Convergence monitor of the time loop (dugksFoam.C:88-107, fields of createFields.H:40-82) on the
 * device: change[0] = gSum(mag(T - Told)) / gSum(T), change[1] = gSum(mag(rho - rhoOld)) / gSum(rho),
 * change[2] = gSum(mag(U - Uold)) / gSum(mag(U)) over the cells, where the "old" fields are the cell macros
 * at the previous call (at dugks_create for the first one); then Told = T, rhoOld = rho, Uold = U as in
 * the reference.  Replaces the D2H of three macro fields per check by one of 24 bytes; the macros are
 * replicated over ranks, so there is no collective.  A zero denominator gives the IEEE quotient, as in
 * the reference.  Synchronises. */
int dugks_convergence(dugks_handle_t* h, double change[3]);

/* gTildeVol/hTildeVol of one cell for all GLOBAL discrete velocities, gathered
 * over ranks (fvDVM::writeDFonCell, fvDVM.C:820-875; every rank must call).
 * g,h: [nXi] each, h_ may be NULL. */
int dugks_get_df(dugks_handle_t* h, int32_t cell, double* g, double* h_);

/* The rank-local slice of gTildeVol/hTildeVol, DV-major like the reference's
 * PtrList<discreteVelocity>: g[i*nCells + c] for local DV i.  For parity
 * tests and DF probes; they move the distribution functions only (use
 * dugks_checkpoint_save/load for a restart).  Either pointer may be NULL. */
int dugks_get_state(dugks_handle_t* h, double* g, double* h_);
int dugks_set_state(dugks_handle_t* h, const double* g, const double* h_);

/* Exact restart.  The reference's own restart is lossy: the distribution functions are NO_READ/NO_WRITE
 * (discreteVelocity.C:74-205), a restarted run re-initialises them to an equilibrium of the saved macros
 * (:220-249) and the wall density to 1 (calculatedMaxwellFvPatchField.C:80).  The blob holds everything the
 * next evolution() reads — gTilde/hTilde, boundary-face values (incoming half of "mixed" patches, :556-573),
 * the lagged boundary gradient (:462-468), cell / face / boundary macros with q and tau, wall constants, the
 * convergence monitor's old fields — so that save -> load -> step reproduces step bit for bit.  Rank-local and
 * layout-bound: load checks that case sizes, rank count and rank match.  buf is host memory. */
int dugks_checkpoint_size(dugks_handle_t* h, uint64_t* bytes);
int dugks_checkpoint_save(dugks_handle_t* h, void* buf, uint64_t bytes);
int dugks_checkpoint_load(dugks_handle_t* h, const void* buf, uint64_t bytes);

/* Host-only (no device needed): the global DV ids rank `rank` of `nRanks` owns under the
 * library's partition (contiguous blocks of the slow velocity indices; the reference's
 * round-robin is fvDVM.C:228-260 — any assignment is valid because the sums commute).
 * ids may be NULL to query the count. */
int dugks_partition(int32_t nXiPerDim, int32_t nSolutionD, int32_t nRanks, int32_t rank,
                    int32_t* ids, int32_t* n);

/* Host-only (no device needed): the velocity-row layout rank `rank` of `nRanks` uses for an
 * nXiPerDim^D velocity grid: rows of *L consecutive ix points, 32 rows per slab;
 * when *Lt > 0 the last slab holds short rows of *Lt points.  On entry *nRows is the capacity of the
 * four arrays (any of which may be NULL), on return the number of rows; row k covers
 * ix in [row_first[k], row_first[k] + row_len[k]) at (row_iy[k], row_iz[k]). */
int dugks_row_layout(int32_t nXiPerDim, int32_t nSolutionD, int32_t nRanks, int32_t rank,
                     int32_t* nch, int32_t* L, int32_t* Lt, int32_t* nRows,
                     int32_t* row_iy, int32_t* row_iz, int32_t* row_first, int32_t* row_len);

/* Host-only (no device needed): the order in which the cell kernels walk the mesh, a
 * permutation of the cells computed from the cell centres C [nCells][3].  kind: "wave" (what dugks_create
 * uses unless DUGKS_ORDER says otherwise), "tiled", "morton" or "natural"; nWarps: persistent warps in flight
 * (only "wave" uses it); first_class (may be NULL): cells with a non-zero entry come first, as the axis-only
 * launch of phase 1 needs them.  The reference walks cells in label order (discreteVelocity.C:346-410 and the
 * face loops :491-530, :948-956); any order gives the same result. */
int dugks_cell_order(int32_t nCells, int32_t nSolutionD, const double* C, const uint8_t* first_class,
                     const char* kind, int32_t nWarps, int32_t* order);

/* Host-only (no device needed): the CTA-pencil plan of phase 1 for a 3-D mesh: x-lines of
 * axis-aligned interior cells in 2 x 2 bundles, cut into work items for nCtas persistent CTAs.  On entry *nItems /
 * *nPencilCells are the capacities of item_first, item_steps / cells (any array may be NULL), on return the
 * counts.  cells: the pencil cells in traversal order — item k covers cells[item_first[k] .. + 4 item_steps[k]),
 * step-major, the four lines of the bundle within a step.  *nAxisCells (may be NULL): cells recognised as
 * axis-aligned (any dimension; exact zeros in the LS vectors and face offsets).  The reference visits cells in label order
 * (discreteVelocity.C:491-530); any order gives the same result. */
int dugks_pencil_plan(const dugks_mesh_t* mesh, int32_t nCtas, int32_t* nItems, int32_t* nPencilCells,
                      int32_t* cells, int32_t* item_first, int32_t* item_steps, int32_t* nAxisCells);

/* Boundary-face values gSurf/hSurf of the local DVs on all boundary faces,
 * g[j*nBoundaryFaces + b] for local DV j (DVi(j).gSurf().boundaryField(),
 * discreteVelocity.H:230-239).  Either pointer may be NULL. */
int dugks_get_boundary_df(dugks_handle_t* h, double* g, double* h_);

/* Global ids of the local DVs, in local order; returns count via *n.
 * ids may be NULL to query the count. */
int dugks_local_dvs(dugks_handle_t* h, int32_t* ids, int32_t* n);

/* nXi() (fvDVM.H:341) and problem sizes. */
int dugks_sizes(dugks_handle_t* h, int32_t* nXi, int32_t* nXiLocal,
                int32_t* nCells, int32_t* nFaces);

/* ---- instrumentation (not in the reference) ------------------------------ */

typedef struct dugks_stats_t {
    uint64_t kernel_launches;   /* kernels of this library launched so far  */
    uint64_t steps;
    uint64_t device_bytes;      /* device memory held by the handle         */
    int32_t  h_elided;          /* 1 when h is identically zero and elided  */
    int32_t  n_slabs;           /* DV slabs per phase                       */
    int32_t  slab_dvs;          /* DVs per slab                             */
    int32_t  keep_slabs;        /* slabs whose face values are kept between the two phases
                                   (fused relax+update; limited by device memory) */
    int32_t  pencil_cells;      /* cells phase 1 advances as CTA pencils (2 x 2 bundles of x-lines), 0 = none */
    int32_t  pencil_mode;       /* 0 off, 1 pencils over gBarP, 2 pencils that apply the half step themselves */
} dugks_stats_t;

int dugks_get_stats(dugks_handle_t* h, dugks_stats_t* out);

/* cudaStream_t the step is enqueued on (so a harness can record events on it). */
void* dugks_stream(dugks_handle_t* h);

/* Accumulated device time (ms, CUDA events) of the dominant kernel family
 * since the last reset, and its launch count; enable before the timed region.
 * which: 0 = cell_outgoing (gradient + reconstruction + face moments / flux),
 *        1 = cell_update, 2 = cell_halfstep, 3 = the all-reduces of the moment
 *        slots (fvDVM.C:363,487-489,519,626-628,725), waiting for the slowest
 *        rank included. */
int dugks_kernel_timing(dugks_handle_t* h, int enable, int which,
                        double* total_ms, uint64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* DUGKS_H */
