// dugks_capi.cu — extern "C" boundary (include/dugks.h) and host orchestration of the
// sm_100a kernels in dugks_kernels.cuh.  No CPU fallback: every entry point needs a
// CUDA device and fails loudly (DUGKS_ERR_NO_DEVICE) without one.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <string>
#include <vector>

#include "../../include/dugks.h"
#include "dugks_hot.cuh"
#include "dugks_pencil.cuh"
#include "dugks_pencil_ws.cuh"

// ------------------------------------------------------------------------------
// NCCL through dlopen (the library is present in every torch install and on the
// system; binding at run time keeps libdugks.so loadable on CPU-only boxes for the
// symbol checks).
typedef struct { char internal[128]; } nccl_uid_t;
typedef int (*nccl_get_uid_fn)(nccl_uid_t*);
typedef int (*nccl_init_rank_fn)(void** comm, int nranks, nccl_uid_t id, int rank);
typedef int (*nccl_allreduce_fn)(const void* send, void* recv, size_t count, int dtype, int op, void* comm,
                                 cudaStream_t stream);
typedef int (*nccl_destroy_fn)(void* comm);
typedef const char* (*nccl_errstr_fn)(int);

struct Nccl {
    void* lib = nullptr;
    nccl_get_uid_fn get_uid = nullptr;
    nccl_init_rank_fn init_rank = nullptr;
    nccl_allreduce_fn allreduce = nullptr;
    nccl_destroy_fn destroy = nullptr;
    nccl_errstr_fn errstr = nullptr;
    bool load(std::string& err) {
        if (lib) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so", nullptr};
        for (int i = 0; names[i] && !lib; i++) lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
        if (!lib) { err = std::string("cannot dlopen libnccl.so.2: ") + dlerror(); return false; }
        get_uid = (nccl_get_uid_fn)dlsym(lib, "ncclGetUniqueId");
        init_rank = (nccl_init_rank_fn)dlsym(lib, "ncclCommInitRank");
        allreduce = (nccl_allreduce_fn)dlsym(lib, "ncclAllReduce");
        destroy = (nccl_destroy_fn)dlsym(lib, "ncclCommDestroy");
        errstr = (nccl_errstr_fn)dlsym(lib, "ncclGetErrorString");
        if (!get_uid || !init_rank || !allreduce || !destroy) { err = "libnccl lacks required symbols"; return false; }
        return true;
    }
};
static Nccl g_nccl;

static thread_local std::string g_create_error = "no error";
#define COURANT_BLOCKS 592   // partial results of the two-stage Courant-number reduction

// ------------------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
};

struct dugks_handle {
    std::string err = "no error";
    int device = 0;
    cudaStream_t stream = nullptr;
    // sizes
    int nc = 0, nif = 0, nbf = 0, nf = 0, D = 3;
    int n1d = 0, nxi = 0;          // global DV grid
    int nch = 1, L = 0, Lt = 0, Rs = 32, nslab = 0, ntab = 0, tabw = 0;
    int nrows = 0;                 // padded local rows = nslab*Rs
    int nvl = 0;                   // local (real) DVs
    bool hasH = true;
    int nm = NM_MAX;
    std::vector<double> tx_host;   // [5][ntab] abscissae and moment weights of a row (DevDV::tx)
    bool pen_ws = false;           // phase 1 of the fused slabs by the warp-specialised pencil kernel
    size_t pen_ws_smem = 0;
    int rank = 0, nranks = 1;
    double xiMax = 0;
    DevGas gas{};
    std::vector<dugks_patch_t> patches;
    // local DV bookkeeping: for flat padded index k = s*L*Rs + i*Rs + r -> global id or -1
    std::vector<int> flat_gid;
    std::vector<int> local_gids;       // sorted global ids of local DVs
    std::vector<int> local_flat;       // flat index of local_gids[j]
    std::vector<int> owner_rank_of_gid;
    // device
    std::vector<DevBuf> bufs;
    uint64_t dev_bytes = 0;
    StepArgs A{};                      // template argument block (slab, dt patched per launch)
    double *gam_a_g = nullptr, *gam_a_h = nullptr, *gam_b_g = nullptr, *gam_b_h = nullptr;
    bool gam_flip = false;
    bool gam_single = false;   // one copy of the lagged boundary gradient, updated in place (second-generation kernels only)
    double *wall_cin = nullptr, *wall_in = nullptr;
    int* d_bc = nullptr;
    double* d_pres = nullptr;
    // symmetry patches (k_sym_pack / k_sym_apply): X = the reference's dfContainer (fvDVM.C:398-449)
    int *d_xrow = nullptr, *d_xmir = nullptr;   // [nflat], [3][nflat]: row of X of a local DV / of its x, y, z mirror DV
    int* d_symface = nullptr;                   // [nsym] boundary faces of all symmetry patches
    double *sym_Xg = nullptr, *sym_Xh = nullptr;
    int nsym = 0, sym_rows = 0;
    bool sym_exchange = false;                  // mirror partners on other ranks: X is indexed by global DV id and all-reduced
    std::vector<SymPatch> sym_patches;
    double* d_co = nullptr;
    double *d_conv_old = nullptr, *d_conv = nullptr;   // convergence monitor: snapshot [nc][5], sums [6] + partials
    double* d_bstage = nullptr;        // [5 nbf] staging of dugks_set_boundary_macros
    double* fslot_fold = nullptr;      // [nf][nm] sharded runs: face moment slots with both sides added (what the all-reduce carries)
    bool has_sym = false, has_wall = false;
    size_t nflat = 0;                  // nslab*L*Rs
    // collective
    dugks_allreduce_fn reduce = nullptr;
    void* reduce_user = nullptr;
    void* nccl_comm = nullptr;
    // host side of the accessors: pinned staging buffer, last boundary macros the caller set
    double* pin = nullptr;
    size_t pin_count = 0;
    double* d_soa = nullptr;       // accessor staging on the device: the macro fields in the caller's (per-field) layout
    size_t soa_count = 0;
    std::vector<double> last_rho_b, last_U_b, last_T_b;
    // stats
    uint64_t launches = 0, steps = 0;
    // kernel timing
    bool timing = false;
    struct Ev { cudaEvent_t a, b; int which; };
    std::vector<Ev> events;
    std::vector<Ev> pool;
    size_t smem_out1 = 0, smem_out2 = 0, smem_bnd = 0, smem_upd = 0;
    size_t fsmem_out1 = 0, fsmem_out2 = 0, fsmem_upd = 0;
    int n_big = 0;          // cells with more than FAST_NE faces (generic kernels)
    bool use_fast = true;
    bool use_tma = true;    // bulk-async staged kernels (dugks_tma.cuh)
    int ci = 4, max_ne_fast = 0, tma_tw = 32;
    size_t tsmem_out1 = 0, tsmem_out2 = 0, tsmem_upd = 0;
    // second-generation kernels (dugks_hot.cuh)
    bool use_hot = true, has_far = false;
    int hot_ne = 6, hot_grid_out1 = 148, hot_grid_out2 = 148, hot_grid_upd = 148, hot_grid_rlx = 148;
    size_t hsmem_rlx = 0, hsmem_half = 0;
    int hot_grid_half = 148, hot_grid_axis = 148;
    bool split_axis = false, want_split = false;
    // gBarP storage: one block per slab, or (wmode) ONE transient block shared by the face-storage slabs
    // (their update reads w, left in place of gTilde by the half-step kernel) + one block per other slab
    double *gb_store = nullptr, *hb_store = nullptr;
    bool wmode = false;
    size_t hsmem_rlx_w = 0;
    int hot_grid_rlx_w = 148;
    size_t hsmem_axis = 0;
    std::vector<char> slab_pair_ok;   // per slab: the axis-only launch of phase 1 is valid (hot_axis_item)
    int n_axis = 0, axis_ne = 0;
    // CTA pencils of phase 1 (dugks_pencil.cuh): 2 x 2 bundles of x-lines of axis-aligned interior cells
    int pen_mode = 0;                  // 0 off, 1 pencil reads gBarP, 2 fused: the pencil applies the half step itself
    int n_pen = 0;                     // traversal items [0, n_pen) are pencil cells, [n_pen, n_axis) the other axis-aligned cells
    PenArgs pen{};
    int pen_grid = 0;
    size_t pen_smem = 0;
    int* d_halflist = nullptr;         // fused mode: cells whose gBarP the other kernels still read (non-pencil cells + their neighbours)
    int n_halflist = 0;
    size_t hsmem_rlx_g = 0;            // relax+update that forms w from gTilde (WMODE 2)
    int hot_grid_rlx_g = 148;
    // face-storage slabs: phase 1 keeps the reconstructed face values of slabs [0, n_keep) so that
    // phase 2 is ONE fused relax+update pass for them (no second gradient, no flux-buffer round trip)
    int n_keep = 0;
    double *fkeep_g = nullptr, *fkeep_h = nullptr;
    size_t hsmem_out1 = 0, hsmem_out2 = 0, hsmem_upd = 0;
};

static int fail(dugks_handle* h, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf; else g_create_error = buf;
    return code;
}

#define CUDA_TRY(h, call)                                                                      \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return fail(h, (e__ == cudaErrorMemoryAllocation) ? DUGKS_ERR_NOMEM : DUGKS_ERR_CUDA, \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

template <class T>
static int dev_alloc(dugks_handle* h, T** out, size_t count, bool zero = true) {
    size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess)
        return fail(h, DUGKS_ERR_NOMEM, "cudaMalloc of %zu bytes failed: %s (device memory held so far: %llu bytes)",
                    bytes, cudaGetErrorString(e), (unsigned long long)h->dev_bytes);
    if (zero) {
        e = cudaMemsetAsync(p, 0, bytes, h->stream);
        if (e != cudaSuccess) return fail(h, DUGKS_ERR_CUDA, "cudaMemset failed: %s", cudaGetErrorString(e));
    }
    h->bufs.push_back({p, bytes});
    h->dev_bytes += bytes;
    *out = (T*)p;
    return 0;
}

template <class T>
static int dev_upload(dugks_handle* h, T** out, const std::vector<T>& v) {
    int rc = dev_alloc(h, out, v.size(), false);
    if (rc) return rc;
    if (!v.empty()) CUDA_TRY(h, cudaMemcpyAsync(*out, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));   // v may be a temporary
    return 0;
}

// ------------------------------------------------------------------------------
// launch helpers
static int grid_for(long long items) {
    long long ctas = (items + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
    const long long cap = 148LL * 16 * 8;   // multiple of the SM count; grid-stride loops cover the rest
    return (int)std::max<long long>(1, std::min(ctas, cap));
}

struct Timed {
    dugks_handle* h; int which; dugks_handle::Ev ev{};
    bool on;
    Timed(dugks_handle* h_, int which_) : h(h_), which(which_), on(h_->timing) {
        if (on) {
            if (!h->pool.empty()) { ev = h->pool.back(); h->pool.pop_back(); }
            else { cudaEventCreate(&ev.a); cudaEventCreate(&ev.b); }
            ev.which = which;
            cudaEventRecord(ev.a, h->stream);
        }
    }
    ~Timed() {
        if (on) { cudaEventRecord(ev.b, h->stream); h->events.push_back(ev); }
    }
};

static int check_launch(dugks_handle* h, const char* name) {
    h->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(h, DUGKS_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(e));
    return 0;
}

static int do_allreduce(dugks_handle* h, double* buf, size_t n) {
    if (h->nranks <= 1) return 0;
    Timed t(h, 3);   // dugks_kernel_timing class 3: the collectives (incl. waiting for the slowest rank)
    if (h->reduce) {
        int rc = h->reduce(h->reduce_user, buf, n, (void*)h->stream);
        if (rc) return fail(h, DUGKS_ERR_COMM, "user allreduce callback returned %d", rc);
        return 0;
    }
    if (!h->nccl_comm) return fail(h, DUGKS_ERR_COMM, "nRanks > 1 but no reducer and no NCCL communicator");
    int rc = g_nccl.allreduce(buf, buf, n, /*ncclDouble*/ 8, /*ncclSum*/ 0, h->nccl_comm, h->stream);
    if (rc) return fail(h, DUGKS_ERR_COMM, "ncclAllReduce failed: %s", g_nccl.errstr ? g_nccl.errstr(rc) : "?");
    return 0;
}

// ---- second-generation kernels: NE (faces staged per cell) and TW (equilibrium table stride)
// are compile-time; create() picks the smallest that fits the mesh / velocity layout.
// Points per chunk: 4 where the moment accumulators set the register count anyway (2 CTAs/SM),
// 2 for the table-heavy phase-2 kernels so that they run 3 CTAs/SM (profiles/r01_occupancy_ab.txt)
#ifndef DUGKS_CI_OUT1
#define DUGKS_CI_OUT1 4
#endif
#ifndef DUGKS_CI_RLX
#define DUGKS_CI_RLX 2
#endif
constexpr int CI_OUT1 = DUGKS_CI_OUT1, CI_OUT2 = 2, CI_RLX = DUGKS_CI_RLX;
// cells with up to 8 faces (polygonal meshes): 9 staged streams per field; with 4 points per chunk one CTA fills an SM's
// shared memory (154 KB with h), with 2 points two fit
#define CI_OUT1_NE(NE_) ((NE_) == 8 ? 2 : CI_OUT1)
// axis-only launch of phase 1 (hot_axis_item): chunks of 4 points are staged, HOT_AXIS_CU = 2 points are advanced
// together, which keeps it at 3 CTAs/SM without spills
#ifndef DUGKS_CI_AXIS
#define DUGKS_CI_AXIS 4
#endif
constexpr int CI_AXIS = DUGKS_CI_AXIS;
// update kernel of the flux-buffer path: 4 points per chunk, 2 when h doubles the streams (shared memory per CTA)
#define CI_UPD (H ? 2 : 4)
#ifndef DUGKS_PENCIL_WS_DEFAULT
#define DUGKS_PENCIL_WS_DEFAULT 0
#endif
// DUGKS_DEV_BUILD (experiment builds, profiles/ab_bench.py): only the instantiations of the 3-D, h-elided case with
// unchunked rows are compiled (a tenth of the build time); every other case fails with DUGKS_ERR_UNSUPPORTED.
#ifdef DUGKS_DEV_BUILD
#define DUGKS_BY_H(h, fn, ...) ((h)->hasH ? fail((h), DUGKS_ERR_UNSUPPORTED, "DUGKS_DEV_BUILD holds the h-elided kernels only") : fn<false>(__VA_ARGS__))
#else
#define DUGKS_BY_H(h, fn, ...) ((h)->hasH ? fn<true>(__VA_ARGS__) : fn<false>(__VA_ARGS__))
#endif
// How phase 1 of a slab treats the pencil cells: 0 = no pencil (warp-per-cell kernels), 1 = pencil over gBarP
// written by the half-step kernel, 2 = fused: the pencil reads gTilde and applies the half step itself (face-
// storage slabs only: the recompute path needs gBarP of every cell again in phase 2).
static int pencil_mode(const dugks_handle* h, int slab) {
    if (!h->pen_mode || !h->split_axis || slab >= (int)h->slab_pair_ok.size() || !h->slab_pair_ok[slab]) return 0;
    return (h->pen_mode == 2 && slab < h->n_keep) ? 2 : 1;
}

template <int PHASE, bool H>
static void launch_hot_outgoing(dugks_handle* h, const StepArgs& a) {
    const size_t sm = PHASE == 1 ? h->hsmem_out1 : h->hsmem_out2;
    const int tw = PHASE == 1 ? 32 : h->tma_tw;
    const int grid = PHASE == 1 ? h->hot_grid_out1 : h->hot_grid_out2;
    if (PHASE == 1 && h->split_axis && a.slab < (int)h->slab_pair_ok.size() && h->slab_pair_ok[a.slab]) {
        // mostly axis-aligned mesh: the light variant (hot_axis_item) alone runs 3 CTAs/SM, a second launch
        // takes the rest.  Slabs with a tie on a y/z face or rows that are not whole chunks take the unified launch.
        StepArgs a1 = a, a2 = a;
        const int pm = pencil_mode(h, a.slab);
        a1.item0 = pm ? h->n_pen : 0; a1.item1 = h->n_axis;
        a2.item0 = h->n_axis; a2.item1 = h->nc;
        const int grid2 = std::max(1, std::min(grid, (h->nc - h->n_axis + HOT_WARPS - 1) / HOT_WARPS));
        const int grid1 = std::max(1, std::min(h->hot_grid_axis, (a1.item1 - a1.item0 + HOT_WARPS - 1) / HOT_WARPS));
        if (!H && pm) {
            // CTA pencils (dugks_pencil.cuh) take the bundled x-lines; the axis-only launch what is left of the axis-aligned cells
            if (pm == 2 && h->pen_ws) k_pencil_ws<<<h->pen_grid, 2 * PEN_WARPS * 32, h->pen_ws_smem, h->stream>>>(a, h->pen);
            else if (pm == 2) k_pencil_phase1<true><<<h->pen_grid, PEN_WARPS * 32, h->pen_smem, h->stream>>>(a, h->pen);
            else k_pencil_phase1<false><<<h->pen_grid, PEN_WARPS * 32, h->pen_smem, h->stream>>>(a, h->pen);
            h->launches++;
        }
#ifndef DUGKS_DEV_BUILD
        if (h->hot_ne == 4) {
            if (a1.item0 < a1.item1) k_hot_outgoing<1, H, 4, 32, CI_AXIS, 1><<<grid1, HOT_WARPS * 32, h->hsmem_axis, h->stream>>>(a1);
            if (h->n_axis < h->nc) k_hot_outgoing<1, H, 4, 32, CI_OUT1, 2><<<grid2, HOT_WARPS * 32, sm, h->stream>>>(a2);
        } else
#endif
        {
            if (a1.item0 < a1.item1) k_hot_outgoing<1, H, 6, 32, CI_AXIS, 1><<<grid1, HOT_WARPS * 32, h->hsmem_axis, h->stream>>>(a1);
            if (h->n_axis < h->nc) k_hot_outgoing<1, H, 6, 32, CI_OUT1, 2><<<grid2, HOT_WARPS * 32, sm, h->stream>>>(a2);
        }
        h->launches++;
        return;
    }
#define DUGKS_HOT_OUT(NE_, TW_) k_hot_outgoing<PHASE, H, NE_, TW_, (PHASE == 1 ? CI_OUT1_NE(NE_) : CI_OUT2), 0><<<grid, HOT_WARPS * 32, sm, h->stream>>>(a)
#ifdef DUGKS_DEV_BUILD
    DUGKS_HOT_OUT(6, 32);
#else
    if (h->hot_ne == 4) { if (tw == 32) DUGKS_HOT_OUT(4, 32); else DUGKS_HOT_OUT(4, 64); }
    else if (h->hot_ne == 6) { if (tw == 32) DUGKS_HOT_OUT(6, 32); else DUGKS_HOT_OUT(6, 64); }
    else { if (tw == 32) DUGKS_HOT_OUT(8, 32); else DUGKS_HOT_OUT(8, 64); }
#endif
#undef DUGKS_HOT_OUT
}
template <bool H>
static void launch_hot_update(dugks_handle* h, const StepArgs& a) {
#ifndef DUGKS_DEV_BUILD
    if (h->hot_ne == 4) k_hot_update<H, 4, CI_UPD><<<h->hot_grid_upd, HOT_WARPS * 32, h->hsmem_upd, h->stream>>>(a);
    else
#endif
    if (h->hot_ne == 6) k_hot_update<H, 6, CI_UPD><<<h->hot_grid_upd, HOT_WARPS * 32, h->hsmem_upd, h->stream>>>(a);
#ifndef DUGKS_DEV_BUILD
    else k_hot_update<H, 8, CI_UPD><<<h->hot_grid_upd, HOT_WARPS * 32, h->hsmem_upd, h->stream>>>(a);
#endif
}
template <bool H>
static void launch_hot_relax(dugks_handle* h, const StepArgs& a) {
    const int tw = h->tma_tw;
#define DUGKS_HOT_RLX(NE_, TW_)                                                                                            \
    do {                                                                                                                   \
        if (pencil_mode(h, a.slab) == 2) k_hot_relax_update<H, NE_, TW_, CI_RLX, 2><<<h->hot_grid_rlx_g, HOT_WARPS * 32, h->hsmem_rlx_g, h->stream>>>(a); \
        else if (h->wmode) k_hot_relax_update<H, NE_, TW_, CI_RLX, 1><<<h->hot_grid_rlx_w, HOT_WARPS * 32, h->hsmem_rlx_w, h->stream>>>(a); \
        else k_hot_relax_update<H, NE_, TW_, CI_RLX, 0><<<h->hot_grid_rlx, HOT_WARPS * 32, h->hsmem_rlx, h->stream>>>(a);             \
    } while (0)
#ifdef DUGKS_DEV_BUILD
    DUGKS_HOT_RLX(6, 32);
#else
    if (h->hot_ne == 4) { if (tw == 32) DUGKS_HOT_RLX(4, 32); else DUGKS_HOT_RLX(4, 64); }
    else if (h->hot_ne == 6) { if (tw == 32) DUGKS_HOT_RLX(6, 32); else DUGKS_HOT_RLX(6, 64); }
    else { if (tw == 32) DUGKS_HOT_RLX(8, 32); else DUGKS_HOT_RLX(8, 64); }
#endif
#undef DUGKS_HOT_RLX
}
template <bool H, int NE, int TW>
static cudaError_t hot_cfg_rlx(dugks_handle* h, int* occ) {
    h->hsmem_rlx = HotRelaxPlan<H, NE, TW, CI_RLX, 0>::total(h->ntab);
    h->hsmem_rlx_w = HotRelaxPlan<H, NE, TW, CI_RLX, 1>::total(h->ntab);
    h->hsmem_rlx_g = HotRelaxPlan<H, NE, TW, CI_RLX, 2>::total(h->ntab);
    cudaError_t e = cudaFuncSetAttribute(k_hot_relax_update<H, NE, TW, CI_RLX, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->hsmem_rlx);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, k_hot_relax_update<H, NE, TW, CI_RLX, 0>, HOT_WARPS * 32, h->hsmem_rlx);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_hot_relax_update<H, NE, TW, CI_RLX, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->hsmem_rlx_w);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ + 1, k_hot_relax_update<H, NE, TW, CI_RLX, 1>, HOT_WARPS * 32, h->hsmem_rlx_w);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_hot_relax_update<H, NE, TW, CI_RLX, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->hsmem_rlx_g);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ + 2, k_hot_relax_update<H, NE, TW, CI_RLX, 2>, HOT_WARPS * 32, h->hsmem_rlx_g);
    return e;
}
template <int PHASE, bool H, int NE, int TW>
static cudaError_t hot_attr_out(size_t bytes) {
    return cudaFuncSetAttribute(k_hot_outgoing<PHASE, H, NE, TW, (PHASE == 1 ? CI_OUT1_NE(NE) : CI_OUT2), 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}
template <bool H>
static int hot_configure(dugks_handle* h) {
    const int ntab = h->ntab, tw = h->tma_tw;
    cudaError_t e = cudaSuccess;
    int occ[6] = {1, 1, 1, 1, 1, 1};
#define DUGKS_HOT_CFG(NE_)                                                                                   \
    do {                                                                                                     \
        h->hsmem_out1 = HotPlan<1, H, NE_, 32, CI_OUT1_NE(NE_)>::total(ntab);                                                 \
        h->hsmem_out2 = tw == 32 ? HotPlan<2, H, NE_, 32, CI_OUT2>::total(ntab) : HotPlan<2, H, NE_, 64, CI_OUT2>::total(ntab); \
        h->hsmem_upd = HotUpdPlan<H, NE_, CI_UPD>::total(ntab);                                                      \
        if (std::max(std::max(h->hsmem_out1, h->hsmem_out2), h->hsmem_upd) > 220 * 1024) { h->use_hot = false; return 0; } \
        e = hot_attr_out<1, H, NE_, 32>(h->hsmem_out1);                                                      \
        if (e == cudaSuccess) e = tw == 32 ? hot_attr_out<2, H, NE_, 32>(h->hsmem_out2) : hot_attr_out<2, H, NE_, 64>(h->hsmem_out2); \
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_hot_update<H, NE_, CI_UPD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->hsmem_upd); \
        /* persistent grids: CTAs the SM can hold (registers and shared memory) times the SM count */    \
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[0], k_hot_outgoing<1, H, NE_, 32, CI_OUT1_NE(NE_), 0>, HOT_WARPS * 32, h->hsmem_out1); \
        if (e == cudaSuccess) e = tw == 32 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[1], k_hot_outgoing<2, H, NE_, 32, CI_OUT2, 0>, HOT_WARPS * 32, h->hsmem_out2) \
                                           : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[1], k_hot_outgoing<2, H, NE_, 64, CI_OUT2, 0>, HOT_WARPS * 32, h->hsmem_out2); \
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[2], k_hot_update<H, NE_, CI_UPD>, HOT_WARPS * 32, h->hsmem_upd); \
        if (e == cudaSuccess) e = tw == 32 ? hot_cfg_rlx<H, NE_, 32>(h, &occ[3]) : hot_cfg_rlx<H, NE_, 64>(h, &occ[3]); \
    } while (0)
#ifdef DUGKS_DEV_BUILD
    if (h->hot_ne != 6 || tw != 32) return fail(h, DUGKS_ERR_UNSUPPORTED, "DUGKS_DEV_BUILD holds the 3-D hex, h-elided, unchunked-row kernels only");
    DUGKS_HOT_CFG(6);
#else
    if (h->hot_ne == 4) DUGKS_HOT_CFG(4);
    else if (h->hot_ne == 6) DUGKS_HOT_CFG(6);
    else DUGKS_HOT_CFG(8);
#endif
#undef DUGKS_HOT_CFG
    if (e != cudaSuccess) return fail(h, DUGKS_ERR_CUDA, "cudaFuncSetAttribute (hot kernels): %s", cudaGetErrorString(e));
    int occ_axis = 1;
    h->split_axis = false;
    if (e == cudaSuccess && (h->hot_ne == 4 || h->hot_ne == 6) && h->want_split && h->axis_ne == h->hot_ne) {
        // axis-aligned cells in their own launch (hot_axis_item, 3 CTAs/SM; 64^3 x 28^3: phase 1 2.90 -> 2.13 ms per slab)
#ifndef DUGKS_DEV_BUILD
        if (h->hot_ne == 4) {
            h->hsmem_axis = HotPlan<1, H, 4, 32, CI_AXIS>::total(ntab);
            e = cudaFuncSetAttribute(k_hot_outgoing<1, H, 4, 32, CI_AXIS, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->hsmem_axis);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k_hot_outgoing<1, H, 4, 32, CI_OUT1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->hsmem_out1);
            if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_axis, k_hot_outgoing<1, H, 4, 32, CI_AXIS, 1>, HOT_WARPS * 32, h->hsmem_axis);
        } else
#endif
        {
            h->hsmem_axis = HotPlan<1, H, 6, 32, CI_AXIS>::total(ntab);
            e = cudaFuncSetAttribute(k_hot_outgoing<1, H, 6, 32, CI_AXIS, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->hsmem_axis);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k_hot_outgoing<1, H, 6, 32, CI_OUT1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->hsmem_out1);
            if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_axis, k_hot_outgoing<1, H, 6, 32, CI_AXIS, 1>, HOT_WARPS * 32, h->hsmem_axis);
        }
        if (e != cudaSuccess) return fail(h, DUGKS_ERR_CUDA, "axis-split kernel configuration: %s", cudaGetErrorString(e));
        h->split_axis = occ_axis >= 3;
    }
    int occ_half = 1;
    h->hsmem_half = HotHalfPlan<H>::total(h->L, tw, ntab);
    if (h->hsmem_half <= 200 * 1024) {
        e = cudaFuncSetAttribute(k_hot_halfstep<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->hsmem_half);
        if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_half, k_hot_halfstep<H>, HOT_WARPS * 32, h->hsmem_half);
        if (e != cudaSuccess) return fail(h, DUGKS_ERR_CUDA, "k_hot_halfstep configuration: %s", cudaGetErrorString(e));
    } else h->hsmem_half = 0;
    int dev_sms = 148;
    cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, h->device);
    const int max_ctas = (h->nc + HOT_WARPS - 1) / HOT_WARPS;
    h->hot_grid_out1 = std::max(1, std::min(dev_sms * std::max(occ[0], 1), max_ctas));
    h->hot_grid_out2 = std::max(1, std::min(dev_sms * std::max(occ[1], 1), max_ctas));
    h->hot_grid_upd = std::max(1, std::min(dev_sms * std::max(occ[2], 1), max_ctas));
    if (const char* e = getenv("DUGKS_RLX_CTAS"))   // experiment hook
        for (int k = 3; k < 6; k++) occ[k] = std::max(1, std::min(occ[k], atoi(e)));
    if (const char* e = getenv("DUGKS_UPD_CTAS")) occ[2] = std::max(1, std::min(occ[2], atoi(e)));
    h->hot_grid_upd = std::max(1, std::min(dev_sms * std::max(occ[2], 1), max_ctas));
    h->hot_grid_rlx = std::max(1, std::min(dev_sms * std::max(occ[3], 1), max_ctas));
    h->hot_grid_rlx_w = std::max(1, std::min(dev_sms * std::max(occ[4], 1), max_ctas));
    h->hot_grid_rlx_g = std::max(1, std::min(dev_sms * std::max(occ[5], 1), max_ctas));
    // CTA pencils: the shared-memory window has to fit twice per SM to be worth it
    if (h->pen_mode) {
        int occ_pen = 0;
        h->pen_smem = PenPlan::total(h->L, ntab, h->tabw);
        if (H || h->pen_smem > 113 * 1024 || !h->split_axis) h->pen_mode = 0;
        else {
            e = cudaFuncSetAttribute(k_pencil_phase1<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->pen_smem);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(k_pencil_phase1<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->pen_smem);
            if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_pen, k_pencil_phase1<true>, PEN_WARPS * 32, h->pen_smem);
            if (e != cudaSuccess) return fail(h, DUGKS_ERR_CUDA, "pencil kernel configuration: %s", cudaGetErrorString(e));
            if (occ_pen < 1) h->pen_mode = 0;
            // warp-specialised variant (dugks_pencil_ws.cuh): producer + consumer warp per line, same shared-memory budget
            h->pen_ws = false;
            const char* ws_env = getenv("DUGKS_PENCIL_WS");
            const int min_len = h->Lt > 0 ? std::min(h->L, h->Lt) : h->L;
            if (h->pen_mode == 2 && (ws_env ? atoi(ws_env) != 0 : DUGKS_PENCIL_WS_DEFAULT) && min_len >= PWS_STAGES * PWS_CH && min_len % PWS_CH == 0) {
                h->pen_ws_smem = PwsPlan::total(h->L, ntab, h->tabw);
                int occ_ws = 0;
                e = cudaFuncSetAttribute(k_pencil_ws, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->pen_ws_smem);
                if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_ws, k_pencil_ws, 2 * PEN_WARPS * 32, h->pen_ws_smem);
                if (e != cudaSuccess) { (void)cudaGetLastError(); occ_ws = 0; }
                h->pen_ws = occ_ws >= 2;
                if (getenv("DUGKS_VERBOSE"))
                    fprintf(stderr, "dugks: warp-specialised pencils: %zu B shared memory per CTA, %d CTAs/SM -> %s\n", h->pen_ws_smem, occ_ws, h->pen_ws ? "on" : "off");
            }
        }
        if (getenv("DUGKS_VERBOSE"))
            fprintf(stderr, "dugks: pencils mode %d, %d of %d axis-aligned cells in %d work items, %zu B shared memory per CTA, %d CTAs/SM, %d cells in the half-step list\n",
                    h->pen_mode, h->n_pen, h->n_axis, h->pen.nitems, h->pen_smem, occ_pen, h->n_halflist);
    }
    h->hot_grid_half = std::max(1, std::min(dev_sms * std::max(occ_half, 1), max_ctas));
    h->hot_grid_axis = std::max(1, std::min(dev_sms * std::max(occ_axis, 1), max_ctas));
    if (getenv("DUGKS_VERBOSE"))
        fprintf(stderr, "dugks: rows of %d points, %d slabs, short-row tail slab: %d points per row\n", h->L, h->nslab, h->Lt);
    if (getenv("DUGKS_VERBOSE"))
        fprintf(stderr, "dugks: hot kernels NE=%d smem %zu/%zu/%zu/%zu B, CTAs per SM %d/%d/%d/%d, axis cells %d of %d, split %d (%d CTAs/SM)\n", h->hot_ne, h->hsmem_out1,
                h->hsmem_out2, h->hsmem_upd, h->hsmem_rlx, occ[0], occ[1], occ[2], occ[3], h->n_axis, h->nc, (int)h->split_axis, occ_axis);
    return 0;
}

// Where the gBarP block of slab a.slab lives: every kernel addresses it as a.gb + slab * nCells * L * 32,
// so the base pointer is shifted instead of touching the kernels.
static void map_gb(const dugks_handle* h, StepArgs& a) {
    if (!h->wmode) return;
    const long long stride = (long long)h->nc * h->L * h->Rs;
    if (a.slab < h->n_keep) {
        // transient block of the face-storage slabs: block 0 of the store
        a.gb = h->gb_store - a.slab * stride;
        if (h->hb_store) a.hb = h->hb_store - a.slab * stride;
    } else {
        const long long block = 1 + (a.slab - h->n_keep);
        a.gb = h->gb_store + (block - a.slab) * stride;
        if (h->hb_store) a.hb = h->hb_store + (block - a.slab) * stride;
    }
}

template <bool H>
static int launch_slab_kernels_phase1(dugks_handle* h, StepArgs a) {
    int rc;
    long long items = (long long)h->nc * (h->Rs / 32);
    map_gb(h, a);
    {
        Timed t(h, 2);
        if (h->use_hot && h->hsmem_half > 0) {
            if (pencil_mode(h, a.slab) == 2)   // only gBarP of the cells the warp-per-cell kernels still read; gTilde stays
                k_hot_halfstep<H><<<std::max(1, std::min(h->hot_grid_half, (h->n_halflist + HOT_WARPS - 1) / HOT_WARPS)), HOT_WARPS * 32, h->hsmem_half, h->stream>>>(
                    a, h->tma_tw, 0, h->d_halflist, h->n_halflist);
            else
                k_hot_halfstep<H><<<h->hot_grid_half, HOT_WARPS * 32, h->hsmem_half, h->stream>>>(a, h->tma_tw, (h->wmode && a.slab < h->n_keep) ? 1 : 0, nullptr, 0);
        } else
            k_cell_halfstep<H><<<grid_for(items), WARPS_PER_CTA * 32, 0, h->stream>>>(a, 0);
    }
    if ((rc = check_launch(h, "k_cell_halfstep"))) return rc;
    if (h->use_hot) {
        {
            Timed t(h, 0);
            const bool keep = a.slab < h->n_keep;
            a.fkeep_g = keep ? h->fkeep_g : nullptr;
            a.fkeep_h = keep ? h->fkeep_h : nullptr;
            launch_hot_outgoing<1, H>(h, a);
        }
        if ((rc = check_launch(h, "k_hot_outgoing<1>"))) return rc;
        if (h->n_big > 0) {
            Timed t(h, 0);
            a.skip_small = 1;
            k_cell_outgoing<1, H><<<grid_for(items), WARPS_PER_CTA * 32, h->smem_out1, h->stream>>>(a);
            if ((rc = check_launch(h, "k_cell_outgoing<1>"))) return rc;
        }
        if (h->nbf > 0 && (h->has_far || h->n_big > 0)) {
            // incoming half of far-field patches; every boundary face of cells the generic kernel took
            long long bitems = (long long)h->nbf * (h->Rs / 32);
            k_bnd_outgoing<H><<<grid_for(bitems), WARPS_PER_CTA * 32, h->smem_bnd, h->stream>>>(a, h->n_big > 0 ? 0 : 1);
            if ((rc = check_launch(h, "k_bnd_outgoing"))) return rc;
        }
        return 0;
    }
    if (h->use_fast && h->use_tma) {
        Timed t(h, 0);
        k_cell_outgoing_tma<1, H, TMA_CI, 32><<<grid_for(items), WARPS_PER_CTA * 32, h->tsmem_out1, h->stream>>>(a, h->max_ne_fast + 1);
        if ((rc = check_launch(h, "k_cell_outgoing_tma<1>"))) return rc;
    } else if (h->use_fast) {
        Timed t(h, 0);
        k_cell_outgoing_fast<1, H><<<grid_for(items), WARPS_PER_CTA * 32, h->fsmem_out1, h->stream>>>(a);
        if ((rc = check_launch(h, "k_cell_outgoing_fast<1>"))) return rc;
    }
    if (!h->use_fast || h->n_big > 0) {
        Timed t(h, 0);
        a.skip_small = h->use_fast ? 1 : 0;
        k_cell_outgoing<1, H><<<grid_for(items), WARPS_PER_CTA * 32, h->smem_out1, h->stream>>>(a);
        if ((rc = check_launch(h, "k_cell_outgoing<1>"))) return rc;
    }
    if (h->nbf > 0) {
        long long bitems = (long long)h->nbf * (h->Rs / 32);
        k_bnd_outgoing<H><<<grid_for(bitems), WARPS_PER_CTA * 32, h->smem_bnd, h->stream>>>(a, 0);
        if ((rc = check_launch(h, "k_bnd_outgoing"))) return rc;
    }
    return 0;
}

template <bool H>
static int launch_slab_kernels_phase2(dugks_handle* h, StepArgs a) {
    int rc;
    long long items = (long long)h->nc * (h->Rs / 32);
    map_gb(h, a);
    if (h->nbf > 0) {
        long long bitems = (long long)h->nbf * (h->Rs / 32);
        if (h->use_hot) {
            const size_t sm = ((size_t)((h->ntab + 1) & ~1) + (size_t)WARPS_PER_CTA * h->tma_tw * 4) * sizeof(double);
            k_hot_bnd_relax<H><<<grid_for(bitems), WARPS_PER_CTA * 32, sm, h->stream>>>(a, h->tma_tw);
        } else
            k_bnd_relax<H><<<grid_for(bitems), WARPS_PER_CTA * 32, 0, h->stream>>>(a);
        if ((rc = check_launch(h, "k_bnd_relax"))) return rc;
    }
    if (h->use_hot && a.slab < h->n_keep) {
        {
            Timed t(h, 1);
            a.fkeep_g = h->fkeep_g; a.fkeep_h = h->fkeep_h;
            launch_hot_relax<H>(h, a);
        }
        if ((rc = check_launch(h, "k_hot_relax_update"))) return rc;
        return 0;
    }
    if (h->use_hot) {
        {
            Timed t(h, 0);
            launch_hot_outgoing<2, H>(h, a);
        }
        if ((rc = check_launch(h, "k_hot_outgoing<2>"))) return rc;
        if (h->n_big > 0) {
            Timed t(h, 0);
            a.skip_small = 1;
            k_cell_outgoing<2, H><<<grid_for(items), WARPS_PER_CTA * 32, h->smem_out2, h->stream>>>(a);
            if ((rc = check_launch(h, "k_cell_outgoing<2>"))) return rc;
        }
        {
            Timed t(h, 1);
            launch_hot_update<H>(h, a);
        }
        if ((rc = check_launch(h, "k_hot_update"))) return rc;
        if (h->n_big > 0) {
            Timed t(h, 1);
            a.skip_small = 1;
            k_cell_update<H><<<grid_for(items), WARPS_PER_CTA * 32, h->smem_upd, h->stream>>>(a);
            if ((rc = check_launch(h, "k_cell_update"))) return rc;
        }
        return 0;
    }
    if (h->use_fast && h->use_tma) {
        Timed t(h, 0);
        if (h->tma_tw == 32)
            k_cell_outgoing_tma<2, H, TMA_CI, 32><<<grid_for(items), WARPS_PER_CTA * 32, h->tsmem_out2, h->stream>>>(a, h->max_ne_fast + 1);
        else
            k_cell_outgoing_tma<2, H, TMA_CI, 64><<<grid_for(items), WARPS_PER_CTA * 32, h->tsmem_out2, h->stream>>>(a, h->max_ne_fast + 1);
        if ((rc = check_launch(h, "k_cell_outgoing_tma<2>"))) return rc;
    } else if (h->use_fast) {
        Timed t(h, 0);
        k_cell_outgoing_fast<2, H><<<grid_for(items), WARPS_PER_CTA * 32, h->fsmem_out2, h->stream>>>(a);
        if ((rc = check_launch(h, "k_cell_outgoing_fast<2>"))) return rc;
    }
    if (!h->use_fast || h->n_big > 0) {
        Timed t(h, 0);
        a.skip_small = h->use_fast ? 1 : 0;
        k_cell_outgoing<2, H><<<grid_for(items), WARPS_PER_CTA * 32, h->smem_out2, h->stream>>>(a);
        if ((rc = check_launch(h, "k_cell_outgoing<2>"))) return rc;
    }
    if (h->use_fast && h->use_tma) {
        Timed t(h, 1);
        k_cell_update_tma<H><<<grid_for(items), WARPS_PER_CTA * 32, h->tsmem_upd, h->stream>>>(a, h->ci, h->max_ne_fast + 2);
        if ((rc = check_launch(h, "k_cell_update_tma"))) return rc;
    } else if (h->use_fast) {
        Timed t(h, 1);
        k_cell_update_fast<H><<<grid_for(items), WARPS_PER_CTA * 32, h->fsmem_upd, h->stream>>>(a);
        if ((rc = check_launch(h, "k_cell_update_fast"))) return rc;
    }
    if (!h->use_fast || h->n_big > 0) {
        Timed t(h, 1);
        a.skip_small = h->use_fast ? 1 : 0;
        k_cell_update<H><<<grid_for(items), WARPS_PER_CTA * 32, h->smem_upd, h->stream>>>(a);
        if ((rc = check_launch(h, "k_cell_update"))) return rc;
    }
    return 0;
}

template <bool H>
static int symmetry_stage(dugks_handle* h, StepArgs a) {
    // X = the reference's dfContainer (fvDVM.C:398-431); the MPI_Allgatherv of :439-449 is a sum all-reduce of
    // the rows every rank owns (zeros elsewhere), and only when mirror partners live on other ranks.
    // Nothing here synchronises with the host: mirror axes and Sf0 were fixed at create.
    const size_t nx = (size_t)h->sym_rows * h->nsym;
    int rc;
    if (h->sym_exchange) {
        CUDA_TRY(h, cudaMemsetAsync(h->sym_Xg, 0, nx * (H ? 2 : 1) * sizeof(double), h->stream));
    }
    {
        const long long total = (long long)h->nsym * (long long)h->nflat;
        const int grid = (int)std::min<long long>((total + 255) / 256, 148LL * 32);
        k_sym_pack<H><<<grid, 256, 0, h->stream>>>(a, h->d_symface, h->nsym, h->d_xrow, (int)h->nflat, h->sym_Xg, h->sym_Xh);
        if ((rc = check_launch(h, "k_sym_pack"))) return rc;
    }
    if (h->sym_exchange && (rc = do_allreduce(h, h->sym_Xg, nx * (H ? 2 : 1)))) return rc;   // g and h rows are one buffer
    for (const SymPatch& P : h->sym_patches) {
        const long long total = (long long)P.size * (long long)h->nflat;
        const int grid = (int)std::min<long long>((total + 255) / 256, 148LL * 32);
        k_sym_apply<H><<<grid, 256, 0, h->stream>>>(a, P, h->nsym, h->d_xrow, h->d_xmir, (int)h->nflat, h->sym_Xg, h->sym_Xh);
        if ((rc = check_launch(h, "k_sym_apply"))) return rc;
    }
    return 0;
}

template <bool H>
static int compute_wall_constants(dugks_handle* h) {
    if (!h->has_wall) return 0;
    StepArgs a = h->A;
    CUDA_TRY(h, cudaMemsetAsync(h->wall_cin, 0, (size_t)h->nbf * h->nm * sizeof(double), h->stream));
    CUDA_TRY(h, cudaMemsetAsync(h->wall_in, 0, (size_t)h->nbf * sizeof(double), h->stream));
    long long bitems = (long long)h->nbf * (h->Rs / 32);
    for (int s = 0; s < h->nslab; s++) {
        a.slab = s;
        if (h->use_hot && h->tabw <= 64) {
            const int tw = h->tabw <= 32 ? 32 : 64;
            const size_t sm = ((size_t)5 * h->ntab + (size_t)WARPS_PER_CTA * tw) * sizeof(double);
            k_hot_wall_constants<H><<<grid_for(bitems), WARPS_PER_CTA * 32, sm, h->stream>>>(a, h->wall_cin, h->wall_in, tw);
        } else
            k_wall_constants<H><<<grid_for(bitems), WARPS_PER_CTA * 32, 0, h->stream>>>(a, h->wall_cin, h->wall_in);
        int rc;
        if ((rc = check_launch(h, "k_wall_constants"))) return rc;
    }
    int rc;
    if ((rc = do_allreduce(h, h->wall_cin, (size_t)h->nbf * h->nm))) return rc;   // fvDVM.C:304-305
    if ((rc = do_allreduce(h, h->wall_in, (size_t)h->nbf))) return rc;
    return 0;
}

template <bool H>
static int step_impl(dugks_handle* h, double dt) {
    int rc;
    StepArgs a = h->A;
    a.dt = dt;
    // lagged boundary gradient: last step's new values become this step's old ones
    if (!h->gam_single) h->gam_flip = !h->gam_flip;
    a.gam_old_g = h->gam_flip ? h->gam_b_g : h->gam_a_g;
    a.gam_old_h = h->gam_flip ? h->gam_b_h : h->gam_a_h;
    a.gam_new_g = h->gam_flip ? h->gam_a_g : h->gam_b_g;
    a.gam_new_h = h->gam_flip ? h->gam_a_h : h->gam_b_h;
    size_t nslots = (size_t)2 * h->nif + h->nbf;
    CUDA_TRY(h, cudaMemsetAsync(a.fslot, 0, nslots * h->nm * sizeof(double), h->stream));
    CUDA_TRY(h, cudaMemsetAsync(a.cslot, 0, (size_t)h->nc * h->nm * sizeof(double), h->stream));
    if (h->pen_mode) {
        // the per-point constants of a row for the pencils' stencil loop (PenArgs::tx6), this step's dt folded in
        const int nt = h->ntab;
        for (int k = 0; k < nt + HOT_CI_MAX && k < 32 + HOT_CI_MAX; k++) {
            const int kk = std::min(k, nt - 1);
            const bool real = k < nt;
            h->pen.tx6[k * 6 + 0] = -0.5 * dt * h->tx_host[kk];
            for (int m = 1; m <= 4; m++) h->pen.tx6[k * 6 + m] = real ? h->tx_host[(size_t)m * nt + kk] : 0.0;
            h->pen.tx6[k * 6 + 5] = h->tx_host[kk];
        }
    }
    if (h->pen_mode == 2) {
        k_cell_coef<<<(h->nc + 127) / 128, 128, 0, h->stream>>>(a);
        if ((rc = check_launch(h, "k_cell_coef"))) return rc;
    }
    for (int s = 0; s < h->nslab; s++) {
        a.slab = s;
        if ((rc = launch_slab_kernels_phase1<H>(h, a))) return rc;
    }
    if (h->has_sym) {
        if ((rc = symmetry_stage<H>(h, a))) return rc;
    }
    if (h->nbf > 0) {
        long long bitems = (long long)h->nbf * (h->Rs / 32);
        for (int s = 0; s < h->nslab; s++) {
            a.slab = s;
            k_bnd_moments<H><<<grid_for(bitems), WARPS_PER_CTA * 32, 0, h->stream>>>(a);
            if ((rc = check_launch(h, "k_bnd_moments"))) return rc;
        }
    }
    const double* fsum = nullptr;
    if (h->nranks > 1) {
        // the two sides of an internal face are added before the collective: nf instead of 2 nif + nbf slots on the wire
        k_fold_fslot<<<148 * 8, 256, 0, h->stream>>>(a, h->fslot_fold);
        if ((rc = check_launch(h, "k_fold_fslot"))) return rc;
        if ((rc = do_allreduce(h, h->fslot_fold, (size_t)h->nf * h->nm))) return rc;   // fvDVM.C:363,487-489,519
        fsum = h->fslot_fold;
    }
    k_face_macros<<<(h->nf + 127) / 128, 128, 0, h->stream>>>(a, fsum);
    if ((rc = check_launch(h, "k_face_macros"))) return rc;
    for (int s = 0; s < h->nslab; s++) {
        a.slab = s;
        if ((rc = launch_slab_kernels_phase2<H>(h, a))) return rc;
    }
    if ((rc = do_allreduce(h, a.cslot, (size_t)h->nc * h->nm))) return rc;  // fvDVM.C:626-628,725
    k_cell_macros<<<(h->nc + 127) / 128, 128, 0, h->stream>>>(a);
    if ((rc = check_launch(h, "k_cell_macros"))) return rc;
    if (h->nbf > 0) {
        k_bnd_macros<<<(h->nbf + 127) / 128, 128, 0, h->stream>>>(a, h->d_bc, h->d_pres);
        if ((rc = check_launch(h, "k_bnd_macros"))) return rc;
    }
    h->steps++;
    return 0;
}

// ------------------------------------------------------------------------------
// velocity-space layout
struct RowDesc {
    int iy, iz, chunk;
    double y, z, w;
    int cbase = 0, len = 0;   // first ix of the row and its number of points
    bool pad = false;         // filler row (weight 0, owns no velocity): keeps a slab to two adjacent ix-chunks
};

// Slabs needed by nch ix-chunks of nbl rows each when a slab (32 rows, one warp) may only hold rows of at
// most two ADJACENT chunks (the equilibrium tables of a warp then span at most 2 L <= 64 entries): chunks are
// laid down one after the other and a slab is padded out as soon as a third chunk would enter it.
static long long packed_slabs(long long nbl, int nch) {
    if (nch == 1 || nbl >= 32) return (nbl * nch + 31) / 32;   // 32+ rows per chunk: never three chunks in a slab
    long long slabs = 0, fill = 0;
    int c0 = -1;
    for (int ch = 0; ch < nch; ch++)
        for (long long r = 0; r < nbl; r++) {
            if (fill == 32) { fill = 0; c0 = -1; }
            if (c0 >= 0 && ch > c0 + 1) { fill = 0; c0 = -1; }
            if (c0 < 0) { c0 = ch; slabs++; }
            fill++;
        }
    return slabs;
}

static int choose_chunks(int n, int D, long long base_rows_local) {
    // Number of ix-chunks per row: rows of L = ceil(n / nch) <= MAX_L points, 32 rows per slab.
    // Cost of a slab pass ~ (per-cell overhead + L): setting up a cell (upwind codes, equilibrium tables,
    // moment reduction) costs about as much as 8 velocity points, so short rows only pay when they
    // remove a lot of padding (measured at 8 GPUs: L = 4 gave 41.8 ms per step, L = 28 about 26 ms).
    int best = -1;
    double best_cost = 1e300;
    for (int nch = 1; nch <= n; nch++) {
        int L = (n + nch - 1) / nch;
        if (L > MAX_L) continue;
        if (L < 4 && nch > 1) break;
        if ((long long)nch * L > NT_MAX) continue;
        long long slabs = packed_slabs(base_rows_local, nch);
        double cost = (double)slabs * (8.0 + L);
        if (cost < best_cost * 0.98) { best_cost = cost; best = nch; }   // ties: the longer rows
    }
    return best;
}

// first base row (iy,iz) owned by rank r: contiguous, balanced blocks
static int row_begin_of(int nbase, int nranks, int r) {
    long long q = nbase / nranks, m = nbase % nranks;
    return (int)(r * q + std::min<long long>(r, m));
}

// Velocity-row layout of one rank (host only): rows of L = ceil(n / nch) points over the rank's block of
// (iy, iz) base rows, sorted so that warps are sign-coherent, plus the optional short-row tail slab.
struct RowLayout {
    int nch = 1, L = 0, Lt = 0, ntab = 0;
    std::vector<RowDesc> rows;
};

static int build_row_layout(int n, int D, int nranks, int rank, const double* Xis, const double* weights, RowLayout& out) {
    const int ny = (D >= 2) ? n : 1, nz = (D == 3) ? n : 1;
    const int nbase = ny * nz, Rs = 32;
    if (nbase < nranks) return -1;
    const int rb0 = row_begin_of(nbase, nranks, rank), rb1 = row_begin_of(nbase, nranks, rank + 1);
    int max_nbl = 0;
    for (int r = 0; r < nranks; r++) max_nbl = std::max(max_nbl, row_begin_of(nbase, nranks, r + 1) - row_begin_of(nbase, nranks, r));
    int nch = choose_chunks(n, D, max_nbl);
    if (const char* e = getenv("DUGKS_NCH")) {   // experiment hook: force the number of ix-chunks per row
        const int want = atoi(e);
        if (want >= 1 && (n + want - 1) / want <= MAX_L && (long long)want * ((n + want - 1) / want) <= NT_MAX) nch = want;
    }
    if (nch < 1) return -1;
    const int L = (n + nch - 1) / nch;
    out.nch = nch; out.L = L; out.ntab = nch * L; out.Lt = 0;
    std::vector<RowDesc>& rows = out.rows;
    rows.clear();
    for (int ch = 0; ch < nch; ch++)
        for (int br = rb0; br < rb1; br++) {
            int iy = (D >= 2) ? br % n : 0, iz = (D == 3) ? br / n : 0;
            RowDesc rd;
            rd.iy = iy; rd.iz = iz; rd.chunk = ch;
            rd.y = (D >= 2) ? Xis[iy] : 0.0;
            rd.z = (D == 3) ? Xis[iz] : 0.0;
            // weight = w[iz]*w[iy]*w[ix] (fvDVM.C:158,187,211): the row carries w[iz]*w[iy]
            rd.w = (D == 3) ? weights[iz] * weights[iy] : ((D == 2) ? weights[iy] : 1.0);
            rows.push_back(rd);
        }
    // keep warps sign-coherent in (xi_y, xi_z): less divergence in the upwind test
    std::stable_sort(rows.begin(), rows.end(), [](const RowDesc& a, const RowDesc& b) {
        if (a.chunk != b.chunk) return a.chunk < b.chunk;
        int sa = (a.z < 0 ? 0 : 2) + (a.y < 0 ? 0 : 1), sb = (b.z < 0 ? 0 : 2) + (b.y < 0 ? 0 : 1);
        return sa < sb;
    });
    for (auto& rd : rows) { rd.cbase = rd.chunk * L; rd.len = L; }
    if (nch > 1) {
        // a slab holds rows of at most two adjacent ix-chunks (packed_slabs): pad it out before a third one enters
        std::vector<RowDesc> packed;
        int fill = 0, c0 = -1;
        for (const RowDesc& rd : rows) {
            if (fill == Rs) { fill = 0; c0 = -1; }
            if (c0 >= 0 && rd.chunk > c0 + 1) {
                for (; fill < Rs; fill++) { RowDesc p = packed.back(); p.w = 0.0; p.pad = true; packed.push_back(p); }
                fill = 0; c0 = -1;
            }
            if (c0 < 0) c0 = rd.chunk;
            packed.push_back(rd);
            fill++;
        }
        rows.swap(packed);
    }
    // short-row tail slab (dv_len, dugks_device.cuh): the rows that would leave the last slab mostly
    // empty are cut into ix-chunks of Lt points and fill the lanes of one short slab
    const int r_last = (int)rows.size() % Rs;
    const bool allow = nch == 1 && getenv("DUGKS_NO_TAIL") == nullptr && getenv("DUGKS_NO_HOT") == nullptr;
    if (allow && r_last > 0 && Rs / r_last >= 2 && n >= 2) {
        int ncht = std::min(Rs / r_last, n);
        const int Lt = (n + ncht - 1) / ncht;
        ncht = (n + Lt - 1) / Lt;
        if (r_last * ncht <= Rs && Lt < L && ncht * Lt <= NT_MAX) {
            std::vector<RowDesc> tail(rows.end() - r_last, rows.end());
            rows.resize(rows.size() - r_last);
            for (int ct = 0; ct < ncht; ct++)
                for (RowDesc rd : tail) { rd.chunk = ct; rd.cbase = ct * Lt; rd.len = Lt; rows.push_back(rd); }
            out.Lt = Lt;
            out.ntab = std::max(out.ntab, ncht * Lt);
        }
    }
    return 0;
}

extern "C" int dugks_row_layout(int32_t nXiPerDim, int32_t nSolutionD, int32_t nRanks, int32_t rank, int32_t* nch,
                                int32_t* L, int32_t* Lt, int32_t* nRows, int32_t* row_iy, int32_t* row_iz,
                                int32_t* row_first, int32_t* row_len) {
    if (nXiPerDim < 1 || nSolutionD < 1 || nSolutionD > 3 || nRanks < 1 || rank < 0 || rank >= nRanks || !nRows)
        return fail(nullptr, DUGKS_ERR_INVALID, "dugks_row_layout: bad argument");
    std::vector<double> x(nXiPerDim), w(nXiPerDim, 1.0);
    for (int i = 0; i < nXiPerDim; i++) x[i] = i - 0.5 * (nXiPerDim - 1);   // symmetric placeholder abscissae
    RowLayout lay;
    if (build_row_layout(nXiPerDim, nSolutionD, nRanks, rank, x.data(), w.data(), lay) != 0)
        return fail(nullptr, DUGKS_ERR_UNSUPPORTED, "nDV = %d cannot be laid out over %d ranks", nXiPerDim, nRanks);
    if (nch) *nch = lay.nch;
    if (L) *L = lay.L;
    if (Lt) *Lt = lay.Lt;
    const int cap = *nRows;
    *nRows = (int)lay.rows.size();
    if (row_iy && row_iz && row_first && row_len)
        for (int k = 0; k < (int)lay.rows.size() && k < cap; k++) {
            row_iy[k] = lay.rows[k].iy; row_iz[k] = lay.rows[k].iz;
            row_first[k] = lay.rows[k].cbase; row_len[k] = lay.rows[k].pad ? 0 : lay.rows[k].len;   // filler rows own nothing
        }
    return 0;
}

// ------------------------------------------------------------------------------
// Traversal order of the cell kernels: a permutation of the cells, computed from the cell centres, so it
// applies to unstructured meshes as well.  Persistent warps take items w, w + nWarps, w + 2 nWarps, ...
//   tiled (default): strips of T rows in y, swept layer by layer in z, x fastest inside a row: the +-z
//     neighbours of a cell are nx*T cells apart instead of nx*ny (64^3: 512 instead of 4096), so that every
//     neighbour block is touched again within about one wave of warps; the rows on strip borders (2 in T)
//     are fetched twice.
//   wave: lines of cells along x are dealt to the warps in flight, nWarps lines at a time, and the order
//     runs through all of them x position by x position: at any time the warps work on one y-z patch of
//     cells at the SAME x, so that the y/z neighbours of a cell are in flight together with it and its x
//     neighbours are the previous / next cell of the same warp.  A probe of where the L2 misses of the row
//     blocks come from; opt-in until measured.
//   morton, natural: for comparison.
// first != nullptr: cells with first[c] != 0 come first (the axis-only launch takes them as one item range);
// the pattern is built inside that class, the other cells follow in tiled order.
static void build_cell_order(int nc, int D, const double* C, const unsigned char* first, const std::string& ord,
                             int tile, int nwarps, std::vector<int>& order) {
    order.resize(nc);
    for (int c = 0; c < nc; c++) order[c] = c;
    if (nc > 1 && ord != "natural") {
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        for (int c = 0; c < nc; c++)
            for (int d = 0; d < 3; d++) {
                lo[d] = std::min(lo[d], C[(size_t)c * 3 + d]);
                hi[d] = std::max(hi[d], C[(size_t)c * 3 + d]);
            }
        std::vector<unsigned long long> key(nc);
        // cells per direction if the mesh were a uniform block
        const int n1 = std::max(1, (int)std::lround(std::pow((double)nc, 1.0 / std::max(D, 1))));
        auto quant = [&](int c, int d) -> unsigned long long {
            const double w = hi[d] - lo[d];
            if (!(w > 0)) return 0ull;
            // centres of a uniform block sit at (i + 0.5) / n1 of the centre-to-centre extent + half a cell
            const double t = (C[(size_t)c * 3 + d] - lo[d]) / w * (n1 - 1) + 0.5;
            return (unsigned long long)std::min<double>(n1 - 1, std::max(0.0, std::floor(t)));
        };
        if (ord == "morton") {
            auto spread = [](unsigned v) {   // 10 bits -> every third bit
                unsigned long long x = v & 0x3ff;
                x = (x | (x << 16)) & 0x30000ffull;
                x = (x | (x << 8)) & 0x300f00full;
                x = (x | (x << 4)) & 0x30c30c3ull;
                x = (x | (x << 2)) & 0x9249249ull;
                return x;
            };
            for (int c = 0; c < nc; c++) {
                unsigned long long k = 0;
                for (int d = 0; d < 3; d++) {
                    const double w = hi[d] - lo[d];
                    unsigned q = w > 0 ? (unsigned)std::min(1023.0, (C[(size_t)c * 3 + d] - lo[d]) / w * 1024.0) : 0u;
                    k |= spread(q) << d;
                }
                key[c] = k;
            }
        } else {
            // strip height for ~512 cells per layer
            const int T = tile > 0 ? tile : std::max(1, 512 / n1);
            for (int c = 0; c < nc; c++) {
                const unsigned long long qx = quant(c, 0), qy = quant(c, 1), qz = quant(c, 2);
                key[c] = (((qy / T) * n1 + qz) * T + (qy % T)) * n1 + qx;
            }
        }
        std::stable_sort(order.begin(), order.end(), [&](int x, int y2) { return key[x] < key[y2]; });
        if (ord == "wave" && nwarps > 0) {
            // the cells the pattern is built over, line by line (z, y), x ascending inside a line
            std::vector<int> sub;
            for (int c = 0; c < nc; c++) if (!first || first[c]) sub.push_back(c);
            std::vector<unsigned long long> line(nc, 0), qxs(nc, 0);
            for (int c : sub) { line[c] = quant(c, 2) * (unsigned long long)n1 + quant(c, 1); qxs[c] = quant(c, 0); }
            std::stable_sort(sub.begin(), sub.end(), [&](int x, int y2) {
                return line[x] != line[y2] ? line[x] < line[y2] : qxs[x] < qxs[y2];
            });
            // key = (round of nWarps lines, position inside the line, line inside the round)
            long long li = -1, pos = 0;
            unsigned long long prev = ~0ull;
            for (int c : sub) {
                if (line[c] != prev) { li++; pos = 0; prev = line[c]; }
                key[c] = ((unsigned long long)(li / nwarps) << 44) | ((unsigned long long)pos << 22) | (unsigned long long)(li % nwarps);
                pos++;
            }
            std::stable_sort(sub.begin(), sub.end(), [&](int x, int y2) { return key[x] < key[y2]; });
            if (!first) order = sub;
            else {
                // the other cells keep their tiled order, behind the pattern
                std::vector<int> rest;
                for (int c : order) if (!first[c]) rest.push_back(c);
                order = sub;
                order.insert(order.end(), rest.begin(), rest.end());
            }
            return;
        }
    }
    if (first) std::stable_partition(order.begin(), order.end(), [&](int c) { return first[c] != 0; });
}

// introspection (host only, no device needed): the traversal order dugks_create would use
extern "C" int dugks_cell_order(int32_t nCells, int32_t nSolutionD, const double* C, const uint8_t* first_class,
                                const char* kind, int32_t nWarps, int32_t* order) {
    if (nCells <= 0 || !C || !order || !kind || nWarps <= 0 || nSolutionD < 1 || nSolutionD > 3)
        return fail(nullptr, DUGKS_ERR_INVALID, "dugks_cell_order: bad argument");
    std::vector<int> o;
    build_cell_order(nCells, nSolutionD, C, first_class, kind, 0, nWarps, o);
    for (int i = 0; i < nCells; i++) order[i] = o[i];
    return 0;
}

// ------------------------------------------------------------------------------
// Host-side mesh analysis shared by dugks_create and the host-only introspection entry points.
// Cell -> face CSR (internal faces first) with one geometry record per (cell, face) entry: LS vector G(3),
// r = Cf - C_cell (3), Sf (3).  Axis-aligned cells: every LS vector and face offset of the cell has exactly ONE
// non-zero component (exact zeros, as orthogonal hexahedra give), two faces per axis, all internal; their
// entries are put in the canonical order x-, x+, y-, y+[, z-, z+] so that the kernels can skip the products
// with the exact zeros at compile time (bit-identical to the general path) and know a neighbour's direction
// from its entry index.
struct HostCsr {
    std::vector<int> cnt, cnt_int, off, e_other, e_face, e_owner;
    std::vector<double> e_geo;
    std::vector<unsigned char> cell_cls;
    int n_axis = 0, axis_ne = 0;     // axis_ne: 4 or 6 when all axis-aligned cells have that many faces, -1 if mixed
};

static int build_host_csr(const dugks_mesh_t* mesh, bool want_axis, HostCsr& out, std::string& err) {
    const int nc = mesh->nCells, nif = mesh->nInternalFaces, nbf = mesh->nBoundaryFaces;
    char buf[256];
    std::vector<int>&cnt = out.cnt, &cnt_int = out.cnt_int, &off = out.off, &e_other = out.e_other, &e_face = out.e_face, &e_owner = out.e_owner;
    std::vector<double>& e_geo = out.e_geo;
    cnt.assign(nc, 0); cnt_int.assign(nc, 0);
    for (int f = 0; f < nif; f++) {
        int o = mesh->owner[f], nb = mesh->neighbour[f];
        if (o < 0 || o >= nc || nb < 0 || nb >= nc) { snprintf(buf, sizeof buf, "face %d: owner/neighbour out of range", f); err = buf; return DUGKS_ERR_INVALID; }
        cnt[o]++; cnt[nb]++; cnt_int[o]++; cnt_int[nb]++;
    }
    for (int b = 0; b < nbf; b++) {
        int o = mesh->owner[nif + b];
        if (o < 0 || o >= nc) { snprintf(buf, sizeof buf, "boundary face %d: owner out of range", b); err = buf; return DUGKS_ERR_INVALID; }
        cnt[o]++;
    }
    off.assign(nc + 1, 0);
    for (int c = 0; c < nc; c++) {
        if (cnt[c] > MAX_CELL_FACES) { snprintf(buf, sizeof buf, "cell %d has %d faces (limit %d)", c, cnt[c], MAX_CELL_FACES); err = buf; return DUGKS_ERR_UNSUPPORTED; }
        off[c + 1] = off[c] + cnt[c];
    }
    const int ne = off[nc];
    e_other.assign(ne, 0); e_face.assign(ne, 0); e_owner.assign(ne, 0);
    e_geo.assign((size_t)ne * 9, 0.0);
    std::vector<int> fill_i(nc, 0), fill_b(nc, 0);
    auto put = [&](int c, int pos, int other, int face, int isown, const double* G) {
        int e = off[c] + pos;
        e_other[e] = other; e_face[e] = face; e_owner[e] = isown;
        for (int d = 0; d < 3; d++) {
            e_geo[(size_t)e * 9 + d] = G[d];
            e_geo[(size_t)e * 9 + 3 + d] = mesh->Cf[(size_t)face * 3 + d] - mesh->C[(size_t)c * 3 + d];
            e_geo[(size_t)e * 9 + 6 + d] = mesh->Sf[(size_t)face * 3 + d];
        }
    };
    for (int f = 0; f < nif; f++) {
        int o = mesh->owner[f], nb = mesh->neighbour[f];
        put(o, fill_i[o]++, nb, f, 1, mesh->ownLs + (size_t)f * 3);
        // grad[nei] -= neiLs*(v_nei - v_own) == neiLs*(v_own - v_nei): the sign is inside neiLs
        // (zeroBoundaryGrad.C:98, zeroBoundaryVectors.C:183-184)
        put(nb, fill_i[nb]++, o, f, 0, mesh->neiLs + (size_t)f * 3);
    }
    for (int b = 0; b < nbf; b++) {
        int f = nif + b, o = mesh->owner[f];
        put(o, cnt_int[o] + fill_b[o]++, -1 - b, f, 1, mesh->patchLs + (size_t)b * 3);
    }
    // Rounding residue of the geometry formulas: on an orthogonal mesh whose coordinates are not exact binary
    // fractions (60 x 60 cells on [0,1]^2, a plate of thickness 0.1) the components of Cf - C and of an LS vector
    // that vanish in exact arithmetic come out as a few ulp of the COORDINATES they were differenced from, i.e.
    // amplified by |C| / |Cf - C| relative to the vector itself.  Components below 16 ulp of that scale are set
    // to zero HERE, for every kernel path alike.  What is dropped is rounding noise of the same size as the noise
    // left in the other components (the reference computes with it; neither is "the" value), at worst a few
    // 1e-13 of a face value, below the per-step tolerance; without it hardly any cell of such a mesh would be
    // recognised as axis-aligned.  Components along a direction in which every LS vector of the cell vanishes
    // (the empty direction of a 2-D case) multiply a gradient component that is exactly zero: zeroed as well.
    if (want_axis) {
        const double ulp16 = 16.0 * 2.220446049250313e-16;
        for (int c = 0; c < nc; c++) {
            bool gzero[3] = {true, true, true};
            for (int e = off[c]; e < off[c + 1]; e++) {
                double* g = &e_geo[(size_t)e * 9];
                const int f = e_face[e];
                double cmax = 0.0, rmax = 0.0, gmax = 0.0;
                for (int d = 0; d < 3; d++) {
                    cmax = std::max(cmax, std::max(std::fabs(mesh->Cf[(size_t)f * 3 + d]), std::fabs(mesh->C[(size_t)c * 3 + d])));
                    rmax = std::max(rmax, std::fabs(g[3 + d]));
                    gmax = std::max(gmax, std::fabs(g[d]));
                }
                const double amp = rmax > 0 ? std::max(1.0, cmax / rmax) : 1.0;
                for (int d = 0; d < 3; d++) {
                    if (std::fabs(g[3 + d]) <= ulp16 * std::max(cmax, rmax)) g[3 + d] = 0.0;
                    if (std::fabs(g[d]) <= ulp16 * amp * gmax) g[d] = 0.0;
                }
                for (int d = 0; d < 3; d++) gzero[d] = gzero[d] && g[d] == 0.0;
            }
            for (int e = off[c]; e < off[c + 1]; e++)
                for (int d = 0; d < 3; d++) if (gzero[d]) e_geo[(size_t)e * 9 + 3 + d] = 0.0;
        }
    }
    out.cell_cls.assign(nc, 0);
    out.n_axis = 0; out.axis_ne = 0;
    std::vector<int> axis_of(MAX_CELL_FACES), order(MAX_CELL_FACES);
    std::vector<double> tmp_geo(MAX_CELL_FACES * 9);
    std::vector<int> tmp_i(MAX_CELL_FACES * 3);
    for (int c = 0; c < nc && want_axis; c++) {
        const int n_e = cnt[c];
        if (n_e != cnt_int[c] || (n_e != 4 && n_e != 6)) continue;
        int per_axis[3] = {0, 0, 0};
        bool ok = true;
        for (int j = 0; j < n_e && ok; j++) {
            const double* g = &e_geo[(size_t)(off[c] + j) * 9];
            int ax = -1;
            for (int d = 0; d < 3; d++)
                if (g[d] != 0.0 || g[3 + d] != 0.0) { if (ax >= 0 && ax != d) ok = false; ax = d; }
            if (ax < 0) ok = false;
            // the flux of the update kernels keeps the general form, so Sf is not constrained
            if (ok) { axis_of[j] = ax; per_axis[ax]++; }
        }
        if (!ok) continue;
        const int naxes = n_e / 2;
        for (int d = 0; d < 3; d++) if (per_axis[d] != (d < naxes ? 2 : 0)) ok = false;
        if (!ok) continue;
        int k = 0;
        for (int d = 0; d < 3; d++) {
            const int k0 = k;
            for (int j = 0; j < n_e; j++) if (axis_of[j] == d) order[k++] = j;
            if (k - k0 == 2 && e_geo[(size_t)(off[c] + order[k0]) * 9 + 3 + d] > e_geo[(size_t)(off[c] + order[k0 + 1]) * 9 + 3 + d])
                std::swap(order[k0], order[k0 + 1]);
        }
        for (int j = 0; j < n_e; j++) {
            const int e = off[c] + order[j];
            for (int d = 0; d < 9; d++) tmp_geo[j * 9 + d] = e_geo[(size_t)e * 9 + d];
            tmp_i[j * 3] = e_other[e]; tmp_i[j * 3 + 1] = e_face[e]; tmp_i[j * 3 + 2] = e_owner[e];
        }
        for (int j = 0; j < n_e; j++) {
            const int e = off[c] + j;
            for (int d = 0; d < 9; d++) e_geo[(size_t)e * 9 + d] = tmp_geo[j * 9 + d];
            e_other[e] = tmp_i[j * 3]; e_face[e] = tmp_i[j * 3 + 1]; e_owner[e] = tmp_i[j * 3 + 2];
        }
        out.cell_cls[c] = 1;
        out.axis_ne = (out.n_axis == 0 || out.axis_ne == n_e) ? n_e : -1;   // -1: mixed face counts
        out.n_axis++;
    }
    return 0;
}

// ------------------------------------------------------------------------------
// CTA pencils of phase 1 (dugks_pencil.cuh), host side: x-lines of axis-aligned interior cells (entries in the
// canonical order x-, x+, y-, y+, z-, z+), 2 x 2 bundles of lines whose cells are mutual y / z neighbours position
// by position, and the work items (runs of consecutive x positions of a bundle) in the order the CTAs take them.
struct PenHost {
    std::vector<PenItem> items;
    std::vector<int> cells, halo;
    std::vector<int> order;          // pencil cells in traversal order: item, step, line
};

static void build_pencils(int nc, const std::vector<int>& off, const std::vector<int>& e_other,
                          const std::vector<unsigned char>& cls, int grid, PenHost& out) {
    auto nbr = [&](int c, int j) { return e_other[off[c] + j]; };
    auto axis = [&](int c) { return c >= 0 && cls[c] != 0; };
    // ---- lines: maximal chains along x+ of axis-aligned cells
    std::vector<int> line_of(nc, -1), pos_of(nc, -1);
    std::vector<std::vector<int>> lines;
    for (int c = 0; c < nc; c++) {
        if (!axis(c) || axis(nbr(c, 0))) continue;            // not a line start
        std::vector<int> ln;
        for (int cur = c; axis(cur) && line_of[cur] < 0; cur = nbr(cur, 1)) {
            line_of[cur] = (int)lines.size();
            pos_of[cur] = (int)ln.size();
            ln.push_back(cur);
        }
        lines.push_back(std::move(ln));
    }
    // ---- bundles: A, B = y+ of A, C = z+ of A, D = y+ of C, equal lengths, aligned position by position
    std::vector<char> used(lines.size(), 0);
    std::vector<std::array<int, 4>> bundles;
    for (size_t la = 0; la < lines.size(); la++) {
        if (used[la]) continue;
        const std::vector<int>& A = lines[la];
        const int b0 = nbr(A[0], 3), c0 = nbr(A[0], 5);
        if (!axis(b0) || !axis(c0) || pos_of[b0] != 0 || pos_of[c0] != 0) continue;
        const int lb = line_of[b0], lc = line_of[c0];
        const int d0 = nbr(c0, 3);
        if (!axis(d0) || pos_of[d0] != 0) continue;
        const int ld = line_of[d0];
        if (lb == (int)la || lc == (int)la || ld == (int)la || lb == lc || lb == ld || lc == ld) continue;
        if (used[lb] || used[lc] || used[ld]) continue;
        const std::vector<int>&B = lines[lb], &C = lines[lc], &Dl = lines[ld];
        if (B.size() != A.size() || C.size() != A.size() || Dl.size() != A.size()) continue;
        bool ok = true;
        for (size_t k = 0; k < A.size() && ok; k++)
            ok = nbr(A[k], 3) == B[k] && nbr(A[k], 5) == C[k] && nbr(C[k], 3) == Dl[k] && nbr(B[k], 5) == Dl[k] &&
                 nbr(B[k], 2) == A[k] && nbr(C[k], 4) == A[k] && nbr(Dl[k], 2) == C[k] && nbr(Dl[k], 4) == B[k];
        if (!ok) continue;
        used[la] = used[lb] = used[lc] = used[ld] = 1;
        bundles.push_back({(int)la, lb, lc, ld});
    }
    if (bundles.empty()) return;
    // ---- work items.  CTA j of `grid` takes items j, j + grid, ...: whole lines for as many full rounds as there
    // are, the remaining bundles cut into pieces so that the last rounds fill the grid too.
    struct Piece { int b, s0, s1; };
    std::vector<Piece> pieces;
    const int nb = (int)bundles.size();
    const int full = (nb / grid) * grid;
    for (int b = 0; b < full; b++) pieces.push_back({b, 0, (int)lines[bundles[b][0]].size()});
    if (nb > full) {
        const int rem = nb - full;
        int nmax = 0;
        for (int b = full; b < nb; b++) nmax = std::max(nmax, (int)lines[bundles[b][0]].size());
        int best_k = 1;
        long long best = -1;
        for (int k = 1; k <= 8; k++) {
            const int len = (nmax + k - 1) / k;
            if (k > 1 && len < 6) break;
            const long long cost = (long long)(((long long)rem * k + grid - 1) / grid) * (len + 1);   // + 1: the prologue of a piece
            if (best < 0 || cost < best) { best = cost; best_k = k; }
        }
        for (int pc = 0; pc < best_k; pc++)
            for (int b = full; b < nb; b++) {
                const int n = (int)lines[bundles[b][0]].size();
                const int s0 = (int)((long long)n * pc / best_k), s1 = (int)((long long)n * (pc + 1) / best_k);
                if (s1 > s0) pieces.push_back({b, s0, s1});
            }
    }
    for (const Piece& pc : pieces) {
        PenItem it;
        it.cells = (int)out.cells.size();
        it.halo = (int)out.halo.size();
        it.item0 = (int)out.order.size();
        it.nsteps = pc.s1 - pc.s0;
        for (int p = pc.s0 - 1; p <= pc.s1; p++)
            for (int l = 0; l < 4; l++) {
                const std::vector<int>& ln = lines[bundles[pc.b][l]];
                out.cells.push_back(p < 0 ? nbr(ln[0], 0) : (p >= (int)ln.size() ? nbr(ln.back(), 1) : ln[p]));
            }
        for (int p = pc.s0; p < pc.s1; p++)
            for (int l = 0; l < 4; l++) {
                const int c = lines[bundles[pc.b][l]][p];
                out.halo.push_back(nbr(c, (l & 1) ? 3 : 2));   // by = 0: y- is outside the bundle, by = 1: y+
                out.halo.push_back(nbr(c, (l & 2) ? 5 : 4));
                out.order.push_back(c);
            }
        out.items.push_back(it);
    }
}

// introspection (host only, no device needed): the pencil work items dugks_create would build for this mesh
extern "C" int dugks_pencil_plan(const dugks_mesh_t* mesh, int32_t nCtas, int32_t* nItems, int32_t* nPencilCells,
                                 int32_t* cells, int32_t* item_first, int32_t* item_steps, int32_t* nAxisCells) {
    if (!mesh || nCtas < 1 || !nItems || !nPencilCells) return fail(nullptr, DUGKS_ERR_INVALID, "dugks_pencil_plan: bad argument");
    HostCsr csr;
    std::string err;
    int rc = build_host_csr(mesh, true, csr, err);
    if (rc) return fail(nullptr, rc, "%s", err.c_str());
    PenHost ph;
    if (mesh->nSolutionD == 3 && csr.axis_ne == PEN_NE) build_pencils(mesh->nCells, csr.off, csr.e_other, csr.cell_cls, nCtas, ph);
    const int cap_items = *nItems, cap_cells = *nPencilCells;
    *nItems = (int)ph.items.size();
    *nPencilCells = (int)ph.order.size();
    if (nAxisCells) *nAxisCells = csr.n_axis;
    if (cells) for (int k = 0; k < (int)ph.order.size() && k < cap_cells; k++) cells[k] = ph.order[k];
    for (int k = 0; k < (int)ph.items.size() && k < cap_items; k++) {
        if (item_first) item_first[k] = ph.items[k].item0;
        if (item_steps) item_steps[k] = ph.items[k].nsteps;
    }
    return 0;
}

extern "C" int dugks_abi_version(void) { return DUGKS_ABI_VERSION; }

extern "C" int dugks_partition(int32_t nXiPerDim, int32_t nSolutionD, int32_t nRanks, int32_t rank, int32_t* ids,
                               int32_t* n) {
    if (nXiPerDim < 1 || nSolutionD < 1 || nSolutionD > 3 || nRanks < 1 || rank < 0 || rank >= nRanks || !n)
        return fail(nullptr, DUGKS_ERR_INVALID, "dugks_partition: bad argument");
    const int nn = nXiPerDim, ny = (nSolutionD >= 2) ? nn : 1, nz = (nSolutionD == 3) ? nn : 1;
    const int nbase = ny * nz;
    if (nbase < nRanks) return fail(nullptr, DUGKS_ERR_UNSUPPORTED, "%d velocity rows cannot be split over %d ranks", nbase, nRanks);
    int b0 = row_begin_of(nbase, nRanks, rank), b1 = row_begin_of(nbase, nRanks, rank + 1);
    *n = (b1 - b0) * nn;
    if (ids)
        for (int k = 0; k < *n; k++) ids[k] = b0 * nn + k;
    return 0;
}

extern "C" const char* dugks_last_error(const dugks_handle_t* h) {
    return h ? h->err.c_str() : g_create_error.c_str();
}

extern "C" int dugks_nccl_unique_id(void* out128) {
    if (!out128) return fail(nullptr, DUGKS_ERR_INVALID, "dugks_nccl_unique_id: NULL output");
    std::string err;
    if (!g_nccl.load(err)) return fail(nullptr, DUGKS_ERR_COMM, "%s", err.c_str());
    nccl_uid_t id;
    int rc = g_nccl.get_uid(&id);
    if (rc) return fail(nullptr, DUGKS_ERR_COMM, "ncclGetUniqueId failed (%d)", rc);
    memcpy(out128, &id, sizeof id);
    return 0;
}

extern "C" void dugks_destroy(dugks_handle_t* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->nccl_comm && g_nccl.destroy) g_nccl.destroy(h->nccl_comm);
    for (auto& b : h->bufs) cudaFree(b.p);
    if (h->pin) cudaFreeHost(h->pin);
    if (h->d_soa) cudaFree(h->d_soa);
    for (auto& e : h->events) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    for (auto& e : h->pool) { cudaEventDestroy(e.a); cudaEventDestroy(e.b); }
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

template <bool H>
static int init_state(dugks_handle* h) {
    StepArgs a = h->A;
    a.dt = 0.0;
    int rc;
    k_init_tau<<<(h->nc + 127) / 128, 128, 0, h->stream>>>(a);                  // fvDVM.C:1073
    if ((rc = check_launch(h, "k_init_tau"))) return rc;
    if (h->nbf > 0) {
        k_bnd_macros<<<(h->nbf + 127) / 128, 128, 0, h->stream>>>(a, h->d_bc, h->d_pres);   // fvDVM.C:1072
        if ((rc = check_launch(h, "k_bnd_macros"))) return rc;
    }
    long long items = (long long)h->nc * (h->Rs / 32);
    for (int s = 0; s < h->nslab; s++) {
        a.slab = s;
        k_cell_halfstep<H><<<grid_for(items), WARPS_PER_CTA * 32, 0, h->stream>>>(a, 1);   // initDFtoEq
        if ((rc = check_launch(h, "k_cell_halfstep(init)"))) return rc;
        if (h->nbf > 0) {
            long long total = (long long)h->nbf * h->L * h->Rs;
            int grid = (int)std::min<long long>((total + 255) / 256, 148LL * 32);
            k_bnd_init_mixed<H><<<grid, 256, 0, h->stream>>>(a);                            // initBoundaryField
            if ((rc = check_launch(h, "k_bnd_init_mixed"))) return rc;
        }
    }
    return compute_wall_constants<H>(h);                                                   // fvDVM.C:1069
}

extern "C" int dugks_create(const dugks_mesh_t* mesh, const dugks_patch_t* patches, int32_t nPatches,
                            const dugks_dvset_t* dvset, const dugks_gas_t* gas, const dugks_par_t* par,
                            const double* rho, const double* U, const double* T, const double* rho_b,
                            const double* U_b, const double* T_b, dugks_handle_t** out) {
    if (!out) return fail(nullptr, DUGKS_ERR_INVALID, "dugks_create: out is NULL");
    *out = nullptr;
    if (!mesh || !dvset || !gas || !rho || !U || !T)
        return fail(nullptr, DUGKS_ERR_INVALID, "dugks_create: NULL argument");
    if (mesh->nCells <= 0 || mesh->nInternalFaces < 0 || mesh->nBoundaryFaces < 0)
        return fail(nullptr, DUGKS_ERR_INVALID, "dugks_create: bad mesh sizes");
    if (mesh->nSolutionD < 1 || mesh->nSolutionD > 3)
        return fail(nullptr, DUGKS_ERR_INVALID, "dugks_create: nSolutionD must be 1, 2 or 3");
    if (dvset->nXiPerDim < 1 || !dvset->Xis || !dvset->weights)
        return fail(nullptr, DUGKS_ERR_INVALID, "dugks_create: bad discrete-velocity set");
    if (mesh->nBoundaryFaces > 0 && (!rho_b || !U_b || !T_b || !patches))
        return fail(nullptr, DUGKS_ERR_INVALID, "dugks_create: boundary fields missing");
    dugks_par_t pdef{};
    pdef.nRanks = 1; pdef.device = -1;
    if (!par) par = &pdef;
    int nranks = par->nRanks < 1 ? 1 : par->nRanks;
    if (par->rank < 0 || par->rank >= nranks) return fail(nullptr, DUGKS_ERR_INVALID, "dugks_create: bad rank");
    if (par->partition != 0 || par->dv_chunk != 0)
        return fail(nullptr, DUGKS_ERR_INVALID, "dugks_create: dugks_par_t.partition and .dv_chunk are reserved and must be 0");
    if (!(par->limiter_k >= 0.0))
        return fail(nullptr, DUGKS_ERR_INVALID, "dugks_create: dugks_par_t.limiter_k must be >= 0 (VenkatakrishnanSlopeMulti.C:70-76)");

    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return fail(nullptr, DUGKS_ERR_NO_DEVICE,
                    "no CUDA device available (%s); dugksfoam_b200 has no CPU fallback",
                    ce != cudaSuccess ? cudaGetErrorString(ce) : "device count is 0");
    dugks_handle* h = new dugks_handle();
    auto bail = [&](int rc) { g_create_error = h->err; dugks_destroy(h); return rc; };
    int dev = par->device;
    if (dev < 0) cudaGetDevice(&dev);
    if (dev >= ndev) { fail(h, DUGKS_ERR_INVALID, "device %d out of range (%d devices)", dev, ndev); return bail(DUGKS_ERR_INVALID); }
    h->device = dev;
    int rc;
#define TRYB(x) do { rc = (x); if (rc) return bail(rc); } while (0)
#define CUDAB(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { fail(h, DUGKS_ERR_CUDA, "%s: %s", #x, cudaGetErrorString(e__)); return bail(DUGKS_ERR_CUDA); } } while (0)
    CUDAB(cudaSetDevice(dev));
    CUDAB(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));

    const int nc = mesh->nCells, nif = mesh->nInternalFaces, nbf = mesh->nBoundaryFaces, nf = nif + nbf;
    const int D = mesh->nSolutionD, n = dvset->nXiPerDim;
    h->nc = nc; h->nif = nif; h->nbf = nbf; h->nf = nf; h->D = D; h->n1d = n;
    h->rank = par->rank; h->nranks = nranks; h->xiMax = dvset->xiMax;
    h->gas = DevGas{gas->R, gas->omega, gas->Tref, gas->muRef, gas->Pr, gas->KInner, D};
    h->patches.assign(patches, patches + nPatches);
    h->reduce = par->reduce; h->reduce_user = par->reduce_user;
    // h == 0 identically when K+3-D == 0 (discreteVelocity.C:1043,1059): elide it unless asked not to
    h->hasH = !((gas->KInner + 3 - D) == 0 && par->store_h == 0);
    h->nm = h->hasH ? NM_MAX : NM_G;

    // ---- patches: validate, per-face kind / bc tables
    std::vector<int> b_kind(nbf, -1), b_bc(nbf, 0);
    std::vector<double> b_pres(nbf, 0.0);
    for (int p = 0; p < nPatches; p++) {
        const dugks_patch_t& P = patches[p];
        if (P.start < 0 || P.size < 0 || P.start + P.size > nbf) { fail(h, DUGKS_ERR_INVALID, "patch %d out of range", p); return bail(DUGKS_ERR_INVALID); }
        if (P.kind < 0 || P.kind > DUGKS_PATCH_PRESSURE_OUT) { fail(h, DUGKS_ERR_INVALID, "patch %d: unknown kind %d", p, P.kind); return bail(DUGKS_ERR_INVALID); }
        for (int j = 0; j < P.size; j++) {
            b_kind[P.start + j] = P.kind;
            b_bc[P.start + j] = (P.U_bc == DUGKS_BC_ZERO_GRADIENT ? 1 : 0) | (P.T_bc == DUGKS_BC_ZERO_GRADIENT ? 2 : 0);
            b_pres[P.start + j] = P.pressure;
        }
        if (P.size > 0 && (P.kind == DUGKS_PATCH_DVM_SYMMETRY || P.kind == DUGKS_PATCH_SYMMETRY_PLANE)) h->has_sym = true;
        if (P.size > 0 && P.kind == DUGKS_PATCH_MAXWELL_WALL) h->has_wall = true;
    }
    for (int b = 0; b < nbf; b++)
        if (b_kind[b] < 0) { fail(h, DUGKS_ERR_INVALID, "boundary face %d belongs to no patch", b); return bail(DUGKS_ERR_INVALID); }

    // ---- discrete-velocity layout (global grid as fvDVM.C:140-220)
    const int ny = (D >= 2) ? n : 1, nz = (D == 3) ? n : 1;
    h->nxi = n * ny * nz;
    const int nbase = ny * nz;                      // base rows (iy, iz)
    // contiguous blocks of base rows per rank (any assignment is valid: sums commute;
    // the reference's round-robin is fvDVM.C:228-260)
    if (nbase < nranks) { fail(h, DUGKS_ERR_UNSUPPORTED, "%d velocity rows cannot be split over %d ranks", nbase, nranks); return bail(DUGKS_ERR_UNSUPPORTED); }
    auto row_begin = [&](int r) { return row_begin_of(nbase, nranks, r); };
    const int rb0 = row_begin(h->rank), rb1 = row_begin(h->rank + 1);
    int max_nbl = 0;
    for (int r = 0; r < nranks; r++) max_nbl = std::max(max_nbl, row_begin(r + 1) - row_begin(r));
    h->owner_rank_of_gid.resize(h->nxi);
    for (int r = 0; r < nranks; r++)
        for (int br = row_begin(r); br < row_begin(r + 1); br++)
            for (int ix = 0; ix < n; ix++) h->owner_rank_of_gid[br * n + ix] = r;
    RowLayout lay;
    if (build_row_layout(n, D, nranks, h->rank, dvset->Xis, dvset->weights, lay) != 0) {
        fail(h, DUGKS_ERR_UNSUPPORTED, "nDV = %d cannot be laid out (rows are cut into ix-chunks of <= %d points, <= %d table entries in all)", n, MAX_L, NT_MAX);
        return bail(DUGKS_ERR_UNSUPPORTED);
    }
    const int nch = lay.nch, L = lay.L;
    h->nch = nch; h->L = L; h->ntab = lay.ntab; h->Lt = lay.Lt;
    h->Rs = 32;
    std::vector<RowDesc>& rows = lay.rows;
    const int nreal = (int)rows.size();
    h->nrows = (nreal + h->Rs - 1) / h->Rs * h->Rs;
    h->nslab = h->nrows / h->Rs;
    h->nflat = (size_t)h->nslab * L * h->Rs;
    std::vector<double> row_y(h->nrows), row_z(h->nrows), row_w(h->nrows, 0.0);
    std::vector<int> row_cb(h->nrows);
    for (int k = 0; k < h->nrows; k++) {
        const RowDesc& rd = rows[std::min(k, nreal - 1)];   // padding rows duplicate the last row with weight 0
        row_y[k] = rd.y; row_z[k] = rd.z; row_cb[k] = rd.cbase;
        if (k < nreal && !rd.pad) row_w[k] = rd.w;
    }
    // tables along x: entries beyond n duplicate the last abscissa with weight 0
    std::vector<double> tx((size_t)5 * h->ntab);
    for (int t = 0; t < h->ntab; t++) {
        double x = dvset->Xis[std::min(t, n - 1)], w = (t < n) ? dvset->weights[t] : 0.0;
        tx[t] = x; tx[h->ntab + t] = w; tx[2 * h->ntab + t] = w * x; tx[3 * h->ntab + t] = w * x * x;
        tx[4 * h->ntab + t] = w * x * x * x;
    }
    // largest table span of any warp
    h->tabw = L;
    for (int k0 = 0; k0 < h->nrows; k0 += 32) {
        int mn = 1 << 30, mx = -1;
        for (int k = k0; k < k0 + 32; k++) { mn = std::min(mn, row_cb[k]); mx = std::max(mx, row_cb[k]); }
        const int len = rows[std::min(k0, nreal - 1)].len;
        h->tabw = std::max(h->tabw, mx + len - mn);
    }
    // flat index <-> global id, mirror partners
    h->flat_gid.assign(h->nflat, -1);
    std::vector<int> gid_flat(h->nxi, -1);
    for (int k = 0; k < nreal; k++) {
        int s = k / h->Rs, r = k % h->Rs;
        const RowDesc& rd = rows[k];
        if (rd.pad) continue;
        for (int i = 0; i < rd.len; i++) {
            int ix = rd.cbase + i;
            if (ix >= n) continue;
            int gid = (rd.iz * ny + rd.iy) * n + ix;
            size_t flat = ((size_t)s * L + i) * h->Rs + r;
            h->flat_gid[flat] = gid;
            gid_flat[gid] = (int)flat;
        }
    }
    for (int g = 0; g < h->nxi; g++)
        if (gid_flat[g] >= 0) { h->local_gids.push_back(g); h->local_flat.push_back(gid_flat[g]); }
    h->nvl = (int)h->local_gids.size();
    // mirror partners (fvDVM.C:162-164), as rows of the symmetry exchange buffer X: local flat indices when every
    // partner is local, global DV ids otherwise (the partner's values then arrive by the all-reduce of X)
    std::vector<int> mir_gid((size_t)3 * h->nflat, -1);
    for (size_t flat = 0; flat < h->nflat; flat++) {
        int g = h->flat_gid[flat];
        if (g < 0) continue;
        int ix = g % n, iy = (g / n) % ny, iz = g / (n * ny);
        int mg[3] = {(iz * ny + iy) * n + (n - 1 - ix),                       // fvDVM.C:162
                     (D >= 2) ? (iz * ny + (ny - 1 - iy)) * n + ix : 0,        // :163
                     (D == 3) ? ((nz - 1 - iz) * ny + iy) * n + ix : 0};       // :164
        for (int d = 0; d < 3; d++) mir_gid[(size_t)d * h->nflat + flat] = mg[d];
    }
    std::vector<int> xrow(h->nflat, -1), xmir((size_t)3 * h->nflat, -1), symface;
    if (h->has_sym) {
        bool axis_used[3] = {false, false, false};
        for (int p = 0; p < nPatches; p++) {
            const dugks_patch_t& P = patches[p];
            if ((P.kind != DUGKS_PATCH_DVM_SYMMETRY && P.kind != DUGKS_PATCH_SYMMETRY_PLANE) || P.size <= 0) continue;
            // mirror axis from the normal of the FIRST face of the patch (discreteVelocity.C:763,777-779)
            const double* s0 = mesh->Sf + (size_t)(nif + P.start) * 3;
            const double mag = std::sqrt(s0[0] * s0[0] + s0[1] * s0[1] + s0[2] * s0[2]);
            int axis = 0;
            double best = -1;
            for (int d = 0; d < 3; d++)
                if (std::fabs(s0[d] / mag) > best) { best = std::fabs(s0[d] / mag); axis = d; }
            if (!(best >= 1.0 - 1e-9) || axis >= D) {
                fail(h, DUGKS_ERR_UNSUPPORTED, "symmetry patch %d: normal is not along a solved axis (the reference's mirror-id rule, discreteVelocity.C:777-779, needs it)", p);
                return bail(DUGKS_ERR_UNSUPPORTED);
            }
            axis_used[axis] = true;
            SymPatch sp{P.start, P.size, axis, (int)symface.size(), s0[0], s0[1], s0[2]};
            h->sym_patches.push_back(sp);
            for (int j = 0; j < P.size; j++) symface.push_back(P.start + j);
        }
        h->nsym = (int)symface.size();
        // every rank must take the same decision (the exchange is a collective): does ANY rank miss a partner
        // along an axis some symmetry patch mirrors about?
        h->sym_exchange = false;
        for (int g = 0; g < h->nxi && !h->sym_exchange; g++) {
            const int ix = g % n, iy = (g / n) % ny, iz = g / (n * ny);
            const int mg[3] = {(iz * ny + iy) * n + (n - 1 - ix), (iz * ny + (ny - 1 - iy)) * n + ix,
                               ((nz - 1 - iz) * ny + iy) * n + ix};
            for (int d = 0; d < D; d++)
                if (axis_used[d] && h->owner_rank_of_gid[mg[d]] != h->owner_rank_of_gid[g]) h->sym_exchange = true;
        }
        h->sym_rows = h->sym_exchange ? h->nxi : (int)h->nflat;
        for (size_t flat = 0; flat < h->nflat; flat++) {
            const int g = h->flat_gid[flat];
            if (g < 0) continue;
            xrow[flat] = h->sym_exchange ? g : (int)flat;
            for (int d = 0; d < 3; d++) {
                const int mg = mir_gid[(size_t)d * h->nflat + flat];
                xmir[(size_t)d * h->nflat + flat] = h->sym_exchange ? mg : gid_flat[mg];
            }
        }
    }

    // ---- cell -> face CSR (internal faces first), geometry per entry, axis-aligned cells (build_host_csr)
    HostCsr csr;
    {
        std::string cerr;
        const int crc = build_host_csr(mesh, getenv("DUGKS_NO_AXIS") == nullptr /* test hook: general path everywhere */, csr, cerr);
        if (crc) { fail(h, crc, "%s", cerr.c_str()); return bail(crc); }
    }
    std::vector<int>&cnt = csr.cnt, &cnt_int = csr.cnt_int, &off = csr.off, &e_other = csr.e_other, &e_face = csr.e_face, &e_owner = csr.e_owner;
    std::vector<double>& e_geo = csr.e_geo;
    std::vector<unsigned char>& cell_cls = csr.cell_cls;
    h->n_axis = csr.n_axis; h->axis_ne = csr.axis_ne;
    const int ne = off[nc];
    std::vector<int> b_owner(nbf);
    std::vector<double> b_n((size_t)nbf * 3), b_invdc(nbf), b_Sf((size_t)nbf * 3), b_r((size_t)nbf * 3);
    for (int b = 0; b < nbf; b++) {
        int f = nif + b, o = mesh->owner[f];
        b_owner[b] = o;
        const double* S = mesh->Sf + (size_t)f * 3;
        double mag = std::sqrt(S[0] * S[0] + S[1] * S[1] + S[2] * S[2]);
        for (int d = 0; d < 3; d++) {
            b_n[(size_t)b * 3 + d] = S[d] / mag;                                 // discreteVelocity.C:438-442
            b_Sf[(size_t)b * 3 + d] = S[d];
            b_r[(size_t)b * 3 + d] = mesh->Cf[(size_t)f * 3 + d] - mesh->C[(size_t)o * 3 + d];
        }
        b_invdc[b] = 1.0 / mesh->deltaCoeffs[f];
    }

    // ---- upload static data
    StepArgs& A = h->A;
    memset(&A, 0, sizeof A);
    A.gas = h->gas; A.nm = h->nm;
    A.limiter_k = par->limiter_k;
    {   // L2 cache policies as kernel arguments (dugks_hot.cuh, RLX_POL)
        unsigned long long* d_pol = nullptr;
        unsigned long long pol[2] = {0, 0};
        TRYB(dev_alloc(h, &d_pol, 2));
        k_make_policies<<<1, 1, 0, h->stream>>>(d_pol);
        CUDA_TRY(h, cudaMemcpyAsync(pol, d_pol, sizeof pol, cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        A.pol_ef = pol[0]; A.pol_el = pol[1];
    }
    DevMesh& M = A.m;
    M.nc = nc; M.nif = nif; M.nbf = nbf; M.nf = nf;
    int *d_i; double* d_d;
    TRYB(dev_upload(h, &d_i, off)); M.cell_off = d_i;
    TRYB(dev_upload(h, &d_i, cnt_int)); M.cell_nint = d_i;
    {
        unsigned char* d_c = nullptr;
        TRYB(dev_upload(h, &d_c, cell_cls)); A.cell_cls = d_c;
        // traversal order of the cell kernels (build_cell_order): DUGKS_ORDER = wave (default) | tiled | morton | natural
        h->want_split = 2 * (long long)h->n_axis > nc && getenv("DUGKS_NO_SPLIT_AXIS") == nullptr;
        const char* ord_env = getenv("DUGKS_ORDER");
        int dev_sms = 148;
        cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, h->device);
        int nwarps = dev_sms * 3 * HOT_WARPS;   // warps in flight of the 3-CTAs/SM kernels (phase 1 axis-only, relax+update)
        if (const char* e = getenv("DUGKS_WAVE")) nwarps = std::max(1, atoi(e));
        int tile = 0;
        if (const char* e = getenv("DUGKS_TILE")) tile = std::max(1, atoi(e));
        std::vector<int> order, order2;   // order2: traversal of the phase-2 kernels of face-storage slabs where it differs
        build_cell_order(nc, D, mesh->C, h->want_split ? cell_cls.data() : nullptr, ord_env ? ord_env : "wave", tile, nwarps, order);
        // CTA pencils of phase 1 (3-D, h elided, every cell within the compile-time face count): DUGKS_PENCIL =
        // 2 (default: the pencil applies the half step itself) | 1 (pencil reads gBarP) | 0 (off)
        {
            int want_pen = 2;
            if (const char* e = getenv("DUGKS_PENCIL")) want_pen = std::max(0, std::min(2, atoi(e)));
            bool big = false;
            for (int c = 0; c < nc; c++) big = big || cnt[c] > FAST_NE;
            if (want_pen && h->want_split && D == 3 && !h->hasH && h->axis_ne == PEN_NE && !big && nch == 1 && L % PEN_CI == 0 &&
                getenv("DUGKS_NO_HOT") == nullptr) {
                PenHost ph;
                h->pen_grid = 2 * dev_sms;
                build_pencils(nc, off, e_other, cell_cls, h->pen_grid, ph);
                if (!ph.items.empty()) {
                    h->pen_mode = want_pen;
                    h->n_pen = (int)ph.order.size();
                    std::vector<char> inpen(nc, 0);
                    for (int c : ph.order) inpen[c] = 1;
                    std::vector<int> rest;
                    for (int c : order) if (!inpen[c]) rest.push_back(c);
                    order = ph.order;
                    order.insert(order.end(), rest.begin(), rest.end());
                    // Phase 2 (k_hot_relax_update: warp w takes items w, w + nWarps, ...) walks the pencil cells with
                    // nWarps / 4 work items in lock-step along x: both readers of a face then run within one sweep of
                    // the grid (50 MB of traffic apart).  In work-item order the z neighbour of a bundle was 4 sweeps
                    // away and the second read of its faces came from DRAM (1.2 of 9 GB per launch).
                    if (getenv("DUGKS_ORDER2_OFF") == nullptr) {
                        const int per_round = std::max(1, nwarps / PEN_WARPS);
                        for (size_t i0 = 0; i0 < ph.items.size(); i0 += per_round) {
                            const size_t i1 = std::min(ph.items.size(), i0 + per_round);
                            int smax = 0;
                            for (size_t i = i0; i < i1; i++) smax = std::max(smax, ph.items[i].nsteps);
                            for (int k = 0; k < smax; k++)
                                for (size_t i = i0; i < i1; i++)
                                    if (k < ph.items[i].nsteps)
                                        for (int l = 0; l < PEN_WARPS; l++) order2.push_back(ph.order[ph.items[i].item0 + PEN_WARPS * k + l]);
                        }
                        order2.insert(order2.end(), rest.begin(), rest.end());
                    }
                    PenItem* d_items = nullptr;
                    int *d_pc = nullptr, *d_ph = nullptr;
                    TRYB(dev_upload(h, &d_items, ph.items));
                    TRYB(dev_upload(h, &d_pc, ph.cells));
                    TRYB(dev_upload(h, &d_ph, ph.halo));
                    h->pen = PenArgs{};
                    h->pen.items = d_items; h->pen.nitems = (int)ph.items.size(); h->pen.cells = d_pc; h->pen.halo = d_ph;
                    h->pen_grid = std::min(h->pen_grid, (int)ph.items.size());
                    // fused mode: the kernels of the other cells still read gBarP of those cells and of their neighbours
                    std::vector<char> need(nc, 0);
                    for (int c = 0; c < nc; c++) {
                        if (inpen[c]) continue;
                        need[c] = 1;
                        for (int j = 0; j < cnt_int[c]; j++) need[e_other[off[c] + j]] = 1;
                    }
                    std::vector<int> hl;
                    for (int c : order) if (need[c]) hl.push_back(c);
                    h->n_halflist = (int)hl.size();
                    TRYB(dev_upload(h, &h->d_halflist, hl));
                }
            }
        }
        // per-cell record (CMETA_N ints) in traversal order: everything a warp needs to start a cell in one load level
        auto build_cmeta = [&](const std::vector<int>& ord, std::vector<int>& cmeta) {
            cmeta.assign((size_t)nc * CMETA_N, 0);
            for (int item = 0; item < nc; item++) {
                const int c = ord[item];
                int* rec = &cmeta[(size_t)item * CMETA_N];
                rec[20] = c;
                rec[0] = off[c];
                rec[1] = (cnt[c] & 0xff) | ((cnt_int[c] & 0xff) << 8) | ((int)cell_cls[c] << 16);
                unsigned char* kinds = reinterpret_cast<unsigned char*>(rec + 18);
                for (int j = 0; j < 8; j++) kinds[j] = 0xff;
                for (int j = 0; j < cnt[c] && j < 8; j++) {
                    const int e = off[c] + j;
                    rec[2 + j] = e_other[e];
                    rec[10 + j] = e_face[e] | (e_owner[e] ? (int)0x80000000u : 0);
                    if (e_other[e] < 0) kinds[j] = (unsigned char)b_kind[-1 - e_other[e]];
                }
            }
        };
        std::vector<int> cmeta;
        build_cmeta(order, cmeta);
        TRYB(dev_upload(h, &d_i, cmeta)); A.cmeta = d_i;
        A.cmeta2 = A.cmeta;
        if ((int)order2.size() == nc) {
            build_cmeta(order2, cmeta);
            TRYB(dev_upload(h, &d_i, cmeta)); A.cmeta2 = d_i;
        }
    }
    TRYB(dev_upload(h, &d_i, e_other)); M.e_other = d_i;
    TRYB(dev_upload(h, &d_i, e_face)); M.e_face = d_i;
    TRYB(dev_upload(h, &d_i, e_owner)); M.e_owner = d_i;
    TRYB(dev_upload(h, &d_d, e_geo)); M.e_geo = d_d;
    {
        std::vector<double> g12((size_t)ne * GEO12, 0.0);
        for (int e = 0; e < ne; e++) {
            const double* g = &e_geo[(size_t)e * 9];
            double* o = &g12[(size_t)e * GEO12];
            o[0] = g[0]; o[1] = g[1]; o[2] = g[2]; o[3] = g[6];
            o[4] = g[3]; o[5] = g[4]; o[6] = g[5]; o[7] = g[7];
            o[8] = g[8]; o[9] = (e_other[e] < 0) ? b_invdc[-1 - e_other[e]] : 0.0;
        }
        TRYB(dev_upload(h, &d_d, g12)); M.e_geo12 = d_d;
    }
    {
        // records of the second-generation kernels (dugks_hot.cuh): branch-free gradient coefficients
        std::vector<double> geo6((size_t)(ne + nc) * 6, 0.0), geoS((size_t)ne * 4, 0.0);
        for (int c = 0; c < nc; c++) {
            double* rec = &geo6[(size_t)(off[c] + c) * 6];
            for (int j = 0; j < cnt[c]; j++) {
                const int e = off[c] + j;
                const double* g = &e_geo[(size_t)e * 9];
                double* o = rec + 6 * (1 + j);
                double scale = 1.0;
                if (e_other[e] < 0) {
                    const int b = -1 - e_other[e];
                    // boundary value = cell + gamma/deltaCoeffs (fixedGradient); symmetryPlane adds nothing
                    scale = (b_kind[b] == K_SYMMETRY_PLANE) ? 0.0 : b_invdc[b];
                } else {
                    for (int d = 0; d < 3; d++) rec[d] -= g[d];
                }
                for (int d = 0; d < 3; d++) { o[d] = g[d] * scale; o[3 + d] = g[3 + d]; }
                const double sgn = e_owner[e] ? 1.0 : -1.0;
                for (int d = 0; d < 3; d++) geoS[(size_t)e * 4 + d] = sgn * g[6 + d];
            }
        }
        TRYB(dev_upload(h, &d_d, geo6)); A.geo6 = d_d;
        TRYB(dev_upload(h, &d_d, geoS)); A.geoS = d_d;
    }
    for (int c = 0; c < nc; c++) {
        if (cnt[c] > FAST_NE) h->n_big++;
        else h->max_ne_fast = std::max(h->max_ne_fast, cnt[c]);
    }
    for (int b = 0; b < nbf; b++)
        if (b_kind[b] == K_FAR_FIELD || b_kind[b] == K_PRESSURE_IN || b_kind[b] == K_PRESSURE_OUT) h->has_far = true;
    h->use_hot = getenv("DUGKS_NO_HOT") == nullptr && !(par->limiter_k > 0.0);   // test hook: first-generation kernels; limited gradient: generic kernels
    h->hot_ne = h->max_ne_fast <= 4 ? 4 : (h->max_ne_fast <= 6 ? 6 : 8);
    h->use_tma = getenv("DUGKS_NO_TMA") == nullptr;            // test hook: per-element LDG kernels instead
    h->ci = TMA_CI;
    h->use_fast = getenv("DUGKS_FORCE_GENERIC") == nullptr && !(par->limiter_k > 0.0);   // test hook: run every cell through the generic kernels
    if (h->Lt > 0) h->use_fast = false;                        // only the generic and second-generation kernels know short rows
    TRYB(dev_upload(h, &d_d, std::vector<double>(mesh->V, mesh->V + nc))); M.V = d_d;
    TRYB(dev_upload(h, &d_i, b_owner)); M.b_owner = d_i;
    TRYB(dev_upload(h, &d_i, b_kind)); M.b_kind = d_i;
    TRYB(dev_upload(h, &d_d, b_n)); M.b_n = d_d;
    TRYB(dev_upload(h, &d_d, b_invdc)); M.b_invdc = d_d;
    TRYB(dev_upload(h, &d_d, b_Sf)); M.b_Sf = d_d;
    TRYB(dev_upload(h, &d_d, b_r)); M.b_r = d_d;
    TRYB(dev_upload(h, &d_d, std::vector<double>(mesh->deltaCoeffs, mesh->deltaCoeffs + nif))); M.dcoef_int = d_d;
    DevDV& V = A.dv;
    V.L = L; V.Lt = h->Lt; V.Rs = h->Rs; V.nslab = h->nslab; V.ntab = h->ntab; V.hasH = h->hasH; V.tabw = h->tabw;
    TRYB(dev_upload(h, &d_d, tx)); V.tx = d_d;
    h->tx_host = tx;
    TRYB(dev_upload(h, &d_d, row_y)); V.row_y = d_d;
    TRYB(dev_upload(h, &d_d, row_z)); V.row_z = d_d;
    TRYB(dev_upload(h, &d_d, row_w)); V.row_w = d_d;
    TRYB(dev_upload(h, &d_i, row_cb)); V.row_cbase = d_i;
    TRYB(dev_upload(h, &h->d_bc, b_bc));
    TRYB(dev_upload(h, &h->d_pres, b_pres));
    TRYB(dev_upload(h, &h->d_xrow, xrow));
    TRYB(dev_upload(h, &h->d_xmir, xmir));
    TRYB(dev_upload(h, &h->d_symface, symface));

    // ---- macros
    std::vector<double> cmac((size_t)nc * MAC_N, 0.0), bmac((size_t)nbf * 5, 0.0), fmac((size_t)nf * MAC_N, 0.0);
    for (int c = 0; c < nc; c++) {
        cmac[(size_t)c * MAC_N] = rho[c];
        for (int d = 0; d < 3; d++) cmac[(size_t)c * MAC_N + 1 + d] = U[(size_t)c * 3 + d];
        cmac[(size_t)c * MAC_N + 4] = T[c];
    }
    for (int b = 0; b < nbf; b++) {
        bmac[(size_t)b * 5] = rho_b[b];
        for (int d = 0; d < 3; d++) bmac[(size_t)b * 5 + 1 + d] = U_b[(size_t)b * 3 + d];
        bmac[(size_t)b * 5 + 4] = T_b[b];
    }
    // Usurf = fvc::interpolate(Uvol, "linear") for the first Courant number (fvDVM.C:1074) [OF-lib]
    for (int f = 0; f < nif; f++) {
        int o = mesh->owner[f], nb = mesh->neighbour[f];
        const double *S = mesh->Sf + (size_t)f * 3, *Cf = mesh->Cf + (size_t)f * 3;
        double sn = 0, sp = 0;
        for (int d = 0; d < 3; d++) {
            sn += S[d] * (mesh->C[(size_t)nb * 3 + d] - Cf[d]);
            sp += S[d] * (Cf[d] - mesh->C[(size_t)o * 3 + d]);
        }
        sn = std::fabs(sn); sp = std::fabs(sp);
        double w = sn / (sp + sn);
        for (int d = 0; d < 3; d++) fmac[(size_t)f * MAC_N + 1 + d] = w * U[(size_t)o * 3 + d] + (1 - w) * U[(size_t)nb * 3 + d];
    }
    for (int b = 0; b < nbf; b++)
        for (int d = 0; d < 3; d++) fmac[(size_t)(nif + b) * MAC_N + 1 + d] = U_b[(size_t)b * 3 + d];
    if (nbf > 0) {
        h->last_rho_b.assign(rho_b, rho_b + nbf);
        h->last_U_b.assign(U_b, U_b + (size_t)3 * nbf);
        h->last_T_b.assign(T_b, T_b + nbf);
    }
    TRYB(dev_upload(h, &A.cmac, cmac));
    TRYB(dev_upload(h, &A.bmac, bmac));
    TRYB(dev_upload(h, &A.fmac, fmac));

    // ---- state
    size_t free_b = 0, total_b = 0;
    CUDAB(cudaMemGetInfo(&free_b, &total_b));
    const size_t ncell_dv = (size_t)nc * h->nflat, nb_dv = (size_t)nbf * h->nflat;
    const int nfld = h->hasH ? 2 : 1;
    // gBarP: at least one slab block (face-storage slabs share a transient one), at most one per slab
    size_t need = (ncell_dv + (size_t)nc * L * h->Rs + 3 * nb_dv + (size_t)h->sym_rows * h->nsym + (size_t)nif * L * h->Rs) * nfld * sizeof(double) +
                  ((size_t)(2 * nif + nbf) + nc) * h->nm * sizeof(double);
    if (need > free_b) {
        fail(h, DUGKS_ERR_NOMEM, "state needs %.2f GB of device memory, only %.2f GB free (nCells=%d, local DVs=%d, h %s)",
             need / 1e9, free_b / 1e9, nc, h->nvl, h->hasH ? "stored" : "elided");
        return bail(DUGKS_ERR_NOMEM);
    }
    // HOT_PAD: the bulk copies of a tail chunk always fetch HOT_CI velocity points
    TRYB(dev_alloc(h, &A.gt, ncell_dv + HOT_PAD));
    // gBarP (A.gb, A.hb) is allocated with the face storage below: how much of it must persist depends on it
    TRYB(dev_alloc(h, &A.gsb, nb_dv + HOT_PAD));
    TRYB(dev_alloc(h, &h->gam_a_g, nb_dv + HOT_PAD));   // the second copy (ping-pong), where one is needed, follows the kernel choice below
    // the flux buffer of the recompute path (A.fbuf_*) is placed with the face storage below
    if (h->hasH) {
        TRYB(dev_alloc(h, &A.ht, ncell_dv + HOT_PAD));
        TRYB(dev_alloc(h, &A.hsb, nb_dv + HOT_PAD));
        TRYB(dev_alloc(h, &h->gam_a_h, nb_dv + HOT_PAD));
    }
    TRYB(dev_alloc(h, &A.fcoef, (size_t)nf * FCOEF_N));
    TRYB(dev_alloc(h, &A.ccoef, (size_t)nc * FCOEF_N));
    if (h->has_sym) {
        // g rows, then h rows: one buffer, one collective
        TRYB(dev_alloc(h, &h->sym_Xg, (size_t)h->sym_rows * h->nsym * nfld));
        if (h->hasH) h->sym_Xh = h->sym_Xg + (size_t)h->sym_rows * h->nsym;
    }
    TRYB(dev_alloc(h, &A.fslot, ((size_t)2 * nif + nbf) * h->nm));
    TRYB(dev_alloc(h, &A.cslot, (size_t)nc * h->nm));
    if (nranks > 1) TRYB(dev_alloc(h, &h->fslot_fold, (size_t)nf * h->nm));
    TRYB(dev_alloc(h, &h->wall_cin, (size_t)nbf * h->nm));
    TRYB(dev_alloc(h, &h->wall_in, (size_t)nbf));
    TRYB(dev_alloc(h, &A.wall_diag, (size_t)nbf * 12));
    TRYB(dev_alloc(h, &h->d_co, 2 + 2 * COURANT_BLOCKS));
    TRYB(dev_alloc(h, &h->d_conv_old, (size_t)5 * nc));
    TRYB(dev_alloc(h, &h->d_conv, 6 + 6 * COURANT_BLOCKS));
    TRYB(dev_alloc(h, &h->d_bstage, (size_t)5 * nbf));
    A.wall_cin = h->wall_cin; A.wall_in = h->wall_in;
    A.gam_old_g = h->gam_a_g; A.gam_old_h = h->gam_a_h;

    // ---- dynamic shared memory sizes
    h->smem_out1 = (size_t)5 * NT_MAX * 8 + (size_t)WARPS_PER_CTA * STAGE_BYTES;
    h->smem_out2 = h->smem_out1 + (size_t)WARPS_PER_CTA * ACC_FACES * 3 * h->tabw * 8;
    h->smem_bnd = (size_t)NT_MAX * 8 + (size_t)WARPS_PER_CTA * STAGE_BYTES;
    h->smem_upd = h->smem_out1;
    h->fsmem_out1 = (size_t)NT_MAX * 6 * 8 + (size_t)WARPS_PER_CTA * FAST_STAGE_BYTES;
    h->fsmem_out2 = h->fsmem_out1 + (size_t)WARPS_PER_CTA * ((size_t)ACC_FACES * 3 * h->tabw + ACC_FACES * 3 * 32 + ACC_FACES * 2) * 8;
    h->fsmem_upd = h->fsmem_out1;
    {
        const int nfld2 = h->hasH ? 2 : 1;
        auto tma_bytes = [&](int nstream, size_t extra_d) {
            return (size_t)NT_MAX * 6 * 8 + (size_t)WARPS_PER_CTA * (TMA_META_BYTES + (TMA_STAGES * tma_stage_doubles(nstream * nfld2, h->ci) + extra_d) * 8);
        };
        h->tma_tw = h->tabw <= 32 ? 32 : 64;
        const size_t extra2 = (size_t)ACC_FACES * 3 * h->tma_tw + ACC_FACES * 3 * 32 + ACC_FACES * 2;
        if (h->tabw > 64) h->use_tma = false;
        h->tsmem_out1 = tma_bytes(h->max_ne_fast + 1, 0);
        h->tsmem_out2 = tma_bytes(h->max_ne_fast + 1, extra2);
        h->tsmem_upd = tma_bytes(h->max_ne_fast + 2, 0);
        if (std::max(h->tsmem_out2, h->tsmem_upd) > 220 * 1024) h->use_tma = false;
    }
    if ((double)std::max(nc, nf) * L * h->Rs >= 2147483647.0) {
        fail(h, DUGKS_ERR_UNSUPPORTED, "mesh too large for 32-bit slab offsets (%d faces x %d slab DVs)", nf, L * h->Rs);
        return bail(DUGKS_ERR_UNSUPPORTED);
    }
    if (h->hasH) {
        CUDAB(cudaFuncSetAttribute(k_cell_outgoing<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_out2));
        CUDAB(cudaFuncSetAttribute(k_cell_outgoing_fast<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->fsmem_out2));
        if (h->use_tma) {
            CUDAB(cudaFuncSetAttribute(k_cell_outgoing_tma<1, true, TMA_CI, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->tsmem_out1));
            CUDAB(cudaFuncSetAttribute(k_cell_outgoing_tma<2, true, TMA_CI, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->tsmem_out2));
            CUDAB(cudaFuncSetAttribute(k_cell_outgoing_tma<2, true, TMA_CI, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->tsmem_out2));
            CUDAB(cudaFuncSetAttribute(k_cell_update_tma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->tsmem_upd));
        }
    } else {
        if (h->use_tma) {
            CUDAB(cudaFuncSetAttribute(k_cell_outgoing_tma<1, false, TMA_CI, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->tsmem_out1));
            CUDAB(cudaFuncSetAttribute(k_cell_outgoing_tma<2, false, TMA_CI, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->tsmem_out2));
            CUDAB(cudaFuncSetAttribute(k_cell_outgoing_tma<2, false, TMA_CI, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->tsmem_out2));
            CUDAB(cudaFuncSetAttribute(k_cell_update_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->tsmem_upd));
        }
        CUDAB(cudaFuncSetAttribute(k_cell_outgoing<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_out2));
        CUDAB(cudaFuncSetAttribute(k_cell_outgoing_fast<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->fsmem_out2));
    }

    // ---- second-generation kernels: shared-memory plan, static upwind range codes
    if (h->tabw > 64 || L > 32) h->use_hot = false;
    if (h->use_hot) TRYB(DUGKS_BY_H(h, hot_configure, h));
    if (h->use_hot) {
        uint4* d_upw = nullptr;
        int* d_bad = nullptr;
        TRYB(dev_alloc(h, &d_upw, (size_t)h->nslab * nc * 32, false));
        TRYB(dev_alloc(h, &d_bad, 1 + (size_t)h->nslab));   // [0]: not range shaped; [1 + slab]: tie on a y/z face of an axis-aligned cell
        long long total = (long long)h->nslab * nc * 32;
        int grid = (int)std::min<long long>((total + 255) / 256, 148LL * 64);
        k_build_upwind<<<grid, 256, 0, h->stream>>>(A, d_upw, d_bad);
        TRYB(check_launch(h, "k_build_upwind"));
        std::vector<int> flags(1 + (size_t)h->nslab, 0);
        CUDAB(cudaMemcpyAsync(flags.data(), d_bad, flags.size() * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CUDAB(cudaStreamSynchronize(h->stream));
        const int bad = flags[0];
        h->slab_pair_ok.assign(h->nslab, 0);
        for (int s = 0; s < h->nslab; s++)
            // rows cut into ix-chunks (more than 32 points per direction): the x masks differ from lane to lane,
            // hot_axis_item would spend its time in the slow x path; those layouts keep the unified launch
            h->slab_pair_ok[s] = flags[1 + s] == 0 && dv_len(A.dv, s) % HOT_AXIS_CU == 0 && h->nch == 1;
        // abscissae that are not ascending give upwind sets that are not ranges: first-generation kernels
        if (bad) h->use_hot = false;
        A.upw = d_upw;
    }
    if (!h->use_hot || h->hsmem_half == 0) h->pen_mode = 0;
    // Lagged boundary gradient (discreteVelocity.C:462-468): the second-generation kernels update it IN PLACE - the warp
    // that owns a boundary cell reads the old value of a (face, velocity) before it writes the new one, in phase 1 for
    // the face-storage slabs and in the second gradient pass (phase 2) for the recompute slabs, which read the old value
    // twice.  The first-generation kernels and k_bnd_outgoing (far-field / pressure patches) read the old array after the
    // cell kernel has run: they keep two copies.  One copy less is 4.4 GB at 64^3 x 28^3: one more face-storage slab.
    h->gam_single = h->use_hot && h->n_big == 0 && !h->has_far && getenv("DUGKS_GAM_PINGPONG") == nullptr;
    if (h->gam_single) { h->gam_b_g = h->gam_a_g; h->gam_b_h = h->gam_a_h; }
    else {
        TRYB(dev_alloc(h, &h->gam_b_g, nb_dv + HOT_PAD));
        if (h->hasH) TRYB(dev_alloc(h, &h->gam_b_h, nb_dv + HOT_PAD));
    }
    A.gam_new_g = h->gam_b_g; A.gam_new_h = h->gam_b_h;
    A.gam_late = h->gam_single ? 1 : 0;

    // ---- face-storage slabs: as many as device memory allows (all of them from 2 GPUs up at the
    // 64^3 x 28^3 size; about half on one GPU).  Cells with too many faces need the flux-buffer path.
    // A face-storage slab needs its gBarP only during phase 1 (its update reads w = -1/3 gTilde + 4/3 gBarP,
    // which the half-step kernel leaves in place of gTilde), so all of them share ONE transient gBarP block
    // (wmode); the other slabs keep theirs for the second gradient pass.
    size_t gb_blocks = (size_t)h->nslab;
    if (h->use_hot && h->n_big == 0 && nif > 0) {
        size_t free2 = 0, total2 = 0;
        CUDAB(cudaMemGetInfo(&free2, &total2));
        const size_t per_slab = (size_t)nif * L * h->Rs * sizeof(double) * nfld;      // face values of a slab
        const size_t per_gb = (size_t)nc * L * h->Rs * sizeof(double) * nfld;          // gBarP of a slab
        // left free: CUDA context growth and the caller's own allocations; NCCL buffers when there are peers
        const size_t reserve = nranks > 1 ? (size_t)3 << 30 : (size_t)1 << 30;   // 1 GiB alone (everything else of the handle is allocated by now), 3 GiB with peers
        const bool can_w = h->hsmem_half > 0 && getenv("DUGKS_NO_WMODE") == nullptr;   // test hook: persistent gBarP everywhere
        long long avail = free2 > reserve ? (long long)(free2 - reserve) : 0;
        // dugks_par_t.scratch_bytes: the caller's cap on what the kept face values (and the gBarP blocks that go
        // with them) may take; the slabs that do not fit recompute their face values in phase 2
        if (par->scratch_bytes > 0) avail = std::min<long long>(avail, (long long)std::min<size_t>(par->scratch_bytes, (size_t)1 << 62));
        // k face-storage slabs need their face values and ONE transient gBarP block; the other slabs a gBarP block
        // each and a flux buffer - which is the face storage of slab 0: phase 2 runs the face-storage slabs first, and
        // slab 0's values are consumed by then
        long long fit = 0;
        for (long long k = h->nslab; k > 0; k--) {
            const long long blocks = can_w ? (h->nslab - k) + 1 : h->nslab;
            if (blocks * (long long)per_gb + k * (long long)per_slab <= avail) { fit = k; break; }
        }
        if (getenv("DUGKS_VERBOSE"))
            fprintf(stderr, "dugks: face storage: %.3f GB free of %.3f, reserve %.3f, available %.3f; %.3f GB of face values and %.3f GB of gBarP per slab: %lld of %d slabs keep their face values (held so far %.3f GB)\n",
                    free2 / 1e9, total2 / 1e9, reserve / 1e9, avail / 1e9, per_slab / 1e9, per_gb / 1e9, fit, h->nslab, h->dev_bytes / 1e9);
        if (const char* e = getenv("DUGKS_KEEP_SLABS")) fit = std::min<long long>(fit, atoll(e));   // test hook
        h->n_keep = (int)std::max<long long>(0, std::min<long long>(fit, h->nslab));
        h->wmode = can_w && h->n_keep > 0;
        if (h->wmode) gb_blocks = (size_t)(h->nslab - h->n_keep) + 1;
    }
    TRYB(dev_alloc(h, &h->gb_store, gb_blocks * nc * L * h->Rs + HOT_PAD));
    if (h->hasH) TRYB(dev_alloc(h, &h->hb_store, gb_blocks * nc * L * h->Rs + HOT_PAD));
    A.gb = h->gb_store; A.hb = h->hb_store;
    if (h->n_keep > 0) {
        TRYB(dev_alloc(h, &h->fkeep_g, (size_t)h->n_keep * nif * L * h->Rs + HOT_PAD, false));
        if (h->hasH) TRYB(dev_alloc(h, &h->fkeep_h, (size_t)h->n_keep * nif * L * h->Rs + HOT_PAD, false));
        // tie points are written by one side only and padding rows by nobody: start from zeros
        CUDAB(cudaMemsetAsync(h->fkeep_g, 0, ((size_t)h->n_keep * nif * L * h->Rs + HOT_PAD) * sizeof(double), h->stream));
        if (h->hasH) CUDAB(cudaMemsetAsync(h->fkeep_h, 0, ((size_t)h->n_keep * nif * L * h->Rs + HOT_PAD) * sizeof(double), h->stream));
        A.fbuf_g = h->fkeep_g; A.fbuf_h = h->fkeep_h;      // flux buffer of the recompute slabs = face storage of slab 0 (phase 2 only)
    } else {
        TRYB(dev_alloc(h, &A.fbuf_g, (size_t)nif * L * h->Rs + HOT_PAD));
        if (h->hasH) TRYB(dev_alloc(h, &A.fbuf_h, (size_t)nif * L * h->Rs + HOT_PAD));
    }

    // ---- collective backend
    if (nranks > 1 && !h->reduce) {
        if (!par->nccl_unique_id) { fail(h, DUGKS_ERR_COMM, "nRanks = %d needs either a reduce callback or an NCCL unique id", nranks); return bail(DUGKS_ERR_COMM); }
        std::string err;
        if (!g_nccl.load(err)) { fail(h, DUGKS_ERR_COMM, "%s", err.c_str()); return bail(DUGKS_ERR_COMM); }
        nccl_uid_t id;
        memcpy(&id, par->nccl_unique_id, sizeof id);
        int nrc = g_nccl.init_rank(&h->nccl_comm, nranks, id, h->rank);
        if (nrc) { fail(h, DUGKS_ERR_COMM, "ncclCommInitRank failed: %s", g_nccl.errstr ? g_nccl.errstr(nrc) : "?"); return bail(DUGKS_ERR_COMM); }
    }

    TRYB(DUGKS_BY_H(h, init_state, h));
    // Told = T; rhoOld = rho; Uold = U (createFields.H:80-82)
    k_convergence_init<<<(nc + 255) / 256, 256, 0, h->stream>>>(h->A, h->d_conv_old);
    TRYB(check_launch(h, "k_convergence_init"));
    CUDAB(cudaStreamSynchronize(h->stream));
    CUDAB(cudaGetLastError());
#undef TRYB
#undef CUDAB
    *out = h;
    return 0;
}

extern "C" int dugks_step(dugks_handle_t* h, double dt) {
    if (!h) return DUGKS_ERR_INVALID;
    if (!(dt > 0.0)) return fail(h, DUGKS_ERR_INVALID, "dugks_step: dt must be positive");
    CUDA_TRY(h, cudaSetDevice(h->device));
    return DUGKS_BY_H(h, step_impl, h, dt);
}

extern "C" int dugks_sync(dugks_handle_t* h) {
    if (!h) return DUGKS_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    CUDA_TRY(h, cudaGetLastError());
    return 0;
}

extern "C" void* dugks_stream(dugks_handle_t* h) { return h ? (void*)h->stream : nullptr; }

// D2H through a pinned staging buffer owned by the handle (pageable copies run at a fraction of the
// PCIe rate and the macro fields are read every step by the host's time loop)
static int fetch_pinned(dugks_handle* h, const double* dev, size_t n, const double** out) {
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (n > h->pin_count) {
        if (h->pin) cudaFreeHost(h->pin);
        h->pin = nullptr; h->pin_count = 0;
        CUDA_TRY(h, cudaHostAlloc((void**)&h->pin, n * sizeof(double), cudaHostAllocDefault));
        h->pin_count = n;
    }
    CUDA_TRY(h, cudaMemcpyAsync(h->pin, dev, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    *out = h->pin;
    return 0;
}
static int fetch(dugks_handle* h, const double* dev, size_t n, std::vector<double>& host) {
    const double* p = nullptr;
    int rc = fetch_pinned(h, dev, n, &p);
    if (rc) return rc;
    host.assign(p, p + n);
    return 0;
}

static void unpack_macros(const double* m, size_t n, double* rho, double* U, double* T, double* q, double* tau) {
    for (size_t k = 0; k < n; k++) {
        const double* p = m + k * MAC_N;
        if (rho) rho[k] = p[0];
        if (U) { U[3 * k] = p[1]; U[3 * k + 1] = p[2]; U[3 * k + 2] = p[3]; }
        if (T) T[k] = p[4];
        if (tau) tau[k] = p[5];
        if (q) { q[3 * k] = p[6]; q[3 * k + 1] = p[7]; q[3 * k + 2] = p[8]; }
    }
}

// True if p points into page-locked host memory (cudaHostAlloc / cudaHostRegister, e.g. dugks_host_register).
static bool host_pinned(const void* p) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

__global__ void k_unpack_macros(const double* m, int n, double* soa) {
    // soa = rho[n] | U[n][3] | T[n] | q[n][3] | tau[n]
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const double* p = m + (size_t)k * MAC_N;
    const size_t N = (size_t)n;
    soa[k] = p[0];
    soa[N + 3 * (size_t)k] = p[1]; soa[N + 3 * (size_t)k + 1] = p[2]; soa[N + 3 * (size_t)k + 2] = p[3];
    soa[4 * N + k] = p[4];
    soa[5 * N + 3 * (size_t)k] = p[6]; soa[5 * N + 3 * (size_t)k + 1] = p[7]; soa[5 * N + 3 * (size_t)k + 2] = p[8];
    soa[8 * N + k] = p[5];
}

// Macro fields into the caller's arrays.  Page-locked arrays (dugks_host_register) are filled by asynchronous copies
// straight from the device, laid out per field there; pageable ones through the handle's pinned staging buffer and a
// host loop (4.9 ms instead of 0.5 for the 262,144 cells of the 64^3 case).
static int get_macros(dugks_handle* h, const double* dev, size_t n, double* rho, double* U, double* T, double* q, double* tau) {
    const bool all_pinned = (rho || U || T || q || tau) && (!rho || host_pinned(rho)) && (!U || host_pinned(U)) && (!T || host_pinned(T)) &&
                            (!q || host_pinned(q)) && (!tau || host_pinned(tau));
    if (!all_pinned) {
        const double* m = nullptr;
        int rc = fetch_pinned(h, dev, n * MAC_N, &m);
        if (rc) return rc;
        unpack_macros(m, n, rho, U, T, q, tau);
        return 0;
    }
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (n * MAC_N > h->soa_count) {
        if (h->d_soa) cudaFree(h->d_soa);
        h->d_soa = nullptr; h->soa_count = 0;
        CUDA_TRY(h, cudaMalloc((void**)&h->d_soa, n * MAC_N * sizeof(double)));
        h->soa_count = n * MAC_N;
    }
    k_unpack_macros<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(dev, (int)n, h->d_soa);
    int rc = check_launch(h, "k_unpack_macros");
    if (rc) return rc;
    const double* s = h->d_soa;
    if (rho) CUDA_TRY(h, cudaMemcpyAsync(rho, s, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (U) CUDA_TRY(h, cudaMemcpyAsync(U, s + n, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (T) CUDA_TRY(h, cudaMemcpyAsync(T, s + 4 * n, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (q) CUDA_TRY(h, cudaMemcpyAsync(q, s + 5 * n, 3 * n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (tau) CUDA_TRY(h, cudaMemcpyAsync(tau, s + 8 * n, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int dugks_get_cell_macros(dugks_handle_t* h, double* rho, double* U, double* T, double* q, double* tau) {
    if (!h) return DUGKS_ERR_INVALID;
    return get_macros(h, h->A.cmac, (size_t)h->nc, rho, U, T, q, tau);
}

extern "C" int dugks_get_face_macros(dugks_handle_t* h, double* rho, double* U, double* T, double* q, double* tau) {
    if (!h) return DUGKS_ERR_INVALID;
    return get_macros(h, h->A.fmac, (size_t)h->nf, rho, U, T, q, tau);
}

// Page-lock / release caller-owned host arrays (OpenFOAM field storage) so that the accessors copy straight into them.
extern "C" int dugks_host_register(void* p, size_t bytes) {
    if (!p || !bytes) return DUGKS_ERR_INVALID;
    cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterDefault);
    if (e == cudaErrorHostMemoryAlreadyRegistered) { (void)cudaGetLastError(); return 0; }
    if (e != cudaSuccess) return fail(nullptr, DUGKS_ERR_CUDA, "cudaHostRegister of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    return 0;
}
extern "C" int dugks_host_unregister(void* p) {
    if (!p) return DUGKS_ERR_INVALID;
    cudaError_t e = cudaHostUnregister(p);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return fail(nullptr, DUGKS_ERR_CUDA, "cudaHostUnregister failed: %s", cudaGetErrorString(e)); }
    return 0;
}

extern "C" int dugks_get_boundary_macros(dugks_handle_t* h, double* rho_b, double* U_b, double* T_b) {
    if (!h) return DUGKS_ERR_INVALID;
    std::vector<double> m;
    int rc = fetch(h, h->A.bmac, (size_t)h->nbf * 5, m);
    if (rc) return rc;
    for (int b = 0; b < h->nbf; b++) {
        if (rho_b) rho_b[b] = m[(size_t)b * 5];
        if (U_b) for (int d = 0; d < 3; d++) U_b[3 * b + d] = m[(size_t)b * 5 + 1 + d];
        if (T_b) T_b[b] = m[(size_t)b * 5 + 4];
    }
    return 0;
}

extern "C" int dugks_set_boundary_macros(dugks_handle_t* h, const double* rho_b, const double* U_b, const double* T_b) {
    if (!h) return DUGKS_ERR_INVALID;
    {
        // arrays identical to the ones set last (the usual case: fixedValue patches that do not vary in
        // time) are a no-op: nothing is copied and the wall in-flux constants stay valid
        const size_t nb = (size_t)h->nbf;
        auto same = [](const double* p, const std::vector<double>& last, size_t n) {
            return p == nullptr || (last.size() == n && memcmp(p, last.data(), n * sizeof(double)) == 0);
        };
        if (same(rho_b, h->last_rho_b, nb) && same(U_b, h->last_U_b, 3 * nb) && same(T_b, h->last_T_b, nb)) return 0;
        if (rho_b) h->last_rho_b.assign(rho_b, rho_b + nb);
        if (U_b) h->last_U_b.assign(U_b, U_b + 3 * nb);
        if (T_b) h->last_T_b.assign(T_b, T_b + nb);
    }
    if (h->nbf == 0) return 0;
    CUDA_TRY(h, cudaSetDevice(h->device));
    // H2D of the fields that were given into a staging buffer, scattered into bmac on the device
    // (stream ordered: no device->host round trip, no synchronisation)
    const size_t nb = (size_t)h->nbf;
    double *d_rho = nullptr, *d_U = nullptr, *d_T = nullptr;
    if (rho_b) { d_rho = h->d_bstage; CUDA_TRY(h, cudaMemcpyAsync(d_rho, rho_b, nb * sizeof(double), cudaMemcpyHostToDevice, h->stream)); }
    if (U_b) { d_U = h->d_bstage + nb; CUDA_TRY(h, cudaMemcpyAsync(d_U, U_b, 3 * nb * sizeof(double), cudaMemcpyHostToDevice, h->stream)); }
    if (T_b) { d_T = h->d_bstage + 4 * nb; CUDA_TRY(h, cudaMemcpyAsync(d_T, T_b, nb * sizeof(double), cudaMemcpyHostToDevice, h->stream)); }
    k_set_bmac<<<(h->nbf + 127) / 128, 128, 0, h->stream>>>(h->A, h->d_bc, d_rho, d_U, d_T);
    int rc = check_launch(h, "k_set_bmac");
    if (rc) return rc;
    rc = DUGKS_BY_H(h, compute_wall_constants, h);
    if (rc) return rc;
    // the caller's arrays are borrowed for the call only: copies from pinned memory must have run
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    return 0;
}

extern "C" int dugks_get_wall_diag(dugks_handle_t* h, double* qWall, double* stressWall) {
    if (!h) return DUGKS_ERR_INVALID;
    std::vector<double> m;
    int rc = fetch(h, h->A.wall_diag, (size_t)h->nbf * 12, m);
    if (rc) return rc;
    for (int b = 0; b < h->nbf; b++) {
        if (qWall) for (int d = 0; d < 3; d++) qWall[3 * b + d] = m[(size_t)b * 12 + d];
        if (stressWall) for (int d = 0; d < 9; d++) stressWall[9 * b + d] = m[(size_t)b * 12 + 3 + d];
    }
    return 0;
}

extern "C" int dugks_courant(dugks_handle_t* h, double dt, double* maxCo, double* meanCo) {
    if (!h) return DUGKS_ERR_INVALID;
    if (h->nif == 0) { if (maxCo) *maxCo = 0; if (meanCo) *meanCo = 0; return 0; }
    CUDA_TRY(h, cudaSetDevice(h->device));
    StepArgs a = h->A;
    const int nblk = std::max(1, std::min(COURANT_BLOCKS, (h->nif + 255) / 256));
    k_courant<<<nblk, 256, 0, h->stream>>>(a, std::sqrt((double)h->D) * h->xiMax, h->d_co + 2);
    int rc = check_launch(h, "k_courant");
    if (rc) return rc;
    k_courant_fold<<<1, 32, 0, h->stream>>>(h->d_co + 2, nblk, h->d_co);
    if ((rc = check_launch(h, "k_courant_fold"))) return rc;
    double out[2];
    CUDA_TRY(h, cudaMemcpyAsync(out, h->d_co, sizeof out, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    if (maxCo) *maxCo = out[0] * dt;
    if (meanCo) *meanCo = out[1] / h->nif * dt;
    return 0;
}

extern "C" int dugks_convergence(dugks_handle_t* h, double change[3]) {
    if (!h || !change) return DUGKS_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    const int nblk = std::max(1, std::min(COURANT_BLOCKS, (h->nc + 255) / 256));
    k_convergence<<<nblk, 256, 0, h->stream>>>(h->A, h->d_conv_old, h->d_conv + 6);
    int rc = check_launch(h, "k_convergence");
    if (rc) return rc;
    k_convergence_fold<<<1, 32, 0, h->stream>>>(h->d_conv + 6, nblk, h->d_conv);
    if ((rc = check_launch(h, "k_convergence_fold"))) return rc;
    double out[6];
    CUDA_TRY(h, cudaMemcpyAsync(out, h->d_conv, sizeof out, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    change[0] = out[0] / out[1];   // dugksFoam.C:97
    change[1] = out[2] / out[3];   // :98
    change[2] = out[4] / out[5];   // :99
    return 0;
}

extern "C" int dugks_local_dvs(dugks_handle_t* h, int32_t* ids, int32_t* n) {
    if (!h || !n) return DUGKS_ERR_INVALID;
    *n = h->nvl;
    if (ids) memcpy(ids, h->local_gids.data(), sizeof(int32_t) * h->nvl);
    return 0;
}

extern "C" int dugks_sizes(dugks_handle_t* h, int32_t* nXi, int32_t* nXiLocal, int32_t* nCells, int32_t* nFaces) {
    if (!h) return DUGKS_ERR_INVALID;
    if (nXi) *nXi = h->nxi;
    if (nXiLocal) *nXiLocal = h->nvl;
    if (nCells) *nCells = h->nc;
    if (nFaces) *nFaces = h->nf;
    return 0;
}

// state <-> DV-major host arrays; device layout [slab][cell][i][r]
static int state_io(dugks_handle* h, double* dev, double* host, bool to_host) {
    const size_t slabsz = (size_t)h->L * h->Rs;
    std::vector<double> tmp((size_t)h->nc * slabsz);
    for (int s = 0; s < h->nslab; s++) {
        double* dslab = dev + (size_t)s * h->nc * slabsz;
        // read-modify-write also when storing: padding entries keep their device values
        CUDA_TRY(h, cudaMemcpyAsync(tmp.data(), dslab, tmp.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        for (int j = 0; j < h->nvl; j++) {
            size_t flat = h->local_flat[j];
            if (flat / slabsz != (size_t)s) continue;
            size_t rem = flat % slabsz;
            double* hj = host + (size_t)j * h->nc;
            if (to_host) for (int c = 0; c < h->nc; c++) hj[c] = tmp[(size_t)c * slabsz + rem];
            else for (int c = 0; c < h->nc; c++) tmp[(size_t)c * slabsz + rem] = hj[c];
        }
        if (!to_host) {
            CUDA_TRY(h, cudaMemcpyAsync(dslab, tmp.data(), tmp.size() * sizeof(double), cudaMemcpyHostToDevice, h->stream));
            CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        }
    }
    return 0;
}

extern "C" int dugks_get_state(dugks_handle_t* h, double* g, double* h_) {
    if (!h) return DUGKS_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    int rc;
    if (g && (rc = state_io(h, h->A.gt, g, true))) return rc;
    if (h_) {
        if (h->hasH) { if ((rc = state_io(h, h->A.ht, h_, true))) return rc; }
        else memset(h_, 0, sizeof(double) * (size_t)h->nvl * h->nc);   // h == 0 exactly when elided
    }
    return 0;
}

extern "C" int dugks_set_state(dugks_handle_t* h, const double* g, const double* h_) {
    if (!h) return DUGKS_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    int rc;
    if (g && (rc = state_io(h, h->A.gt, const_cast<double*>(g), false))) return rc;
    if (h_ && h->hasH && (rc = state_io(h, h->A.ht, const_cast<double*>(h_), false))) return rc;
    return 0;
}

extern "C" int dugks_get_df(dugks_handle_t* h, int32_t cell, double* g, double* h_) {
    if (!h || cell < 0 || cell >= h->nc) return h ? fail(h, DUGKS_ERR_INVALID, "dugks_get_df: bad cell") : DUGKS_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    // gather this rank's slice into a zero-filled global vector on the device, then sum over
    // ranks (an all-gather expressed with the one collective the reducer offers)
    const size_t slabsz = (size_t)h->L * h->Rs;
    for (int pass = 0; pass < 2; pass++) {
        double* dst = pass == 0 ? g : h_;
        if (!dst) continue;
        const double* src = pass == 0 ? h->A.gt : h->A.ht;
        std::vector<double> glob(h->nxi, 0.0);
        if (src) {
            std::vector<double> tmp(slabsz);
            for (int s = 0; s < h->nslab; s++) {
                CUDA_TRY(h, cudaMemcpyAsync(tmp.data(), src + ((size_t)s * h->nc + cell) * slabsz, slabsz * sizeof(double),
                                            cudaMemcpyDeviceToHost, h->stream));
                CUDA_TRY(h, cudaStreamSynchronize(h->stream));
                for (size_t k = 0; k < slabsz; k++) {
                    int gid = h->flat_gid[(size_t)s * slabsz + k];
                    if (gid >= 0) glob[gid] = tmp[k];
                }
            }
        }
        if (h->nranks > 1) {
            double* d = nullptr;
            CUDA_TRY(h, cudaMalloc(&d, sizeof(double) * h->nxi));
            CUDA_TRY(h, cudaMemcpyAsync(d, glob.data(), sizeof(double) * h->nxi, cudaMemcpyHostToDevice, h->stream));
            int rc = do_allreduce(h, d, h->nxi);
            if (!rc) {
                cudaMemcpyAsync(glob.data(), d, sizeof(double) * h->nxi, cudaMemcpyDeviceToHost, h->stream);
                cudaStreamSynchronize(h->stream);
            }
            cudaFree(d);
            if (rc) return rc;
        }
        memcpy(dst, glob.data(), sizeof(double) * h->nxi);
    }
    return 0;
}

// ------------------------------------------------------------------------------
// Exact restart.  The reference's restart is lossy: every distribution-function field
// is NO_READ/NO_WRITE (discreteVelocity.C:74-205), so a restarted run re-initialises gTilde/hTilde to an
// equilibrium of the saved macros (:220-249) and rho_w to 1 (calculatedMaxwellFvPatchField.C:80).  The blob
// carries everything the next evolution() reads: gTilde/hTilde, the boundary-face values (the incoming half of
// "mixed" patches is never recomputed, :556-573), the lagged boundary gradient (:462-468), cell/face/boundary
// macros (q and tau enter the next half step, :393-396; Usurf the next Courant number), the wall constants
// and the convergence monitor's old fields.  Rank-local, in device layout: valid for the same case, rank
// count and rank only (checked).
struct CkHeader {
    uint64_t magic;
    int32_t abi, nc, nbf, nf, nm, hasH, nranks, rank, L, Rs, nslab, Lt;
    uint64_t nflat, steps, total_bytes;
};
#define CK_MAGIC 0x44554732434b5054ull   // "DUG2CKPT"

struct CkPiece { void* dev; size_t bytes; };
static std::vector<CkPiece> ck_pieces(dugks_handle* h) {
    const size_t ncell_dv = (size_t)h->nc * h->nflat, nb_dv = (size_t)h->nbf * h->nflat;
    // the array the NEXT step reads as the lagged gradient (step_impl flips first)
    const bool flip_next = !h->gam_flip;
    double* gam_g = flip_next ? h->gam_b_g : h->gam_a_g;
    double* gam_h = flip_next ? h->gam_b_h : h->gam_a_h;
    std::vector<CkPiece> v;
    auto add = [&](double* p, size_t n) { if (p && n) v.push_back({p, n * sizeof(double)}); };
    add(h->A.gt, ncell_dv);
    if (h->hasH) add(h->A.ht, ncell_dv);
    add(h->A.gsb, nb_dv);
    if (h->hasH) add(h->A.hsb, nb_dv);
    add(gam_g, nb_dv);
    if (h->hasH) add(gam_h, nb_dv);
    add(h->A.cmac, (size_t)h->nc * MAC_N);
    add(h->A.fmac, (size_t)h->nf * MAC_N);
    add(h->A.bmac, (size_t)h->nbf * 5);
    add(h->A.wall_diag, (size_t)h->nbf * 12);
    add(h->wall_cin, (size_t)h->nbf * h->nm);
    add(h->wall_in, (size_t)h->nbf);
    add(h->d_conv_old, (size_t)5 * h->nc);
    return v;
}
static CkHeader ck_header(dugks_handle* h) {
    CkHeader k{};
    k.magic = CK_MAGIC; k.abi = DUGKS_ABI_VERSION; k.nc = h->nc; k.nbf = h->nbf; k.nf = h->nf; k.nm = h->nm;
    k.hasH = h->hasH; k.nranks = h->nranks; k.rank = h->rank; k.L = h->L; k.Rs = h->Rs; k.nslab = h->nslab; k.Lt = h->Lt;
    k.nflat = h->nflat; k.steps = h->steps;
    k.total_bytes = sizeof(CkHeader);
    for (const CkPiece& p : ck_pieces(h)) k.total_bytes += p.bytes;
    return k;
}

extern "C" int dugks_checkpoint_size(dugks_handle_t* h, uint64_t* bytes) {
    if (!h || !bytes) return DUGKS_ERR_INVALID;
    *bytes = ck_header(h).total_bytes;
    return 0;
}

extern "C" int dugks_checkpoint_save(dugks_handle_t* h, void* buf, uint64_t bytes) {
    if (!h || !buf) return DUGKS_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    const CkHeader k = ck_header(h);
    if (bytes < k.total_bytes) return fail(h, DUGKS_ERR_INVALID, "dugks_checkpoint_save: buffer of %llu bytes, %llu needed", (unsigned long long)bytes, (unsigned long long)k.total_bytes);
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    char* out = (char*)buf;
    memcpy(out, &k, sizeof k);
    out += sizeof k;
    for (const CkPiece& p : ck_pieces(h)) {
        CUDA_TRY(h, cudaMemcpy(out, p.dev, p.bytes, cudaMemcpyDeviceToHost));
        out += p.bytes;
    }
    return 0;
}

extern "C" int dugks_checkpoint_load(dugks_handle_t* h, const void* buf, uint64_t bytes) {
    if (!h || !buf) return DUGKS_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (bytes < sizeof(CkHeader)) return fail(h, DUGKS_ERR_INVALID, "dugks_checkpoint_load: truncated blob");
    CkHeader got;
    memcpy(&got, buf, sizeof got);
    const CkHeader want = ck_header(h);
    if (got.magic != CK_MAGIC || got.abi != want.abi) return fail(h, DUGKS_ERR_INVALID, "dugks_checkpoint_load: not a checkpoint of this library version");
    if (got.nc != want.nc || got.nbf != want.nbf || got.nf != want.nf || got.nm != want.nm || got.hasH != want.hasH ||
        got.nranks != want.nranks || got.rank != want.rank || got.L != want.L || got.Rs != want.Rs || got.nslab != want.nslab ||
        got.Lt != want.Lt || got.nflat != want.nflat || got.total_bytes != want.total_bytes)
        return fail(h, DUGKS_ERR_INVALID, "dugks_checkpoint_load: the blob belongs to another case, rank count or rank "
                    "(cells %d/%d, boundary faces %d/%d, ranks %d/%d, rank %d/%d, local DV slots %llu/%llu)", got.nc, want.nc, got.nbf,
                    want.nbf, got.nranks, want.nranks, got.rank, want.rank, (unsigned long long)got.nflat, (unsigned long long)want.nflat);
    if (bytes < got.total_bytes) return fail(h, DUGKS_ERR_INVALID, "dugks_checkpoint_load: truncated blob");
    CUDA_TRY(h, cudaStreamSynchronize(h->stream));
    const char* in = (const char*)buf + sizeof got;
    for (const CkPiece& p : ck_pieces(h)) {
        CUDA_TRY(h, cudaMemcpy(p.dev, in, p.bytes, cudaMemcpyHostToDevice));
        in += p.bytes;
    }
    h->steps = got.steps;
    // the caller's last boundary arrays are unknown now: the next dugks_set_boundary_macros always applies
    h->last_rho_b.clear(); h->last_U_b.clear(); h->last_T_b.clear();
    return 0;
}

extern "C" int dugks_get_stats(dugks_handle_t* h, dugks_stats_t* out) {
    if (!h || !out) return DUGKS_ERR_INVALID;
    out->kernel_launches = h->launches;
    out->steps = h->steps;
    out->device_bytes = h->dev_bytes;
    out->h_elided = h->hasH ? 0 : 1;
    out->n_slabs = h->nslab;
    out->slab_dvs = h->L * h->Rs;
    out->keep_slabs = h->n_keep;
    out->pencil_mode = h->split_axis ? h->pen_mode : 0;
    out->pencil_cells = out->pencil_mode ? h->n_pen : 0;
    return 0;
}

extern "C" int dugks_kernel_timing(dugks_handle_t* h, int enable, int which, double* total_ms, uint64_t* launches) {
    if (!h) return DUGKS_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    if (total_ms || launches) {
        CUDA_TRY(h, cudaStreamSynchronize(h->stream));
        double tot = 0;
        uint64_t n = 0;
        for (auto& e : h->events)
            if (e.which == which) {
                float ms = 0;
                cudaEventElapsedTime(&ms, e.a, e.b);
                tot += ms;
                n++;
            }
        if (total_ms) *total_ms = tot;
        if (launches) *launches = n;
    }
    if (enable == 0 || enable == 2) {   // 0: stop and clear, 2: clear and keep running
        for (auto& e : h->events) h->pool.push_back(e);
        h->events.clear();
    }
    if (enable == 0) h->timing = false;
    if (enable == 1 || enable == 2) h->timing = true;
    return 0;
}

// debug accessor for parity tests: boundary-face values gSurf/hSurf of local DV j (sorted
// global-id order) on all boundary faces, [nvl][nbf]
extern "C" int dugks_get_boundary_df(dugks_handle_t* h, double* g, double* h_) {
    if (!h) return DUGKS_ERR_INVALID;
    CUDA_TRY(h, cudaSetDevice(h->device));
    const size_t slabsz = (size_t)h->L * h->Rs;
    for (int pass = 0; pass < 2; pass++) {
        double* dst = pass == 0 ? g : h_;
        const double* src = pass == 0 ? h->A.gsb : h->A.hsb;
        if (!dst) continue;
        if (!src) { memset(dst, 0, sizeof(double) * (size_t)h->nvl * h->nbf); continue; }
        std::vector<double> tmp;
        int rc = fetch(h, src, (size_t)h->nbf * h->nflat, tmp);
        if (rc) return rc;
        for (int j = 0; j < h->nvl; j++) {
            size_t flat = h->local_flat[j], s = flat / slabsz, rem = flat % slabsz;
            for (int b = 0; b < h->nbf; b++) dst[(size_t)j * h->nbf + b] = tmp[(s * h->nbf + b) * slabsz + rem];
        }
    }
    return 0;
}
