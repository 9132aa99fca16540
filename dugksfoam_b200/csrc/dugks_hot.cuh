// dugks_hot.cuh — second-generation cell kernels (the ones the step spends its time in).
//
// Same mapping as dugks_tma.cuh (a warp advances one cell, a lane owns one velocity row and walks
// the ix points serially, all inputs arrive as bulk-async copies into a per-warp shared-memory
// ring), rebuilt around the instruction count per (cell, velocity) update:
//   * upwind sets are STATIC (they depend on xi and Sf only): create() evaluates the reference's
//     exact tests once (k_build_upwind) and stores, per (slab, cell, lane), eight 13-bit range codes;
//     the step kernels decode them to bit masks — no FP64 dot products or compares per face and DV;
//   * the least-squares gradient is branch-free: grad = G0*v_c + sum_j G'_j * s_j, where s_j is the
//     j-th staged stream (neighbour gBarP or the lagged boundary gradient), G'_j has the boundary
//     1/deltaCoeffs folded in, G0 = -sum of the internal G_j, symmetry planes carry G' = 0;
//   * four velocity points are advanced together (independent FP64 chains, geometry reads amortised);
//   * boundary faces of a cell are finished in the same pass (lagged normal gradient, outgoing face
//     values), so k_bnd_outgoing only remains for the incoming half of far-field patches;
//   * warps are persistent and prefetch the first chunk (and the geometry record) of their NEXT
//     cell while they finish the current one, so the copy engine never idles at cell boundaries;
//   * face equilibria of phase 2 come from per-face coefficient records written by k_face_macros.
#pragma once
#include <type_traits>

#include "dugks_tma.cuh"

#ifndef HOT_WARPS
#define HOT_WARPS 4
#endif
// velocity points per chunk (= per copy and per unrolled group) are a template parameter CI (2 or 4):
// 4 amortises the geometry reads best (phase 1, 2 CTAs/SM at ~240 registers); 2 halves the stages
// and the working set so that the table-heavy phase-2 kernels run 3 CTAs/SM (12 warps)
#define HOT_CI_MAX 4
#ifndef HOT_AXIS_CU
#define HOT_AXIS_CU 2     // points advanced together by hot_axis_item (its staged chunk may hold more)
#endif
#define HOT_STAGES 2
#define HOT_PAD (HOT_CI_MAX * 32)   // doubles of slack behind every streamed array (tail chunks copy full size)
#define FCOEF_N 12        // per-face equilibrium record: Ux Uy Uz a pre qx qy qz omrf RT 0 0

// ---- upwind range codes -----------------------------------------------------------------------
// bits 0-5: a, bits 6-11: b, bit 12: dir.   dir = 0: [0,a) none, [a,b) tie, [b,L) full
//                                           dir = 1: [0,a) full, [a,b) tie, [b,L) none
__host__ __device__ __forceinline__ unsigned hot_lowmask(unsigned n) {
#ifdef __CUDA_ARCH__
    return __funnelshift_lc(0xffffffffu, 0u, n);   // low n bits set, n <= 32 (one SHF)
#else
    return n >= 32u ? 0xffffffffu : ((1u << n) - 1u);
#endif
}
__host__ __device__ __forceinline__ void hot_decode(unsigned code, int L, unsigned& full, unsigned& tie) {
    const unsigned la = hot_lowmask(code & 63u), lb = hot_lowmask((code >> 6) & 63u);
    tie = lb & ~la;
    full = (code & 0x1000u) ? la : (hot_lowmask((unsigned)L) & ~lb);
}
__host__ __device__ __forceinline__ bool hot_encode(unsigned full, unsigned tie, int L, unsigned& code) {
    // try both directions; the masks must be (prefix | run | suffix) shaped
    const int nfull = __builtin_popcount(full), ntie = __builtin_popcount(tie);
    for (unsigned dir = 0; dir < 2; dir++) {
        unsigned a, b;
        if (dir == 0) { b = (unsigned)(L - nfull); a = b - (unsigned)ntie; }
        else { a = (unsigned)nfull; b = a + (unsigned)ntie; }
        unsigned c = a | (b << 6) | (dir << 12), f2, t2;
        hot_decode(c, L, f2, t2);
        if (f2 == full && t2 == tie) { code = c; return true; }
    }
    return false;
}

// One thread per (slab, cell, lane): evaluates xi.Sf with the reference's operation order for every
// face entry and every ix, classifies (discreteVelocity.C:495,506 internal; :562,580,614,675 boundary)
// and stores the range codes.  bad[0] is raised if a set is not range shaped (unsorted abscissae),
// bad[1 + slab] if a row of the slab has a tie on a y/z face of an axis-aligned cell (hot_axis_item
// cannot take that slab).
__global__ void k_build_upwind(StepArgs a, uint4* out, int* bad) {
    const DevDV& dv = a.dv;
    const long long total = (long long)dv.nslab * a.m.nc * 32;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int lane = (int)(t & 31);
        const long long sc = t >> 5;
        const int item = (int)(sc % a.m.nc), slab = (int)(sc / a.m.nc);
        const int c = a.cmeta[(size_t)item * 24 + 20];   // codes are stored in traversal order
        const int grow = slab * 32 + lane;
        const double y = dv.row_y[grow], z = dv.row_z[grow];
        const int cb = dv.row_cbase[grow];
        const int e0 = a.m.cell_off[c], ne = a.m.cell_off[c + 1] - e0;
        unsigned short codes[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const bool axis_cell = ((a.cmeta[(size_t)item * 24 + 1] >> 16) & 0xff) != 0;
        for (int j = 0; j < ne && j < 8; j++) {
            const double* g = a.m.e_geo + (size_t)(e0 + j) * 9;
            const int o = a.m.e_other[e0 + j];
            const bool isown = a.m.e_owner[e0 + j] != 0;
            unsigned full = 0, tie = 0;
            for (int i = 0; i < dv_len(dv, slab); i++) {
                const double phi = dot_exact(dv.tx[cb + i], y, z, g[6], g[7], g[8]);
                if (o >= 0) {
                    const bool neg = phi < -DUGKS_VSMALL, pos = phi >= DUGKS_VSMALL;
                    const bool f = isown ? pos : neg, none = isown ? neg : pos;
                    if (f) full |= 1u << i;
                    else if (!none) tie |= 1u << i;
                } else {
                    const int kind = a.m.b_kind[-1 - o];
                    bool outgoing;
                    if (kind == K_ZERO_GRADIENT) outgoing = true;
                    else if (kind == K_DVM_SYMMETRY || kind == K_SYMMETRY_PLANE) outgoing = phi > -DUGKS_VSMALL;
                    else outgoing = phi > 0;
                    if (outgoing) full |= 1u << i;
                }
            }
            unsigned code = 0;
            if (!hot_encode(full, tie, dv_len(dv, slab), code)) atomicOr(bad, 1);
            if (axis_cell && j >= 2 && tie != 0) atomicOr(bad + 1 + slab, 1);
            codes[j] = (unsigned short)code;
        }
        uint4 w;
        w.x = codes[0] | ((unsigned)codes[1] << 16);
        w.y = codes[2] | ((unsigned)codes[3] << 16);
        w.z = codes[4] | ((unsigned)codes[5] << 16);
        w.w = codes[6] | ((unsigned)codes[7] << 16);
        out[t] = w;
    }
}

// ---- asynchronous staging (LDGSTS: 16 bytes per lane and instruction, L2 -> shared, bypassing L1) ----
// Seven-plus streams of 1 KB per chunk are cheaper to launch as 2 warp-wide cp.async per stream (one
// address add each) than as one bulk copy per stream issued by a single elected lane (uniform-register
// set-up + mbarrier bookkeeping per copy), see profiles/r01_ncu_summary_r2a.txt.
__device__ __forceinline__ void cp_async16(uint32_t sdst, const void* gsrc) {
    // no "memory" clobber: the stage is only read after cp_async_wait + __syncwarp, and refilled after
    // the __syncwarp that ends the chunk which read it
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sdst), "l"(gsrc));
}
// single-use data (state rows read once per pass, upwind codes): L2 evict-first, so that the row blocks
// every neighbour re-reads stay resident
__device__ __forceinline__ unsigned long long l2_evict_first_policy() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ unsigned long long l2_evict_last_policy() {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__global__ void k_make_policies(unsigned long long* out) {
    out[0] = l2_evict_first_policy();
    out[1] = l2_evict_last_policy();
}
__device__ __forceinline__ void cp_async16_ef(uint32_t sdst, const void* gsrc, unsigned long long pol) {
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(sdst), "l"(gsrc), "l"(pol));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Loads that must be ISSUED where they are written (the next cell's record, long before its use): as
// plain C++ the compiler sinks them to their first use to save registers, which puts the full memory
// latency back on the critical path (profiles/r01_ncu_summary_r3c.txt).  Read-only data only (.nc).
__device__ __forceinline__ int ldg_early(const int* p) {
    int v;
    asm volatile("ld.global.nc.b32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int2 ldg_early2(const int* p) {
    int2 v;
    asm volatile("ld.global.nc.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ int ldg_early_u8(const unsigned char* p) {
    int v;
    asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 ldg_early4(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
}

// ---- per-warp item bookkeeping ------------------------------------------------------------------
struct HotMeta {
    int c, e0, ne, nint, cls;
    int other, face, own, kind;   // lane j < ne: entry j (kind: -1 internal, else patch kind)
    uint4 mw;                     // upwind range codes of this (cell, lane)
};

// CTAs per SM a kernel is compiled for (register budget 65536 / (128 * n)): 2 with CI = 4, 3 with CI = 2
#define HOT_MINB(CI) ((CI) == 4 ? 2 : 3)
#define HOT_PTRS 20   // stream pointer table entries per buffer (>= 2 * (2 + 8))

// Per-cell record `cmeta` (24 ints, built by create()): [0] first entry e0, [1] ne | nint << 8 | cls << 16,
// [2..9] other cell (or -1-b) of entry j, [10..17] face id | owner << 31, [18..19] patch kind bytes
// (0xff = internal), [20] the cell.  One dependency level instead of cell_off -> e_other -> b_kind.
// Records (and the upwind codes) are stored in TRAVERSAL order: warps walk the mesh along a space-
// filling curve of the cell centres so that all neighbours of a cell are visited close in time and
// their row blocks are still in L2 (natural order: +-z neighbours of a 64^3 mesh are 29 MB apart).
#define CMETA_N 24

// Issues the loads of cell c's record and upwind codes; nothing is consumed here, so the latency hides
// behind the cell the warp is working on.
__device__ __forceinline__ void hot_meta_issue(const StepArgs& a, int item, int lane, HotMeta& M) {
    const int* rec = a.cmeta + (size_t)item * CMETA_N;
    const int l8 = lane & 7;
    M.c = ldg_early(rec + 20);
    const int2 h2 = ldg_early2(rec);
    M.e0 = h2.x;
    M.ne = h2.y;                   // packed, unpacked by hot_meta_commit
    M.other = ldg_early(rec + 2 + l8);
    M.face = ldg_early(rec + 10 + l8);         // | owner << 31
    M.kind = ldg_early_u8(reinterpret_cast<const unsigned char*>(rec + 18) + l8);
    M.mw = ldg_early4(a.upw + ((size_t)a.slab * a.m.nc + item) * 32 + lane);
}

// Unpacks the record and writes the stream sources of phase-1/2 into the pointer table `sp`
// (slot 0 = own cell, 1+j = entry j; + NSLOT for the h field).
template <bool HAS_H>
__device__ __forceinline__ void hot_meta_commit(const StepArgs& a, int lane, const double* gbs, const double* hbs,
                                                const double* gam_g, const double* gam_h, int NE,
                                                unsigned long long* sp, HotMeta& M) {
    const int blk = a.dv.L * 32;
    const int NSLOT = 1 + NE;
    const int pk = M.ne;
    M.ne = pk & 0xff; M.nint = (pk >> 8) & 0xff; M.cls = (pk >> 16) & 0xff;
    const bool valid = lane < M.ne && M.ne <= NE;
    M.own = valid ? (int)((unsigned)M.face >> 31) : 0;
    M.face = valid ? (M.face & 0x7fffffff) : 0;
    M.kind = (valid && M.other < 0) ? M.kind : -1;
    if (!valid) M.other = 0;
    if (M.ne <= NE) {
        // lane j+1 <- entry j
        const int o = __shfl_up_sync(0xffffffffu, M.other, 1), k = __shfl_up_sync(0xffffffffu, M.kind, 1);
        if (lane <= M.ne) {
#pragma unroll
            for (int fld = 0; fld < (HAS_H ? 2 : 1); fld++) {
                const double* cellsrc = fld ? hbs : gbs;
                const double* src;
                if (lane == 0) src = cellsrc + (size_t)M.c * blk;
                else if (o >= 0) src = cellsrc + (size_t)o * blk;
                else if (k == K_SYMMETRY_PLANE) src = cellsrc + (size_t)M.c * blk;   // G' = 0: any finite data
                else src = (fld ? gam_h : gam_g) + (size_t)(-1 - o) * blk;
                sp[fld * NSLOT + lane] = (unsigned long long)src;
            }
        }
    }
    __syncwarp();
}

// chunk `ch` of the item whose pointers are in `sp` into `stage`; one commit group (all lanes call)
template <int CI, int NTOT, int NSLOT, bool ALL>
__device__ __forceinline__ void hot_stage(const unsigned long long* sp, int ne, int ch, double* stage, int lane) {
    const uint32_t loff = (uint32_t)lane * 16u + (uint32_t)ch * (CI * 256u);
    const uint32_t sdst = smem_u32(stage) + (uint32_t)lane * 16u;
#pragma unroll
    for (int k = 0; k < NTOT; k++) {
        if (ALL || (k % NSLOT) <= ne) {
            const char* p = reinterpret_cast<const char*>(sp[k]) + loff;
#pragma unroll
            for (int part = 0; part < CI * 256 / 512; part++)
                cp_async16(sdst + k * (CI * 256) + part * 512, p + part * 512);
        }
    }
}

// one stream block of a chunk from `base` + off16 * 16 bytes (off16 already holds the lane's 16 bytes)
template <int CI>
__device__ __forceinline__ void hot_stage_one(uint32_t sdst, const double* base, uint32_t off16, uint32_t coff) {
    const char* p = reinterpret_cast<const char*>(base) + (size_t)off16 * 16u + coff;
#pragma unroll
    for (int part = 0; part < CI * 256 / 512; part++) cp_async16(sdst + part * 512, p + part * 512);
}
template <int CI>
__device__ __forceinline__ void hot_stage_one_ef(uint32_t sdst, const double* base, uint32_t off16, uint32_t coff,
                                                 unsigned long long pol) {
    const char* p = reinterpret_cast<const char*>(base) + (size_t)off16 * 16u + coff;
#pragma unroll
    for (int part = 0; part < CI * 256 / 512; part++) cp_async16_ef(sdst + part * 512, p + part * 512, pol);
}

// one stream block of a chunk from the byte address `p` (the lane's 16 bytes included)
#ifdef DUGKS_RLX_NOHINT
#define RLX_HINT false
#else
#define RLX_HINT true
#endif
#ifndef RLX_L2_AHEAD
#define RLX_L2_AHEAD 0   // 1: also pull the chunk after the one being staged into L2 (prefetch.global.L2)
#endif
template <int CI>
__device__ __forceinline__ void hot_stage_at(uint32_t sdst, unsigned long long p, unsigned long long pol, bool ahead = false) {
#pragma unroll
    for (int part = 0; part < CI * 256 / 512; part++) {
        if (RLX_HINT) cp_async16_ef(sdst + part * 512, reinterpret_cast<const char*>(p) + part * 512, pol);
        else cp_async16(sdst + part * 512, reinterpret_cast<const char*>(p) + part * 512);
    }
    if (RLX_L2_AHEAD && ahead) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + CI * 256));
}
// L2 hints of the relax+update kernel come from the kernel arguments (uniform registers; a policy made with
// createpolicy inside the kernel lives in a vector register and costs two R2UR per LDGSTS)
#define RLX_POL(x) (x)

// interior cells: every stream lives in the same slab array, so 32-bit offsets (16-byte units, already
// including the lane's 16 bytes) in registers replace the pointer table: one 64-bit multiply-add + CI/2
// LDGSTS per stream
// EVL: mark the lines evict-last in L2 (row blocks that six neighbours read again while a stream of
// single-use stores passes through the same cache)
template <int CI, int NFLD, int NSLOT, bool EVL = false>
__device__ __forceinline__ void hot_stage_off(const double* gbs, const double* hbs, const uint32_t (&soff)[NSLOT],
                                              int ch, double* stage, int lane, unsigned long long pol = 0) {
    const uint32_t sdst = smem_u32(stage) + (uint32_t)lane * 16u;
    const uint32_t coff = (uint32_t)ch * (CI * 256u);
#pragma unroll
    for (int fld = 0; fld < NFLD; fld++) {
        // array base + chunk offset once per chunk (kept opaque: otherwise the compiler folds the chunk
        // offset into every stream's 32-bit product and adds the 64-bit base per stream, 3 instructions each)
        unsigned long long cbase = (unsigned long long)(fld ? hbs : gbs) + coff;
        asm volatile("" : "+l"(cbase));
#pragma unroll
        for (int k = 0; k < NSLOT; k++) {
            const char* p = reinterpret_cast<const char*>(cbase + (unsigned long long)soff[k] * 16u);
#pragma unroll
            for (int part = 0; part < CI * 256 / 512; part++) {
                if (EVL) cp_async16_ef(sdst + (fld * NSLOT + k) * (CI * 256) + part * 512, p + part * 512, pol);
                else cp_async16(sdst + (fld * NSLOT + k) * (CI * 256) + part * 512, p + part * 512);
            }
        }
    }
}

// geometry record of a cell ((1 + ne) * 48 bytes) into `dst`, same commit group as the chunk that follows
__device__ __forceinline__ void hot_stage_geo(const double* src, int ne, double* dst, int lane) {
    if (lane < (1 + ne) * 3) cp_async16(smem_u32(dst) + lane * 16, reinterpret_cast<const char*>(src) + lane * 16);
}

// shared-memory plan of one warp
template <int PHASE, bool HAS_H, int NE, int TW, int CI>
struct HotPlan {
    static constexpr int NFLD = HAS_H ? 2 : 1;
    static constexpr int NSLOT = 1 + NE;
    static constexpr int STAGE_D = NFLD * NSLOT * CI * 32;
    static constexpr int GEO_D = NSLOT * 6;
    // PHASE 1 reduces its moments through the stage that was consumed last when that is large enough
    // (32 x 17 doubles for one face at a time, 32 x (2 NV + 1) for hot_reduce_faces2)
    static constexpr int RED_D = 32 * (2 * (HAS_H ? 17 : 13) + 1);
    static constexpr bool RED_IN_STAGE = STAGE_D >= RED_D;
    static constexpr int FREC_D = NE * FCOEF_N;   // PHASE 2: face equilibrium records of a cell (double-buffered)
    static constexpr int EXTRA_D = (PHASE == 1) ? (RED_IN_STAGE ? 0 : RED_D) : (NE * 4 * TW + NE * 2 + 2 * FREC_D);
    static constexpr int PER_WARP_D = 2 * HOT_PTRS + 2 * GEO_D + HOT_STAGES * STAGE_D + EXTRA_D;
    static constexpr size_t PER_WARP = ((size_t)PER_WARP_D * 8 + 127) / 128 * 128;
    static __host__ __device__ size_t txs_bytes(int ntab) { return ((size_t)(ntab + HOT_CI_MAX) * 48 + 127) / 128 * 128; }
    static __host__ size_t total(int ntab) { return txs_bytes(ntab) + HOT_WARPS * PER_WARP; }
};

// txs[t] = { -0.5 dt x, W0, W1, W2, W3, x }, zero weights past the table
__device__ __forceinline__ void hot_fill_txs(const DevDV& dv, double hd, double* txs) {
    for (int k = threadIdx.x; k < dv.ntab + HOT_CI_MAX; k += blockDim.x) {
        const int kk = min(k, dv.ntab - 1);
        const bool real = k < dv.ntab;
        txs[k * 6 + 0] = hd * dv.tx[kk];
        txs[k * 6 + 1] = real ? dv.tx[dv.ntab + kk] : 0.0;
        txs[k * 6 + 2] = real ? dv.tx[2 * dv.ntab + kk] : 0.0;
        txs[k * 6 + 3] = real ? dv.tx[3 * dv.ntab + kk] : 0.0;
        txs[k * 6 + 4] = real ? dv.tx[4 * dv.ntab + kk] : 0.0;
        txs[k * 6 + 5] = dv.tx[kk];
    }
}

// bit 4*ch of the result: some / every point of chunk ch is set in m
template <int CI>
__device__ __forceinline__ unsigned hot_spread_any(unsigned m) {
    static_assert(CI == 2 || CI == 4, "CI must be 2 or 4");
    return CI == 4 ? ((m | (m >> 1) | (m >> 2) | (m >> 3)) & 0x11111111u) : ((m | (m >> 1)) & 0x55555555u);
}
template <int CI>
__device__ __forceinline__ unsigned hot_spread_all(unsigned m) {
    return CI == 4 ? ((m & (m >> 1) & (m >> 2) & (m >> 3)) & 0x11111111u) : ((m & (m >> 1)) & 0x55555555u);
}

// per-warp constants of the outgoing kernel
struct HotCtx {
    const double* txs;
    const double* geo;     // geometry record of the current cell
    double* red;           // PHASE 1 reduction scratch
    double* xtab;          // PHASE 2 [NE][TW][4]
    double* unic;          // PHASE 2 [NE][2]
    const double* frec;    // PHASE 2 face equilibrium records of the current cell [NE][FCOEF_N]
    double y, z, wr, yh, zh, kd;
    int cb, tmin, span, lane, L, blk, nm, nchunk;
    size_t slab_b;
};

// One cell of the outgoing kernel.  INTERIOR: every one of the NE entries is an internal face, so all
// face predicates are compile-time.  `prefetch(ch)` stages the chunk that follows chunk ch.
// AXIS (implies INTERIOR): entries come in axis order with a single non-zero component each, so the
// gradient is 2 products per axis and a face value 1 (exact zeros skipped: bit-identical results).
template <int PHASE, bool HAS_H, int NE, int TW, int CI, bool INTERIOR, bool AXIS, class Prefetch>
__device__ __forceinline__ void hot_out_item(const StepArgs& a, const HotCtx& x, const HotMeta& cur,
                                             double* stages, int stage_d, uint32_t& q, Prefetch&& prefetch) {
    constexpr int NSLOT = 1 + NE, NFLD = HAS_H ? 2 : 1;
    const int lane = x.lane, L = x.L, blk = x.blk;
    // INTERIOR: all NE entries exist (compile-time stream count); AXIS additionally: all of them internal
    const int ne = INTERIOR ? NE : cur.ne, nint = AXIS ? NE : cur.nint;
    const double* gb_ = x.geo;
    const double* txs = x.txs;
    // ---- upwind masks of this (cell, lane) and their warp-uniform per-chunk summaries
    unsigned full[NE], tie[NE], anyc[NE], allc[NE];
    const unsigned ownmask = __ballot_sync(0xffffffffu, cur.own != 0);
    int fbase[NE];
    {
        const unsigned w4[4] = {cur.mw.x, cur.mw.y, cur.mw.z, cur.mw.w};
#pragma unroll
        for (int j = 0; j < NE; j++) {
            full[j] = 0; tie[j] = 0; anyc[j] = 0; allc[j] = 0;
            fbase[j] = __shfl_sync(0xffffffffu, cur.face, j);
            if (j < ne) {
                hot_decode((w4[j >> 1] >> ((j & 1) * 16)) & 0xffffu, L, full[j], tie[j]);
                if (PHASE == 2 && j < nint) {
                    // exactly one side writes the face value: the owner unless phi < -VSMALL (ties: the
                    // owner; their flux is zero).  From here on full[] holds the write mask.
                    if ((ownmask >> j) & 1u) full[j] |= tie[j];
                    tie[j] = 0;
                }
                if (j < nint) {   // boundary entries keep their out-set in full[] but take no part in the face loop
                    anyc[j] = __reduce_or_sync(0xffffffffu, hot_spread_any<CI>(full[j] | tie[j]));
                    allc[j] = __reduce_and_sync(0xffffffffu, hot_spread_all<CI>(full[j]));
                }
            }
        }
    }
    double* const fkeep_g = (PHASE == 1 && a.fkeep_g) ? a.fkeep_g + (size_t)a.slab * a.m.nif * blk : nullptr;
    double* const fkeep_h = (PHASE == 1 && HAS_H && a.fkeep_h) ? a.fkeep_h + (size_t)a.slab * a.m.nif * blk : nullptr;
    if (PHASE == 2) {   // the face records were staged with this cell's first chunk (the only group in flight)
        cp_async_wait<0>();
        __syncwarp();
    }
    // ---- accumulators (PHASE 1) / face equilibrium tables (PHASE 2)
    double accg[NE][4], acch[NE][2];
    double EYZ[NE], YZ2[NE], QYZ[NE];
#pragma unroll
    for (int j = 0; j < NE; j++) {
        accg[j][0] = accg[j][1] = accg[j][2] = accg[j][3] = 0.0;
        acch[j][0] = acch[j][1] = 0.0;
        EYZ[j] = YZ2[j] = QYZ[j] = 0.0;
        if (PHASE == 2 && j < nint && anyc[j] != 0) {
            const double* fc = x.frec + j * FCOEF_N;
            const double Ux = fc[0], Uy = fc[1], Uz = fc[2], ia = fc[3], pre = fc[4];
            const double qx = fc[5], qy = fc[6], qz = fc[7];
            for (int tt = lane; tt < x.span; tt += 32) {
                const double cx = txs[(x.tmin + tt) * 6 + 5] - Ux;
                const double x2 = cx * cx * ia;
                double* xt = x.xtab + ((size_t)j * TW + tt) * 4;
                xt[0] = exp(-0.5 * x2); xt[1] = x2; xt[2] = cx * qx; xt[3] = 0.0;
            }
            const double cy = x.y - Uy, cz = x.z - Uz;
            const double yz2 = (cy * cy + cz * cz) * ia;
            EYZ[j] = pre * exp(-0.5 * yz2);
            YZ2[j] = yz2 - a.gas.D - 2.0;
            QYZ[j] = cy * qy + cz * qz;
            if (lane == 0) { x.unic[j * 2] = fc[8]; x.unic[j * 2 + 1] = fc[9]; }
        }
    }
    if (PHASE == 2) __syncwarp();

    for (int ch = 0; ch < x.nchunk; ch++) {
        prefetch(ch);
        cp_async_commit();
        cp_async_wait<1>();
        __syncwarp();
        const double* sg = stages + (q & 1) * stage_d;
        const int i0 = ch * CI;
        const int tb = x.cb + i0;
#pragma unroll
        for (int fld = 0; fld < NFLD; fld++) {
            const double* sf = sg + fld * NSLOT * CI * 32 + lane;
            // ---- gradient (stock leastSquaresGrad, zeroBoundaryGrad.C:90-99,126-133)
            double v[CI], gx[CI], gy[CI], gz[CI], base[CI];
            {
                const double2 G01 = lds2(gb_);
                const double G2 = gb_[2];
#pragma unroll
                for (int u = 0; u < CI; u++) {
                    v[u] = sf[u * 32];
                    gx[u] = G01.x * v[u]; gy[u] = G01.y * v[u]; gz[u] = G2 * v[u];
                }
            }
            double rax[NE];   // AXIS: the one non-zero component of Cf - C per face
#pragma unroll
            for (int j = 0; j < NE; j++) {
                rax[j] = 0.0;
                if (AXIS) {
                    const int d = j >> 1;
                    const double G = gb_[6 * (1 + j) + d];
                    rax[j] = gb_[6 * (1 + j) + 3 + d];
#pragma unroll
                    for (int u = 0; u < CI; u++) {
                        const double sv = sf[((1 + j) * CI + u) * 32];
                        if (d == 0) gx[u] = fma(G, sv, gx[u]);
                        else if (d == 1) gy[u] = fma(G, sv, gy[u]);
                        else gz[u] = fma(G, sv, gz[u]);
                    }
                } else if (j < ne) {
                    const double2 G01 = lds2(gb_ + 6 * (1 + j));
                    const double G2 = gb_[6 * (1 + j) + 2];
#pragma unroll
                    for (int u = 0; u < CI; u++) {
                        const double sv = sf[((1 + j) * CI + u) * 32];
                        gx[u] = fma(G01.x, sv, gx[u]); gy[u] = fma(G01.y, sv, gy[u]); gz[u] = fma(G2, sv, gz[u]);
                    }
                }
            }
            // ---- value at the cell centre moved back by half a step: v - 0.5 dt xi.grad (:498-502)
            double W[CI][4];
#pragma unroll
            for (int u = 0; u < CI; u++) {
                const double2 t0 = lds2(txs + (tb + u) * 6);
                base[u] = (AXIS && NE == 4) ? fma(t0.x, gx[u], fma(x.yh, gy[u], v[u]))
                                            : fma(t0.x, gx[u], fma(x.yh, gy[u], fma(x.zh, gz[u], v[u])));
                if (PHASE == 1) {
                    const double2 t1 = lds2(txs + (tb + u) * 6 + 2);
                    W[u][0] = t0.y; W[u][1] = t1.x; W[u][2] = t1.y; W[u][3] = txs[(tb + u) * 6 + 4];
                }
            }
            // ---- internal faces
#pragma unroll
            for (int j = 0; j < NE; j++) {
                if (j >= nint) continue;
                if (!((anyc[j] >> i0) & 1u)) continue;                      // warp-uniform
                double r0 = 0.0, r1 = 0.0, r2 = 0.0;
                if (!AXIS) {
                    const double2 Gr = lds2(gb_ + 6 * (1 + j) + 2), r12 = lds2(gb_ + 6 * (1 + j) + 4);
                    r0 = Gr.y; r1 = r12.x; r2 = r12.y;
                }
                // value at the face centre: base + (Cf - C).grad
                auto face_val = [&](int u) {
                    if (AXIS) return fma(rax[j], (j >> 1) == 0 ? gx[u] : ((j >> 1) == 1 ? gy[u] : gz[u]), base[u]);
                    return fma(r0, gx[u], fma(r1, gy[u], fma(r2, gz[u], base[u])));
                };
                const bool allfull = (allc[j] >> i0) & 1u;                  // warp-uniform
                if (PHASE == 1) {
                    // slabs in face-storage mode keep the reconstructed value for the fused relax+update
                    // kernel: the owner writes unless phi < -VSMALL, then the neighbour does
                    double* keep = (fld == 0 ? fkeep_g : fkeep_h);
                    if (keep) keep += (size_t)fbase[j] * blk + i0 * 32 + lane;
                    if (allfull) {
#pragma unroll
                        for (int u = 0; u < CI; u++) {
                            const double val = face_val(u);
                            if (keep) __stcs(keep + u * 32, val);
                            if (fld == 0) {
                                accg[j][0] = fma(W[u][0], val, accg[j][0]); accg[j][1] = fma(W[u][1], val, accg[j][1]);
                                accg[j][2] = fma(W[u][2], val, accg[j][2]); accg[j][3] = fma(W[u][3], val, accg[j][3]);
                            } else {
                                acch[j][0] = fma(W[u][0], val, acch[j][0]); acch[j][1] = fma(W[u][1], val, acch[j][1]);
                            }
                        }
                    } else {
                        const unsigned fb = full[j] >> i0, tbits = tie[j] >> i0;
                        const unsigned wbk = ((ownmask >> j) & 1u) ? (fb | tbits) : fb;
#pragma unroll
                        for (int u = 0; u < CI; u++) {
                            double val = face_val(u);
                            if (keep && ((wbk >> u) & 1u)) __stcs(keep + u * 32, val);
                            // this side's share: all of it, half of it on a tie (:513-529), or none
                            const int hi = ((fb >> u) & 1u) ? 0x3ff00000 : (((tbits >> u) & 1u) ? 0x3fe00000 : 0);
                            val *= __hiloint2double(hi, 0);
                            if (fld == 0) {
                                accg[j][0] = fma(W[u][0], val, accg[j][0]); accg[j][1] = fma(W[u][1], val, accg[j][1]);
                                accg[j][2] = fma(W[u][2], val, accg[j][2]); accg[j][3] = fma(W[u][3], val, accg[j][3]);
                            } else {
                                acch[j][0] = fma(W[u][0], val, acch[j][0]); acch[j][1] = fma(W[u][1], val, acch[j][1]);
                            }
                        }
                    }
                } else {
                    const double omrf = x.unic[j * 2], frt = x.unic[j * 2 + 1];
                    double* dst = (fld == 0 ? a.fbuf_g : a.fbuf_h) + (size_t)fbase[j] * blk + i0 * 32 + lane;
                    const unsigned wb = full[j] >> i0;
#pragma unroll
                    for (int u = 0; u < CI; u++) {
                        const double val = face_val(u);
                        const double* xt = x.xtab + ((size_t)j * TW + (tb + u - x.tmin)) * 4;
                        const double2 x01 = lds2(xt);
                        const double cc = x01.y + YZ2[j];
                        const double cq = xt[2] + QYZ[j];
                        const double gM = x01.x * EYZ[j];
                        double eq;
                        if (fld == 0) eq = fma(cq, cc, 1.0) * gM;                                      // :1042
                        else eq = (x.kd + cq * ((cc + 2.0) * x.kd - 2.0 * a.gas.K)) * gM * frt;        // :1043
                        if (allfull || ((wb >> u) & 1u)) __stcs(dst + u * 32, fma(omrf, val, eq));             // :880-881
                    }
                }
            }
            // ---- boundary faces of this cell (PHASE 1): lagged normal gradient (:436-470) and the
            // outgoing half of the patch rules (:533-690)
            // gam_late (one in-place copy of the lagged gradient): a slab that recomputes its gradient in phase 2 still
            // needs the old values then, so it writes the new ones there (same gx, gy, gz: same bits)
            const bool gam_here = PHASE == 1 ? (!a.gam_late || fkeep_g != nullptr) : a.gam_late != 0;
            if (!AXIS && (PHASE == 1 || gam_here) && nint < ne) {
                for (int j = nint; j < ne; j++) {
                    const int kind = __shfl_sync(0xffffffffu, cur.kind, j);
                    const int b = -1 - __shfl_sync(0xffffffffu, cur.other, j);
                    const double r0 = gb_[6 * (1 + j) + 3], r1 = gb_[6 * (1 + j) + 4], r2 = gb_[6 * (1 + j) + 5];
                    const size_t bo = x.slab_b + (size_t)b * blk + i0 * 32 + lane;
                    double* gam_new = fld == 0 ? a.gam_new_g : a.gam_new_h;
                    double* gsb = fld == 0 ? a.gsb : a.hsb;
                    unsigned fj = 0;
#pragma unroll
                    for (int jj = 0; jj < NE; jj++) if (jj == j) fj = full[jj];
                    const unsigned ob = fj >> i0;
                    if (gam_here && kind != K_SYMMETRY_PLANE) {
                        const double n0 = a.m.b_n[(size_t)b * 3], n1 = a.m.b_n[(size_t)b * 3 + 1], n2 = a.m.b_n[(size_t)b * 3 + 2];
#pragma unroll
                        for (int u = 0; u < CI; u++)
                            if (i0 + u < L) gam_new[bo + u * 32] = gx[u] * n0 + gy[u] * n1 + gz[u] * n2;
                    }
                    if (PHASE != 1) continue;
#pragma unroll
                    for (int u = 0; u < CI; u++) {
                        if ((ob >> u) & 1u) {
                            const double val = (kind == K_ZERO_GRADIENT)
                                                   ? v[u]
                                                   : fma(r0, gx[u], fma(r1, gy[u], fma(r2, gz[u], base[u])));
                            gsb[bo + u * 32] = val;
                        }
                    }
                }
            }
        }
        __syncwarp();   // every lane is done with this stage before it is refilled
        q++;
    }

    if (PHASE == 1) {
#pragma unroll
        for (int j = 0; j < NE; j++) {
            if (j < nint && anyc[j] != 0) {   // warp-uniform; a side that is nowhere upwind adds nothing
                double vv[16];
                expand_g(accg[j], x.wr, x.y, x.z, vv);
                double uu[NM_H] = {0, 0, 0, 0};
                if (HAS_H) expand_h(acch[j], x.wr, x.y, x.z, uu);
                vv[13] = uu[0]; vv[14] = uu[1]; vv[15] = uu[2];
                const double tot = warp_reduce16_smem(vv, stage_d >= 32 * 17 ? stages + ((q & 1) ^ 1) * stage_d : x.red, lane);
                const size_t slot = (size_t)2 * fbase[j] + (((ownmask >> j) & 1u) ? 0 : 1);
                // one warp owns a slot per launch: the fire-and-forget add keeps the sum deterministic
                if (lane < 16 && lane < x.nm) atomicAdd(a.fslot + slot * x.nm + lane, tot);
                if (HAS_H) {
                    const double t3 = warp_sum(uu[3]);
                    if (lane == 0) atomicAdd(a.fslot + slot * x.nm + 16, t3);
                }
            }
        }
    }
}

// Moments of two faces in one pass through shared memory.  Every lane brings raw accumulators for face A
// and face B (zeros where it takes no part); 2 x NV columns, row stride 2 NV + 1 doubles (odd: no bank
// conflicts), every lane sums half a column (16 rows) per round of 16 columns.  red: 32 x (2 NV + 1) doubles.
template <bool HAS_H>
__device__ __forceinline__ void hot_reduce_faces2(const StepArgs& a, const HotCtx& x, const double (&mA)[4],
                                                  const double (&mB)[4], const double (&hA)[2], const double (&hB)[2],
                                                  size_t slotA, size_t slotB, bool onA, bool onB, double* red, int lane) {
    constexpr int NV = HAS_H ? 17 : 13, NCOL = 2 * NV, RS = NCOL + 1;
    double* row = red + lane * RS;
#pragma unroll
    for (int f = 0; f < 2; f++) {
        double vv[NM_G];
        expand_g(f ? mB : mA, x.wr, x.y, x.z, vv);
#pragma unroll
        for (int k = 0; k < 13; k++) row[f * NV + k] = vv[k];
        if (HAS_H) {
            double uu[NM_H];
            expand_h(f ? hB : hA, x.wr, x.y, x.z, uu);
#pragma unroll
            for (int k = 0; k < 4; k++) row[f * NV + 13 + k] = uu[k];
        }
    }
    __syncwarp();
    const int col = lane & 15, half = lane >> 4;
#pragma unroll
    for (int c0 = 0; c0 < NCOL; c0 += 16) {
        const int c = c0 + col;
        const bool live = c < NCOL;                    // compile-time true except in the last round
        const double* src = red + (half * 16) * RS + (live ? c : 0);
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int rr = 0; rr < 16; rr += 2) {
            s0 += src[rr * RS];
            s1 += src[(rr + 1) * RS];
        }
        double t = s0 + s1;
        t += __shfl_xor_sync(0xffffffffu, t, 16);
        // one warp owns a slot per launch: the fire-and-forget add keeps the sum deterministic
        const bool second = c >= NV;
        const size_t slot = second ? slotB : slotA;
        if (half == 0 && live && (second ? onB : onA)) atomicAdd(a.fslot + slot * x.nm + (second ? c - NV : c), t);
    }
    __syncwarp();
}

// One axis-aligned interior cell of PHASE 1 in the axis-only launch (SEL = 1).  Same arithmetic, in the
// same order, as hot_out_item<1, ..., AXIS = true> (bit-identical face values and moment sums), built
// around what the static upwind sets of such a cell look like:
//   * x faces (entries 0, 1): xi.Sf = xi_x * Sx, the same for every lane -> warp-uniform per point;
//   * y / z faces (entries 2.., in pairs): xi.Sf = xi_y * Sy (xi_z * Sz) does not depend on the point, so a
//     lane is upwind of exactly ONE face of a pair for the whole row.  The lane reconstructs only that
//     face (one face offset, one store address and ONE set of moment accumulators per pair, selected once
//     per cell) and the two faces' moments are separated at the end by masking lanes.
// Needs: no tie (|xi.Sf| < VSMALL) on a y / z face of an axis-aligned cell for any row of the slab
// (k_build_upwind reports that per slab; such slabs take the unified launch) and rows that are whole CU-point groups.
// CI: points per staged chunk (copy granularity, kernel template parameter); CU: points advanced together
// (2 keeps the function at 3 CTAs/SM; a 4-point stage doubles the bytes in flight per warp and halves the
// per-chunk bookkeeping).
template <bool HAS_H, int NE, int CI, int CU, class Prefetch>
__device__ __forceinline__ void hot_axis_item(const StepArgs& a, const HotCtx& x, const HotMeta& cur, double* stages,
                                              int stage_d, uint32_t& q, double* red, Prefetch&& prefetch) {
    static_assert(NE == 4 || NE == 6, "axis-aligned cells have 4 or 6 faces");
    static_assert(CI % CU == 0, "a staged chunk is a whole number of compute groups");
    constexpr int NSLOT = 1 + NE, NFLD = HAS_H ? 2 : 1, NP = NE / 2 - 1;   // NP: y (and z) pairs
    const int lane = x.lane, L = x.L, blk = x.blk;
    const double* gb_ = x.geo;
    const double* txs = x.txs;
    // The cell's first chunk and its geometry record are one commit group, possibly still in flight: stage
    // the second chunk, then wait for the first BEFORE the set-up below reads the record.
    prefetch(0);
    cp_async_commit();
    cp_async_wait<1>();
    __syncwarp();
    const unsigned ownmask = __ballot_sync(0xffffffffu, cur.own != 0);
    const unsigned w4[4] = {cur.mw.x, cur.mw.y, cur.mw.z, cur.mw.w};
    // ---- x faces: masks per point and their warp-uniform summaries per compute group
    unsigned fullx[2], tiex[2], anyx[2], allx[2];
#pragma unroll
    for (int j = 0; j < 2; j++) {
        hot_decode((w4[0] >> (j * 16)) & 0xffffu, L, fullx[j], tiex[j]);
        anyx[j] = __reduce_or_sync(0xffffffffu, hot_spread_any<CU>(fullx[j] | tiex[j]));
        allx[j] = __reduce_and_sync(0xffffffffu, hot_spread_all<CU>(fullx[j]));
    }
    // ---- face storage: running store pointers of this lane's rows (the pair's face is chosen per lane)
    double* const fk_g = a.fkeep_g ? a.fkeep_g + (size_t)a.slab * a.m.nif * blk : nullptr;
    double* const fk_h = (HAS_H && a.fkeep_h) ? a.fkeep_h + (size_t)a.slab * a.m.nif * blk : nullptr;
    const bool keep_on = fk_g != nullptr;
    const ptrdiff_t fk_gh = HAS_H ? fk_h - fk_g : 0;   // the h copy of a face row sits this far behind the g copy
    double* kx[2];
    double* kp[NP];
#pragma unroll
    for (int j = 0; j < 2; j++) kx[j] = fk_g + ((size_t)__shfl_sync(0xffffffffu, cur.face, j) * blk + lane);
    bool sel[NP], act[NP];
    double rsel[NP];
#pragma unroll
    for (int p = 0; p < NP; p++) {
        unsigned fa, ta, fb, tb2;
        hot_decode((w4[1 + p] & 0xffffu), L, fa, ta);
        hot_decode((w4[1 + p] >> 16) & 0xffffu, L, fb, tb2);
        sel[p] = fa != 0;
        act[p] = (fa | fb) != 0;
        const int d = 1 + p;
        const double ra = gb_[6 * (3 + 2 * p) + 3 + d], rb = gb_[6 * (4 + 2 * p) + 3 + d];
        rsel[p] = sel[p] ? ra : rb;
        const int fa_id = __shfl_sync(0xffffffffu, cur.face, 2 + 2 * p), fb_id = __shfl_sync(0xffffffffu, cur.face, 3 + 2 * p);
        kp[p] = fk_g + ((size_t)(sel[p] ? fa_id : fb_id) * blk + lane);
    }
    // ---- moment accumulators: two x faces, one per pair
    double ax[2][4], bx[2][2], ap[NP][4], bp[NP][2];
#pragma unroll
    for (int j = 0; j < 2; j++) {
        ax[j][0] = ax[j][1] = ax[j][2] = ax[j][3] = 0.0;
        bx[j][0] = bx[j][1] = 0.0;
    }
#pragma unroll
    for (int p = 0; p < NP; p++) {
        ap[p][0] = ap[p][1] = ap[p][2] = ap[p][3] = 0.0;
        bp[p][0] = bp[p][1] = 0.0;
    }

    for (int ch = 0; ch < x.nchunk; ch++) {
        if (ch > 0) {
            prefetch(ch);
            cp_async_commit();
            cp_async_wait<1>();
            __syncwarp();
        }
        const double* sg = stages + (q & 1) * stage_d + lane;
        const int nsub = min(CI / CU, (L - ch * CI) / CU);   // the last chunk of a row may be short (rows are whole groups)
#pragma unroll 1
        for (int sub = 0; sub < nsub; sub++) {
        const int i0 = ch * CI + sub * CU;
        const int tb = x.cb + i0;
#pragma unroll
        for (int fld = 0; fld < NFLD; fld++) {
            const double* sf = sg + fld * NSLOT * CI * 32 + sub * CU * 32;
            // ---- gradient (stock leastSquaresGrad, zeroBoundaryGrad.C:90-99): one component per face
            double v[CU], g[3][CU], base[CU];
            {
                const double2 G01 = lds2(gb_);
                const double G2 = gb_[2];
#pragma unroll
                for (int u = 0; u < CU; u++) {
                    v[u] = sf[u * 32];
                    g[0][u] = G01.x * v[u]; g[1][u] = G01.y * v[u]; g[2][u] = G2 * v[u];
                }
            }
#pragma unroll
            for (int j = 0; j < NE; j++) {
                const int d = j >> 1;
                const double G = gb_[6 * (1 + j) + d];
#pragma unroll
                for (int u = 0; u < CU; u++) g[d][u] = fma(G, sf[((1 + j) * CI + u) * 32], g[d][u]);
            }
            // ---- value at the cell centre moved back by half a step (discreteVelocity.C:498-502)
            double W[CU][4];
#pragma unroll
            for (int u = 0; u < CU; u++) {
                const double2 t0 = lds2(txs + (tb + u) * 6), t1 = lds2(txs + (tb + u) * 6 + 2);
                base[u] = NE == 4 ? fma(t0.x, g[0][u], fma(x.yh, g[1][u], v[u]))
                                  : fma(t0.x, g[0][u], fma(x.yh, g[1][u], fma(x.zh, g[2][u], v[u])));
                W[u][0] = t0.y; W[u][1] = t1.x; W[u][2] = t1.y; W[u][3] = txs[(tb + u) * 6 + 4];
            }
            // ---- x faces
#pragma unroll
            for (int j = 0; j < 2; j++) {
                if (!((anyx[j] >> i0) & 1u)) continue;                      // warp-uniform
                const double r = gb_[6 * (1 + j) + 3];
                double* const keep = fld == 0 ? kx[j] : kx[j] + fk_gh;
                if ((allx[j] >> i0) & 1u) {                                  // warp-uniform
#pragma unroll
                    for (int u = 0; u < CU; u++) {
                        const double val = fma(r, g[0][u], base[u]);
                        if (keep_on) __stcs(keep + u * 32, val);
                        if (fld == 0) {
                            ax[j][0] = fma(W[u][0], val, ax[j][0]); ax[j][1] = fma(W[u][1], val, ax[j][1]);
                            ax[j][2] = fma(W[u][2], val, ax[j][2]); ax[j][3] = fma(W[u][3], val, ax[j][3]);
                        } else {
                            bx[j][0] = fma(W[u][0], val, bx[j][0]); bx[j][1] = fma(W[u][1], val, bx[j][1]);
                        }
                    }
                } else {
                    // the group that holds the sign change of xi_x: this side's share is all, half (tie,
                    // :513-529) or none; the owner keeps the value unless phi < -VSMALL
                    const unsigned fb = fullx[j] >> i0, tbits = tiex[j] >> i0;
                    const unsigned wbk = ((ownmask >> j) & 1u) ? (fb | tbits) : fb;
#pragma unroll
                    for (int u = 0; u < CU; u++) {
                        double val = fma(r, g[0][u], base[u]);
                        if (keep_on && ((wbk >> u) & 1u)) __stcs(keep + u * 32, val);
                        const int hi = ((fb >> u) & 1u) ? 0x3ff00000 : (((tbits >> u) & 1u) ? 0x3fe00000 : 0);
                        val *= __hiloint2double(hi, 0);
                        if (fld == 0) {
                            ax[j][0] = fma(W[u][0], val, ax[j][0]); ax[j][1] = fma(W[u][1], val, ax[j][1]);
                            ax[j][2] = fma(W[u][2], val, ax[j][2]); ax[j][3] = fma(W[u][3], val, ax[j][3]);
                        } else {
                            bx[j][0] = fma(W[u][0], val, bx[j][0]); bx[j][1] = fma(W[u][1], val, bx[j][1]);
                        }
                    }
                }
            }
            // ---- y / z pairs: the one face of the pair this lane is upwind of
#pragma unroll
            for (int p = 0; p < NP; p++) {
                double* const keep = fld == 0 ? kp[p] : kp[p] + fk_gh;
                const bool st = keep_on && act[p];
#pragma unroll
                for (int u = 0; u < CU; u++) {
                    const double val = fma(rsel[p], g[1 + p][u], base[u]);
                    if (st) __stcs(keep + u * 32, val);
                    if (fld == 0) {
                        ap[p][0] = fma(W[u][0], val, ap[p][0]); ap[p][1] = fma(W[u][1], val, ap[p][1]);
                        ap[p][2] = fma(W[u][2], val, ap[p][2]); ap[p][3] = fma(W[u][3], val, ap[p][3]);
                    } else {
                        bp[p][0] = fma(W[u][0], val, bp[p][0]); bp[p][1] = fma(W[u][1], val, bp[p][1]);
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 2; j++) kx[j] += CU * 32;
#pragma unroll
        for (int p = 0; p < NP; p++) kp[p] += CU * 32;
        }
        __syncwarp();   // every lane is done with this stage before it is refilled
        q++;
    }

    // ---- face moments, two faces per pass: (x-, x+), then each pair with its lanes split by the face they chose
    {
        const int f0 = __shfl_sync(0xffffffffu, cur.face, 0), f1 = __shfl_sync(0xffffffffu, cur.face, 1);
        hot_reduce_faces2<HAS_H>(a, x, ax[0], ax[1], bx[0], bx[1], (size_t)2 * f0 + ((ownmask & 1u) ? 0 : 1),
                                 (size_t)2 * f1 + ((ownmask & 2u) ? 0 : 1), anyx[0] != 0, anyx[1] != 0, red, lane);
    }
#pragma unroll
    for (int p = 0; p < NP; p++) {
        const bool la = act[p] && sel[p], lb = act[p] && !sel[p];
        double mA[4], mB[4], hA[2], hB[2];
#pragma unroll
        for (int k = 0; k < 4; k++) { mA[k] = la ? ap[p][k] : 0.0; mB[k] = lb ? ap[p][k] : 0.0; }
#pragma unroll
        for (int k = 0; k < 2; k++) { hA[k] = la ? bp[p][k] : 0.0; hB[k] = lb ? bp[p][k] : 0.0; }
        const bool onA = __any_sync(0xffffffffu, la), onB = __any_sync(0xffffffffu, lb);
        const int fa_id = __shfl_sync(0xffffffffu, cur.face, 2 + 2 * p), fb_id = __shfl_sync(0xffffffffu, cur.face, 3 + 2 * p);
        hot_reduce_faces2<HAS_H>(a, x, mA, mB, hA, hB, (size_t)2 * fa_id + (((ownmask >> (2 + 2 * p)) & 1u) ? 0 : 1),
                                 (size_t)2 * fb_id + (((ownmask >> (3 + 2 * p)) & 1u) ? 0 : 1), onA, onB, red, lane);
    }
}

// -------------------------------------------------------------------------------------------------
// stages 2.1 + 3 (PHASE 1: face moments, boundary-face values, lagged boundary gradient) and
// stage 4 (PHASE 2: relaxed internal-face values into the slab flux buffer).
// discreteVelocity.C:412-691 / fvDVM.C:473-516 / discreteVelocity.C:867-881
// SEL: 0 = every cell; 1 = only axis-aligned cells (the light variant alone fits 3 CTAs/SM);
//      2 = only the other cells (run as a second launch when SEL = 1 is used)
#ifndef HOT_K1_MINB
#define HOT_K1_MINB 2
#endif
template <int PHASE, bool HAS_H, int NE, int TW, int CI, int SEL>
__global__ void __launch_bounds__(HOT_WARPS * 32, (SEL == 1 ? 3 : (PHASE == 1 ? HOT_K1_MINB : HOT_MINB(CI))))
k_hot_outgoing(StepArgs a) {
    using P = HotPlan<PHASE, HAS_H, NE, TW, CI>;
    auto mine = [](const HotMeta& M) {
        if (M.ne > NE) return false;
        const bool axis = M.cls != 0 && M.ne == NE && M.nint == NE && (NE == 4 || NE == 6);
        return SEL == 0 || (SEL == 1 ? axis : !axis);
    };
    constexpr int NSLOT = P::NSLOT, NTOT = P::NFLD * P::NSLOT;
    extern __shared__ __align__(128) unsigned char dyn[];
    const DevDV& dv = a.dv;
    const int L = dv.L, nc = a.m.nc, blk = L * 32;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const double hd = -0.5 * a.dt;
    double* txs = reinterpret_cast<double*>(dyn);
    hot_fill_txs(dv, hd, txs);
    unsigned char* wbase = dyn + P::txs_bytes(dv.ntab) + wib * P::PER_WARP;
    unsigned long long* sptr = reinterpret_cast<unsigned long long*>(wbase);   // [2][HOT_PTRS]
    double* geo = reinterpret_cast<double*>(wbase) + 2 * HOT_PTRS;             // [2][GEO_D]
    double* stages = geo + 2 * P::GEO_D;                                       // [HOT_STAGES][STAGE_D]
    double* extra = stages + HOT_STAGES * P::STAGE_D;
    __syncthreads();

    const size_t slab_c = (size_t)a.slab * nc * blk, slab_b = (size_t)a.slab * a.m.nbf * blk;
    const double* gbs = a.gb + slab_c;
    const double* hbs = HAS_H ? a.hb + slab_c : nullptr;
    const double* gam_g = a.gam_old_g + slab_b;
    const double* gam_h = HAS_H ? a.gam_old_h + slab_b : nullptr;
    const int grow = a.slab * 32 + lane;
    HotCtx x;
    x.txs = txs;
    x.red = extra; x.xtab = extra; x.unic = extra + NE * 4 * TW;
    double* frecs = extra + NE * 4 * TW + NE * 2;   // PHASE 2: [2][NE][FCOEF_N]
    auto stage_frec = [&](const HotMeta& M, double* dst) {
        if (PHASE != 2) return;
        const uint32_t sd = smem_u32(dst);
#pragma unroll
        for (int r = 0; r < (NE * 6 + 31) / 32; r++) {
            const int p = lane + 32 * r;
            const int jf = min(p / 6, NE - 1);
            const int f = __shfl_sync(0xffffffffu, M.face, jf);
            if (p < NE * 6 && jf < M.nint)
                cp_async16(sd + p * 16, reinterpret_cast<const char*>(a.fcoef + (size_t)f * FCOEF_N) + (p % 6) * 16);
        }
    };
    x.y = dv.row_y[grow]; x.z = dv.row_z[grow]; x.wr = dv.row_w[grow];
    x.yh = hd * x.y; x.zh = hd * x.z;
    x.kd = (double)(a.gas.K + 3 - a.gas.D);
    x.cb = dv.row_cbase[grow];
    x.tmin = 0; x.span = 0;
    const int Ln = dv_len(dv, a.slab);     // points per row of THIS slab (strides use L)
    if (PHASE == 2) table_range(dv, x.cb, x.tmin, x.span, Ln);
    x.lane = lane; x.L = Ln; x.blk = blk; x.nm = a.nm;
    x.nchunk = (Ln + CI - 1) / CI;
    x.slab_b = slab_b;

    const unsigned long long pol_el = l2_evict_last_policy();
    const int nw = gridDim.x * HOT_WARPS;
    // items [a.item0, a.item1) of the traversal order (the whole mesh unless the launch is split by cell class)
    const int item_end = a.item1 > 0 ? a.item1 : nc;
    int item = a.item0 + blockIdx.x * HOT_WARPS + wib;
    uint32_t q = 0;        // flat chunk counter of this warp: stage = q & 1
    int gsel = 0;          // pointer / geometry buffer of the current item
    HotMeta cur{}, nxt{};
    if (item < item_end) {
        hot_meta_issue(a, item, lane, cur);
        hot_meta_commit<HAS_H>(a, lane, gbs, hbs, gam_g, gam_h, NE, sptr, cur);
        if (mine(cur)) {
            hot_stage<CI, NTOT, NSLOT, false>(sptr, cur.ne, 0, stages, lane);
            hot_stage_geo(a.geo6 + (size_t)(cur.e0 + cur.c) * 6, cur.ne, geo, lane);
            stage_frec(cur, frecs);
        }
        cp_async_commit();
    }
    while (item < item_end) {
        const int nitem = item + nw;
        const bool has_next = nitem < item_end;
        unsigned long long* sp_cur = sptr + gsel * HOT_PTRS;
        unsigned long long* sp_nxt = sptr + (gsel ^ 1) * HOT_PTRS;
        if (has_next) hot_meta_issue(a, nitem, lane, nxt);
        bool next_ok = false;
        // called once, right before the last chunk of the current cell is computed
        auto stage_next_item = [&](double* stage) {
            if (has_next) {
                hot_meta_commit<HAS_H>(a, lane, gbs, hbs, gam_g, gam_h, NE, sp_nxt, nxt);
                next_ok = mine(nxt);
            }
            if (next_ok) {
                hot_stage<CI, NTOT, NSLOT, false>(sp_nxt, nxt.ne, 0, stage, lane);
                hot_stage_geo(a.geo6 + (size_t)(nxt.e0 + nxt.c) * 6, nxt.ne, geo + (gsel ^ 1) * P::GEO_D, lane);
                stage_frec(nxt, frecs + (gsel ^ 1) * P::FREC_D);
            }
        };
        if (!mine(cur)) {       // too many faces (the generic kernels take it) or the other launch's cell
            __syncwarp();
            stage_next_item(stages + (q & 1) * P::STAGE_D);
            cp_async_commit();
        } else {
            x.geo = geo + gsel * P::GEO_D;
            x.frec = frecs + gsel * P::FREC_D;
            const bool interior = cur.ne == NE && cur.nint == NE;
            if (SEL != 2 && interior && cur.cls && (NE == 4 || NE == 6)) {
                uint32_t soff[NSLOT];   // 16-byte units
                soff[0] = (uint32_t)cur.c * (uint32_t)(blk / 2) + (uint32_t)lane;
#pragma unroll
                for (int j = 0; j < NE; j++)
                    soff[1 + j] = (uint32_t)__shfl_sync(0xffffffffu, cur.other, j) * (uint32_t)(blk / 2) + (uint32_t)lane;
#ifdef HOT_P1_EVICT_LAST
                constexpr bool EVL = SEL == 1 && PHASE == 1;
#else
                constexpr bool EVL = false;
#endif
                auto prefetch = [&](int ch) {
                    double* st = stages + ((q & 1) ^ 1) * P::STAGE_D;
                    if (ch + 1 < x.nchunk) hot_stage_off<CI, P::NFLD, NSLOT, EVL>(gbs, hbs, soff, ch + 1, st, lane, pol_el);
                    else stage_next_item(st);
                };
                if constexpr (SEL == 1 && PHASE == 1 && (NE == 4 || NE == 6)) {
                    // reduction scratch: the stage consumed last (the chunk count per cell is fixed, so its
                    // parity after the loop is known now) or the plan's own scratch
                    double* red = P::RED_IN_STAGE ? stages + (((q + x.nchunk) & 1) ^ 1) * P::STAGE_D : x.red;
                    hot_axis_item<HAS_H, NE, CI, HOT_AXIS_CU>(a, x, cur, stages, P::STAGE_D, q, red, prefetch);
                }
                else
                    hot_out_item<PHASE, HAS_H, NE, TW, CI, true, (NE == 4 || NE == 6)>(a, x, cur, stages, P::STAGE_D, q, prefetch);
            } else if (SEL != 1 && cur.ne == NE) {
                // all NE entries exist; boundary entries (if any) stream the lagged gradient from another
                // array, so only all-internal cells can use the register offsets
                uint32_t soff[NSLOT];   // 16-byte units
                soff[0] = (uint32_t)cur.c * (uint32_t)(blk / 2) + (uint32_t)lane;
#pragma unroll
                for (int j = 0; j < NE; j++)
                    soff[1 + j] = (uint32_t)__shfl_sync(0xffffffffu, cur.other, j) * (uint32_t)(blk / 2) + (uint32_t)lane;
                auto prefetch = [&](int ch) {
                    double* st = stages + ((q & 1) ^ 1) * P::STAGE_D;
                    if (ch + 1 < x.nchunk) {
                        if (interior) hot_stage_off<CI, P::NFLD, NSLOT>(gbs, hbs, soff, ch + 1, st, lane);
                        else hot_stage<CI, NTOT, NSLOT, true>(sp_cur, NE, ch + 1, st, lane);
                    } else stage_next_item(st);
                };
                hot_out_item<PHASE, HAS_H, NE, TW, CI, true, false>(a, x, cur, stages, P::STAGE_D, q, prefetch);
            } else if (SEL != 1) {
                auto prefetch = [&](int ch) {
                    double* st = stages + ((q & 1) ^ 1) * P::STAGE_D;
                    if (ch + 1 < x.nchunk) hot_stage<CI, NTOT, NSLOT, false>(sp_cur, cur.ne, ch + 1, st, lane);
                    else stage_next_item(st);
                };
                hot_out_item<PHASE, HAS_H, NE, TW, CI, false, false>(a, x, cur, stages, P::STAGE_D, q, prefetch);
            }
        }
        cur = nxt; item = nitem; gsel ^= 1;
    }
    cp_async_wait<0>();
}

// -------------------------------------------------------------------------------------------------
// stage 5 + the cell moments of stage 6: gTilde <- -1/3 gTilde + 4/3 gBarP - dt/V sum_f +-(xi.Sf) g_f
// (discreteVelocity.C:934-978, fvDVM.C:612-622,712-721).  Streams: gTilde, gBarP, one per face.
// w = -1/3 gTilde + 4/3 gBarP (discreteVelocity.C:937): the cell part of the update.  One definition, so that
// the kernels that read (gTilde, gBarP) and the ones that read a stored w (face-storage slabs: the half-step
// kernel leaves w in place of gTilde and gBarP in a transient slab buffer) give the same bits.
__device__ __forceinline__ double hot_w_combine(double gt, double gb) { return fma(4.0 / 3, gb, (-1.0 / 3) * gt); }

template <bool HAS_H, int NE, int CI>
struct HotUpdPlan {
    static constexpr int NFLD = HAS_H ? 2 : 1;
    static constexpr int NSLOT = 2 + NE;
    static constexpr int STAGE_D = NFLD * NSLOT * CI * 32;
    static constexpr int PER_WARP_D = 2 * HOT_PTRS + HOT_STAGES * STAGE_D + 32 * 17;
    static constexpr size_t PER_WARP = ((size_t)PER_WARP_D * 8 + 127) / 128 * 128;
    static __host__ __device__ size_t txs_bytes(int ntab) { return ((size_t)(ntab + HOT_CI_MAX) * 48 + 127) / 128 * 128; }
    static __host__ size_t total(int ntab) { return txs_bytes(ntab) + HOT_WARPS * PER_WARP; }
};

struct HotUpdMeta {
    int c, e0, ne, nint, face;   // face: lane j < ne: face id of entry j
    int cls;                     // 1: axis-aligned interior cell (entries x-, x+, y-, y+[, z-, z+])
    int so, sf;                  // this lane's stream slot (lane - 2): other cell / face id of that entry
};

// loads only (see hot_meta_issue).  NPRE: cell streams ahead of the face streams (gTilde, gBarP: 2; or the
// combination w = -1/3 gTilde + 4/3 gBarP alone: 1)
template <int NPRE = 2>
__device__ __forceinline__ void hot_upd_issue(const int* cmeta, int item, int lane, HotUpdMeta& M) {
    const int* rec = cmeta + (size_t)item * CMETA_N;
    M.c = ldg_early(rec + 20);
    const int2 h2 = ldg_early2(rec);
    M.e0 = h2.x;
    M.ne = h2.y;
    M.face = ldg_early(rec + 10 + (lane & 7));
    M.so = ldg_early(rec + 2 + ((lane - NPRE) & 7));
    M.sf = ldg_early(rec + 10 + ((lane - NPRE) & 7));
}

// fsrc_g/h: where internal-face values come from (flux buffer or kept face values of the slab)
template <bool HAS_H, int NPRE = 2>
__device__ __forceinline__ void hot_upd_commit(const StepArgs& a, int lane, const double* gts, const double* hts,
                                               const double* gbs, const double* hbs, const double* gsbs,
                                               const double* hsbs, const double* fsrc_g, const double* fsrc_h, int NE,
                                               unsigned long long* sp, HotUpdMeta& M) {
    const int blk = a.dv.L * 32;
    const int NSLOT = NPRE + NE;
    const int pk = M.ne;
    M.ne = pk & 0xff; M.nint = (pk >> 8) & 0xff; M.cls = (pk >> 16) & 0xff;
    M.face = (lane < M.ne && M.ne <= NE) ? (M.face & 0x7fffffff) : 0;
    if (M.ne <= NE && lane < NPRE + M.ne) {
        const int o = M.so, f = M.sf & 0x7fffffff;
#pragma unroll
        for (int fld = 0; fld < (HAS_H ? 2 : 1); fld++) {
            const double* src;
            if (lane == 0) src = (fld ? hts : gts) + (size_t)M.c * blk;
            else if (NPRE == 2 && lane == 1) src = (fld ? hbs : gbs) + (size_t)M.c * blk;
            else if (o >= 0) src = (fld ? fsrc_h : fsrc_g) + (size_t)f * blk;
            else src = (fld ? hsbs : gsbs) + (size_t)(-1 - o) * blk;
            sp[fld * NSLOT + lane] = (unsigned long long)src;
        }
    }
    __syncwarp();
}

template <bool HAS_H, int NE, int CI>
__global__ void __launch_bounds__(HOT_WARPS * 32, 3)
k_hot_update(StepArgs a) {
    using P = HotUpdPlan<HAS_H, NE, CI>;
    constexpr int NSLOT = P::NSLOT, NTOT = P::NFLD * P::NSLOT;
    extern __shared__ __align__(128) unsigned char dyn[];
    const DevDV& dv = a.dv;
    const int L = dv.L, nc = a.m.nc, blk = L * 32;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double* txs = reinterpret_cast<double*>(dyn);
    hot_fill_txs(dv, 1.0, txs);
    unsigned char* wbase = dyn + P::txs_bytes(dv.ntab) + wib * P::PER_WARP;
    unsigned long long* sptr = reinterpret_cast<unsigned long long*>(wbase);
    double* stages = reinterpret_cast<double*>(wbase) + 2 * HOT_PTRS;
    double* red = stages + HOT_STAGES * P::STAGE_D;
    __syncthreads();

    const size_t slab_c = (size_t)a.slab * nc * blk, slab_b = (size_t)a.slab * a.m.nbf * blk;
    double* gts = a.gt + slab_c;
    double* hts = HAS_H ? a.ht + slab_c : nullptr;
    const double* gbs = a.gb + slab_c;
    const double* hbs = HAS_H ? a.hb + slab_c : nullptr;
    const double* gsbs = a.gsb + slab_b;
    const double* hsbs = HAS_H ? a.hsb + slab_b : nullptr;
    const int grow = a.slab * 32 + lane;
    const double y = dv.row_y[grow], z = dv.row_z[grow], wr = dv.row_w[grow];
    const int cb = dv.row_cbase[grow];
    const int Ln = dv_len(dv, a.slab);     // points per row of THIS slab (strides use L)
    const int nchunk = (Ln + CI - 1) / CI;
    const int nm = a.nm;

    const unsigned long long pol_ef = l2_evict_first_policy();
    const int nw = gridDim.x * HOT_WARPS;
    int item = blockIdx.x * HOT_WARPS + wib;
    uint32_t q = 0;
    int gsel = 0;
    HotUpdMeta cur{}, nxt{};
    if (item < nc) {
        hot_upd_issue(a.cmeta, item, lane, cur);
        hot_upd_commit<HAS_H>(a, lane, gts, hts, gbs, hbs, gsbs, hsbs, a.fbuf_g, a.fbuf_h, NE, sptr, cur);
        if (cur.ne <= NE) hot_stage<CI, NTOT, NSLOT, false>(sptr, cur.ne + 1, 0, stages, lane);
        cp_async_commit();
    }
    while (item < nc) {
        const int nitem = item + nw;
        const bool has_next = nitem < nc;
        unsigned long long* sp_cur = sptr + gsel * HOT_PTRS;
        unsigned long long* sp_nxt = sptr + (gsel ^ 1) * HOT_PTRS;
        if (has_next) hot_upd_issue(a.cmeta, nitem, lane, nxt);
        // called once, right before the last chunk of the current cell is computed
        auto stage_next_item = [&](double* st) {
            if (!has_next) return;
            hot_upd_commit<HAS_H>(a, lane, gts, hts, gbs, hbs, gsbs, hsbs, a.fbuf_g, a.fbuf_h, NE, sp_nxt, nxt);
            if (nxt.ne <= NE) hot_stage<CI, NTOT, NSLOT, false>(sp_nxt, nxt.ne + 1, 0, st, lane);
        };
        if (cur.ne > NE) {
            stage_next_item(stages + (q & 1) * P::STAGE_D);
            cp_async_commit();
            cur = nxt; item = nitem; gsel ^= 1;
            continue;
        }
        const int ne = cur.ne, c = cur.c;
        // interior cells stage from register offsets (16-byte units), see hot_stage_off
        const bool interior = ne == NE && cur.nint == NE;
        const uint32_t offc = (uint32_t)c * (uint32_t)(blk / 2) + (uint32_t)lane;
        uint32_t offf[NE];
#pragma unroll
        for (int j = 0; j < NE; j++) offf[j] = (uint32_t)__shfl_sync(0xffffffffu, cur.face, j) * (uint32_t)(blk / 2) + (uint32_t)lane;
        auto stage_interior = [&](int ch, double* st) {
            const uint32_t sdst = smem_u32(st) + (uint32_t)lane * 16u, coff = (uint32_t)ch * (CI * 256u);
#pragma unroll
            for (int fld = 0; fld < P::NFLD; fld++) {
                hot_stage_one_ef<CI>(sdst + (fld * NSLOT + 0) * (CI * 256), fld ? hts : gts, offc, coff, pol_ef);
                hot_stage_one_ef<CI>(sdst + (fld * NSLOT + 1) * (CI * 256), fld ? hbs : gbs, offc, coff, pol_ef);
#pragma unroll
                for (int j = 0; j < NE; j++)
                    hot_stage_one<CI>(sdst + (fld * NSLOT + 2 + j) * (CI * 256), fld ? a.fbuf_h : a.fbuf_g, offf[j], coff);
            }
        };
        // outward area vectors of the cell's faces (sign folded in), flux = x*Sx + (y*Sy + z*Sz)
        double Sx[NE], cyz[NE];
#pragma unroll
        for (int j = 0; j < NE; j++) {
            Sx[j] = 0.0; cyz[j] = 0.0;
            if (j < ne) {
                const double* S = a.geoS + (size_t)(cur.e0 + j) * 4;
                Sx[j] = S[0];
                cyz[j] = fma(y, S[1], z * S[2]);
            }
        }
        const double dtv = a.dt / a.m.V[c];
        double A[4] = {0, 0, 0, 0}, B[2] = {0, 0};
        double* gdst = gts + (size_t)c * blk + lane;
        double* hdst = HAS_H ? hts + (size_t)c * blk + lane : nullptr;
        for (int ch = 0; ch < nchunk; ch++) {
            {
                double* st = stages + ((q & 1) ^ 1) * P::STAGE_D;
                if (ch + 1 < nchunk) {
                    if (interior) stage_interior(ch + 1, st);
                    else hot_stage<CI, NTOT, NSLOT, false>(sp_cur, ne + 1, ch + 1, st, lane);
                } else stage_next_item(st);
            }
            cp_async_commit();
            cp_async_wait<1>();
            __syncwarp();
            const double* sg = stages + (q & 1) * P::STAGE_D + lane;
            const int i0 = ch * CI, tb = cb + i0;
            double xx[CI], W[CI][4];
#pragma unroll
            for (int u = 0; u < CI; u++) {
                const double2 t0 = lds2(txs + (tb + u) * 6), t1 = lds2(txs + (tb + u) * 6 + 2), t2 = lds2(txs + (tb + u) * 6 + 4);
                W[u][0] = t0.y; W[u][1] = t1.x; W[u][2] = t1.y; W[u][3] = t2.x; xx[u] = t2.y;
            }
#pragma unroll
            for (int fld = 0; fld < P::NFLD; fld++) {
                const double* sf = sg + fld * NSLOT * CI * 32;
                double sum[CI];
#pragma unroll
                for (int u = 0; u < CI; u++) sum[u] = 0.0;
#pragma unroll
                for (int j = 0; j < NE; j++) {
                    if (j < ne) {
#pragma unroll
                        for (int u = 0; u < CI; u++)
                            sum[u] = fma(fma(xx[u], Sx[j], cyz[j]), sf[((2 + j) * CI + u) * 32], sum[u]);   // :952-955
                    }
                }
#pragma unroll
                for (int u = 0; u < CI; u++) {
                    const double vnew = fma(-sum[u], dtv, hot_w_combine(sf[u * 32], sf[(CI + u) * 32]));   // :937,952
                    if (i0 + u >= Ln) continue;   // tail chunk (warp-uniform)
                    __stcs((fld == 0 ? gdst : hdst) + (i0 + u) * 32, vnew);
                    if (fld == 0) {
                        A[0] = fma(W[u][0], vnew, A[0]); A[1] = fma(W[u][1], vnew, A[1]);
                        A[2] = fma(W[u][2], vnew, A[2]); A[3] = fma(W[u][3], vnew, A[3]);
                    } else {
                        B[0] = fma(W[u][0], vnew, B[0]); B[1] = fma(W[u][1], vnew, B[1]);
                    }
                }
            }
            __syncwarp();
            q++;
        }
        double vv[16];
        expand_g(A, wr, y, z, vv);
        double uu[NM_H] = {0, 0, 0, 0};
        if (HAS_H) expand_h(B, wr, y, z, uu);
        vv[13] = uu[0]; vv[14] = uu[1]; vv[15] = uu[2];
        const double tot = warp_reduce16_smem(vv, red, lane);
        if (lane < 16 && lane < nm) atomicAdd(a.cslot + (size_t)c * nm + lane, tot);
        if (HAS_H) {
            const double t3 = warp_sum(uu[3]);
            if (lane == 0) atomicAdd(a.cslot + (size_t)c * nm + 16, t3);
        }
        cur = nxt; item = nitem; gsel ^= 1;
    }
    cp_async_wait<0>();
}

// -------------------------------------------------------------------------------------------------
// Face-storage slabs: stage 4 and stage 5 in one pass.  Every face value gBar_f kept by phase 1 is read
// by the two cells that share the face; each relaxes it to g_f with the face equilibrium
// (discreteVelocity.C:867-881) and adds its flux (:934-978).  Boundary-face values come relaxed from
// k_bnd_relax.  Replaces k_hot_outgoing<2> + k_hot_update and their flux buffer round trip.
// WMODE 1: the cell stream is w = -1/3 gTilde + 4/3 gBarP, left in place of gTilde by the half-step kernel
// (one stream and 8 bytes per update less; gBarP of the slab need not outlive phase 1).
// WMODE 2: the cell stream is gTilde itself and w is formed here, gBarP = (1 - rf) gTilde + rf gS from the cell's
// macros of the step start (discreteVelocity.C:393-406; they are only replaced after phase 2): slabs whose phase 1
// applied the half step on the fly (dugks_pencil.cuh) never wrote gBarP or w.  Same operations as
// k_hot_halfstep + hot_w_combine, so the bits are those of WMODE 1.
template <bool HAS_H, int NE, int TW, int CI, int WMODE = 0>
struct HotRelaxPlan {
    static constexpr int NFLD = HAS_H ? 2 : 1;
    static constexpr int NPRE = WMODE ? 1 : 2;
    static constexpr int NSLOT = NPRE + NE;
    static constexpr int STAGE_D = NFLD * NSLOT * CI * 32;
    static constexpr int NTABS = WMODE == 2 ? NE + 1 : NE;   // one more equilibrium table: the cell's own
    // the moment reduction (32 x 17 doubles) runs through the face tables, which are dead by then
    static constexpr int TAB_D = (NTABS * 4 * TW + NE * 2) > 32 * 17 ? (NTABS * 4 * TW + NE * 2) : 32 * 17;
    // face equilibrium records + outward area vectors of a cell (+ the cell's macro record, WMODE 2)
    static constexpr int REC_D = NE * (FCOEF_N + 4) + (WMODE == 2 ? FCOEF_N : 0);
    static constexpr int PER_WARP_D = 2 * HOT_PTRS + 2 * REC_D + HOT_STAGES * STAGE_D + TAB_D;
    static constexpr size_t PER_WARP = ((size_t)PER_WARP_D * 8 + 127) / 128 * 128;
    static __host__ __device__ size_t txs_bytes(int ntab) { return ((size_t)(ntab + HOT_CI_MAX) * 48 + 127) / 128 * 128; }
    static __host__ size_t total(int ntab) { return txs_bytes(ntab) + HOT_WARPS * PER_WARP; }
};

template <bool HAS_H, int NE, int TW, int CI, int WMODE>
__global__ void __launch_bounds__(HOT_WARPS * 32, HOT_MINB(CI))
k_hot_relax_update(StepArgs a) {
    using P = HotRelaxPlan<HAS_H, NE, TW, CI, WMODE>;
    constexpr int NSLOT = P::NSLOT, NTOT = P::NFLD * P::NSLOT, NPRE = P::NPRE;
    extern __shared__ __align__(128) unsigned char dyn[];
    const DevDV& dv = a.dv;
    const int L = dv.L, nc = a.m.nc, blk = L * 32;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double* txs = reinterpret_cast<double*>(dyn);
    hot_fill_txs(dv, 1.0, txs);
    unsigned char* wbase = dyn + P::txs_bytes(dv.ntab) + wib * P::PER_WARP;
    unsigned long long* sptr = reinterpret_cast<unsigned long long*>(wbase);
    double* recs = reinterpret_cast<double*>(wbase) + 2 * HOT_PTRS;   // [2][NE][FCOEF_N] | [NE][4] per buffer
    double* stages = recs + 2 * P::REC_D;
    double* xtab = stages + HOT_STAGES * P::STAGE_D;       // [NTABS][TW][4]
    double* unic = xtab + P::NTABS * 4 * TW;               // [NE][2] omrf, RT
    double* ctab = xtab + NE * 4 * TW;                     // WMODE 2: [TW][4] half-step table of the cell
    __syncthreads();

    const size_t slab_c = (size_t)a.slab * nc * blk, slab_b = (size_t)a.slab * a.m.nbf * blk;
    const size_t slab_f = (size_t)a.slab * a.m.nif * blk;
    double* gts = a.gt + slab_c;
    double* hts = HAS_H ? a.ht + slab_c : nullptr;
    const double* gbs = a.gb + slab_c;
    const double* hbs = HAS_H ? a.hb + slab_c : nullptr;
    const double* gsbs = a.gsb + slab_b;
    const double* hsbs = HAS_H ? a.hsb + slab_b : nullptr;
    const double* fk_g = a.fkeep_g + slab_f;
    const double* fk_h = HAS_H ? a.fkeep_h + slab_f : nullptr;
    const int grow = a.slab * 32 + lane;
    const double y = dv.row_y[grow], z = dv.row_z[grow], wr = dv.row_w[grow];
    const int cb = dv.row_cbase[grow];
    const int Ln = dv_len(dv, a.slab);     // points per row of THIS slab (strides use L)
    int tmin, span;
    table_range(dv, cb, tmin, span, Ln);
    const int nchunk = (Ln + CI - 1) / CI;
    const int nm = a.nm;
    const double kd = (double)(a.gas.K + 3 - a.gas.D);

    // face equilibrium records (16-byte pieces: 6 per face) and outward area vectors (2 per entry) of a
    // cell into `dst`, same commit group as its first chunk
    auto stage_records = [&](const HotUpdMeta& M, double* dst) {
        const uint32_t sd = smem_u32(dst);
#pragma unroll
        for (int r = 0; r < (NE * 8 + 31) / 32; r++) {
            const int p = lane + 32 * r;
            const int jf = min(p / 6, NE - 1);
            const int f = __shfl_sync(0xffffffffu, M.face, jf);
            if (p < NE * 6) {
                if (jf < M.ne) cp_async16(sd + p * 16, reinterpret_cast<const char*>(a.fcoef + (size_t)f * FCOEF_N) + (p % 6) * 16);
            } else if (p < NE * 8) {
                const int pe = p - NE * 6;
                if (pe / 2 < M.ne)
                    cp_async16(sd + NE * FCOEF_N * 8 + pe * 16, reinterpret_cast<const char*>(a.geoS + (size_t)M.e0 * 4) + pe * 16);
            }
        }
        if (WMODE == 2 && lane < FCOEF_N / 2)   // half-step coefficient record of the cell (k_cell_coef)
            cp_async16(sd + (NE * (FCOEF_N + 4) + 2 * lane) * 8, a.ccoef + (size_t)M.c * FCOEF_N + 2 * lane);
    };
    const int nw = gridDim.x * HOT_WARPS;
    int item = blockIdx.x * HOT_WARPS + wib;
    uint32_t q = 0;
    int gsel = 0;
    HotUpdMeta cur{}, nxt{};
    if (item < nc) {
        hot_upd_issue<NPRE>(a.cmeta2, item, lane, cur);
        hot_upd_commit<HAS_H, NPRE>(a, lane, gts, hts, gbs, hbs, gsbs, hsbs, fk_g, fk_h, NE, sptr, cur);
        if (cur.ne <= NE) {
            hot_stage<CI, NTOT, NSLOT, false>(sptr, cur.ne + NPRE - 1, 0, stages, lane);
            stage_records(cur, recs);
        }
        cp_async_commit();
    }
    while (item < nc) {
        const int nitem = item + nw;
        const bool has_next = nitem < nc;
        unsigned long long* sp_cur = sptr + gsel * HOT_PTRS;
        unsigned long long* sp_nxt = sptr + (gsel ^ 1) * HOT_PTRS;
        if (has_next) hot_upd_issue<NPRE>(a.cmeta2, nitem, lane, nxt);
        // called once, right before the last chunk of the current cell is computed
        auto stage_next_item = [&](double* st) {
            if (!has_next) return;
            hot_upd_commit<HAS_H, NPRE>(a, lane, gts, hts, gbs, hbs, gsbs, hsbs, fk_g, fk_h, NE, sp_nxt, nxt);
            if (nxt.ne <= NE) {
                hot_stage<CI, NTOT, NSLOT, false>(sp_nxt, nxt.ne + NPRE - 1, 0, st, lane);
                stage_records(nxt, recs + (gsel ^ 1) * P::REC_D);
            }
        };
        if (cur.ne > NE) {
            stage_next_item(stages + (q & 1) * P::STAGE_D);
            cp_async_commit();
            cur = nxt; item = nitem; gsel ^= 1;
            continue;
        }
        // FULL: all NE entries are internal faces (compile-time face predicates)
        // AXIS (implies FULL): every outward area vector has one component, so the flux coefficient xi.Sf of a y / z face
        // is a lane constant and that of an x face x * Sx: folded into the relaxation factor and the y/z part of the
        // equilibrium (A_j, B_j below), a face costs 6 instead of 8 FP64 instructions per point
        auto run_item = [&](auto full_c, auto axis_c) {
        constexpr bool FULL = decltype(full_c)::value;
        constexpr bool AXIS = decltype(axis_c)::value;
        const int ne = FULL ? NE : cur.ne, nint = FULL ? NE : cur.nint, c = cur.c;
        // interior cells stage from register offsets (16-byte units), see hot_stage_off
        const bool interior = FULL;
        const uint32_t offc = (uint32_t)c * (uint32_t)(blk / 2) + (uint32_t)lane;
        uint32_t offf[NE];
#pragma unroll
        for (int j = 0; j < NE; j++) offf[j] = (uint32_t)__shfl_sync(0xffffffffu, cur.face, j) * (uint32_t)(blk / 2) + (uint32_t)lane;
        auto stage_interior = [&](int ch, double* st) {
            const uint32_t sdst = smem_u32(st) + (uint32_t)lane * 16u, coff = (uint32_t)ch * (CI * 256u);
            const bool ahead = ch + 1 < nchunk;
#pragma unroll
            for (int fld = 0; fld < P::NFLD; fld++) {
                // array base + chunk offset once per chunk and array, kept opaque: a stream then costs one 64-bit
                // multiply-add and its LDGSTS (folded into every stream it was three integer instructions each)
                unsigned long long bc = (unsigned long long)(fld ? hts : gts) + coff;
                unsigned long long bf = (unsigned long long)(fld ? fk_h : fk_g) + coff;
                asm volatile("" : "+l"(bc));
                asm volatile("" : "+l"(bf));
                hot_stage_at<CI>(sdst + (fld * NSLOT + 0) * (CI * 256), bc + (unsigned long long)offc * 16u, RLX_POL(a.pol_ef), ahead);
                if (WMODE == 0) {
                    unsigned long long bb = (unsigned long long)(fld ? hbs : gbs) + coff;
                    asm volatile("" : "+l"(bb));
                    hot_stage_at<CI>(sdst + (fld * NSLOT + 1) * (CI * 256), bb + (unsigned long long)offc * 16u, RLX_POL(a.pol_ef), ahead);
                }
#pragma unroll
                for (int j = 0; j < NE; j++)   // every face block is read by two cells: keep it in L2 for the second one
                    hot_stage_at<CI>(sdst + (fld * NSLOT + NPRE + j) * (CI * 256), bf + (unsigned long long)offf[j] * 16u, RLX_POL(a.pol_el), ahead);
            }
        };
        // outward area vectors (sign folded in) and the face equilibria of the internal faces
        // the records of this cell were staged together with its first chunk (the only group in flight)
        cp_async_wait<0>();
        __syncwarp();
        const double* rec_f = recs + gsel * P::REC_D;
        const double* rec_s = rec_f + NE * FCOEF_N;
        double Sx[NE], cyz[NE], EYZ[NE], YZ2[NE], QYZ[NE];
#pragma unroll
        for (int j = 0; j < NE; j++) {
            Sx[j] = 0.0; cyz[j] = 0.0; EYZ[j] = YZ2[j] = QYZ[j] = 0.0;
            if (j < ne) {
                const double* S = rec_s + j * 4;
                Sx[j] = S[0];
                cyz[j] = fma(y, S[1], z * S[2]);
            }
            if (j < nint) {
                const double* fc = rec_f + j * FCOEF_N;
                const double Ux = fc[0], Uy = fc[1], Uz = fc[2], ia = fc[3], pre = fc[4];
                const double qx = fc[5], qy = fc[6], qz = fc[7];
                for (int tt = lane; tt < span; tt += 32) {
                    const double cx = txs[(tmin + tt) * 6 + 5] - Ux;
                    const double x2 = cx * cx * ia;
                    double* xt = xtab + ((size_t)j * TW + tt) * 4;
                    xt[0] = exp(-0.5 * x2); xt[1] = x2; xt[2] = cx * qx; xt[3] = 0.0;
                }
                const double cy = y - Uy, cz = z - Uz;
                const double yz2 = (cy * cy + cz * cz) * ia;
                EYZ[j] = pre * exp(-0.5 * yz2);
                YZ2[j] = yz2 - a.gas.D - 2.0;
                QYZ[j] = cy * qy + cz * qz;
                if (lane == 0) { unic[j * 2] = fc[8]; unic[j * 2 + 1] = fc[9]; }
            }
        }
        double Aj[NE], Bj[NE];
        if (AXIS) {
#pragma unroll
            for (int j = 0; j < NE; j++) {
                const double cj = (j < 2) ? Sx[j] : cyz[j];
                Aj[j] = cj * rec_f[j * FCOEF_N + 8];       // coefficient * (1 - rf) of the face
                Bj[j] = cj * EYZ[j];
            }
        }
        // WMODE 2: half-step table of the cell itself (the operations of k_hot_halfstep)
        double cEYZ = 0.0, cYZ2 = 0.0, cQYZ = 0.0, comrf = 0.0, cRT = 0.0, cqx = 0.0;
        if (WMODE == 2) {
            const double* rc = rec_s + NE * 4;                            // Ux Uy Uz a pre qx qy qz omrf RT (k_cell_coef)
            for (int tt = lane; tt < span; tt += 32) {
                const double cx = txs[(tmin + tt) * 6 + 5] - rc[0];
                const double x2 = cx * cx * rc[3];
                double* xt = ctab + tt * 4;
                xt[0] = exp(-0.5 * x2); xt[1] = x2; xt[2] = cx * rc[5]; xt[3] = 0.0;
            }
            const double cy = y - rc[1], cz = z - rc[2];
            const double yz2 = (cy * cy + cz * cz) * rc[3];
            cEYZ = rc[4] * exp(-0.5 * yz2);
            cYZ2 = yz2 - a.gas.D - 2.0;
            cQYZ = fma(-rc[0], rc[5], cy * rc[6] + cz * rc[7]);           // (xi - U).q = x qx + cQYZ, as the pencils form it
            cqx = rc[5];
            comrf = rc[8];
            cRT = rc[9];
        }
        __syncwarp();
        const double dtv = a.dt / a.m.V[c];
        double A[4] = {0, 0, 0, 0}, B[2] = {0, 0};
        double* gdst = gts + (size_t)c * blk + lane;
        double* hdst = HAS_H ? hts + (size_t)c * blk + lane : nullptr;
        for (int ch = 0; ch < nchunk; ch++) {
            {
                double* st = stages + ((q & 1) ^ 1) * P::STAGE_D;
                if (ch + 1 < nchunk) {
                    if (interior) stage_interior(ch + 1, st);
                    else hot_stage<CI, NTOT, NSLOT, false>(sp_cur, ne + NPRE - 1, ch + 1, st, lane);
                } else stage_next_item(st);
            }
            cp_async_commit();
            cp_async_wait<1>();
            __syncwarp();
            const double* sg = stages + (q & 1) * P::STAGE_D + lane;
            const int i0 = ch * CI, tb = cb + i0;
            double xx[CI], W[CI][4];
#pragma unroll
            for (int u = 0; u < CI; u++) {
                const double2 t0 = lds2(txs + (tb + u) * 6), t1 = lds2(txs + (tb + u) * 6 + 2), t2 = lds2(txs + (tb + u) * 6 + 4);
                W[u][0] = t0.y; W[u][1] = t1.x; W[u][2] = t1.y; W[u][3] = t2.x; xx[u] = t2.y;
            }
#pragma unroll
            for (int fld = 0; fld < P::NFLD; fld++) {
                const double* sf = sg + fld * NSLOT * CI * 32;
                double sum[CI], sum2[CI];
                if (AXIS) {
                    // x faces: sx = sum_j Sx_j g_f (times x at the end); y / z faces straight into the flux sum;
                    // relaxed and equilibrium parts in separate chains
                    double sxa[CI], sxb[CI];
#pragma unroll
                    for (int u = 0; u < CI; u++) sum[u] = sum2[u] = sxa[u] = sxb[u] = 0.0;
#pragma unroll
                    for (int j = 0; j < NE; j++) {
                        const double frt = fld ? unic[j * 2 + 1] : 0.0;
#pragma unroll
                        for (int u = 0; u < CI; u++) {
                            const double* xt = xtab + ((size_t)j * TW + (tb + u - tmin)) * 4;
                            const double2 x01 = lds2(xt);
                            const double cc = x01.y + YZ2[j];
                            const double cq = xt[2] + QYZ[j];
                            double e;
                            if (fld == 0) e = fma(cq, cc, 1.0) * x01.x;                                // :1042 without the y/z factor
                            else e = (kd + cq * ((cc + 2.0) * kd - 2.0 * a.gas.K)) * frt * x01.x;      // :1043
                            const double f = sf[((NPRE + j) * CI + u) * 32];
                            if (j < 2) { sxa[u] = fma(Aj[j], f, sxa[u]); sxb[u] = fma(Bj[j], e, sxb[u]); }   // :880-881, :952-955
                            else { sum[u] = fma(Aj[j], f, sum[u]); sum2[u] = fma(Bj[j], e, sum2[u]); }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < CI; u++) sum[u] = fma(xx[u], sxa[u] + sxb[u], sum[u]);
                } else {
                // two partial sums per point (even / odd entries): half the depth of the dependent FMA chain
#pragma unroll
                for (int u = 0; u < CI; u++) sum[u] = sum2[u] = 0.0;
#pragma unroll
                for (int j = 0; j < NE; j++) {
                    if (j < nint) {
                        const double omrf = unic[j * 2], frt = unic[j * 2 + 1];
#pragma unroll
                        for (int u = 0; u < CI; u++) {
                            const double* xt = xtab + ((size_t)j * TW + (tb + u - tmin)) * 4;
                            const double2 x01 = lds2(xt);
                            const double cc = x01.y + YZ2[j];
                            const double cq = xt[2] + QYZ[j];
                            const double gM = x01.x * EYZ[j];
                            double eq;
                            if (fld == 0) eq = fma(cq, cc, 1.0) * gM;                                  // :1042
                            else eq = (kd + cq * ((cc + 2.0) * kd - 2.0 * a.gas.K)) * gM * frt;        // :1043
                            const double gf = fma(omrf, sf[((NPRE + j) * CI + u) * 32], eq);          // :880-881
                            double& acc = (j & 1) ? sum2[u] : sum[u];
                            acc = fma(fma(xx[u], Sx[j], cyz[j]), gf, acc);                             // :952-955
                        }
                    } else if (j < ne) {
#pragma unroll
                        for (int u = 0; u < CI; u++) {
                            double& acc = (j & 1) ? sum2[u] : sum[u];
                            acc = fma(fma(xx[u], Sx[j], cyz[j]), sf[((NPRE + j) * CI + u) * 32], acc);
                        }
                    }
                }
                }
                // new values of the whole chunk first (one basic block: the chains of its points interleave), stores and
                // moment sums after
                double vnew[CI];
#pragma unroll
                for (int u = 0; u < CI; u++) {
                    double wcell;
                    if (WMODE == 1) wcell = sf[u * 32];
                    else if (WMODE == 0) wcell = hot_w_combine(sf[u * 32], sf[(CI + u) * 32]);
                    else {
                        const double* xt = ctab + (size_t)(tb + u - tmin) * 4;
                        const double2 x01 = lds2(xt);
                        const double cc = x01.y + cYZ2, cq = fma(xx[u], cqx, cQYZ), gM = x01.x * cEYZ;
                        const double t_i = sf[u * 32];
                        const double b_i = fld == 0 ? fma(comrf, t_i, fma(cq, cc, 1.0) * gM)                                      // :405,1042
                                                    : fma(comrf, t_i, (kd + cq * ((cc + 2.0) * kd - 2.0 * a.gas.K)) * gM * cRT);   // :406,1043
                        wcell = hot_w_combine(t_i, b_i);
                    }
                    vnew[u] = fma(-(sum[u] + sum2[u]), dtv, wcell);                                    // :937,952
                }
#pragma unroll
                for (int u = 0; u < CI; u++) {
                    if (i0 + u >= Ln) continue;   // tail chunk (warp-uniform)
                    __stcs((fld == 0 ? gdst : hdst) + (i0 + u) * 32, vnew[u]);
                    if (fld == 0) {
                        A[0] = fma(W[u][0], vnew[u], A[0]); A[1] = fma(W[u][1], vnew[u], A[1]);
                        A[2] = fma(W[u][2], vnew[u], A[2]); A[3] = fma(W[u][3], vnew[u], A[3]);
                    } else {
                        B[0] = fma(W[u][0], vnew[u], B[0]); B[1] = fma(W[u][1], vnew[u], B[1]);
                    }
                }
            }
            __syncwarp();
            q++;
        }
        double vv[16];
        expand_g(A, wr, y, z, vv);
        double uu[NM_H] = {0, 0, 0, 0};
        if (HAS_H) expand_h(B, wr, y, z, uu);
        vv[13] = uu[0]; vv[14] = uu[1]; vv[15] = uu[2];
        __syncwarp();
        const double tot = warp_reduce16_smem(vv, xtab, lane);
        if (lane < 16 && lane < nm) atomicAdd(a.cslot + (size_t)c * nm + lane, tot);
        if (HAS_H) {
            const double t3 = warp_sum(uu[3]);
            if (lane == 0) atomicAdd(a.cslot + (size_t)c * nm + 16, t3);
        }
        };
        if (cur.ne == NE && cur.nint == NE) {
            if (cur.cls == 1) run_item(std::true_type{}, std::true_type{});
            else run_item(std::true_type{}, std::false_type{});
        } else run_item(std::false_type{}, std::false_type{});
        cur = nxt; item = nitem; gsel ^= 1;
    }
    cp_async_wait<0>();
}

// -------------------------------------------------------------------------------------------------
// stage 1, gTilde -> gBarP (discreteVelocity.C:346-410), as a pure stream: persistent warps, the whole
// row block of the NEXT cell (and its macro record) is in flight while the current cell's equilibrium
// tables are built and applied; output goes straight to global memory.
template <bool HAS_H>
struct HotHalfPlan {
    static constexpr int NFLD = HAS_H ? 2 : 1;
    static __host__ __device__ size_t stage_d(int L) { return (size_t)NFLD * L * 32; }                 // doubles
    static __host__ __device__ size_t per_warp(int L, int tw) {
        return ((2 * stage_d(L) + 2 * 16 /*macro records*/ + 4 * (size_t)tw) * 8 + 127) / 128 * 128;
    }
    static __host__ size_t total(int L, int tw, int ntab) { return ((size_t)ntab * 8 + 127) / 128 * 128 + HOT_WARPS * per_warp(L, tw); }
};

template <bool HAS_H>
__global__ void __launch_bounds__(HOT_WARPS * 32, 4)
k_hot_halfstep(StepArgs a, int tw, int wmode /* also leave w = -1/3 gTilde + 4/3 gBarP in place of gTilde */,
               const int* list /* cells to convert, or null: all */, int nlist) {
    using P = HotHalfPlan<HAS_H>;
    extern __shared__ __align__(128) unsigned char dyn[];
    const DevDV& dv = a.dv;
    const int L = dv.L, nc = a.m.nc, blk = L * 32;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double* txs = reinterpret_cast<double*>(dyn);                    // abscissae
    for (int k = threadIdx.x; k < dv.ntab; k += blockDim.x) txs[k] = dv.tx[k];
    unsigned char* wbase = dyn + ((size_t)dv.ntab * 8 + 127) / 128 * 128 + wib * P::per_warp(L, tw);
    const size_t stage_d = P::stage_d(L);
    double* stages = reinterpret_cast<double*>(wbase);               // [2][NFLD][L][32]
    double* mrec = stages + 2 * stage_d;                             // [2][16] cell macro records (9 used)
    double* xtab = mrec + 2 * 16;                                    // [tw][4]
    __syncthreads();
    const size_t slab_c = (size_t)a.slab * nc * blk;
    const double* gts = a.gt + slab_c;
    const double* hts = HAS_H ? a.ht + slab_c : nullptr;
    double* gbs = a.gb + slab_c;
    double* hbs = HAS_H ? a.hb + slab_c : nullptr;
    const int grow = a.slab * 32 + lane;
    const double y = dv.row_y[grow], z = dv.row_z[grow];
    const int cb = dv.row_cbase[grow];
    const int Ln = dv_len(dv, a.slab);                               // points per row of THIS slab (strides use L)
    int tmin, span;
    table_range(dv, cb, tmin, span, Ln);
    const double kd = (double)(a.gas.K + 3 - a.gas.D);
    const int npiece = Ln * 16;                                      // 16-byte pieces per field block
    const unsigned long long pol_ef = l2_evict_first_policy();

    auto stage_cell = [&](int c, int buf) {
        const uint32_t sd = smem_u32(stages + buf * stage_d);
#pragma unroll
        for (int fld = 0; fld < P::NFLD; fld++) {
            const char* src = reinterpret_cast<const char*>((fld ? hts : gts) + (size_t)c * blk);
            for (int p = lane; p < npiece; p += 32) cp_async16_ef(sd + (fld * npiece + p) * 16, src + p * 16, pol_ef);
        }
        if (lane < MAC_N)   // 72-byte records are only 8-byte aligned
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(mrec + buf * 16 + lane)),
                         "l"(a.cmac + (size_t)c * MAC_N + lane));
    };
    const int nw = gridDim.x * HOT_WARPS;
    const int nitems = list ? nlist : nc;
    int idx = blockIdx.x * HOT_WARPS + wib, buf = 0;
    if (idx < nitems) stage_cell(list ? list[idx] : idx, 0);
    cp_async_commit();
    for (; idx < nitems; idx += nw, buf ^= 1) {
        const int c = list ? list[idx] : idx;
        if (idx + nw < nitems) stage_cell(list ? list[idx + nw] : idx + nw, buf ^ 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncwarp();
        const double* mc = mrec + buf * 16;
        const double rf = 1.5 * a.dt / (2.0 * mc[5] + a.dt);         // discreteVelocity.C:393
        const EqCoef e = make_eq(a.gas, mc, rf);
        for (int tt = lane; tt < span; tt += 32) {
            const double cx = txs[tmin + tt] - e.Ux;
            const double x2 = cx * cx * e.a;
            double* xt = xtab + tt * 4;
            xt[0] = exp(-0.5 * x2); xt[1] = x2; xt[2] = cx * e.qx; xt[3] = 0.0;
        }
        const double cy = y - e.Uy, cz = z - e.Uz;
        const double yz2 = (cy * cy + cz * cz) * e.a;
        const double EYZ = e.pre * exp(-0.5 * yz2);
        const double YZ2 = yz2 - a.gas.D - 2.0;
        const double QYZ = cy * e.qy + cz * e.qz;
        const double omrf = 1.0 - rf;
        __syncwarp();
        const double* sg = stages + buf * stage_d + lane;
        double* dg = gbs + (size_t)c * blk + lane;
        double* dh = HAS_H ? hbs + (size_t)c * blk + lane : nullptr;
        double* wg = a.gt + slab_c + (size_t)c * blk + lane;
        double* wh = HAS_H ? a.ht + slab_c + (size_t)c * blk + lane : nullptr;
        const double* xt0 = xtab + (cb - tmin) * 4;
#pragma unroll 4
        for (int i = 0; i < Ln; i++) {
            const double2 x01 = lds2(xt0 + i * 4);
            const double cc = x01.y + YZ2;                           // cSqrByRT - D - 2
            const double cq = xt0[i * 4 + 2] + QYZ;                  // (1-Pr) cqBy5pRT
            const double gM = x01.x * EYZ;                           // rf * gEqBGK
            const double gt_i = sg[i * 32];
            const double gb_i = fma(omrf, gt_i, fma(cq, cc, 1.0) * gM);                                          // :405,1042
            __stcs(dg + i * 32, gb_i);
            if (wmode) __stcs(wg + i * 32, hot_w_combine(gt_i, gb_i));
            if (HAS_H) {
                const double ht_i = sg[(Ln + i) * 32];
                const double hb_i = fma(omrf, ht_i, (kd + cq * ((cc + 2.0) * kd - 2.0 * a.gas.K)) * gM * e.RT);  // :406,1043
                __stcs(dh + i * 32, hb_i);
                if (wmode) __stcs(wh + i * 32, hot_w_combine(ht_i, hb_i));
            }
        }
        __syncwarp();   // stage and tables are free again
    }
    cp_async_wait<0>();
}

// -------------------------------------------------------------------------------------------------
// stage 2.3 (wall incoming Maxwellian, discreteVelocity.C:693-731) and stage 4 on boundary faces
// (:886-931) with the separable equilibrium tables: one exp per table entry and lane instead of two per
// (face, DV).  Same rules as k_bnd_relax (dugks_kernels.cuh).  item = boundary face.
template <bool HAS_H>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
k_hot_bnd_relax(StepArgs a, int tw) {
    extern __shared__ __align__(16) unsigned char dyn[];
    const DevDV& dv = a.dv;
    double* txs = reinterpret_cast<double*>(dyn);                          // abscissae
    for (int k = threadIdx.x; k < dv.ntab; k += blockDim.x) txs[k] = dv.tx[k];
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double* xtab = txs + ((dv.ntab + 1) & ~1) + (size_t)wib * tw * 4;     // [tw][4]: EX, X2, QX, EXwall
    const int L = dv.L, blk = L * 32;
    const double hstep = 0.5 * a.dt;
    const int grow = a.slab * 32 + lane;
    const double y = dv.row_y[grow], z = dv.row_z[grow];
    const int cb = dv.row_cbase[grow];
    const int Ln = dv_len(dv, a.slab);
    int tmin, span;
    table_range(dv, cb, tmin, span, Ln);
    const double kd = (double)(a.gas.K + 3 - a.gas.D);
    for (int b = blockIdx.x * WARPS_PER_CTA + wib; b < a.m.nbf; b += gridDim.x * WARPS_PER_CTA) {
        const int kind = a.m.b_kind[b];
        const double sx = a.m.b_Sf[(size_t)b * 3], sy = a.m.b_Sf[(size_t)b * 3 + 1], sz = a.m.b_Sf[(size_t)b * 3 + 2];
        const double* bm = a.bmac + (size_t)b * 5;
        const double* mf = a.fmac + (size_t)(a.m.nif + b) * MAC_N;
        const double rf = hstep / (2.0 * mf[5] + hstep);                   // discreteVelocity.C:867
        const EqCoef e = make_eq(a.gas, mf, rf);
        const double omrf = 1.0 - rf;
        const bool wall = kind == K_MAXWELL_WALL;
        // wall Maxwellian rho_w / (2 pi R T_w)^(D/2) exp(-|xi - U_w|^2 / 2 R T_w), :1063-1075
        const double RTw = a.gas.R * bm[4], aw = 1.0 / RTw;
        double prew = 0.0;
        if (wall) {
            const double sq = sqrt(2.0 * DUGKS_PI * RTw);
            prew = bm[0] / ((a.gas.D == 3) ? sq * sq * sq : ((a.gas.D == 2) ? sq * sq : sq));
        }
        for (int tt = lane; tt < span; tt += 32) {
            const double xv = txs[tmin + tt];
            const double cx = xv - e.Ux, x2 = cx * cx * e.a;
            double* xt = xtab + tt * 4;
            xt[0] = exp(-0.5 * x2); xt[1] = x2; xt[2] = cx * e.qx;
            const double cw = xv - bm[1];
            xt[3] = wall ? exp(-0.5 * cw * cw * aw) : 0.0;
        }
        const double cy = y - e.Uy, cz = z - e.Uz;
        const double yz2 = (cy * cy + cz * cz) * e.a;
        const double EYZ = e.pre * exp(-0.5 * yz2), YZ2 = yz2 - a.gas.D - 2.0, QYZ = cy * e.qy + cz * e.qz;
        double EYZw = 0.0;
        if (wall) {
            const double wy = y - bm[2], wz = z - bm[3];
            EYZw = prew * exp(-0.5 * (wy * wy + wz * wz) * aw);
        }
        const double hfw = RTw * kd;
        const double ySy = __dmul_rn(y, sy), zSz = __dmul_rn(z, sz);
        __syncwarp();
        const size_t bbase = ((size_t)a.slab * a.m.nbf + b) * blk + lane;
        const double* xt0 = xtab + (cb - tmin) * 4;
        for (int i = 0; i < Ln; i++) {
            const double phi = __dadd_rn(__dadd_rn(__dmul_rn(txs[cb + i], sx), ySy), zSz);
            const double2 x01 = lds2(xt0 + i * 4), x23 = lds2(xt0 + i * 4 + 2);
            double g = a.gsb[bbase + (size_t)i * 32];
            double hh = HAS_H ? a.hsb[bbase + (size_t)i * 32] : 0.0;
            if (wall && phi <= 0) {                                       // :713-727
                g = x23.y * EYZw;
                hh = g * hfw;
            }
            const double cc = x01.y + YZ2, cq = x23.x + QYZ, gM = x01.x * EYZ;
            const double gS = fma(cq, cc, 1.0) * gM;
            const double hS = HAS_H ? (kd + cq * ((cc + 2.0) * kd - 2.0 * a.gas.K)) * gM * e.RT : 0.0;
            if (kind == K_SYMMETRY_PLANE) { g = fma(omrf, g, gS); hh = fma(omrf, hh, hS); }   // :880-881 [OF-lib]
            if (phi > 0) { g = fma(omrf, g, gS); hh = fma(omrf, hh, hS); }                    // :907-919
            if (kind == K_DVM_SYMMETRY) { g = fma(omrf, g, gS); hh = fma(omrf, hh, hS); }     // :922-930
            a.gsb[bbase + (size_t)i * 32] = g;
            if (HAS_H) a.hsb[bbase + (size_t)i * 32] = hh;
        }
        __syncwarp();
    }
}

// -------------------------------------------------------------------------------------------------
// Wall constants with the separable Maxwellian (same sums as k_wall_constants, dugks_kernels.cuh;
// fvDVM.C:263-309): moments of the incoming half-space Maxwellian per unit rho_w and inComingByRho.
// Recomputed whenever the caller changes the wall velocity / temperature, so it is on the
// end-to-end path of a time-varying boundary.
template <bool HAS_H>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
k_hot_wall_constants(StepArgs a, double* cin, double* win, int tw) {
    extern __shared__ __align__(16) unsigned char dyn[];
    const DevDV& dv = a.dv;
    double* txs = reinterpret_cast<double*>(dyn);                         // [5][ntab]
    for (int k = threadIdx.x; k < 5 * dv.ntab; k += blockDim.x) txs[k] = dv.tx[k];
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    double* xtab = txs + 5 * dv.ntab + (size_t)wib * tw;                  // [tw] exp(-(x - U_x)^2 / 2RT)
    const int nt = dv.ntab;
    const int grow = a.slab * 32 + lane;
    const double y = dv.row_y[grow], z = dv.row_z[grow], wr = dv.row_w[grow];
    const int cb = dv.row_cbase[grow];
    const int Ln = dv_len(dv, a.slab);
    int tmin, span;
    table_range(dv, cb, tmin, span, Ln);
    const int nm = a.nm;
    for (int b = blockIdx.x * WARPS_PER_CTA + wib; b < a.m.nbf; b += gridDim.x * WARPS_PER_CTA) {
        if (a.m.b_kind[b] != K_MAXWELL_WALL) continue;   // warp-uniform
        const double sx = a.m.b_Sf[(size_t)b * 3], sy = a.m.b_Sf[(size_t)b * 3 + 1], sz = a.m.b_Sf[(size_t)b * 3 + 2];
        const double* bm = a.bmac + (size_t)b * 5;
        const double RT = a.gas.R * bm[4], aw = 1.0 / RT;
        const double sq = sqrt(2.0 * DUGKS_PI * RT);
        const double p = (a.gas.D == 3) ? sq * sq * sq : ((a.gas.D == 2) ? sq * sq : sq);
        for (int tt = lane; tt < span; tt += 32) {
            const double cw = txs[tmin + tt] - bm[1];
            xtab[tt] = exp(-0.5 * cw * cw * aw);
        }
        const double wy = y - bm[2], wz = z - bm[3];
        const double EYZ = exp(-0.5 * (wy * wy + wz * wz) * aw) / p;
        const double hfac = RT * (a.gas.K + 3 - a.gas.D);
        const double ySy = __dmul_rn(y, sy), zSz = __dmul_rn(z, sz);
        __syncwarp();
        double A[4] = {0, 0, 0, 0}, B[2] = {0, 0}, inb = 0.0;
        for (int i = 0; i < Ln; i++) {
            const int t = cb + i;
            const double phi = __dadd_rn(__dadd_rn(__dmul_rn(txs[t], sx), ySy), zSz);
            if (phi <= 0) {                               // discreteVelocity.C:713
                const double M = xtab[t - tmin] * EYZ;
                A[0] = fma(txs[nt + t], M, A[0]);
                A[1] = fma(txs[2 * nt + t], M, A[1]);
                A[2] = fma(txs[3 * nt + t], M, A[2]);
                A[3] = fma(txs[4 * nt + t], M, A[3]);
                if (HAS_H) {
                    B[0] = fma(txs[nt + t], M * hfac, B[0]);
                    B[1] = fma(txs[2 * nt + t], M * hfac, B[1]);
                }
                if (phi < 0) inb += -(txs[nt + t] * wr) * phi * M;   // fvDVM.C:291-299
            }
        }
        double v[16];
        expand_g(A, wr, y, z, v);
        double u[NM_H] = {0, 0, 0, 0};
        if (HAS_H) expand_h(B, wr, y, z, u);
        v[13] = u[0]; v[14] = u[1]; v[15] = u[2];
        const double tot = warp_reduce16(v, lane);
        const double t3 = warp_sum(u[3]);
        const double tin = warp_sum(inb);
        const int idx = reduce16_index(lane);
        if ((lane & 1) == 0 && idx < nm) cin[(size_t)b * nm + idx] += tot;
        if (HAS_H && lane == 0) cin[(size_t)b * nm + 16] += t3;
        if (lane == 0) win[b] += tin;
        __syncwarp();
    }
}
