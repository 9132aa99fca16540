// dugks_pencil.cuh — phase 1 of the step for axis-aligned interior cells of a 3-D mesh, as a CTA-level PENCIL:
// the four warps of a CTA take a 2 x 2 bundle of x-lines of cells and advance along x together, one cell per
// warp and step, behind ONE barrier per step.
//
// Why: with one independent warp per cell (dugks_hot.cuh) every row block (L x 32 doubles,
// 7 KB at L = 28) is fetched from L2 by seven warps - its own and the six neighbours - and phase 1 runs at the
// L2 -> SM bandwidth, not at the DRAM rate.  Here
//   * the blocks of a line's cells live in a shared-memory WINDOW of three x positions (previous, current,
//     next): the x neighbours of a cell are the window's other two slots, the y / z neighbours inside the
//     bundle are the other warps' current slots.  A block is fetched ONCE per CTA (cp.async, 16 B per lane) and
//     then read from shared memory by up to five (cell, neighbour) uses;
//   * only the two neighbours outside the bundle (the "halo": one in y, one in z) are streamed chunk by chunk,
//     so a cell costs 3 block reads from L2 instead of 7;
//   * FUSE: the window is filled with gTilde and the half step gBarP = (1 - rf) gTilde + rf gS
//     (discreteVelocity.C:393-406) is applied in shared memory when a block enters the window (once per cell by
//     the warp that owns the line; on the fly for the two halo blocks), so gBarP is never written to or read
//     from global memory: the half-step kernel and its 16 B per update disappear for these cells (the update
//     kernel forms w = -1/3 gTilde + 4/3 gBarP from gTilde and the cell's equilibrium itself).
// The arithmetic per (cell, velocity) is that of hot_axis_item (same operations; only the order in which the
// two y (z) neighbours enter the gradient sum may differ), the upwind sets are the static codes of
// k_build_upwind, face values go to the slab's face storage, moments to the face slots.
#pragma once
#include "dugks_hot.cuh"

#define PEN_WARPS 4      // 2 x 2 bundle: line l = by + 2 bz
#define PEN_CI 4         // points per staged chunk (1 KB per block and chunk)
#define PEN_CU 2         // points advanced together
#define PEN_NE 6
#ifndef PEN_CTXS
#define PEN_CTXS 1       // per-point row constants of the stencil loop from the constant bank (PenArgs::tx6) instead of shared memory
#endif

// One work item: `nsteps` consecutive x positions of a bundle.
//   cells: offset into pen_cells of [(nsteps + 2)][4] cell ids, positions -1 .. nsteps of the four lines
//   halo : offset into pen_halo of [nsteps][4][2]: the y and the z neighbour outside the bundle
//   item0: traversal item of (step 0, line 0); (step k, line l) is item0 + 4 k + l (cmeta / upwind codes)
struct PenItem { int cells, halo, item0, nsteps; };

#define PEN_TX6 ((32 + HOT_CI_MAX) * 6)
struct PenArgs {
    const PenItem* items;
    int nitems;
    const int* cells;
    const int* halo;
    // per-point constants of a row {-dt/2 x, w, w x, w x^2, w x^3, x} (pencils run on unchunked rows: at most 32 points):
    // kernel parameters live in the constant bank, so the stencil loop reads them without a shared-memory instruction
    double tx6[PEN_TX6];
};

// shared-memory plan (bytes), L = points per row
struct PenPlan {
    static __host__ __device__ size_t txs_bytes(int ntab) { return ((size_t)(ntab + HOT_CI_MAX) * 48 + 127) / 128 * 128; }
    static __host__ __device__ size_t win_bytes(int L) { return (size_t)3 * PEN_WARPS * L * 32 * 8; }
    // per warp: halo stages [2][2][CI][32], geometry records [2][7 * 6], half-step coefficient records
    // [2][3][FCOEF_N], tables [3][tw][2]  (tw: table entries a warp spans, DevDV::tabw)
    static __host__ __device__ size_t warp_bytes(int tw) {
        return ((size_t)(2 * 2 * PEN_CI * 32 + 2 * (1 + PEN_NE) * 6 + 2 * 3 * FCOEF_N + 3 * tw * 2) * 8 + 127) / 128 * 128;
    }
    static __host__ size_t total(int L, int ntab, int tw) { return txs_bytes(ntab) + win_bytes(L) + PEN_WARPS * warp_bytes(tw); }
};

// Sum of 13 values per lane over the warp through 16 x 13 doubles of scratch; lanes 0..12 return the totals.
__device__ __forceinline__ double pen_reduce13(double (&v)[NM_G], double* scratch, int lane) {
#pragma unroll
    for (int k = 0; k < NM_G; k++) v[k] += __shfl_xor_sync(0xffffffffu, v[k], 16);
    if (lane < 16) {
#pragma unroll
        for (int k = 0; k < NM_G; k++) scratch[lane * NM_G + k] = v[k];
    }
    __syncwarp();
    const int col = lane < NM_G ? lane : (lane < 2 * NM_G ? lane - NM_G : 0), half = lane >= NM_G ? 1 : 0;
    const double* src = scratch + (half * 8) * NM_G + col;
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int rr = 0; rr < 8; rr += 2) {
        s0 += src[rr * NM_G];
        s1 += src[(rr + 1) * NM_G];
    }
    double t = s0 + s1;
    t += __shfl_down_sync(0xffffffffu, t, NM_G);
    __syncwarp();
    return t;
}

// coefficients of one half-step conversion table (per lane: the y/z factor of the separable equilibrium)
struct PenEq {
    double EYZ, YZ2, QYZ, omrf, qx;   // QYZ = (y - Uy) qy + (z - Uz) qz - Ux qx: (xi - U).q = x qx + QYZ is one FMA per point
};

template <bool FUSE>
__global__ void __launch_bounds__(PEN_WARPS * 32, 2)
k_pencil_phase1(StepArgs a, PenArgs P) {
    extern __shared__ __align__(128) unsigned char dyn[];
    const DevDV& dv = a.dv;
    const int L = dv.L, nc = a.m.nc, blk = L * 32;
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const int by = wl & 1, bz = wl >> 1;
    const double hd = -0.5 * a.dt;
    double* txs = reinterpret_cast<double*>(dyn);
    hot_fill_txs(dv, hd, txs);
    double* win = reinterpret_cast<double*>(dyn + PenPlan::txs_bytes(dv.ntab));                 // [3][4][blk]
    const int TW = dv.tabw;
    double* wbase = reinterpret_cast<double*>(dyn + PenPlan::txs_bytes(dv.ntab) + PenPlan::win_bytes(L) + wl * PenPlan::warp_bytes(TW));
    double* halo = wbase;                                   // [2 stages][2 (y, z)][CI][32]
    double* geo = halo + 2 * 2 * PEN_CI * 32;               // [2][7 * 6]
    double* mrec = geo + 2 * (1 + PEN_NE) * 6;              // [2][3][FCOEF_N]: records of own(k+1), halo y(k), halo z(k)
    double* xtab = mrec + 2 * 3 * FCOEF_N;                  // [3][TW][2]: EX, X2
    __syncthreads();

    const size_t slab_c = (size_t)a.slab * nc * blk;
    const double* src = (FUSE ? a.gt : a.gb) + slab_c;      // gTilde (converted here) or gBarP (half-step kernel)
    const int grow = a.slab * 32 + lane;
    const double y = dv.row_y[grow], z = dv.row_z[grow], wr = dv.row_w[grow];
    const double yh = hd * y, zh = hd * z;
    const int cb = dv.row_cbase[grow];
    const int Ln = dv_len(dv, a.slab);
    const int nchunk = (Ln + PEN_CI - 1) / PEN_CI;
    int tmin = 0, span = 0;
    table_range(dv, cb, tmin, span, Ln);
    const int nm = a.nm;
    double* const fk_g = a.fkeep_g ? a.fkeep_g + (size_t)a.slab * a.m.nif * blk : nullptr;
    const bool keep_on = fk_g != nullptr;

    // block of cell c, rows [r0, r0 + nr), into dst (shared): 16 bytes per lane and instruction
    auto load_rows = [&](int c, int r0, int nr, double* dst) {
        const char* g = reinterpret_cast<const char*>(src + (size_t)c * blk + r0 * 32) + lane * 16;
        const uint32_t s = smem_u32(dst + r0 * 32) + lane * 16;
        for (int p = 0; p < nr * 256; p += 512) cp_async16(s + p, g + p);
    };
    auto load_chunk = [&](int c, int ch, double* dst /* block base */) {
        const char* g = reinterpret_cast<const char*>(src + (size_t)c * blk) + ch * (PEN_CI * 256) + lane * 16;
        const uint32_t s = smem_u32(dst) + ch * (PEN_CI * 256) + lane * 16;
#pragma unroll
        for (int part = 0; part < PEN_CI * 256 / 512; part++) cp_async16(s + part * 512, g + part * 512);
    };
    auto load_halo = [&](int cy, int cz, int ch, double* stage) {
        const uint32_t s = smem_u32(stage) + lane * 16;
        const char* gy = reinterpret_cast<const char*>(src + (size_t)cy * blk) + ch * (PEN_CI * 256) + lane * 16;
        const char* gz = reinterpret_cast<const char*>(src + (size_t)cz * blk) + ch * (PEN_CI * 256) + lane * 16;
#pragma unroll
        for (int part = 0; part < PEN_CI * 256 / 512; part++) {
            cp_async16(s + part * 512, gy + part * 512);
            cp_async16(s + PEN_CI * 256 + part * 512, gz + part * 512);
        }
    };
    // half-step coefficient record of cell c (k_cell_coef: Ux Uy Uz a pre qx qy qz omrf RT, 96 bytes) into dst
    auto load_mrec = [&](int c, double* dst, int l0) {
        if (lane >= l0 && lane < l0 + FCOEF_N / 2)
            cp_async16(smem_u32(dst + 2 * (lane - l0)), a.ccoef + (size_t)c * FCOEF_N + 2 * (lane - l0));
    };
    // half-step conversion table of the cell whose record is rc (discreteVelocity.C:393-406, :1033-1043)
    auto build_table = [&](const double* rc, double* xt, PenEq& E) {
        const double Ux = rc[0], ia = rc[3];
        for (int tt = lane; tt < span; tt += 32) {
            const double cx = txs[(tmin + tt) * 6 + 5] - Ux;
            const double x2 = cx * cx * ia;
            xt[tt * 2] = exp(-0.5 * x2);
            xt[tt * 2 + 1] = x2;
        }
        const double cy = y - rc[1], cz = z - rc[2];
        const double yz2 = (cy * cy + cz * cz) * ia;
        E.EYZ = rc[4] * exp(-0.5 * yz2);
        E.YZ2 = yz2 - a.gas.D - 2.0;
        E.QYZ = fma(-Ux, rc[5], cy * rc[6] + cz * rc[7]);
        E.omrf = rc[8];
        E.qx = rc[5];
    };
    // gBarP of one value (the half step of k_hot_halfstep; (xi - U).q formed as x qx + const): xt = table row of the point, x = its abscissa
    auto convert = [&](double raw, const double* xt, double x, const PenEq& E) {
        const double2 x01 = lds2(xt);
        const double cc = x01.y + E.YZ2;
        const double cq = fma(x, E.qx, E.QYZ);
        const double gM = x01.x * E.EYZ;
        return fma(E.omrf, raw, fma(cq, cc, 1.0) * gM);
    };

    uint32_t q = 0;   // flat chunk counter of this warp: halo stage = q & 1
    for (int it = blockIdx.x; it < P.nitems; it += gridDim.x) {
        const PenItem I = P.items[it];
        const int* ctab = P.cells + I.cells + wl;            // own line: ctab[(p + 1) * 4], p = -1 .. ns
        const int* htab = P.halo + I.halo + wl * 2;          // htab[k * 8 + {0, 1}]
        const int ns = I.nsteps;
        auto wslot = [&](int p, int line) { return win + ((size_t)((p + 3) % 3) * PEN_WARPS + line) * blk; };

        // ---- prologue: positions -1, 0, 1 of the own line, halo chunk 0 / geometry / macros of step 0
        __syncthreads();                                      // the previous item's readers are done with the window
        {
            const int cm1 = ctab[0], c0 = ctab[4], c1 = ctab[8];
            load_rows(cm1, 0, Ln, wslot(-1, wl));
            load_rows(c0, 0, Ln, wslot(0, wl));
            load_rows(c1, 0, Ln, wslot(1, wl));
            load_halo(htab[0], htab[1], 0, halo + (q & 1) * (2 * PEN_CI * 32));
            if (FUSE) {
                load_mrec(cm1, mrec, 0);                       // buffer 0: own(-1), own(0) for the prologue conversion
                load_mrec(c0, mrec + FCOEF_N, 6);
            }
        }
        HotMeta cur{}, nxt{};
        hot_meta_issue(a, I.item0 + wl, lane, cur);
        // cell ids one step ahead of their first use (read where they are needed, each is an exposed L2 round trip per
        // step or chunk: 10 % of the stall samples, profiles/r02/ncu48s/): halo cells of this step and of the next one,
        // own cell two positions ahead
        int hy_c = htab[0], hz_c = htab[1];
        int hy_nn = (1 < ns) ? ldg_early(htab + 8) : -1, hz_nn = (1 < ns) ? ldg_early(htab + 9) : -1;
        int own2_n = (2 <= ns) ? ldg_early(ctab + 3 * 4) : -1;
        cp_async_commit();
        cp_async_wait<0>();
        __syncwarp();
        if (FUSE) {
            // own(-1) and own(0) enter the window converted; own(1) is converted in step 0 like every later block
            PenEq E;
#pragma unroll 1
            for (int t = 0; t < 2; t++) {
                build_table(mrec + t * FCOEF_N, xtab, E);
                __syncwarp();
                double* blkp = wslot(t - 1, wl) + lane;
                const double* xt0 = xtab + (cb - tmin) * 2;
                for (int i = 0; i < Ln; i++) blkp[i * 32] = convert(blkp[i * 32], xt0 + i * 2, txs[(cb + i) * 6 + 5], E);
                __syncwarp();
            }
            // macros of step 0: own(1), halo(0); geometry of step 0
            load_mrec(ctab[8], mrec, 0);
            load_mrec(htab[0], mrec + FCOEF_N, 6);
            load_mrec(htab[1], mrec + 2 * FCOEF_N, 12);
        }
        hot_stage_geo(a.geo6 + (size_t)(cur.e0 + cur.c) * 6, PEN_NE, geo, lane);
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();                                      // every line's slot 0 is in place (and converted)

        for (int k = 0; k < ns; k++) {
            const int gsel = k & 1;
            const double* gb_ = geo + gsel * ((1 + PEN_NE) * 6);
            // ---- meta of this cell (issued one step ahead), of the next one
            const int own_k2m = own2_n;                                         // position k + 2 (window prefetch), -1 past the item
            const int hy_n = hy_nn, hz_n = hz_nn;                               // halo cells of step k + 1
            own2_n = (k + 3 <= ns) ? ldg_early(ctab + (k + 4) * 4) : -1;
            hy_nn = (k + 2 < ns) ? ldg_early(htab + (k + 2) * 8) : -1;
            hz_nn = (k + 2 < ns) ? ldg_early(htab + (k + 2) * 8 + 1) : -1;
            if (k + 1 < ns) hot_meta_issue(a, I.item0 + (k + 1) * 4 + wl, lane, nxt);
            {   // unpack the record of the current cell
                const bool valid = lane < PEN_NE;
                cur.own = valid ? (int)((unsigned)cur.face >> 31) : 0;
                cur.face = valid ? (cur.face & 0x7fffffff) : 0;
            }
            // ---- half-step tables of the blocks converted in this step: own(k + 1), halo y(k), halo z(k)
            PenEq Exp{}, Ehy{}, Ehz{};
            if (FUSE) {
                const double* mr = mrec + gsel * (3 * FCOEF_N);
                build_table(mr, xtab, Exp);
                build_table(mr + FCOEF_N, xtab + TW * 2, Ehy);
                build_table(mr + 2 * FCOEF_N, xtab + 2 * TW * 2, Ehz);
                __syncwarp();
            }
            // ---- upwind sets (static codes), store pointers, accumulators: as hot_axis_item
            const unsigned ownmask = __ballot_sync(0xffffffffu, cur.own != 0);
            const unsigned w4[3] = {cur.mw.x, cur.mw.y, cur.mw.z};
            unsigned fullx[2], tiex[2], anyx[2], allx[2];
#pragma unroll
            for (int j = 0; j < 2; j++) {
                hot_decode((w4[0] >> (j * 16)) & 0xffffu, Ln, fullx[j], tiex[j]);
                anyx[j] = __reduce_or_sync(0xffffffffu, hot_spread_any<PEN_CU>(fullx[j] | tiex[j]));
                allx[j] = __reduce_and_sync(0xffffffffu, hot_spread_all<PEN_CU>(fullx[j]));
            }
            double* kx[2];
            double* kp[2];
#pragma unroll
            for (int j = 0; j < 2; j++) kx[j] = fk_g + ((size_t)__shfl_sync(0xffffffffu, cur.face, j) * blk + lane);
            bool sel[2], act[2];
            double rsel[2];
#pragma unroll
            for (int p = 0; p < 2; p++) {
                unsigned fa, ta, fb, tb2;
                hot_decode((w4[1 + p] & 0xffffu), Ln, fa, ta);
                hot_decode((w4[1 + p] >> 16) & 0xffffu, Ln, fb, tb2);
                sel[p] = fa != 0;
                act[p] = (fa | fb) != 0;
                const int d = 1 + p;
                const double ra = gb_[6 * (3 + 2 * p) + 3 + d], rb = gb_[6 * (4 + 2 * p) + 3 + d];
                rsel[p] = sel[p] ? ra : rb;
                const int fa_id = __shfl_sync(0xffffffffu, cur.face, 2 + 2 * p), fb_id = __shfl_sync(0xffffffffu, cur.face, 3 + 2 * p);
                kp[p] = fk_g + ((size_t)(sel[p] ? fa_id : fb_id) * blk + lane);
            }
            double ax[2][4], ap[2][4];
#pragma unroll
            for (int j = 0; j < 2; j++) {
                ax[j][0] = ax[j][1] = ax[j][2] = ax[j][3] = 0.0;
                ap[j][0] = ap[j][1] = ap[j][2] = ap[j][3] = 0.0;
            }
            // ---- geometry of the cell: entries x-, x+, y-, y+, z-, z+ (canonical order of axis-aligned cells);
            // inside the bundle the y neighbour is y+ for by = 0 and y- for by = 1 (same for z)
            const double G0x = gb_[0], G0y = gb_[1], G0z = gb_[2];
            const double Gxm = gb_[6 * 1 + 0], Gxp = gb_[6 * 2 + 0];
            const double rxm = gb_[6 * 1 + 3], rxp = gb_[6 * 2 + 3];
            const int jyi = by ? 2 : 3, jyh = by ? 3 : 2, jzi = bz ? 4 : 5, jzh = bz ? 5 : 4;
            const double Gyi = gb_[6 * (1 + jyi) + 1], Gyh = gb_[6 * (1 + jyh) + 1];
            const double Gzi = gb_[6 * (1 + jzi) + 2], Gzh = gb_[6 * (1 + jzh) + 2];

            const double* wk = wslot(k, wl) + lane;
            const double* wm = wslot(k - 1, wl) + lane;
            double* wp = wslot(k + 1, wl) + lane;
            const double* wy = wslot(k, wl ^ 1) + lane;
            const double* wz = wslot(k, wl ^ 2) + lane;
            double* wnew = wslot(k + 2, wl);                 // = the slot of position k - 1, refilled chunk by chunk

            for (int ch = 0; ch < nchunk; ch++) {
                // ---- prefetch: window block of position k + 2 (this chunk: its x- use is over), next halo chunk
                {
                    double* st_n = halo + ((q & 1) ^ 1) * (2 * PEN_CI * 32);
                    if (ch + 1 < nchunk) load_halo(hy_c, hz_c, ch + 1, st_n);
                    else if (k + 1 < ns) {
                        load_halo(hy_n, hz_n, 0, st_n);
                        hot_stage_geo(a.geo6 + (size_t)(nxt.e0 + nxt.c) * 6, PEN_NE, geo + (gsel ^ 1) * ((1 + PEN_NE) * 6), lane);
                    }
                    if (ch > 0 && own_k2m >= 0) load_chunk(own_k2m, ch - 1, wnew);
                    if (FUSE && ch == 0 && k + 1 < ns) {
                        // macros of the blocks converted in step k + 1: own(k + 2), halo(k + 1)
                        double* mr = mrec + (gsel ^ 1) * (3 * FCOEF_N);
                        if (own_k2m >= 0) load_mrec(own_k2m, mr, 0);
                        load_mrec(hy_n, mr + FCOEF_N, 6);
                        load_mrec(hz_n, mr + 2 * FCOEF_N, 12);
                    }
                }
                cp_async_commit();
                cp_async_wait<1>();
                __syncwarp();
                const double* sg = halo + (q & 1) * (2 * PEN_CI * 32) + lane;
                const int nsub = min(PEN_CI / PEN_CU, (Ln - ch * PEN_CI) / PEN_CU);
                // one base per stream and chunk; everything inside the (unrolled) chunk is a compile-time offset from it
                const int c0 = ch * PEN_CI;
                const double* const ck = wk + c0 * 32;
                const double* const cm = wm + c0 * 32;
                double* const cp = wp + c0 * 32;
                const double* const cy_ = wy + c0 * 32;
                const double* const cz_ = wz + c0 * 32;
                const double* const ctx = txs + (cb + c0) * 6;
#if PEN_CTXS
                const double* const cct = P.tx6 + (cb + c0) * 6;  // uniform except in the short-row tail slab (two chunk bases per warp)
#endif
                const double* const cxt = xtab + (cb + c0 - tmin) * 2;
                double* const kxc[2] = {kx[0] + c0 * 32, kx[1] + c0 * 32};
                double* const kpc[2] = {kp[0] + c0 * 32, kp[1] + c0 * 32};
#pragma unroll
                for (int sub = 0; sub < PEN_CI / PEN_CU; sub++) {
                    if (sub >= nsub) break;                                        // warp-uniform (short last chunk)
                    const int i0 = c0 + sub * PEN_CU;
                    const int so = sub * PEN_CU;                                    // compile-time after unrolling
                    double v[PEN_CU], g0[PEN_CU], g1[PEN_CU], g2[PEN_CU], base[PEN_CU], W[PEN_CU][4];
#pragma unroll
                    for (int u = 0; u < PEN_CU; u++) {
#if PEN_CTXS
                        double2 t0, t1, t2;
                        t0.x = cct[(so + u) * 6]; t0.y = cct[(so + u) * 6 + 1]; t1.x = cct[(so + u) * 6 + 2];
                        t1.y = cct[(so + u) * 6 + 3]; t2.x = cct[(so + u) * 6 + 4]; t2.y = cct[(so + u) * 6 + 5];
#else
                        const double2 t0 = lds2(ctx + (so + u) * 6), t1 = lds2(ctx + (so + u) * 6 + 2), t2 = lds2(ctx + (so + u) * 6 + 4);
#endif
                        W[u][0] = t0.y; W[u][1] = t1.x; W[u][2] = t1.y; W[u][3] = t2.x;
                        const double xq = t2.y;
                        v[u] = ck[(so + u) * 32];
                        const double vxm = cm[(so + u) * 32];
                        double vxp = cp[(so + u) * 32];
                        const double vyi = cy_[(so + u) * 32], vzi = cz_[(so + u) * 32];
                        double vyh = sg[(so + u) * 32], vzh = sg[(PEN_CI + so + u) * 32];
                        if (FUSE) {
                            vxp = convert(vxp, cxt + (so + u) * 2, xq, Exp);
                            cp[(so + u) * 32] = vxp;                      // from now on the block holds gBarP
                            vyh = convert(vyh, cxt + (TW + so + u) * 2, xq, Ehy);
                            vzh = convert(vzh, cxt + (2 * TW + so + u) * 2, xq, Ehz);
                        }
                        // gradient (stock leastSquaresGrad, zeroBoundaryGrad.C:90-99): one component per face
                        g0[u] = fma(Gxp, vxp, fma(Gxm, vxm, G0x * v[u]));
                        g1[u] = fma(Gyh, vyh, fma(Gyi, vyi, G0y * v[u]));
                        g2[u] = fma(Gzh, vzh, fma(Gzi, vzi, G0z * v[u]));
                        // value at the cell centre moved back by half a step (discreteVelocity.C:498-502)
                        base[u] = fma(t0.x, g0[u], fma(yh, g1[u], fma(zh, g2[u], v[u])));
                    }
                    // ---- x faces
#pragma unroll
                    for (int j = 0; j < 2; j++) {
                        if (!((anyx[j] >> i0) & 1u)) continue;                      // warp-uniform
                        const double r = j ? rxp : rxm;
                        double* const keep = kxc[j] + so * 32;
                        if ((allx[j] >> i0) & 1u) {                                  // warp-uniform
#pragma unroll
                            for (int u = 0; u < PEN_CU; u++) {
                                const double val = fma(r, g0[u], base[u]);
                                if (keep_on) __stcs(keep + u * 32, val);
                                ax[j][0] = fma(W[u][0], val, ax[j][0]); ax[j][1] = fma(W[u][1], val, ax[j][1]);
                                ax[j][2] = fma(W[u][2], val, ax[j][2]); ax[j][3] = fma(W[u][3], val, ax[j][3]);
                            }
                        } else {
                            // the group that holds the sign change of xi_x: all, half (tie, :513-529) or none
                            const unsigned fb = fullx[j] >> i0, tbits = tiex[j] >> i0;
                            const unsigned wbk = ((ownmask >> j) & 1u) ? (fb | tbits) : fb;
#pragma unroll
                            for (int u = 0; u < PEN_CU; u++) {
                                double val = fma(r, g0[u], base[u]);
                                if (keep_on && ((wbk >> u) & 1u)) __stcs(keep + u * 32, val);
                                const int hi = ((fb >> u) & 1u) ? 0x3ff00000 : (((tbits >> u) & 1u) ? 0x3fe00000 : 0);
                                val *= __hiloint2double(hi, 0);
                                ax[j][0] = fma(W[u][0], val, ax[j][0]); ax[j][1] = fma(W[u][1], val, ax[j][1]);
                                ax[j][2] = fma(W[u][2], val, ax[j][2]); ax[j][3] = fma(W[u][3], val, ax[j][3]);
                            }
                        }
                    }
                    // ---- y / z pairs: the one face of the pair this lane is upwind of
#pragma unroll
                    for (int p = 0; p < 2; p++) {
                        const bool st = keep_on && act[p];
#pragma unroll
                        for (int u = 0; u < PEN_CU; u++) {
                            const double val = fma(rsel[p], p ? g2[u] : g1[u], base[u]);
                            if (st) __stcs(kpc[p] + (so + u) * 32, val);
                            ap[p][0] = fma(W[u][0], val, ap[p][0]); ap[p][1] = fma(W[u][1], val, ap[p][1]);
                            ap[p][2] = fma(W[u][2], val, ap[p][2]); ap[p][3] = fma(W[u][3], val, ap[p][3]);
                        }
                    }
                }
                __syncwarp();   // every lane is done with this halo stage (and this chunk of the window) before the refill
                q++;
            }
            // the last chunk of position k - 1 is free now; its own commit group, so that "all but the newest group"
            // below and in the next step covers everything the next step reads first (geometry, macros, halo chunk 0)
            if (own_k2m >= 0) load_chunk(own_k2m, nchunk - 1, wnew);
            cp_async_commit();
            cp_async_wait<1>();
            __syncwarp();

            // ---- face moments: one face at a time through the halo stage consumed last
            {
                double* scratch = halo + ((q & 1) ^ 1) * (2 * PEN_CI * 32);
                // the other stage holds the first halo chunk of the next step, possibly in flight: untouched
                auto reduce_face = [&](const double (&acc)[4], bool on, int face, int owns) {
                    if (!on) return;                                                 // warp-uniform
                    double vv[NM_G];
                    expand_g(acc, wr, y, z, vv);
                    const double tot = pen_reduce13(vv, scratch, lane);
                    const size_t slot = (size_t)2 * face + (owns ? 0 : 1);
                    // one warp owns a slot per launch: the fire-and-forget add keeps the sum deterministic
                    if (lane < NM_G) atomicAdd(a.fslot + slot * nm + lane, tot);
                };
#pragma unroll
                for (int j = 0; j < 2; j++)
                    reduce_face(ax[j], anyx[j] != 0, __shfl_sync(0xffffffffu, cur.face, j), (ownmask >> j) & 1u);
#pragma unroll
                for (int p = 0; p < 2; p++) {
                    const bool la = act[p] && sel[p], lb = act[p] && !sel[p];
                    double mA[4], mB[4];
#pragma unroll
                    for (int t = 0; t < 4; t++) { mA[t] = la ? ap[p][t] : 0.0; mB[t] = lb ? ap[p][t] : 0.0; }
                    const bool onA = __any_sync(0xffffffffu, la), onB = __any_sync(0xffffffffu, lb);
                    reduce_face(mA, onA, __shfl_sync(0xffffffffu, cur.face, 2 + 2 * p), (ownmask >> (2 + 2 * p)) & 1u);
                    reduce_face(mB, onB, __shfl_sync(0xffffffffu, cur.face, 3 + 2 * p), (ownmask >> (3 + 2 * p)) & 1u);
                }
            }
            cur = nxt;
            hy_c = hy_n; hz_c = hz_n;
            // next step: the other warps read this line's slot of position k + 1 (converted above, landed long ago)
            __syncthreads();
        }
        cp_async_wait<0>();
    }
}
