// dugks_pencil_ws.cuh — the fused CTA pencil of dugks_pencil.cuh, WARP-SPECIALISED: every x-line of the 2 x 2 bundle
// has two warps,
//   * a PRODUCER (warp group 1, 96 registers per thread) that fetches the halo chunks and the window block two
//     positions ahead (cp.async), builds the half-step tables (the exponentials) and converts gTilde -> gBarP in
//     shared memory (discreteVelocity.C:393-406), and
//   * a CONSUMER (warp group 0, 160 registers) that runs the stencil on converted values only: least-squares
//     gradient, upwind reconstruction, face values, face moments (:412-530, fvDVM.C:473-516).
// Why: the one-warp-per-line pencil is latency bound at 8 warps per SM (255 registers, 114 KB of
// shared memory per CTA): 40 % of its stall samples sit on the conversion and the table builds, which feed the stencil's
// dependency chain.  Two warps per line at the SAME register and shared-memory budget (setmaxnreg moves registers from
// the producers to the consumers) run the two halves concurrently: 16 warps per SM.
// Hand-over per chunk of PWS_CH points through two mbarriers per ring stage (FULL: producer -> consumer, the halo
// chunk and the chunk of the own block at position k + 1 are converted; EMPTY: consumer -> producer, the stage and
// the same chunk of the block at position k - 1 are free, the latter is refilled with position k + 2).
// Every value is written and read by the SAME lane index in all warps, so a warp-level arrive (32 arrivals) orders it.
#pragma once
#include "dugks_pencil.cuh"

#define PWS_CH 2          // points per hand-over chunk
#define PWS_STAGES 3      // halo ring: one stage in use by the consumer, one converted / converting, one in flight
#define PWS_CONS_REGS 160
#define PWS_PROD_REGS 96

struct PwsPlan {
    static __host__ __device__ size_t txs_bytes(int ntab) { return PenPlan::txs_bytes(ntab); }
    static __host__ __device__ size_t win_bytes(int L) { return PenPlan::win_bytes(L); }
    // per line: ring [STAGES][2 (y, z)][CH][32], geometry [2][7 * 6], coefficient records [2][3][FCOEF_N],
    // tables [3][tw][2], reduction scratch [8][NM_G], mbarriers FULL[STAGES] EMPTY[STAGES]
    static __host__ __device__ size_t line_bytes(int tw) {
        return ((size_t)(PWS_STAGES * 2 * PWS_CH * 32 + 2 * (1 + PEN_NE) * 6 + 2 * 3 * FCOEF_N + 3 * tw * 2 + 8 * NM_G + 2 * PWS_STAGES + 1) * 8 + 127) / 128 * 128;
    }
    static __host__ size_t total(int L, int ntab, int tw) { return txs_bytes(ntab) + win_bytes(L) + PEN_WARPS * line_bytes(tw); }
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <int N>
__device__ __forceinline__ void pws_setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void pws_setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ void pws_bar_consumers() { asm volatile("bar.sync 1, 128;" ::: "memory"); }
// all eight warps, from either role's code path: an mbarrier all 256 threads arrive at (bar.sync reached from two
// program locations is what compute-sanitizer's synccheck calls a divergent barrier)
__device__ __forceinline__ void pws_bar_all(uint64_t* bar, uint32_t& phase) {
    mbar_arrive(bar);
    mbar_wait(bar, phase);
    phase ^= 1u;
}

// Sum of 13 values per lane over the warp through 8 x 13 doubles of scratch; lanes 0..12 return the totals.
__device__ __forceinline__ double pws_reduce13(double (&v)[NM_G], double* scratch, int lane) {
#pragma unroll
    for (int k = 0; k < NM_G; k++) {
        v[k] += __shfl_xor_sync(0xffffffffu, v[k], 16);
        v[k] += __shfl_xor_sync(0xffffffffu, v[k], 8);
    }
    if (lane < 8) {
#pragma unroll
        for (int k = 0; k < NM_G; k++) scratch[lane * NM_G + k] = v[k];
    }
    __syncwarp();
    double s0 = 0.0, s1 = 0.0;
    if (lane < NM_G) {
#pragma unroll
        for (int rr = 0; rr < 8; rr += 2) {
            s0 += scratch[rr * NM_G + lane];
            s1 += scratch[(rr + 1) * NM_G + lane];
        }
    }
    __syncwarp();
    return s0 + s1;
}

__global__ void __launch_bounds__(2 * PEN_WARPS * 32, 2)
k_pencil_ws(StepArgs a, PenArgs P) {
    extern __shared__ __align__(128) unsigned char dyn[];
    const DevDV& dv = a.dv;
    const int L = dv.L, nc = a.m.nc, blk = L * 32;
    // the warp index as a broadcast: the compiler then knows that the role branch below is warp-uniform and does not wrap
    // every shuffle / ballot / __syncwarp inside it in its divergent-collective sequence (WARPSYNC ... ENDCOLLECTIVE)
    const int lane = threadIdx.x & 31, wid = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const bool producer = wid >= PEN_WARPS;
    const int wl = wid & (PEN_WARPS - 1);                   // line of the bundle: by + 2 bz
    const int by = wl & 1, bz = wl >> 1;
    const double hd = -0.5 * a.dt;
    double* txs = reinterpret_cast<double*>(dyn);
    hot_fill_txs(dv, hd, txs);
    double* win = reinterpret_cast<double*>(dyn + PwsPlan::txs_bytes(dv.ntab));                 // [3][4][blk]
    const int TW = dv.tabw;
    double* lbase = reinterpret_cast<double*>(dyn + PwsPlan::txs_bytes(dv.ntab) + PwsPlan::win_bytes(L) + wl * PwsPlan::line_bytes(TW));
    double* ring = lbase;                                    // [STAGES][2][CH][32]
    double* geo = ring + PWS_STAGES * 2 * PWS_CH * 32;       // [2][7 * 6]
    double* mrec = geo + 2 * (1 + PEN_NE) * 6;               // [2][3][FCOEF_N]: records of own(k+1), halo y(k), halo z(k)
    double* xtab = mrec + 2 * 3 * FCOEF_N;                   // [3][TW][2]: EX, X2
    double* scratch = xtab + 3 * TW * 2;                     // [8][NM_G]
    uint64_t* full = reinterpret_cast<uint64_t*>(scratch + 8 * NM_G);   // [STAGES]
    uint64_t* empty = full + PWS_STAGES;                     // [STAGES]
    // the CTA-wide barrier lives in line 0's block
    uint64_t* cta_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<double*>(dyn + PwsPlan::txs_bytes(dv.ntab) + PwsPlan::win_bytes(L)) +
                                                    PWS_STAGES * 2 * PWS_CH * 32 + 2 * (1 + PEN_NE) * 6 + 2 * 3 * FCOEF_N + 3 * TW * 2 + 8 * NM_G) + 2 * PWS_STAGES;
    uint32_t cta_phase = 0;
    if (producer && lane == 0) {
        if (wl == 0) mbar_init(cta_bar, 2 * PEN_WARPS * 32);
#pragma unroll
        for (int s = 0; s < PWS_STAGES; s++) { mbar_init(full + s, 32); mbar_init(empty + s, 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const size_t slab_c = (size_t)a.slab * nc * blk;
    const double* src = a.gt + slab_c;
    const int grow = a.slab * 32 + lane;
    const double y = dv.row_y[grow], z = dv.row_z[grow];
    const int cb = dv.row_cbase[grow];
    const int Ln = dv_len(dv, a.slab);
    const int nchunk = (Ln + PWS_CH - 1) / PWS_CH;
    auto wslot = [&](int p, int line) { return win + ((size_t)((p + 3) % 3) * PEN_WARPS + line) * blk; };
    uint32_t q = 0;     // hand-over chunks so far (both warps of a line count alike): stage q % STAGES, use q / STAGES

    if (producer) {
        // =============================================================================== producer
        pws_setmaxnreg_dec<PWS_PROD_REGS>();
        int tmin = 0, span = 0;
        table_range(dv, cb, tmin, span, Ln);
        auto load_rows = [&](int c, double* dst) {
            const char* g = reinterpret_cast<const char*>(src + (size_t)c * blk) + lane * 16;
            const uint32_t s = smem_u32(dst) + lane * 16;
            for (int p = 0; p < Ln * 256; p += 512) cp_async16(s + p, g + p);
        };
        auto load_chunk = [&](int c, int ch, double* dst /* block base */) {   // 512 bytes: one instruction
            cp_async16(smem_u32(dst) + ch * (PWS_CH * 256) + lane * 16,
                       reinterpret_cast<const char*>(src + (size_t)c * blk) + ch * (PWS_CH * 256) + lane * 16);
        };
        auto load_halo = [&](int cy, int cz, int ch, double* stage) {
            const uint32_t s = smem_u32(stage) + lane * 16;
            cp_async16(s, reinterpret_cast<const char*>(src + (size_t)cy * blk) + ch * (PWS_CH * 256) + lane * 16);
            cp_async16(s + PWS_CH * 256, reinterpret_cast<const char*>(src + (size_t)cz * blk) + ch * (PWS_CH * 256) + lane * 16);
        };
        auto load_mrec = [&](int c, double* dst, int l0) {
            if (lane >= l0 && lane < l0 + FCOEF_N / 2)
                cp_async16(smem_u32(dst + 2 * (lane - l0)), a.ccoef + (size_t)c * FCOEF_N + 2 * (lane - l0));
        };
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
        auto convert = [&](double raw, const double* xt, double x, const PenEq& E) {
            const double2 x01 = lds2(xt);
            const double cc = x01.y + E.YZ2;
            const double cq = fma(x, E.qx, E.QYZ);
            const double gM = x01.x * E.EYZ;
            return fma(E.omrf, raw, fma(cq, cc, 1.0) * gM);
        };
        for (int it = blockIdx.x; it < P.nitems; it += gridDim.x) {
            const PenItem I = P.items[it];
            const int* ctab = P.cells + I.cells + wl;            // own line: ctab[(p + 1) * 4], p = -1 .. ns
            const int* htab = P.halo + I.halo + wl * 2;          // htab[k * 8 + {0, 1}]
            const int ns = I.nsteps;
            pws_bar_all(cta_bar, cta_phase);                                        // (A) the previous item is finished by all eight warps
            // ---- prologue: positions -1, 0, 1 of the own line; -1 and 0 enter converted
            const int cm1 = ctab[0], c0 = ctab[4], c1 = ctab[8];
            load_rows(cm1, wslot(-1, wl));
            load_rows(c0, wslot(0, wl));
            load_rows(c1, wslot(1, wl));
            load_mrec(cm1, mrec, 0);
            load_mrec(c0, mrec + FCOEF_N, 6);
            cp_async_commit();
            cp_async_wait<0>();
            __syncwarp();
            {
                PenEq E;
#pragma unroll 1
                for (int t = 0; t < 2; t++) {
                    build_table(mrec + t * FCOEF_N, xtab, E);
                    __syncwarp();
                    double* blkp = wslot(t - 1, wl) + lane;
                    const double* xt0 = xtab + (cb - tmin) * 2;
#pragma unroll 4
                    for (int i = 0; i < Ln; i++) blkp[i * 32] = convert(blkp[i * 32], xt0 + i * 2, txs[(cb + i) * 6 + 5], E);
                    __syncwarp();
                }
            }
            int hy_c = htab[0], hz_c = htab[1];
            // records of step 0: own(1), halo(0)
            load_mrec(c1, mrec, 0);
            load_mrec(hy_c, mrec + FCOEF_N, 6);
            load_mrec(hz_c, mrec + 2 * FCOEF_N, 12);
            cp_async_commit();
            pws_bar_all(cta_bar, cta_phase);                                        // (B) slots -1 and 0 of every line are converted
            int own2 = (2 <= ns) ? ldg_early(ctab + 3 * 4) : -1;  // position k + 2
            int own_tail = -1;                                    // position k + 1 when its last STAGES chunks are still to be fetched
            int hy_n = (1 < ns) ? ldg_early(htab + 8) : -1, hz_n = (1 < ns) ? ldg_early(htab + 9) : -1;
            for (int k = 0; k < ns; k++) {
                const int gsel = k & 1;
                double* wp = wslot(k + 1, wl);                    // converted chunk by chunk in this step
                double* wnew = wslot(k + 2, wl);                  // = the slot of position k - 1: refilled behind the consumer
                // issue(ch): once the consumer has released hand-over chunk q + ch - STAGES - the ring stage and, with it,
                // the same chunk of the window block it read as its x- neighbour - fetch the halo chunk and refill the window:
                // position k + 2 over the chunks of position k - 1 released in THIS step (ch >= STAGES), the last STAGES
                // chunks of position k + 1 over those released at the end of the previous step (ch < STAGES)
                auto issue = [&](int ch) {
                    const uint32_t qq = q + ch;
                    const int st = qq % PWS_STAGES;
                    mbar_wait(empty + st, ((qq / PWS_STAGES) & 1u) ^ 1u);
                    load_halo(hy_c, hz_c, ch, ring + st * (2 * PWS_CH * 32));
                    if (ch >= PWS_STAGES) { if (own2 >= 0) load_chunk(own2, ch - PWS_STAGES, wnew); }
                    else if (own_tail >= 0) load_chunk(own_tail, nchunk - PWS_STAGES + ch, wp);
                    cp_async_commit();
                };
                issue(0);
                issue(1);
                cp_async_wait<2>();                               // the records of this step (older groups)
                __syncwarp();
                PenEq Exp, Ehy, Ehz;
                {
                    const double* mr = mrec + gsel * (3 * FCOEF_N);
                    build_table(mr, xtab, Exp);
                    build_table(mr + FCOEF_N, xtab + TW * 2, Ehy);
                    build_table(mr + 2 * FCOEF_N, xtab + 2 * TW * 2, Ehz);
                    __syncwarp();
                }
                if (k + 1 < ns) {                                 // records of step k + 1: own(k + 2), halo(k + 1); committed with the next group
                    double* mr = mrec + (gsel ^ 1) * (3 * FCOEF_N);
                    if (own2 >= 0) load_mrec(own2, mr, 0);
                    load_mrec(hy_n, mr + FCOEF_N, 6);
                    load_mrec(hz_n, mr + 2 * FCOEF_N, 12);
                }
                const double* xt0 = xtab + (cb - tmin) * 2;
                for (int ch = 0; ch < nchunk; ch++) {
                    cp_async_wait<1>();                           // chunk ch is here (chunk ch + 1 may be in flight)
                    __syncwarp();
                    const int st = (q + ch) % PWS_STAGES;
                    double* sg = ring + st * (2 * PWS_CH * 32) + lane;
                    const int i0 = ch * PWS_CH;
#pragma unroll
                    for (int u = 0; u < PWS_CH; u++) {
                        if (i0 + u < Ln) {
                            const double xq = txs[(cb + i0 + u) * 6 + 5];
                            const double* xt = xt0 + (i0 + u) * 2;
                            wp[(i0 + u) * 32 + lane] = convert(wp[(i0 + u) * 32 + lane], xt, xq, Exp);
                            sg[u * 32] = convert(sg[u * 32], xt + TW * 2, xq, Ehy);
                            sg[(PWS_CH + u) * 32] = convert(sg[(PWS_CH + u) * 32], xt + 2 * TW * 2, xq, Ehz);
                        }
                    }
                    mbar_arrive(full + st);
                    if (ch + 2 < nchunk) issue(ch + 2);
                    else cp_async_commit();                       // keeps "all but the newest group" meaning chunk ch + 1
                }
                q += nchunk;
                own_tail = own2;
                own2 = (k + 3 <= ns) ? ldg_early(ctab + (k + 4) * 4) : -1;
                hy_c = hy_n; hz_c = hz_n;
                hy_n = (k + 2 < ns) ? ldg_early(htab + (k + 2) * 8) : -1;
                hz_n = (k + 2 < ns) ? ldg_early(htab + (k + 2) * 8 + 1) : -1;
            }
            cp_async_wait<0>();
        }
    } else {
        // =============================================================================== consumer
        pws_setmaxnreg_inc<PWS_CONS_REGS>();
        const double wr = dv.row_w[grow];
        const double yh = hd * y, zh = hd * z;
        const int nm = a.nm;
        double* const fk_g = a.fkeep_g ? a.fkeep_g + (size_t)a.slab * a.m.nif * blk : nullptr;
        const bool keep_on = fk_g != nullptr;
        for (int it = blockIdx.x; it < P.nitems; it += gridDim.x) {
            const PenItem I = P.items[it];
            const int ns = I.nsteps;
            pws_bar_all(cta_bar, cta_phase);                                        // (A)
            HotMeta cur{}, nxt{};
            hot_meta_issue(a, I.item0 + wl, lane, cur);
            hot_stage_geo(a.geo6 + (size_t)(cur.e0 + cur.c) * 6, PEN_NE, geo, lane);
            cp_async_commit();
            pws_bar_all(cta_bar, cta_phase);                                        // (B)
            for (int k = 0; k < ns; k++) {
                const int gsel = k & 1;
                const double* gb_ = geo + gsel * ((1 + PEN_NE) * 6);
                if (k + 1 < ns) {
                    hot_meta_issue(a, I.item0 + (k + 1) * 4 + wl, lane, nxt);
                }
                {   // unpack the record of the current cell
                    const bool valid = lane < PEN_NE;
                    cur.own = valid ? (int)((unsigned)cur.face >> 31) : 0;
                    cur.face = valid ? (cur.face & 0x7fffffff) : 0;
                }
                cp_async_wait<0>();                               // geometry of this cell
                __syncwarp();
                const unsigned ownmask = __ballot_sync(0xffffffffu, cur.own != 0);
                const unsigned w4[3] = {cur.mw.x, cur.mw.y, cur.mw.z};
                unsigned fullx[2], tiex[2], anyx[2], allx[2];
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    hot_decode((w4[0] >> (j * 16)) & 0xffffu, Ln, fullx[j], tiex[j]);
                    anyx[j] = __reduce_or_sync(0xffffffffu, hot_spread_any<PWS_CH>(fullx[j] | tiex[j]));
                    allx[j] = __reduce_and_sync(0xffffffffu, hot_spread_all<PWS_CH>(fullx[j]));
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
                const double G0x = gb_[0], G0y = gb_[1], G0z = gb_[2];
                const double Gxm = gb_[6 * 1 + 0], Gxp = gb_[6 * 2 + 0];
                const double rxm = gb_[6 * 1 + 3], rxp = gb_[6 * 2 + 3];
                const int jyi = by ? 2 : 3, jyh = by ? 3 : 2, jzi = bz ? 4 : 5, jzh = bz ? 5 : 4;
                const double Gyi = gb_[6 * (1 + jyi) + 1], Gyh = gb_[6 * (1 + jyh) + 1];
                const double Gzi = gb_[6 * (1 + jzi) + 2], Gzh = gb_[6 * (1 + jzh) + 2];
                // geometry of the next cell while this one is computed
                if (k + 1 < ns) hot_stage_geo(a.geo6 + (size_t)(nxt.e0 + nxt.c) * 6, PEN_NE, geo + (gsel ^ 1) * ((1 + PEN_NE) * 6), lane);
                cp_async_commit();

                const double* wk = wslot(k, wl) + lane;
                const double* wm = wslot(k - 1, wl) + lane;
                const double* wp = wslot(k + 1, wl) + lane;
                const double* wy = wslot(k, wl ^ 1) + lane;
                const double* wz = wslot(k, wl ^ 2) + lane;
                for (int ch = 0; ch < nchunk; ch++, q++) {
                    const int st = q % PWS_STAGES;
                    mbar_wait(full + st, (q / PWS_STAGES) & 1u);
                    const double* sg = ring + st * (2 * PWS_CH * 32) + lane;
                    const int i0 = ch * PWS_CH;
                    const double* const cct = P.tx6 + (cb + i0) * 6;
                    double v[PWS_CH], g0[PWS_CH], g1[PWS_CH], g2[PWS_CH], base[PWS_CH], W[PWS_CH][4];
#pragma unroll
                    for (int u = 0; u < PWS_CH; u++) {
                        const double hx = cct[u * 6];
                        W[u][0] = cct[u * 6 + 1]; W[u][1] = cct[u * 6 + 2]; W[u][2] = cct[u * 6 + 3]; W[u][3] = cct[u * 6 + 4];
                        v[u] = wk[(i0 + u) * 32];
                        const double vxm = wm[(i0 + u) * 32], vxp = wp[(i0 + u) * 32];
                        const double vyi = wy[(i0 + u) * 32], vzi = wz[(i0 + u) * 32];
                        const double vyh = sg[u * 32], vzh = sg[(PWS_CH + u) * 32];
                        g0[u] = fma(Gxp, vxp, fma(Gxm, vxm, G0x * v[u]));
                        g1[u] = fma(Gyh, vyh, fma(Gyi, vyi, G0y * v[u]));
                        g2[u] = fma(Gzh, vzh, fma(Gzi, vzi, G0z * v[u]));
                        base[u] = fma(hx, g0[u], fma(yh, g1[u], fma(zh, g2[u], v[u])));
                    }
                    // the values of this chunk are in registers: release the stage and the chunk of position k - 1
                    __syncwarp();
                    mbar_arrive(empty + st);
                    // ---- x faces
#pragma unroll
                    for (int j = 0; j < 2; j++) {
                        if (!((anyx[j] >> i0) & 1u)) continue;                      // warp-uniform
                        const double r = j ? rxp : rxm;
                        double* const keep = kx[j] + i0 * 32;
                        if ((allx[j] >> i0) & 1u) {                                  // warp-uniform
#pragma unroll
                            for (int u = 0; u < PWS_CH; u++) {
                                const double val = fma(r, g0[u], base[u]);
                                if (keep_on) __stcs(keep + u * 32, val);
                                ax[j][0] = fma(W[u][0], val, ax[j][0]); ax[j][1] = fma(W[u][1], val, ax[j][1]);
                                ax[j][2] = fma(W[u][2], val, ax[j][2]); ax[j][3] = fma(W[u][3], val, ax[j][3]);
                            }
                        } else {
                            const unsigned fb = fullx[j] >> i0, tbits = tiex[j] >> i0;
                            const unsigned wbk = ((ownmask >> j) & 1u) ? (fb | tbits) : fb;
#pragma unroll
                            for (int u = 0; u < PWS_CH; u++) {
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
                        const bool st_ = keep_on && act[p];
#pragma unroll
                        for (int u = 0; u < PWS_CH; u++) {
                            const double val = fma(rsel[p], p ? g2[u] : g1[u], base[u]);
                            if (st_ && i0 + u < Ln) __stcs(kp[p] + (i0 + u) * 32, val);
                            ap[p][0] = fma(W[u][0], val, ap[p][0]); ap[p][1] = fma(W[u][1], val, ap[p][1]);
                            ap[p][2] = fma(W[u][2], val, ap[p][2]); ap[p][3] = fma(W[u][3], val, ap[p][3]);
                        }
                    }
                }
                // ---- face moments: one face at a time
                {
                    auto reduce_face = [&](const double (&acc)[4], bool on, int face, int owns) {
                        if (!on) return;                                                 // warp-uniform
                        double vv[NM_G];
                        expand_g(acc, wr, y, z, vv);
                        const double tot = pws_reduce13(vv, scratch, lane);
                        const size_t slot = (size_t)2 * face + (owns ? 0 : 1);
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
                // next step: the other consumers read this line's slot of position k + 1 (converted by its producer
                // before the FULL arrivals this warp has waited for)
                pws_bar_consumers();
            }
            cp_async_wait<0>();
        }
    }
}
