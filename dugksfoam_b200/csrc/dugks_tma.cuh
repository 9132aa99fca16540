// dugks_tma.cuh — bulk-async-copy (TMA engine, cp.async.bulk + mbarrier) staged variants of the
// two heavy cell kernels for cells with at most FAST_NE faces.
//
// A warp advances one cell at a time.  Every input it needs is a CONTIGUOUS run in HBM: the
// cell's own row block and one row block per face (neighbour cell, lagged boundary gradient,
// face flux), each L x 32 doubles.  The i-range is cut into chunks of CI points; for every chunk
// one lane per stream issues a single `cp.async.bulk.shared::cluster.global` of CI*256 bytes into
// the warp's shared-memory stage and the warp waits on an mbarrier (complete_tx).  Two stages are
// kept in flight, so the copy of chunk k+1 overlaps the FP64 work on chunk k.  Compared with
// per-element LDG this removes all per-load address arithmetic, keeps >= 2 x streams x CI*256 B in
// flight per warp and hands the DRAM controller long sequential bursts.
#pragma once
#include <cuda/std/cstdint>

#include "dugks_fast.cuh"

#define TMA_STAGES 2
#define TMA_CI 4        // velocity points per bulk copy (1 KB per stream)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// shared memory per warp: FastStage | pad to 128 | stages[TMA_STAGES][nstream][CI][32] doubles | extra
__host__ __device__ inline size_t tma_stage_doubles(int nstream, int ci) { return (size_t)nstream * ci * 32; }
#define TMA_META_BYTES 1024   // FastStage (832 B) + 2 mbarriers, padded
#define TMA_BAR_OFFSET 896
static_assert(FAST_STAGE_BYTES <= TMA_BAR_OFFSET, "FastStage overlaps the mbarriers");

// ---------------------------------------------------------------------------------
// warp reduction of 16 values per lane through shared memory (row stride 17 doubles):
// every lane returns the warp total of value index (lane & 15).  ~50 instructions.
__device__ __forceinline__ double warp_reduce16_smem(const double v[16], double* red, int lane) {
#pragma unroll
    for (int k = 0; k < 16; k++) red[lane * 17 + k] = v[k];
    __syncwarp();
    const int col = lane & 15, half = lane >> 4;
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int rr = 0; rr < 16; rr += 2) {
        s0 += red[(half * 16 + rr) * 17 + col];
        s1 += red[(half * 16 + rr + 1) * 17 + col];
    }
    double s = s0 + s1;
    s += __shfl_xor_sync(0xffffffffu, s, 16);
    __syncwarp();
    return s;
}

// Work of one warp on one cell for the bulk-staged outgoing kernel.
// NE_T >= 0: all faces internal and ne == nint == NE_T (interior cell, fully unrolled,
// no run-time face predicates); NE_T < 0: run-time ne / nint (boundary cells).
template <int PHASE, bool HAS_H, int CI, int TW, int NE_T>
struct OutgoingCell {
    template <class Issue>
    static __device__ __forceinline__ void run(const StepArgs& a, const FastStage& st, const double* txs,
                                               double* stages, size_t stage_d, int nstream, uint64_t* bars,
                                               uint32_t& phase, double* ftab, double* lanec, double* unic,
                                               int ne_rt, int nint_rt, unsigned ownmask, unsigned symmask,
                                               int c, int r, int lane, Issue&& issue) {
        const DevDV& dv = a.dv;
        const int L = dv.L, Rs = dv.Rs, nm = a.nm;
        constexpr int tw = TW;
        const int ne = NE_T >= 0 ? NE_T : ne_rt;
        const int nint = NE_T >= 0 ? NE_T : nint_rt;
        const double kd = (double)(a.gas.K + 3 - a.gas.D);
        const double hstep = 0.5 * a.dt, hd = -0.5 * a.dt;
        const int grow = a.slab * Rs + r;
        const double y = dv.row_y[grow], z = dv.row_z[grow], wr = dv.row_w[grow];
        const int cb = dv.row_cbase[grow];
        int tmin = 0, span = 0;
        if (PHASE == 2) table_range(dv, cb, tmin, span, dv.L);
        const int nchunk = (L + CI - 1) / CI;
        // exact upwind thresholds on the OUTWARD flux phi' = +-(xi.Sf) of this cell
        // (discreteVelocity.C:495,506: owner side full if phi >= VSMALL, neighbour side full if phi < -VSMALL)
        const double V0 = DUGKS_VSMALL;
        const double Vp = __longlong_as_double(__double_as_longlong(V0) + 1);   // next above VSMALL
        const double Vm = __longlong_as_double(__double_as_longlong(V0) - 1);   // next below VSMALL

        for (int p0 = 0; p0 < nint; p0 += ACC_FACES) {
            issue(0, 0);
            double ySy[ACC_FACES], zSz[ACC_FACES];
            double accg[ACC_FACES][4], acch[ACC_FACES][2];
#pragma unroll
            for (int jj = 0; jj < ACC_FACES; jj++) {
                const int j = p0 + jj;
                ySy[jj] = zSz[jj] = 0.0;
                accg[jj][0] = accg[jj][1] = accg[jj][2] = accg[jj][3] = 0.0;
                acch[jj][0] = acch[jj][1] = 0.0;
                if (j < nint) {
                    const double* G = st.geo + j * GEO12;
                    const bool isown = (ownmask >> j) & 1u;
                    const double sgn = isown ? 1.0 : -1.0;
                    ySy[jj] = __dmul_rn(y, sgn * G[7]);
                    zSz[jj] = __dmul_rn(z, sgn * G[8]);
                    if (PHASE == 2) {
                        const double* mf = a.fmac + (size_t)st.face[j] * MAC_N;
                        double rf = hstep / (2.0 * mf[5] + hstep);            // discreteVelocity.C:867
                        EqCoef e = make_eq(a.gas, mf, rf);
                        for (int tt = lane; tt < span; tt += 32) {
                            double cx = txs[(tmin + tt) * 6] - e.Ux;
                            double x2 = cx * cx * e.a;
                            ftab[(jj * 3 + 0) * tw + tt] = exp(-0.5 * x2);
                            ftab[(jj * 3 + 1) * tw + tt] = x2;
                            ftab[(jj * 3 + 2) * tw + tt] = cx * e.qx;
                        }
                        double cy = y - e.Uy, cz = z - e.Uz;
                        double yz2 = (cy * cy + cz * cz) * e.a;
                        lanec[(jj * 3 + 0) * 32 + lane] = e.pre * exp(-0.5 * yz2);
                        lanec[(jj * 3 + 1) * 32 + lane] = yz2 - a.gas.D - 2.0;
                        lanec[(jj * 3 + 2) * 32 + lane] = cy * e.qy + cz * e.qz;
                        if (lane == 0) { unic[jj * 2] = 1.0 - rf; unic[jj * 2 + 1] = e.RT; }
                    }
                }
            }
            if (PHASE == 2) __syncwarp();

            for (int ch = 0; ch < nchunk; ch++) {
                const int s = ch & 1;
                if (ch + 1 < nchunk) issue(ch + 1, s ^ 1);
                mbar_wait(&bars[s], (phase >> s) & 1u);
                phase ^= (1u << s);
                const double* sg = stages + s * stage_d;            // [stream][CI][32]
                const double* sh = sg + (size_t)nstream * CI * 32;
                const int ilen = min(CI, L - ch * CI);
#pragma unroll
                for (int ii = 0; ii < CI; ii += 2) {
                    if (ii >= ilen) break;                          // warp-uniform
                    const bool two = ii + 1 < ilen;
                    const int i0 = ch * CI + ii;
                    const int iu1 = two ? ii + 1 : ii;
                    double vc[2], wc[2];
                    double gg[2][3] = {{0, 0, 0}, {0, 0, 0}}, gh[2][3] = {{0, 0, 0}, {0, 0, 0}};
                    vc[0] = sg[ii * 32 + lane]; vc[1] = sg[iu1 * 32 + lane];
                    wc[0] = HAS_H ? sh[ii * 32 + lane] : 0.0; wc[1] = HAS_H ? sh[iu1 * 32 + lane] : 0.0;
                    // ---- least-squares gradient (see cell_gradient in dugks_kernels.cuh)
#pragma unroll
                    for (int j = 0; j < FAST_NE; j++) {
                        if (j < nint) {                              // internal face: neighbour cell value
                            const double2 G01 = lds2(st.geo + j * GEO12), G2s = lds2(st.geo + j * GEO12 + 2);
#pragma unroll
                            for (int u = 0; u < 2; u++) {
                                const int iu = u ? iu1 : ii;
                                const double dg = sg[((1 + j) * CI + iu) * 32 + lane] - vc[u];
                                gg[u][0] = fma(G01.x, dg, gg[u][0]); gg[u][1] = fma(G01.y, dg, gg[u][1]);
                                gg[u][2] = fma(G2s.x, dg, gg[u][2]);
                                if (HAS_H) {
                                    const double dh = sh[((1 + j) * CI + iu) * 32 + lane] - wc[u];
                                    gh[u][0] = fma(G01.x, dh, gh[u][0]); gh[u][1] = fma(G01.y, dh, gh[u][1]);
                                    gh[u][2] = fma(G2s.x, dh, gh[u][2]);
                                }
                            }
                        } else if (NE_T < 0 && j < ne) {             // boundary face: lagged normal gradient
                            if ((symmask >> j) & 1u) continue;
                            const double2 G01 = lds2(st.geo + j * GEO12), G2s = lds2(st.geo + j * GEO12 + 2);
                            const double idc = st.geo[j * GEO12 + 9];
#pragma unroll
                            for (int u = 0; u < 2; u++) {
                                const int iu = u ? iu1 : ii;
                                const double dg = (vc[u] + sg[((1 + j) * CI + iu) * 32 + lane] * idc) - vc[u];
                                gg[u][0] = fma(G01.x, dg, gg[u][0]); gg[u][1] = fma(G01.y, dg, gg[u][1]);
                                gg[u][2] = fma(G2s.x, dg, gg[u][2]);
                                if (HAS_H) {
                                    const double dh = (wc[u] + sh[((1 + j) * CI + iu) * 32 + lane] * idc) - wc[u];
                                    gh[u][0] = fma(G01.x, dh, gh[u][0]); gh[u][1] = fma(G01.y, dh, gh[u][1]);
                                    gh[u][2] = fma(G2s.x, dh, gh[u][2]);
                                }
                            }
                        }
                    }
                    // ---- per-point constants
                    double x[2], W[2][4], xg[2], xh[2];
#pragma unroll
                    for (int u = 0; u < 2; u++) {
                        const int t = cb + i0 + ((u == 1 && two) ? 1 : 0);
                        const double2 t0 = lds2(txs + t * 6), t1 = lds2(txs + t * 6 + 2), t2 = lds2(txs + t * 6 + 4);
                        x[u] = t0.x; W[u][0] = t0.y; W[u][1] = t1.x; W[u][2] = t1.y; W[u][3] = t2.x;
                        xg[u] = (x[u] * gg[u][0] + y * gg[u][1] + z * gg[u][2]) * hd;   // discreteVelocity.C:498-502
                        xh[u] = HAS_H ? (x[u] * gh[u][0] + y * gh[u][1] + z * gh[u][2]) * hd : 0.0;
                    }
                    // ---- faces
#pragma unroll
                    for (int jj = 0; jj < ACC_FACES; jj++) {
                        const int j = p0 + jj;
                        if (j < nint) {
                            const double* G = st.geo + j * GEO12;
                            const bool isown = (ownmask >> j) & 1u;
                            const double Sxj = isown ? G[3] : -G[3];
                            const double tFj = isown ? V0 : Vp;      // full  <=> phi' >= tF
                            const double tNj = isown ? -V0 : -Vm;    // none  <=> phi' <  tN
                            double phi[2];
                            phi[0] = __dadd_rn(__dadd_rn(__dmul_rn(x[0], Sxj), ySy[jj]), zSz[jj]);
                            phi[1] = __dadd_rn(__dadd_rn(__dmul_rn(x[1], Sxj), ySy[jj]), zSz[jj]);
                            if (PHASE == 1) {
                                const bool act0 = !(phi[0] < tNj);             // this side contributes (full or tie)
                                const bool act1 = two && !(phi[1] < tNj);
                                if (!__any_sync(0xffffffffu, act0 || act1)) continue;
                                const double2 r01 = lds2(G + 4);
                                const double r2 = G[6];
#pragma unroll
                                for (int u = 0; u < 2; u++) {
                                    if (u ? act1 : act0) {
                                        double val = vc[u] + (gg[u][0] * r01.x + gg[u][1] * r01.y + gg[u][2] * r2) + xg[u];
                                        if (!(phi[u] >= tFj)) val *= 0.5;                         // tie :513-529
                                        accg[jj][0] = fma(W[u][0], val, accg[jj][0]);
                                        accg[jj][1] = fma(W[u][1], val, accg[jj][1]);
                                        accg[jj][2] = fma(W[u][2], val, accg[jj][2]);
                                        accg[jj][3] = fma(W[u][3], val, accg[jj][3]);
                                        if (HAS_H) {
                                            double vh = wc[u] + (gh[u][0] * r01.x + gh[u][1] * r01.y + gh[u][2] * r2) + xh[u];
                                            if (!(phi[u] >= tFj)) vh *= 0.5;
                                            acch[jj][0] = fma(W[u][0], vh, acch[jj][0]);
                                            acch[jj][1] = fma(W[u][1], vh, acch[jj][1]);
                                        }
                                    }
                                }
                            } else {
                                // exactly one side writes the face value: the owner unless phi < -VSMALL
                                // (then the neighbour, for which that is the "full" test)
                                const double tW = isown ? tNj : tFj;
                                const bool wr0 = phi[0] >= tW;
                                const bool wr1 = two && (phi[1] >= tW);
                                if (!__any_sync(0xffffffffu, wr0 || wr1)) continue;
                                const double2 r01 = lds2(G + 4);
                                const double r2 = G[6];
                                const double EYZ = lanec[(jj * 3 + 0) * 32 + lane], YZ2 = lanec[(jj * 3 + 1) * 32 + lane],
                                             QYZ = lanec[(jj * 3 + 2) * 32 + lane];
                                const double omrf = unic[jj * 2], frt = unic[jj * 2 + 1];
                                const int fbase = st.face[j] * L * Rs + r;
#pragma unroll
                                for (int u = 0; u < 2; u++) {
                                    if (u ? wr1 : wr0) {
                                        const int tt = cb + i0 + u - tmin;
                                        double val = vc[u] + (gg[u][0] * r01.x + gg[u][1] * r01.y + gg[u][2] * r2) + xg[u];
                                        double cc = ftab[(jj * 3 + 1) * tw + tt] + YZ2;
                                        double cq = ftab[(jj * 3 + 2) * tw + tt] + QYZ;
                                        double gM = ftab[(jj * 3 + 0) * tw + tt] * EYZ;
                                        double gS = fma(cq, cc, 1.0) * gM;
                                        const int fo = fbase + (i0 + u) * Rs;
                                        a.fbuf_g[fo] = fma(omrf, val, gS);                       // :880
                                        if (HAS_H) {
                                            double vh = wc[u] + (gh[u][0] * r01.x + gh[u][1] * r01.y + gh[u][2] * r2) + xh[u];
                                            double hS = (kd + cq * ((cc + 2.0) * kd - 2.0 * a.gas.K)) * gM * frt;
                                            a.fbuf_h[fo] = fma(omrf, vh, hS);                    // :881
                                        }
                                    }
                                }
                            }
                        }
                    }
                }
                __syncwarp();   // all lanes are done with stage s before it is refilled
            }

            if (PHASE == 1) {
                double* red = stages;   // both stages are idle here
#pragma unroll
                for (int jj = 0; jj < ACC_FACES; jj++) {
                    const int j = p0 + jj;
                    if (j < nint) {   // warp-uniform
                        double v[16];
                        expand_g(accg[jj], wr, y, z, v);
                        double uu[NM_H] = {0, 0, 0, 0};
                        if (HAS_H) expand_h(acch[jj], wr, y, z, uu);
                        v[13] = uu[0]; v[14] = uu[1]; v[15] = uu[2];
                        const double tot = warp_reduce16_smem(v, red, lane);
                        const size_t slot = (size_t)2 * st.face[j] + (((ownmask >> j) & 1u) ? 0 : 1);
                        if (lane < 16 && lane < nm) a.fslot[slot * nm + lane] += tot;
                        if (HAS_H) {
                            const double t3 = warp_sum(uu[3]);
                            if (lane == 0) a.fslot[slot * nm + 16] += t3;
                        }
                    }
                }
            }
            __syncwarp();
        }
    }
};

// TW: stride of the per-face equilibrium tables (>= dv.tabw; 32 or 64)
template <int PHASE, bool HAS_H, int CI, int TW>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, HAS_H ? 2 : 3)
k_cell_outgoing_tma(StepArgs a, int nstream /* streams per field incl. own */) {
    extern __shared__ __align__(128) unsigned char dyn[];
    // layout: txs[NT_MAX][6] | per warp: meta(1 KB) | stages | (PHASE 2) ftab, lanec, unic
    double* txs = reinterpret_cast<double*>(dyn);
    const DevDV& dv = a.dv;
    for (int k = threadIdx.x; k < dv.ntab; k += blockDim.x) {
        txs[k * 6 + 0] = dv.tx[k];
        txs[k * 6 + 1] = dv.tx[dv.ntab + k];
        txs[k * 6 + 2] = dv.tx[2 * dv.ntab + k];
        txs[k * 6 + 3] = dv.tx[3 * dv.ntab + k];
        txs[k * 6 + 4] = dv.tx[4 * dv.ntab + k];
        txs[k * 6 + 5] = 0.0;
    }
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    constexpr int tw = TW;
    const int nfld = HAS_H ? 2 : 1;
    const size_t stage_d = tma_stage_doubles(nstream * nfld, CI);
    const size_t extra_d = (PHASE == 2) ? ((size_t)ACC_FACES * 3 * tw + ACC_FACES * 3 * 32 + ACC_FACES * 2) : 0;
    const size_t per_warp = TMA_META_BYTES + (TMA_STAGES * stage_d + extra_d) * 8;
    unsigned char* wbase = dyn + (size_t)NT_MAX * 6 * 8 + wib * per_warp;
    FastStage st = carve_fast(wbase);
    uint64_t* bars = reinterpret_cast<uint64_t*>(wbase + TMA_BAR_OFFSET);
    double* stages = reinterpret_cast<double*>(wbase + TMA_META_BYTES);
    double* ftab = stages + TMA_STAGES * stage_d;
    double* lanec = ftab + (size_t)ACC_FACES * 3 * tw;
    double* unic = lanec + ACC_FACES * 3 * 32;
    if (lane == 0) {
        for (int s = 0; s < TMA_STAGES; s++) mbar_init(&bars[s], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    uint32_t phase = 0;   // parity bit per stage
    const int L = dv.L, Rs = dv.Rs;
    const int nwr = Rs >> 5;
    const long long nitems = (long long)a.m.nc * nwr;
    const size_t slab_c = (size_t)a.slab * a.m.nc * L * Rs, slab_b = (size_t)a.slab * a.m.nbf * L * Rs;
    const double* gbs = a.gb + slab_c;
    const double* hbs = HAS_H ? a.hb + slab_c : nullptr;
    const double* gam_g = a.gam_old_g + slab_b;
    const double* gam_h = HAS_H ? a.gam_old_h + slab_b : nullptr;

    for (long long item = (long long)blockIdx.x * WARPS_PER_CTA + wib; item < nitems;
         item += (long long)gridDim.x * WARPS_PER_CTA) {
        const int c = (int)(item / nwr), r0 = (int)(item % nwr) * 32, r = r0 + lane;
        unsigned intmask, symmask, ownmask;
        const int ne = stage_fast(a, c, lane, st, intmask, symmask, ownmask);
        if (ne > FAST_NE) continue;               // handled by the generic kernel
        const int nint = a.m.cell_nint[c];
        if (nint == 0) continue;
        __syncwarp();
        // ---- stream sources (per lane): lane 0 = own row block, lane 1+j = face j; lanes 16.. = h
        // (a row block is contiguous because a slab is one warp of rows, Rs == 32)
        const int sl = lane & 15;                 // stream slot
        const bool hl = lane >= 16;               // this lane copies an h stream
        const double* src = nullptr;
        if (sl <= ne && (HAS_H || !hl)) {
            if (sl == 0) src = (hl ? hbs : gbs) + (size_t)c * L * Rs + r0;
            else {
                const int j = sl - 1;
                if ((intmask >> j) & 1u) src = (hl ? hbs : gbs) + (size_t)st.oidx[j] + r0;
                else if (!((symmask >> j) & 1u)) src = (hl ? gam_h : gam_g) + (size_t)st.oidx[j] + r0;
            }
        }
        const unsigned active = __ballot_sync(0xffffffffu, src != nullptr);
        const int nactive = __popc(active);
        auto issue = [&](int ch, int s) {
            const int ilen = min(CI, L - ch * CI);
            const uint32_t bytes = (uint32_t)ilen * 32 * 8;
            if (lane == 0) mbar_expect_tx(&bars[s], bytes * nactive);
            __syncwarp();
            if (src != nullptr) {
                double* dst = stages + s * stage_d + ((size_t)(hl ? nstream : 0) + sl) * CI * 32;
                bulk_g2s(dst, src + (size_t)ch * CI * Rs, bytes, &bars[s]);
            }
        };
#define DUGKS_RUN(NE_T)                                                                                      \
    OutgoingCell<PHASE, HAS_H, CI, TW, NE_T>::run(a, st, txs, stages, stage_d, nstream, bars, phase, ftab, lanec, \
                                               unic, ne, nint, ownmask, symmask, c, r, lane, issue)
        if (nint == ne && ne == 6) DUGKS_RUN(6);        // interior hexahedron
        else if (nint == ne && ne == 4) DUGKS_RUN(4);   // interior 2-D quad
        else DUGKS_RUN(-1);
#undef DUGKS_RUN
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------
// stage 5 + cell moments with bulk-staged inputs: streams = own gTilde, own gBarP, one per face
template <bool HAS_H>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, 2)
k_cell_update_tma(StepArgs a, int CI, int nstream /* 2 + max faces */) {
    extern __shared__ __align__(128) unsigned char dyn[];
    double* txs = reinterpret_cast<double*>(dyn);
    const DevDV& dv = a.dv;
    for (int k = threadIdx.x; k < dv.ntab; k += blockDim.x) {
        txs[k * 6 + 0] = dv.tx[k];
        txs[k * 6 + 1] = dv.tx[dv.ntab + k];
        txs[k * 6 + 2] = dv.tx[2 * dv.ntab + k];
        txs[k * 6 + 3] = dv.tx[3 * dv.ntab + k];
        txs[k * 6 + 4] = dv.tx[4 * dv.ntab + k];
        txs[k * 6 + 5] = 0.0;
    }
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int nfld = HAS_H ? 2 : 1;
    const size_t stage_d = tma_stage_doubles(nstream * nfld, CI);
    const size_t per_warp = TMA_META_BYTES + TMA_STAGES * stage_d * 8;
    unsigned char* wbase = dyn + (size_t)NT_MAX * 6 * 8 + wib * per_warp;
    FastStage st = carve_fast(wbase);
    uint64_t* bars = reinterpret_cast<uint64_t*>(wbase + TMA_BAR_OFFSET);
    double* stages = reinterpret_cast<double*>(wbase + TMA_META_BYTES);
    if (lane == 0) {
        for (int s = 0; s < TMA_STAGES; s++) mbar_init(&bars[s], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    uint32_t phase = 0;
    const int L = dv.L, Rs = dv.Rs;
    const int nwr = Rs >> 5;
    const long long nitems = (long long)a.m.nc * nwr;
    const int nm = a.nm;
    const size_t slab_c = (size_t)a.slab * a.m.nc * L * Rs, slab_b = (size_t)a.slab * a.m.nbf * L * Rs;
    double* gts = a.gt + slab_c;
    double* hts = HAS_H ? a.ht + slab_c : nullptr;
    const double* gbs = a.gb + slab_c;
    const double* hbs = HAS_H ? a.hb + slab_c : nullptr;
    const double* gsbs = a.gsb + slab_b;
    const double* hsbs = HAS_H ? a.hsb + slab_b : nullptr;
    const int nchunk = (L + CI - 1) / CI;

    for (long long item = (long long)blockIdx.x * WARPS_PER_CTA + wib; item < nitems;
         item += (long long)gridDim.x * WARPS_PER_CTA) {
        const int c = (int)(item / nwr), r0 = (int)(item % nwr) * 32, r = r0 + lane;
        unsigned intmask, symmask, ownmask;
        const int ne = stage_fast(a, c, lane, st, intmask, symmask, ownmask);
        if (ne > FAST_NE) continue;
        __syncwarp();
        // streams: slot 0 = gTilde, 1 = gBarP, 2+j = face j (internal: slab flux buffer, boundary: gsb)
        const int sl = lane & 15;
        const bool hl = lane >= 16;
        const double* src = nullptr;
        if (sl < ne + 2 && (HAS_H || !hl)) {
            if (sl == 0) src = (hl ? hts : gts) + (size_t)c * L * Rs + r0;
            else if (sl == 1) src = (hl ? hbs : gbs) + (size_t)c * L * Rs + r0;
            else {
                const int j = sl - 2;
                if ((intmask >> j) & 1u) src = (hl ? a.fbuf_h : a.fbuf_g) + (size_t)st.face[j] * L * Rs + r0;
                else src = (hl ? hsbs : gsbs) + (size_t)st.oidx[j] + r0;
            }
        }
        const unsigned active = __ballot_sync(0xffffffffu, src != nullptr);
        const int nactive = __popc(active);
        auto issue = [&](int ch, int s) {
            const int ilen = min(CI, L - ch * CI);
            const uint32_t bytes = (uint32_t)ilen * 32 * 8;
            if (lane == 0) mbar_expect_tx(&bars[s], bytes * nactive);
            __syncwarp();
            if (src != nullptr) {
                double* dst = stages + s * stage_d + ((size_t)(hl ? nstream : 0) + sl) * CI * 32;
                bulk_g2s(dst, src + (size_t)ch * CI * Rs, bytes, &bars[s]);
            }
        };
        issue(0, 0);
        const int grow = a.slab * Rs + r;
        const double y = dv.row_y[grow], z = dv.row_z[grow], wr = dv.row_w[grow];
        const int cb = dv.row_cbase[grow];
        const size_t cidx = (size_t)c * L * Rs + r;
        double ySy[FAST_NE], zSz[FAST_NE], Sx[FAST_NE];
#pragma unroll
        for (int j = 0; j < FAST_NE; j++) {
            ySy[j] = zSz[j] = Sx[j] = 0.0;
            if (j < ne) {
                const double* G = st.geo + j * GEO12;
                // flux sign folded into the area vector: -(xi.Sf) for the neighbour side is exact
                const double sgn = ((ownmask >> j) & 1u) ? 1.0 : -1.0;
                Sx[j] = sgn * G[3];
                ySy[j] = __dmul_rn(y, sgn * G[7]);
                zSz[j] = __dmul_rn(z, sgn * G[8]);
            }
        }
        const double dtv = a.dt / a.m.V[c];
        double A[4] = {0, 0, 0, 0}, B[2] = {0, 0};
        for (int ch = 0; ch < nchunk; ch++) {
            const int s = ch & 1;
            if (ch + 1 < nchunk) issue(ch + 1, s ^ 1);
            mbar_wait(&bars[s], (phase >> s) & 1u);
            phase ^= (1u << s);
            const double* sg = stages + s * stage_d;
            const double* sh = sg + (size_t)nstream * CI * 32;
            const int ilen = min(CI, L - ch * CI);
            for (int ii = 0; ii < ilen; ii++) {
                const int i = ch * CI + ii;
                const int t = cb + i;
                const double2 t0 = lds2(txs + t * 6), t1 = lds2(txs + t * 6 + 2), t2 = lds2(txs + t * 6 + 4);
                const double x = t0.x;
                double sumg = 0.0, sumh = 0.0, sumg2 = 0.0, sumh2 = 0.0;
#pragma unroll
                for (int j = 0; j < FAST_NE; j++) {
                    if (j < ne) {
                        const double sphi = __dadd_rn(__dadd_rn(__dmul_rn(x, Sx[j]), ySy[j]), zSz[j]);   // discreteVelocity.C:952-955
                        const double gf = sg[((2 + j) * CI + ii) * 32 + lane];
                        if (j & 1) sumg2 = fma(sphi, gf, sumg2); else sumg = fma(sphi, gf, sumg);
                        if (HAS_H) {
                            const double hf = sh[((2 + j) * CI + ii) * 32 + lane];
                            if (j & 1) sumh2 = fma(sphi, hf, sumh2); else sumh = fma(sphi, hf, sumh);
                        }
                    }
                }
                const double g0 = sg[ii * 32 + lane], gb0 = sg[(CI + ii) * 32 + lane];
                const double gnew = (-1.0 / 3) * g0 + (4.0 / 3) * gb0 - (sumg + sumg2) * dtv;   // :937,952
                gts[cidx + (size_t)i * Rs] = gnew;
                A[0] = fma(t0.y, gnew, A[0]); A[1] = fma(t1.x, gnew, A[1]);
                A[2] = fma(t1.y, gnew, A[2]); A[3] = fma(t2.x, gnew, A[3]);
                if (HAS_H) {
                    const double h0 = sh[ii * 32 + lane], hb0 = sh[(CI + ii) * 32 + lane];
                    const double hnew = (-1.0 / 3) * h0 + (4.0 / 3) * hb0 - (sumh + sumh2) * dtv;
                    hts[cidx + (size_t)i * Rs] = hnew;
                    B[0] = fma(t0.y, hnew, B[0]); B[1] = fma(t1.x, hnew, B[1]);
                }
            }
            __syncwarp();
        }
        double v[16];
        expand_g(A, wr, y, z, v);
        double uu[NM_H] = {0, 0, 0, 0};
        if (HAS_H) expand_h(B, wr, y, z, uu);
        v[13] = uu[0]; v[14] = uu[1]; v[15] = uu[2];
        double tot = warp_reduce16(v, lane);
        const int idx16 = reduce16_index(lane);
        if ((lane & 1) == 0 && idx16 < nm) a.cslot[(size_t)c * nm + idx16] += tot;
        if (HAS_H) {
            double t3 = warp_sum(uu[3]);
            if (lane == 0) a.cslot[(size_t)c * nm + 16] += t3;
        }
        __syncwarp();
    }
}
