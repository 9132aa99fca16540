// dugks_fast.cuh — tuned variants of the two heavy cell kernels for cells with at most
// FAST_NE faces (hexahedra, prisms, 2-D quads/triangles ...).  Same arithmetic as the generic
// kernels in dugks_kernels.cuh (which remain the path for cells with more faces), but
//   * the per-(cell, face) constants are staged as 16-byte pairs and read with LDS.128,
//   * face kinds / ownership are warp-uniform bit masks in registers (no per-face branches on
//     shared-memory flags),
//   * neighbour row offsets live in registers as 32-bit element indices,
//   * two consecutive velocity points (i, i+1) are processed together so that every dependent
//     FP64 chain has an independent twin (ILP) and 2 x (1 + faces) loads are in flight per thread,
//   * faces for which no lane of the warp is the upwind side are skipped with one vote.
#pragma once
#include "dugks_kernels.cuh"

#define GEO12 12   // doubles per staged entry: G0 G1 | G2 Sx | r0 r1 | r2 Sy | Sz invdc | pad pad

struct FastStage {
    double* geo;   // [FAST_NE][12]
    int* oidx;     // [FAST_NE] element offset of the other cell's / boundary face's row block in its slab array
    int* face;     // [FAST_NE]
};
#define FAST_STAGE_BYTES (FAST_NE * GEO12 * 8 + FAST_NE * 4 * 2)

__device__ __forceinline__ FastStage carve_fast(unsigned char* p) {
    FastStage s;
    s.geo = reinterpret_cast<double*>(p);
    s.oidx = reinterpret_cast<int*>(s.geo + FAST_NE * GEO12);
    s.face = s.oidx + FAST_NE;
    return s;
}

// warp-cooperative staging; returns ne, fills the three masks (bit j: entry j is an internal face /
// a symmetryPlane boundary face / owned by this cell)
__device__ __forceinline__ int stage_fast(const StepArgs& a, int c, int lane, FastStage& s, unsigned& intmask,
                                          unsigned& symmask, unsigned& ownmask) {
    const DevMesh& m = a.m;
    int e0 = m.cell_off[c];
    int ne = m.cell_off[c + 1] - e0;
    bool isint = false, issym = false, isown = false;
    if (ne <= FAST_NE) {
        for (int k = lane; k < ne * GEO12; k += 32) s.geo[k] = m.e_geo12[(size_t)e0 * GEO12 + k];
        if (lane < ne) {
            int o = m.e_other[e0 + lane];
            s.face[lane] = m.e_face[e0 + lane];
            isown = m.e_owner[e0 + lane] != 0;
            if (o >= 0) {
                isint = true;
                s.oidx[lane] = o * a.dv.L * a.dv.Rs;
            } else {
                int b = -1 - o;
                issym = m.b_kind[b] == K_SYMMETRY_PLANE;
                s.oidx[lane] = b * a.dv.L * a.dv.Rs;
            }
        }
    }
    intmask = __ballot_sync(0xffffffffu, isint);
    symmask = __ballot_sync(0xffffffffu, issym);
    ownmask = __ballot_sync(0xffffffffu, isown);
    return ne;
}

__device__ __forceinline__ double2 lds2(const double* p) { return *reinterpret_cast<const double2*>(p); }

// ---------------------------------------------------------------------------------
template <int PHASE, bool HAS_H>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, HAS_H ? 2 : 3)
k_cell_outgoing_fast(StepArgs a) {
    extern __shared__ __align__(16) unsigned char dyn[];
    // layout: txs[NT_MAX][6] (x, W0 | W1, W2 | W3, pad) | per warp: FastStage |
    //         (PHASE 2) ftab[ACC_FACES][3][tabw], lanec[ACC_FACES][3][32], unic[ACC_FACES][2]
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
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int tw = dv.tabw;
    const size_t per_warp = FAST_STAGE_BYTES + (PHASE == 2 ? ((size_t)ACC_FACES * 3 * tw + ACC_FACES * 3 * 32 + ACC_FACES * 2) * 8 : 0);
    unsigned char* wbase = dyn + (size_t)NT_MAX * 6 * 8 + wib * per_warp;
    FastStage st = carve_fast(wbase);
    double* ftab = reinterpret_cast<double*>(wbase + FAST_STAGE_BYTES);
    double* lanec = ftab + (size_t)ACC_FACES * 3 * tw;
    double* unic = lanec + ACC_FACES * 3 * 32;
    const int L = dv.L, Rs = dv.Rs;
    const int nwr = Rs >> 5;
    const long long nitems = (long long)a.m.nc * nwr;
    const double kd = (double)(a.gas.K + 3 - a.gas.D);
    const double hstep = 0.5 * a.dt;
    const double hd = -0.5 * a.dt;
    const int nm = a.nm;
    const size_t slab_c = (size_t)a.slab * a.m.nc * L * Rs, slab_b = (size_t)a.slab * a.m.nbf * L * Rs;
    const double* __restrict__ gbs = a.gb + slab_c;
    const double* __restrict__ hbs = HAS_H ? a.hb + slab_c : nullptr;
    const double* __restrict__ gam_g = a.gam_old_g + slab_b;
    const double* __restrict__ gam_h = HAS_H ? a.gam_old_h + slab_b : nullptr;

    for (long long item = (long long)blockIdx.x * WARPS_PER_CTA + wib; item < nitems;
         item += (long long)gridDim.x * WARPS_PER_CTA) {
        const int c = (int)(item / nwr), r = (int)(item % nwr) * 32 + lane;
        unsigned intmask, symmask, ownmask;
        const int ne = stage_fast(a, c, lane, st, intmask, symmask, ownmask);
        if (ne > FAST_NE) continue;               // handled by the generic kernel
        const int nint = a.m.cell_nint[c];
        if (nint == 0) continue;
        const int grow = a.slab * Rs + r;
        const double y = dv.row_y[grow], z = dv.row_z[grow], wr = dv.row_w[grow];
        const int cb = dv.row_cbase[grow];
        int tmin, span;
        table_range(dv, cb, tmin, span, dv.L);
        __syncwarp();
        const int cidx = c * L * Rs + r;
        int oidx[FAST_NE];
#pragma unroll
        for (int j = 0; j < FAST_NE; j++) oidx[j] = (j < ne) ? st.oidx[j] + r : 0;

        for (int p0 = 0; p0 < nint; p0 += ACC_FACES) {
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
                    ySy[jj] = __dmul_rn(y, G[7]);
                    zSz[jj] = __dmul_rn(z, G[8]);
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

            for (int i0 = 0; i0 < L; i0 += 2) {
                const bool two = i0 + 1 < L;       // warp-uniform
                // ---- loads: own value and all neighbours for the two velocity points
                double vc[2], wc[2], vn[2][FAST_NE], wn[2][FAST_NE];
#pragma unroll
                for (int u = 0; u < 2; u++) {
                    const int off = (i0 + ((u == 1 && !two) ? 0 : u)) * Rs;
                    vc[u] = gbs[cidx + off];
                    wc[u] = HAS_H ? hbs[cidx + off] : 0.0;
#pragma unroll
                    for (int j = 0; j < FAST_NE; j++) {
                        vn[u][j] = 0.0; wn[u][j] = 0.0;
                        if (j < ne) {
                            if ((intmask >> j) & 1u) {
                                vn[u][j] = gbs[oidx[j] + off];
                                if (HAS_H) wn[u][j] = hbs[oidx[j] + off];
                            } else if (!((symmask >> j) & 1u)) {
                                vn[u][j] = gam_g[oidx[j] + off];
                                if (HAS_H) wn[u][j] = gam_h[oidx[j] + off];
                            }
                        }
                    }
                }
                // ---- least-squares gradient (see cell_gradient in dugks_kernels.cuh)
                double gg[2][3] = {{0, 0, 0}, {0, 0, 0}}, gh[2][3] = {{0, 0, 0}, {0, 0, 0}};
#pragma unroll
                for (int j = 0; j < FAST_NE; j++) {
                    if (j < ne) {
                        const double2 G01 = lds2(st.geo + j * GEO12), G2s = lds2(st.geo + j * GEO12 + 2);
                        const bool isint = (intmask >> j) & 1u, issym = (symmask >> j) & 1u;
                        const double idc = isint ? 0.0 : st.geo[j * GEO12 + 9];
#pragma unroll
                        for (int u = 0; u < 2; u++) {
                            double dg, dh = 0.0;
                            if (isint) {
                                dg = vn[u][j] - vc[u];
                                if (HAS_H) dh = wn[u][j] - wc[u];
                            } else if (issym) {
                                dg = 0.0;
                            } else {
                                dg = (vc[u] + vn[u][j] * idc) - vc[u];
                                if (HAS_H) dh = (wc[u] + wn[u][j] * idc) - wc[u];
                            }
                            gg[u][0] = fma(G01.x, dg, gg[u][0]); gg[u][1] = fma(G01.y, dg, gg[u][1]);
                            gg[u][2] = fma(G2s.x, dg, gg[u][2]);
                            if (HAS_H) {
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
                    const int t = cb + i0 + ((u == 1 && !two) ? 0 : u);
                    const double2 t0 = lds2(txs + t * 6), t1 = lds2(txs + t * 6 + 2), t2 = lds2(txs + t * 6 + 4);
                    x[u] = t0.x; W[u][0] = t0.y; W[u][1] = t1.x; W[u][2] = t1.y; W[u][3] = t2.x;
                    if (u == 1 && !two) { W[u][0] = W[u][1] = W[u][2] = W[u][3] = 0.0; }
                    xg[u] = (x[u] * gg[u][0] + y * gg[u][1] + z * gg[u][2]) * hd;   // -0.5 dt xi.grad (discreteVelocity.C:498-502)
                    xh[u] = HAS_H ? (x[u] * gh[u][0] + y * gh[u][1] + z * gh[u][2]) * hd : 0.0;
                }
                // ---- faces
#pragma unroll
                for (int jj = 0; jj < ACC_FACES; jj++) {
                    const int j = p0 + jj;
                    if (j < nint) {
                        const double* G = st.geo + j * GEO12;
                        const double Sx = G[3];
                        const bool isown = (ownmask >> j) & 1u;
                        bool full[2], none[2], neg[2];
#pragma unroll
                        for (int u = 0; u < 2; u++) {
                            const double phi = __dadd_rn(__dadd_rn(__dmul_rn(x[u], Sx), ySy[jj]), zSz[jj]);
                            neg[u] = phi < -DUGKS_VSMALL;                    // discreteVelocity.C:506
                            const bool pos = phi >= DUGKS_VSMALL;            // :495
                            full[u] = isown ? pos : neg[u];
                            none[u] = isown ? neg[u] : pos;
                        }
                        if (PHASE == 1) {
                            if (__all_sync(0xffffffffu, none[0] && none[1])) continue;
                            const double2 r01 = lds2(G + 4);
                            const double r2 = G[6];
#pragma unroll
                            for (int u = 0; u < 2; u++) {
                                double val = vc[u] + (gg[u][0] * r01.x + gg[u][1] * r01.y + gg[u][2] * r2) + xg[u];
                                val = none[u] ? 0.0 : (full[u] ? val : 0.5 * val);          // :513-529
                                accg[jj][0] = fma(W[u][0], val, accg[jj][0]);
                                accg[jj][1] = fma(W[u][1], val, accg[jj][1]);
                                accg[jj][2] = fma(W[u][2], val, accg[jj][2]);
                                accg[jj][3] = fma(W[u][3], val, accg[jj][3]);
                                if (HAS_H) {
                                    double vh = wc[u] + (gh[u][0] * r01.x + gh[u][1] * r01.y + gh[u][2] * r2) + xh[u];
                                    vh = none[u] ? 0.0 : (full[u] ? vh : 0.5 * vh);
                                    acch[jj][0] = fma(W[u][0], vh, acch[jj][0]);
                                    acch[jj][1] = fma(W[u][1], vh, acch[jj][1]);
                                }
                            }
                        } else {
                            // exactly one side writes the face value (ties: the owner)
                            bool wr_[2];
                            wr_[0] = isown ? !neg[0] : neg[0];
                            wr_[1] = two && (isown ? !neg[1] : neg[1]);
                            if (!__any_sync(0xffffffffu, wr_[0] || wr_[1])) continue;
                            const double2 r01 = lds2(G + 4);
                            const double r2 = G[6];
                            const double EYZ = lanec[(jj * 3 + 0) * 32 + lane], YZ2 = lanec[(jj * 3 + 1) * 32 + lane],
                                         QYZ = lanec[(jj * 3 + 2) * 32 + lane];
                            const double omrf = unic[jj * 2], frt = unic[jj * 2 + 1];
                            const size_t fbase = (size_t)st.face[j] * L * Rs + r;
#pragma unroll
                            for (int u = 0; u < 2; u++) {
                                if (wr_[u]) {
                                    const int tt = cb + i0 + u - tmin;
                                    double val = vc[u] + (gg[u][0] * r01.x + gg[u][1] * r01.y + gg[u][2] * r2) + xg[u];
                                    double cc = ftab[(jj * 3 + 1) * tw + tt] + YZ2;
                                    double cq = ftab[(jj * 3 + 2) * tw + tt] + QYZ;
                                    double gM = ftab[(jj * 3 + 0) * tw + tt] * EYZ;
                                    double gS = fma(cq, cc, 1.0) * gM;
                                    const size_t fo = fbase + (size_t)(i0 + u) * Rs;
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

            if (PHASE == 1) {
#pragma unroll
                for (int jj = 0; jj < ACC_FACES; jj++) {
                    const int j = p0 + jj;
                    if (j < nint) {   // warp-uniform
                        double v[16];
                        expand_g(accg[jj], wr, y, z, v);
                        v[13] = v[14] = v[15] = 0.0;
                        double tot = warp_reduce16(v, lane);
                        const size_t slot = (size_t)2 * st.face[j] + (((ownmask >> j) & 1u) ? 0 : 1);
                        const int idx = reduce16_index(lane);
                        if ((lane & 1) == 0 && idx < NM_G) a.fslot[slot * nm + idx] += tot;
                        if (HAS_H) {
                            double uu[16];
                            expand_h(acch[jj], wr, y, z, uu);
#pragma unroll
                            for (int k = NM_H; k < 16; k++) uu[k] = 0.0;
                            double toth = warp_reduce16(uu, lane);
                            if ((lane & 1) == 0 && idx < NM_H) a.fslot[slot * nm + NM_G + idx] += toth;
                        }
                    }
                }
            }
            __syncwarp();
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------
template <bool HAS_H>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32, HAS_H ? 3 : 4)
k_cell_update_fast(StepArgs a) {
    extern __shared__ __align__(16) unsigned char dyn[];
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
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    FastStage st = carve_fast(dyn + (size_t)NT_MAX * 6 * 8 + wib * FAST_STAGE_BYTES);
    const int L = dv.L, Rs = dv.Rs;
    const int nwr = Rs >> 5;
    const long long nitems = (long long)a.m.nc * nwr;
    const int nm = a.nm;
    const size_t slab_c = (size_t)a.slab * a.m.nc * L * Rs, slab_b = (size_t)a.slab * a.m.nbf * L * Rs;
    double* __restrict__ gts = a.gt + slab_c;
    double* __restrict__ hts = HAS_H ? a.ht + slab_c : nullptr;
    const double* __restrict__ gbs = a.gb + slab_c;
    const double* __restrict__ hbs = HAS_H ? a.hb + slab_c : nullptr;
    const double* __restrict__ gsbs = a.gsb + slab_b;
    const double* __restrict__ hsbs = HAS_H ? a.hsb + slab_b : nullptr;

    for (long long item = (long long)blockIdx.x * WARPS_PER_CTA + wib; item < nitems;
         item += (long long)gridDim.x * WARPS_PER_CTA) {
        const int c = (int)(item / nwr), r = (int)(item % nwr) * 32 + lane;
        unsigned intmask, symmask, ownmask;
        const int ne = stage_fast(a, c, lane, st, intmask, symmask, ownmask);
        if (ne > FAST_NE) continue;
        const int grow = a.slab * Rs + r;
        const double y = dv.row_y[grow], z = dv.row_z[grow], wr = dv.row_w[grow];
        const int cb = dv.row_cbase[grow];
        __syncwarp();
        const int cidx = c * L * Rs + r;
        // face value source: internal -> fbuf[face] (slab scratch), boundary -> gsb
        int fidx[FAST_NE];
        double ySy[FAST_NE], zSz[FAST_NE], Sx[FAST_NE];
#pragma unroll
        for (int j = 0; j < FAST_NE; j++) {
            fidx[j] = 0; ySy[j] = zSz[j] = Sx[j] = 0.0;
            if (j < ne) {
                fidx[j] = (((intmask >> j) & 1u) ? st.face[j] * L * Rs : st.oidx[j]) + r;
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
        for (int i0 = 0; i0 < L; i0 += 2) {
            const bool two = i0 + 1 < L;
            double g0[2], gb0[2], h0[2], hb0[2], gf[2][FAST_NE], hf[2][FAST_NE];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int off = (i0 + ((u == 1 && !two) ? 0 : u)) * Rs;
                g0[u] = gts[cidx + off]; gb0[u] = gbs[cidx + off];
                h0[u] = HAS_H ? hts[cidx + off] : 0.0; hb0[u] = HAS_H ? hbs[cidx + off] : 0.0;
#pragma unroll
                for (int j = 0; j < FAST_NE; j++) {
                    gf[u][j] = 0.0; hf[u][j] = 0.0;
                    if (j < ne) {
                        if ((intmask >> j) & 1u) {
                            gf[u][j] = a.fbuf_g[fidx[j] + off];
                            if (HAS_H) hf[u][j] = a.fbuf_h[fidx[j] + off];
                        } else {
                            gf[u][j] = gsbs[fidx[j] + off];
                            if (HAS_H) hf[u][j] = hsbs[fidx[j] + off];
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 2; u++) {
                if (u == 1 && !two) break;
                const int t = cb + i0 + u;
                const double2 t0 = lds2(txs + t * 6), t1 = lds2(txs + t * 6 + 2), t2 = lds2(txs + t * 6 + 4);
                const double x = t0.x;
                double sumg = 0.0, sumh = 0.0;
#pragma unroll
                for (int j = 0; j < FAST_NE; j++) {
                    if (j < ne) {
                        const double sphi = __dadd_rn(__dadd_rn(__dmul_rn(x, Sx[j]), ySy[j]), zSz[j]);   // discreteVelocity.C:952-955
                        sumg = fma(sphi, gf[u][j], sumg);
                        if (HAS_H) sumh = fma(sphi, hf[u][j], sumh);
                    }
                }
                const double gnew = (-1.0 / 3) * g0[u] + (4.0 / 3) * gb0[u] - sumg * dtv;   // :937,952
                gts[cidx + (i0 + u) * Rs] = gnew;
                A[0] = fma(t0.y, gnew, A[0]); A[1] = fma(t1.x, gnew, A[1]);
                A[2] = fma(t1.y, gnew, A[2]); A[3] = fma(t2.x, gnew, A[3]);
                if (HAS_H) {
                    const double hnew = (-1.0 / 3) * h0[u] + (4.0 / 3) * hb0[u] - sumh * dtv;
                    hts[cidx + (i0 + u) * Rs] = hnew;
                    B[0] = fma(t0.y, hnew, B[0]); B[1] = fma(t1.x, hnew, B[1]);
                }
            }
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
