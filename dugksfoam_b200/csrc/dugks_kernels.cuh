// dugks_kernels.cuh — the sm_100a kernels of the discrete-velocity update.
//
// Stage map (reference fvDVM::evolution, fvDVM.C:1086-1108):
//   k_cell_halfstep      updateGHbarPvol                 discreteVelocity.C:346-410
//   k_cell_outgoing<1>   updateGHbarSurf (internal faces) + face moments of updateMacroSurf
//                                                        discreteVelocity.C:412-530, fvDVM.C:473-483,503-516
//   k_bnd_outgoing       updateGHbarSurf (boundary faces, lagged gradient) discreteVelocity.C:436-470,533-690
//   k_bnd_symmetry       updateGHbarSurfSymmetryIn       discreteVelocity.C:733-817, fvDVM.C:375-454
//   k_bnd_moments        boundary part of updateMacroSurf + wall out-flux  discreteVelocity.C:623-624
//   k_face_macros        updateMaxwellWallRho + updateMacroSurf  fvDVM.C:347-367,493-581
//   k_bnd_wall_in        updateGHbarSurfMaxwellWallIn    discreteVelocity.C:693-731
//   k_bnd_relax          updateGHsurf (boundary rules)   discreteVelocity.C:886-931
//   k_cell_outgoing<2>   updateGHsurf (internal faces)   discreteVelocity.C:867-881
//   k_cell_update        updateGHtildeVol + cell moments discreteVelocity.C:934-978, fvDVM.C:612-622,712-721
//   k_cell_macros        updateMacroVol                  fvDVM.C:694-727
//   k_bnd_macros         U,T.correctBoundaryConditions + updatePressureInOutBC  fvDVM.C:698-699,730-806
#pragma once
#include "dugks_device.cuh"

#define NT_MAX 128  // max table length (n1d padded to chunks)
#define WARPS_PER_CTA 4

struct StepArgs {
    DevMesh m;
    DevDV dv;
    DevGas gas;
    int slab;
    int nm;            // moments per slot: 13 or 17
    int skip_small;    // generic cell kernels: skip cells the fast kernels (dugks_fast.cuh) handle
    double dt;
    double *gt, *ht, *gb, *hb;              // [nslab][nc][L][Rs]
    double *gsb, *hsb;                      // [nslab][nbf][L][Rs]
    const double *gam_old_g, *gam_old_h;    // [nslab][nbf][L][Rs]
    double *gam_new_g, *gam_new_h;
    double *fbuf_g, *fbuf_h;                // [nif][L][Rs] slab scratch
    double *fslot;                          // [2 nif + nbf][nm]
    double *cslot;                          // [nc][nm]
    double *cmac;                           // [nc][9]
    double *fmac;                           // [nf][9]
    double *bmac;                           // [nbf][5] rho_b, U_b, T_b
    const double *wall_cin;                 // [nbf][nm] incoming-half-space constants per unit rho_w
    const double *wall_in;                  // [nbf] inComingByRho
    double *wall_diag;                      // [nbf][12] qWall(3), stressWall(9)
    // second-generation kernels (dugks_hot.cuh)
    const double *geo6;                     // per cell (cell_off[c]+c)*6: [G0(3) 0 0 0] then per face [G'(3) r(3)]
    const double *geoS;                     // [ne][4] outward-oriented Sf of every (cell, face) entry
    const uint4 *upw;                       // [nslab][nc][32] upwind range codes (k_build_upwind)
    double *fcoef;                          // [nf][12] face equilibrium records (k_face_macros)
    int item0, item1;                       // item range of a split launch (0, 0 = all cells)
    const int *cmeta;                       // [nc][24] per-cell record of the second-generation kernels (dugks_hot.cuh)
    const int *cmeta2;                      // the same records in the traversal order of k_hot_relax_update (= cmeta unless pencils are on)
    const unsigned char *cell_cls;          // [nc] 1: axis-aligned interior cell (entries in axis order), else 0
    double *fkeep_g, *fkeep_h;              // [n_keep_slabs][nif][L][32] reconstructed face values (face-storage slabs) or null
    double *ccoef;                          // [nc][12] half-step equilibrium records of the cells (k_cell_coef)
    double limiter_k;                       // > 0: Venkatakrishnan-limited least-squares gradient (generic kernels only)
    int gam_late;                           // 1: one copy of the lagged boundary gradient; recompute slabs write it in phase 2
    unsigned long long pol_ef, pol_el;      // L2 cache policies (createpolicy evict_first / evict_last), made once at create
};

__device__ __forceinline__ size_t dv_index(const DevDV& dv, int slab, int n_outer, int outer, int i, int r) {
    return (((size_t)slab * n_outer + outer) * dv.L + i) * dv.Rs + r;
}

// per-warp shared staging of one cell's (cell, face) entries
struct CellStage {
    double* geo;        // [MAX_CELL_FACES][9]
    long long* obase;   // [MAX_CELL_FACES] base offset of the other cell / boundary face row block
    int* kind;          // -1 internal, else patch kind
    int* face;
    int* own;
    double* invdc;
};

#define STAGE_DOUBLES (MAX_CELL_FACES * 9 + MAX_CELL_FACES + MAX_CELL_FACES)
#define STAGE_INTS (MAX_CELL_FACES * 3)
#define STAGE_BYTES (STAGE_DOUBLES * 8 + STAGE_INTS * 4)

__device__ __forceinline__ CellStage carve_stage(unsigned char* p) {
    CellStage s;
    s.geo = reinterpret_cast<double*>(p);
    s.obase = reinterpret_cast<long long*>(s.geo + MAX_CELL_FACES * 9);
    s.invdc = reinterpret_cast<double*>(s.obase + MAX_CELL_FACES);
    s.kind = reinterpret_cast<int*>(s.invdc + MAX_CELL_FACES);
    s.face = s.kind + MAX_CELL_FACES;
    s.own = s.face + MAX_CELL_FACES;
    return s;
}

// stage the entries of cell c (warp-cooperative); returns number of entries
__device__ __forceinline__ int stage_cell(const StepArgs& a, int c, int lane, CellStage& s) {
    const DevMesh& m = a.m;
    int e0 = m.cell_off[c];
    int ne = m.cell_off[c + 1] - e0;
    for (int k = lane; k < ne * 9; k += 32) s.geo[k] = m.e_geo[(size_t)e0 * 9 + k];
    for (int k = lane; k < ne; k += 32) {
        int o = m.e_other[e0 + k];
        s.face[k] = m.e_face[e0 + k];
        s.own[k] = m.e_owner[e0 + k];
        if (o >= 0) {
            s.kind[k] = -1;
            s.obase[k] = (long long)(((size_t)a.slab * m.nc + o) * a.dv.L) * a.dv.Rs;
            s.invdc[k] = 0.0;
        } else {
            int b = -1 - o;
            s.kind[k] = m.b_kind[b];
            s.obase[k] = (long long)(((size_t)a.slab * m.nbf + b) * a.dv.L) * a.dv.Rs;
            s.invdc[k] = m.b_invdc[b];
        }
    }
    __syncwarp();
    return ne;
}

// least-squares gradient of gBarP (and hBarP) of the staged cell at (i, r):
// stock leastSquaresGrad [OF-lib]; in-tree twin zeroBoundaryGrad.C:90-99 plus the
// boundary contribution kept in comments at :126-133; boundary value of a
// fixedGradient patch = cell + gradient/deltaCoeffs (lagged gradient, discreteVelocity.C:408-409)
// VenkatakrishnanSlopeMultiLimiter::limitFace (VenkatakrishnanSlopeMulti.C:83-128): sqrEps = k^3 V
__device__ __forceinline__ double venkat_limit_face(double sqrEps, double dMax, double dMin, double d2) {
    const double two = 2.0 * d2 * d2;
    if (d2 > 0.0) {
        double den = dMax * dMax + two + dMax * d2 + sqrEps;
        if (fabs(den) < DUGKS_VSMALL) den = den < 0 ? -DUGKS_VSMALL : DUGKS_VSMALL;   // stabilise() [OF-lib]
        return ((dMax * dMax + sqrEps) + 2.0 * d2 * dMax) / den;
    } else if (d2 < 0.0) {
        double den = dMin * dMin + two + dMin * d2 + sqrEps;
        if (fabs(den) < DUGKS_VSMALL) den = den < 0 ? -DUGKS_VSMALL : DUGKS_VSMALL;
        return ((dMin * dMin + sqrEps) + 2.0 * d2 * dMin) / den;
    }
    return 1.0;
}

// a.limiter_k > 0: gradSchemes "VenkatakrishnanLimited leastSquares k" AS IT IS MEANT TO WORK
// (VenkatakrishnanLimitedGrads.C:59-226: limiter = min over the faces of the cell of limitFace(V, max - phi,
// min - phi, (Cf - C).grad) with max / min over the face neighbours and patch values, grad *= limiter; the
// reference itself limits a copy and returns the unlimited gradient, :76,:225, i.e. behaves like limiter_k = 0).
// V: volume of the staged cell.
template <bool HAS_H>
__device__ __forceinline__ void cell_gradient(const StepArgs& a, const CellStage& s, int ne, size_t ir,
                                              double v0, double w0, double g[3], double h[3], double V = 0.0) {
    g[0] = g[1] = g[2] = 0.0;
    h[0] = h[1] = h[2] = 0.0;
    double gmax = 0.0, gmin = 0.0, hmax = 0.0, hmin = 0.0;   // max / min of (neighbour - cell), the cell itself included
    for (int j = 0; j < ne; j++) {
        int kind = s.kind[j];
        size_t off = (size_t)s.obase[j] + ir;
        double dg, dh = 0.0;
        if (kind < 0) {
            dg = a.gb[off] - v0;
            if (HAS_H) dh = a.hb[off] - w0;
        } else if (kind == K_SYMMETRY_PLANE) {
            dg = 0.0;
        } else {
            double idc = s.invdc[j];
            dg = (v0 + a.gam_old_g[off] * idc) - v0;
            if (HAS_H) dh = (w0 + a.gam_old_h[off] * idc) - w0;
        }
        const double* G = s.geo + j * 9;
        g[0] += G[0] * dg; g[1] += G[1] * dg; g[2] += G[2] * dg;
        if (HAS_H) { h[0] += G[0] * dh; h[1] += G[1] * dh; h[2] += G[2] * dh; }
        gmax = fmax(gmax, dg); gmin = fmin(gmin, dg);
        hmax = fmax(hmax, dh); hmin = fmin(hmin, dh);
    }
    if (a.limiter_k > 0.0) {
        const double sqrEps = a.limiter_k * a.limiter_k * a.limiter_k * V;
        double lg = 1.0, lh = 1.0;
        for (int j = 0; j < ne; j++) {
            const double* G = s.geo + j * 9;
            lg = fmin(lg, venkat_limit_face(sqrEps, gmax, gmin, G[3] * g[0] + G[4] * g[1] + G[5] * g[2]));
            if (HAS_H) lh = fmin(lh, venkat_limit_face(sqrEps, hmax, hmin, G[3] * h[0] + G[4] * h[1] + G[5] * h[2]));
        }
        g[0] *= lg; g[1] *= lg; g[2] *= lg;
        if (HAS_H) { h[0] *= lh; h[1] *= lh; h[2] *= lh; }
    }
}

// ---- fast path for cells with at most FAST_NE faces: all neighbour values of one (i, r) are
// fetched into registers with independent loads (memory-level parallelism), one iteration ahead
#define FAST_NE 8
template <bool HAS_H>
struct NbrVals {
    double v0, w0;
    double vn[FAST_NE], wn[FAST_NE];
};

template <bool HAS_H>
__device__ __forceinline__ void nbr_load(const StepArgs& a, const CellStage& s, int ne, size_t cbase, size_t ir,
                                         NbrVals<HAS_H>& o) {
    o.v0 = a.gb[cbase + ir];
    o.w0 = HAS_H ? a.hb[cbase + ir] : 0.0;
#pragma unroll
    for (int j = 0; j < FAST_NE; j++) {
        o.vn[j] = 0.0;
        o.wn[j] = 0.0;
        if (j < ne) {
            int kind = s.kind[j];
            size_t off = (size_t)s.obase[j] + ir;
            if (kind < 0) {
                o.vn[j] = a.gb[off];
                if (HAS_H) o.wn[j] = a.hb[off];
            } else if (kind != K_SYMMETRY_PLANE) {
                o.vn[j] = a.gam_old_g[off];
                if (HAS_H) o.wn[j] = a.gam_old_h[off];
            }
        }
    }
}

template <bool HAS_H>
__device__ __forceinline__ void nbr_gradient(const CellStage& s, int ne, const NbrVals<HAS_H>& o, double g[3],
                                             double h[3]) {
    g[0] = g[1] = g[2] = 0.0;
    h[0] = h[1] = h[2] = 0.0;
#pragma unroll
    for (int j = 0; j < FAST_NE; j++) {
        if (j < ne) {
            int kind = s.kind[j];
            double dg, dh = 0.0;
            if (kind < 0) {
                dg = o.vn[j] - o.v0;
                if (HAS_H) dh = o.wn[j] - o.w0;
            } else if (kind == K_SYMMETRY_PLANE) {
                dg = 0.0;
            } else {
                double idc = s.invdc[j];
                dg = (o.v0 + o.vn[j] * idc) - o.v0;
                if (HAS_H) dh = (o.w0 + o.wn[j] * idc) - o.w0;
            }
            const double* G = s.geo + j * 9;
            g[0] += G[0] * dg; g[1] += G[1] * dg; g[2] += G[2] * dg;
            if (HAS_H) { h[0] += G[0] * dh; h[1] += G[1] * dh; h[2] += G[2] * dh; }
        }
    }
}

// warp-uniform table range of the rows held by this warp
__device__ __forceinline__ void table_range(const DevDV& dv, int cb, int& tmin, int& span, int len) {
    int mn = cb, mx = cb;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    tmin = mn;
    span = mx + len - mn;
    if (span > dv.tabw) span = dv.tabw;  // create() sizes tabw as the largest span of any warp
}

// equilibrium tables along x for one macro state: tab[0]=exp(-cx^2 a/2), tab[1]=cx^2 a, tab[2]=cx q'x
__device__ __forceinline__ void build_xtables(const double* txs, const EqCoef& e, int tmin, int span, int lane,
                                              double* tab, int tw) {
    for (int tt = lane; tt < span; tt += 32) {
        double cx = txs[tmin + tt] - e.Ux;
        double x2 = cx * cx * e.a;
        tab[tt] = exp(-0.5 * x2);
        tab[tw + tt] = x2;
        tab[2 * tw + tt] = cx * e.qx;
    }
}

// ---------------------------------------------------------------------------------
// stage 1 (and initialisation when init != 0: gTilde = Shakhov(rho,U,T,q=0),
// discreteVelocity.C:220-249)
template <bool HAS_H>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
k_cell_halfstep(StepArgs a, int init) {
    __shared__ double txs[NT_MAX];
    __shared__ double tabs[WARPS_PER_CTA][3 * NT_MAX];
    const DevDV& dv = a.dv;
    for (int k = threadIdx.x; k < dv.ntab; k += blockDim.x) txs[k] = dv.tx[k];
    __syncthreads();
    int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    int nwr = dv.Rs >> 5;
    long long nitems = (long long)a.m.nc * nwr;
    double* tab = tabs[wib];
    const int tw = dv.tabw;
    const double kd = (double)(a.gas.K + 3 - a.gas.D);
    for (long long item = (long long)blockIdx.x * WARPS_PER_CTA + wib; item < nitems;
         item += (long long)gridDim.x * WARPS_PER_CTA) {
        int c = (int)(item / nwr), r = (int)(item % nwr) * 32 + lane;
        int grow = a.slab * dv.Rs + r;
        double y = dv.row_y[grow], z = dv.row_z[grow];
        int cb = dv.row_cbase[grow];
        int tmin, span;
        table_range(dv, cb, tmin, span, dv_len(dv, a.slab));
        const double* mc = a.cmac + (size_t)c * MAC_N;
        double rf = init ? 1.0 : 1.5 * a.dt / (2.0 * mc[5] + a.dt);   // discreteVelocity.C:393
        EqCoef e = make_eq(a.gas, mc, rf);
        if (init) { e.qx = e.qy = e.qz = 0.0; }
        build_xtables(txs, e, tmin, span, lane, tab, tw);
        __syncwarp();
        double cy = y - e.Uy, cz = z - e.Uz;
        double yz2 = (cy * cy + cz * cz) * e.a;
        double EYZ = e.pre * exp(-0.5 * yz2);
        double YZ2 = yz2 - a.gas.D - 2.0;
        double QYZ = cy * e.qy + cz * e.qz;
        double omrf = 1.0 - rf;
        size_t base = dv_index(dv, a.slab, a.m.nc, c, 0, r);
        int t0 = cb - tmin;
        const double* src_g = a.gt;
        const double* src_h = a.ht;
        double* dst_g = init ? a.gt : a.gb;
        double* dst_h = init ? a.ht : a.hb;
        for (int i = 0; i < dv_len(dv, a.slab); i++) {
            size_t idx = base + (size_t)i * dv.Rs;
            double cc = tab[tw + t0 + i] + YZ2;          // cSqrByRT - D - 2
            double cq = tab[2 * tw + t0 + i] + QYZ;      // (1-Pr) cqBy5pRT
            double gM = tab[t0 + i] * EYZ;                    // rf * gEqBGK
            double gS = fma(cq, cc, 1.0) * gM;                // discreteVelocity.C:1042
            double g0 = init ? 0.0 : src_g[idx];
            dst_g[idx] = fma(omrf, g0, gS);                   // :405
            if (HAS_H) {
                double hS = (kd + cq * ((cc + 2.0) * kd - 2.0 * a.gas.K)) * gM * e.RT;  // :1043
                double h0 = init ? 0.0 : src_h[idx];
                dst_h[idx] = fma(omrf, h0, hS);               // :406
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------
// stages 2.1 / 3 (PHASE 1) and 4 (PHASE 2) on internal faces, cell-centred:
// every cell evaluates the faces for which it is the upwind side.
template <int PHASE, bool HAS_H>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
k_cell_outgoing(StepArgs a) {
    extern __shared__ __align__(16) unsigned char dyn[];
    // layout: txs[5][NT_MAX] | per warp: stage | (PHASE 2) ftab[ACC_FACES][3][tabw]
    double* txs = reinterpret_cast<double*>(dyn);
    const DevDV& dv = a.dv;
    for (int k = threadIdx.x; k < 5 * dv.ntab; k += blockDim.x) txs[(k / dv.ntab) * NT_MAX + (k % dv.ntab)] = dv.tx[k];
    __syncthreads();
    int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int tw = dv.tabw;
    const size_t per_warp = STAGE_BYTES + (PHASE == 2 ? (size_t)ACC_FACES * 3 * tw * 8 : 0);
    unsigned char* wbase = dyn + 5 * NT_MAX * 8 + wib * per_warp;
    CellStage st = carve_stage(wbase);
    double* ftab = reinterpret_cast<double*>(wbase + STAGE_BYTES);
    int nwr = dv.Rs >> 5;
    long long nitems = (long long)a.m.nc * nwr;
    const double kd = (double)(a.gas.K + 3 - a.gas.D);
    const double hstep = 0.5 * a.dt;
    const int nm = a.nm;

    for (long long item = (long long)blockIdx.x * WARPS_PER_CTA + wib; item < nitems;
         item += (long long)gridDim.x * WARPS_PER_CTA) {
        int c = (int)(item / nwr), r = (int)(item % nwr) * 32 + lane;
        int grow = a.slab * dv.Rs + r;
        double y = dv.row_y[grow], z = dv.row_z[grow], wr = dv.row_w[grow];
        int cb = dv.row_cbase[grow];
        int tmin, span;
        table_range(dv, cb, tmin, span, dv_len(dv, a.slab));
        int ne = stage_cell(a, c, lane, st);
        if (a.skip_small && ne <= FAST_NE) continue;   // warp-uniform
        int nint = a.m.cell_nint[c];
        size_t base = dv_index(dv, a.slab, a.m.nc, c, 0, r);

        for (int p0 = 0; p0 < nint; p0 += ACC_FACES) {
            double ySy[ACC_FACES], zSz[ACC_FACES];
            double accg[ACC_FACES][4];
            double acch[ACC_FACES][2];
            double EYZ[ACC_FACES], YZ2[ACC_FACES], QYZ[ACC_FACES], OMRF[ACC_FACES], FRT[ACC_FACES];
#pragma unroll
            for (int jj = 0; jj < ACC_FACES; jj++) {
                int j = p0 + jj;
                ySy[jj] = zSz[jj] = 0.0;
                accg[jj][0] = accg[jj][1] = accg[jj][2] = accg[jj][3] = 0.0;
                acch[jj][0] = acch[jj][1] = 0.0;
                EYZ[jj] = YZ2[jj] = QYZ[jj] = OMRF[jj] = FRT[jj] = 0.0;
                if (j < nint) {
                    const double* G = st.geo + j * 9;
                    ySy[jj] = __dmul_rn(y, G[7]);
                    zSz[jj] = __dmul_rn(z, G[8]);
                    if (PHASE == 2) {
                        const double* mf = a.fmac + (size_t)st.face[j] * MAC_N;
                        double rf = hstep / (2.0 * mf[5] + hstep);            // discreteVelocity.C:867
                        EqCoef e = make_eq(a.gas, mf, rf);
                        build_xtables(txs, e, tmin, span, lane, ftab + jj * 3 * tw, tw);
                        double cy = y - e.Uy, cz = z - e.Uz;
                        double yz2 = (cy * cy + cz * cz) * e.a;
                        EYZ[jj] = e.pre * exp(-0.5 * yz2);
                        YZ2[jj] = yz2 - a.gas.D - 2.0;
                        QYZ[jj] = cy * e.qy + cz * e.qz;
                        OMRF[jj] = 1.0 - rf;
                        FRT[jj] = e.RT;
                    }
                }
            }
            if (PHASE == 2) __syncwarp();

            auto body = [&](int i, double v0, double w0, const double* gg, const double* gh) {
                int t = cb + i;
                double x = txs[t];
                // reconstruction point: Cf - C - 0.5 xi dt  (discreteVelocity.C:498-502)
                double hd = -0.5 * a.dt;
                double xg = (x * gg[0] + y * gg[1] + z * gg[2]) * hd;
                double xh = HAS_H ? (x * gh[0] + y * gh[1] + z * gh[2]) * hd : 0.0;
                double W0 = 0, W1 = 0, W2 = 0, W3 = 0;
                if (PHASE == 1) {
                    W0 = txs[NT_MAX + t]; W1 = txs[2 * NT_MAX + t]; W2 = txs[3 * NT_MAX + t]; W3 = txs[4 * NT_MAX + t];
                }
#pragma unroll
                for (int jj = 0; jj < ACC_FACES; jj++) {
                    int j = p0 + jj;
                    if (j < nint) {
                        const double* G = st.geo + j * 9;
                        double phi = __dadd_rn(__dadd_rn(__dmul_rn(x, G[6]), ySy[jj]), zSz[jj]);
                        bool isown = st.own[j] != 0;
                        bool neg = phi < -DUGKS_VSMALL, pos = phi >= DUGKS_VSMALL;   // discreteVelocity.C:495,506
                        bool full = isown ? pos : neg;
                        bool none = isown ? neg : pos;
                        if (PHASE == 1) {
                            if (!none) {
                                double val = v0 + (gg[0] * G[3] + gg[1] * G[4] + gg[2] * G[5]) + xg;
                                if (!full) val *= 0.5;                                  // :513-529
                                accg[jj][0] = fma(W0, val, accg[jj][0]);
                                accg[jj][1] = fma(W1, val, accg[jj][1]);
                                accg[jj][2] = fma(W2, val, accg[jj][2]);
                                accg[jj][3] = fma(W3, val, accg[jj][3]);
                                if (HAS_H) {
                                    double vh = w0 + (gh[0] * G[3] + gh[1] * G[4] + gh[2] * G[5]) + xh;
                                    if (!full) vh *= 0.5;
                                    acch[jj][0] = fma(W0, vh, acch[jj][0]);
                                    acch[jj][1] = fma(W1, vh, acch[jj][1]);
                                }
                            }
                        } else {
                            bool writer = isown ? !neg : neg;   // exactly one side writes (ties: owner)
                            if (writer) {
                                const double* ft = ftab + jj * 3 * tw;
                                int tt = t - tmin;
                                double val = v0 + (gg[0] * G[3] + gg[1] * G[4] + gg[2] * G[5]) + xg;
                                double cc = ft[tw + tt] + YZ2[jj];
                                double cq = ft[2 * tw + tt] + QYZ[jj];
                                double gM = ft[tt] * EYZ[jj];
                                double gS = fma(cq, cc, 1.0) * gM;
                                size_t fo = ((size_t)st.face[j] * dv.L + i) * dv.Rs + r;
                                a.fbuf_g[fo] = fma(OMRF[jj], val, gS);                  // :880
                                if (HAS_H) {
                                    double vh = w0 + (gh[0] * G[3] + gh[1] * G[4] + gh[2] * G[5]) + xh;
                                    double hS = (kd + cq * ((cc + 2.0) * kd - 2.0 * a.gas.K)) * gM * FRT[jj];
                                    a.fbuf_h[fo] = fma(OMRF[jj], vh, hS);               // :881
                                }
                            }
                        }
                    }
                }
            };
            const bool fast = ne <= FAST_NE && !(a.limiter_k > 0.0);   // the limited gradient lives in cell_gradient only
            const size_t cellbase = base - r;   // base offset of the cell's row block (without r)
            if (fast) {
                NbrVals<HAS_H> A, B;
                nbr_load<HAS_H>(a, st, ne, cellbase, (size_t)r, A);
                for (int i = 0; i < dv_len(dv, a.slab); i += 2) {
                    double gg[3], gh[3];
                    bool hasB = i + 1 < dv_len(dv, a.slab);
                    if (hasB) nbr_load<HAS_H>(a, st, ne, cellbase, (size_t)(i + 1) * dv.Rs + r, B);
                    nbr_gradient<HAS_H>(st, ne, A, gg, gh);
                    body(i, A.v0, A.w0, gg, gh);
                    if (i + 2 < dv_len(dv, a.slab)) nbr_load<HAS_H>(a, st, ne, cellbase, (size_t)(i + 2) * dv.Rs + r, A);
                    if (hasB) {
                        nbr_gradient<HAS_H>(st, ne, B, gg, gh);
                        body(i + 1, B.v0, B.w0, gg, gh);
                    }
                }
            } else {
                for (int i = 0; i < dv_len(dv, a.slab); i++) {
                    size_t ir = (size_t)i * dv.Rs + r;
                    double v0 = a.gb[base + (size_t)i * dv.Rs];
                    double w0 = HAS_H ? a.hb[base + (size_t)i * dv.Rs] : 0.0;
                    double gg[3], gh[3];
                    cell_gradient<HAS_H>(a, st, ne, ir, v0, w0, gg, gh, a.m.V[c]);
                    body(i, v0, w0, gg, gh);
                }
            }

            if (PHASE == 1) {
#pragma unroll
                for (int jj = 0; jj < ACC_FACES; jj++) {
                    int j = p0 + jj;
                    if (j < nint) {   // warp-uniform
                        double v[16];
                        expand_g(accg[jj], wr, y, z, v);
                        v[13] = v[14] = v[15] = 0.0;
                        double tot = warp_reduce16(v, lane);
                        size_t slot = (size_t)2 * st.face[j] + (st.own[j] ? 0 : 1);
                        int idx = reduce16_index(lane);
                        if ((lane & 1) == 0 && idx < NM_G) a.fslot[slot * nm + idx] += tot;
                        if (HAS_H) {
                            double u[16];
                            expand_h(acch[jj], wr, y, z, u);
#pragma unroll
                            for (int k = NM_H; k < 16; k++) u[k] = 0.0;
                            double toth = warp_reduce16(u, lane);
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
// boundary faces of stage 2.1: lagged normal gradient + patch rules.
// item = (boundary face, row-warp)
template <bool HAS_H>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
k_bnd_outgoing(StepArgs a, int far_only) {
    extern __shared__ __align__(16) unsigned char dyn[];
    double* txs = reinterpret_cast<double*>(dyn);
    const DevDV& dv = a.dv;
    for (int k = threadIdx.x; k < dv.ntab; k += blockDim.x) txs[k] = dv.tx[k];
    __syncthreads();
    int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    CellStage st = carve_stage(dyn + NT_MAX * 8 + wib * STAGE_BYTES);
    int nwr = dv.Rs >> 5;
    long long nitems = (long long)a.m.nbf * nwr;
    for (long long item = (long long)blockIdx.x * WARPS_PER_CTA + wib; item < nitems;
         item += (long long)gridDim.x * WARPS_PER_CTA) {
        int b = (int)(item / nwr), r = (int)(item % nwr) * 32 + lane;
        int grow = a.slab * dv.Rs + r;
        double y = dv.row_y[grow], z = dv.row_z[grow];
        int cb = dv.row_cbase[grow];
        int c = a.m.b_owner[b];
        int kind = a.m.b_kind[b];
        // with the fused cell kernels (dugks_hot.cuh) only the incoming half of far-field patches is left
        if (far_only && kind != K_FAR_FIELD && kind != K_PRESSURE_IN && kind != K_PRESSURE_OUT) continue;
        int ne = stage_cell(a, c, lane, st);
        const double* Sf = a.m.b_Sf + (size_t)b * 3;
        const double* rr = a.m.b_r + (size_t)b * 3;
        const double* nn = a.m.b_n + (size_t)b * 3;
        double sx = Sf[0], sy = Sf[1], sz = Sf[2];
        size_t cbase = dv_index(dv, a.slab, a.m.nc, c, 0, r);
        size_t bbase = dv_index(dv, a.slab, a.m.nbf, b, 0, r);
        const double* bm = a.bmac + (size_t)b * 5;
        const double* mc = a.cmac + (size_t)c * MAC_N;
        for (int i = 0; i < dv_len(dv, a.slab); i++) {
            size_t ir = (size_t)i * dv.Rs + r;
            double x = txs[cb + i];
            double v0 = a.gb[cbase + (size_t)i * dv.Rs];
            double w0 = HAS_H ? a.hb[cbase + (size_t)i * dv.Rs] : 0.0;
            double gg[3], gh[3];
            cell_gradient<HAS_H>(a, st, ne, ir, v0, w0, gg, gh, a.m.V[c]);
            size_t bo = bbase + (size_t)i * dv.Rs;
            if (kind != K_SYMMETRY_PLANE) {   // discreteVelocity.C:444-468
                a.gam_new_g[bo] = gg[0] * nn[0] + gg[1] * nn[1] + gg[2] * nn[2];
                if (HAS_H) a.gam_new_h[bo] = gh[0] * nn[0] + gh[1] * nn[1] + gh[2] * nn[2];
            }
            double phi = dot_exact(x, y, z, sx, sy, sz);
            double hd = -0.5 * a.dt;
            double gOut = v0 + (gg[0] * rr[0] + gg[1] * rr[1] + gg[2] * rr[2]) + (x * gg[0] + y * gg[1] + z * gg[2]) * hd;
            double hOut = HAS_H ? w0 + (gh[0] * rr[0] + gh[1] * rr[1] + gh[2] * rr[2]) + (x * gh[0] + y * gh[1] + z * gh[2]) * hd : 0.0;
            switch (kind) {
            case K_ZERO_GRADIENT:                       // :551-555
                a.gsb[bo] = v0;
                if (HAS_H) a.hsb[bo] = w0;
                break;
            case K_MIXED:                               // :556-573
            case K_MAXWELL_WALL:                        // :605-627 (out-flux is summed in k_bnd_moments)
                if (phi > 0) { a.gsb[bo] = gOut; if (HAS_H) a.hsb[bo] = hOut; }
                break;
            case K_FAR_FIELD:
            case K_PRESSURE_IN:
            case K_PRESSURE_OUT:                        // :574-604
                if (phi > 0) { a.gsb[bo] = gOut; if (HAS_H) a.hsb[bo] = hOut; }
                else {
                    double gi = bm[0] * maxwell_by_rho(a.gas, x, y, z, mc[1], mc[2], mc[3], bm[4]);
                    a.gsb[bo] = gi;
                    if (HAS_H) a.hsb[bo] = gi * (a.gas.R * bm[4]) * (a.gas.K + 3 - a.gas.D);
                }
                break;
            case K_DVM_SYMMETRY:
            case K_SYMMETRY_PLANE:                      // :673-689
                if (phi > -DUGKS_VSMALL) { a.gsb[bo] = gOut; if (HAS_H) a.hsb[bo] = hOut; }
                break;
            default: break;
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------
// Symmetry patches: incoming DVs copy the patch values of their mirror DV
// (discreteVelocity.C:733-817).  Two kernels around an optional collective, as in the reference
// (fvDVM.C:375-454: every rank copies its DVs' patch values into dfContainer, MPI_Allgatherv, then every
// incoming DV reads its mirror partner's entry):
//   k_sym_pack   X[xrow[k]][j] = gSurf of local DV k on symmetry face j   (the dfContainer; rows the rank
//                does not own stay zero, so a SUM all-reduce over the ranks is the all-gather)
//   k_sym_apply  incoming DVs (xi.Sf0 <= 0, :775, Sf0 = first face of the patch) take X[xmir[axis][k]][j]
// xrow / xmir index X by GLOBAL DV id when mirror partners may live on another rank, by the local flat index
// (slab*L*Rs + i*Rs + r) when every partner is local (no collective then).  X: [nX][nsym] (+ the same for h).
struct SymPatch {
    int start, size, axis, xoff;     // boundary-face range, mirror axis (:777-779), first column in X
    double s0x, s0y, s0z;            // Sf of the first face of the patch
};

template <bool HAS_H>
__global__ void k_sym_pack(StepArgs a, const int* symface, int nsym, const int* xrow, int ndvpad, double* Xg, double* Xh) {
    const DevDV& dv = a.dv;
    const long long total = (long long)nsym * ndvpad;
    const int slabsz = dv.L * dv.Rs;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(t / ndvpad), k = (int)(t % ndvpad);
        const int row = xrow[k];
        if (row < 0) continue;                                   // padding entry: owns no velocity
        const int s = k / slabsz, rem = k % slabsz;
        const size_t src = ((size_t)s * a.m.nbf + symface[j]) * slabsz + rem;
        Xg[(size_t)row * nsym + j] = a.gsb[src];
        if (HAS_H) Xh[(size_t)row * nsym + j] = a.hsb[src];
    }
}

template <bool HAS_H>
__global__ void k_sym_apply(StepArgs a, SymPatch P, int nsym, const int* xrow, const int* xmir, int ndvpad,
                            const double* Xg, const double* Xh) {
    const DevDV& dv = a.dv;
    const long long total = (long long)P.size * ndvpad;
    const int slabsz = dv.L * dv.Rs;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(t / ndvpad), k = (int)(t % ndvpad);
        if (xrow[k] < 0) continue;
        const int s = k / slabsz, rem = k % slabsz, i = rem / dv.Rs, r = rem % dv.Rs;
        const int grow = s * dv.Rs + r;
        const double x = dv.tx[dv.row_cbase[grow] + i], y = dv.row_y[grow], z = dv.row_z[grow];
        if (dot_exact(x, y, z, P.s0x, P.s0y, P.s0z) <= 0) {      // :775 (first face of the patch)
            const int mrow = xmir[(size_t)P.axis * ndvpad + k];
            if (mrow < 0) continue;
            const size_t dst = ((size_t)s * a.m.nbf + P.start + j) * slabsz + rem;
            a.gsb[dst] = Xg[(size_t)mrow * nsym + P.xoff + j];
            if (HAS_H) a.hsb[dst] = Xh[(size_t)mrow * nsym + P.xoff + j];
        }
    }
}

// ---------------------------------------------------------------------------------
// boundary-face moments from the stored face values.  Wall faces: outgoing half only
// (the incoming half is rho_w * wall_cin, added in k_face_macros).
template <bool HAS_H>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
k_bnd_moments(StepArgs a) {
    __shared__ double txs[5 * NT_MAX];
    const DevDV& dv = a.dv;
    for (int k = threadIdx.x; k < 5 * dv.ntab; k += blockDim.x) txs[(k / dv.ntab) * NT_MAX + (k % dv.ntab)] = dv.tx[k];
    __syncthreads();
    int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    int nwr = dv.Rs >> 5;
    long long nitems = (long long)a.m.nbf * nwr;
    const int nm = a.nm;
    for (long long item = (long long)blockIdx.x * WARPS_PER_CTA + wib; item < nitems;
         item += (long long)gridDim.x * WARPS_PER_CTA) {
        int b = (int)(item / nwr), r = (int)(item % nwr) * 32 + lane;
        int grow = a.slab * dv.Rs + r;
        double y = dv.row_y[grow], z = dv.row_z[grow], wr = dv.row_w[grow];
        int cb = dv.row_cbase[grow];
        int kind = a.m.b_kind[b];
        const double* Sf = a.m.b_Sf + (size_t)b * 3;
        double sx = Sf[0], sy = Sf[1], sz = Sf[2];
        size_t bbase = dv_index(dv, a.slab, a.m.nbf, b, 0, r);
        double A[4] = {0, 0, 0, 0}, B[2] = {0, 0};
        for (int i = 0; i < dv_len(dv, a.slab); i++) {
            int t = cb + i;
            double x = txs[t];
            bool use = true;
            if (kind == K_MAXWELL_WALL) use = dot_exact(x, y, z, sx, sy, sz) > 0;
            if (use) {
                double val = a.gsb[bbase + (size_t)i * dv.Rs];
                A[0] = fma(txs[NT_MAX + t], val, A[0]);
                A[1] = fma(txs[2 * NT_MAX + t], val, A[1]);
                A[2] = fma(txs[3 * NT_MAX + t], val, A[2]);
                A[3] = fma(txs[4 * NT_MAX + t], val, A[3]);
                if (HAS_H) {
                    double vh = a.hsb[bbase + (size_t)i * dv.Rs];
                    B[0] = fma(txs[NT_MAX + t], vh, B[0]);
                    B[1] = fma(txs[2 * NT_MAX + t], vh, B[1]);
                }
            }
        }
        double v[16];
        expand_g(A, wr, y, z, v);
        if (HAS_H) {
            double u[NM_H];
            expand_h(B, wr, y, z, u);
            v[13] = u[0]; v[14] = u[1]; v[15] = u[2];
            double tot = warp_reduce16(v, lane);
            double t3 = warp_sum(u[3]);
            size_t slot = (size_t)2 * a.m.nif + b;
            int idx = reduce16_index(lane);
            if ((lane & 1) == 0) a.fslot[slot * nm + idx] += tot;
            if (lane == 0) a.fslot[slot * nm + 16] += t3;
        } else {
            v[13] = v[14] = v[15] = 0.0;
            double tot = warp_reduce16(v, lane);
            size_t slot = (size_t)2 * a.m.nif + b;
            int idx = reduce16_index(lane);
            if ((lane & 1) == 0 && idx < NM_G) a.fslot[slot * nm + idx] += tot;
        }
    }
}

// ---------------------------------------------------------------------------------
// wall constants: moments of the incoming half-space Maxwellian per unit rho_w
// (cin, [nbf][nm]) and inComingByRho (fvDVM.C:263-309), accumulated over slabs.
template <bool HAS_H>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
k_wall_constants(StepArgs a, double* cin, double* win) {
    __shared__ double txs[5 * NT_MAX];
    const DevDV& dv = a.dv;
    for (int k = threadIdx.x; k < 5 * dv.ntab; k += blockDim.x) txs[(k / dv.ntab) * NT_MAX + (k % dv.ntab)] = dv.tx[k];
    __syncthreads();
    int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    int nwr = dv.Rs >> 5;
    long long nitems = (long long)a.m.nbf * nwr;
    const int nm = a.nm;
    for (long long item = (long long)blockIdx.x * WARPS_PER_CTA + wib; item < nitems;
         item += (long long)gridDim.x * WARPS_PER_CTA) {
        int b = (int)(item / nwr), r = (int)(item % nwr) * 32 + lane;
        if (a.m.b_kind[b] != K_MAXWELL_WALL) continue;   // warp-uniform
        int grow = a.slab * dv.Rs + r;
        double y = dv.row_y[grow], z = dv.row_z[grow], wr = dv.row_w[grow];
        int cb = dv.row_cbase[grow];
        const double* Sf = a.m.b_Sf + (size_t)b * 3;
        double sx = Sf[0], sy = Sf[1], sz = Sf[2];
        const double* bm = a.bmac + (size_t)b * 5;
        double A[4] = {0, 0, 0, 0}, B[2] = {0, 0}, inb = 0.0;
        double hfac = (a.gas.R * bm[4]) * (a.gas.K + 3 - a.gas.D);
        for (int i = 0; i < dv_len(dv, a.slab); i++) {
            int t = cb + i;
            double x = txs[t];
            double phi = dot_exact(x, y, z, sx, sy, sz);
            if (phi <= 0) {                               // discreteVelocity.C:713
                double M = maxwell_by_rho(a.gas, x, y, z, bm[1], bm[2], bm[3], bm[4]);
                A[0] = fma(txs[NT_MAX + t], M, A[0]);
                A[1] = fma(txs[2 * NT_MAX + t], M, A[1]);
                A[2] = fma(txs[3 * NT_MAX + t], M, A[2]);
                A[3] = fma(txs[4 * NT_MAX + t], M, A[3]);
                if (HAS_H) {
                    B[0] = fma(txs[NT_MAX + t], M * hfac, B[0]);
                    B[1] = fma(txs[2 * NT_MAX + t], M * hfac, B[1]);
                }
                if (phi < 0) inb += -(txs[NT_MAX + t] * wr) * phi * M;   // fvDVM.C:291-299
            }
        }
        double v[16];
        expand_g(A, wr, y, z, v);
        double u[NM_H] = {0, 0, 0, 0};
        if (HAS_H) expand_h(B, wr, y, z, u);
        v[13] = u[0]; v[14] = u[1]; v[15] = u[2];
        double tot = warp_reduce16(v, lane);
        double t3 = warp_sum(u[3]);
        double tin = warp_sum(inb);
        int idx = reduce16_index(lane);
        if ((lane & 1) == 0 && idx < nm) cin[(size_t)b * nm + idx] += tot;
        if (HAS_H && lane == 0) cin[(size_t)b * nm + 16] += t3;
        if (lane == 0) win[b] += tin;
    }
}

// ---------------------------------------------------------------------------------
// face macros from the (all-reduced) moment slots; wall density; wall diagnostics.
// Owner-side + neighbour-side slot of every internal face, boundary slots copied: [nf][nm].  Run before the
// face all-reduce of a sharded step so that the collective carries nf instead of 2 nif + nbf slots.
__global__ void k_fold_fslot(StepArgs a, double* out) {
    const long long total = (long long)a.m.nf * a.nm;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int f = (int)(t / a.nm), k = (int)(t % a.nm);
        out[t] = f < a.m.nif ? a.fslot[(size_t)(2 * f) * a.nm + k] + a.fslot[(size_t)(2 * f + 1) * a.nm + k]
                             : a.fslot[((size_t)a.m.nif + f) * a.nm + k];
    }
}

// fsum: folded (and all-reduced) slots [nf][nm] of a sharded step, or null: the two sides are added here
__global__ void k_face_macros(StepArgs a, const double* fsum) {
    int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= a.m.nf) return;
    const int nm = a.nm;
    double M[NM_MAX];
    for (int k = 0; k < NM_MAX; k++) M[k] = 0.0;
    if (fsum) {
        for (int k = 0; k < nm; k++) M[k] = fsum[(size_t)f * nm + k];
    }
    if (f < a.m.nif) {
        if (!fsum) {
        const double* s0 = a.fslot + (size_t)(2 * f) * nm;
        const double* s1 = s0 + nm;
        for (int k = 0; k < nm; k++) M[k] = s0[k] + s1[k];
        }
    } else {
        int b = f - a.m.nif;
        if (!fsum) {
            const double* s0 = a.fslot + ((size_t)2 * a.m.nif + b) * nm;
            for (int k = 0; k < nm; k++) M[k] = s0[k];
        }
        if (a.m.b_kind[b] == K_MAXWELL_WALL) {
            const double* Sf = a.m.b_Sf + (size_t)b * 3;
            // outGoing = sum_{out} w (xi.Sf) gSurf (discreteVelocity.C:623-624);
            // rho_w = outGoing/|inComingByRho| (calculatedMaxwellFvPatchField.C:158)
            double out = Sf[0] * M[1] + Sf[1] * M[2] + Sf[2] * M[3];
            double rhow = out / fabs(a.wall_in[b]);
            a.bmac[(size_t)b * 5] = rhow;
            const double* ci = a.wall_cin + (size_t)b * nm;
            for (int k = 0; k < nm; k++) M[k] = fma(rhow, ci[k], M[k]);
        }
    }
    double out[MAC_N];
    macros_from_moments(a.gas, M, 0.5 * a.dt, out);     // fvDVM.C:493-522
    for (int k = 0; k < MAC_N; k++) a.fmac[(size_t)f * MAC_N + k] = out[k];
    if (a.fcoef) {
        // relaxation factor and Shakhov coefficients of updateGHsurf (discreteVelocity.C:867-870)
        const double hstep = 0.5 * a.dt;
        const double rf = hstep / (2.0 * out[5] + hstep);
        const EqCoef e = make_eq(a.gas, out, rf);
        double* fc = a.fcoef + (size_t)f * 12;
        fc[0] = e.Ux; fc[1] = e.Uy; fc[2] = e.Uz; fc[3] = e.a; fc[4] = e.pre;
        fc[5] = e.qx; fc[6] = e.qy; fc[7] = e.qz; fc[8] = 1.0 - rf; fc[9] = e.RT; fc[10] = 0.0; fc[11] = 0.0;
    }
    if (f >= a.m.nif) {
        int b = f - a.m.nif;
        double* wd = a.wall_diag + (size_t)b * 12;
        if (a.m.b_kind[b] == K_MAXWELL_WALL) {              // fvDVM.C:546-581
            // same moments, peculiar velocity taken about the WALL velocity
            const double* bm = a.bmac + (size_t)b * 5;
            double U[3] = {bm[1], bm[2], bm[3]};
            double U2 = U[0] * U[0] + U[1] * U[1] + U[2] * U[2];
            double trM2 = M[4] + M[7] + M[9];
            double UM1 = U[0] * M[1] + U[1] * M[2] + U[2] * M[3];
            double M2U[3] = {M[4] * U[0] + M[5] * U[1] + M[6] * U[2], M[5] * U[0] + M[7] * U[1] + M[8] * U[2],
                             M[6] * U[0] + M[8] * U[1] + M[9] * U[2]};
            double tau = out[5];
            double fq = 2.0 * tau / (2.0 * tau + 0.5 * a.dt * a.gas.Pr);
            double fs = 2.0 * tau / (2.0 * tau + 0.5 * a.dt);
            for (int i = 0; i < 3; i++) {
                double qg = M[10 + i] - 2.0 * M2U[i] + U2 * M[1 + i] - U[i] * trM2 + 2.0 * U[i] * UM1 - U[i] * U2 * M[0];
                double qh = M[14 + i] - U[i] * M[13];
                wd[i] = fq * 0.5 * (qg + qh);
            }
            wd[3] = fs * M[4]; wd[4] = fs * M[5]; wd[5] = fs * M[6];
            wd[6] = fs * M[5]; wd[7] = fs * M[7]; wd[8] = fs * M[8];
            wd[9] = fs * M[6]; wd[10] = fs * M[8]; wd[11] = fs * M[9];
        } else {
            for (int k = 0; k < 12; k++) wd[k] = 0.0;
        }
    }
}

// ---------------------------------------------------------------------------------
// stage 2.3 (wall incoming Maxwellian) and stage 4 on boundary faces.
// item = (boundary face, row-warp).  Runs after k_face_macros.
template <bool HAS_H>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
k_bnd_relax(StepArgs a) {
    __shared__ double txs[NT_MAX];
    const DevDV& dv = a.dv;
    for (int k = threadIdx.x; k < dv.ntab; k += blockDim.x) txs[k] = dv.tx[k];
    __syncthreads();
    int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    int nwr = dv.Rs >> 5;
    long long nitems = (long long)a.m.nbf * nwr;
    const double hstep = 0.5 * a.dt;
    for (long long item = (long long)blockIdx.x * WARPS_PER_CTA + wib; item < nitems;
         item += (long long)gridDim.x * WARPS_PER_CTA) {
        int b = (int)(item / nwr), r = (int)(item % nwr) * 32 + lane;
        int grow = a.slab * dv.Rs + r;
        double y = dv.row_y[grow], z = dv.row_z[grow];
        int cb = dv.row_cbase[grow];
        int kind = a.m.b_kind[b];
        const double* Sf = a.m.b_Sf + (size_t)b * 3;
        double sx = Sf[0], sy = Sf[1], sz = Sf[2];
        const double* bm = a.bmac + (size_t)b * 5;
        const double* mf = a.fmac + (size_t)(a.m.nif + b) * MAC_N;
        double rf = hstep / (2.0 * mf[5] + hstep);               // discreteVelocity.C:867
        EqCoef e = make_eq(a.gas, mf, rf);
        double omrf = 1.0 - rf;
        size_t bbase = dv_index(dv, a.slab, a.m.nbf, b, 0, r);
        for (int i = 0; i < dv_len(dv, a.slab); i++) {
            double x = txs[cb + i];
            size_t bo = bbase + (size_t)i * dv.Rs;
            double phi = dot_exact(x, y, z, sx, sy, sz);
            double g = a.gsb[bo];
            double h = HAS_H ? a.hsb[bo] : 0.0;
            if (kind == K_MAXWELL_WALL && phi <= 0) {           // :713-727
                g = bm[0] * maxwell_by_rho(a.gas, x, y, z, bm[1], bm[2], bm[3], bm[4]);
                h = g * (a.gas.R * bm[4]) * (a.gas.K + 3 - a.gas.D);
            }
            double gS, hS;
            shakhov_direct(a.gas, e, x, y, z, gS, hS);          // scaled by rf
            if (kind == K_SYMMETRY_PLANE) { g = fma(omrf, g, gS); h = fma(omrf, h, hS); }   // :880-881 [OF-lib]
            if (phi > 0) { g = fma(omrf, g, gS); h = fma(omrf, h, hS); }                    // :907-919
            if (kind == K_DVM_SYMMETRY) { g = fma(omrf, g, gS); h = fma(omrf, h, hS); }     // :922-930
            a.gsb[bo] = g;
            if (HAS_H) a.hsb[bo] = h;
        }
    }
}

// ---------------------------------------------------------------------------------
// stage 5 + the cell moments of stage 6: atomic-free gather over the cell's faces.
template <bool HAS_H>
__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
k_cell_update(StepArgs a) {
    extern __shared__ __align__(16) unsigned char dyn[];
    double* txs = reinterpret_cast<double*>(dyn);
    const DevDV& dv = a.dv;
    for (int k = threadIdx.x; k < 5 * dv.ntab; k += blockDim.x) txs[(k / dv.ntab) * NT_MAX + (k % dv.ntab)] = dv.tx[k];
    __syncthreads();
    int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    CellStage st = carve_stage(dyn + 5 * NT_MAX * 8 + wib * STAGE_BYTES);
    int nwr = dv.Rs >> 5;
    long long nitems = (long long)a.m.nc * nwr;
    const int nm = a.nm;
    for (long long item = (long long)blockIdx.x * WARPS_PER_CTA + wib; item < nitems;
         item += (long long)gridDim.x * WARPS_PER_CTA) {
        int c = (int)(item / nwr), r = (int)(item % nwr) * 32 + lane;
        int grow = a.slab * dv.Rs + r;
        double y = dv.row_y[grow], z = dv.row_z[grow], wr = dv.row_w[grow];
        int cb = dv.row_cbase[grow];
        int ne = stage_cell(a, c, lane, st);
        if (a.skip_small && ne <= FAST_NE) continue;   // warp-uniform
        // face value source: internal -> fbuf[face], boundary -> gsb (obase already points there)
        size_t base = dv_index(dv, a.slab, a.m.nc, c, 0, r);
        double dtv = a.dt / a.m.V[c];
        double A[4] = {0, 0, 0, 0}, B[2] = {0, 0};
        const size_t slabsz = (size_t)dv.L * dv.Rs;
        auto finish = [&](int i, double g0, double gb0, double h0, double hb0, double sumg, double sumh) {
            int t = cb + i;
            size_t idx = base + (size_t)i * dv.Rs;
            double gnew = (-1.0 / 3) * g0 + (4.0 / 3) * gb0 - sumg * dtv;   // discreteVelocity.C:937,952
            a.gt[idx] = gnew;
            A[0] = fma(txs[NT_MAX + t], gnew, A[0]);
            A[1] = fma(txs[2 * NT_MAX + t], gnew, A[1]);
            A[2] = fma(txs[3 * NT_MAX + t], gnew, A[2]);
            A[3] = fma(txs[4 * NT_MAX + t], gnew, A[3]);
            if (HAS_H) {
                double hnew = (-1.0 / 3) * h0 + (4.0 / 3) * hb0 - sumh * dtv;
                a.ht[idx] = hnew;
                B[0] = fma(txs[NT_MAX + t], hnew, B[0]);
                B[1] = fma(txs[2 * NT_MAX + t], hnew, B[1]);
            }
        };
        if (ne <= FAST_NE) {
            // per-thread y*Sy, z*Sz of every face (exact product, see dot_exact)
            double ySy[FAST_NE], zSz[FAST_NE];
#pragma unroll
            for (int j = 0; j < FAST_NE; j++) {
                ySy[j] = zSz[j] = 0.0;
                if (j < ne) { ySy[j] = __dmul_rn(y, st.geo[j * 9 + 7]); zSz[j] = __dmul_rn(z, st.geo[j * 9 + 8]); }
            }
            struct Vals { double g0, gb0, h0, hb0, gf[FAST_NE], hf[FAST_NE]; };
            auto load = [&](int i, Vals& o) {
                size_t ir = (size_t)i * dv.Rs + r;
                size_t idx = base + (size_t)i * dv.Rs;
                o.g0 = a.gt[idx]; o.gb0 = a.gb[idx];
                o.h0 = HAS_H ? a.ht[idx] : 0.0; o.hb0 = HAS_H ? a.hb[idx] : 0.0;
#pragma unroll
                for (int j = 0; j < FAST_NE; j++) {
                    o.gf[j] = 0.0; o.hf[j] = 0.0;
                    if (j < ne) {
                        if (st.kind[j] < 0) {
                            size_t fo = (size_t)st.face[j] * slabsz + ir;
                            o.gf[j] = a.fbuf_g[fo];
                            if (HAS_H) o.hf[j] = a.fbuf_h[fo];
                        } else {
                            size_t bo = (size_t)st.obase[j] + ir;
                            o.gf[j] = a.gsb[bo];
                            if (HAS_H) o.hf[j] = a.hsb[bo];
                        }
                    }
                }
            };
            auto compute = [&](int i, const Vals& o) {
                double x = txs[cb + i];
                double sumg = 0.0, sumh = 0.0;
#pragma unroll
                for (int j = 0; j < FAST_NE; j++) {
                    if (j < ne) {
                        double phi = __dadd_rn(__dadd_rn(__dmul_rn(x, st.geo[j * 9 + 6]), ySy[j]), zSz[j]);
                        double sphi = st.own[j] ? phi : -phi;       // discreteVelocity.C:952-955
                        sumg = fma(sphi, o.gf[j], sumg);
                        if (HAS_H) sumh = fma(sphi, o.hf[j], sumh);
                    }
                }
                finish(i, o.g0, o.gb0, o.h0, o.hb0, sumg, sumh);
            };
            Vals P, Q;
            load(0, P);
            for (int i = 0; i < dv_len(dv, a.slab); i += 2) {
                bool hasQ = i + 1 < dv_len(dv, a.slab);
                if (hasQ) load(i + 1, Q);
                compute(i, P);
                if (i + 2 < dv_len(dv, a.slab)) load(i + 2, P);
                if (hasQ) compute(i + 1, Q);
            }
        } else {
            for (int i = 0; i < dv_len(dv, a.slab); i++) {
                double x = txs[cb + i];
                size_t ir = (size_t)i * dv.Rs + r;
                size_t idx = base + (size_t)i * dv.Rs;
                double sumg = 0.0, sumh = 0.0;
                for (int j = 0; j < ne; j++) {
                    const double* G = st.geo + j * 9;
                    double phi = dot_exact(x, y, z, G[6], G[7], G[8]);
                    double gf, hf = 0.0;
                    if (st.kind[j] < 0) {
                        size_t fo = (size_t)st.face[j] * slabsz + ir;
                        gf = a.fbuf_g[fo];
                        if (HAS_H) hf = a.fbuf_h[fo];
                    } else {
                        size_t bo = (size_t)st.obase[j] + ir;
                        gf = a.gsb[bo];
                        if (HAS_H) hf = a.hsb[bo];
                    }
                    double sphi = st.own[j] ? phi : -phi;
                    sumg = fma(sphi, gf, sumg);
                    if (HAS_H) sumh = fma(sphi, hf, sumh);
                }
                finish(i, a.gt[idx], a.gb[idx], HAS_H ? a.ht[idx] : 0.0, HAS_H ? a.hb[idx] : 0.0, sumg, sumh);
            }
        }
        double v[16];
        expand_g(A, wr, y, z, v);
        double u[NM_H] = {0, 0, 0, 0};
        if (HAS_H) expand_h(B, wr, y, z, u);
        v[13] = u[0]; v[14] = u[1]; v[15] = u[2];
        double tot = warp_reduce16(v, lane);
        int idx16 = reduce16_index(lane);
        if ((lane & 1) == 0 && idx16 < nm) a.cslot[(size_t)c * nm + idx16] += tot;
        if (HAS_H) {
            double t3 = warp_sum(u[3]);
            if (lane == 0) a.cslot[(size_t)c * nm + 16] += t3;
        }
        __syncwarp();
    }
}

// Half-step coefficients of every cell for this step (discreteVelocity.C:393-404): relaxation factor
// rf = 1.5 dt / (2 tau + dt) and the Shakhov coefficients scaled by it, in the layout of the face records:
// Ux Uy Uz a pre qx qy qz omrf RT 0 0.  One evaluation per cell and step instead of one per warp that needs the
// cell's equilibrium (CTA pencils and the update kernel of fused slabs).
__global__ void k_cell_coef(StepArgs a) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.m.nc) return;
    const double* mc = a.cmac + (size_t)c * MAC_N;
    const double rf = 1.5 * a.dt / (2.0 * mc[5] + a.dt);
    const EqCoef e = make_eq(a.gas, mc, rf);
    double* o = a.ccoef + (size_t)c * 12;
    o[0] = e.Ux; o[1] = e.Uy; o[2] = e.Uz; o[3] = e.a; o[4] = e.pre; o[5] = e.qx; o[6] = e.qy; o[7] = e.qz;
    o[8] = 1.0 - rf; o[9] = e.RT; o[10] = 0.0; o[11] = 0.0;
}

// ---------------------------------------------------------------------------------
__global__ void k_cell_macros(StepArgs a) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.m.nc) return;
    double M[NM_MAX];
    for (int k = 0; k < NM_MAX; k++) M[k] = (k < a.nm) ? a.cslot[(size_t)c * a.nm + k] : 0.0;
    double out[MAC_N];
    macros_from_moments(a.gas, M, a.dt, out);            // fvDVM.C:694-727
    for (int k = 0; k < MAC_N; k++) a.cmac[(size_t)c * MAC_N + k] = out[k];
}

// U,T.correctBoundaryConditions() for zeroGradient patches (fvDVM.C:698-699) and
// updatePressureInOutBC (fvDVM.C:730-806).  bc[b] = U_bc | T_bc<<1; pres[b] = patch pressure
__global__ void k_bnd_macros(StepArgs a, const int* bc, const double* pres) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.m.nbf) return;
    int c = a.m.b_owner[b];
    const double* mc = a.cmac + (size_t)c * MAC_N;
    double* bm = a.bmac + (size_t)b * 5;
    int kind = a.m.b_kind[b];
    if (bc[b] & 1) { bm[1] = mc[1]; bm[2] = mc[2]; bm[3] = mc[3]; }
    if (bc[b] & 2) bm[4] = mc[4];
    if (kind == K_PRESSURE_IN || kind == K_PRESSURE_OUT) {
        const double R = a.gas.R;
        const int K = a.gas.K;
        const double* n = a.m.b_n + (size_t)b * 3;
        double pr = pres[b];
        double Ti = mc[4], rhoi = mc[0];
        double ai = sqrt(R * Ti * (K + 5) / (K + 3));
        double Un = mc[1] * n[0] + mc[2] * n[1] + mc[3] * n[2];
        double UnIn;
        if (kind == K_PRESSURE_IN) {
            bm[0] = pr / R / bm[4];
            UnIn = Un + (pr - rhoi * R * Ti) / rhoi / ai;
        } else {
            bm[0] = rhoi + (pr - rhoi * R * Ti) / ai / ai;
            bm[4] = pr / (R * rhoi);
            UnIn = Un + (rhoi * R * Ti - pr) / rhoi / ai;
        }
        for (int d = 0; d < 3; d++) bm[1 + d] = UnIn * n[d] + (mc[1 + d] - Un * n[d]);
    }
}

// "mixed" patches: initial Maxwellian of the boundary macros (discreteVelocity.C:312-344,1046-1060)
template <bool HAS_H>
__global__ void k_bnd_init_mixed(StepArgs a) {
    const DevDV& dv = a.dv;
    int slabsz = dv.L * dv.Rs;
    long long total = (long long)a.m.nbf * slabsz;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        int b = (int)(t / slabsz), rem = (int)(t % slabsz), i = rem / dv.Rs, r = rem % dv.Rs;
        if (a.m.b_kind[b] != K_MIXED || i >= dv_len(dv, a.slab)) continue;
        int grow = a.slab * dv.Rs + r;
        double x = dv.tx[dv.row_cbase[grow] + i], y = dv.row_y[grow], z = dv.row_z[grow];
        const double* bm = a.bmac + (size_t)b * 5;
        double g = bm[0] * maxwell_by_rho(a.gas, x, y, z, bm[1], bm[2], bm[3], bm[4]);
        size_t bo = ((size_t)a.slab * a.m.nbf + b) * slabsz + rem;
        a.gsb[bo] = g;
        if (HAS_H) a.hsb[bo] = (a.gas.K + 3 - a.gas.D) * a.gas.R * bm[4] * g;
    }
}

// tau of the initial macro state (fvDVM.C:1073) and first face velocities (:1074)
__global__ void k_init_tau(StepArgs a) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.m.nc) return;
    double* mc = a.cmac + (size_t)c * MAC_N;
    mc[5] = dugks_tau(a.gas, mc[4], mc[0]);
}

// fvDVM::getCoNum (fvDVM.C:1111-1119): out[0] = max, out[1] = sum over internal faces.
// Two deterministic stages: every block reduces a fixed set of faces to part[2*blockIdx.x..], the last
// launch (one block) folds the partials in index order.
__global__ void k_courant(StepArgs a, double sqrtD_xiMax, double* part) {
    __shared__ double smax[32], ssum[32];
    double mx = 0.0, sm = 0.0;
    for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < a.m.nif; f += gridDim.x * blockDim.x) {
        const double* mf = a.fmac + (size_t)f * MAC_N;
        double v = a.m.dcoef_int[f] * (sqrt(mf[1] * mf[1] + mf[2] * mf[2] + mf[3] * mf[3]) + sqrtD_xiMax);
        mx = fmax(mx, v);
        sm += v;
    }
    for (int o = 16; o > 0; o >>= 1) {
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        sm += __shfl_xor_sync(0xffffffffu, sm, o);
    }
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) { smax[w] = mx; ssum[w] = sm; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int nw = blockDim.x >> 5;
        for (int k = 1; k < nw; k++) { mx = fmax(mx, smax[k]); sm += ssum[k]; }
        part[2 * blockIdx.x] = mx; part[2 * blockIdx.x + 1] = sm;
    }
}
__global__ void k_courant_fold(const double* part, int nblocks, double* out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double mx = 0.0, sm = 0.0;
    for (int k = 0; k < nblocks; k++) { mx = fmax(mx, part[2 * k]); sm += part[2 * k + 1]; }
    out[0] = mx; out[1] = sm;
}

// Convergence monitor (dugksFoam.C:88-107): per block the six sums over a fixed set of cells
// { |T - Told|, T, |rho - rhoOld|, rho, |U - Uold|, |U| } -> part[6 * blockIdx.x ..], and the snapshot
// old[c] = { rho, Ux, Uy, Uz, T } is replaced by the current macros in the same pass.  k_convergence_fold
// adds the partials in index order (deterministic).
__global__ void k_convergence(StepArgs a, double* old, double* part) {
    __shared__ double sh[6][32];
    double s[6] = {0, 0, 0, 0, 0, 0};
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < a.m.nc; c += gridDim.x * blockDim.x) {
        const double* mc = a.cmac + (size_t)c * MAC_N;     // rho, U(3), T, tau, q(3)
        double* o = old + (size_t)c * 5;
        const double rho = mc[0], ux = mc[1], uy = mc[2], uz = mc[3], T = mc[4];
        const double dx = ux - o[1], dy = uy - o[2], dz = uz - o[3];
        s[0] += fabs(T - o[4]); s[1] += T;
        s[2] += fabs(rho - o[0]); s[3] += rho;
        s[4] += sqrt(dx * dx + dy * dy + dz * dz); s[5] += sqrt(ux * ux + uy * uy + uz * uz);
        o[0] = rho; o[1] = ux; o[2] = uy; o[3] = uz; o[4] = T;
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 6; k++) {
        for (int off = 16; off > 0; off >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], off);
        if (lane == 0) sh[k][w] = s[k];
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        double t = 0.0;
        for (int k = 0; k < (int)(blockDim.x >> 5); k++) t += sh[threadIdx.x][k];
        part[6 * blockIdx.x + threadIdx.x] = t;
    }
}
__global__ void k_convergence_fold(const double* part, int nblocks, double* out) {
    if (blockIdx.x != 0 || threadIdx.x >= 6) return;
    double t = 0.0;
    for (int k = 0; k < nblocks; k++) t += part[6 * k + threadIdx.x];
    out[threadIdx.x] = t;
}
__global__ void k_convergence_init(StepArgs a, double* old) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= a.m.nc) return;
    const double* mc = a.cmac + (size_t)c * MAC_N;
    double* o = old + (size_t)c * 5;
    o[0] = mc[0]; o[1] = mc[1]; o[2] = mc[2]; o[3] = mc[3]; o[4] = mc[4];
}

// dugks_set_boundary_macros: scatter the caller's boundary fields (any may be null) into bmac.  Only values
// the caller owns are taken: components the library evolves itself keep their device values — U / T of
// zeroGradient patch fields (bc bits, fvDVM.C:698-699), rho_w of Maxwell walls
// (calculatedMaxwellFvPatchField.C:158), and on pressure patches U and rho (and T on outlets), which
// updatePressureInOutBC recomputes every step (fvDVM.C:743-804).
__global__ void k_set_bmac(StepArgs a, const int* bc, const double* rho_b, const double* U_b, const double* T_b) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.m.nbf) return;
    double* bm = a.bmac + (size_t)b * 5;
    const int kind = a.m.b_kind[b];
    const bool pres = kind == K_PRESSURE_IN || kind == K_PRESSURE_OUT;
    if (rho_b && !pres && kind != K_MAXWELL_WALL) bm[0] = rho_b[b];
    if (U_b && !pres && !(bc[b] & 1)) { bm[1] = U_b[3 * b]; bm[2] = U_b[3 * b + 1]; bm[3] = U_b[3 * b + 2]; }
    if (T_b && kind != K_PRESSURE_OUT && !(bc[b] & 2)) bm[4] = T_b[b];
}
