// dugks_device.cuh — device-side data model and math shared by all kernels.
//
// Layout (DESIGN.md "Data layout in HBM"): the rank-local discrete velocities are
// arranged as ROWS (all DVs sharing xi_y, xi_z and an ix-chunk) of L inner points
// (consecutive ix).  32*m rows form a SLAB; one kernel launch processes one slab
// over the whole mesh.  Every per-(cell|face, DV) array is
//     a[slab][cell][i][r]      i = inner index 0..L-1,  r = row in slab 0..Rs-1
// so the innermost (fastest) index is a discrete-velocity index and a warp's 32
// lanes read 32 consecutive doubles; a thread owns one row and walks i serially,
// which makes every ix-dependent quantity warp-uniform and turns the velocity
// moments into in-register sums (sum factorisation over the tensor-product grid).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define DUGKS_VSMALL 1.0e-300
#define DUGKS_PI 3.14159265358979323846

// moment vector of one face side / cell: 13 raw moments of g and 4 of h
//  0: 1   1-3: xi_i   4-9: xx xy xz yy yz zz   10-12: xi_i |xi|^2   13: h   14-16: h xi_i
#define NM_G 13
#define NM_H 4
#define NM_MAX 17

#define MAX_L 32
#define MAX_CELL_FACES 24 // faces per cell staged in shared memory
#define ACC_FACES 6      // internal faces with register accumulators per pass

enum {
    K_ZERO_GRADIENT = 0, K_MIXED = 1, K_MAXWELL_WALL = 2, K_FAR_FIELD = 3,
    K_DVM_SYMMETRY = 4, K_SYMMETRY_PLANE = 5, K_PRESSURE_IN = 6, K_PRESSURE_OUT = 7
};

struct DevGas {
    double R, omega, Tref, muRef, Pr;
    int K, D;
};

struct DevMesh {
    int nc, nif, nbf, nf;
    const int* cell_off;     // [nc+1] CSR over (cell, face) entries, internal faces first
    const int* cell_nint;    // [nc]   number of internal-face entries of the cell
    const int* e_other;      // [ne]   other cell (internal) or -1-b (boundary face b)
    const int* e_face;       // [ne]   face index f (internal) or nif+b
    const int* e_owner;      // [ne]   1 if the cell owns the face
    const double* e_geo;     // [ne][9] G(3) LS vector, r(3) = Cf - C_cell, Sf(3) owner-oriented
    const double* e_geo12;   // [ne][12] same data as 16-byte pairs: G0 G1 | G2 Sx | r0 r1 | r2 Sy | Sz invdc | 0 0
    const double* V;         // [nc]
    const int* b_owner;      // [nbf]
    const int* b_kind;       // [nbf]
    const double* b_n;       // [nbf][3] Sf/|Sf|
    const double* b_invdc;   // [nbf] 1/deltaCoeffs
    const double* b_Sf;      // [nbf][3]
    const double* b_r;       // [nbf][3] Cf - C_owner
    const double* dcoef_int; // [nif] deltaCoeffs on internal faces (Courant number)
};

struct DevDV {
    int L, Rs, nslab, ntab, hasH;
    int Lt;                  // > 0: the LAST slab holds short rows of Lt points (stride stays L), see dv_len
    int tabw;                // table width: largest span (in table entries) any warp needs
    const double* tx;        // [5][ntab]: xi_x, w_x, w_x xi_x, w_x xi_x^2, w_x xi_x^3
    const double* row_y;     // [nslab*Rs] xi_y of the row
    const double* row_z;     // [nslab*Rs] xi_z
    const double* row_w;     // [nslab*Rs] w_z*w_y (0 for padding rows)
    const int* row_cbase;    // [nslab*Rs] first table index of the row's chunk
};

// Points per row of a slab.  All slabs use rows of L points except, optionally, the last one: when a
// rank's row count leaves a nearly empty last slab (98 rows = 3 slabs + 2 rows at 8 GPUs), those few
// rows are cut into ix-chunks of Lt points so that they fill the 32 lanes of ONE short slab instead of
// idling 30 of them for L points.  Array strides are L everywhere; only loop counts use dv_len.
__host__ __device__ __forceinline__ int dv_len(const DevDV& dv, int slab) {
    return (dv.Lt > 0 && slab == dv.nslab - 1) ? dv.Lt : dv.L;
}

// ---- macro arrays: 9 doubles per cell / face: rho, Ux,Uy,Uz, T, tau, qx,qy,qz
#define MAC_N 9

__host__ __device__ inline double dugks_tau(const DevGas& g, double T, double rho) {
    // fvDVM::updateTau, fvDVM.C:816
    return g.muRef * exp(g.omega * log(T / g.Tref)) / rho / T / g.R;
}

// exact evaluation order of OpenFOAM's vector dot product x*Sx + y*Sy + z*Sz without
// FMA contraction (the upwind side is decided by its sign)
__device__ __forceinline__ double dot_exact(double x, double y, double z, double sx, double sy, double sz) {
    return __dadd_rn(__dadd_rn(__dmul_rn(x, sx), __dmul_rn(y, sy)), __dmul_rn(z, sz));
}

// Coefficients of the Shakhov equilibrium for one macro state, scaled by `scale`
// (the relaxation factor), discreteVelocity.C:1033-1043.
struct EqCoef {
    double a;        // 1/(R T)
    double pre;      // scale * rho / (2 pi R T)^(D/2)
    double qx, qy, qz; // (1-Pr) q / (5 rho (RT)^2)
    double RT;
    double Ux, Uy, Uz;
};

__host__ __device__ inline EqCoef make_eq(const DevGas& g, const double* m, double scale) {
    EqCoef e;
    double rho = m[0], T = m[4];
    double RT = g.R * T;
    e.RT = RT;
    e.a = 1.0 / RT;
    double s = sqrt(2.0 * DUGKS_PI * RT);
    double p = (g.D == 3) ? s * s * s : ((g.D == 2) ? s * s : s);
    e.pre = scale * rho / p;
    double c = (1.0 - g.Pr) / (5.0 * rho * RT * RT);
    e.qx = c * m[6]; e.qy = c * m[7]; e.qz = c * m[8];
    e.Ux = m[1]; e.Uy = m[2]; e.Uz = m[3];
    return e;
}

// Direct (non-tabulated) Shakhov g,h at one velocity, scaled; used by the small
// boundary kernels.
__device__ inline void shakhov_direct(const DevGas& g, const EqCoef& e, double x, double y, double z,
                                      double& gS, double& hS) {
    double cx = x - e.Ux, cy = y - e.Uy, cz = z - e.Uz;
    double cc = (cx * cx + cy * cy + cz * cz) * e.a;
    double cq = cx * e.qx + cy * e.qy + cz * e.qz;
    double gM = e.pre * exp(-0.5 * cc);
    double kd = (double)(g.K + 3 - g.D);
    gS = (1.0 + cq * (cc - g.D - 2.0)) * gM;
    hS = (kd + cq * ((cc - g.D) * kd - 2.0 * g.K)) * gM * e.RT;
}

// Maxwellian divided by rho, discreteVelocity.C:1063-1075
__device__ inline double maxwell_by_rho(const DevGas& g, double x, double y, double z, double Ux, double Uy,
                                        double Uz, double T) {
    double RT = g.R * T;
    double s = sqrt(2.0 * DUGKS_PI * RT);
    double p = (g.D == 3) ? s * s * s : ((g.D == 2) ? s * s : s);
    double cx = Ux - x, cy = Uy - y, cz = Uz - z;
    return 1.0 / p * exp(-(cx * cx + cy * cy + cz * cz) / (2.0 * RT));
}

// Macros from the 17 raw moments (fvDVM.C:493-522 faces, :694-727 cells).
// qfac_dt: the dt that multiplies Pr in the bar->original heat-flux factor
// (0.5*dt at faces :522, dt in cells :727).
__host__ __device__ inline void macros_from_moments(const DevGas& g, const double* M, double qfac_dt, double* out) {
    double rho = M[0];
    double Ux = M[1] / rho, Uy = M[2] / rho, Uz = M[3] / rho;
    double rE = 0.5 * ((M[4] + M[7] + M[9]) + M[13]);
    double U2 = Ux * Ux + Uy * Uy + Uz * Uz;
    double T = (rE - 0.5 * rho * U2) / ((g.K + 3) / 2.0 * g.R * rho);
    double tau = dugks_tau(g, T, rho);
    // q_i = 1/2 sum w c_i (|c|^2 g + h), c = xi - U, expanded in raw moments
    double trM2 = M[4] + M[7] + M[9];
    double UM1 = Ux * M[1] + Uy * M[2] + Uz * M[3];
    double M2U[3] = {M[4] * Ux + M[5] * Uy + M[6] * Uz, M[5] * Ux + M[7] * Uy + M[8] * Uz,
                     M[6] * Ux + M[8] * Uy + M[9] * Uz};
    double U[3] = {Ux, Uy, Uz};
    double fac = 2.0 * tau / (2.0 * tau + qfac_dt * g.Pr);
    out[0] = rho; out[1] = Ux; out[2] = Uy; out[3] = Uz; out[4] = T; out[5] = tau;
    for (int i = 0; i < 3; i++) {
        double qg = M[10 + i] - 2.0 * M2U[i] + U2 * M[1 + i] - U[i] * trM2 + 2.0 * U[i] * UM1 - U[i] * U2 * rho;
        double qh = M[14 + i] - U[i] * M[13];
        out[6 + i] = fac * 0.5 * (qg + qh);
    }
}

// level-2 expansion of the in-row sums A_a = sum_i w_x xi_x^a val into the 13 raw
// moments of g (times the row weight); y,z = xi_y, xi_z of the row.
__device__ __forceinline__ void expand_g(const double A[4], double wr, double y, double z, double out[NM_G]) {
    double A0 = wr * A[0], A1 = wr * A[1], A2 = wr * A[2], A3 = wr * A[3];
    double yz2 = y * y + z * z;
    out[0] = A0;
    out[1] = A1; out[2] = y * A0; out[3] = z * A0;
    out[4] = A2; out[5] = y * A1; out[6] = z * A1;
    out[7] = y * y * A0; out[8] = y * z * A0; out[9] = z * z * A0;
    out[10] = A3 + yz2 * A1;
    double t = A2 + yz2 * A0;
    out[11] = y * t; out[12] = z * t;
}
__device__ __forceinline__ void expand_h(const double B[2], double wr, double y, double z, double out[NM_H]) {
    double B0 = wr * B[0], B1 = wr * B[1];
    out[0] = B0; out[1] = B1; out[2] = y * B0; out[3] = z * B0;
}

// Warp reduction of 16 values per lane: after the call every lane with (lane&1)==0
// holds, in the return value, the warp total of value index (lane>>1).
// 16 double-shuffles instead of 16*5.
__device__ __forceinline__ double warp_reduce16(double v[16], int lane) {
    const unsigned full = 0xffffffffu;
    bool up;
    up = lane & 16;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        double mine = up ? v[k + 8] : v[k];
        double theirs = up ? v[k] : v[k + 8];
        v[k] = mine + __shfl_xor_sync(full, theirs, 16);
    }
    up = lane & 8;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        double mine = up ? v[k + 4] : v[k];
        double theirs = up ? v[k] : v[k + 4];
        v[k] = mine + __shfl_xor_sync(full, theirs, 8);
    }
    up = lane & 4;
#pragma unroll
    for (int k = 0; k < 2; k++) {
        double mine = up ? v[k + 2] : v[k];
        double theirs = up ? v[k] : v[k + 2];
        v[k] = mine + __shfl_xor_sync(full, theirs, 4);
    }
    up = lane & 2;
    {
        double mine = up ? v[1] : v[0];
        double theirs = up ? v[0] : v[1];
        v[0] = mine + __shfl_xor_sync(full, theirs, 2);
    }
    v[0] += __shfl_xor_sync(full, v[0], 1);
    return v[0];
}
// value index held by `lane` after warp_reduce16: bit4->8, bit3->4, bit2->2, bit1->1
__device__ __forceinline__ int reduce16_index(int lane) { return (lane >> 1) & 15; }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
