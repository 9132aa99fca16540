/*
 * dugks_oracle.c — CPU restatement of dugksFoam's per-time-step discrete-velocity
 * update.  TEST INFRASTRUCTURE ONLY: nothing under dugksfoam_b200/ may import,
 * link or execute this file; it is the checker for tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs.
 *
 * PARITY UNPINNED: the reference ships no tests, golden fields or asserted
 * numbers for this path and cannot be compiled here (needs
 * OpenFOAM + MPI, neither is in this image).  This file follows the reference
 * source stage by stage, DV-outermost and field-at-a-time exactly like
 * fvDVM / discreteVelocity do; it is validated by the physical identities in
 * tests/test_oracle.py (mass conservation, equilibrium steady state,
 * full-vs-half symmetry, serial vs velocity-partitioned agreement) and by the
 * reference's only golden vector on this path, the shipped Gauss-Hermite
 * Xis/weights (tests/golden/).
 *
 * All `file:line` citations are relative to /root/reference/src/fvDVM/.
 * OpenFOAM library behaviour that is not in the reference tree is marked
 * [OF-lib].
 *
 * Build: see oracle/Makefile  (gcc -O2 -fopenmp -ffp-contract=off -shared).
 */
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/dugks.h"

#define VSMALL 1.0e-300 /* [OF-lib] double-precision build */
#define PI 3.14159265358979323846 /* constant::mathematical::pi */

typedef struct oracle {
    /* mesh (copied) */
    int nc, nif, nbf, nf, D;
    int32_t *owner, *neigh;
    double *C, *V, *Cf, *Sf, *ownLs, *neiLs, *patchLs, *dcoef;
    int npatch;
    dugks_patch_t *patch;
    /* gas */
    double R, omega, Tref, muRef, Pr;
    int K;
    double xiMax;
    /* global DV set, fvDVM/fvDVM.C:140-220 */
    int nxi, n1d;
    double *xi; /* [nxi][3] */
    double *w;  /* [nxi] */
    int32_t *symX, *symY, *symZ;
    /* virtual ranks of the -dvParallel decomposition, fvDVM/fvDVM.C:228-260 */
    int P, partition;
    int32_t *rank_of; /* [nxi] */
    /* per-DV fields, discreteVelocity/discreteVelocity.H:86-103 */
    double *gTilde, *hTilde, *gBarP, *hBarP; /* [nxi][nc] */
    double *gSurf, *hSurf;                   /* [nxi][nf] */
    double *gGrad, *hGrad;                   /* [nxi][nc][3] */
    double *gamG, *gamH;                     /* [nxi][nbf] fixedGradient gradient() */
    /* macros */
    double *rho, *U, *T, *q, *tau;           /* cells */
    double *rhoS, *US, *TS, *qS, *tauS;      /* faces */
    double *rho_b, *U_b, *T_b;               /* boundary values of rho,U,T */
    double *inByRho, *outGoing;              /* calculatedMaxwell patch data */
    double *qWall, *stressWall;              /* [nbf][3], [nbf][9] */
    double *part;                            /* scratch for rank partials */
    double *Told, *rhoOld, *Uold;            /* convergence monitor snapshot, createFields.H:40-82 */
    int grad_limiter;                        /* 0: gradSchemes leastSquares; 1: VenkatakrishnanLimited leastSquares k, as meant */
    double limiter_k;
    long steps;
} oracle_t;

static void *xcalloc(size_t n, size_t s) {
    void *p = calloc(n ? n : 1, s);
    if (!p) { fprintf(stderr, "oracle: out of memory (%zu x %zu)\n", n, s); abort(); }
    return p;
}
static void *xdup(const void *src, size_t bytes) {
    void *p = xcalloc(bytes ? bytes : 1, 1);
    if (bytes) memcpy(p, src, bytes);
    return p;
}

static inline double dot3(const double *a, const double *b) {
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}

/* discreteVelocity.C:1016-1044 equilibriumShakhov, one element */
static inline void shakhov(const oracle_t *o, const double *xi, double rho, const double *U,
                           double T, const double *q, double *gEq, double *hEq) {
    const double R = o->R, Pr = o->Pr;
    const int D = o->D, K = o->K;
    double d0 = U[0] - xi[0], d1 = U[1] - xi[1], d2 = U[2] - xi[2];
    double cSqrByRT = (d0 * d0 + d1 * d1 + d2 * d2) / (R * T);             /* :1033-1034 */
    double cqBy5pRT = ((xi[0] - U[0]) * q[0] + (xi[1] - U[1]) * q[1] + (xi[2] - U[2]) * q[2]) /
                      (5.0 * rho * R * T * R * T);                           /* :1036-1037 */
    double gEqBGK = rho / pow(sqrt(2.0 * PI * R * T), D) * exp(-cSqrByRT / 2.0); /* :1039-1040 */
    *gEq = (1.0 + (1.0 - Pr) * cqBy5pRT * (cSqrByRT - D - 2.0)) * gEqBGK; /* :1042 */
    *hEq = ((K + 3.0 - D) + (1.0 - Pr) * cqBy5pRT * ((cSqrByRT - D) * (K + 3.0 - D) - 2 * K)) *
           gEqBGK * R * T;                                                   /* :1043 */
}

/* discreteVelocity.C:1063-1075 equilibriumMaxwellByRho */
static inline double maxwellByRho(const oracle_t *o, const double *xi, const double *U, double T) {
    double d0 = U[0] - xi[0], d1 = U[1] - xi[1], d2 = U[2] - xi[2];
    return 1.0 / pow(sqrt(2.0 * PI * o->R * T), o->D) *
           exp(-(d0 * d0 + d1 * d1 + d2 * d2) / (2.0 * o->R * T));
}

/* fvDVM.C:808-817 updateTau */
static inline double tau_of(const oracle_t *o, double T, double rho) {
    return o->muRef * exp(o->omega * log(T / o->Tref)) / rho / T / o->R;
}

/* fvDVM.C:140-220: tensor-product grid, weights, mirror ids */
static void build_dvs(oracle_t *o, const dugks_dvset_t *dv) {
    int n = dv->nXiPerDim, D = o->D;
    o->n1d = n;
    int nx = n, ny = (D >= 2) ? n : 1, nz = (D == 3) ? n : 1;
    o->nxi = nx * ny * nz;
    o->xi = xcalloc((size_t)o->nxi * 3, sizeof(double));
    o->w = xcalloc(o->nxi, sizeof(double));
    o->symX = xcalloc(o->nxi, sizeof(int32_t));
    o->symY = xcalloc(o->nxi, sizeof(int32_t));
    o->symZ = xcalloc(o->nxi, sizeof(int32_t));
    int i = 0;
    for (int iz = 0; iz < nz; iz++)
        for (int iy = 0; iy < ny; iy++)
            for (int ix = 0; ix < nx; ix++) {
                double wt;
                if (D == 3) wt = dv->weights[iz] * dv->weights[iy] * dv->weights[ix]; /* :158 */
                else if (D == 2) wt = dv->weights[iy] * dv->weights[ix] * 1;         /* :187 */
                else wt = dv->weights[ix] * 1 * 1;                                     /* :211 */
                o->w[i] = wt;
                o->xi[3 * i + 0] = dv->Xis[ix];
                o->xi[3 * i + 1] = (D >= 2) ? dv->Xis[iy] : 0.0;
                o->xi[3 * i + 2] = (D == 3) ? dv->Xis[iz] : 0.0;
                if (D == 3) {                                                          /* :162-164 */
                    o->symX[i] = iz * ny * nx + iy * nx + (nx - ix - 1);
                    o->symY[i] = iz * ny * nx + (ny - iy - 1) * nx + ix;
                    o->symZ[i] = (nz - iz - 1) * ny * nx + iy * nx + ix;
                } else if (D == 2) {                                                   /* :191-193 */
                    o->symX[i] = iy * nx + (nx - ix - 1);
                    o->symY[i] = (ny - iy - 1) * nx + ix;
                    o->symZ[i] = 0;
                } else {                                                               /* :215-217 */
                    o->symX[i] = nx - ix - 1;
                    o->symY[i] = 0;
                    o->symZ[i] = 0;
                }
                i++;
            }
}

/* rank that owns global DV gid.  partition 1 = fvDVM.C:228-260 (gid = i*P + rank);
 * partition 0 = contiguous blocks (what the CUDA library uses by default). */
static void build_partition(oracle_t *o) {
    o->rank_of = xcalloc(o->nxi, sizeof(int32_t));
    int P = o->P, n = o->nxi;
    for (int g = 0; g < n; g++) {
        if (o->partition == 1) o->rank_of[g] = g % P;
        else {
            int base = n / P, rem = n % P; /* first rem ranks hold base+1 */
            int cut = rem * (base + 1);
            o->rank_of[g] = (g < cut) ? g / (base + 1) : rem + (g - cut) / (base ? base : 1);
        }
    }
}

/* fvDVM.C:263-309 setCalculatedMaxwellRhoBC */
static void set_wall_incoming(oracle_t *o) {
    for (int p = 0; p < o->npatch; p++) {
        if (o->patch[p].kind != DUGKS_PATCH_MAXWELL_WALL) continue;
        for (int j = 0; j < o->patch[p].size; j++) {
            int b = o->patch[p].start + j;
            const double *Sf = o->Sf + 3 * (size_t)(o->nif + b);
            double tot = 0.0;
            for (int r = 0; r < o->P; r++) {
                double acc = 0.0;
                for (int k = 0; k < o->nxi; k++) {
                    if (o->rank_of[k] != r) continue;
                    const double *xi = o->xi + 3 * k;
                    double phi = dot3(xi, Sf);
                    if (phi < 0) /* :291 */
                        acc += -o->w[k] * phi * maxwellByRho(o, xi, o->U_b + 3 * b, o->T_b[b]);
                }
                tot += acc; /* :304-305 allreduce */
            }
            o->inByRho[b] = tot;
        }
    }
}

/* fvDVM.C:730-806 updatePressureInOutBC */
static void update_pressure_bc(oracle_t *o) {
    const double R = o->R;
    const int K = o->K;
    for (int p = 0; p < o->npatch; p++) {
        int kind = o->patch[p].kind;
        if (kind != DUGKS_PATCH_PRESSURE_IN && kind != DUGKS_PATCH_PRESSURE_OUT) continue;
        double pr = o->patch[p].pressure;
        for (int j = 0; j < o->patch[p].size; j++) {
            int b = o->patch[p].start + j;
            int own = o->owner[o->nif + b];
            const double *Sf = o->Sf + 3 * (size_t)(o->nif + b);
            double magSf = sqrt(dot3(Sf, Sf));
            double norm[3] = {Sf[0] / magSf, Sf[1] / magSf, Sf[2] / magSf};
            const double *Ui = o->U + 3 * own;
            double Ti = o->T[own], rhoi = o->rho[own];
            double ai = sqrt(R * Ti * (K + 5) / (K + 3)); /* :765,:792 */
            double Un = dot3(Ui, norm);
            double UnIn;
            if (kind == DUGKS_PATCH_PRESSURE_IN) {
                double Tin = o->T_b[b];
                o->rho_b[b] = pr / R / Tin;                        /* :758 */
                UnIn = Un + (pr - rhoi * R * Ti) / rhoi / ai;      /* :770 */
            } else {
                o->rho_b[b] = rhoi + (pr - rhoi * R * Ti) / ai / ai; /* :795 */
                o->T_b[b] = pr / (R * rhoi);                       /* :796 */
                UnIn = Un + (rhoi * R * Ti - pr) / rhoi / ai;      /* :801 */
            }
            for (int d = 0; d < 3; d++)
                o->U_b[3 * b + d] = UnIn * norm[d] + (Ui[d] - Un * norm[d]); /* :771,:802 */
        }
    }
}

/* U,T.correctBoundaryConditions() (fvDVM.C:698-699) for zeroGradient patches [OF-lib] */
static void correct_macro_bcs(oracle_t *o) {
    for (int p = 0; p < o->npatch; p++)
        for (int j = 0; j < o->patch[p].size; j++) {
            int b = o->patch[p].start + j;
            int own = o->owner[o->nif + b];
            if (o->patch[p].U_bc == DUGKS_BC_ZERO_GRADIENT)
                for (int d = 0; d < 3; d++) o->U_b[3 * b + d] = o->U[3 * own + d];
            if (o->patch[p].T_bc == DUGKS_BC_ZERO_GRADIENT) o->T_b[b] = o->T[own];
        }
}

oracle_t *oracle_create(const dugks_mesh_t *m, const dugks_patch_t *patches, int npatch,
                        const dugks_dvset_t *dv, const dugks_gas_t *gas, int nranks, int partition,
                        const double *rho, const double *U, const double *T, const double *rho_b,
                        const double *U_b, const double *T_b) {
    oracle_t *o = xcalloc(1, sizeof(*o));
    o->nc = m->nCells; o->nif = m->nInternalFaces; o->nbf = m->nBoundaryFaces;
    o->nf = o->nif + o->nbf; o->D = m->nSolutionD;
    int nc = o->nc, nf = o->nf, nif = o->nif, nbf = o->nbf;
    o->owner = xdup(m->owner, sizeof(int32_t) * nf);
    o->neigh = xdup(m->neighbour, sizeof(int32_t) * nif);
    o->C = xdup(m->C, sizeof(double) * 3 * nc);
    o->V = xdup(m->V, sizeof(double) * nc);
    o->Cf = xdup(m->Cf, sizeof(double) * 3 * nf);
    o->Sf = xdup(m->Sf, sizeof(double) * 3 * nf);
    o->ownLs = xdup(m->ownLs, sizeof(double) * 3 * nif);
    o->neiLs = xdup(m->neiLs, sizeof(double) * 3 * nif);
    o->patchLs = xdup(m->patchLs, sizeof(double) * 3 * nbf);
    o->dcoef = xdup(m->deltaCoeffs, sizeof(double) * nf);
    o->npatch = npatch;
    o->patch = xdup(patches, sizeof(dugks_patch_t) * npatch);
    o->R = gas->R; o->omega = gas->omega; o->Tref = gas->Tref; o->muRef = gas->muRef;
    o->Pr = gas->Pr; o->K = gas->KInner; o->xiMax = dv->xiMax;
    o->P = nranks < 1 ? 1 : nranks; o->partition = partition;
    build_dvs(o, dv);
    build_partition(o);
    size_t nx = o->nxi;
    o->gTilde = xcalloc(nx * nc, sizeof(double)); o->hTilde = xcalloc(nx * nc, sizeof(double));
    o->gBarP = xcalloc(nx * nc, sizeof(double));  o->hBarP = xcalloc(nx * nc, sizeof(double));
    o->gSurf = xcalloc(nx * nf, sizeof(double));  o->hSurf = xcalloc(nx * nf, sizeof(double));
    o->gGrad = xcalloc(nx * nc * 3, sizeof(double)); o->hGrad = xcalloc(nx * nc * 3, sizeof(double));
    o->gamG = xcalloc(nx * nbf, sizeof(double));  o->gamH = xcalloc(nx * nbf, sizeof(double));
    o->rho = xdup(rho, sizeof(double) * nc); o->U = xdup(U, sizeof(double) * 3 * nc);
    o->T = xdup(T, sizeof(double) * nc);
    o->q = xcalloc(3 * (size_t)nc, sizeof(double)); o->tau = xcalloc(nc, sizeof(double));
    o->rhoS = xcalloc(nf, sizeof(double)); o->US = xcalloc(3 * (size_t)nf, sizeof(double));
    o->TS = xcalloc(nf, sizeof(double)); o->qS = xcalloc(3 * (size_t)nf, sizeof(double));
    o->tauS = xcalloc(nf, sizeof(double));
    o->rho_b = xdup(rho_b, sizeof(double) * nbf); o->U_b = xdup(U_b, sizeof(double) * 3 * nbf);
    o->T_b = xdup(T_b, sizeof(double) * nbf);
    o->inByRho = xcalloc(nbf, sizeof(double)); o->outGoing = xcalloc(nbf, sizeof(double));
    o->qWall = xcalloc(3 * (size_t)nbf, sizeof(double));
    o->stressWall = xcalloc(9 * (size_t)nbf, sizeof(double));
    o->part = xcalloc((size_t)o->P * 16, sizeof(double));

    /* discreteVelocity ctor, discreteVelocity.C:206-210 */
    const double q0[3] = {0, 0, 0};
#pragma omp parallel for schedule(static)
    for (int k = 0; k < o->nxi; k++) {
        const double *xi = o->xi + 3 * k;
        /* initDFtoEq :220-249 — Shakhov with q = 0 */
        for (int c = 0; c < nc; c++)
            shakhov(o, xi, o->rho[c], o->U + 3 * c, o->T[c], q0, &o->gTilde[(size_t)k * nc + c],
                    &o->hTilde[(size_t)k * nc + c]);
        /* initBoundaryField :312-344 — "mixed" patches get a Maxwellian of the boundary macros */
        for (int p = 0; p < npatch; p++) {
            if (patches[p].kind != DUGKS_PATCH_MIXED) continue;
            for (int j = 0; j < patches[p].size; j++) {
                int b = patches[p].start + j;
                double Tb = o->T_b[b];
                const double *Ub = o->U_b + 3 * b;
                double d0 = Ub[0] - xi[0], d1 = Ub[1] - xi[1], d2 = Ub[2] - xi[2];
                double geq = o->rho_b[b] / pow(sqrt(2.0 * PI * o->R * Tb), o->D) *
                             exp(-(d0 * d0 + d1 * d1 + d2 * d2) / (2.0 * o->R * Tb)); /* :1058 */
                o->gSurf[(size_t)k * nf + nif + b] = geq;
                o->hSurf[(size_t)k * nf + nif + b] = (o->K + 3 - o->D) * o->R * Tb * geq; /* :1059 */
            }
        }
    }
    /* fvDVM ctor body, fvDVM.C:1067-1074 */
    set_wall_incoming(o);
    update_pressure_bc(o);
    for (int c = 0; c < nc; c++) o->tau[c] = tau_of(o, o->T[c], o->rho[c]);
    /* wall rho starts at 1, BCs/calculatedMaxwellFvPatchField/calculatedMaxwellFvPatchField.C:80 */
    for (int p = 0; p < npatch; p++)
        if (patches[p].kind == DUGKS_PATCH_MAXWELL_WALL)
            for (int j = 0; j < patches[p].size; j++) o->rho_b[patches[p].start + j] = 1.0;
    /* Usurf = fvc::interpolate(Uvol, "linear") :1074 [OF-lib]: w U_P + (1-w) U_N, boundary = patch value */
    for (int f = 0; f < nif; f++) {
        int own = o->owner[f], nei = o->neigh[f];
        const double *Sf = o->Sf + 3 * f, *Cf = o->Cf + 3 * f;
        double dn[3] = {o->C[3 * nei] - Cf[0], o->C[3 * nei + 1] - Cf[1], o->C[3 * nei + 2] - Cf[2]};
        double dp[3] = {Cf[0] - o->C[3 * own], Cf[1] - o->C[3 * own + 1], Cf[2] - o->C[3 * own + 2]};
        double sn = fabs(dot3(Sf, dn)), sp = fabs(dot3(Sf, dp));
        double wgt = sn / (sp + sn);
        for (int d = 0; d < 3; d++) o->US[3 * f + d] = wgt * o->U[3 * own + d] + (1 - wgt) * o->U[3 * nei + d];
    }
    for (int b = 0; b < nbf; b++)
        for (int d = 0; d < 3; d++) o->US[3 * (nif + b) + d] = o->U_b[3 * b + d];
    /* Told = T; rhoOld = rho; Uold = U, createFields.H:80-82 */
    o->Told = (double *)malloc(sizeof(double) * nc);
    o->rhoOld = (double *)malloc(sizeof(double) * nc);
    o->Uold = (double *)malloc(sizeof(double) * 3 * nc);
    memcpy(o->Told, o->T, sizeof(double) * nc);
    memcpy(o->rhoOld, o->rho, sizeof(double) * nc);
    memcpy(o->Uold, o->U, sizeof(double) * 3 * nc);
    return o;
}

void oracle_destroy(oracle_t *o) {
    if (!o) return;
    void *ptrs[] = {o->owner, o->neigh, o->C, o->V, o->Cf, o->Sf, o->ownLs, o->neiLs, o->patchLs,
                    o->dcoef, o->patch, o->xi, o->w, o->symX, o->symY, o->symZ, o->rank_of,
                    o->gTilde, o->hTilde, o->gBarP, o->hBarP, o->gSurf, o->hSurf, o->gGrad, o->hGrad,
                    o->gamG, o->gamH, o->rho, o->U, o->T, o->q, o->tau, o->rhoS, o->US, o->TS, o->qS,
                    o->tauS, o->rho_b, o->U_b, o->T_b, o->inByRho, o->outGoing, o->qWall,
                    o->stressWall, o->part, o->Told, o->rhoOld, o->Uold};
    for (size_t i = 0; i < sizeof(ptrs) / sizeof(ptrs[0]); i++) free(ptrs[i]);
    free(o);
}

/* ------------------------------------------------------------------------- */
/* stage 1: discreteVelocity::updateGHbarPvol, discreteVelocity.C:346-410       */
static void stage_barPvol(oracle_t *o, double dt) {
    int nc = o->nc;
#pragma omp parallel for schedule(static)
    for (int k = 0; k < o->nxi; k++) {
        const double *xi = o->xi + 3 * k;
        double *gB = o->gBarP + (size_t)k * nc, *hB = o->hBarP + (size_t)k * nc;
        const double *gT = o->gTilde + (size_t)k * nc, *hT = o->hTilde + (size_t)k * nc;
        for (int c = 0; c < nc; c++) {
            double rf = 1.5 * dt / (2.0 * o->tau[c] + dt); /* :393 */
            double gEq, hEq;
            shakhov(o, xi, o->rho[c], o->U + 3 * c, o->T[c], o->q + 3 * c, &gEq, &hEq); /* :396-404 */
            gB[c] = (1.0 - rf) * gT[c] + rf * gEq; /* :405 */
            hB[c] = (1.0 - rf) * hT[c] + rf * hEq; /* :406 */
        }
        /* :408-409 correctBoundaryConditions(): fixedGradient patch value is formed on
         * the fly in stage_barSurf (phi_b = phi_c + gradient/deltaCoeffs [OF-lib]). */
    }
}

/* boundary value of gBarPvol on boundary face b (fixedGradient / symmetryPlane evaluate [OF-lib]) */
static inline double bvalue(const oracle_t *o, int kind, double cellv, double gamma, double dcoef) {
    if (kind == DUGKS_PATCH_SYMMETRY_PLANE) return cellv; /* constraint patch: scalar mirror = cell value */
    return cellv + gamma / dcoef;
}

/* VenkatakrishnanSlopeMultiLimiter::limitFace, VenkatakrishnanSlopeMulti.C:83-128 (Blazek chap. 5) */
static inline double venkat_limit_face(double k, double V, double dMax, double dMin, double d2) {
    double sqrEps = k * k * k * V;
    double two = 2.0 * d2 * d2;
    if (d2 > 0.0) {
        double den = dMax * dMax + two + dMax * d2 + sqrEps;
        if (fabs(den) < VSMALL) den = den < 0 ? -VSMALL : VSMALL; /* stabilise() [OF-lib] */
        return ((dMax * dMax + sqrEps) + 2.0 * d2 * dMax) / den;
    } else if (d2 < 0.0) {
        double den = dMin * dMin + two + dMin * d2 + sqrEps;
        if (fabs(den) < VSMALL) den = den < 0 ? -VSMALL : VSMALL;
        return ((dMin * dMin + sqrEps) + 2.0 * d2 * dMin) / den;
    }
    return 1.0;
}

/* VenkatakrishnanLimitedGrad<scalar>::calcGrad, VenkatakrishnanLimitedGrads.C:59-226, AS IT IS MEANT TO WORK:
 * the reference limits a COPY of the gradient and returns the unlimited one (:76 `volVectorField g = tGrad()`,
 * :225 `return tGrad`), and starts the per-cell limiter from 0 instead of 1 (:140-151; the commented line :152
 * shows the intent), so in the reference the scheme has no effect at all (= grad_limiter 0).  Here: limiter = 1,
 * min over the faces of the cell of limitFace(V, max - phi, min - phi, (Cf - C).grad), grad *= limiter.
 * vb: boundary value of the field on every boundary face. */
static void venkat_limit(const oracle_t *o, const double *v, const double *vb, double *grad,
                         double *maxV, double *minV, double *lim) {
    int nc = o->nc, nif = o->nif, nbf = o->nbf;
    for (int c = 0; c < nc; c++) { maxV[c] = v[c]; minV[c] = v[c]; lim[c] = 1.0; }
    for (int f = 0; f < nif; f++) { /* :88-101 */
        int own = o->owner[f], nei = o->neigh[f];
        if (v[nei] > maxV[own]) maxV[own] = v[nei];
        if (v[nei] < minV[own]) minV[own] = v[nei];
        if (v[own] > maxV[nei]) maxV[nei] = v[own];
        if (v[own] < minV[nei]) minV[nei] = v[own];
    }
    for (int b = 0; b < nbf; b++) { /* :126-134 non-coupled patches: the patch value */
        int own = o->owner[nif + b];
        if (vb[b] > maxV[own]) maxV[own] = vb[b];
        if (vb[b] < minV[own]) minV[own] = vb[b];
    }
    for (int c = 0; c < nc; c++) { maxV[c] -= v[c]; minV[c] -= v[c]; } /* :137-138 */
    for (int f = 0; f < nif + nbf; f++) { /* :156-207 */
        const double *Cf = o->Cf + 3 * (size_t)f;
        int own = o->owner[f];
        double r[3] = {Cf[0] - o->C[3 * own], Cf[1] - o->C[3 * own + 1], Cf[2] - o->C[3 * own + 2]};
        double l = venkat_limit_face(o->limiter_k, o->V[own], maxV[own], minV[own], dot3(r, grad + 3 * own));
        if (l < lim[own]) lim[own] = l;
        if (f < nif) {
            int nei = o->neigh[f];
            double rn[3] = {Cf[0] - o->C[3 * nei], Cf[1] - o->C[3 * nei + 1], Cf[2] - o->C[3 * nei + 2]};
            l = venkat_limit_face(o->limiter_k, o->V[nei], maxV[nei], minV[nei], dot3(rn, grad + 3 * nei));
            if (l < lim[nei]) lim[nei] = l;
        }
    }
    for (int c = 0; c < nc; c++)
        for (int d = 0; d < 3; d++) grad[3 * c + d] *= lim[c]; /* :219 */
}

/* stage 2.1: discreteVelocity::updateGHbarSurf, discreteVelocity.C:412-691 */
static void stage_barSurf(oracle_t *o, double dt) {
    int nc = o->nc, nif = o->nif, nf = o->nf, nbf = o->nbf;
    memset(o->outGoing, 0, sizeof(double) * nbf);
    double *outPart = xcalloc((size_t)o->nxi * (nbf ? nbf : 1), sizeof(double));
#pragma omp parallel for schedule(static)
    for (int k = 0; k < o->nxi; k++) {
        const double *xi = o->xi + 3 * k;
        const double wk = o->w[k];
        const double *gB = o->gBarP + (size_t)k * nc, *hB = o->hBarP + (size_t)k * nc;
        double *gG = o->gGrad + (size_t)k * nc * 3, *hG = o->hGrad + (size_t)k * nc * 3;
        double *gS = o->gSurf + (size_t)k * nf, *hS = o->hSurf + (size_t)k * nf;
        double *gamG = o->gamG + (size_t)k * nbf, *gamH = o->gamH + (size_t)k * nbf;
        /* :420-421 fvc::grad -> stock leastSquaresGrad [OF-lib]; in-tree twin
         * zeroBoundaryGrad/zeroBoundaryGrad.C:90-99 plus the stock boundary lines kept in
         * comments at :126-133 */
        memset(gG, 0, sizeof(double) * 3 * nc);
        memset(hG, 0, sizeof(double) * 3 * nc);
        for (int f = 0; f < nif; f++) {
            int own = o->owner[f], nei = o->neigh[f];
            double dg = gB[nei] - gB[own], dh = hB[nei] - hB[own];
            for (int d = 0; d < 3; d++) {
                gG[3 * own + d] += o->ownLs[3 * f + d] * dg;
                gG[3 * nei + d] -= o->neiLs[3 * f + d] * dg;
                hG[3 * own + d] += o->ownLs[3 * f + d] * dh;
                hG[3 * nei + d] -= o->neiLs[3 * f + d] * dh;
            }
        }
        for (int p = 0; p < o->npatch; p++) {
            int kind = o->patch[p].kind;
            for (int j = 0; j < o->patch[p].size; j++) {
                int b = o->patch[p].start + j, own = o->owner[nif + b];
                double dc = o->dcoef[nif + b];
                double dg = bvalue(o, kind, gB[own], gamG[b], dc) - gB[own];
                double dh = bvalue(o, kind, hB[own], gamH[b], dc) - hB[own];
                for (int d = 0; d < 3; d++) {
                    gG[3 * own + d] += o->patchLs[3 * b + d] * dg;
                    hG[3 * own + d] += o->patchLs[3 * b + d] * dh;
                }
            }
        }
        if (o->grad_limiter == 1 && o->limiter_k >= 1e-15) { /* gradSchemes: VenkatakrishnanLimited leastSquares k (:70 k < SMALL: off) */
            double *vbg = xcalloc((size_t)(nbf ? nbf : 1) * 2 + 3 * (size_t)nc, sizeof(double));
            double *vbh = vbg + (nbf ? nbf : 1), *wrk = vbh + (nbf ? nbf : 1);
            for (int p = 0; p < o->npatch; p++)
                for (int j = 0; j < o->patch[p].size; j++) {
                    int b = o->patch[p].start + j, own = o->owner[nif + b];
                    vbg[b] = bvalue(o, o->patch[p].kind, gB[own], gamG[b], o->dcoef[nif + b]);
                    vbh[b] = bvalue(o, o->patch[p].kind, hB[own], gamH[b], o->dcoef[nif + b]);
                }
            venkat_limit(o, gB, vbg, gG, wrk, wrk + nc, wrk + 2 * nc);
            venkat_limit(o, hB, vbh, hG, wrk, wrk + nc, wrk + 2 * nc);
            free(vbg);
        }
        /* :427-470 boundary grad = cell grad (zeroGradient); its normal component becomes
         * the fixedGradient gradient() used NEXT step; skipped for constraint patches */
        for (int p = 0; p < o->npatch; p++) {
            if (o->patch[p].kind == DUGKS_PATCH_SYMMETRY_PLANE) continue; /* :447 */
            for (int j = 0; j < o->patch[p].size; j++) {
                int b = o->patch[p].start + j, own = o->owner[nif + b];
                const double *Sf = o->Sf + 3 * (size_t)(nif + b);
                double magSf = sqrt(dot3(Sf, Sf));
                double n[3] = {Sf[0] / magSf, Sf[1] / magSf, Sf[2] / magSf}; /* :438-442 */
                gamG[b] = dot3(gG + 3 * own, n); /* :464-465 */
                gamH[b] = dot3(hG + 3 * own, n); /* :466-467 */
            }
        }
        /* :491-530 internal faces */
        for (int f = 0; f < nif; f++) {
            int own = o->owner[f], nei = o->neigh[f];
            const double *Sf = o->Sf + 3 * f, *Cf = o->Cf + 3 * f;
            double phi = dot3(xi, Sf);
            double ro[3], rn[3];
            for (int d = 0; d < 3; d++) {
                ro[d] = Cf[d] - o->C[3 * own + d] - 0.5 * xi[d] * dt;
                rn[d] = Cf[d] - o->C[3 * nei + d] - 0.5 * xi[d] * dt;
            }
            if (phi >= VSMALL) { /* :495 */
                gS[f] = gB[own] + dot3(gG + 3 * own, ro);
                hS[f] = hB[own] + dot3(hG + 3 * own, ro);
            } else if (phi < -VSMALL) { /* :506 */
                gS[f] = gB[nei] + dot3(gG + 3 * nei, rn);
                hS[f] = hB[nei] + dot3(hG + 3 * nei, rn);
            } else { /* :513-529 */
                gS[f] = 0.5 * (gB[nei] + dot3(gG + 3 * nei, rn) + gB[own] + dot3(gG + 3 * own, ro));
                hS[f] = 0.5 * (hB[nei] + dot3(hG + 3 * nei, rn) + hB[own] + dot3(hG + 3 * own, ro));
            }
        }
        /* :533-690 boundary faces */
        for (int p = 0; p < o->npatch; p++) {
            int kind = o->patch[p].kind;
            for (int j = 0; j < o->patch[p].size; j++) {
                int b = o->patch[p].start + j, f = nif + b, own = o->owner[f];
                const double *Sf = o->Sf + 3 * (size_t)f, *Cf = o->Cf + 3 * (size_t)f;
                double phi = dot3(xi, Sf);
                double r[3];
                for (int d = 0; d < 3; d++) r[d] = Cf[d] - o->C[3 * own + d] - 0.5 * xi[d] * dt;
                double gOut = gB[own] + dot3(gG + 3 * own, r);
                double hOut = hB[own] + dot3(hG + 3 * own, r);
                switch (kind) {
                case DUGKS_PATCH_ZERO_GRADIENT: /* :551-555 patchInternalField */
                    gS[f] = gB[own]; hS[f] = hB[own];
                    break;
                case DUGKS_PATCH_MIXED: /* :556-573 */
                    if (phi > 0) { gS[f] = gOut; hS[f] = hOut; }
                    break;
                case DUGKS_PATCH_FAR_FIELD:
                case DUGKS_PATCH_PRESSURE_IN:
                case DUGKS_PATCH_PRESSURE_OUT: /* :574-604 */
                    if (phi > 0) { gS[f] = gOut; hS[f] = hOut; }
                    else {
                        gS[f] = o->rho_b[b] * maxwellByRho(o, xi, o->U + 3 * own, o->T_b[b]); /* :593-598 */
                        hS[f] = gS[f] * (o->R * o->T_b[b]) * (o->K + 3 - o->D);                /* :599-601 */
                    }
                    break;
                case DUGKS_PATCH_MAXWELL_WALL: /* :605-627 */
                    if (phi > 0) {
                        gS[f] = gOut; hS[f] = hOut;
                        outPart[(size_t)k * nbf + b] = wk * phi * gS[f]; /* :623-624 */
                    }
                    break;
                case DUGKS_PATCH_DVM_SYMMETRY:
                case DUGKS_PATCH_SYMMETRY_PLANE: /* :673-689 */
                    if (phi > -VSMALL) { gS[f] = gOut; hS[f] = hOut; }
                    break;
                default: break;
                }
            }
        }
    }
    /* the += at :623 runs over the local DVs in order; :363 all-reduces over ranks */
    for (int b = 0; b < nbf; b++) {
        double tot = 0.0;
        for (int r = 0; r < o->P; r++) {
            double acc = 0.0;
            for (int k = 0; k < o->nxi; k++)
                if (o->rank_of[k] == r) acc += outPart[(size_t)k * nbf + b];
            tot += acc;
        }
        o->outGoing[b] = tot;
    }
    free(outPart);
}

/* stage 2.2: fvDVM::updateMaxwellWallRho fvDVM.C:347-367 +
 * calculatedMaxwellFvPatchField::evaluate BCs/calculatedMaxwellFvPatchField/...C:139-164 */
static void stage_wallRho(oracle_t *o) {
    for (int p = 0; p < o->npatch; p++) {
        if (o->patch[p].kind != DUGKS_PATCH_MAXWELL_WALL) continue;
        for (int j = 0; j < o->patch[p].size; j++) {
            int b = o->patch[p].start + j;
            o->rho_b[b] = o->outGoing[b] / fabs(o->inByRho[b]); /* :158 */
            o->outGoing[b] = 0.0;                                 /* :163 */
        }
    }
}

/* stage 2.3: discreteVelocity::updateGHbarSurfMaxwellWallIn discreteVelocity.C:693-731 */
static void stage_wallIn(oracle_t *o) {
    int nif = o->nif, nf = o->nf;
#pragma omp parallel for schedule(static)
    for (int k = 0; k < o->nxi; k++) {
        const double *xi = o->xi + 3 * k;
        double *gS = o->gSurf + (size_t)k * nf, *hS = o->hSurf + (size_t)k * nf;
        for (int p = 0; p < o->npatch; p++) {
            if (o->patch[p].kind != DUGKS_PATCH_MAXWELL_WALL) continue;
            for (int j = 0; j < o->patch[p].size; j++) {
                int b = o->patch[p].start + j, f = nif + b;
                if (dot3(xi, o->Sf + 3 * (size_t)f) <= 0) { /* :713 */
                    gS[f] = o->rho_b[b] * maxwellByRho(o, xi, o->U_b + 3 * b, o->T_b[b]); /* :716-721 */
                    hS[f] = gS[f] * (o->R * o->T_b[b]) * (o->K + 3 - o->D);               /* :724-726 */
                }
            }
        }
    }
}

/* stage 2.4: fvDVM::updateGHbarSurfSymmetryIn fvDVM.C:375-454 +
 * discreteVelocity::updateGHbarSurfSymmetryIn discreteVelocity.C:733-817.
 * The dfContainer snapshot (fvDVM.C:423-449) is reproduced by copying from a
 * snapshot of the patch values of all DVs. */
static void stage_symmetryIn(oracle_t *o) {
    int nif = o->nif, nf = o->nf;
    for (int p = 0; p < o->npatch; p++) {
        int kind = o->patch[p].kind;
        if (kind != DUGKS_PATCH_DVM_SYMMETRY && kind != DUGKS_PATCH_SYMMETRY_PLANE) continue;
        int ps = o->patch[p].size, b0 = o->patch[p].start;
        if (ps <= 0) continue; /* :747 */
        double *snapG = xcalloc((size_t)o->nxi * ps, sizeof(double));
        double *snapH = xcalloc((size_t)o->nxi * ps, sizeof(double));
        for (int k = 0; k < o->nxi; k++) {
            memcpy(snapG + (size_t)k * ps, o->gSurf + (size_t)k * nf + nif + b0, sizeof(double) * ps);
            memcpy(snapH + (size_t)k * ps, o->hSurf + (size_t)k * nf + nif + b0, sizeof(double) * ps);
        }
        const double *Sf0 = o->Sf + 3 * (size_t)(nif + b0); /* :763 first face of the patch */
        double mag = sqrt(dot3(Sf0, Sf0));
        double n[3] = {Sf0[0] / mag, Sf0[1] / mag, Sf0[2] / mag};
        for (int k = 0; k < o->nxi; k++) {
            const double *xi = o->xi + 3 * k;
            if (dot3(xi, Sf0) <= 0) { /* :775 */
                double t = n[0] * o->symX[k] + n[1] * o->symY[k] + n[2] * o->symZ[k];
                int tgt = (int)lround(fabs(t)); /* :778-779 */
                memcpy(o->gSurf + (size_t)k * nf + nif + b0, snapG + (size_t)tgt * ps, sizeof(double) * ps);
                memcpy(o->hSurf + (size_t)k * nf + nif + b0, snapH + (size_t)tgt * ps, sizeof(double) * ps);
            }
        }
        free(snapG); free(snapH);
    }
}

/* sum over the virtual ranks of the rank-local sums (local DV order), i.e. the
 * forAll(DV_) accumulation followed by MPI_Allreduce */
#define RANK_SUM(result, expr)                                   \
    do {                                                         \
        double tot__ = 0.0;                                      \
        for (int r__ = 0; r__ < o->P; r__++) {                   \
            double acc__ = 0.0;                                  \
            for (int k = 0; k < o->nxi; k++) {                   \
                if (o->rank_of[k] != r__) continue;              \
                acc__ += (expr);                                 \
            }                                                    \
            tot__ += acc__;                                      \
        }                                                        \
        (result) = tot__;                                        \
    } while (0)

/* The DV sums of fvDVM::updateMacroSurf / updateMacroVol (fvDVM.C:473-483,503-516 and :612-622,712-721), for the
 * n <= MBLK consecutive entries [e0, e0 + n) of a DV-major field pair g, h (leading dimension ld).  The reference
 * streams one whole field per discrete velocity (DV outermost); this is that loop restricted to a block of entries
 * that stays in cache, so every memory access is contiguous.  Per entry the terms are added in the order of the
 * reference (k ascending inside a rank, ranks in order: forAll(DV_) + MPI_Allreduce): same bits as the
 * entry-by-entry RANK_SUM form. */
#define MBLK 512
static void dv_moments(const oracle_t *o, const double *g, const double *h, size_t ld, int e0, int n,
                       double *rho, double *rU, double *rE) {
    double a0[MBLK], a1[MBLK], a2[MBLK], a3[MBLK], aE[MBLK];
    for (int j = 0; j < n; j++) { rho[j] = 0.0; rU[3 * j] = rU[3 * j + 1] = rU[3 * j + 2] = 0.0; rE[j] = 0.0; }
    for (int r = 0; r < o->P; r++) {
        for (int j = 0; j < n; j++) a0[j] = a1[j] = a2[j] = a3[j] = aE[j] = 0.0;
        for (int k = 0; k < o->nxi; k++) {
            if (o->rank_of[k] != r) continue;
            const double w = o->w[k], x = o->xi[3 * k], y = o->xi[3 * k + 1], z = o->xi[3 * k + 2];
            const double x2 = dot3(o->xi + 3 * k, o->xi + 3 * k);
            const double *gk = g + (size_t)k * ld + e0, *hk = h + (size_t)k * ld + e0;
            for (int j = 0; j < n; j++) {
                a0[j] += w * gk[j];
                a1[j] += w * gk[j] * x;
                a2[j] += w * gk[j] * y;
                a3[j] += w * gk[j] * z;
                aE[j] += 0.5 * w * (gk[j] * x2 + hk[j]);
            }
        }
        for (int j = 0; j < n; j++) {
            rho[j] += a0[j]; rU[3 * j] += a1[j]; rU[3 * j + 1] += a2[j]; rU[3 * j + 2] += a3[j]; rE[j] += aE[j];
        }
    }
}

/* second pass: 1/2 sum w c (|c|^2 g + h), c = xi - U[entry] */
static void dv_heat_flux(const oracle_t *o, const double *g, const double *h, size_t ld, int e0, int n,
                         const double *U /* [n][3] */, double *q /* [n][3] */) {
    double a[3][MBLK];
    for (int j = 0; j < 3 * n; j++) q[j] = 0.0;
    for (int r = 0; r < o->P; r++) {
        for (int d = 0; d < 3; d++) for (int j = 0; j < n; j++) a[d][j] = 0.0;
        for (int k = 0; k < o->nxi; k++) {
            if (o->rank_of[k] != r) continue;
            const double w = o->w[k], x = o->xi[3 * k], y = o->xi[3 * k + 1], z = o->xi[3 * k + 2];
            const double *gk = g + (size_t)k * ld + e0, *hk = h + (size_t)k * ld + e0;
            for (int j = 0; j < n; j++) {
                const double cx = x - U[3 * j], cy = y - U[3 * j + 1], cz = z - U[3 * j + 2];
                const double m = (cx * cx + cy * cy + cz * cz) * gk[j] + hk[j];
                a[0][j] += 0.5 * w * cx * m;
                a[1][j] += 0.5 * w * cy * m;
                a[2][j] += 0.5 * w * cz * m;
            }
        }
        for (int d = 0; d < 3; d++) for (int j = 0; j < n; j++) q[3 * j + d] += a[d][j];
    }
}

/* stage 3: fvDVM::updateMacroSurf fvDVM.C:456-582 */
static void stage_macroSurf(oracle_t *o, double dt) {
    int nf = o->nf, nif = o->nif;
    const double R = o->R;
    const int K = o->K;
#pragma omp parallel for schedule(dynamic)
    for (int f0 = 0; f0 < nf; f0 += MBLK) {
        const int n = nf - f0 < MBLK ? nf - f0 : MBLK;
        double rho[MBLK], rU[3 * MBLK], rE[MBLK], qs[3 * MBLK];
        dv_moments(o, o->gSurf, o->hSurf, (size_t)nf, f0, n, rho, rU, rE); /* :473-490 */
        for (int j = 0; j < n; j++) {
            const int f = f0 + j;
            double *Us = o->US + 3 * f;
            for (int d = 0; d < 3; d++) Us[d] = rU[3 * j + d] / rho[j]; /* :493 */
            o->rhoS[f] = rho[j];
            o->TS[f] = (rE[j] - 0.5 * rho[j] * dot3(Us, Us)) / ((K + 3) / 2.0 * R * rho[j]); /* :495 */
            o->tauS[f] = tau_of(o, o->TS[f], rho[j]); /* :497 */
        }
        dv_heat_flux(o, o->gSurf, o->hSurf, (size_t)nf, f0, n, o->US + 3 * (size_t)f0, qs); /* :503-519 */
        for (int j = 0; j < n; j++) {
            const int f = f0 + j;
            double fac = 2.0 * o->tauS[f] / (2.0 * o->tauS[f] + 0.5 * dt * o->Pr); /* :522 */
            for (int d = 0; d < 3; d++) o->qS[3 * f + d] = fac * qs[3 * j + d];
        }
    }
    /* :539-581 wall diagnostics: per wall face 1/2 sum w c (|c|^2 g + h) with c = xi - U_wall, and sum w g xi xi;
     * blocks of boundary faces, DV outermost inside a block (contiguous reads), terms added per face in the
     * reference's order (k ascending inside a rank, ranks in order) */
    memset(o->qWall, 0, sizeof(double) * 3 * o->nbf);
    memset(o->stressWall, 0, sizeof(double) * 9 * o->nbf);
    for (int p = 0; p < o->npatch; p++) {
        if (o->patch[p].kind != DUGKS_PATCH_MAXWELL_WALL) continue;
        const int pb0 = o->patch[p].start, pn = o->patch[p].size;
#pragma omp parallel for schedule(dynamic)
        for (int j0 = 0; j0 < pn; j0 += MBLK) {
            const int n = pn - j0 < MBLK ? pn - j0 : MBLK;
            const int b0 = pb0 + j0;
            double acc[12][MBLK], tot[12][MBLK];
            for (int m = 0; m < 12; m++) for (int j = 0; j < n; j++) tot[m][j] = 0.0;
            for (int r = 0; r < o->P; r++) {
                for (int m = 0; m < 12; m++) for (int j = 0; j < n; j++) acc[m][j] = 0.0;
                for (int k = 0; k < o->nxi; k++) {
                    if (o->rank_of[k] != r) continue;
                    const double w = o->w[k], *xi = o->xi + 3 * k;
                    const double *gk = o->gSurf + (size_t)k * nf + nif + b0, *hk = o->hSurf + (size_t)k * nf + nif + b0;
                    for (int j = 0; j < n; j++) {
                        const double *Up = o->U_b + 3 * (b0 + j);
                        const double cx = xi[0] - Up[0], cy = xi[1] - Up[1], cz = xi[2] - Up[2];
                        const double m = (cx * cx + cy * cy + cz * cz) * gk[j] + hk[j]; /* :563-567 */
                        acc[0][j] += 0.5 * w * cx * m;
                        acc[1][j] += 0.5 * w * cy * m;
                        acc[2][j] += 0.5 * w * cz * m;
                        for (int a = 0; a < 3; a++)
                            for (int c = 0; c < 3; c++) acc[3 + 3 * a + c][j] += w * gk[j] * xi[a] * xi[c]; /* :568-569 */
                    }
                }
                for (int m = 0; m < 12; m++) for (int j = 0; j < n; j++) tot[m][j] += acc[m][j];
            }
            for (int j = 0; j < n; j++) {
                const int b = b0 + j, f = nif + b;
                double taup = o->tauS[f];
                double fq = 2.0 * taup / (2.0 * taup + 0.5 * dt * o->Pr); /* :571 */
                double fs = 2.0 * taup / (2.0 * taup + 0.5 * dt);         /* :573 */
                for (int d = 0; d < 3; d++) o->qWall[3 * b + d] = fq * tot[d][j];
                for (int m = 0; m < 9; m++) o->stressWall[9 * b + m] = fs * tot[3 + m][j];
            }
        }
    }
}

/* stage 4: discreteVelocity::updateGHsurf discreteVelocity.C:819-932 */
static void stage_surf(oracle_t *o, double dt) {
    int nf = o->nf, nif = o->nif;
    double h = 0.5 * dt; /* :822 */
#pragma omp parallel for schedule(static)
    for (int k = 0; k < o->nxi; k++) {
        const double *xi = o->xi + 3 * k;
        double *gS = o->gSurf + (size_t)k * nf, *hS = o->hSurf + (size_t)k * nf;
        for (int f = 0; f < nif; f++) { /* :867-881 internal field */
            double rf = h / (2 * o->tauS[f] + h);
            double gEq, hEq;
            shakhov(o, xi, o->rhoS[f], o->US + 3 * f, o->TS[f], o->qS + 3 * f, &gEq, &hEq);
            gS[f] = (1.0 - rf) * gS[f] + rf * gEq;
            hS[f] = (1.0 - rf) * hS[f] + rf * hEq;
        }
        for (int p = 0; p < o->npatch; p++) { /* :886-931 */
            int kind = o->patch[p].kind;
            for (int j = 0; j < o->patch[p].size; j++) {
                int f = nif + o->patch[p].start + j;
                double rf = h / (2 * o->tauS[f] + h);
                double gEq, hEq;
                shakhov(o, xi, o->rhoS[f], o->US + 3 * f, o->TS[f], o->qS + 3 * f, &gEq, &hEq);
                double phi = dot3(xi, o->Sf + 3 * (size_t)f);
                if (kind == DUGKS_PATCH_SYMMETRY_PLANE) {
                    /* stock symmetryPlane fvsPatchField takes part in the whole-field
                     * expression at :880-881 [OF-lib] */
                    gS[f] = (1.0 - rf) * gS[f] + rf * gEq;
                    hS[f] = (1.0 - rf) * hS[f] + rf * hEq;
                }
                if (phi > 0) { /* :907-919 outgoing only */
                    gS[f] = (1.0 - rf) * gS[f] + rf * gEq;
                    hS[f] = (1.0 - rf) * hS[f] + rf * hEq;
                }
                if (kind == DUGKS_PATCH_DVM_SYMMETRY) { /* :922-930 whole patch again */
                    gS[f] = (1.0 - rf) * gS[f] + rf * gEq;
                    hS[f] = (1.0 - rf) * hS[f] + rf * hEq;
                }
            }
        }
    }
}

/* stage 5: discreteVelocity::updateGHtildeVol discreteVelocity.C:934-978 */
static void stage_tildeVol(oracle_t *o, double dt) {
    int nc = o->nc, nf = o->nf, nif = o->nif;
#pragma omp parallel for schedule(static)
    for (int k = 0; k < o->nxi; k++) {
        const double *xi = o->xi + 3 * k;
        double *gT = o->gTilde + (size_t)k * nc, *hT = o->hTilde + (size_t)k * nc;
        const double *gB = o->gBarP + (size_t)k * nc, *hB = o->hBarP + (size_t)k * nc;
        const double *gS = o->gSurf + (size_t)k * nf, *hS = o->hSurf + (size_t)k * nf;
        for (int c = 0; c < nc; c++) { /* :937-938 */
            gT[c] = -1.0 / 3 * gT[c] + 4.0 / 3 * gB[c];
            hT[c] = -1.0 / 3 * hT[c] + 4.0 / 3 * hB[c];
        }
        for (int f = 0; f < nif; f++) { /* :948-956 */
            int own = o->owner[f], nei = o->neigh[f];
            double phi = dot3(xi, o->Sf + 3 * (size_t)f);
            gT[own] -= (phi * gS[f] * dt / o->V[own]);
            gT[nei] += (phi * gS[f] * dt / o->V[nei]);
            hT[own] -= (phi * hS[f] * dt / o->V[own]);
            hT[nei] += (phi * hS[f] * dt / o->V[nei]);
        }
        for (int f = nif; f < nf; f++) { /* :959-976 */
            int own = o->owner[f];
            double phi = dot3(xi, o->Sf + 3 * (size_t)f);
            gT[own] -= phi * gS[f] * dt / o->V[own];
            hT[own] -= phi * hS[f] * dt / o->V[own];
        }
    }
}

/* stage 6: fvDVM::updateMacroVol fvDVM.C:597-728 (macroFlux == "no") */
static void stage_macroVol(oracle_t *o, double dt) {
    int nc = o->nc;
    const double R = o->R;
    const int K = o->K;
#pragma omp parallel for schedule(dynamic)
    for (int c0 = 0; c0 < nc; c0 += MBLK) {
        const int n = nc - c0 < MBLK ? nc - c0 : MBLK;
        double rho[MBLK], rU[3 * MBLK], rE[MBLK];
        dv_moments(o, o->gTilde, o->hTilde, (size_t)nc, c0, n, rho, rU, rE); /* :612-628 */
        for (int j = 0; j < n; j++) {
            const int c = c0 + j;
            o->rho[c] = rho[j];
            for (int d = 0; d < 3; d++) o->U[3 * c + d] = rU[3 * j + d] / rho[j]; /* :694 */
            o->T[c] = (rE[j] - 0.5 * rho[j] * dot3(o->U + 3 * c, o->U + 3 * c)) / ((K + 3) / 2.0 * R * rho[j]); /* :695 */
            o->tau[c] = tau_of(o, o->T[c], rho[j]); /* :707 */
        }
    }
    correct_macro_bcs(o); /* :698-699 */
#pragma omp parallel for schedule(dynamic)
    for (int c0 = 0; c0 < nc; c0 += MBLK) {
        const int n = nc - c0 < MBLK ? nc - c0 : MBLK;
        double qs[3 * MBLK];
        dv_heat_flux(o, o->gTilde, o->hTilde, (size_t)nc, c0, n, o->U + 3 * (size_t)c0, qs); /* :712-725 */
        for (int j = 0; j < n; j++) {
            const int c = c0 + j;
            double fac = 2.0 * o->tau[c] / (2.0 * o->tau[c] + dt * o->Pr); /* :727 */
            for (int d = 0; d < 3; d++) o->q[3 * c + d] = fac * qs[3 * j + d];
        }
    }
}

/* gradSchemes of system/fvSchemes (doc/usage.tex:169-192): 0 = leastSquares (what the demos use, and what the
 * reference's VenkatakrishnanLimited amounts to because of its copy bug), 1 = VenkatakrishnanLimited leastSquares k
 * as it is meant to work */
void oracle_set_grad_scheme(oracle_t *o, int limiter, double k) {
    o->grad_limiter = limiter;
    o->limiter_k = k;
}

/* fvDVM::evolution fvDVM.C:1086-1108 */
void oracle_step(oracle_t *o, double dt) {
    /* ORACLE_TIMING=1: wall time per stage on stderr (where the CPU baseline of bench.py spends its time) */
    const int timing = getenv("ORACLE_TIMING") != NULL;
    double t[11];
#define STAGE(i, call) do { call; if (timing) t[i] = omp_get_wtime(); } while (0)
    if (timing) t[0] = omp_get_wtime();
    STAGE(1, stage_barPvol(o, dt));     /* :1089 */
    STAGE(2, stage_barSurf(o, dt));     /* :1091 */
    STAGE(3, stage_wallRho(o));         /* :1093 */
    STAGE(4, stage_wallIn(o));          /* :1095 */
    STAGE(5, stage_symmetryIn(o));      /* :1097 */
    STAGE(6, stage_macroSurf(o, dt));   /* :1099 */
    STAGE(7, stage_surf(o, dt));        /* :1101 */
    STAGE(8, stage_tildeVol(o, dt));    /* :1103 */
    STAGE(9, stage_macroVol(o, dt));    /* :1105 */
    STAGE(10, update_pressure_bc(o));   /* :1107 */
#undef STAGE
    if (timing) {
        static const char *nm[] = {"", "barPvol", "barSurf", "wallRho", "wallIn", "symmetryIn", "macroSurf", "surf", "tildeVol", "macroVol", "pressureBC"};
        fprintf(stderr, "oracle_step:");
        for (int i = 1; i <= 10; i++) fprintf(stderr, " %s %.3f", nm[i], t[i] - t[i - 1]);
        fprintf(stderr, " s\n");
    }
    o->steps++;
}

/* fvDVM::getCoNum fvDVM.C:1111-1119 (internal faces: surfaceScalarField -> scalarField) */
void oracle_courant(const oracle_t *o, double dt, double *maxCo, double *meanCo) {
    double mx = -1e300, sum = 0.0;
    for (int f = 0; f < o->nif; f++) {
        const double *u = o->US + 3 * f;
        double v = o->dcoef[f] * (sqrt(dot3(u, u)) + sqrt((double)o->D) * o->xiMax);
        if (v > mx) mx = v;
        sum += v;
    }
    *maxCo = mx * dt;
    *meanCo = sum / o->nif * dt;
}

/* Convergence monitor of the time loop, dugksFoam.C:88-107: relative change of T, rho and U since the
 * previous check (gSum(mag(T - Told)) / gSum(T), ..., gSum(mag(U - Uold)) / gSum(mag(U))), then
 * Told = T; rhoOld = rho; Uold = U.  out = {TemperatureChange, rhoChange, Uchange}. */
void oracle_convergence(oracle_t *o, double *out) {
    double dT = 0, sT = 0, dR = 0, sR = 0, dU = 0, sU = 0;
    for (int c = 0; c < o->nc; c++) {
        dT += fabs(o->T[c] - o->Told[c]);
        sT += o->T[c];
        dR += fabs(o->rho[c] - o->rhoOld[c]);
        sR += o->rho[c];
        const double *u = o->U + 3 * c, *v = o->Uold + 3 * c;
        double d[3] = {u[0] - v[0], u[1] - v[1], u[2] - v[2]};
        dU += sqrt(dot3(d, d));
        sU += sqrt(dot3(u, u));
    }
    out[0] = dT / sT;
    out[1] = dR / sR;
    out[2] = dU / sU;
    memcpy(o->Told, o->T, sizeof(double) * o->nc);
    memcpy(o->rhoOld, o->rho, sizeof(double) * o->nc);
    memcpy(o->Uold, o->U, sizeof(double) * 3 * o->nc);
}

/* ---- accessors ---------------------------------------------------------- */
int oracle_nxi(const oracle_t *o) { return o->nxi; }
void oracle_get_cell_macros(const oracle_t *o, double *rho, double *U, double *T, double *q, double *tau) {
    if (rho) memcpy(rho, o->rho, sizeof(double) * o->nc);
    if (U) memcpy(U, o->U, sizeof(double) * 3 * o->nc);
    if (T) memcpy(T, o->T, sizeof(double) * o->nc);
    if (q) memcpy(q, o->q, sizeof(double) * 3 * o->nc);
    if (tau) memcpy(tau, o->tau, sizeof(double) * o->nc);
}
void oracle_get_face_macros(const oracle_t *o, double *rho, double *U, double *T, double *q, double *tau) {
    if (rho) memcpy(rho, o->rhoS, sizeof(double) * o->nf);
    if (U) memcpy(U, o->US, sizeof(double) * 3 * o->nf);
    if (T) memcpy(T, o->TS, sizeof(double) * o->nf);
    if (q) memcpy(q, o->qS, sizeof(double) * 3 * o->nf);
    if (tau) memcpy(tau, o->tauS, sizeof(double) * o->nf);
}
void oracle_get_boundary_macros(const oracle_t *o, double *rho_b, double *U_b, double *T_b) {
    if (rho_b) memcpy(rho_b, o->rho_b, sizeof(double) * o->nbf);
    if (U_b) memcpy(U_b, o->U_b, sizeof(double) * 3 * o->nbf);
    if (T_b) memcpy(T_b, o->T_b, sizeof(double) * o->nbf);
}
void oracle_get_wall_diag(const oracle_t *o, double *qWall, double *stressWall) {
    if (qWall) memcpy(qWall, o->qWall, sizeof(double) * 3 * o->nbf);
    if (stressWall) memcpy(stressWall, o->stressWall, sizeof(double) * 9 * o->nbf);
}
/* state, DV-major [nxi][nc] with GLOBAL DV ids */
void oracle_get_state(const oracle_t *o, double *g, double *h) {
    if (g) memcpy(g, o->gTilde, sizeof(double) * (size_t)o->nxi * o->nc);
    if (h) memcpy(h, o->hTilde, sizeof(double) * (size_t)o->nxi * o->nc);
}
void oracle_set_state(oracle_t *o, const double *g, const double *h) {
    if (g) memcpy(o->gTilde, g, sizeof(double) * (size_t)o->nxi * o->nc);
    if (h) memcpy(o->hTilde, h, sizeof(double) * (size_t)o->nxi * o->nc);
}
/* face values of one DV (gSurf/hSurf), [nf] */
void oracle_get_surf(const oracle_t *o, int k, double *g, double *h) {
    if (g) memcpy(g, o->gSurf + (size_t)k * o->nf, sizeof(double) * o->nf);
    if (h) memcpy(h, o->hSurf + (size_t)k * o->nf, sizeof(double) * o->nf);
}
void oracle_get_dvs(const oracle_t *o, double *xi, double *w, int32_t *symX, int32_t *symY, int32_t *symZ) {
    if (xi) memcpy(xi, o->xi, sizeof(double) * 3 * o->nxi);
    if (w) memcpy(w, o->w, sizeof(double) * o->nxi);
    if (symX) memcpy(symX, o->symX, sizeof(int32_t) * o->nxi);
    if (symY) memcpy(symY, o->symY, sizeof(int32_t) * o->nxi);
    if (symZ) memcpy(symZ, o->symZ, sizeof(int32_t) * o->nxi);
}
void oracle_get_wall_incoming(const oracle_t *o, double *inByRho) {
    memcpy(inByRho, o->inByRho, sizeof(double) * o->nbf);
}
