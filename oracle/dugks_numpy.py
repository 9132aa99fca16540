"""Second, independent CPU restatement of the reference's time step — TEST INFRASTRUCTURE ONLY.

Written field-at-a-time from the reference text (one discrete velocity after the other, whole-field NumPy
expressions, exactly as fvDVM / discreteVelocity operate on OpenFOAM fields), sharing NO code with
oracle/dugks_oracle.c: its purpose is to catch common-mode misreadings of the reference in the C oracle
(tests/test_oracle.py cross-checks the two at 1e-13).  PARITY UNPINNED, like the C oracle: the reference ships
no golden outputs for this path and cannot be compiled here (OpenFOAM + MPI).

Only tests/ may import this module.  Stage -> reference lines are given at each method.  Supported DF boundary
kinds: zeroGradient, mixed, maxwellWall, farField, DVMsymmetry, pressureIn/Out (symmetryPlane is a constraint
patch whose behaviour lives in OpenFOAM, not in the reference tree: the C oracle documents its reading).
"""
from __future__ import annotations

import numpy as np

VSMALL = 1.0e-300


def xi_dot_sf(xi, Sf):
    """xi & Sf evaluated like OpenFOAM's vector inner product, x*Sx + y*Sy + z*Sz, term by term without fused
    multiply-adds (a BLAS matrix-vector product may sum in another order): the SIGN of this number picks the upwind
    cell, and on a face whose normal is at 45 degrees a velocity with xi_x = xi_y gives exactly 0 one way and 1e-18 the
    other - a tie averaged from both sides (discreteVelocity.C:513-529) or a one-sided value."""
    return xi[0] * Sf[:, 0] + xi[1] * Sf[:, 1] + xi[2] * Sf[:, 2]


ZERO_GRADIENT, MIXED, MAXWELL_WALL, FAR_FIELD, DVM_SYMMETRY, SYMMETRY_PLANE, PRESSURE_IN, PRESSURE_OUT = range(8)


class NumpyDVM:
    def __init__(self, case):
        g = case.geom
        self.case = case
        self.nc, self.nif, self.nbf, self.D = g.nCells, g.nInternalFaces, g.nBoundaryFaces, g.nSolutionD
        self.own, self.nei = np.asarray(g.owner[: self.nif]), np.asarray(g.neighbour)
        self.bown = np.asarray(g.owner[self.nif:])
        self.C, self.V, self.Cf, self.Sf = g.C, g.V, g.Cf, g.Sf
        self.ownLs, self.neiLs, self.patchLs, self.dc = g.ownLs, g.neiLs, g.patchLs, g.deltaCoeffs
        gas = case.gas
        self.R, self.omega, self.Tref, self.muRef, self.Pr = (gas[k] for k in ("R", "omega", "Tref", "muRef", "Pr"))
        self.K = int(gas.get("KInner", 0))
        self.patches = case.patches
        # fvDVM::initialiseDV, fvDVM.C:140-220: id = iz n n + iy n + ix, weight = product of the 1-D weights
        X, W, n, D = np.asarray(case.Xis), np.asarray(case.weights), len(case.Xis), self.D
        ny, nz = (n if D >= 2 else 1), (n if D == 3 else 1)
        iz, iy, ix = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(n), indexing="ij")
        ix, iy, iz = ix.ravel(), iy.ravel(), iz.ravel()
        self.xi = np.stack([X[ix], X[iy] if D >= 2 else 0 * X[ix], X[iz] if D == 3 else 0 * X[ix]], axis=1)
        self.w = W[ix] * (W[iy] if D >= 2 else 1.0) * (W[iz] if D == 3 else 1.0)
        self.mirror = np.stack([(iz * ny + iy) * n + (n - 1 - ix), (iz * ny + (ny - 1 - iy)) * n + ix,
                                ((nz - 1 - iz) * ny + iy) * n + ix], axis=1)     # :162-164
        self.nxi = len(self.w)
        self.xiMax = case.xiMax
        # macro fields
        self.rho, self.U, self.T = case.rho.copy(), case.U.copy(), case.T.copy()
        self.q = np.zeros((self.nc, 3))
        self.rho_b, self.U_b, self.T_b = case.rho_b.copy(), case.U_b.copy(), case.T_b.copy()
        self.kind = np.zeros(self.nbf, dtype=int)
        for p in self.patches:
            self.kind[p.start:p.start + p.size] = p.kind
        self._pressure_bc()                                                    # fvDVM.C:1072
        self.tau = self._tau(self.T, self.rho)                                 # :1073
        # discreteVelocity ctor: initDFtoEq (q = 0), initBoundaryField ("mixed" patches)  discreteVelocity.C:220-249,312-344
        self.gT = np.empty((self.nxi, self.nc)); self.hT = np.empty((self.nxi, self.nc))
        nf = self.nif + self.nbf
        self.gS = np.zeros((self.nxi, nf)); self.hS = np.zeros((self.nxi, nf))
        self.gam_g = np.zeros((self.nxi, self.nbf)); self.gam_h = np.zeros((self.nxi, self.nbf))
        mixed = self.kind == MIXED
        for k in range(self.nxi):
            self.gT[k], self.hT[k] = self._shakhov(self.xi[k], self.rho, self.U, self.T, self.q)
            if mixed.any():
                geq = self.rho_b[mixed] * self._maxwell_by_rho(self.xi[k], self.U_b[mixed], self.T_b[mixed])
                self.gS[k, self.nif:][mixed] = geq
                self.hS[k, self.nif:][mixed] = (self.K + 3 - self.D) * self.R * self.T_b[mixed] * geq
        self._wall_incoming()                                                  # fvDVM.C:1069 -> :263-309
        # Usurf = fvc::interpolate(U, "linear") for the first Courant number (:1074) [OF-lib linear weights]
        sn = np.abs(np.einsum("ij,ij->i", self.Sf[: self.nif], self.C[self.nei] - self.Cf[: self.nif]))
        sp = np.abs(np.einsum("ij,ij->i", self.Sf[: self.nif], self.Cf[: self.nif] - self.C[self.own]))
        wl = (sn / (sp + sn))[:, None]
        self.Usurf = np.vstack([wl * self.U[self.own] + (1 - wl) * self.U[self.nei], self.U_b])
        self.rhoSurf = np.zeros(nf); self.Tsurf = np.zeros(nf); self.qSurf = np.zeros((nf, 3)); self.tauSurf = np.zeros(nf)
        self.qWall = np.zeros((self.nbf, 3)); self.stressWall = np.zeros((self.nbf, 3, 3))

    # ---- small pieces -------------------------------------------------------------------------------
    def _tau(self, T, rho):                                                    # fvDVM::updateTau, fvDVM.C:816
        return self.muRef * np.exp(self.omega * np.log(T / self.Tref)) / rho / T / self.R

    def _shakhov(self, xi, rho, U, T, q):                                      # discreteVelocity.C:1016-1044
        R, D, K, Pr = self.R, self.D, self.K, self.Pr
        cSqrByRT = ((U - xi) ** 2).sum(axis=1) / (R * T)
        cqBy5pRT = ((xi - U) * q).sum(axis=1) / (5.0 * rho * R * T * R * T)
        gEqBGK = rho / np.sqrt(2.0 * np.pi * R * T) ** D * np.exp(-cSqrByRT / 2.0)
        gEq = (1.0 + (1.0 - Pr) * cqBy5pRT * (cSqrByRT - D - 2.0)) * gEqBGK
        hEq = ((K + 3.0 - D) + (1.0 - Pr) * cqBy5pRT * ((cSqrByRT - D) * (K + 3.0 - D) - 2 * K)) * gEqBGK * R * T
        return gEq, hEq

    def _maxwell_by_rho(self, xi, U, T):                                       # discreteVelocity.C:1063-1075
        return 1.0 / np.sqrt(2.0 * np.pi * self.R * T) ** self.D * np.exp(-((U - xi) ** 2).sum(axis=-1) / (2.0 * self.R * T))

    def _wall_incoming(self):                                                  # fvDVM::setCalculatedMaxwellRhoBC, fvDVM.C:263-309
        self.inByRho = np.zeros(self.nbf)
        wall = np.where(self.kind == MAXWELL_WALL)[0]
        Sfb = self.Sf[self.nif:]
        for k in range(self.nxi):
            phi = xi_dot_sf(self.xi[k], Sfb[wall])
            inc = phi < 0
            b = wall[inc]
            self.inByRho[b] += -self.w[k] * phi[inc] * self._maxwell_by_rho(self.xi[k], self.U_b[b], self.T_b[b])

    def _pressure_bc(self):                                                    # fvDVM::updatePressureInOutBC, fvDVM.C:730-806
        R, K = self.R, self.K
        Sfb = self.Sf[self.nif:]
        for p in self.patches:
            if p.kind not in (PRESSURE_IN, PRESSURE_OUT):
                continue
            sl = slice(p.start, p.start + p.size)
            own = self.bown[sl]
            Ui, Ti, rhoi = self.U[own], self.T[own], self.rho[own]
            ai = np.sqrt(R * Ti * (K + 5) / (K + 3))
            norm = Sfb[sl] / np.linalg.norm(Sfb[sl], axis=1)[:, None]
            Un = (Ui * norm).sum(axis=1)
            if p.kind == PRESSURE_IN:
                self.rho_b[sl] = p.pressure / R / self.T_b[sl]
                UnIn = Un + (p.pressure - rhoi * R * Ti) / rhoi / ai
            else:
                self.rho_b[sl] = rhoi + (p.pressure - rhoi * R * Ti) / ai / ai
                self.T_b[sl] = p.pressure / (R * rhoi)
                UnIn = Un + (rhoi * R * Ti - p.pressure) / rhoi / ai
            self.U_b[sl] = UnIn[:, None] * norm + (Ui - Un[:, None] * norm)

    def _grad(self, v, vb):
        """stock leastSquaresGrad [OF-lib]; in-tree twin zeroBoundaryGrad.C:90-99 + the boundary lines kept in
        comments at :126-133."""
        grad = np.zeros((self.nc, 3))
        d = v[self.nei] - v[self.own]
        np.add.at(grad, self.own, self.ownLs * d[:, None])
        np.add.at(grad, self.nei, -self.neiLs * d[:, None])
        np.add.at(grad, self.bown, self.patchLs * (vb - v[self.bown])[:, None])
        return grad

    # ---- one time step = fvDVM::evolution(), fvDVM.C:1086-1108 -------------------------------------------
    def step(self, dt):
        nif, nbf, nc = self.nif, self.nbf, self.nc
        Sfi, Sfb = self.Sf[:nif], self.Sf[nif:]
        nb = Sfb / np.linalg.norm(Sfb, axis=1)[:, None]
        kind = self.kind
        is_wall, is_mixed, is_zg, is_sym = kind == MAXWELL_WALL, kind == MIXED, kind == ZERO_GRADIENT, kind == DVM_SYMMETRY
        is_far = (kind == FAR_FIELD) | (kind == PRESSURE_IN) | (kind == PRESSURE_OUT)
        if (kind == SYMMETRY_PLANE).any():
            raise NotImplementedError("symmetryPlane is a constraint patch [OF-lib]; see the C oracle")
        gB = np.empty((self.nxi, nc)); hB = np.empty((self.nxi, nc))
        outGoing = np.zeros(nbf)
        rf_c = 1.5 * dt / (2.0 * self.tau + dt)                                  # discreteVelocity.C:393
        for k in range(self.nxi):
            xi = self.xi[k]
            # 1  updateGHbarPvol  discreteVelocity.C:346-410
            gEq, hEq = self._shakhov(xi, self.rho, self.U, self.T, self.q)
            gB[k] = (1.0 - rf_c) * self.gT[k] + rf_c * gEq
            hB[k] = (1.0 - rf_c) * self.hT[k] + rf_c * hEq
            # correctBoundaryConditions of a fixedGradient patch [OF-lib]: value = cell + gradient / deltaCoeffs
            gBb = gB[k][self.bown] + self.gam_g[k] / self.dc[nif:]
            hBb = hB[k][self.bown] + self.gam_h[k] / self.dc[nif:]
            # 2.1  updateGHbarSurf  :412-691
            gG, hG = self._grad(gB[k], gBb), self._grad(hB[k], hBb)               # :420-421
            # boundary value of the gradient = cell value (zeroGradient); its normal part is next step's gradient()
            self.gam_g[k] = (gG[self.bown] * nb).sum(axis=1)                      # :462-468
            self.gam_h[k] = (hG[self.bown] * nb).sum(axis=1)
            phi = xi_dot_sf(xi, Sfi)
            ro = self.Cf[:nif] - self.C[self.own] - 0.5 * xi * dt
            rn = self.Cf[:nif] - self.C[self.nei] - 0.5 * xi * dt
            go = gB[k][self.own] + (gG[self.own] * ro).sum(axis=1); gn = gB[k][self.nei] + (gG[self.nei] * rn).sum(axis=1)
            ho = hB[k][self.own] + (hG[self.own] * ro).sum(axis=1); hn = hB[k][self.nei] + (hG[self.nei] * rn).sum(axis=1)
            up_o, up_n = phi >= VSMALL, phi < -VSMALL                             # :495, :506, else :513-529
            self.gS[k, :nif] = np.where(up_o, go, np.where(up_n, gn, 0.5 * (gn + go)))
            self.hS[k, :nif] = np.where(up_o, ho, np.where(up_n, hn, 0.5 * (hn + ho)))
            phib = xi_dot_sf(xi, Sfb)
            rb = self.Cf[nif:] - self.C[self.bown] - 0.5 * xi * dt
            gout = gB[k][self.bown] + (gG[self.bown] * rb).sum(axis=1)
            hout = hB[k][self.bown] + (hG[self.bown] * rb).sum(axis=1)
            gSb, hSb = self.gS[k, nif:], self.hS[k, nif:]                         # views
            gSb[is_zg] = gB[k][self.bown][is_zg]; hSb[is_zg] = hB[k][self.bown][is_zg]    # :551-555
            m = (is_mixed | is_far | is_wall) & (phib > 0)                        # :562, :580, :614
            gSb[m] = gout[m]; hSb[m] = hout[m]
            m = is_far & ~(phib > 0)                                              # :590-603: rho_b, owner-cell U, T_b
            geq = self.rho_b[m] * self._maxwell_by_rho(xi, self.U[self.bown[m]], self.T_b[m])
            gSb[m] = geq; hSb[m] = geq * (self.R * self.T_b[m]) * (self.K + 3 - self.D)
            m = is_wall & (phib > 0)
            outGoing[m] += self.w[k] * phib[m] * gSb[m]                           # :623-624
            m = is_sym & (phib > -VSMALL)                                         # :675
            gSb[m] = gout[m]; hSb[m] = hout[m]
        # 2.2  updateMaxwellWallRho  fvDVM.C:347-367, calculatedMaxwellFvPatchField.C:139-164
        self.rho_b[is_wall] = outGoing[is_wall] / np.abs(self.inByRho[is_wall])
        # 2.3  updateGHbarSurfMaxwellWallIn  discreteVelocity.C:693-731
        for k in range(self.nxi):
            m = is_wall & (xi_dot_sf(self.xi[k], Sfb) <= 0)
            geq = self.rho_b[m] * self._maxwell_by_rho(self.xi[k], self.U_b[m], self.T_b[m])
            self.gS[k, nif:][m] = geq
            self.hS[k, nif:][m] = geq * (self.R * self.T_b[m]) * (self.K + 3 - self.D)
        # 2.4  updateGHbarSurfSymmetryIn  fvDVM.C:375-454, discreteVelocity.C:733-817
        for p in self.patches:
            if p.kind != DVM_SYMMETRY or p.size <= 0:
                continue
            sl = slice(nif + p.start, nif + p.start + p.size)
            snapG, snapH = self.gS[:, sl].copy(), self.hS[:, sl].copy()           # dfContainer after the Allgatherv
            Sf0 = Sfb[p.start]
            n0 = Sf0 / np.linalg.norm(Sf0)
            for k in range(self.nxi):
                if xi_dot_sf(self.xi[k], Sf0[None, :])[0] <= 0:                   # :775
                    tgt = int(round(abs(n0 @ self.mirror[k])))                    # :777-779
                    self.gS[k, sl] = snapG[tgt]; self.hS[k, sl] = snapH[tgt]
        # 3  updateMacroSurf  fvDVM.C:456-582
        w, xi = self.w, self.xi
        rhoS = np.einsum("k,kf->f", w, self.gS)
        rhoUS = np.einsum("k,kf,kd->fd", w, self.gS, xi)
        rhoES = 0.5 * np.einsum("k,kf->f", w, self.gS * (xi ** 2).sum(axis=1)[:, None] + self.hS)
        US = rhoUS / rhoS[:, None]
        TS = (rhoES - 0.5 * rhoS * (US ** 2).sum(axis=1)) / ((self.K + 3) / 2.0 * self.R * rhoS)
        tauS = self._tau(TS, rhoS)
        qS = np.zeros_like(US)
        for k in range(self.nxi):
            c = xi[k] - US
            qS += 0.5 * w[k] * c * ((c ** 2).sum(axis=1) * self.gS[k] + self.hS[k])[:, None]     # :503-516
        qS *= (2.0 * tauS / (2.0 * tauS + 0.5 * dt * self.Pr))[:, None]            # :522
        self.rhoSurf, self.Usurf, self.Tsurf, self.tauSurf, self.qSurf = rhoS, US, TS, tauS, qS
        self.qWall[:] = 0.0; self.stressWall[:] = 0.0                              # :539-581
        bw = np.where(is_wall)[0]
        for k in range(self.nxi):
            c = xi[k] - self.U_b[bw]
            gk, hk = self.gS[k, nif + bw], self.hS[k, nif + bw]
            self.qWall[bw] += 0.5 * w[k] * c * ((c ** 2).sum(axis=1) * gk + hk)[:, None]
            self.stressWall[bw] += (w[k] * gk)[:, None, None] * np.outer(xi[k], xi[k])[None]
        tw = tauS[nif + bw]
        self.qWall[bw] *= (2.0 * tw / (2.0 * tw + 0.5 * dt * self.Pr))[:, None]
        self.stressWall[bw] *= (2.0 * tw / (2.0 * tw + 0.5 * dt))[:, None, None]
        # 4  updateGHsurf  discreteVelocity.C:819-932;  5  updateGHtildeVol  :934-978
        h = 0.5 * dt
        rfS = h / (2.0 * tauS + h)                                                 # :867
        for k in range(self.nxi):
            gEq, hEq = self._shakhov(xi[k], rhoS, US, TS, qS)                      # :870-878
            gI = (1.0 - rfS[:nif]) * self.gS[k, :nif] + rfS[:nif] * gEq[:nif]      # :880-881 (internal faces)
            hI = (1.0 - rfS[:nif]) * self.hS[k, :nif] + rfS[:nif] * hEq[:nif]
            self.gS[k, :nif] = gI; self.hS[k, :nif] = hI
            phib = xi_dot_sf(xi[k], Sfb)
            gb, hb = self.gS[k, nif:], self.hS[k, nif:]
            rb_, geb, heb = rfS[nif:], gEq[nif:], hEq[nif:]
            m = phib > 0                                                           # :905-920 outgoing only
            gb[m] = (1.0 - rb_[m]) * gb[m] + rb_[m] * geb[m]
            hb[m] = (1.0 - rb_[m]) * hb[m] + rb_[m] * heb[m]
            m = is_sym                                                             # :922-930 whole patch, again
            gb[m] = (1.0 - rb_[m]) * gb[m] + rb_[m] * geb[m]
            hb[m] = (1.0 - rb_[m]) * hb[m] + rb_[m] * heb[m]
            gt = -1.0 / 3 * self.gT[k] + 4.0 / 3 * gB[k]                           # :937-938
            ht = -1.0 / 3 * self.hT[k] + 4.0 / 3 * hB[k]
            phi = xi_dot_sf(xi[k], Sfi)
            np.subtract.at(gt, self.own, phi * gI * dt / self.V[self.own])         # :948-956
            np.add.at(gt, self.nei, phi * gI * dt / self.V[self.nei])
            np.subtract.at(ht, self.own, phi * hI * dt / self.V[self.own])
            np.add.at(ht, self.nei, phi * hI * dt / self.V[self.nei])
            np.subtract.at(gt, self.bown, phib * gb * dt / self.V[self.bown])      # :959-975
            np.subtract.at(ht, self.bown, phib * hb * dt / self.V[self.bown])
            self.gT[k], self.hT[k] = gt, ht
        # 6  updateMacroVol  fvDVM.C:597-728 (macroFlux "no")
        rho = np.einsum("k,kc->c", w, self.gT)
        rhoU = np.einsum("k,kc,kd->cd", w, self.gT, xi)
        rhoE = 0.5 * np.einsum("k,kc->c", w, (xi ** 2).sum(axis=1)[:, None] * self.gT + self.hT)
        self.rho = rho
        self.U = rhoU / rho[:, None]
        self.T = (rhoE - 0.5 * rho * (self.U ** 2).sum(axis=1)) / ((self.K + 3) / 2.0 * self.R * rho)
        for p in self.patches:                                                     # :698-699 correctBoundaryConditions
            sl = slice(p.start, p.start + p.size)
            if p.U_bc == 1:
                self.U_b[sl] = self.U[self.bown[sl]]
            if p.T_bc == 1:
                self.T_b[sl] = self.T[self.bown[sl]]
        self.tau = self._tau(self.T, self.rho)                                     # :706
        self.q = np.zeros((nc, 3))
        for k in range(self.nxi):
            c = xi[k] - self.U
            self.q += 0.5 * w[k] * c * ((c ** 2).sum(axis=1) * self.gT[k] + self.hT[k])[:, None]   # :712-721
        self.q *= (2.0 * self.tau / (2.0 * self.tau + dt * self.Pr))[:, None]      # :727
        # 7  updatePressureInOutBC  :730-806
        self._pressure_bc()

    def courant(self, dt):                                                         # fvDVM::getCoNum, fvDVM.C:1111-1119
        UbyDx = self.dc[: self.nif] * (np.linalg.norm(self.Usurf[: self.nif], axis=1) + np.sqrt(self.D) * self.xiMax)
        return float(UbyDx.max() * dt), float(UbyDx.mean() * dt)
