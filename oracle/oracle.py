"""ctypes wrapper of oracle/libdugks_oracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module (see dugks_oracle.c header).  PARITY UNPINNED: the
reference ships no golden outputs for this path.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from dugksfoam_b200.abi import DvsetT, GasT, Marshalled, MeshT, PatchT, c_double_p, c_int32_p, dptr, iptr
from dugksfoam_b200.case import Case

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libdugks_oracle.so")
    src = os.path.join(_HERE, "dugks_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "dugks.h")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libdugks_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "libdugks_oracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.POINTER(MeshT), C.POINTER(PatchT), C.c_int, C.POINTER(DvsetT),
                                    C.POINTER(GasT), C.c_int, C.c_int] + [c_double_p] * 6
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_step.argtypes = [C.c_void_p, C.c_double]
        L.oracle_courant.argtypes = [C.c_void_p, C.c_double, c_double_p, c_double_p]
        L.oracle_convergence.argtypes = [C.c_void_p, c_double_p]
        L.oracle_nxi.argtypes = [C.c_void_p]
        L.oracle_nxi.restype = C.c_int
        for name, n in (("oracle_get_cell_macros", 5), ("oracle_get_face_macros", 5),
                        ("oracle_get_boundary_macros", 3), ("oracle_get_wall_diag", 2),
                        ("oracle_get_state", 2), ("oracle_set_state", 2)):
            getattr(L, name).argtypes = [C.c_void_p] + [c_double_p] * n
        L.oracle_get_surf.argtypes = [C.c_void_p, C.c_int, c_double_p, c_double_p]
        L.oracle_get_dvs.argtypes = [C.c_void_p, c_double_p, c_double_p, c_int32_p, c_int32_p, c_int32_p]
        L.oracle_get_wall_incoming.argtypes = [C.c_void_p, c_double_p]
        L.oracle_set_grad_scheme.argtypes = [C.c_void_p, C.c_int, C.c_double]
        _LIB = L
    return _LIB


class Oracle:
    """Stage-by-stage CPU restatement of fvDVM (fvDVM.C / discreteVelocity.C)."""

    def __init__(self, case: Case, nranks: int = 1, partition: int = 0):
        self.case = case
        self.m = Marshalled(case)
        L = lib()
        self.h = L.oracle_create(C.byref(self.m.mesh), self.m.patches, self.m.npatch, C.byref(self.m.dvset),
                                 C.byref(self.m.gas), nranks, partition, *self.m.fields)
        self.nc = case.geom.nCells
        self.nf = case.geom.nFaces
        self.nbf = case.geom.nBoundaryFaces
        self.nxi = L.oracle_nxi(self.h)

    def close(self):
        if self.h:
            lib().oracle_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_grad_scheme(self, limiter: int, k: float = 0.0):
        """0: leastSquares; 1: VenkatakrishnanLimited leastSquares k as it is meant to work (the reference's own
        implementation is inert, VenkatakrishnanLimitedGrads.C:76,225)."""
        lib().oracle_set_grad_scheme(self.h, int(limiter), float(k))

    def step(self, dt: float):
        lib().oracle_step(self.h, dt)

    def courant(self, dt: float):
        a, b = C.c_double(), C.c_double()
        lib().oracle_courant(self.h, dt, C.byref(a), C.byref(b))
        return a.value, b.value

    def convergence(self):
        """(TemperatureChange, rhoChange, Uchange) since the previous call (dugksFoam.C:88-107); first call: since create."""
        out = np.empty(3)
        lib().oracle_convergence(self.h, dptr(out))
        return tuple(float(v) for v in out)

    def cell_macros(self):
        nc = self.nc
        rho, U, T, q, tau = np.empty(nc), np.empty((nc, 3)), np.empty(nc), np.empty((nc, 3)), np.empty(nc)
        lib().oracle_get_cell_macros(self.h, dptr(rho), dptr(U), dptr(T), dptr(q), dptr(tau))
        return dict(rho=rho, U=U, T=T, q=q, tau=tau)

    def face_macros(self):
        nf = self.nf
        rho, U, T, q, tau = np.empty(nf), np.empty((nf, 3)), np.empty(nf), np.empty((nf, 3)), np.empty(nf)
        lib().oracle_get_face_macros(self.h, dptr(rho), dptr(U), dptr(T), dptr(q), dptr(tau))
        return dict(rho=rho, U=U, T=T, q=q, tau=tau)

    def boundary_macros(self):
        n = self.nbf
        rho, U, T = np.empty(n), np.empty((n, 3)), np.empty(n)
        lib().oracle_get_boundary_macros(self.h, dptr(rho), dptr(U), dptr(T))
        return dict(rho=rho, U=U, T=T)

    def wall_diag(self):
        n = self.nbf
        q, s = np.empty((n, 3)), np.empty((n, 9))
        lib().oracle_get_wall_diag(self.h, dptr(q), dptr(s))
        return dict(qWall=q, stressWall=s)

    def state(self):
        g, h = np.empty((self.nxi, self.nc)), np.empty((self.nxi, self.nc))
        lib().oracle_get_state(self.h, dptr(g), dptr(h))
        return g, h

    def set_state(self, g, h):
        g = np.ascontiguousarray(g, dtype=np.float64)
        h = np.ascontiguousarray(h, dtype=np.float64)
        lib().oracle_set_state(self.h, dptr(g), dptr(h))

    def surf(self, k: int):
        g, h = np.empty(self.nf), np.empty(self.nf)
        lib().oracle_get_surf(self.h, k, dptr(g), dptr(h))
        return g, h

    def dvs(self):
        n = self.nxi
        xi, w = np.empty((n, 3)), np.empty(n)
        sx, sy, sz = (np.empty(n, dtype=np.int32) for _ in range(3))
        lib().oracle_get_dvs(self.h, dptr(xi), dptr(w), iptr(sx), iptr(sy), iptr(sz))
        return xi, w, sx, sy, sz

    def wall_incoming(self):
        a = np.empty(self.nbf)
        lib().oracle_get_wall_incoming(self.h, dptr(a))
        return a
