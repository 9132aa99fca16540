"""ctypes mirror of include/dugks.h (struct layouts and marshalling of a Case)."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import numpy as np

from .case import Case

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)

ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)


class PatchT(C.Structure):
    _fields_ = [("kind", C.c_int32), ("start", C.c_int32), ("size", C.c_int32),
                ("U_bc", C.c_int32), ("T_bc", C.c_int32), ("reserved", C.c_int32),
                ("pressure", C.c_double)]


class MeshT(C.Structure):
    _fields_ = [("nCells", C.c_int32), ("nInternalFaces", C.c_int32),
                ("nBoundaryFaces", C.c_int32), ("nSolutionD", C.c_int32),
                ("owner", c_int32_p), ("neighbour", c_int32_p),
                ("C", c_double_p), ("V", c_double_p), ("Cf", c_double_p), ("Sf", c_double_p),
                ("ownLs", c_double_p), ("neiLs", c_double_p), ("patchLs", c_double_p),
                ("deltaCoeffs", c_double_p)]


class DvsetT(C.Structure):
    _fields_ = [("nXiPerDim", C.c_int32), ("reserved", C.c_int32),
                ("Xis", c_double_p), ("weights", c_double_p),
                ("xiMax", C.c_double), ("xiMin", C.c_double)]


class GasT(C.Structure):
    _fields_ = [("R", C.c_double), ("omega", C.c_double), ("Tref", C.c_double),
                ("muRef", C.c_double), ("Pr", C.c_double), ("KInner", C.c_int32),
                ("reserved", C.c_int32)]


class ParT(C.Structure):
    _fields_ = [("rank", C.c_int32), ("nRanks", C.c_int32), ("device", C.c_int32),
                ("partition", C.c_int32), ("reduce", ALLREDUCE_FN), ("reduce_user", C.c_void_p),
                ("nccl_unique_id", C.c_void_p), ("scratch_bytes", C.c_size_t),
                ("store_h", C.c_int32), ("dv_chunk", C.c_int32), ("limiter_k", C.c_double)]


class StatsT(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("steps", C.c_uint64),
                ("device_bytes", C.c_uint64), ("h_elided", C.c_int32), ("n_slabs", C.c_int32),
                ("slab_dvs", C.c_int32), ("keep_slabs", C.c_int32), ("pencil_cells", C.c_int32),
                ("pencil_mode", C.c_int32)]


def dptr(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_double_p)


def iptr(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(c_int32_p)


class Marshalled:
    """ctypes views of a Case; keeps the numpy arrays alive."""

    def __init__(self, case: Case):
        g = case.geom
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
        self.keep = dict(owner=i32(g.owner), neighbour=i32(g.neighbour), C=f64(g.C), V=f64(g.V),
                         Cf=f64(g.Cf), Sf=f64(g.Sf), ownLs=f64(g.ownLs), neiLs=f64(g.neiLs),
                         patchLs=f64(g.patchLs), deltaCoeffs=f64(g.deltaCoeffs),
                         Xis=f64(case.Xis), weights=f64(case.weights),
                         rho=f64(case.rho), U=f64(case.U), T=f64(case.T),
                         rho_b=f64(case.rho_b), U_b=f64(case.U_b), T_b=f64(case.T_b))
        k = self.keep
        self.mesh = MeshT(g.nCells, g.nInternalFaces, g.nBoundaryFaces, g.nSolutionD,
                          iptr(k["owner"]), iptr(k["neighbour"]), dptr(k["C"]), dptr(k["V"]),
                          dptr(k["Cf"]), dptr(k["Sf"]), dptr(k["ownLs"]), dptr(k["neiLs"]),
                          dptr(k["patchLs"]), dptr(k["deltaCoeffs"]))
        self.npatch = len(case.patches)
        self.patches = (PatchT * max(self.npatch, 1))()
        for i, p in enumerate(case.patches):
            self.patches[i] = PatchT(p.kind, p.start, p.size, p.U_bc, p.T_bc, 0, p.pressure)
        self.dvset = DvsetT(len(case.Xis), 0, dptr(k["Xis"]), dptr(k["weights"]), case.xiMax, case.xiMin)
        gs = case.gas
        self.gas = GasT(gs["R"], gs["omega"], gs["Tref"], gs["muRef"], gs["Pr"], int(gs.get("KInner", 0)), 0)
        self.fields = (dptr(k["rho"]), dptr(k["U"]), dptr(k["T"]), dptr(k["rho_b"]), dptr(k["U_b"]), dptr(k["T_b"]))
