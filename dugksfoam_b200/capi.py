"""Python host side over the C-ABI (include/dugks.h): `fvDVM`, a mirror of the
reference's Foam::fvDVM facade (fvDVM.H:269-376) for the standalone harness.

The product path is the CUDA library `libdugks.so`; there is NO CPU fallback: if
the library is missing or no CUDA device is present, construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Callable, Optional

import numpy as np

from .abi import (ALLREDUCE_FN, DvsetT, GasT, Marshalled, MeshT, ParT, PatchT, StatsT, c_double_p,
                  c_int32_p, dptr, iptr)
from .case import Case

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DUGKS_LIB", os.path.join(_HERE, "libdugks.so"))   # DUGKS_LIB: A/B builds
_LIB = None

# 1886: the dynamic shared-memory array is declared with 16- and 128-byte alignment in different kernels of the
# one translation unit (a warning per instantiation, nothing else)
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-diag-suppress", "1886", "-Xcompiler", "-fPIC", "-shared"]

# every symbol include/dugks.h declares
EXPORTS = ["dugks_abi_version", "dugks_nccl_unique_id", "dugks_create", "dugks_destroy", "dugks_last_error",
           "dugks_step", "dugks_sync", "dugks_set_boundary_macros", "dugks_get_cell_macros",
           "dugks_get_face_macros", "dugks_get_boundary_macros", "dugks_get_wall_diag", "dugks_courant",
           "dugks_get_df", "dugks_get_state", "dugks_set_state", "dugks_local_dvs", "dugks_sizes",
           "dugks_get_stats", "dugks_stream", "dugks_kernel_timing", "dugks_partition",
           "dugks_get_boundary_df", "dugks_row_layout", "dugks_convergence", "dugks_cell_order",
           "dugks_checkpoint_size", "dugks_checkpoint_save", "dugks_checkpoint_load", "dugks_pencil_plan",
           "dugks_host_register", "dugks_host_unregister"]


class DugksError(RuntimeError):
    pass


def build_library(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> dugksfoam_b200/libdugks.so (in tree)."""
    src = os.path.join(_HERE, "csrc", "dugks_capi.cu")
    deps = [src, os.path.join(_HERE, "csrc", "dugks_kernels.cuh"), os.path.join(_HERE, "csrc", "dugks_device.cuh"),
            os.path.join(_HERE, "csrc", "dugks_fast.cuh"), os.path.join(_HERE, "csrc", "dugks_tma.cuh"),
            os.path.join(_HERE, "csrc", "dugks_hot.cuh"), os.path.join(_HERE, "csrc", "dugks_pencil.cuh"),
            os.path.join(_HERE, "..", "include", "dugks.h")]
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= max(os.path.getmtime(d) for d in deps):
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH, src, "-ldl"]
    subprocess.check_call(cmd)
    return LIB_PATH


def load_library():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise DugksError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                         "(dugksfoam_b200 has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    L.dugks_abi_version.restype = C.c_int
    L.dugks_last_error.restype = C.c_char_p
    L.dugks_last_error.argtypes = [C.c_void_p]
    L.dugks_nccl_unique_id.argtypes = [C.c_void_p]
    L.dugks_create.argtypes = [C.POINTER(MeshT), C.POINTER(PatchT), C.c_int32, C.POINTER(DvsetT), C.POINTER(GasT),
                               C.POINTER(ParT)] + [c_double_p] * 6 + [C.POINTER(C.c_void_p)]
    L.dugks_destroy.argtypes = [C.c_void_p]
    L.dugks_destroy.restype = None
    L.dugks_step.argtypes = [C.c_void_p, C.c_double]
    L.dugks_sync.argtypes = [C.c_void_p]
    L.dugks_set_boundary_macros.argtypes = [C.c_void_p] + [c_double_p] * 3
    L.dugks_get_cell_macros.argtypes = [C.c_void_p] + [c_double_p] * 5
    L.dugks_get_face_macros.argtypes = [C.c_void_p] + [c_double_p] * 5
    L.dugks_host_register.argtypes = [C.c_void_p, C.c_size_t]
    L.dugks_host_unregister.argtypes = [C.c_void_p]
    L.dugks_get_boundary_macros.argtypes = [C.c_void_p] + [c_double_p] * 3
    L.dugks_get_wall_diag.argtypes = [C.c_void_p] + [c_double_p] * 2
    L.dugks_courant.argtypes = [C.c_void_p, C.c_double, c_double_p, c_double_p]
    L.dugks_convergence.argtypes = [C.c_void_p, c_double_p]
    L.dugks_get_df.argtypes = [C.c_void_p, C.c_int32, c_double_p, c_double_p]
    L.dugks_get_state.argtypes = [C.c_void_p, c_double_p, c_double_p]
    L.dugks_set_state.argtypes = [C.c_void_p, c_double_p, c_double_p]
    L.dugks_local_dvs.argtypes = [C.c_void_p, c_int32_p, c_int32_p]
    L.dugks_sizes.argtypes = [C.c_void_p] + [c_int32_p] * 4
    L.dugks_get_stats.argtypes = [C.c_void_p, C.POINTER(StatsT)]
    L.dugks_stream.argtypes = [C.c_void_p]
    L.dugks_stream.restype = C.c_void_p
    L.dugks_kernel_timing.argtypes = [C.c_void_p, C.c_int, C.c_int, c_double_p, C.POINTER(C.c_uint64)]
    L.dugks_partition.argtypes = [C.c_int32] * 4 + [c_int32_p, c_int32_p]
    L.dugks_get_boundary_df.argtypes = [C.c_void_p, c_double_p, c_double_p]
    L.dugks_row_layout.argtypes = [C.c_int32] * 4 + [c_int32_p] * 8
    L.dugks_cell_order.argtypes = [C.c_int32, C.c_int32, c_double_p, C.POINTER(C.c_uint8), C.c_char_p, C.c_int32, c_int32_p]
    L.dugks_pencil_plan.argtypes = [C.POINTER(MeshT), C.c_int32, c_int32_p, c_int32_p, c_int32_p, c_int32_p, c_int32_p, c_int32_p]
    L.dugks_checkpoint_size.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    L.dugks_checkpoint_save.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    L.dugks_checkpoint_load.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
    _LIB = L
    return L


def partition(nXiPerDim: int, nSolutionD: int, nRanks: int, rank: int) -> np.ndarray:
    """Global ids of the discrete velocities rank `rank` owns (host-only, needs no device)."""
    L = load_library()
    n = C.c_int32()
    rc = L.dugks_partition(nXiPerDim, nSolutionD, nRanks, rank, None, C.byref(n))
    if rc:
        raise DugksError(f"dugks_partition failed ({rc}): {L.dugks_last_error(None).decode()}")
    ids = np.empty(n.value, dtype=np.int32)
    L.dugks_partition(nXiPerDim, nSolutionD, nRanks, rank, iptr(ids), C.byref(n))
    return ids


def row_layout(nXiPerDim: int, nSolutionD: int, nRanks: int, rank: int) -> dict:
    """Velocity-row layout of one rank (host-only, needs no device): see dugks_row_layout."""
    L = load_library()
    nch, ll, lt, nrows = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32(0)
    rc = L.dugks_row_layout(nXiPerDim, nSolutionD, nRanks, rank, C.byref(nch), C.byref(ll), C.byref(lt), C.byref(nrows),
                            None, None, None, None)
    if rc:
        raise DugksError(f"dugks_row_layout failed ({rc}): {L.dugks_last_error(None).decode()}")
    n = nrows.value
    arr = [np.empty(n, dtype=np.int32) for _ in range(4)]
    cap = C.c_int32(n)
    L.dugks_row_layout(nXiPerDim, nSolutionD, nRanks, rank, C.byref(nch), C.byref(ll), C.byref(lt), C.byref(cap),
                       *[iptr(a) for a in arr])
    return dict(nch=nch.value, L=ll.value, Lt=lt.value, iy=arr[0], iz=arr[1], first=arr[2], len=arr[3])


def cell_order(centres: np.ndarray, nSolutionD: int, kind: str = "tiled", nWarps: int = 1776,
               first_class: Optional[np.ndarray] = None) -> np.ndarray:
    """Traversal order of the cell kernels (host-only, needs no device): see dugks_cell_order."""
    L = load_library()
    Cc = np.ascontiguousarray(centres, dtype=np.float64)
    n = Cc.shape[0]
    out = np.empty(n, dtype=np.int32)
    fc = None
    if first_class is not None:
        fc = np.ascontiguousarray(first_class, dtype=np.uint8)
    rc = L.dugks_cell_order(n, nSolutionD, dptr(Cc), None if fc is None else fc.ctypes.data_as(C.POINTER(C.c_uint8)),
                            kind.encode(), nWarps, iptr(out))
    if rc:
        raise DugksError(f"dugks_cell_order failed ({rc}): {L.dugks_last_error(None).decode()}")
    return out


def pencil_plan(case: Case, nCtas: int = 296) -> dict:
    """CTA-pencil plan of phase 1 (host-only, needs no device): see dugks_pencil_plan."""
    L = load_library()
    m = Marshalled(case)
    ni, npc = C.c_int32(0), C.c_int32(0)
    nax = C.c_int32(0)
    rc = L.dugks_pencil_plan(C.byref(m.mesh), nCtas, C.byref(ni), C.byref(npc), None, None, None, C.byref(nax))
    if rc:
        raise DugksError(f"dugks_pencil_plan failed ({rc}): {L.dugks_last_error(None).decode()}")
    cells = np.empty(npc.value, dtype=np.int32)
    first, steps = np.empty(ni.value, dtype=np.int32), np.empty(ni.value, dtype=np.int32)
    L.dugks_pencil_plan(C.byref(m.mesh), nCtas, C.byref(ni), C.byref(npc), iptr(cells), iptr(first), iptr(steps), None)
    return dict(cells=cells, item_first=first, item_steps=steps, n_axis=nax.value)


def nccl_unique_id() -> bytes:
    L = load_library()
    buf = C.create_string_buffer(128)
    rc = L.dugks_nccl_unique_id(buf)
    if rc:
        raise DugksError(f"dugks_nccl_unique_id failed ({rc}): {L.dugks_last_error(None).decode()}")
    return buf.raw


class fvDVM:
    """Mirror of Foam::fvDVM (fvDVM.H): construct from the macro fields, call
    evolution() once per time step, read the macro fields back.

    rank/nranks: velocity-space decomposition (the reference's -dvParallel,
    fvDVM.C:228-260).  The collective is either the library's own NCCL
    communicator (pass `nccl_id`, 128 bytes broadcast from rank 0) or a Python
    callable `reduce(ptr, n, stream)` that sum-reduces n doubles at device
    pointer ptr in place (the role of fieldMPIreducer::reduceField)."""

    def __init__(self, case: Case, *, rank: int = 0, nranks: int = 1, device: int = -1,
                 nccl_id: Optional[bytes] = None, reduce: Optional[Callable[[int, int, int], int]] = None,
                 store_h: bool = False, scratch_bytes: int = 0, limiter_k: float = 0.0):
        self.L = load_library()
        self.case = case
        self._m = Marshalled(case)
        self._cb = None
        par = ParT()
        par.rank, par.nRanks, par.device, par.partition = rank, nranks, device, 0
        par.scratch_bytes = int(scratch_bytes)
        par.store_h = 1 if store_h else 0
        par.dv_chunk = 0
        par.limiter_k = float(limiter_k)
        if reduce is not None:
            def _cb(user, ptr, n, stream):
                try:
                    return int(reduce(ptr, n, stream) or 0)
                except Exception:  # never let an exception cross the ABI
                    import traceback
                    traceback.print_exc()
                    return -1
            self._cb = ALLREDUCE_FN(_cb)
            par.reduce = self._cb
        self._idbuf = None
        if nccl_id is not None:
            self._idbuf = C.create_string_buffer(nccl_id, 128)
            par.nccl_unique_id = C.cast(self._idbuf, C.c_void_p)
        self._par = par
        h = C.c_void_p()
        rc = self.L.dugks_create(C.byref(self._m.mesh), self._m.patches, self._m.npatch, C.byref(self._m.dvset),
                                 C.byref(self._m.gas), C.byref(par), *self._m.fields, C.byref(h))
        if rc:
            raise DugksError(f"dugks_create failed ({rc}): {self.L.dugks_last_error(None).decode()}")
        self.h = h
        n = (C.c_int32 * 4)()
        self.L.dugks_sizes(self.h, *(C.cast(C.byref(n, 4 * i), c_int32_p) for i in range(4)))
        self._nXi, self.nXiLocal, self.nCells, self.nFaces = (int(v) for v in n)
        self.nBoundaryFaces = case.geom.nBoundaryFaces
        self._pinned_cm = None

    # -- lifetime -----------------------------------------------------------
    def close(self):
        if getattr(self, "_pinned_cm", None):
            for a in self._pinned_cm.values():
                self.L.dugks_host_unregister(a.ctypes.data_as(C.c_void_p))   # a no-op error for arrays that never got locked
            self._pinned_cm = None
        if getattr(self, "h", None):
            self.L.dugks_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc: int, what: str):
        if rc:
            raise DugksError(f"{what} failed ({rc}): {self.L.dugks_last_error(self.h).decode()}")

    # -- the hot path ---------------------------------------------------------
    def evolution(self, dt: float):
        """fvDVM::evolution() (fvDVM.C:1086-1108); dt = runTime.deltaTValue()."""
        self._chk(self.L.dugks_step(self.h, float(dt)), "dugks_step")

    def sync(self):
        self._chk(self.L.dugks_sync(self.h), "dugks_sync")

    def getCoNum(self, dt: float):
        """fvDVM::getCoNum (fvDVM.C:1111-1119) -> (maxCoNum, meanCoNum)."""
        a, b = C.c_double(), C.c_double()
        self._chk(self.L.dugks_courant(self.h, float(dt), C.byref(a), C.byref(b)), "dugks_courant")
        return a.value, b.value

    def convergence(self):
        """(TemperatureChange, rhoChange, Uchange) of the time loop's convergence monitor
        (dugksFoam.C:88-107) since the previous call; the first call compares with the initial fields."""
        out = np.empty(3)
        self._chk(self.L.dugks_convergence(self.h, dptr(out)), "dugks_convergence")
        return tuple(float(v) for v in out)

    # -- accessors (fvDVM.H:309-364) -----------------------------------------
    def nXi(self) -> int:
        return self._nXi

    def nXiPerDim(self) -> int:
        return self.case.nXiPerDim

    def _macros(self, fn, n):
        rho, U, T, q, tau = np.empty(n), np.empty((n, 3)), np.empty(n), np.empty((n, 3)), np.empty(n)
        self._chk(fn(self.h, dptr(rho), dptr(U), dptr(T), dptr(q), dptr(tau)), fn.__name__)
        return dict(rho=rho, U=U, T=T, q=q, tau=tau)

    def cell_macros(self, pinned: bool = False):
        """rhoVol(), Uvol(), Tvol(), qVol(), tauVol().  pinned: fill (and return) one set of page-locked arrays owned by
        this object - what a solver does with its field storage (dugks_host_register); valid until the next call."""
        if not pinned:
            return self._macros(self.L.dugks_get_cell_macros, self.nCells)
        if self._pinned_cm is None:
            n = self.nCells
            out = dict(rho=np.zeros(n), U=np.zeros((n, 3)), T=np.zeros(n), q=np.zeros((n, 3)), tau=np.zeros(n))
            self._pinned_ok = True
            for a in out.values():
                # page-locking is an optimisation: where the host refuses it (locked-memory limit) the accessor stages
                if self.L.dugks_host_register(a.ctypes.data_as(C.c_void_p), a.nbytes):
                    self._pinned_ok = False
            self._pinned_cm = out
        o = self._pinned_cm
        self._chk(self.L.dugks_get_cell_macros(self.h, dptr(o["rho"]), dptr(o["U"]), dptr(o["T"]), dptr(o["q"]), dptr(o["tau"])),
                  "dugks_get_cell_macros")
        return o

    def face_macros(self):
        """rhoSurf(), Usurf(), Tsurf(), qSurf(), tauSurf() on internal + non-empty boundary faces."""
        return self._macros(self.L.dugks_get_face_macros, self.nFaces)

    def rhoVol(self): return self.cell_macros()["rho"]
    def Uvol(self): return self.cell_macros()["U"]
    def Tvol(self): return self.cell_macros()["T"]
    def qVol(self): return self.cell_macros()["q"]
    def tauVol(self): return self.cell_macros()["tau"]

    def boundary_macros(self):
        n = self.nBoundaryFaces
        rho, U, T = np.empty(n), np.empty((n, 3)), np.empty(n)
        self._chk(self.L.dugks_get_boundary_macros(self.h, dptr(rho), dptr(U), dptr(T)), "dugks_get_boundary_macros")
        return dict(rho=rho, U=U, T=T)

    def set_boundary_macros(self, rho_b=None, U_b=None, T_b=None):
        f = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        rho_b, U_b, T_b = f(rho_b), f(U_b), f(T_b)
        self._chk(self.L.dugks_set_boundary_macros(self.h, dptr(rho_b), dptr(U_b), dptr(T_b)),
                  "dugks_set_boundary_macros")

    def wall_diag(self):
        n = self.nBoundaryFaces
        q, s = np.empty((n, 3)), np.empty((n, 9))
        self._chk(self.L.dugks_get_wall_diag(self.h, dptr(q), dptr(s)), "dugks_get_wall_diag")
        return dict(qWall=q, stressWall=s)

    def local_dvs(self) -> np.ndarray:
        n = C.c_int32()
        self.L.dugks_local_dvs(self.h, None, C.byref(n))
        ids = np.empty(n.value, dtype=np.int32)
        self._chk(self.L.dugks_local_dvs(self.h, iptr(ids), C.byref(n)), "dugks_local_dvs")
        return ids

    def state(self):
        """(gTildeVol, hTildeVol) of the local DVs, [nXiLocal, nCells] each, ordered by global DV id."""
        g = np.empty((self.nXiLocal, self.nCells))
        h = np.empty((self.nXiLocal, self.nCells))
        self._chk(self.L.dugks_get_state(self.h, dptr(g), dptr(h)), "dugks_get_state")
        return g, h

    def set_state(self, g, h=None):
        g = np.ascontiguousarray(g, dtype=np.float64)
        h = None if h is None else np.ascontiguousarray(h, dtype=np.float64)
        self._chk(self.L.dugks_set_state(self.h, dptr(g), dptr(h)), "dugks_set_state")

    def checkpoint(self) -> np.ndarray:
        """Everything the next evolution() reads, as one opaque rank-local blob (dugks_checkpoint_save)."""
        n = C.c_uint64()
        self._chk(self.L.dugks_checkpoint_size(self.h, C.byref(n)), "dugks_checkpoint_size")
        buf = np.empty(n.value, dtype=np.uint8)
        self._chk(self.L.dugks_checkpoint_save(self.h, buf.ctypes.data_as(C.c_void_p), n.value), "dugks_checkpoint_save")
        return buf

    def restore(self, blob: np.ndarray):
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        self._chk(self.L.dugks_checkpoint_load(self.h, blob.ctypes.data_as(C.c_void_p), blob.size), "dugks_checkpoint_load")

    def writeDFonCell(self, cell: int):
        """fvDVM::writeDFonCell (fvDVM.C:820-875): gTilde (and hTilde) of one cell for all global DVs."""
        g, h = np.empty(self._nXi), np.empty(self._nXi)
        self._chk(self.L.dugks_get_df(self.h, int(cell), dptr(g), dptr(h)), "dugks_get_df")
        return g, h

    def boundary_surf(self):
        g = np.empty((self.nXiLocal, self.nBoundaryFaces))
        h = np.empty((self.nXiLocal, self.nBoundaryFaces))
        self._chk(self.L.dugks_get_boundary_df(self.h, dptr(g), dptr(h)), "dugks_get_boundary_df")
        return g, h

    # -- instrumentation ------------------------------------------------------
    def stats(self) -> dict:
        s = StatsT()
        self._chk(self.L.dugks_get_stats(self.h, C.byref(s)), "dugks_get_stats")
        return {k: int(getattr(s, k)) for k, _ in StatsT._fields_ if k != "reserved"}

    def stream(self) -> int:
        return int(self.L.dugks_stream(self.h) or 0)

    def kernel_timing(self, enable: int, which: int = 0):
        ms, n = C.c_double(), C.c_uint64()
        self._chk(self.L.dugks_kernel_timing(self.h, enable, which, C.byref(ms), C.byref(n)), "dugks_kernel_timing")
        return ms.value, int(n.value)
