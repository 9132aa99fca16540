"""Case description = what the fvDVM constructor sees (fvDVM.C:886-1075): mesh,
DVMProperties, Xis/weights, the rho/U/T fields and their patch types.

read_case() loads a reference case directory (demo/cavity ...); the *_case()
builders generate the synthetic configurations of BASELINE.json.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

from . import dvset as _dvset
from . import foam
from .polymesh import (Geometry, PolyMesh, compute_geometry, hex_block, ogrid_cylinder, ratchet_profile, tri_prism_2d,
                       voronoi_prism_2d)

# dugks_patch_kind (include/dugks.h)
PATCH_ZERO_GRADIENT, PATCH_MIXED, PATCH_MAXWELL_WALL, PATCH_FAR_FIELD = 0, 1, 2, 3
PATCH_DVM_SYMMETRY, PATCH_SYMMETRY_PLANE, PATCH_PRESSURE_IN, PATCH_PRESSURE_OUT = 4, 5, 6, 7
BC_FIXED_VALUE, BC_ZERO_GRADIENT = 0, 1

# rho patch-field type -> DF boundary kind, discreteVelocity.C:274-281
RHO_BC_TO_KIND = {
    "fixedValue": PATCH_MIXED,
    "zeroGradient": PATCH_ZERO_GRADIENT,
    "calculatedMaxwell": PATCH_MAXWELL_WALL,
    "farField": PATCH_FAR_FIELD,
    "symmetryMod": PATCH_DVM_SYMMETRY,
    "pressureIn": PATCH_PRESSURE_IN,
    "pressureOut": PATCH_PRESSURE_OUT,
    "symmetryPlane": PATCH_SYMMETRY_PLANE,   # constraint type, assigned automatically
}

# argon at Kn = 0.075, demo/cavity/constant/DVMProperties and demo/cavity/0/*
ARGON = dict(R=208.244343891, omega=0.81, Tref=273.0, muRef=1.60281882485e-05,
             Pr=0.666666666666667, KInner=0)
RHO0 = 1.14423514927e-06
T0 = 273.0


@dataclass
class PatchSpec:
    name: str
    kind: int
    start: int
    size: int
    U_bc: int = BC_FIXED_VALUE
    T_bc: int = BC_FIXED_VALUE
    pressure: float = 0.0


@dataclass
class Case:
    geom: Geometry
    patches: List[PatchSpec]
    Xis: np.ndarray
    weights: np.ndarray
    xiMax: float
    xiMin: float
    gas: Dict[str, float]
    rho: np.ndarray      # [nc]
    U: np.ndarray        # [nc, 3]
    T: np.ndarray        # [nc]
    rho_b: np.ndarray    # [nbf]
    U_b: np.ndarray      # [nbf, 3]
    T_b: np.ndarray      # [nbf]
    deltaT: float = 0.0
    maxCo: float = 0.8
    adjustTimeStep: bool = False
    name: str = ""
    mesh: Optional[PolyMesh] = field(default=None, repr=False)

    @property
    def nXiPerDim(self) -> int:
        return len(self.Xis)

    @property
    def nXi(self) -> int:
        return len(self.Xis) ** self.geom.nSolutionD

    @property
    def nCells(self) -> int:
        return self.geom.nCells

    def faces_per_cell(self) -> float:
        return self.geom.nFaces / self.geom.nCells

    def courant_dt(self, co: float) -> float:
        """dt such that fvDVM::getCoNum (fvDVM.C:1111-1119) with U = 0 equals co."""
        g = self.geom
        ubydx = g.deltaCoeffs[: g.nInternalFaces].max() * np.sqrt(g.nSolutionD) * self.xiMax
        return co / ubydx


def _bc_kind(tp: str) -> int:
    return BC_ZERO_GRADIENT if tp == "zeroGradient" else BC_FIXED_VALUE


def read_case(case_dir: str) -> Case:
    """Loads a reference case (e.g. /root/reference/demo/cavity)."""
    mesh = foam.read_polymesh(case_dir)
    geom = compute_geometry(mesh)
    props = foam.read_dict(os.path.join(case_dir, "constant", "DVMProperties"))
    paras, gasd = props["fvDVMparas"], props["gasProperties"]
    nDV = int(paras["nDV"][0])
    Xis = foam.read_scalar_list(os.path.join(case_dir, "constant", "Xis"))[:nDV]       # fvDVM.C:112-116
    weights = foam.read_scalar_list(os.path.join(case_dir, "constant", "weights"))[:nDV]
    gas = dict(R=foam.dimensioned_value(gasd["R"]), omega=float(gasd["omega"][0]),
               Tref=foam.dimensioned_value(gasd["Tref"]), muRef=foam.dimensioned_value(gasd["muRef"]),
               Pr=float(gasd["Pr"][0]), KInner=int(gasd["KInner"][0]) if "KInner" in gasd else 0)
    nc = geom.nCells
    rho, rho_bf = foam.read_field(os.path.join(case_dir, "0", "rho"), nc, 1)
    U, U_bf = foam.read_field(os.path.join(case_dir, "0", "U"), nc, 3)
    T, T_bf = foam.read_field(os.path.join(case_dir, "0", "T"), nc, 1)
    rho = rho[:, 0].copy(); T = T[:, 0].copy()
    nbf = geom.nBoundaryFaces
    rho_b = np.zeros(nbf); U_b = np.zeros((nbf, 3)); T_b = np.zeros(nbf)
    patches: List[PatchSpec] = []
    for name, ptype, start, size in zip(geom.patch_names, geom.patch_types, geom.patch_start, geom.patch_size):
        sl = slice(start, start + size)
        own = geom.owner[geom.nInternalFaces + start: geom.nInternalFaces + start + size]
        if ptype == "symmetryPlane":
            kind, ubc, tbc, pr = PATCH_SYMMETRY_PLANE, BC_ZERO_GRADIENT, BC_ZERO_GRADIENT, 0.0
            rho_b[sl] = rho[own]; U_b[sl] = U[own]; T_b[sl] = T[own]
        else:
            re, ue, te = rho_bf[name], U_bf[name], T_bf[name]
            rtype = re["type"][0]
            if rtype not in RHO_BC_TO_KIND:
                raise ValueError(f"patch {name}: unsupported rho boundary type {rtype}")
            kind = RHO_BC_TO_KIND[rtype]
            ubc, tbc = _bc_kind(ue["type"][0]), _bc_kind(te["type"][0])
            pr = 0.0
            if rtype == "pressureIn":
                pr = float(re["pressureIn"][0])
            elif rtype == "pressureOut":
                pr = float(re["pressureOut"][0])
            v = foam.patch_field_values(re, size, 1)
            rho_b[sl] = v[:, 0] if v is not None else rho[own]
            v = foam.patch_field_values(ue, size, 3)
            U_b[sl] = v if (v is not None and ubc == BC_FIXED_VALUE) else U[own]
            v = foam.patch_field_values(te, size, 1)
            T_b[sl] = v[:, 0] if (v is not None and tbc == BC_FIXED_VALUE) else T[own]
        patches.append(PatchSpec(name, kind, start, size, ubc, tbc, pr))
    ctl = foam.read_dict(os.path.join(case_dir, "system", "controlDict"))
    return Case(geom=geom, patches=patches, Xis=Xis, weights=weights,
                xiMax=foam.dimensioned_value(paras["xiMax"]), xiMin=foam.dimensioned_value(paras["xiMin"]),
                gas=gas, rho=rho, U=U, T=T, rho_b=rho_b, U_b=U_b, T_b=T_b,
                deltaT=float(ctl["deltaT"][0]), maxCo=float(ctl.get("maxCo", ["0.8"])[0]),
                adjustTimeStep=ctl.get("adjustTimeStep", ["no"])[0] in ("yes", "on", "true"),
                name=os.path.basename(os.path.normpath(case_dir)), mesh=mesh)


def _uniform_case(mesh: PolyMesh, Xis, weights, kinds: Dict[str, int], *, lid_patch="movingWall",
                  lid_U=(50.0, 0.0, 0.0), wall_T: Optional[Dict[str, float]] = None, gas=None,
                  rho0=RHO0, T0_=T0, U0=(0.0, 0.0, 0.0), name="", bc_overrides: Optional[Dict[str, dict]] = None,
                  perturb: float = 0.0, seed: int = 20260101) -> Case:
    geom = compute_geometry(mesh)
    nc, nbf = geom.nCells, geom.nBoundaryFaces
    rho = np.full(nc, rho0); U = np.tile(np.asarray(U0, dtype=np.float64), (nc, 1)); T = np.full(nc, T0_)
    if perturb > 0:
        rng = np.random.default_rng(seed)
        rho *= 1.0 + perturb * (rng.random(nc) - 0.5) * 2
        T *= 1.0 + perturb * (rng.random(nc) - 0.5) * 2
        U[:, : geom.nSolutionD] += (rng.random((nc, geom.nSolutionD)) - 0.5) * 2 * 5.0
    rho_b = np.full(nbf, rho0); U_b = np.zeros((nbf, 3)); T_b = np.full(nbf, T0_)
    patches = []
    for pname, ptype, start, size in zip(geom.patch_names, geom.patch_types, geom.patch_start, geom.patch_size):
        kind = PATCH_SYMMETRY_PLANE if ptype == "symmetryPlane" else kinds.get(pname, PATCH_MAXWELL_WALL)
        spec = PatchSpec(pname, kind, start, size)
        sl = slice(start, start + size)
        if pname == lid_patch:
            U_b[sl] = np.asarray(lid_U)
        if wall_T and pname in wall_T:
            T_b[sl] = wall_T[pname]
        if kind in (PATCH_DVM_SYMMETRY, PATCH_SYMMETRY_PLANE, PATCH_ZERO_GRADIENT):
            spec.U_bc = spec.T_bc = BC_ZERO_GRADIENT
        if bc_overrides and pname in bc_overrides:
            ov = bc_overrides[pname]
            for k, v in ov.items():
                if k == "U":
                    U_b[sl] = np.asarray(v)
                elif k == "T":
                    T_b[sl] = v
                elif k == "rho":
                    rho_b[sl] = v
                else:
                    setattr(spec, k, v)
        patches.append(spec)
    return Case(geom=geom, patches=patches, Xis=np.asarray(Xis, dtype=np.float64),
                weights=np.asarray(weights, dtype=np.float64), xiMax=float(np.max(Xis)),
                xiMin=float(np.min(Xis)), gas=dict(gas or ARGON), rho=rho, U=U, T=T, rho_b=rho_b,
                U_b=U_b, T_b=T_b, name=name, mesh=mesh)


def gh_set(nDV: int, gas=None, T=T0, stable: bool = False):
    g = gas or ARGON
    return _dvset.dvGH(float(np.sqrt(2.0 * g["R"] * T)), nDV, stable=stable)


def cavity2d_case(n: int, nDV: int = 28, quad: str = "GH", *, distort: float = 0.0, perturb: float = 0.0,
                  xiMax: Optional[float] = None, name: Optional[str] = None) -> Case:
    """2-D lid-driven cavity n x n hexes on [0,1]^2 x [0,0.1], Maxwell walls, lid on top
    (the shape of demo/cavity and of BASELINE config 2)."""
    mesh = hex_block(n, n, 1, (1.0, 1.0, 0.1), two_d=True, distort=distort)
    if quad in ("GH", "GHs"):     # GHs: extended-precision recurrence (sets setDV.py itself cannot produce, nDV > 32)
        Xis, w = gh_set(nDV, stable=quad == "GHs")
    else:
        Xis, w = _dvset.dvNC(xiMax if xiMax is not None else 4.0 * np.sqrt(2 * ARGON["R"] * T0), nDV)
    c = _uniform_case(mesh, Xis, w, {}, name=name or f"cavity2d_{n}x{n}_{quad}{nDV}", perturb=perturb)
    return c


def cavity3d_case(n: int, nDV: int = 28, *, distort: float = 0.0, perturb: float = 0.0,
                  name: Optional[str] = None, length: float = 1.0) -> Case:
    """3-D lid-driven cavity n^3 hexes on [0,length]^3, six Maxwell walls, lid = top (BASELINE config 3).
    With n a power of two (or length = n / 2^k) the point coordinates are exact binary fractions, the
    geometry formulas give exact zeros and the cells are recognised as axis-aligned."""
    mesh = hex_block(n, n, n, (length, length, length), distort=distort)
    Xis, w = gh_set(nDV)
    return _uniform_case(mesh, Xis, w, {}, name=name or f"cavity3d_{n}^3_GH{nDV}", perturb=perturb)


def tri_cavity_case(n: int, nDV: int = 28, *, distort: float = 0.15, wall_T=None, perturb: float = 0.0) -> Case:
    """2-D unstructured triangular-prism cavity with Maxwell walls at different temperatures
    (the shape of BASELINE config 4)."""
    mesh = tri_prism_2d(n, n, (1.0, 1.0, 0.1), distort=distort)
    Xis, w = gh_set(nDV)
    return _uniform_case(mesh, Xis, w, {}, wall_T=wall_T or {"movingWall": 300.0}, name=f"tri_{n}x{n}_GH{nDV}",
                         perturb=perturb)


def poly_cavity_case(n: int, nDV: int = 28, *, jitter: float = 0.25, wall_T=None, perturb: float = 0.0) -> Case:
    """2-D cavity on an unstructured POLYGONAL mesh (Voronoi cells with 4 to 8 sides, the "poly" of BASELINE config 4),
    Maxwell walls at different temperatures, lid on top."""
    mesh = voronoi_prism_2d(n, n, (1.0, 1.0, 0.1), jitter=jitter)
    Xis, w = gh_set(nDV)
    return _uniform_case(mesh, Xis, w, {}, wall_T=wall_T or {"movingWall": 300.0}, name=f"poly_{n}x{n}_GH{nDV}", perturb=perturb)


def ratchet_channel_case(nx: int, ny: int, nDV: int = 28, *, teeth: int = 4, tooth_height: float = 0.3, aspect: float = 4.0,
                         T_ratchet: float = 0.9 * T0, T_top: float = 1.1 * T0, distort: float = 0.1, perturb: float = 0.0) -> Case:
    """2-D micro-channel with a ratchet (saw-tooth) lower wall on an unstructured triangular mesh, every solid wall a
    diffuse Maxwell wall at its own temperature (BASELINE config 4): the ratchet at `T_ratchet`, the flat upper wall at
    `T_top`, the two end walls at the mean - a thermally driven (Knudsen-pump) flow, no moving wall.  The reference case
    closes the channel with cyclic patches, which are out of scope: walls stand in for them.
    2 nx ny triangular prisms: nx = 632, ny = 158 gives 199,712 cells."""
    L = aspect
    mesh = tri_prism_2d(nx, ny, (L, 1.0, 0.1), distort=distort, bottom=ratchet_profile(teeth, tooth_height, L),
                        patch_names={"ymax": "topWall", "ymin": "ratchet", "xmin": "endWalls", "xmax": "endWalls"})
    Xis, w = gh_set(nDV)
    return _uniform_case(mesh, Xis, w, {}, lid_patch="none",
                         wall_T={"ratchet": T_ratchet, "topWall": T_top, "endWalls": 0.5 * (T_ratchet + T_top)},
                         name=f"ratchet_{nx}x{ny}_GH{nDV}", perturb=perturb)


def cylinder_case(ntheta: int, nr: int, nDV: int = 81, *, mach: float = 5.0, quad: str = "NC", xiMax: Optional[float] = None,
                  r_in: float = 0.5, r_out: float = 15.0, T_wall: float = T0, perturb: float = 0.0,
                  blend: float = 3.0) -> Case:
    """2-D hypersonic flow past a circular cylinder (BASELINE config 5): O-type quadrilateral mesh, free stream at
    `mach` along +x (argon, Pr = 2/3, T = 273 K), outer boundary with fixedValue rho / U / T - the reference's way
    of imposing a free stream: DF boundary type "mixed" (doc/usage.tex:85-89, discreteVelocity.C:556-573: the
    incoming half keeps the free-stream Maxwellian of t = 0) - and a diffuse (calculatedMaxwell) cylinder.
    Velocity set: compound Newton-Cotes on [-xiMax, xiMax] wide enough for the shifted Maxwellian (4Z + 1 points,
    setDV.py:153; the config's "80 x 80" is taken as 81 x 81), or stable Gauss-Hermite with quad = "GHs".
    Initial field: the free stream, slowed linearly to rest over `blend` cylinder radii from the surface (0 = the
    uniform free stream; at Ma = 5 that leaves a vacuum on the lee side after one step - wall density 1e-17 of
    the free stream - and the face temperature there is a quotient of round-off, in the reference as here)."""
    mesh = ogrid_cylinder(ntheta, nr, r_in, r_out)
    a = float(np.sqrt(5.0 / 3.0 * ARGON["R"] * T0))
    Uinf = mach * a
    c = float(np.sqrt(2.0 * ARGON["R"] * T0))
    if quad == "NC":
        Xis, w = _dvset.dvNC(xiMax if xiMax is not None else Uinf + 5.0 * c, nDV)
    else:
        Xis, w = gh_set(nDV, stable=True)
    c_ = _uniform_case(mesh, Xis, w, {"farField": PATCH_MIXED, "cylinder": PATCH_MAXWELL_WALL}, lid_patch="none",
                       U0=(Uinf, 0.0, 0.0), wall_T={"cylinder": T_wall}, name=f"cylinder_{ntheta}x{nr}_{quad}{nDV}_Ma{mach:g}",
                       perturb=perturb, bc_overrides={"farField": dict(U=(Uinf, 0.0, 0.0))})
    if blend > 0:
        r = np.hypot(c_.geom.C[:, 0], c_.geom.C[:, 1])
        c_.U *= np.clip((r - r_in) / (blend * r_in), 0.0, 1.0)[:, None]
    return c_
