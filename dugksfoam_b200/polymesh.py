"""polyMesh container, OpenFOAM finite-volume geometry and least-squares vectors.

Host-side pre-processing for the standalone harness (tests, bench.py): inside
OpenFOAM the adapter takes ``mesh.C()/V()/Cf()/Sf()`` and
``leastSquaresVectors`` straight from the library; here the same
quantities are rebuilt from ``points/faces/owner/neighbour`` with the formulas of
the OpenFOAM Foundation releases the reference supports (README.md:19-26).

[OF-lib] = behaviour of OpenFOAM code that is not in the reference tree
(SURVEY.md Appendix C items 1-8); version notes are kept next to each formula.
Reference call sites: discreteVelocity.C:472-476,942-943 (C, Cf, Sf, V),
:420-421 (fvc::grad -> leastSquares), zeroBoundaryGrad/zeroBoundaryVectors.C
:90-230 (in-tree twin of the stock least-squares vectors).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

SMALL = 1.0e-15       # [OF-lib] double-precision build
VSMALL = 1.0e-300
ROOTVSMALL = 1.0e-150


@dataclass
class Patch:
    name: str
    type: str          # polyMesh/boundary type: wall, patch, empty, symmetryPlane ...
    nFaces: int
    startFace: int


@dataclass
class PolyMesh:
    """constant/polyMesh: all faces, including those of ``empty`` patches."""

    points: np.ndarray                 # [nPoints, 3]
    face_verts: np.ndarray             # flat vertex list
    face_offsets: np.ndarray           # [nFaces + 1]
    owner: np.ndarray                  # [nFaces] int32
    neighbour: np.ndarray              # [nInternalFaces] int32
    patches: List[Patch]
    nCells: int = 0
    geom: Optional["Geometry"] = field(default=None, repr=False)

    def __post_init__(self):
        self.owner = np.ascontiguousarray(self.owner, dtype=np.int32)
        self.neighbour = np.ascontiguousarray(self.neighbour, dtype=np.int32)
        self.face_offsets = np.ascontiguousarray(self.face_offsets, dtype=np.int64)
        self.face_verts = np.ascontiguousarray(self.face_verts, dtype=np.int64)
        self.points = np.ascontiguousarray(self.points, dtype=np.float64)
        if not self.nCells:
            self.nCells = int(self.owner.max()) + 1

    @property
    def nFaces(self) -> int:
        return len(self.owner)

    @property
    def nInternalFaces(self) -> int:
        return len(self.neighbour)

    def geometry(self) -> "Geometry":
        if self.geom is None:
            self.geom = compute_geometry(self)
        return self.geom


@dataclass
class Geometry:
    """Everything dugks_mesh_t needs, on the faces that take part in the solver
    (internal faces + faces of non-empty patches, in that order)."""

    nCells: int
    nInternalFaces: int
    nBoundaryFaces: int
    nSolutionD: int
    owner: np.ndarray        # [nif + nbf]
    neighbour: np.ndarray    # [nif]
    C: np.ndarray            # [nc, 3]
    V: np.ndarray            # [nc]
    Cf: np.ndarray           # [nf, 3]
    Sf: np.ndarray           # [nf, 3]
    ownLs: np.ndarray        # [nif, 3]
    neiLs: np.ndarray        # [nif, 3]
    patchLs: np.ndarray      # [nbf, 3]
    deltaCoeffs: np.ndarray  # [nf]
    weights: np.ndarray      # [nif] linear interpolation weights
    patch_names: List[str]
    patch_types: List[str]
    patch_start: List[int]   # in boundary-face numbering
    patch_size: List[int]
    empty_dirs: np.ndarray   # bool[3]

    @property
    def nFaces(self) -> int:
        return self.nInternalFaces + self.nBoundaryFaces


# ---------------------------------------------------------------------------
# primitiveMesh face / cell geometry  [OF-lib] primitiveMeshFaceCentresAndAreas.C,
# primitiveMeshCellCentresAndVols.C (Foundation 2.3 - 6)
# ---------------------------------------------------------------------------

def face_centres_and_areas(mesh: PolyMesh):
    """Triangle fan about the vertex average."""
    nF = mesh.nFaces
    offs = mesh.face_offsets
    nv = np.diff(offs)
    Cf = np.zeros((nF, 3))
    Sf = np.zeros((nF, 3))
    pts = mesh.points
    for n in np.unique(nv):
        idx = np.nonzero(nv == n)[0]
        verts = mesh.face_verts[offs[idx][:, None] + np.arange(n)[None, :]]  # [m, n]
        P = pts[verts]                                                       # [m, n, 3]
        if n == 3:
            Cf[idx] = (1.0 / 3.0) * (P[:, 0] + P[:, 1] + P[:, 2])
            Sf[idx] = 0.5 * np.cross(P[:, 1] - P[:, 0], P[:, 2] - P[:, 0])
            continue
        fCentre = P[:, 0].copy()
        for i in range(1, n):
            fCentre += P[:, i]
        fCentre /= n
        sumN = np.zeros((len(idx), 3))
        sumA = np.zeros(len(idx))
        sumAc = np.zeros((len(idx), 3))
        for i in range(n):
            p = P[:, i]
            pn = P[:, (i + 1) % n]
            c = p + pn + fCentre
            nrm = np.cross(pn - p, fCentre - p)
            a = np.sqrt((nrm * nrm).sum(axis=1))
            sumN += nrm
            sumA += a
            sumAc += a[:, None] * c
        ok = sumA > ROOTVSMALL
        cf = fCentre.copy()
        cf[ok] = (1.0 / 3.0) * sumAc[ok] / sumA[ok][:, None]
        sf = np.zeros_like(sumN)
        sf[ok] = 0.5 * sumN[ok]
        Cf[idx] = cf
        Sf[idx] = sf
    return Cf, Sf


def cell_centres_and_volumes(mesh: PolyMesh, Cf: np.ndarray, Sf: np.ndarray):
    """Pyramid decomposition about the face-centre average.

    Version note: Foundation >= 2.3 clamps pyr3Vol with max(., VSMALL); on valid
    meshes every pyramid volume is positive, so the clamp is inactive.
    """
    nC, nif = mesh.nCells, mesh.nInternalFaces
    own, nei = mesh.owner, mesh.neighbour
    cEst = np.zeros((nC, 3))
    nCellFaces = np.zeros(nC)
    for d in range(3):
        cEst[:, d] = np.bincount(own, weights=Cf[:, d], minlength=nC)
        cEst[:, d] += np.bincount(nei, weights=Cf[:nif, d], minlength=nC)
    nCellFaces = np.bincount(own, minlength=nC) + np.bincount(nei, minlength=nC)
    cEst /= nCellFaces[:, None]

    ctr = np.zeros((nC, 3))
    vol = np.zeros(nC)
    # owner side
    pyr = np.einsum("ij,ij->i", Sf, Cf - cEst[own])
    pyr = np.maximum(pyr, VSMALL)
    pc = 0.75 * Cf + 0.25 * cEst[own]
    for d in range(3):
        ctr[:, d] += np.bincount(own, weights=pyr * pc[:, d], minlength=nC)
    vol += np.bincount(own, weights=pyr, minlength=nC)
    # neighbour side
    pyrn = np.einsum("ij,ij->i", Sf[:nif], cEst[nei] - Cf[:nif])
    pyrn = np.maximum(pyrn, VSMALL)
    pcn = 0.75 * Cf[:nif] + 0.25 * cEst[nei]
    for d in range(3):
        ctr[:, d] += np.bincount(nei, weights=pyrn * pcn[:, d], minlength=nC)
    vol += np.bincount(nei, weights=pyrn, minlength=nC)
    ctr /= vol[:, None]
    vol *= 1.0 / 3.0
    return ctr, vol


# ---------------------------------------------------------------------------
# symmTensor helpers (xx, xy, xz, yy, yz, zz)
# ---------------------------------------------------------------------------

def _sqr(d):
    return np.stack([d[:, 0] * d[:, 0], d[:, 0] * d[:, 1], d[:, 0] * d[:, 2],
                     d[:, 1] * d[:, 1], d[:, 1] * d[:, 2], d[:, 2] * d[:, 2]], axis=1)


def _symm_dot(t, v):
    return np.stack([t[:, 0] * v[:, 0] + t[:, 1] * v[:, 1] + t[:, 2] * v[:, 2],
                     t[:, 1] * v[:, 0] + t[:, 3] * v[:, 1] + t[:, 4] * v[:, 2],
                     t[:, 2] * v[:, 0] + t[:, 4] * v[:, 1] + t[:, 5] * v[:, 2]], axis=1)


def _symm_inv_raw(t):
    """[OF-lib] SymmTensorI.H inv(): cofactors / det."""
    xx, xy, xz, yy, yz, zz = (t[:, i] for i in range(6))
    det = (xx * yy * zz + xy * yz * xz + xz * xy * yz
           - xx * yz * yz - xy * xy * zz - xz * yy * xz)
    inv = np.stack([yy * zz - yz * yz, xz * yz - xy * zz, xy * yz - xz * yy,
                    xx * zz - xz * xz, xy * xz - xx * yz, xx * yy - xy * xy], axis=1)
    return inv / det[:, None]


def symm_inv_field(t):
    """[OF-lib] symmTensorField.C inv(tmp<symmTensorField>): the empty directions are
    detected from ELEMENT 0 only (scale = magSqr(tf[0]); a diagonal entry with
    entry/scale < SMALL is regularised by adding 1 before and removing it after the
    inversion) — SURVEY App. C item 7; used at zeroBoundaryVectors.C:171."""
    if len(t) == 0:
        return t.copy()
    t0 = t[0]
    scale = (t0[0] ** 2 + 2 * t0[1] ** 2 + 2 * t0[2] ** 2 + t0[3] ** 2 + 2 * t0[4] ** 2 + t0[5] ** 2)
    remove = [t0[0] / scale < SMALL, t0[3] / scale < SMALL, t0[5] / scale < SMALL]
    if any(remove):
        adj = np.zeros(6)
        for flag, i in zip(remove, (0, 3, 5)):
            if flag:
                adj[i] = 1.0
        return _symm_inv_raw(t + adj[None, :]) - adj[None, :]
    return _symm_inv_raw(t)


# ---------------------------------------------------------------------------
# full geometry
# ---------------------------------------------------------------------------

def compute_geometry(mesh: PolyMesh, patch_delta: str = "normal") -> Geometry:
    """patch_delta: "normal" = Foundation >= 2.3 fvPatch::delta() = n (n . (Cf - C))
    (the versions the reference lists); "full" = <= 2.2 (Cf - C).  Identical on
    orthogonal meshes."""
    CfAll, SfAll = face_centres_and_areas(mesh)
    C, V = cell_centres_and_volumes(mesh, CfAll, SfAll)
    nif = mesh.nInternalFaces

    keep = [np.arange(nif)]
    names, types, starts, sizes = [], [], [], []
    empty_dirs = np.zeros(3, dtype=bool)
    nb = 0
    for p in mesh.patches:
        sl = np.arange(p.startFace, p.startFace + p.nFaces)
        if p.type == "empty":
            # [OF-lib] polyMesh::calcDirections: a direction is empty when the empty
            # patches' normals point along it
            if p.nFaces:
                nrm = np.abs(SfAll[sl]).sum(axis=0)
                empty_dirs |= nrm > 1e-6 * nrm.max()
            continue
        keep.append(sl)
        names.append(p.name)
        types.append(p.type)
        starts.append(nb)
        sizes.append(p.nFaces)
        nb += p.nFaces
    keep = np.concatenate(keep)
    owner = mesh.owner[keep].astype(np.int32)
    nei = mesh.neighbour
    Cf = np.ascontiguousarray(CfAll[keep])
    Sf = np.ascontiguousarray(SfAll[keep])
    nf = len(keep)
    nbf = nf - nif
    magSf = np.sqrt((Sf * Sf).sum(axis=1))
    own_i = owner[:nif]

    # surfaceInterpolation::makeWeights [OF-lib] (App. C item 3)
    SfdOwn = np.abs(np.einsum("ij,ij->i", Sf[:nif], Cf[:nif] - C[own_i]))
    SfdNei = np.abs(np.einsum("ij,ij->i", Sf[:nif], C[nei] - Cf[:nif]))
    w = SfdNei / (SfdOwn + SfdNei)

    # deltas
    d = C[nei] - C[own_i]
    own_b = owner[nif:]
    nHat = Sf[nif:] / magSf[nif:, None]
    dfull = Cf[nif:] - C[own_b]
    if patch_delta == "normal":
        pd = nHat * np.einsum("ij,ij->i", nHat, dfull)[:, None]
    else:
        pd = dfull
    deltaCoeffs = np.empty(nf)
    deltaCoeffs[:nif] = 1.0 / np.sqrt((d * d).sum(axis=1))          # App. C item 4
    deltaCoeffs[nif:] = 1.0 / np.sqrt((pd * pd).sum(axis=1))

    # leastSquaresVectors::calcLeastSquaresVectors [OF-lib]; in-tree twin
    # zeroBoundaryVectors.C:113-123 (internal) + the stock boundary lines kept in
    # comments at :159-165
    nC = mesh.nCells
    magSqrD = (d * d).sum(axis=1)
    wdd = (magSf[:nif] / magSqrD)[:, None] * _sqr(d)
    dd = np.zeros((nC, 6))
    for i in range(6):
        dd[:, i] += np.bincount(own_i, weights=(1 - w) * wdd[:, i], minlength=nC)
        dd[:, i] += np.bincount(nei, weights=w * wdd[:, i], minlength=nC)
    magSqrPd = (pd * pd).sum(axis=1)
    bdd = (magSf[nif:] / magSqrPd)[:, None] * _sqr(pd)
    for i in range(6):
        dd[:, i] += np.bincount(own_b, weights=bdd[:, i], minlength=nC)
    invDd = symm_inv_field(dd)                                       # :171
    ownLs = ((1 - w) * magSf[:nif] / magSqrD)[:, None] * _symm_dot(invDd[own_i], d)   # :182
    neiLs = (-w * magSf[:nif] / magSqrD)[:, None] * _symm_dot(invDd[nei], d)          # :183
    patchLs = (magSf[nif:] * (1.0 / magSqrPd))[:, None] * _symm_dot(invDd[own_b], pd)  # :213-220

    nSolutionD = 3 - int(empty_dirs.sum())
    return Geometry(
        nCells=nC, nInternalFaces=nif, nBoundaryFaces=nbf, nSolutionD=nSolutionD,
        owner=np.ascontiguousarray(owner), neighbour=np.ascontiguousarray(nei, dtype=np.int32),
        C=np.ascontiguousarray(C), V=np.ascontiguousarray(V), Cf=Cf, Sf=Sf,
        ownLs=np.ascontiguousarray(ownLs), neiLs=np.ascontiguousarray(neiLs),
        patchLs=np.ascontiguousarray(patchLs), deltaCoeffs=deltaCoeffs, weights=w,
        patch_names=names, patch_types=types, patch_start=starts, patch_size=sizes,
        empty_dirs=empty_dirs)


# ---------------------------------------------------------------------------
# synthetic meshes in OpenFOAM (blockMesh / upper-triangular) ordering
# ---------------------------------------------------------------------------

def hex_block(nx: int, ny: int, nz: int, lengths=(1.0, 1.0, 1.0), *, two_d: bool = False,
              patch_names: Optional[Dict[str, str]] = None,
              distort: float = 0.0, seed: int = 20260101) -> PolyMesh:
    """nx*ny*nz hexahedra on [0,Lx]x[0,Ly]x[0,Lz]; cell id = i + nx*(j + ny*k).

    Internal faces are in OpenFOAM's upper-triangular order (by owner, then by
    neighbour).  Boundary faces are grouped into patches; ``patch_names`` maps the
    six sides ``xmin,xmax,ymin,ymax,zmin,zmax`` to patch names (sides sharing a name
    are merged in the order given).  ``two_d`` marks zmin/zmax as one ``empty``
    patch ``frontAndBack`` (nz must be 1).  ``distort`` moves interior points
    randomly by that fraction of the cell size (non-orthogonal test meshes).
    """
    Lx, Ly, Lz = lengths
    px, py, pz = nx + 1, ny + 1, nz + 1
    xs = np.linspace(0.0, Lx, px)
    ys = np.linspace(0.0, Ly, py)
    zs = np.linspace(0.0, Lz, pz)
    Z, Y, X = np.meshgrid(zs, ys, xs, indexing="ij")
    points = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    if distort > 0.0:
        rng = np.random.default_rng(seed)
        K, J, I = np.meshgrid(np.arange(pz), np.arange(py), np.arange(px), indexing="ij")
        interior = ((I > 0) & (I < nx) & (J > 0) & (J < ny)).ravel()
        h = np.array([Lx / nx, Ly / ny, Lz / nz])
        disp = (rng.random((len(points), 3)) - 0.5) * 2.0 * distort * h[None, :]
        if two_d or nz == 1:
            disp[:, 2] = 0.0
            # keep the extrusion straight: same displacement on both z-planes
            disp = disp.reshape(pz, py * px, 3)
            disp[:] = disp[0][None]
            disp = disp.reshape(-1, 3)
        else:
            interior &= ((K > 0) & (K < nz)).ravel()
        points[interior] += disp[interior]

    def pid(i, j, k):
        return i + px * (j + py * k)

    def cid(i, j, k):
        return i + nx * (j + ny * k)

    kk, jj, ii = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    ii, jj, kk = ii.ravel(), jj.ravel(), kk.ravel()
    cells = cid(ii, jj, kk)

    def xface(i, j, k):  # plane x = i, normal +x
        return np.stack([pid(i, j, k), pid(i, j + 1, k), pid(i, j + 1, k + 1), pid(i, j, k + 1)], axis=1)

    def yface(i, j, k):  # plane y = j, normal +y
        return np.stack([pid(i, j, k), pid(i, j, k + 1), pid(i + 1, j, k + 1), pid(i + 1, j, k)], axis=1)

    def zface(i, j, k):  # plane z = k, normal +z
        return np.stack([pid(i, j, k), pid(i + 1, j, k), pid(i + 1, j + 1, k), pid(i, j + 1, k)], axis=1)

    # internal faces: per cell x+, y+, z+ (ascending neighbour id), sorted by owner
    fv, fo, fn, order = [], [], [], []
    m = ii < nx - 1
    fv.append(xface(ii[m] + 1, jj[m], kk[m])); fo.append(cells[m]); fn.append(cells[m] + 1); order.append(cells[m] * 3 + 0)
    m = jj < ny - 1
    fv.append(yface(ii[m], jj[m] + 1, kk[m])); fo.append(cells[m]); fn.append(cells[m] + nx); order.append(cells[m] * 3 + 1)
    m = kk < nz - 1
    fv.append(zface(ii[m], jj[m], kk[m] + 1)); fo.append(cells[m]); fn.append(cells[m] + nx * ny); order.append(cells[m] * 3 + 2)
    fv = np.concatenate(fv); fo = np.concatenate(fo); fn = np.concatenate(fn)
    perm = np.argsort(np.concatenate(order), kind="stable")
    fv, fo, fn = fv[perm], fo[perm], fn[perm]

    sides = {}
    m = ii == 0
    sides["xmin"] = (xface(ii[m], jj[m], kk[m])[:, ::-1], cells[m])
    m = ii == nx - 1
    sides["xmax"] = (xface(ii[m] + 1, jj[m], kk[m]), cells[m])
    m = jj == 0
    sides["ymin"] = (yface(ii[m], jj[m], kk[m])[:, ::-1], cells[m])
    m = jj == ny - 1
    sides["ymax"] = (yface(ii[m], jj[m] + 1, kk[m]), cells[m])
    m = kk == 0
    sides["zmin"] = (zface(ii[m], jj[m], kk[m])[:, ::-1], cells[m])
    m = kk == nz - 1
    sides["zmax"] = (zface(ii[m], jj[m], kk[m] + 1), cells[m])

    if patch_names is None:
        patch_names = {"ymax": "movingWall", "xmin": "fixedWalls", "xmax": "fixedWalls",
                       "ymin": "fixedWalls"}
        if not two_d:
            patch_names.update({"zmin": "fixedWalls", "zmax": "fixedWalls"})
    if two_d:
        assert nz == 1
        patch_names = dict(patch_names)
        patch_names.pop("zmin", None)
        patch_names.pop("zmax", None)
    grouped: Dict[str, list] = {}
    for side, name in patch_names.items():
        grouped.setdefault(name, []).append(side)
    patches = []
    all_fv, all_fo = [fv], [fo]
    start = len(fo)
    for name, sds in grouped.items():
        n = 0
        for s in sds:
            all_fv.append(sides[s][0]); all_fo.append(sides[s][1]); n += len(sides[s][1])
        ptype = "symmetryPlane" if name.startswith("symmetryPlane") else "wall"
        patches.append(Patch(name, ptype, n, start))
        start += n
    if two_d:
        n = 0
        for s in ("zmin", "zmax"):
            all_fv.append(sides[s][0]); all_fo.append(sides[s][1]); n += len(sides[s][1])
        patches.append(Patch("frontAndBack", "empty", n, start))
    fv_all = np.concatenate(all_fv)
    owner = np.concatenate(all_fo)
    offsets = np.arange(len(owner) + 1, dtype=np.int64) * 4
    return PolyMesh(points=points, face_verts=fv_all.ravel(), face_offsets=offsets,
                    owner=owner, neighbour=fn, patches=patches, nCells=nx * ny * nz)


def ogrid_cylinder(ntheta: int, nr: int, r_in: float = 0.5, r_out: float = 15.0, *, growth: Optional[float] = None,
                   thickness: float = 0.1) -> PolyMesh:
    """2-D O-type quadrilateral mesh around a circular cylinder (BASELINE config 5), extruded one cell in z:
    ntheta cells around, nr cells from the cylinder (radius r_in, patch ``cylinder``) to the far boundary (radius
    r_out, patch ``farField``), radial spacing in geometric progression (``growth`` = ratio of successive cell
    heights; default: the ratio that makes the first cell square).  The ring is closed: the faces between the last
    and the first sector are ordinary internal faces.  Cell id = i + ntheta * j with i running CLOCKWISE (so that
    (i, j, z) is right-handed like hex_block's (x, y, z)) and j outwards; internal faces in OpenFOAM's
    upper-triangular order; front and back are one ``empty`` patch."""
    if ntheta < 3 or nr < 1:
        raise ValueError("ogrid_cylinder: need ntheta >= 3 and nr >= 1")
    if growth is None:
        # first cell about square: h0 = r_in * 2 pi / ntheta; solve h0 (g^nr - 1) / (g - 1) = r_out - r_in for g
        h0 = r_in * 2.0 * np.pi / ntheta
        lo, hi = 1.0 + 1e-12, 4.0
        for _ in range(200):
            g = 0.5 * (lo + hi)
            if h0 * (g ** nr - 1.0) / (g - 1.0) > r_out - r_in:
                hi = g
            else:
                lo = g
        growth = 0.5 * (lo + hi) if h0 * nr < r_out - r_in else 1.0
    if abs(growth - 1.0) < 1e-12:
        radii = np.linspace(r_in, r_out, nr + 1)
    else:
        hs = growth ** np.arange(nr)
        radii = r_in + (r_out - r_in) * np.concatenate([[0.0], np.cumsum(hs)]) / hs.sum()
    theta = -2.0 * np.pi * np.arange(ntheta) / ntheta
    pr = nr + 1
    J, I = np.meshgrid(np.arange(pr), np.arange(ntheta), indexing="ij")
    xy = np.stack([radii[J] * np.cos(theta[I]), radii[J] * np.sin(theta[I])], axis=-1).reshape(-1, 2)
    points = np.concatenate([np.c_[xy, np.zeros(len(xy))], np.c_[xy, np.full(len(xy), thickness)]])

    def pid(i, j, k):
        return (i % ntheta) + ntheta * (j + pr * k)

    jj, ii = np.meshgrid(np.arange(nr), np.arange(ntheta), indexing="ij")
    ii, jj = ii.ravel(), jj.ravel()
    cells = ii + ntheta * jj

    def iface(i, j):   # between sectors i - 1 and i, normal along +i
        return np.stack([pid(i, j, 0), pid(i, j + 1, 0), pid(i, j + 1, 1), pid(i, j, 1)], axis=1)

    def jface(i, j):   # radius index j, normal outwards
        return np.stack([pid(i, j, 0), pid(i, j, 1), pid(i + 1, j, 1), pid(i + 1, j, 0)], axis=1)

    def kface(i, j, k):  # normal +z
        return np.stack([pid(i, j, k), pid(i + 1, j, k), pid(i + 1, j + 1, k), pid(i, j + 1, k)], axis=1)

    # internal faces: i+ face of every cell (the last sector's closes the ring: its neighbour has the smaller id, so
    # that cell owns the face and the vertex order is reversed), j+ face below the outer ring
    nxt = ((ii + 1) % ntheta) + ntheta * jj
    wrap = nxt < cells
    fv_i = iface(ii + 1, jj)
    fv_i[wrap] = fv_i[wrap][:, ::-1]
    fo_i = np.where(wrap, nxt, cells)
    fn_i = np.where(wrap, cells, nxt)
    m = jj < nr - 1
    fv = np.concatenate([fv_i, jface(ii[m], jj[m] + 1)])
    fo = np.concatenate([fo_i, cells[m]])
    fn = np.concatenate([fn_i, cells[m] + ntheta])
    perm = np.lexsort((fn, fo))
    fv, fo, fn = fv[perm], fo[perm], fn[perm]
    inner, outer = jj == 0, jj == nr - 1
    bnd = [("cylinder", "wall", jface(ii[inner], jj[inner])[:, ::-1], cells[inner]),
           ("farField", "patch", jface(ii[outer], jj[outer] + 1), cells[outer]),
           ("frontAndBack", "empty", np.concatenate([kface(ii, jj, 0)[:, ::-1], kface(ii, jj, 1)]), np.concatenate([cells, cells]))]
    patches, all_fv, all_fo, start = [], [fv], [fo], len(fo)
    for name, ptype, v, o in bnd:
        patches.append(Patch(name, ptype, len(o), start))
        all_fv.append(v); all_fo.append(o); start += len(o)
    owner = np.concatenate(all_fo)
    return PolyMesh(points=points, face_verts=np.concatenate(all_fv).ravel(),
                    face_offsets=np.arange(len(owner) + 1, dtype=np.int64) * 4, owner=owner, neighbour=fn,
                    patches=patches, nCells=ntheta * nr)


def ratchet_profile(teeth: int, height: float, length: float, rise: float = 0.8):
    """Saw-tooth (ratchet) wall y_b(x): `teeth` asymmetric teeth over [0, length], each rising linearly to
    `height` over the fraction `rise` of its period and falling back over the rest."""
    period = length / teeth

    def yb(x):
        s = np.mod(x, period) / period
        s = np.where(np.isclose(x, length), 0.0, s)
        return height * np.where(s <= rise, s / rise, (1.0 - s) / (1.0 - rise))
    return yb


def tri_prism_2d(nx: int, ny: int, lengths=(1.0, 1.0, 0.1), *, distort: float = 0.0,
                 seed: int = 20260101, patch_names: Optional[Dict[str, str]] = None,
                 bottom=None) -> PolyMesh:
    """2-D unstructured triangular mesh (each quad of an nx*ny grid split along
    alternating diagonals, optionally distorted), extruded one cell in z with an
    ``empty`` frontAndBack patch — the shape of BASELINE config 4 (tri mesh,
    Maxwell walls).  Cells are triangular prisms: 3 quad side faces + 2 triangles.
    `bottom(x)`: profile of the lower wall (ratchet_profile); the columns of the mesh are
    compressed between it and the flat upper wall."""
    Lx, Ly, Lz = lengths
    px, py = nx + 1, ny + 1
    xs = np.linspace(0, Lx, px); ys = np.linspace(0, Ly, py)
    Y, X = np.meshgrid(ys, xs, indexing="ij")
    xy = np.stack([X.ravel(), Y.ravel()], axis=1)
    if distort > 0:
        rng = np.random.default_rng(seed)
        J, I = np.meshgrid(np.arange(py), np.arange(px), indexing="ij")
        interior = ((I > 0) & (I < nx) & (J > 0) & (J < ny)).ravel()
        h = np.array([Lx / nx, Ly / ny])
        xy[interior] += ((rng.random((len(xy), 2)) - 0.5) * 2 * distort * h)[interior]
    npl = len(xy)
    xyp = xy.copy()    # patches are told apart on the unmapped rectangle
    if bottom is not None:
        yb = np.asarray(bottom(xyp[:, 0]), dtype=np.float64)
        xyp[:, 1] = yb + xy[:, 1] * (Ly - yb) / Ly
    points = np.concatenate([np.column_stack([xyp, np.zeros(npl)]), np.column_stack([xyp, np.full(npl, Lz)])])

    def p2(i, j):
        return i + px * j

    tris = []  # counter-clockwise vertex triples
    for j in range(ny):
        for i in range(nx):
            a, b, c, d = p2(i, j), p2(i + 1, j), p2(i + 1, j + 1), p2(i, j + 1)
            if (i + j) % 2 == 0:
                tris.append((a, b, c)); tris.append((a, c, d))
            else:
                tris.append((a, b, d)); tris.append((b, c, d))
    tris = np.array(tris, dtype=np.int64)
    nC = len(tris)
    # edges -> cells
    edge_map: Dict[tuple, list] = {}
    for c, t in enumerate(tris):
        for e in range(3):
            u, v = int(t[e]), int(t[(e + 1) % 3])
            edge_map.setdefault((min(u, v), max(u, v)), []).append((c, u, v))
    internal, boundary = [], {}
    if patch_names is None:
        patch_names = {"ymax": "movingWall", "xmin": "fixedWalls", "xmax": "fixedWalls", "ymin": "fixedWalls"}
    tol = 1e-12
    for key, lst in edge_map.items():
        if len(lst) == 2:
            (c0, u, v), (c1, _, _) = sorted(lst)
            internal.append((c0, c1, u, v))       # u->v is CCW in c0: outward normal of c0
        else:
            c0, u, v = lst[0]
            mx, my = 0.5 * (xy[u] + xy[v])
            if abs(my - Ly) < tol: side = "ymax"
            elif abs(my) < tol: side = "ymin"
            elif abs(mx) < tol: side = "xmin"
            else: side = "xmax"
            boundary.setdefault(patch_names[side], []).append((c0, u, v))
    internal.sort(key=lambda t: (t[0], t[1]))

    def side_face(u, v):  # quad with outward normal for CCW edge u->v: (u, v, v+npl, u+npl)
        return [u, v, v + npl, u + npl]

    verts, offs, owner, neigh = [], [0], [], []
    for c0, c1, u, v in internal:
        verts += side_face(u, v); offs.append(len(verts)); owner.append(c0); neigh.append(c1)
    patches = []
    for name in dict.fromkeys(patch_names.values()):
        lst = sorted(boundary.get(name, []))
        patches.append(Patch(name, "wall", len(lst), len(owner)))
        for c0, u, v in lst:
            verts += side_face(u, v); offs.append(len(verts)); owner.append(c0)
    start = len(owner)
    for c, t in enumerate(tris):   # back (z=0): outward -z -> clockwise
        verts += [int(t[0]), int(t[2]), int(t[1])]; offs.append(len(verts)); owner.append(c)
    for c, t in enumerate(tris):   # front (z=Lz): outward +z
        verts += [int(t[0]) + npl, int(t[1]) + npl, int(t[2]) + npl]; offs.append(len(verts)); owner.append(c)
    patches.append(Patch("frontAndBack", "empty", 2 * nC, start))
    return PolyMesh(points=points, face_verts=np.array(verts), face_offsets=np.array(offs),
                    owner=np.array(owner), neighbour=np.array(neigh), patches=patches, nCells=nC)


def voronoi_prism_2d(nx: int, ny: int, lengths=(1.0, 1.0, 0.1), *, jitter: float = 0.25, seed: int = 20260101,
                     patch_names: Optional[Dict[str, str]] = None) -> PolyMesh:
    """2-D unstructured POLYGONAL mesh (the "poly" of BASELINE config 4): the Voronoi diagram of nx*ny jittered seed
    points in a rectangle, extruded one cell in z with an ``empty`` frontAndBack patch.  The seeds are mirrored in the
    four walls, so the cells of the original seeds are cut exactly by the rectangle.  Cells are prisms over polygons
    with 4 to about 8 sides (needs scipy)."""
    from scipy.spatial import Voronoi
    Lx, Ly, Lz = lengths
    rng = np.random.default_rng(seed)
    hx, hy = Lx / nx, Ly / ny
    I, J = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    pts = np.column_stack([(I.ravel() + 0.5) * hx, (J.ravel() + 0.5) * hy])
    pts += (rng.random(pts.shape) - 0.5) * 2 * jitter * np.array([hx, hy])
    n0 = len(pts)
    mirrored = [pts, pts * [-1, 1], pts * [-1, 1] + [2 * Lx, 0], pts * [1, -1], pts * [1, -1] + [0, 2 * Ly]]
    vor = Voronoi(np.concatenate(mirrored))
    # vertices used by the cells of the original seeds, renumbered; coordinates snapped onto the walls
    used = sorted({v for i in range(n0) for v in vor.regions[vor.point_region[i]]})
    assert -1 not in used, "open Voronoi region (seeds too close to a wall?)"
    vid = {v: k for k, v in enumerate(used)}
    xy = vor.vertices[used].copy()
    tol = 1e-9 * max(Lx, Ly)
    for d, L in ((0, Lx), (1, Ly)):
        xy[np.abs(xy[:, d]) < tol, d] = 0.0
        xy[np.abs(xy[:, d] - L) < tol, d] = L
    npl = len(xy)
    points = np.concatenate([np.column_stack([xy, np.zeros(npl)]), np.column_stack([xy, np.full(npl, Lz)])])
    # polygons, counter-clockwise about their seed
    polys = []
    for i in range(n0):
        vs = [vid[v] for v in vor.regions[vor.point_region[i]]]
        ang = np.arctan2(xy[vs, 1] - pts[i, 1], xy[vs, 0] - pts[i, 0])
        polys.append([vs[k] for k in np.argsort(ang)])
    if patch_names is None:
        patch_names = {"ymax": "movingWall", "xmin": "fixedWalls", "xmax": "fixedWalls", "ymin": "fixedWalls"}
    internal, boundary = [], {}
    for (p, q), rv in zip(vor.ridge_points, vor.ridge_vertices):
        if -1 in rv or (p >= n0 and q >= n0) or rv[0] not in vid or rv[1] not in vid:
            continue
        a, b = vid[rv[0]], vid[rv[1]]
        if np.hypot(*(xy[a] - xy[b])) < tol:
            continue                                            # degenerate ridge (four seeds on a circle)
        c0 = int(min(p, q)) if max(p, q) < n0 else int(p if p < n0 else q)
        # orient the edge counter-clockwise as seen from c0: its outward normal then points away from c0
        e, m = xy[b] - xy[a], 0.5 * (xy[a] + xy[b]) - pts[c0]
        if e[0] * m[1] - e[1] * m[0] > 0:                       # seed on the left of a->b means clockwise about the seed: flip
            a, b = b, a
        if max(p, q) < n0:
            internal.append((c0, int(max(p, q)), a, b))
        else:
            mx, my = 0.5 * (xy[a] + xy[b])
            side = "ymax" if abs(my - Ly) < tol else ("ymin" if abs(my) < tol else ("xmin" if abs(mx) < tol else "xmax"))
            boundary.setdefault(patch_names[side], []).append((c0, a, b))
    internal.sort(key=lambda t: (t[0], t[1]))

    def side_face(u, v):                                        # outward for the counter-clockwise edge u -> v
        return [u, v, v + npl, u + npl]

    verts, offs, owner, neigh = [], [0], [], []
    for c0, c1, u, v in internal:
        verts += side_face(u, v); offs.append(len(verts)); owner.append(c0); neigh.append(c1)
    patches = []
    for name in dict.fromkeys(patch_names.values()):
        lst = sorted(boundary.get(name, []))
        patches.append(Patch(name, "wall", len(lst), len(owner)))
        for c0, u, v in lst:
            verts += side_face(u, v); offs.append(len(verts)); owner.append(c0)
    start = len(owner)
    for c, poly in enumerate(polys):                            # back (z = 0): outward -z -> clockwise
        verts += poly[::-1]; offs.append(len(verts)); owner.append(c)
    for c, poly in enumerate(polys):                            # front (z = Lz)
        verts += [v + npl for v in poly]; offs.append(len(verts)); owner.append(c)
    patches.append(Patch("frontAndBack", "empty", 2 * n0, start))
    return PolyMesh(points=points, face_verts=np.array(verts), face_offsets=np.array(offs),
                    owner=np.array(owner), neighbour=np.array(neigh), patches=patches, nCells=n0)
