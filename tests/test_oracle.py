"""CPU tests (no GPU): the oracle pinned against the reference's golden vector and the
physical / numerical identities that stand in for the golden outputs the reference
does not ship ("parity unpinned")."""
import json
import os

import numpy as np
import pytest

import parity_util as util
from dugksfoam_b200 import case as cs
from dugksfoam_b200 import dvset
from dugksfoam_b200.polymesh import compute_geometry, hex_block, tri_prism_2d, voronoi_prism_2d

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "demo_cavity.json")))
REF_DEMO = "/root/reference/demo/cavity"


# ---- golden vector: the shipped 28-point half-range Gauss-Hermite set -----------------
def test_gh_quadrature_matches_shipped_demo_set():
    g = GOLD["gas"]
    Xis, w = dvset.dvGH(float(np.sqrt(2.0 * g["R"] * GOLD["T0"])), 28)      # doc/demo.tex:28
    assert util.rel_err(Xis, GOLD["Xis"]) < 1e-10
    assert util.rel_err(w, GOLD["weights"]) < 1e-9
    # quadrature identities: moments of a 1-D Maxwellian at the case temperature
    RT = g["R"] * GOLD["T0"]
    X, W = np.array(GOLD["Xis"]), np.array(GOLD["weights"])
    M = np.exp(-X ** 2 / (2 * RT)) / np.sqrt(2 * np.pi * RT)
    assert abs((W * M).sum() - 1.0) < 1e-13
    assert abs((W * M * X ** 2).sum() / RT - 1.0) < 1e-12
    assert abs((W * M * X).sum()) < 1e-10


def test_newton_cotes_rule():
    X, w = dvset.dvNC(1600.0, 41)                                            # setDV.py usage example
    assert len(X) == 41 and abs(X[0] + 1600) < 1e-12 and abs(X[-1] - 1600) < 1e-9
    # Boole's rule integrates polynomials up to degree 5 exactly
    for p in range(6):
        exact = 0.0 if p % 2 else 2 * 1600.0 ** (p + 1) / (p + 1)
        assert abs((w * X ** p).sum() - exact) <= 1e-13 * 2 * 1600.0 ** (p + 1)
    with pytest.raises(ValueError):
        dvset.dvNC(1600.0, 40)


@pytest.mark.skipif(not os.path.isdir(REF_DEMO), reason="reference tree not mounted")
def test_demo_cavity_reader_matches_golden_facts():
    c = cs.read_case(REF_DEMO)
    g = c.geom
    assert (g.nCells, g.nInternalFaces, g.nBoundaryFaces, g.nSolutionD) == (3600, 7080, 240, 2)
    assert g.patch_names == ["movingWall", "fixedWalls"] and g.patch_size == [60, 180]
    assert [p.kind for p in c.patches] == [cs.PATCH_MAXWELL_WALL, cs.PATCH_MAXWELL_WALL]
    assert np.allclose(c.U_b[:60], [50.0, 0, 0]) and np.allclose(c.U_b[60:], 0.0)
    assert np.array_equal(c.Xis, GOLD["Xis"]) and np.array_equal(c.weights, GOLD["weights"])
    assert abs(g.V.sum() - 0.1) < 1e-15
    # the shipped mesh and the generated 60x60 block have the same geometry up to numbering
    s = cs.cavity2d_case(60).geom
    assert np.allclose(np.sort(g.V), np.sort(s.V), rtol=1e-12)
    assert np.allclose(np.sort(g.deltaCoeffs), np.sort(s.deltaCoeffs), rtol=1e-10)


def test_golden_case_facts():
    assert GOLD["nCells"] == 3600 and GOLD["nInternalFaces"] == 7080 and GOLD["nPoints"] == 7442
    assert GOLD["nFacesAll"] == 14520 and GOLD["patch_sizes"] == [60, 180]
    assert abs(GOLD["gas"]["R"] - 208.244343891) < 1e-12 and GOLD["lid_U"] == [50.0, 0.0, 0.0]


def test_ratchet_channel_mesh():
    """BASELINE config 4: channel with a saw-tooth lower wall on triangular prisms, walls at three temperatures."""
    from dugksfoam_b200.polymesh import cell_centres_and_volumes, face_centres_and_areas
    c = cs.ratchet_channel_case(40, 10, 8, teeth=4, tooth_height=0.3)
    m = c.mesh
    Cf, Sf = face_centres_and_areas(m)
    C, V = cell_centres_and_volumes(m, Cf, Sf)
    nif = m.nInternalFaces
    closure = np.zeros((m.nCells, 3))
    np.add.at(closure, m.owner, Sf)
    np.add.at(closure, m.neighbour, -Sf[:nif])
    assert m.nCells == 2 * 40 * 10 and V.min() > 0 and np.abs(closure).max() < 1e-15
    assert abs(V.sum() - (4.0 - 0.5 * 0.3 * 4.0) * 0.1) < 1e-13          # rectangle minus the four triangular teeth
    sizes = {p.name: p.size for p in c.patches}
    assert sizes == {"topWall": 40, "ratchet": 40, "endWalls": 20}
    # the ratchet wall really is a saw tooth: its face normals alternate between the two flanks
    pr = [p for p in c.patches if p.name == "ratchet"][0]
    nx_ = Sf[nif + pr.start: nif + pr.start + pr.size, 0]
    assert (nx_ > 0).sum() == 32 and (nx_ < 0).sum() == 8                # rising flank 80 % of a period (outward normal tilts to +x... of the gas side), falling 20 %
    Tb = {p.name: set(np.round(c.T_b[p.start:p.start + p.size], 9)) for p in c.patches}
    assert Tb == {"topWall": {300.3}, "ratchet": {245.7}, "endWalls": {273.0}}


# ---- geometry identities ---------------------------------------------------------------
@pytest.mark.parametrize("mesh", [hex_block(5, 4, 3, (1.0, 0.8, 0.6), distort=0.2),
                                  hex_block(6, 5, 1, (1.0, 1.0, 0.1), two_d=True, distort=0.2),
                                  tri_prism_2d(5, 4, distort=0.15), voronoi_prism_2d(6, 5, (1.0, 0.8, 0.1))],
                         ids=["hex3d", "hex2d", "tri2d", "poly2d"])
def test_geometry_identities(mesh):
    from dugksfoam_b200.polymesh import cell_centres_and_volumes, face_centres_and_areas
    Cf, Sf = face_centres_and_areas(mesh)
    C, V = cell_centres_and_volumes(mesh, Cf, Sf)
    nif = mesh.nInternalFaces
    closure = np.zeros((mesh.nCells, 3))
    np.add.at(closure, mesh.owner, Sf)
    np.add.at(closure, mesh.neighbour, -Sf[:nif])
    assert np.abs(closure).max() < 1e-15                          # sum of Sf over a closed cell = 0
    lens = mesh.points.max(axis=0) - mesh.points.min(axis=0)
    assert abs(V.sum() - np.prod(lens)) < 1e-14                   # sum V = domain volume
    g = compute_geometry(mesh)
    # least-squares gradient is exact for linear fields, including boundary cells when the
    # boundary value is the exact linear value (stock leastSquares with boundary faces)
    a = np.array([0.3, -1.1, 0.7]) * (~g.empty_dirs)
    phi = C @ a + 2.0
    grad = np.zeros((g.nCells, 3))
    own, nei = g.owner[:nif], g.neighbour
    d = phi[nei] - phi[own]
    np.add.at(grad, own, g.ownLs * d[:, None])
    np.add.at(grad, nei, -g.neiLs * d[:, None])
    ob = g.owner[nif:]
    nHat = g.Sf[nif:] / np.linalg.norm(g.Sf[nif:], axis=1)[:, None]
    delta = nHat * np.einsum("ij,ij->i", nHat, g.Cf[nif:] - g.C[ob])[:, None]
    phib = phi[ob] + delta @ a                                     # value at C + delta
    np.add.at(grad, ob, g.patchLs * (phib - phi[ob])[:, None])
    assert np.abs(grad - a[None, :]).max() < 1e-10


def test_ogrid_cylinder_mesh():
    """The O-type mesh of BASELINE config 5: closed cells, volume of the polygonal annulus, outward boundary normals,
    owner -> neighbour face normals, upper-triangular face order, least-squares gradient exact for linear fields."""
    from dugksfoam_b200.polymesh import ogrid_cylinder
    nt, nr, r0, r1, th = 20, 7, 0.5, 6.0, 0.1
    mesh = ogrid_cylinder(nt, nr, r0, r1, thickness=th)
    g = compute_geometry(mesh)
    nif = g.nInternalFaces
    assert (g.nCells, nif, g.nBoundaryFaces, g.nSolutionD) == (nt * nr, nt * nr + nt * (nr - 1), 2 * nt, 2)
    assert g.patch_names == ["cylinder", "farField"] and (g.owner[:nif] < g.neighbour).all()
    key = g.owner[:nif].astype(np.int64) * g.nCells + g.neighbour
    assert (np.diff(key) > 0).all()                                        # sorted by owner, then neighbour
    closure = np.zeros((g.nCells, 3))
    np.add.at(closure, g.owner[:nif], g.Sf[:nif]); np.add.at(closure, g.neighbour, -g.Sf[:nif])
    np.add.at(closure, g.owner[nif:], g.Sf[nif:])
    assert np.abs(closure[:, :2]).max() < 1e-14
    assert abs(g.V.sum() - th * 0.5 * nt * np.sin(2 * np.pi / nt) * (r1 ** 2 - r0 ** 2)) < 1e-12 and (g.V > 0).all()
    assert (np.einsum("ij,ij->i", g.Sf[nif:], g.Cf[nif:] - g.C[g.owner[nif:]]) > 0).all()
    assert (np.einsum("ij,ij->i", g.Sf[:nif], g.C[g.neighbour] - g.C[g.owner[:nif]]) > 0).all()
    rin = np.hypot(g.Cf[nif:nif + nt, 0], g.Cf[nif:nif + nt, 1])
    assert np.allclose(rin, r0 * np.cos(np.pi / nt))                      # chord mid-points of the cylinder polygon
    a = np.array([0.3, -1.1, 0.0])
    phi = g.C @ a + 2.0
    grad = np.zeros((g.nCells, 3))
    d = phi[g.neighbour] - phi[g.owner[:nif]]
    np.add.at(grad, g.owner[:nif], g.ownLs * d[:, None]); np.add.at(grad, g.neighbour, -g.neiLs * d[:, None])
    ob = g.owner[nif:]
    nh = g.Sf[nif:] / np.linalg.norm(g.Sf[nif:], axis=1)[:, None]
    delta = nh * np.einsum("ij,ij->i", nh, g.Cf[nif:] - g.C[ob])[:, None]
    np.add.at(grad, ob, g.patchLs * (delta @ a)[:, None])
    assert np.abs(grad - a[None, :]).max() < 1e-10


# ---- oracle identities -------------------------------------------------------------------
def _run(oracle_lib, case, nsteps, co=0.5, **kw):
    o = oracle_lib.Oracle(case, **kw)
    dt = case.courant_dt(co)
    for _ in range(nsteps):
        o.step(dt)
    return o


def test_uniform_equilibrium_is_steady(oracle_lib):
    """Gas at rest between isothermal walls at the gas temperature stays at rest up to the
    quadrature error of the 28-point set."""
    case = cs.cavity2d_case(8, 28)
    case.U_b[:] = 0.0
    o = _run(oracle_lib, case, 5)
    m = o.cell_macros()
    assert util.rel_err(m["rho"], case.rho) < 1e-11
    assert util.rel_err(m["T"], case.T) < 1e-11
    assert np.abs(m["U"]).max() < 1e-8 * np.sqrt(2 * case.gas["R"] * 273.0)
    o.close()


@pytest.mark.parametrize("D", [2, 3])
def test_velocity_partition_independence(oracle_lib, D):
    """serial == -dvParallel decomposition up to summation order (fvDVM.C:228-260)."""
    case = cs.cavity2d_case(8, 8, perturb=0.01) if D == 2 else cs.cavity3d_case(4, 8, perturb=0.01)
    ref = _run(oracle_lib, case, 3)
    a = ref.cell_macros()
    for P, part in ((4, 0), (3, 1)):
        o = _run(oracle_lib, case, 3, nranks=P, partition=part)
        b = o.cell_macros()
        sc = util.macro_scales(case)
        assert util.rel_err(b["rho"], a["rho"]) < 1e-13
        assert util.rel_err(b["T"], a["T"]) < 1e-13
        assert util.rel_err(b["U"], a["U"], sc["U"]) < 1e-13
        assert util.rel_err(b["q"], a["q"], sc["q"]) < 1e-13
        assert util.rel_err(o.state()[0], ref.state()[0]) < 1e-13
        o.close()
    ref.close()


def test_mirror_symmetry_of_the_cavity(oracle_lib):
    """A cavity driven by two opposite lids moving in opposite directions is point-symmetric;
    the solution must keep that symmetry to round-off (checks owner/neighbour handling,
    upwinding and wall treatment for both signs of every velocity component)."""
    n = 8
    mesh = hex_block(n, n, 1, (1.0, 1.0, 0.1), two_d=True,
                     patch_names={"ymax": "movingWall", "ymin": "lid2", "xmin": "fixedWalls", "xmax": "fixedWalls"})
    Xis, w = cs.gh_set(8)
    case = cs._uniform_case(mesh, Xis, w, {}, bc_overrides={"lid2": dict(U=(-50.0, 0, 0))})
    o = _run(oracle_lib, case, 6)
    m = o.cell_macros()
    rho = m["rho"].reshape(n, n)
    Ux = m["U"][:, 0].reshape(n, n)
    assert util.rel_err(rho[::-1, ::-1], rho) < 1e-13
    assert util.rel_err(-Ux[::-1, ::-1], Ux) < 1e-12
    o.close()


def test_half_cavity_with_symmetry_patch(oracle_lib):
    """demo/testSymmetry: a half cavity closed by a DVMsymmetry patch reproduces the flow of the
    mirrored full cavity.  Here the full problem is the symmetric double-lid cavity (both lids move
    in +x), whose solution is mirror-symmetric about y = 0.5."""
    n = 8
    Xis, w = cs.gh_set(8)
    full_mesh = hex_block(n, n, 1, (1.0, 1.0, 0.1), two_d=True,
                          patch_names={"ymax": "movingWall", "ymin": "lid2", "xmin": "fixedWalls", "xmax": "fixedWalls"})
    full = cs._uniform_case(full_mesh, Xis, w, {}, bc_overrides={"lid2": dict(U=(50.0, 0, 0))})
    half_mesh = hex_block(n, n // 2, 1, (1.0, 0.5, 0.1), two_d=True,
                          patch_names={"ymax": "symmetryWall", "ymin": "lid2", "xmin": "fixedWalls", "xmax": "fixedWalls"})
    half = cs._uniform_case(half_mesh, Xis, w, {"symmetryWall": cs.PATCH_DVM_SYMMETRY}, lid_patch="lid2")
    dt = full.courant_dt(0.5)
    of, oh = oracle_lib.Oracle(full), oracle_lib.Oracle(half)
    for _ in range(20):
        of.step(dt); oh.step(dt)
    mf, mh = of.cell_macros(), oh.cell_macros()
    # the symmetry treatment is a boundary closure, not an identity: agreement is to the
    # truncation error of the boundary reconstruction, far below the flow signal
    sig = np.abs(mf["U"][:, 0]).max()
    assert np.abs(mf["U"][: n * n // 2, 0] - mh["U"][:, 0]).max() < 0.05 * sig
    assert util.rel_err(mh["rho"], mf["rho"][: n * n // 2]) < 1e-3   # flow signal in rho is ~5e-2
    of.close(); oh.close()


def test_mass_change_equals_boundary_flux_only(oracle_lib):
    """Internal-face fluxes cancel exactly: with zero-velocity walls and a symmetric start the
    total mass changes only through the (tiny) wall-flux imbalance of the scheme."""
    case = cs.cavity2d_case(8, 8, perturb=0.02)
    o = oracle_lib.Oracle(case)
    m0 = (o.cell_macros()["rho"] * case.geom.V).sum()
    dt = case.courant_dt(0.5)
    for _ in range(5):
        o.step(dt)
    m1 = (o.cell_macros()["rho"] * case.geom.V).sum()
    assert abs(m1 - m0) / m0 < 1e-6
    o.close()


def test_venkatakrishnan_limiter(oracle_lib):
    """gradSchemes "VenkatakrishnanLimited leastSquares k" as it is meant to work (oracle_set_grad_scheme): k = 0
    switches it off (VenkatakrishnanLimitedGrads.C:70), a huge k is the unlimited gradient (eps^2 = k^3 V dominates),
    a small k limits the reconstruction at the under-resolved start of the Ma 5 cylinder: the overshoot of the
    face values (negative distribution functions after the first step) shrinks."""
    case = cs.cylinder_case(16, 6, 29, perturb=0.01)
    dt = case.courant_dt(0.5)
    out = {}
    for k in (None, 0.0, 1e3, 1e-9):
        o = oracle_lib.Oracle(case)
        if k is not None:
            o.set_grad_scheme(1, k)
        o.step(dt)
        out[k] = (o.state()[0], o.cell_macros())
        o.close()
    assert np.array_equal(out[None][0], out[0.0][0])
    assert util.rel_err(out[1e3][0], out[None][0]) < 1e-13
    gmax = np.abs(out[None][0]).max()
    assert out[None][0].min() < -0.05 * gmax and out[1e-9][0].min() > 0.5 * out[None][0].min()
    assert util.rel_err(out[1e-9][1]["T"], out[None][1]["T"]) > 1e-6           # it does change the solution
    mass = lambda m: float((m["rho"] * case.geom.V).sum())
    assert abs(mass(out[1e-9][1]) - mass(out[None][1])) < 1e-3 * mass(out[None][1])


def test_courant_number(oracle_lib):
    case = cs.cavity2d_case(8, 8)
    o = oracle_lib.Oracle(case)
    dt = case.courant_dt(0.8)
    mx, mean = o.courant(dt)
    assert abs(mx - 0.8) < 1e-12 and abs(mean - 0.8) < 1e-12          # uniform mesh, U = 0
    o.close()


def test_convergence_monitor(oracle_lib):
    """dugksFoam.C:88-107: relative change of T, rho, U since the previous check; Told = T etc. afterwards."""
    case = cs.cavity2d_case(8, 8, perturb=0.01)
    orc = oracle_lib.Oracle(case)
    m0 = orc.cell_macros()
    dt = case.courant_dt(0.5)
    for _ in range(3):
        orc.step(dt)
    m1 = orc.cell_macros()
    got = orc.convergence()
    want = (np.abs(m1["T"] - m0["T"]).sum() / m1["T"].sum(),
            np.abs(m1["rho"] - m0["rho"]).sum() / m1["rho"].sum(),
            np.linalg.norm(m1["U"] - m0["U"], axis=1).sum() / np.linalg.norm(m1["U"], axis=1).sum())
    assert np.allclose(got, want, rtol=1e-12, atol=0)
    assert all(v > 0 for v in got)
    assert orc.convergence() == (0.0, 0.0, 0.0)      # the snapshot was replaced by the current fields
    orc.close()


# ---- a second, independent restatement (oracle/dugks_numpy.py) ---------------------------------------------
def _xcheck_zoo():
    K = cs
    p0 = K.RHO0 * K.ARGON["R"] * K.T0
    return [
        ("cavity2d_6_gh8_distort", cs.cavity2d_case(6, 8, distort=0.2, perturb=0.02)),
        ("cavity3d_4_gh8_distort", cs.cavity3d_case(4, 8, distort=0.1, perturb=0.02)),
        ("tri_5_gh8", cs.tri_cavity_case(5, 8, perturb=0.02)),
        ("poly_5_gh8", cs.poly_cavity_case(5, 8, perturb=0.02)),
        ("cavity2d_5_nc9_ties", cs.cavity2d_case(5, 9, quad="NC", perturb=0.02)),
        ("farfield_zg_mixed", util.channel_case(
            6, 4, 8, kinds={"inlet": K.PATCH_FAR_FIELD, "outlet": K.PATCH_ZERO_GRADIENT, "top": K.PATCH_MIXED},
            bc_overrides={"inlet": dict(U=(30.0, 0, 0), rho=1.2 * K.RHO0, T=290.0, U_bc=K.BC_ZERO_GRADIENT),
                          "top": dict(U=(20.0, 0, 0), T=280.0)}, perturb=0.01, distort=0.1)),
        ("pressure_in_out", util.channel_case(
            6, 4, 8, kinds={"inlet": K.PATCH_PRESSURE_IN, "outlet": K.PATCH_PRESSURE_OUT},
            bc_overrides={"inlet": dict(pressure=1.1 * p0), "outlet": dict(pressure=0.9 * p0)}, perturb=0.01)),
        # O-type mesh at Ma 5: faces at 45 degrees meet velocities with xi_x = xi_y, xi.Sf = 0 exactly - or 1e-18 if the
        # three products are summed in another order (the tie rule, discreteVelocity.C:513-529, then flips to one-sided)
        ("cylinder_16x6_nc41_ma5", cs.cylinder_case(16, 6, 41, perturb=0.01)),
        ("dvm_symmetry_xy", util.channel_case(
            6, 4, 8, kinds={"inlet": K.PATCH_DVM_SYMMETRY, "bottom": K.PATCH_DVM_SYMMETRY},
            bc_overrides={"top": dict(U=(40.0, 0, 0))}, perturb=0.01)),
    ]


@pytest.mark.parametrize("name,case", _xcheck_zoo(), ids=[c[0] for c in _xcheck_zoo()])
def test_two_independent_restatements_agree(oracle_lib, name, case):
    """The C oracle (stage by stage, scalar loops) and oracle/dugks_numpy.py (field at a time, written separately
    from the reference text, no shared code) give the same step to round-off on every boundary kind: a misreading
    of the reference would have to be made twice, independently, to go unnoticed."""
    from oracle.dugks_numpy import NumpyDVM
    o, p = oracle_lib.Oracle(case), NumpyDVM(case)
    dt = case.courant_dt(0.5)
    sc = util.macro_scales(case)
    assert np.allclose(o.courant(dt), p.courant(dt), rtol=1e-13)
    for s in range(3):
        o.step(dt * (1 + 0.1 * s)); p.step(dt * (1 + 0.1 * s))
    m, f, b, wd = o.cell_macros(), o.face_macros(), o.boundary_macros(), o.wall_diag()
    g, h = o.state()
    errs = dict(rho=util.rel_err(p.rho, m["rho"]), T=util.rel_err(p.T, m["T"]), U=util.rel_err(p.U, m["U"], sc["U"]),
                q=util.rel_err(p.q, m["q"], sc["q"]), tau=util.rel_err(p.tau, m["tau"]),
                face_rho=util.rel_err(p.rhoSurf, f["rho"]), face_q=util.rel_err(p.qSurf, f["q"], sc["q"]),
                face_U=util.rel_err(p.Usurf, f["U"], sc["U"]), bnd_rho=util.rel_err(p.rho_b, b["rho"], sc["rho"]),
                bnd_U=util.rel_err(p.U_b, b["U"], sc["U"]), bnd_T=util.rel_err(p.T_b, b["T"]),
                g=util.rel_err(p.gT, g), h=util.rel_err(p.hT, h, max(np.abs(h).max(), 1e-300)),
                qWall=util.rel_err(p.qWall, wd["qWall"], sc["q"]),
                stressWall=util.rel_err(p.stressWall.reshape(-1, 9), wd["stressWall"], sc["rho"] * sc["U"] ** 2))
    assert max(errs.values()) <= 1e-13, errs
    assert np.allclose(o.courant(dt), p.courant(dt), rtol=1e-12)
    o.close()


def test_geometry_against_independent_formulas():
    """[OF-lib] geometry checked WITHOUT polymesh.py's code path: areas and
    centroids by the shoelace / divergence-theorem formulas from a different apex, interpolation weights, patch
    deltas and the least-squares vectors rebuilt cell by cell with plain loops from their definition."""
    sheared = hex_block(3, 2, 2, (1.0, 0.8, 0.6))
    # an affine map keeps faces planar (a random distortion in 3-D warps them, and the fan about the vertex average
    # that OpenFOAM uses then differs from any other triangulation at second order in the warp)
    sheared.points = np.ascontiguousarray(sheared.points @ np.array([[1.0, 0.3, 0.1], [0.2, 1.0, -0.25], [-0.15, 0.1, 1.0]]).T)
    for mesh in (hex_block(3, 3, 1, (1.0, 0.9, 0.1), two_d=True, distort=0.25), sheared):
        g = compute_geometry(mesh)
        P = mesh.points
        nF = mesh.nFaces
        # faces: area vector = 1/2 sum p_i x p_{i+1}; centroid = area-weighted centroid of the fan about vertex 0
        Sf = np.zeros((nF, 3)); Cf = np.zeros((nF, 3))
        for f in range(nF):
            v = P[mesh.face_verts[mesh.face_offsets[f]:mesh.face_offsets[f + 1]]]
            S = sum(np.cross(v[i], v[(i + 1) % len(v)]) for i in range(len(v))) / 2.0
            num, den = np.zeros(3), 0.0
            for i in range(1, len(v) - 1):
                a = np.cross(v[i] - v[0], v[i + 1] - v[0]) / 2.0
                wgt = a @ S / np.linalg.norm(S)
                num += wgt * (v[0] + v[i] + v[i + 1]) / 3.0; den += wgt
            Sf[f], Cf[f] = S, num / den
        # cells: V = 1/3 sum Cf.Sf (outward); centroid from tetrahedra (apex = a vertex of the cell)
        own_all, nei = mesh.owner, mesh.neighbour
        V = np.zeros(mesh.nCells); Cc = np.zeros((mesh.nCells, 3)); apex = {}
        faces_of = [[] for _ in range(mesh.nCells)]
        for f in range(nF):
            faces_of[own_all[f]].append((f, 1.0))
            if f < len(nei):
                faces_of[nei[f]].append((f, -1.0))
        for c in range(mesh.nCells):
            a0 = P[mesh.face_verts[mesh.face_offsets[faces_of[c][0][0]]]]
            vol, mom = 0.0, np.zeros(3)
            for f, sgn in faces_of[c]:
                v = P[mesh.face_verts[mesh.face_offsets[f]:mesh.face_offsets[f + 1]]]
                for i in range(1, len(v) - 1):
                    t = sgn * np.dot(np.cross(v[i] - v[0], v[i + 1] - v[0]), v[0] - a0) / 6.0   # signed tetrahedron volume
                    vol += t; mom += t * (a0 + v[0] + v[i] + v[i + 1]) / 4.0
            V[c], Cc[c] = vol, mom / vol
        # the solver's face list = internal faces + faces of non-empty patches, in that order
        bfaces = []
        for p in mesh.patches:
            if p.type != "empty":
                bfaces += list(range(p.startFace, p.startFace + p.nFaces))
        sel = np.r_[np.arange(g.nInternalFaces), np.array(bfaces, dtype=int)]
        tolg = 1e-12
        assert np.allclose(g.Sf, Sf[sel], rtol=0, atol=1e-14) and np.allclose(g.Cf, Cf[sel], rtol=0, atol=tolg)
        assert np.allclose(g.V, V, rtol=tolg) and np.allclose(g.C, Cc, rtol=0, atol=tolg)
        # least-squares vectors from their definition, with polymesh's own C, Cf, Sf as input (Appendix C 3-6)
        nif = g.nInternalFaces
        dd = np.zeros((g.nCells, 3, 3))
        wf = np.zeros(nif)
        for f in range(nif):
            o_, n_ = g.owner[f], g.neighbour[f]
            sn, sp = abs(g.Sf[f] @ (g.C[n_] - g.Cf[f])), abs(g.Sf[f] @ (g.Cf[f] - g.C[o_]))
            wf[f] = sn / (sp + sn)
            d = g.C[n_] - g.C[o_]
            wdd = np.linalg.norm(g.Sf[f]) / (d @ d) * np.outer(d, d)
            dd[o_] += (1 - wf[f]) * wdd; dd[n_] += wf[f] * wdd
            assert abs(g.deltaCoeffs[f] - 1 / np.linalg.norm(d)) < 1e-12 / np.linalg.norm(d)
        deltas = []
        for b in range(g.nBoundaryFaces):
            f = nif + b; o_ = g.owner[f]
            nh = g.Sf[f] / np.linalg.norm(g.Sf[f])
            d = nh * (nh @ (g.Cf[f] - g.C[o_]))
            deltas.append(d)
            dd[o_] += np.linalg.norm(g.Sf[f]) / (d @ d) * np.outer(d, d)
            assert abs(g.deltaCoeffs[f] - 1 / np.linalg.norm(d)) < 1e-12 / np.linalg.norm(d)
        solved = ~g.empty_dirs
        inv = np.zeros_like(dd)
        for c in range(g.nCells):
            sub = dd[c][np.ix_(solved, solved)]
            inv[c][np.ix_(solved, solved)] = np.linalg.inv(sub)
        for f in range(nif):
            o_, n_ = g.owner[f], g.neighbour[f]
            d = g.C[n_] - g.C[o_]
            k = np.linalg.norm(g.Sf[f]) / (d @ d)
            assert np.allclose(g.ownLs[f], (1 - wf[f]) * k * (inv[o_] @ d), rtol=1e-10, atol=1e-12 * np.abs(g.ownLs).max())
            assert np.allclose(g.neiLs[f], -wf[f] * k * (inv[n_] @ d), rtol=1e-10, atol=1e-12 * np.abs(g.neiLs).max())
        for b in range(g.nBoundaryFaces):
            f = nif + b; o_ = g.owner[f]; d = deltas[b]
            assert np.allclose(g.patchLs[b], np.linalg.norm(g.Sf[f]) / (d @ d) * (inv[o_] @ d), rtol=1e-10,
                               atol=1e-12 * np.abs(g.patchLs).max())
