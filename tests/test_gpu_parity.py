"""Parity of the CUDA path (through the C-ABI) against the CPU oracle.

Bar (BASELINE.json north_star): 1e-12 relative per step, 1e-9 on rho/U/T/q after
1000 steps, FP64.  Relative = max|a-b| / max|b| per field (vector fields whose
reference value may vanish are scaled by the thermal speed / rho c^3).
"""
import numpy as np
import pytest

from dugksfoam_b200 import capi
from dugksfoam_b200 import case as cs
import parity_util as util

pytestmark = pytest.mark.gpu


def _compare(dv, orc, case, tol, label):
    sc = util.macro_scales(case)
    errs = {}
    cm_g, cm_o = dv.cell_macros(), orc.cell_macros()
    fm_g, fm_o = dv.face_macros(), orc.face_macros()
    for k in ("rho", "T", "tau"):
        errs["cell_" + k] = util.rel_err(cm_g[k], cm_o[k])
        errs["face_" + k] = util.rel_err(fm_g[k], fm_o[k])
    for k in ("U", "q"):
        errs["cell_" + k] = util.rel_err(cm_g[k], cm_o[k], sc[k])
        errs["face_" + k] = util.rel_err(fm_g[k], fm_o[k], sc[k])
    bm_g, bm_o = dv.boundary_macros(), orc.boundary_macros()
    if not label.endswith("init"):   # wall rho starts at 1 in the reference (calculatedMaxwellFvPatchField.C:80), unused
        errs["bnd_rho"] = util.rel_err(bm_g["rho"], bm_o["rho"], sc["rho"])
    errs["bnd_U"] = util.rel_err(bm_g["U"], bm_o["U"], sc["U"])
    errs["bnd_T"] = util.rel_err(bm_g["T"], bm_o["T"], sc["T"])
    wd_g, wd_o = dv.wall_diag(), orc.wall_diag()
    errs["qWall"] = util.rel_err(wd_g["qWall"], wd_o["qWall"], sc["q"])
    errs["stressWall"] = util.rel_err(wd_g["stressWall"], wd_o["stressWall"], sc["rho"] * sc["U"] ** 2)
    g_g, h_g = dv.state()
    g_o, h_o = orc.state()
    ids = dv.local_dvs()
    errs["gTilde"] = util.rel_err(g_g, g_o[ids])
    errs["hTilde"] = util.rel_err(h_g, h_o[ids], max(np.abs(h_o).max(), 1e-300))
    bad = {k: v for k, v in errs.items() if not (v <= tol)}
    assert not bad, f"{label}: parity above {tol:g}: {bad}  (all: {errs})"
    return errs


@pytest.mark.parametrize("name,case,store_h", util.cases_small(), ids=[c[0] for c in util.cases_small()])
def test_step_parity(oracle_lib, name, case, store_h):
    dv = capi.fvDVM(case, store_h=store_h)
    orc = oracle_lib.Oracle(case)
    dt = case.courant_dt(0.5)
    assert np.allclose(dv.getCoNum(dt), orc.courant(dt), rtol=1e-12)
    # initial state (discreteVelocity::initDFtoEq)
    _compare(dv, orc, case, util.TOL_STEP, name + " init")
    for step in range(4):
        dtk = dt * (1.0 + 0.1 * step)           # dt changes every step (adjustTimeStep)
        dv.evolution(dtk)
        orc.step(dtk)
        _compare(dv, orc, case, util.TOL_STEP * (step + 1), f"{name} step {step + 1}")
        assert np.allclose(dv.getCoNum(dtk), orc.courant(dtk), rtol=1e-10)
    # boundary face DFs
    gb, hb = dv.boundary_surf()
    nif = case.geom.nInternalFaces
    ids = dv.local_dvs()
    ref = np.stack([orc.surf(int(k))[0][nif:] for k in ids])
    assert util.rel_err(gb, ref) <= 1e-11
    dv.close(); orc.close()


def _bc_cases():
    K = cs
    out = []
    # far field in / zero-gradient out, mixed (fixedValue rho) top, Maxwell wall bottom
    out.append(("farfield_zg_mixed", util.channel_case(
        10, 6, 8, kinds={"inlet": K.PATCH_FAR_FIELD, "outlet": K.PATCH_ZERO_GRADIENT, "top": K.PATCH_MIXED},
        bc_overrides={"inlet": dict(U=(30.0, 0, 0), rho=1.2 * K.RHO0, T=290.0, U_bc=K.BC_ZERO_GRADIENT),
                      "top": dict(U=(20.0, 0, 0), T=280.0)}, perturb=0.01)))
    # pressure inlet / outlet
    p0 = K.RHO0 * K.ARGON["R"] * K.T0
    out.append(("pressure_in_out", util.channel_case(
        10, 6, 8, kinds={"inlet": K.PATCH_PRESSURE_IN, "outlet": K.PATCH_PRESSURE_OUT},
        bc_overrides={"inlet": dict(pressure=1.1 * p0), "outlet": dict(pressure=0.9 * p0)}, perturb=0.01)))
    # symmetryMod (DVMsymmetry) on the left, symmetryPlane-like on the bottom
    out.append(("dvm_symmetry", util.channel_case(
        8, 6, 8, kinds={"inlet": K.PATCH_DVM_SYMMETRY, "bottom": K.PATCH_SYMMETRY_PLANE},
        bc_overrides={"top": dict(U=(40.0, 0, 0))}, perturb=0.01)))
    return out


@pytest.mark.parametrize("name,case", _bc_cases(), ids=[c[0] for c in _bc_cases()])
def test_boundary_kinds(oracle_lib, name, case):
    dv = capi.fvDVM(case)
    orc = oracle_lib.Oracle(case)
    dt = case.courant_dt(0.5)
    for step in range(4):
        dv.evolution(dt)
        orc.step(dt)
        _compare(dv, orc, case, util.TOL_STEP * (step + 1), f"{name} step {step + 1}")
    dv.close(); orc.close()


_PATHS = [
    ("keep_all", {}),                                   # face-storage slabs: fused relax+update
    ("keep_none", {"DUGKS_KEEP_SLABS": "0"}),           # flux-buffer path (what one GPU uses when memory is short)
    ("keep_one", {"DUGKS_KEEP_SLABS": "1"}),            # both kinds of slab in one step
    ("no_axis", {"DUGKS_NO_AXIS": "1"}),                # general path also for axis-aligned cells
    ("no_axis_keep_none", {"DUGKS_NO_AXIS": "1", "DUGKS_KEEP_SLABS": "0"}),
    ("no_split", {"DUGKS_NO_SPLIT_AXIS": "1"}),         # axis-aligned cells inside the unified phase-1 launch
    ("no_wmode", {"DUGKS_NO_WMODE": "1"}),              # face-storage slabs with persistent gBarP (update reads gTilde, gBarP)
    ("no_wmode_keep_one", {"DUGKS_NO_WMODE": "1", "DUGKS_KEEP_SLABS": "1"}),
    ("order_tiled", {"DUGKS_ORDER": "tiled"}),          # strips of rows (the default is the x-wavefront order, dugks_cell_order)
    ("order_natural", {"DUGKS_ORDER": "natural"}),
    ("pencil_off", {"DUGKS_PENCIL": "0"}),              # phase 1 without CTA pencils (warp-per-cell kernels everywhere)
    ("pencil_unfused", {"DUGKS_PENCIL": "1"}),          # CTA pencils over gBarP written by the half-step kernel
    ("pencil_keep_one", {"DUGKS_KEEP_SLABS": "1"}),     # fused pencils on the face-storage slab, unfused on the others
    ("pencil_ws", {"DUGKS_PENCIL_WS": "1"}),            # warp-specialised fused pencils (producer + consumer warp per line, mbarrier ring)
    ("pencil_ws_keep_one", {"DUGKS_PENCIL_WS": "1", "DUGKS_KEEP_SLABS": "1"}),
    ("gam_pingpong", {"DUGKS_GAM_PINGPONG": "1"}),      # two copies of the lagged boundary gradient (the default updates one in place)
    ("gam_pingpong_keep_none", {"DUGKS_GAM_PINGPONG": "1", "DUGKS_KEEP_SLABS": "0"}),
    ("gen1_tma", {"DUGKS_NO_HOT": "1"}),                # first-generation bulk-copy kernels
    ("gen1_ldg", {"DUGKS_NO_HOT": "1", "DUGKS_NO_TMA": "1"}),
    ("generic", {"DUGKS_NO_HOT": "1", "DUGKS_FORCE_GENERIC": "1"}),   # cells with many faces
]


@pytest.mark.parametrize("path,env", _PATHS, ids=[p[0] for p in _PATHS])
def test_every_kernel_path(oracle_lib, monkeypatch, path, env):
    """Every device code path that can carry the step gives the oracle's answer."""
    for k in ("DUGKS_KEEP_SLABS", "DUGKS_NO_HOT", "DUGKS_NO_TMA", "DUGKS_FORCE_GENERIC", "DUGKS_NO_AXIS",
              "DUGKS_NO_SPLIT_AXIS", "DUGKS_NO_WMODE", "DUGKS_ORDER", "DUGKS_PENCIL", "DUGKS_PENCIL_WS", "DUGKS_GAM_PINGPONG"):
        monkeypatch.delenv(k, raising=False)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    zoo = [("cavity3d_5_gh8_distort", cs.cavity3d_case(5, 8, distort=0.15, perturb=0.01), False),
           ("cavity3d_6_gh28", cs.cavity3d_case(6, 28, perturb=0.01), False),      # two slabs
           ("cavity3d_10_gh8", cs.cavity3d_case(10, 8, perturb=0.01), False),      # mostly interior cells: axis-only launch, CTA pencils
           ("cavity3d_11_gh28", cs.cavity3d_case(11, 28, perturb=0.01), False),    # 9 x 9 interior lines: 16 bundles + lines left to the axis-only launch; 25 slabs with a short-row tail
           ("cavity2d_12_gh28", cs.cavity2d_case(12, 28, perturb=0.01), False),    # same in 2-D, with h
           # more cells than resident warps (148 SMs x 12): persistent warps carry state from cell to cell
           ("cavity3d_20_gh8", cs.cavity3d_case(20, 8, perturb=0.01), False),
           ("cavity2d_48_gh28", cs.cavity2d_case(48, 28, perturb=0.01), False),
           ("cavity3d_10_gh8_storeh", cs.cavity3d_case(10, 8, perturb=0.01), True),   # axis-only launch, 6 faces, with h
           ("cavity2d_9_nc9_ties", cs.cavity2d_case(9, 9, quad="NC", perturb=0.01), False),
           ("tri_8_gh8", cs.tri_cavity_case(8, 8, perturb=0.01), False),
           ("ratchet_20x8_gh8", cs.ratchet_channel_case(20, 8, 8, teeth=2, perturb=0.01), False),   # config 4 as named
           ("poly_7_gh8", cs.poly_cavity_case(7, 8, perturb=0.01), False),         # polygonal cells with up to 8 sides
           ("cavity3d_4_gh8_storeh", cs.cavity3d_case(4, 8, perturb=0.01), True)]
    for name, case, store_h in zoo:
        dv = capi.fvDVM(case, store_h=store_h)
        st = dv.stats()
        if path == "keep_all":
            assert st["keep_slabs"] == st["n_slabs"]
        if name in ("cavity3d_10_gh8", "cavity3d_20_gh8", "cavity3d_11_gh28") and "DUGKS_NO_HOT" not in env and "DUGKS_NO_AXIS" not in env \
                and "DUGKS_NO_SPLIT_AXIS" not in env:
            n_int = {"cavity3d_10_gh8": 8, "cavity3d_20_gh8": 18, "cavity3d_11_gh28": 8}[name]
            want_mode = int(env.get("DUGKS_PENCIL", "2"))
            assert st["pencil_mode"] == want_mode and st["pencil_cells"] == (n_int ** 2 * (n_int if name != "cavity3d_11_gh28" else 9) if want_mode else 0), st
        if path in ("keep_none", "gen1_tma", "gen1_ldg", "generic", "no_axis_keep_none"):
            assert st["keep_slabs"] == 0
        orc = oracle_lib.Oracle(case)
        dt = case.courant_dt(0.5)
        for step in range(3):
            dv.evolution(dt * (1.0 + 0.1 * step))
            orc.step(dt * (1.0 + 0.1 * step))
            _compare(dv, orc, case, util.TOL_STEP * (step + 1), f"{path}/{name} step {step + 1}")
        dv.close(); orc.close()


@pytest.mark.parametrize("which", ["cavity2d", "cavity3d_perturbed"])
def test_thousand_steps(oracle_lib, which):
    """1e-9 on rho/U/T/q after 1000 steps (north_star)."""
    case = cs.cavity2d_case(10, 8) if which == "cavity2d" else cs.cavity3d_case(5, 8, distort=0.1, perturb=0.01)
    dv = capi.fvDVM(case)
    orc = oracle_lib.Oracle(case)
    dt = case.courant_dt(0.8)
    for _ in range(1000):
        dv.evolution(dt)
        orc.step(dt)
    sc = util.macro_scales(case)
    a, b = dv.cell_macros(), orc.cell_macros()
    assert util.rel_err(a["rho"], b["rho"]) <= util.TOL_LONG
    assert util.rel_err(a["T"], b["T"]) <= util.TOL_LONG
    assert util.rel_err(a["U"], b["U"], np.abs(b["U"]).max()) <= util.TOL_LONG
    assert util.rel_err(a["q"], b["q"], np.abs(b["q"]).max()) <= util.TOL_LONG
    dv.close(); orc.close()


def test_convergence_monitor(oracle_lib):
    """dugks_convergence (dugksFoam.C:88-107 on the device) against the oracle: change since create, change
    between two checks, and zero right after a check."""
    case = cs.cavity3d_case(6, 8, perturb=0.01)
    dv = capi.fvDVM(case)
    orc = oracle_lib.Oracle(case)
    dt = case.courant_dt(0.5)
    sc = util.macro_scales(case)
    done = 0
    for nsteps in (2, 3):
        for _ in range(nsteps):
            dv.evolution(dt)
            orc.step(dt)
        done += nsteps
        got, want = dv.convergence(), orc.convergence()
        assert all(w > 0 for w in want)
        # the fields agree to TOL_STEP * steps of their scale; a sum of |differences| over a sum inherits twice that
        m = orc.cell_macros()
        den = (m["T"].sum(), m["rho"].sum(), np.linalg.norm(m["U"], axis=1).sum())
        for g, w, scale, d in zip(got, want, (sc["T"], sc["rho"], sc["U"]), den):
            assert abs(g - w) <= 2 * util.TOL_STEP * done * scale * case.nCells / d + 1e-13 * abs(w), (got, want)
    assert dv.convergence() == (0.0, 0.0, 0.0)
    dv.close(); orc.close()


def test_set_get_state_roundtrip():
    case = cs.cavity2d_case(6, 8)
    dv = capi.fvDVM(case)
    g, h = dv.state()
    rng = np.random.default_rng(1)
    g2 = g * (1 + 0.01 * rng.random(g.shape))
    dv.set_state(g2, h)
    g3, _ = dv.state()
    assert np.array_equal(g2, g3)
    gd, hd = dv.writeDFonCell(3)
    assert np.array_equal(gd[dv.local_dvs()], g3[:, 3])
    st = dv.stats()
    assert st["kernel_launches"] > 0
    dv.close()


def test_page_locked_accessor_arrays():
    """dugks_host_register: macro accessors that copy straight into page-locked caller arrays return the same bits as the
    staged path (include/dugks.h, accessors)."""
    case = cs.cavity3d_case(6, 8, perturb=0.01)
    dv = capi.fvDVM(case)
    dt = case.courant_dt(0.5)
    for _ in range(2):
        dv.evolution(dt)
    a = dv.cell_macros()
    b = dv.cell_macros(pinned=True)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    dv.evolution(dt)
    a = dv.cell_macros()
    b2 = dv.cell_macros(pinned=True)
    assert b2["rho"] is b["rho"]                       # the same page-locked arrays, refilled
    for k in a:
        assert np.array_equal(a[k], b2[k]), k
    dv.close()


def test_fails_loudly_on_bad_input():
    case = cs.cavity2d_case(4, 8)
    case.patches[0].kind = 99
    with pytest.raises(capi.DugksError):
        capi.fvDVM(case)


_CHUNKED = [
    # (label, case builder, environment): more than 32 velocity points per direction -> rows are cut into ix-chunks
    ("nc37", lambda: cs.cavity2d_case(10, 37, quad="NC", perturb=0.01), {}),
    ("nc101_cfg2_grid", lambda: cs.cavity2d_case(6, 101, quad="NC", perturb=0.01), {}),          # BASELINE config 2's velocity grid
    ("nc41_ties", lambda: cs.cavity2d_case(7, 41, quad="NC", xiMax=1600.0, perturb=0.01), {}),    # xi = 0 abscissa: ties on every axis-aligned face
    ("gh28_forced_2_chunks", lambda: cs.cavity2d_case(10, 28, perturb=0.01), {"DUGKS_NCH": "2"}),
    ("gh40_stable_recurrence", lambda: cs.cavity2d_case(6, 40, quad="GHs", perturb=0.01), {}),
    ("nc37_3d", lambda: _cavity3d_nc(4, 33), {}),
    ("nc41_distorted", lambda: cs.cavity2d_case(6, 41, quad="NC", distort=0.2, perturb=0.01), {}),
]


def _cavity3d_nc(n, nDV):
    from dugksfoam_b200 import dvset
    from dugksfoam_b200.polymesh import hex_block
    Xis, w = dvset.dvNC(4.0 * np.sqrt(2 * cs.ARGON["R"] * cs.T0), nDV)
    return cs._uniform_case(hex_block(n, n, n, (1.0, 1.0, 1.0)), Xis, w, {}, name=f"cavity3d_{n}_NC{nDV}", perturb=0.01)


@pytest.mark.parametrize("label,build,env", _CHUNKED, ids=[c[0] for c in _CHUNKED])
def test_chunked_rows_parity(oracle_lib, monkeypatch, label, build, env):
    """More than 32 velocity points per direction (BASELINE config 2 has 101, config 5 about 80)."""
    monkeypatch.delenv("DUGKS_NCH", raising=False)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    case = build()
    dv = capi.fvDVM(case)
    orc = oracle_lib.Oracle(case)
    dt = case.courant_dt(0.5)
    for step in range(3):
        dv.evolution(dt)
        orc.step(dt)
        _compare(dv, orc, case, util.TOL_STEP * (step + 1), f"chunked {label} step {step + 1}")
    dv.close(); orc.close()


def test_shipped_demo_cavity(oracle_lib):
    """BASELINE config 1: the reference's own demo/cavity (shipped polyMesh with its 4-block numbering, shipped
    Xis/weights, DVMProperties, 0/*; tests/golden/demo_cavity_case.npz), stepped as dugksFoam steps it: the first
    step at the controlDict deltaT, then the Courant-limited one (setDeltaTvar.H:34-47)."""
    case = util.demo_cavity_case()
    assert (case.nCells, case.nXi) == (3600, 784)
    dv = capi.fvDVM(case)
    orc = oracle_lib.Oracle(case)
    dt = case.courant_dt(case.maxCo)
    assert abs(dt - 5.3e-6) < 1e-7                  # doc: about 5.3e-6 s from step 2 on
    _compare(dv, orc, case, util.TOL_STEP, "demo/cavity init")
    for step in range(5):
        dv.evolution(dt)
        orc.step(dt)
        errs = _compare(dv, orc, case, util.TOL_STEP * (step + 1), f"demo/cavity step {step + 1}")
    assert np.allclose(dv.getCoNum(dt), orc.courant(dt), rtol=1e-10)
    # closed cavity, diffuse walls: internal fluxes cancel, the total mass only feels the (tiny) wall-flux
    # imbalance of the scheme (tests/test_oracle.py::test_mass_change_equals_boundary_flux_only) - and it is the
    # oracle's mass to round-off
    m0 = float((case.rho * case.geom.V).sum())
    m5 = float((dv.cell_macros()["rho"] * case.geom.V).sum())
    m5o = float((orc.cell_macros()["rho"] * case.geom.V).sum())
    assert abs(m5 - m0) <= 1e-6 * m0 and abs(m5 - m5o) <= 5e-12 * m0, (m0, m5, m5o)
    dv.close(); orc.close()


def test_full_size_properties():
    """BASELINE config 3 at its FULL size (64^3 cells x 28^3 velocities, 5.75e9 updates per step - the oracle would need
    minutes per step) through properties that do not depend on the size:
      * a gas at rest between walls at its own temperature stays where it is (every kernel path of the step: the half
        step, the reconstruction, the wall rule and the update must cancel to round-off);
      * the lid-driven cavity is mirror-symmetric about z = 1/2 (both signs of xi_z, owner and neighbour sides);
      * its total mass changes only by the wall-flux imbalance of the scheme."""
    import torch
    if torch.cuda.get_device_properties(0).total_memory < 150e9:
        pytest.skip("needs the 180 GB of a B200")
    n, nDV = 64, 28
    case = cs.cavity3d_case(n, nDV)
    dt = case.courant_dt(0.8)
    V = case.geom.V
    cth = float(np.sqrt(2.0 * case.gas["R"] * cs.T0))
    # ---- 1. equilibrium at rest
    rest = cs.cavity3d_case(n, nDV)
    rest.U_b[:] = 0.0
    dv = capi.fvDVM(rest)
    st = dv.stats()
    assert st["pencil_mode"] == 2 and st["pencil_cells"] == (n - 2) ** 3 and st["keep_slabs"] >= 19
    for _ in range(3):
        dv.evolution(dt)
    m = dv.cell_macros()
    # the 28-point Gauss-Hermite sums carry the discrete Maxwellian's moments to a few 1e-13 (tests/test_oracle.py:
    # test_uniform_equilibrium_is_steady holds the oracle to the same figures)
    assert util.rel_err(m["rho"], rest.rho) < 1e-11 and util.rel_err(m["T"], rest.T) < 1e-11
    assert np.abs(m["U"]).max() < 1e-8 * cth and np.abs(m["q"]).max() < 1e-7 * cs.RHO0 * cth ** 3
    dv.close()
    # ---- 2. and 3. the driven cavity
    dv = capi.fvDVM(case)
    for _ in range(3):
        dv.evolution(dt)
    m = dv.cell_macros()
    rho = m["rho"].reshape(n, n, n)                   # [k (z)][j (y)][i (x)], polymesh.hex_block numbering
    U = m["U"].reshape(n, n, n, 3)
    assert util.rel_err(rho[::-1], rho) < 1e-12
    assert np.abs(U[::-1, :, :, 0] - U[..., 0]).max() < 1e-11 * cth and np.abs(U[::-1, :, :, 2] + U[..., 2]).max() < 1e-11 * cth
    assert np.abs(U[..., 0]).max() > 1e-3 * 50.0      # the lid does drive the gas
    m0, m3 = float((case.rho * V).sum()), float((m["rho"] * V).sum())
    assert abs(m3 - m0) <= 1e-6 * m0
    assert np.isfinite(m["q"]).all() and np.isfinite(m["tau"]).all()
    dv.close()


@pytest.mark.parametrize("which", ["cfg2_cavity2d_256_nc101", "cfg4_ratchet_632x158_gh28"])
def test_full_size_gas_at_rest(which):
    """BASELINE configs 2 and 4 at their full sizes: a gas at rest between walls at its own temperature stays at rest
    (chunked velocity rows with h stored; 199,712 triangular prisms on the general path)."""
    if which.startswith("cfg2"):
        case = cs.cavity2d_case(256, 101, quad="NC")
    else:
        case = cs.ratchet_channel_case(632, 158, 28, T_ratchet=cs.T0, T_top=cs.T0)
    case.U_b[:] = 0.0
    dv = capi.fvDVM(case)
    dt = case.courant_dt(0.8)
    for _ in range(3):
        dv.evolution(dt)
    m = dv.cell_macros()
    cth = float(np.sqrt(2.0 * case.gas["R"] * cs.T0))
    # Newton-Cotes on [-4 c, 4 c] integrates the Maxwellian to 1e-7 only (the tails), Gauss-Hermite to round-off:
    # the discrete equilibrium is steady to that accuracy
    tol = 1e-6 if which.startswith("cfg2") else 1e-10
    assert util.rel_err(m["rho"], case.rho) < tol and util.rel_err(m["T"], case.T) < tol
    assert np.abs(m["U"]).max() < tol * 1e2 * cth
    assert np.isfinite(m["q"]).all()
    dv.close()


def _hypersonic_channel(nDV=29, nx=10, ny=6):
    """Free stream at Ma = 5 (argon, 273 K: a = 307.8 m/s) through far-field in / zeroGradient out, fixedValue-rho
    ("mixed") top, Maxwell wall bottom; Newton-Cotes grid wide enough for the shifted Maxwellian."""
    from dugksfoam_b200 import dvset
    from dugksfoam_b200.polymesh import hex_block
    K = cs
    Uinf = 5.0 * np.sqrt(5.0 / 3.0 * K.ARGON["R"] * K.T0)
    names = {"xmin": "inlet", "xmax": "outlet", "ymin": "bottom", "ymax": "top"}
    mesh = hex_block(nx, ny, 1, (1.0, 0.6, 0.1), two_d=True, patch_names=names, distort=0.1)
    Xis, w = dvset.dvNC(3200.0, nDV)
    return cs._uniform_case(
        mesh, Xis, w, {"inlet": K.PATCH_FAR_FIELD, "outlet": K.PATCH_ZERO_GRADIENT, "top": K.PATCH_MIXED}, lid_patch="none",
        U0=(Uinf, 0.0, 0.0), name="ma5_channel", perturb=0.01,
        bc_overrides={"inlet": dict(U=(Uinf, 0, 0), U_bc=K.BC_ZERO_GRADIENT), "top": dict(U=(Uinf, 0, 0))}), Uinf


def test_hypersonic_free_stream(oracle_lib):
    """Ma = 5 (BASELINE config 5's regime): the CUDA path forms U, T and the centred heat flux from RAW moments
    about zero and centres them algebraically; the oracle sums (xi - U) directly as fvDVM.C:503-516 does.
    The raw third moment is (U / c)^3 ~ 100 times the heat-flux scale here: the 1e-12 budget must survive it."""
    case, Uinf = _hypersonic_channel()
    dv = capi.fvDVM(case)
    orc = oracle_lib.Oracle(case)
    dt = case.courant_dt(0.5)
    for step in range(4):
        dv.evolution(dt)
        orc.step(dt)
        errs = _compare(dv, orc, case, util.TOL_STEP * (step + 1), f"Ma 5 step {step + 1}")
    m = dv.cell_macros()
    assert np.abs(m["U"][:, 0] - Uinf).max() < 0.5 * Uinf and np.isfinite(m["q"]).all()
    dv.close(); orc.close()


@pytest.mark.parametrize("ntheta,nr,nDV", [(24, 8, 29), (24, 8, 41)], ids=["24x8_nc29", "24x8_nc41_chunked"])
def test_hypersonic_cylinder_ogrid(oracle_lib, ntheta, nr, nDV):
    """BASELINE config 5 in small: Ma = 5 past a cylinder on an O-type mesh (curved, non-orthogonal quadrilaterals, the ring
    closed through ordinary internal faces), free-stream "mixed" outer boundary, Maxwell-wall cylinder.
    (On a 16 x 6 mesh the start field - free stream slowed to rest over three radii - is so under-resolved that one
    reconstructed face value is 36 times the cell values around it; a 1.5-ulp perturbation of the initial distribution
    moves the worst entry of gTilde by 8e-14 of the field maximum in the ORACLE ITSELF, and the two implementations sit
    2e-12 apart there, profiles/diag_cyl.py.  The per-step tolerance is a statement about resolved fields.)"""
    case = cs.cylinder_case(ntheta, nr, nDV, perturb=0.01)
    dv = capi.fvDVM(case)
    orc = oracle_lib.Oracle(case)
    dt = case.courant_dt(0.5)
    for step in range(5):
        dv.evolution(dt)
        orc.step(dt)
        _compare(dv, orc, case, util.TOL_STEP * (step + 1), f"cylinder {ntheta}x{nr} NC{nDV} step {step + 1}")
    m = dv.cell_macros()
    assert np.isfinite(m["q"]).all() and m["T"].max() > 1.2 * cs.T0       # the bow shock heats the gas
    dv.close(); orc.close()


@pytest.mark.parametrize("which", ["cylinder_ma5", "cavity3d", "tri2d"])
def test_venkatakrishnan_limited_gradient(oracle_lib, which):
    """dugks_par_t.limiter_k > 0: gradSchemes "VenkatakrishnanLimited leastSquares k" as it is meant to work
    (VenkatakrishnanLimitedGrads.C:59-226; the reference's own implementation is inert, see include/dugks.h), against the
    oracle's restatement of the same scheme; k is chosen so that the limiter bites (eps^2 = k^3 V against squared
    differences of distribution functions of order 1e-24)."""
    case = {"cylinder_ma5": lambda: cs.cylinder_case(24, 8, 29, perturb=0.01),
            "cavity3d": lambda: cs.cavity3d_case(5, 8, distort=0.15, perturb=0.02),
            "tri2d": lambda: cs.tri_cavity_case(8, 8, perturb=0.02)}[which]()
    k = 1e-9
    dv = capi.fvDVM(case, limiter_k=k)
    free = capi.fvDVM(case)
    orc = oracle_lib.Oracle(case)
    orc.set_grad_scheme(1, k)
    dt = case.courant_dt(0.5)
    for step in range(3):
        dv.evolution(dt); free.evolution(dt)
        orc.step(dt)
        _compare(dv, orc, case, util.TOL_STEP * (step + 1), f"limited gradient {which} step {step + 1}")
    # the limiter is active: the limited run differs from the unlimited one far beyond round-off
    assert util.rel_err(dv.cell_macros()["T"], free.cell_macros()["T"]) > 1e-8
    dv.close(); free.close(); orc.close()


def test_symmetry_patch_on_every_axis(oracle_lib):
    """DVMsymmetry / symmetryPlane patches with x- and y-normals on one rank (the sharded variants are in
    tests/test_multi_gpu.py)."""
    K = cs
    for kinds in ({"bottom": K.PATCH_DVM_SYMMETRY, "inlet": K.PATCH_DVM_SYMMETRY}, {"top": K.PATCH_SYMMETRY_PLANE},
                  {"outlet": K.PATCH_DVM_SYMMETRY, "bottom": K.PATCH_SYMMETRY_PLANE}):
        ov = {} if "top" in kinds else {"top": dict(U=(40.0, 0, 0))}
        case = util.channel_case(8, 6, 8, kinds=kinds, bc_overrides=ov, perturb=0.01)
        dv = capi.fvDVM(case)
        orc = oracle_lib.Oracle(case)
        dt = case.courant_dt(0.5)
        for step in range(3):
            dv.evolution(dt)
            orc.step(dt)
            _compare(dv, orc, case, util.TOL_STEP * (step + 1), f"symmetry {sorted(kinds)} step {step + 1}")
        dv.close(); orc.close()


def test_checkpoint_resume_is_bit_exact():
    """dugks_checkpoint_save/load: a fresh handle restored from the blob continues with the same bits (the
    reference's own restart re-initialises the distribution functions, discreteVelocity.C:220-249)."""
    K = cs
    zoo = [cs.cavity3d_case(6, 8, perturb=0.01),
           util.channel_case(10, 6, 8, kinds={"inlet": K.PATCH_FAR_FIELD, "outlet": K.PATCH_ZERO_GRADIENT, "top": K.PATCH_MIXED},
                             bc_overrides={"inlet": dict(U=(30.0, 0, 0), rho=1.2 * K.RHO0, T=290.0, U_bc=K.BC_ZERO_GRADIENT),
                                           "top": dict(U=(20.0, 0, 0), T=280.0)}, perturb=0.01)]
    for case in zoo:
        dt = case.courant_dt(0.5)
        a = capi.fvDVM(case)
        for _ in range(3):
            a.evolution(dt)
        blob = a.checkpoint()
        conv_a0 = a.convergence()
        for _ in range(2):
            a.evolution(dt * 1.1)
        b = capi.fvDVM(case)                       # starts from the t = 0 fields, then takes the blob
        b.restore(blob)
        conv_b0 = b.convergence()
        for _ in range(2):
            b.evolution(dt * 1.1)
        assert conv_a0 == conv_b0
        ma, mb = a.cell_macros(), b.cell_macros()
        for k in ma:
            assert np.array_equal(ma[k], mb[k]), k
        fa, fb = a.face_macros(), b.face_macros()
        for k in fa:
            assert np.array_equal(fa[k], fb[k]), k
        (ga, ha), (gb, hb) = a.state(), b.state()
        assert np.array_equal(ga, gb) and np.array_equal(ha, hb)
        assert all(np.array_equal(x, y) for x, y in zip(a.boundary_surf(), b.boundary_surf()))
        wa, wb = a.wall_diag(), b.wall_diag()
        assert np.array_equal(wa["qWall"], wb["qWall"]) and np.array_equal(a.boundary_macros()["rho"], b.boundary_macros()["rho"])
        assert a.getCoNum(dt) == b.getCoNum(dt)
        # a blob of another case is refused
        other = capi.fvDVM(cs.cavity2d_case(6, 8))
        with pytest.raises(capi.DugksError, match="another case"):
            other.restore(blob)
        other.close(); a.close(); b.close()


def test_scratch_bytes_caps_the_face_storage():
    """dugks_par_t.scratch_bytes bounds the memory taken for kept face values; the result does not depend on it."""
    case = cs.cavity3d_case(6, 28, perturb=0.01)          # 25 slabs of 216 cells
    free = capi.fvDVM(case)
    st = free.stats()
    assert st["keep_slabs"] == st["n_slabs"]
    per_slab = case.geom.nInternalFaces * st["slab_dvs"] * 8
    capped = capi.fvDVM(case, scratch_bytes=int(3.5 * per_slab) + case.nCells * st["slab_dvs"] * 8 * st["n_slabs"])
    sc = capped.stats()
    assert 0 < sc["keep_slabs"] < st["n_slabs"] and sc["device_bytes"] < st["device_bytes"]
    dt = case.courant_dt(0.5)
    for _ in range(2):
        free.evolution(dt); capped.evolution(dt)
    a, b = free.cell_macros(), capped.cell_macros()
    assert util.rel_err(a["rho"], b["rho"]) <= 1e-13 and util.rel_err(a["T"], b["T"]) <= 1e-13
    free.close(); capped.close()
