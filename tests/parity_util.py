"""Shared helpers of the parity tests: case zoo and comparison metrics."""
import numpy as np

from dugksfoam_b200 import case as cs
from dugksfoam_b200 import dvset
from dugksfoam_b200.polymesh import hex_block

# tolerance of BASELINE.json north_star: 1e-12 relative per step (FP64)
TOL_STEP = 1e-12
# 1e-9 on rho/U/T/q after 1000 steps
TOL_LONG = 1e-9


def rel_err(a, b, scale=None):
    """max |a-b| / max |b| (or / scale): field-wise relative L-infinity error."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    s = np.abs(b).max() if scale is None else scale
    if s == 0.0:
        return float(np.abs(a - b).max())
    return float(np.abs(a - b).max() / s)


def macro_scales(case):
    """Characteristic magnitudes used for the vector fields whose reference value may be ~0."""
    R = case.gas["R"]
    T = float(np.max(case.T))
    rho = float(np.max(case.rho))
    c = np.sqrt(2 * R * T)
    return dict(U=c, q=rho * c ** 3, rho=rho, T=T)


def channel_case(nx=10, ny=6, nDV=8, kinds=None, **kw):
    """2-D channel: xmin = inlet patch, xmax = outlet patch, ymin/ymax walls; used for the
    far-field / mixed / zeroGradient / pressure / symmetry boundary kinds."""
    names = {"xmin": "inlet", "xmax": "outlet", "ymin": "bottom", "ymax": "top"}
    mesh = hex_block(nx, ny, 1, (1.0, 0.6, 0.1), two_d=True, patch_names=names, distort=kw.pop("distort", 0.0))
    Xis, w = cs.gh_set(nDV)
    kinds = kinds or {}
    c = cs._uniform_case(mesh, Xis, w, kinds, lid_patch="none", name="channel", **kw)
    return c


def cases_small():
    """(name, case, store_h) tuples the oracle finishes in well under a second per step."""
    out = []
    out.append(("cavity2d_12_gh8", cs.cavity2d_case(12, 8, perturb=0.01), False))
    out.append(("cavity2d_16_gh28", cs.cavity2d_case(16, 28), False))
    out.append(("cavity2d_10_gh8_distort", cs.cavity2d_case(10, 8, distort=0.2, perturb=0.01), False))
    out.append(("cavity2d_9_nc9_ties", cs.cavity2d_case(9, 9, quad="NC", perturb=0.01), False))
    out.append(("cavity3d_6_gh8", cs.cavity3d_case(6, 8, perturb=0.01), False))
    out.append(("cavity3d_5_gh8_storeh", cs.cavity3d_case(5, 8, perturb=0.01), True))
    out.append(("cavity3d_5_gh8_distort", cs.cavity3d_case(5, 8, distort=0.15, perturb=0.01), False))
    out.append(("tri_8_gh8", cs.tri_cavity_case(8, 8, perturb=0.01), False))
    out.append(("poly_7_gh8", cs.poly_cavity_case(7, 8, perturb=0.01), False))            # config 4: polygonal (Voronoi) cells, 4 to 8 sides
    out.append(("ratchet_20x8_gh8", cs.ratchet_channel_case(20, 8, 8, teeth=2, perturb=0.01), False))   # config 4: saw-tooth wall, three wall temperatures
    return out


def demo_cavity_case():
    """The reference's shipped demo/cavity (BASELINE config 1) from tests/golden/demo_cavity_case.npz (made by
    tests/golden/make_golden.py from /root/reference/demo/cavity): 60 x 60 cells in 4 blockMesh blocks, the
    shipped 28-point Gauss-Hermite set, argon at Kn = 0.075, lid at 50 m/s."""
    import json
    import os
    from dugksfoam_b200.polymesh import Patch, PolyMesh, compute_geometry
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "demo_cavity_case.npz"))
    mesh = PolyMesh(points=z["points"], face_verts=z["face_verts"], face_offsets=z["face_offsets"], owner=z["owner"],
                    neighbour=z["neighbour"], patches=[Patch(n, t, k, s) for n, t, k, s in json.loads(str(z["patch_table"]))])
    geom = compute_geometry(mesh)
    patches = [cs.PatchSpec(n, k, st, sz, ub, tb, pr) for n, k, st, sz, ub, tb, pr in json.loads(str(z["case_patches"]))]
    return cs.Case(geom=geom, patches=patches, Xis=z["Xis"], weights=z["weights"], xiMax=float(z["xiMax"]),
                   xiMin=float(z["xiMin"]), gas=json.loads(str(z["gas"])), rho=z["rho"], U=z["U"], T=z["T"],
                   rho_b=z["rho_b"], U_b=z["U_b"], T_b=z["T_b"], deltaT=float(z["deltaT"]), maxCo=float(z["maxCo"]),
                   name="demo/cavity (shipped)", mesh=mesh)
