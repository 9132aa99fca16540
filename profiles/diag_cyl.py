"""Where does the cylinder 16x6 NC41 case differ from the oracle after one step?  (diagnostic, not a test)"""
import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
from dugksfoam_b200 import capi, case as cs
from oracle import oracle as om
import parity_util as util
for args, env in (((16, 6, 41), {}), ((16, 6, 41), {"DUGKS_NO_HOT": "1"}), ((16, 6, 29), {}), ((24, 8, 41), {}), ((16, 6, 41), {"DUGKS_KEEP_SLABS": "0"})):
    for k in ("DUGKS_NO_HOT", "DUGKS_KEEP_SLABS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    case = cs.cylinder_case(*args, perturb=0.01)
    dv = capi.fvDVM(case); orc = om.Oracle(case)
    dt = case.courant_dt(0.5)
    dv.evolution(dt); orc.step(dt)
    g, h = dv.state(); go, ho = orc.state()
    ids = dv.local_dvs()
    err = np.abs(g - go[ids])
    k, c = np.unravel_index(err.argmax(), err.shape)
    n = case.nXiPerDim
    print(args, env, "gTilde rel", err.max() / np.abs(go).max(), "at dv", k, "(ix, iy) =", (k % n, k // n), "xi =", (case.Xis[k % n], case.Xis[k // n]),
          "cell", c, "(i, j) =", (c % args[0], c // args[0]), "g there", go[k, c], "rel local", err[k, c] / abs(go[k, c]), "max g", np.abs(go).max(), flush=True)
    # distribution of errors by DV row
    rel = err.max(axis=1) / np.abs(go).max()
    bad = np.where(rel > 2e-13)[0]
    print("   DVs above 2e-13:", len(bad), [(int(b % n), int(b // n)) for b in bad[:12]])
    gb, hb = dv.boundary_surf()
    nif = case.geom.nInternalFaces
    ref = np.stack([orc.surf(int(kk))[0][nif:] for kk in ids])
    eb = np.abs(gb - ref)
    kb, b = np.unravel_index(eb.argmax(), eb.shape)
    print("   boundary gSurf rel", eb.max() / np.abs(ref).max(), "at dv", (kb % n, kb // n), "bface", b)
    dv.close(); orc.close()
