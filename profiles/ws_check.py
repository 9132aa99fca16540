"""Device check of the warp-specialised pencil kernel (DUGKS_PENCIL_WS=1) against the oracle, 3-D cavities (dev builds too)."""
import os, sys, time
os.environ["DUGKS_PENCIL_WS"] = os.environ.get("DUGKS_PENCIL_WS", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from dugksfoam_b200 import capi, case as cs
from oracle import oracle as orc_mod
import parity_util as util
t0 = time.time()
for name, case in (("cavity3d_10_gh8", cs.cavity3d_case(10, 8, perturb=0.01)),
                   ("cavity3d_11_gh28", cs.cavity3d_case(11, 28, perturb=0.01)),
                   ("cavity3d_20_gh8", cs.cavity3d_case(20, 8, perturb=0.01))):
    dv, orc = capi.fvDVM(case), orc_mod.Oracle(case)
    dt = case.courant_dt(0.5)
    worst = 0.0
    for step in range(3):
        dv.evolution(dt); orc.step(dt)
        a, b = dv.cell_macros(), orc.cell_macros()
        g, _ = dv.state(); go, _ = orc.state()
        e = max(util.rel_err(a["rho"], b["rho"]), util.rel_err(a["T"], b["T"]), util.rel_err(g, go[dv.local_dvs()]))
        worst = max(worst, e)
    st = dv.stats()
    print(f"WS_CHECK {name}: max rel err {worst:.3e} pencil_mode {st['pencil_mode']} cells {st['pencil_cells']} t {time.time() - t0:.1f}s", flush=True)
    dv.close(); orc.close()
