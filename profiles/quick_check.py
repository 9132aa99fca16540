"""Last-minute device check of two paths added after the GPU budget was nearly spent (no torch, no pytest)."""
import sys, time
t0 = time.time()
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
from dugksfoam_b200 import capi, case as cs
from oracle import oracle as orc_mod
import parity_util as util
case = cs.cavity3d_case(6, 8, perturb=0.01)
dv, orc = capi.fvDVM(case), orc_mod.Oracle(case)
dt = case.courant_dt(0.5)
for _ in range(2):
    dv.evolution(dt); orc.step(dt)
print("convergence gpu", dv.convergence(), "oracle", orc.convergence(), "again", dv.convergence(), flush=True)
dv.close(); orc.close()
case = cs.cavity3d_case(10, 8, perturb=0.01)
dv, orc = capi.fvDVM(case, store_h=True), orc_mod.Oracle(case)
for _ in range(2):
    dv.evolution(dt); orc.step(dt)
a, b = dv.cell_macros(), orc.cell_macros()
g, h = dv.state(); go, ho = orc.state()
print("storeh 10^3: rho", util.rel_err(a["rho"], b["rho"]), "T", util.rel_err(a["T"], b["T"]),
      "g", util.rel_err(g, go[dv.local_dvs()]), "h", util.rel_err(h, ho[dv.local_dvs()], np.abs(go).max()),
      "stats", dv.stats()["keep_slabs"], "t", round(time.time() - t0, 1), flush=True)
