"""Where the host-side time of an end-to-end step goes (one GPU; 32^3 case so that the device part is short)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench
from dugksfoam_b200 import capi
wl = sys.argv[1] if len(sys.argv) > 1 else "cavity3d_64_gh28"
case = bench.build_case(*bench.WORKLOADS[wl])
dv = capi.fvDVM(case, device=0)
dt = case.courant_dt(0.8)
for _ in range(3):
    dv.evolution(dt)
dv.sync()
dv.cell_macros(pinned=True)   # page-locks the arrays once
bm = dv.boundary_macros()
pin = {k: torch.from_numpy(v.copy()).pin_memory() for k, v in bm.items()}
U0 = pin["U"].clone()
acc = {"set_bmac": 0.0, "evolution_enqueue": 0.0, "sync_after_step": 0.0, "cell_macros": 0.0, "courant": 0.0, "cell_macros_pinned": 0.0}
K = 5
for k in range(K):
    torch.mul(U0, 1.0 + 1e-9 * (k + 1), out=pin["U"])
    t0 = time.perf_counter(); dv.set_boundary_macros(None, pin["U"].numpy(), pin["T"].numpy()); t1 = time.perf_counter()
    dv.evolution(dt); t2 = time.perf_counter()
    dv.sync(); t3 = time.perf_counter()
    cm = dv.cell_macros(); t4 = time.perf_counter()
    co = dv.getCoNum(dt); t5 = time.perf_counter()
    cp = dv.cell_macros(pinned=True); t6 = time.perf_counter()
    assert all(np.array_equal(cm[k], cp[k]) for k in cm)
    for name, d in zip(acc, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5)):
        acc[name] += d
print("E2E_PROBE", wl, {k: round(v / K * 1e3, 3) for k, v in acc.items()}, "ms per step")
dv.close()
