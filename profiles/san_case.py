"""Small cases through every hot kernel family, for compute-sanitizer (memcheck / racecheck / synccheck):
CTA pencils with the fused half step (3-D, 16^3: 14^3 interior cells in 49 bundles), the axis-only and general launches,
relax+update in all three cell-stream modes, the recompute path, 2-D with h, chunked rows, triangular prisms."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from dugksfoam_b200 import capi, case as cs

runs = [("cavity3d_16_gh8 pencils fused", cs.cavity3d_case(16, 8, perturb=0.01), {}),
        ("cavity3d_16_gh8 pencils unfused, one face-storage slab", cs.cavity3d_case(16, 8, perturb=0.01), {"DUGKS_PENCIL": "1", "DUGKS_KEEP_SLABS": "1"}),
        ("cavity3d_13 (odd interior: leftover lines)", cs.cavity3d_case(13, 8, perturb=0.01), {}),
        ("cavity2d_24_gh28 axis-only launch with h", cs.cavity2d_case(24, 28, perturb=0.01), {}),
        ("cavity2d_10_nc37 chunked rows", cs.cavity2d_case(10, 37, quad="NC", perturb=0.01), {}),
        ("tri_8_gh8 general path", cs.tri_cavity_case(8, 8, perturb=0.01), {}),
        ("cavity3d_16_gh8 warp-specialised pencils (producer / consumer warps, mbarrier ring)", cs.cavity3d_case(16, 8, perturb=0.01), {"DUGKS_PENCIL_WS": "1"}),
        ("cavity3d_12_gh8 recompute path (lagged boundary gradient written in place in phase 2)", cs.cavity3d_case(12, 8, perturb=0.01), {"DUGKS_KEEP_SLABS": "0"}),
        ("ratchet_20x8_gh8 saw-tooth channel", cs.ratchet_channel_case(20, 8, 8, teeth=2, perturb=0.01), {})]
for name, case, env in runs:
    for k in ("DUGKS_PENCIL", "DUGKS_KEEP_SLABS", "DUGKS_PENCIL_WS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    dv = capi.fvDVM(case)
    dt = case.courant_dt(0.5)
    for _ in range(2):
        dv.evolution(dt)
    dv.sync()
    m = dv.cell_macros()
    st = dv.stats()
    print(f"{name}: rho_sum {m['rho'].sum():.15e} T_sum {m['T'].sum():.15e} pencil cells {st['pencil_cells']} mode {st['pencil_mode']} "
          f"keep {st['keep_slabs']}/{st['n_slabs']} finite {bool(np.isfinite(m['q']).all())}", flush=True)
    dv.close()
