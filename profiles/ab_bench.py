"""A/B timing of library builds on one GPU (experiment tool, not the bench contract).

    python profiles/ab_bench.py [--workload W] [--steps K] label=LIB[,ENV=VAL...] ...

Each variant runs in its own process (the library path is bound at import).  Prints one line per
variant: label, GUPS, ms per step, per-family ms per step, and a checksum of the cell macros after
the timed steps (variants that compute the same thing agree to round-off in `rho_sum`/`T_sum`).
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def worker(workload, steps):
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import bench
    from dugksfoam_b200 import capi
    kind, kw = bench.WORKLOADS[workload]
    case = bench.build_case(kind, kw)
    dv = capi.fvDVM(case, device=0)
    dt = case.courant_dt(0.8)
    for _ in range(3):
        dv.evolution(dt)
    dv.sync()
    stream = torch.cuda.ExternalStream(dv.stream(), device=torch.device("cuda", 0))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        dv.evolution(dt)
    e1.record(stream)
    dv.sync()
    ms = e0.elapsed_time(e1) / steps
    dv.kernel_timing(1)
    for _ in range(2):
        dv.evolution(dt)
    dv.sync()
    fam = {}
    for which, name in ((0, "outgoing"), (1, "update"), (2, "halfstep")):
        t, n = dv.kernel_timing(-1, which)
        fam[name] = round(t / 2, 2)
    dv.kernel_timing(0)
    cm = dv.cell_macros()
    st = dv.stats()
    print(json.dumps({"gups": round(case.nCells * case.nXi / ms / 1e6, 2), "ms": round(ms, 2), "fam": fam,
                      "rho_sum": float(cm["rho"].sum()), "T_sum": float(cm["T"].sum()),
                      "q_abs": float(np.abs(cm["q"]).sum()), "keep": st.get("keep_slabs", -1),
                      "slabs": st.get("n_slabs", -1)}))
    dv.close()


def main():
    args = sys.argv[1:]
    workload, steps = "cavity3d_64_gh28", 5
    variants = []
    i = 0
    while i < len(args):
        if args[i] == "--workload":
            workload = args[i + 1]; i += 2
        elif args[i] == "--steps":
            steps = int(args[i + 1]); i += 2
        elif args[i] == "--worker":
            return worker(args[i + 1], int(args[i + 2]))
        else:
            variants.append(args[i]); i += 1
    for v in variants:
        label, _, rest = v.partition("=")
        parts = rest.split(",")
        env = dict(os.environ)
        if parts[0]:
            env["DUGKS_LIB"] = os.path.join(ROOT, parts[0]) if not os.path.isabs(parts[0]) else parts[0]
        for kv in parts[1:]:
            k, _, val = kv.partition("=")
            env[k] = val
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--worker", workload, str(steps)], env=env,
                           capture_output=True, text=True, timeout=600)
        out = r.stdout.strip().splitlines()
        print(label, workload, out[-1] if out else ("FAILED rc=%d: %s" % (r.returncode, r.stderr[-800:])), flush=True)


if __name__ == "__main__":
    main()
