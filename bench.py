#!/usr/bin/env python
"""bench.py — cell x DV updates/s of the DUGKS discrete-velocity update on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one fvDVM::evolution() (all stages) over the whole mesh and all discrete
velocities.  Workload at every N: BASELINE.json configs[2], the configuration the metric is
quoted on — 3-D cavity 64^3 hexes x 28^3 Gauss-Hermite velocities (h == 0 in 3-D monatomic gas
and is elided, nf = 1) — velocity space sharded over the N GPUs (strong scaling: total work fixed).
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "cell x DV updates/sec (GUPS)"
UNIT = "GUPS"

WORKLOADS = {
    # name: (builder, args)
    "cavity3d_64_gh28": ("cavity3d", dict(n=64, nDV=28)),            # BASELINE configs[2] (headline)
    "cavity2d_256_nc101": ("cavity2d", dict(n=256, nDV=101, quad="NC")),  # BASELINE configs[1]
    "cavity3d_48_gh28": ("cavity3d", dict(n=48, nDV=28)),            # profiling size: every slab keeps its face values, ncu can replay
    "cavity3d_32_gh28": ("cavity3d", dict(n=32, nDV=28)),
    "cavity3d_16_gh16": ("cavity3d", dict(n=16, nDV=16)),
    "cavity2d_60_gh28": ("cavity2d", dict(n=60, nDV=28)),            # demo/cavity shape
    "tri2d_316_gh28": ("tri2d", dict(n=316, nDV=28)),                # ~200k distorted triangles in a square cavity
    # BASELINE configs[3]: micro-channel with a ratchet (saw-tooth) wall, 199,712 triangular prisms, Maxwell walls at three temperatures
    "ratchet_632x158_gh28": ("ratchet", dict(nx=632, ny=158, nDV=28)),
    "poly2d_447_gh28": ("poly2d", dict(n=447, nDV=28)),              # the same size on polygonal (Voronoi) cells: 199,809 cells with 4 to 8 sides
    # BASELINE configs[4]: Ma = 5 past a cylinder, O-type mesh 1000 x 500 quadrilaterals, 81 x 81 Newton-Cotes velocities
    "cylinder_1000x500_nc81": ("cylinder", dict(ntheta=1000, nr=500, nDV=81)),
    "cylinder_200x100_nc81": ("cylinder", dict(ntheta=200, nr=100, nDV=81)),
}
CHECK_STEPS = 3          # the checksum is taken after this many steps (W >= 3 always)
CHECKSUMS = os.path.join(ROOT, "profiles", "bench_checksums.json")
# bounded CPU sample of the same workload shape (3-D cavity, 28^3 GH velocities, same gas/BCs)
CPU_SAMPLE = ("cavity3d", dict(n=16, nDV=28))


def build_case(kind, kw):
    from dugksfoam_b200 import case as cs
    if kind == "cavity3d":
        return cs.cavity3d_case(kw["n"], kw["nDV"])
    if kind == "tri2d":
        return cs.tri_cavity_case(kw["n"], kw["nDV"])
    if kind == "cylinder":
        return cs.cylinder_case(kw["ntheta"], kw["nr"], kw["nDV"])
    if kind == "ratchet":
        return cs.ratchet_channel_case(kw["nx"], kw["ny"], kw["nDV"])
    if kind == "poly2d":
        return cs.poly_cavity_case(kw["n"], kw["nDV"])
    return cs.cavity2d_case(kw["n"], kw["nDV"], quad=kw.get("quad", "GH"))


def b_alg(case, nf):
    """Algorithmic bytes per update, SURVEY.md §8(d): 8 nf (3 + 2F)."""
    return 8.0 * nf * (3.0 + 2.0 * case.faces_per_cell())


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.device)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def _cpu_worker(threads, steps, warmup):
    """One timing of the CPU restatement in its own process (OpenMP reads OMP_NUM_THREADS once per process)."""
    from oracle import oracle as orc
    case = build_case(*CPU_SAMPLE)
    o = orc.Oracle(case)
    dt = case.courant_dt(0.8)
    for _ in range(warmup):
        o.step(dt)
    t0 = time.perf_counter()
    for _ in range(steps):
        o.step(dt)
    el = time.perf_counter() - t0
    print(json.dumps({"gups": case.nCells * o.nxi * steps / el / 1e9, "ms": el / steps * 1e3,
                      "updates": case.nCells * case.nXi}))


def run_cpu_reference(steps, warmup, threads=None, scaling_check=True):
    """Times the CPU restatement of the reference (oracle/: the reference's own loop structure, one discrete
    velocity after the other with OpenMP over the velocities like its -dvParallel ranks, DV sums streamed DV-outer)
    on a bounded sample of the workload, with all host threads; a second, shorter run at half the threads shows
    whether the baseline scales on this host."""
    from oracle import oracle as orc
    orc.build()
    threads = threads or os.cpu_count() or 1

    def timed(nthr, k, w):
        env = dict(os.environ, OMP_NUM_THREADS=str(nthr), OMP_PROC_BIND="spread")
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--cpu-worker", str(nthr), str(k), str(w)],
                             env=env, capture_output=True, text=True, timeout=900)
        return json.loads(out.stdout.strip().splitlines()[-1])
    r = timed(threads, steps, warmup)
    scaling = {str(threads): round(r["gups"], 5)}
    if scaling_check and threads >= 2:
        scaling[str(threads // 2)] = round(timed(threads // 2, 1, 1)["gups"], 5)
    sample = (f"3-D cavity {CPU_SAMPLE[1]['n']}^3 cells x {CPU_SAMPLE[1]['nDV']}^3 GH velocities "
              f"({r['updates']:.3g} updates/step), {steps} steps after {warmup} warm-up, {threads} OpenMP threads; "
              f"GUPS by thread count: {scaling}")
    return r["gups"], r["ms"], threads, sample


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--cpu-worker":
        return _cpu_worker(int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cavity3d_64_gh28", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--record-checksum", action="store_true", help="one GPU: store this run's checksums as the reference")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    kind, kw = WORKLOADS[args.workload]
    if kind == "cylinder":
        wl_name = (f"2-D Ma=5 flow past a cylinder, O-type mesh {kw['ntheta']} x {kw['nr']} quadrilaterals x {kw['nDV']}^2 NC velocities, "
                   "argon Pr=2/3, free-stream (fixedValue -> mixed) outer boundary, Maxwell-wall cylinder")
    elif kind == "ratchet":
        wl_name = (f"2-D micro-channel with a ratchet (saw-tooth) wall, {2 * kw['nx'] * kw['ny']} triangular prisms (unstructured) x "
                   f"{kw['nDV']}^2 GH velocities, Kn=0.075 argon, Maxwell walls at three temperatures")
    else:
        wl_name = f"{'3-D' if kind == 'cavity3d' else '2-D'} cavity {kw['n']}^{3 if kind == 'cavity3d' else 2} " \
                   f"{'triangular-prism (unstructured)' if kind == 'tri2d' else ('polygonal (Voronoi, unstructured)' if kind == 'poly2d' else 'hex')} cells x " \
                  f"{kw['nDV']}^{3 if kind == 'cavity3d' else 2} {kw.get('quad', 'GH')} velocities, Kn=0.075 argon, Maxwell walls"

    if args.impl == "reference":
        # the reference's own CPU path (restated: OpenFOAM + MPI are not available, so the oracle
        # port stands in), rank 0 only
        if rank != 0:
            return
        K = max(1, min(args.steps, 3))
        W = max(1, min(args.warmup, 1))
        gups, ms, cores, sample = run_cpu_reference(K, W)
        line = {
            "impl": "reference", "metric": METRIC, "value": gups, "unit": UNIT, "n_gpus": args.gpus, "steps": K,
            "warmup": W, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl_name, "timed_on": sample, "parallelism": f"openmp{cores}",
                       "same_config": False,
                       "note": "the reference needs OpenFOAM + MPI (absent): oracle/ restates its loops; a bounded sample of the workload"},
            "cpu_baseline": {"value": gups, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": gups, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line), flush=True)
        return

    import torch
    from dugksfoam_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; dugksfoam_b200 has no CPU fallback "
                         "(use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dist = None
    nccl_id = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(idt, 0)
        nccl_id = bytes(idt.cpu().numpy().tobytes())

    case = build_case(kind, kw)
    dv = capi.fvDVM(case, rank=rank, nranks=world, device=local_rank, nccl_id=nccl_id)
    st0 = dv.stats()
    nf = 1 if st0["h_elided"] else 2
    dt = case.courant_dt(0.8)
    updates_per_step = case.nCells * case.nXi
    K, W = args.steps, max(args.warmup, 3)
    stream = torch.cuda.ExternalStream(dv.stream(), device=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- correctness evidence at every N: checksums of the cell macros after exactly CHECK_STEPS steps from the
    # initial state (inside the warm-up, outside every timed region).  The velocity sums commute, so the fields do
    # not depend on how velocity space is split: every N must reproduce the one-GPU values
    # (profiles/bench_checksums.json, written by a one-GPU run with --record-checksum) to 1e-12.
    checksum = None
    for k in range(W):
        dv.evolution(dt)
        if k + 1 == CHECK_STEPS:
            cm3 = dv.cell_macros()
            checksum = {"steps": CHECK_STEPS, "rho_sum": float(cm3["rho"].sum()), "T_sum": float(cm3["T"].sum()),
                        "U_abs_sum": float(np.abs(cm3["U"]).sum()), "q_abs_sum": float(np.abs(cm3["q"]).sum())}
    dv.sync()

    # ---- device-timed region: K steps, state resident in HBM, CUDA events on the library's stream
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = dv.stats()["kernel_launches"]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(K):
        dv.evolution(dt)
    e1.record(stream)
    dv.sync()
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = dv.stats()["kernel_launches"] - l0
    clocks = sampler.stop() if rank == 0 else None
    # per-kernel-family device time: a separate pass with CUDA events around every launch of the
    # three big kernels (kept out of the timed region above)
    dv.kernel_timing(1)
    for _ in range(K):
        dv.evolution(dt)
    dv.sync()
    # device time by timing class of the library (dugks_kernel_timing): 0 = reconstruction / out-flux kernels
    # (k_pencil_phase1, k_hot_outgoing), 1 = relax + update kernels (k_hot_relax_update, k_hot_update), 2 = half step
    tcls = {}
    for which in (0, 1, 2, 3):
        ms, n = dv.kernel_timing(-1, which)
        tcls[which] = (ms / K, n / K)
    dv.kernel_timing(0)
    if dist is not None:
        t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / K
    gups = updates_per_step / (ms_step * 1e-3) / 1e9

    # ---- end-to-end through the host-facing API: every step the host pushes the boundary macro
    # fields (pinned host memory) and reads back the cell macro fields + Courant number, as the
    # reference's time loop does (dugksFoam.C:63-109, CourantNo.H:35)
    bm = dv.boundary_macros()
    pin = {k: torch.from_numpy(v.copy()).pin_memory() for k, v in bm.items()}
    h2d = sum(v.numel() * 8 for v in pin.values())
    d2h = case.nCells * 11 * 8 + 16
    barrier()
    t0 = time.perf_counter()
    U0 = pin["U"].clone()
    for k in range(K):
        # a time-varying wall velocity (1e-9 relative wobble): the boundary fields really change every step,
        # so the H2D copy and the recomputation of the wall in-flux constants are inside the timed region
        torch.mul(U0, 1.0 + 1e-9 * (k + 1), out=pin["U"])
        dv.set_boundary_macros(None, pin["U"].numpy(), pin["T"].numpy())
        dv.evolution(dt)
        if rank == 0:
            # the macro fields go to the host where they are written (dugksFoam.C:63,109: rank 0), into page-locked
            # field storage (dugks_host_register, what the OpenFOAM adapter does with its volFields)
            cm = dv.cell_macros(pinned=True)
        co = dv.getCoNum(dt)
    dv.sync()
    barrier()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_gups = updates_per_step * K / e2e_s / 1e9
    assert co[0] > 0 and (rank != 0 or np.isfinite(cm["rho"]).all())

    if rank != 0:
        dv.close()
        if dist is not None:
            dist.destroy_process_group()
        return

    # partition independence: compare with the one-GPU checksums
    ref_ck, ck_diff = None, None
    try:
        ref_ck = json.load(open(CHECKSUMS)).get(args.workload)
    except Exception:
        ref_ck = None
    if args.record_checksum and world == 1 and checksum is not None:
        allck = json.load(open(CHECKSUMS)) if os.path.exists(CHECKSUMS) else {}
        allck[args.workload] = checksum
        json.dump(allck, open(CHECKSUMS, "w"), indent=1, sort_keys=True)
        ref_ck = checksum
    if ref_ck is not None and checksum is not None and ref_ck.get("steps") == checksum["steps"]:
        # relative to the natural magnitude of each field, as the parity tests do (tests/parity_util.macro_scales): U and q
        # of a cavity at rest are differences of large numbers, their own size is not the scale of their round-off
        cth = float(np.sqrt(2.0 * case.gas["R"] * ref_ck["T_sum"] / case.nCells))
        scale = {"rho_sum": ref_ck["rho_sum"], "T_sum": ref_ck["T_sum"], "U_abs_sum": case.nCells * cth,
                 "q_abs_sum": ref_ck["rho_sum"] * cth ** 3}
        ck_diff = max(abs(checksum[k] - ref_ck[k]) / scale[k] for k in scale)
        assert ck_diff <= 1e-12, f"cell macros after {CHECK_STEPS} steps differ from the one-GPU reference by {ck_diff:.3e}: {checksum} vs {ref_ck}"
    peak, peak_src = peaks()
    balg = b_alg(case, nf)
    achieved = gups / world * balg   # GB/s per GPU (each GPU advances 1/world of the updates) vs one GPU's HBM peak
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "traffic_per_update.json")
    if os.path.exists(tp) and world == 1:
        try:
            t = json.load(open(tp)).get(args.workload)
            traffic, traffic_src = t["dram_gb_per_step"], t["source"]
        except Exception:
            traffic = None
    # The two phases of a step and their algorithmic bytes per update (section 4; they add up to
    # B_alg): phase 1 reads gTilde and writes every face value once, phase 2 reads the face values and gTilde and
    # writes gTilde.  A recompute slab moves more than that (gBarP, the out-flux buffer); the algorithmic figure is the
    # same for every slab.
    F8 = 8.0 * nf * case.faces_per_cell()
    upd_rank = updates_per_step / world
    fam = {
        "phase1: k_pencil_phase1 + k_hot_outgoing<1> + k_hot_halfstep (+ k_hot_outgoing<2> of recompute slabs)":
            {"ms_per_step": tcls[0][0] + tcls[2][0], "launches_per_step": tcls[0][1] + tcls[2][1],
             "alg_bytes_per_update": 8.0 * nf + F8},
        "phase2: k_hot_relax_update (+ k_hot_update of recompute slabs)":
            {"ms_per_step": tcls[1][0], "launches_per_step": tcls[1][1], "alg_bytes_per_update": 16.0 * nf + F8},
    }
    for name, v in fam.items():
        if v["ms_per_step"] > 0:
            v["achieved_gbs"] = upd_rank * v["alg_bytes_per_update"] / (v["ms_per_step"] * 1e-3) / 1e9
            v["frac_of_peak"] = v["achieved_gbs"] / peaks()[0]
    dom = max(fam, key=lambda k: fam[k]["ms_per_step"])
    line = {
        "metric": METRIC, "value": gups, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl_name, "nCells": case.nCells, "nXi": case.nXi, "updates_per_step": updates_per_step,
                   "h_elided": bool(st0["h_elided"]), "nf": nf, "faces_per_cell": case.faces_per_cell(),
                   "parallelism": f"dv{world}", "dt": dt,
                   "l2_policy": "inputs larger than L2 (state is tens of GB per GPU)",
                   "slabs_per_step": st0["n_slabs"], "slab_dvs": st0["slab_dvs"], "face_storage_slabs": st0["keep_slabs"],
                   "device_bytes": st0["device_bytes"]},
        "e2e": {"value": e2e_gups, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_s / K * 1e3},
        "gpu_launches": launches,
        "allreduce": {"ms_per_step": tcls[3][0], "calls_per_step": tcls[3][1],
                      "note": "rank 0, CUDA events around the collectives of a step (waiting for the slowest rank included)"} if world > 1 else None,
        "checksum": checksum, "checksum_rel_diff": ck_diff,
        "checksum_reference": "profiles/bench_checksums.json (one-GPU run)" if ref_ck is not None else None,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_unit": "GB of DRAM reads+writes per step (ncu)", "traffic_source": traffic_src,
                     "peak_source": peak_src,
                     "scope": "whole step (all kernels of one evolution()), per GPU",
                     "algorithmic_bytes_per_update": balg,
                     "dominant_kernel": dom, "kernels": fam},
    }
    if not args.no_cpu_baseline and world == 1:
        cg, cms, cores, sample = run_cpu_reference(min(args.cpu_steps, 2), 1)
        line["cpu_baseline"] = {"value": cg, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                                "ms_per_step": cms}
    print(json.dumps(line), flush=True)
    dv.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
