"""Worker of the sharded-parity tests: run under torchrun, one process per rank.
Velocity space is sharded over the ranks (the reference's -dvParallel, fvDVM.C:228-260); the moment sums
are all-reduced by
  nccl     : the library's own NCCL communicator (one GPU per rank),
  callback : a torch.distributed NCCL all-reduce behind dugks_par_t.reduce (the fieldMPIreducer role),
  gloo     : the same callback staged through host memory and a gloo all-reduce — the ranks may then SHARE
             one GPU, which is how the driver's one-GPU `pytest -m gpu` box runs the sharded path.
Rank 0 compares with the CPU oracle; every rank checks that its slice of the state matches."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import parity_util as util  # noqa: E402
from dugksfoam_b200 import capi  # noqa: E402
from dugksfoam_b200 import case as cs  # noqa: E402


def cases(world):
    K = cs
    out = [("cavity3d_6_gh8", cs.cavity3d_case(6, 8, perturb=0.01)),
           ("cavity2d_12_gh8", cs.cavity2d_case(12, 8, perturb=0.01)),
           ("cavity2d_9_nc9_ties", cs.cavity2d_case(9, 9, quad="NC", perturb=0.01)),
           # rows cut into ix-chunks with fewer than 32 rows per rank (BASELINE configs 2 and 5 on 4-8 GPUs)
           ("cavity2d_8_nc37_chunked", cs.cavity2d_case(8, 37, quad="NC", perturb=0.01)),
           # y-normal symmetry patches: mirror partners live on another rank (fvDVM.C:375-454 exchanges them)
           ("sym_y_dvm", util.channel_case(8, 6, 8, kinds={"bottom": K.PATCH_DVM_SYMMETRY, "inlet": K.PATCH_DVM_SYMMETRY},
                                           bc_overrides={"top": dict(U=(40.0, 0, 0))}, perturb=0.01)),
           ("sym_y_plane", util.channel_case(8, 6, 8, kinds={"bottom": K.PATCH_SYMMETRY_PLANE},
                                             bc_overrides={"top": dict(U=(40.0, 0, 0))}, perturb=0.01))]
    return [c for c in out if c[1].nXiPerDim ** (c[1].geom.nSolutionD - 1) >= world]


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "nccl"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    ndev = torch.cuda.device_count()
    dev = local % ndev
    torch.cuda.set_device(dev)
    if mode == "gloo":
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
    ok = True
    for name, case in cases(world):
        kw = {}
        keep = []
        if mode == "nccl":
            idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                idt = torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8).cuda()
            dist.broadcast(idt, 0)
            kw["nccl_id"] = bytes(idt.cpu().numpy().tobytes())
        else:
            def reduce(ptr, n, stream):
                # wrap the device buffer without copying; sum it over the ranks in place, ordered on the library's stream
                class _Arr:
                    __cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3}
                t = torch.as_tensor(_Arr(), device="cuda")
                ext = torch.cuda.ExternalStream(stream)
                with torch.cuda.stream(ext):
                    if mode == "gloo":
                        hbuf = t.cpu()              # waits for the kernels that produced the sums
                        dist.all_reduce(hbuf)
                        t.copy_(hbuf)
                    else:
                        dist.all_reduce(t)
                keep.append(t)
                return 0
            kw["reduce"] = reduce
        dv = capi.fvDVM(case, rank=rank, nranks=world, device=dev, **kw)
        ids = dv.local_dvs()
        assert np.array_equal(ids, capi.partition(case.nXiPerDim, case.geom.nSolutionD, world, rank))
        dt = case.courant_dt(0.5)
        nsteps = 3
        for _ in range(nsteps):
            dv.evolution(dt)
        dv.sync()
        cm, fm, bm = dv.cell_macros(), dv.face_macros(), dv.boundary_macros()
        g, h = dv.state()
        gdf, _ = dv.writeDFonCell(1)
        from oracle import oracle as orc
        o = orc.Oracle(case)
        for _ in range(nsteps):
            o.step(dt)
        om, of, ob = o.cell_macros(), o.face_macros(), o.boundary_macros()
        go, ho = o.state()
        sc = util.macro_scales(case)
        errs = dict(rho=util.rel_err(cm["rho"], om["rho"]), T=util.rel_err(cm["T"], om["T"]),
                    U=util.rel_err(cm["U"], om["U"], sc["U"]), q=util.rel_err(cm["q"], om["q"], sc["q"]),
                    frho=util.rel_err(fm["rho"], of["rho"]), fq=util.rel_err(fm["q"], of["q"], sc["q"]),
                    brho=util.rel_err(bm["rho"], ob["rho"], sc["rho"]),
                    g=util.rel_err(g, go[ids]), h=util.rel_err(h, ho[ids], max(np.abs(ho).max(), 1e-300)),
                    df=util.rel_err(gdf, go[:, 1]))
        good = all(v <= util.TOL_STEP * nsteps for v in errs.values())
        flag = torch.tensor([1 if good else 0], dtype=torch.int32, device=None if mode == "gloo" else "cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok &= bool(flag.item())
        if rank == 0 or not good:
            print(f"MGPU {mode} {name} world={world} rank={rank} dvs={len(ids)} {'OK' if good else 'FAIL'} "
                  f"{ {k: float('%.2e' % v) for k, v in errs.items()} }", flush=True)
        dv.close()
        o.close()
        dist.barrier()
    if rank == 0:
        print("MGPU_RESULT", "PASS" if ok else "FAIL", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
