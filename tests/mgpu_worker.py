"""Worker for the multi-GPU parity test: run under torchrun, one rank per GPU.
Velocity space is sharded over the ranks (the reference's -dvParallel); the moment sums are
all-reduced by the library's own NCCL communicator (mode "nccl") or by a torch.distributed
callback (mode "callback", the fieldMPIreducer role).  Rank 0 compares with the CPU oracle."""
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


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "nccl"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cases = [("cavity3d_6_gh8", cs.cavity3d_case(6, 8, perturb=0.01)),
             ("cavity2d_12_gh8", cs.cavity2d_case(12, 8, perturb=0.01)),
             ("cavity2d_9_nc9", cs.cavity2d_case(9, 9, quad="NC", perturb=0.01))]
    ok = True
    for name, case in cases:
        kw = {}
        if mode == "nccl":
            idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                idt = torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8).cuda()
            dist.broadcast(idt, 0)
            kw["nccl_id"] = bytes(idt.cpu().numpy().tobytes())
        else:
            keep = []

            def reduce(ptr, n, stream):
                # wrap the device buffer without copying and all-reduce it in place on the library's stream
                class _Arr:
                    __cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3}
                t = torch.as_tensor(_Arr(), device="cuda")
                ext = torch.cuda.ExternalStream(stream)
                with torch.cuda.stream(ext):
                    dist.all_reduce(t)
                keep.append(t)
                return 0
            kw["reduce"] = reduce
        dv = capi.fvDVM(case, rank=rank, nranks=world, device=local, **kw)
        ids = dv.local_dvs()
        assert np.array_equal(ids, capi.partition(case.nXiPerDim, case.geom.nSolutionD, world, rank))
        dt = case.courant_dt(0.5)
        for _ in range(3):
            dv.evolution(dt)
        dv.sync()
        cm = dv.cell_macros()
        g, _ = dv.state()
        gdf, _ = dv.writeDFonCell(1)
        if rank == 0:
            from oracle import oracle as orc
            o = orc.Oracle(case)
            for _ in range(3):
                o.step(dt)
            om = o.cell_macros()
            go, _ = o.state()
            sc = util.macro_scales(case)
            errs = dict(rho=util.rel_err(cm["rho"], om["rho"]), T=util.rel_err(cm["T"], om["T"]),
                        U=util.rel_err(cm["U"], om["U"], sc["U"]), q=util.rel_err(cm["q"], om["q"], sc["q"]),
                        g=util.rel_err(g, go[ids]), df=util.rel_err(gdf, go[:, 1]))
            good = all(v <= 3e-12 for v in errs.values())
            ok &= good
            print(f"MGPU {mode} {name} world={world} {'OK' if good else 'FAIL'} {errs}", flush=True)
        dv.close()
        dist.barrier()
    if rank == 0:
        print("MGPU_RESULT", "PASS" if ok else "FAIL", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
