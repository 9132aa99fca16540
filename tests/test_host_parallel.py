"""CPU tests of the host-side pieces of the velocity-space decomposition: the library's
partition (dugks_partition), and a world_size-2 gloo run in which every rank forms the moment
sums of its own discrete velocities and all-reduces them — the exchange step of the path
(reference: fieldMPIreducer::reduceField, fieldMPIreducer.C:48-150)."""
import os
import subprocess
import sys
import textwrap

import numpy as np

from dugksfoam_b200 import capi

HERE = os.path.dirname(os.path.abspath(__file__))


def test_partition_is_disjoint_and_complete():
    for n, D, P in ((28, 3, 8), (28, 2, 4), (101, 2, 8), (9, 1, 1), (8, 3, 3)):
        ids = [capi.partition(n, D, P, r) for r in range(P)]
        allids = np.concatenate(ids)
        assert len(allids) == n ** D and np.array_equal(np.sort(allids), np.arange(n ** D))
        sizes = [len(i) for i in ids]
        assert max(sizes) - min(sizes) <= n          # balanced to one velocity row
        for i in ids:                                  # whole rows (ix fastest) stay on one rank
            assert len(i) % n == 0 and np.all(np.diff(i) == 1)


def test_partition_rejects_bad_input():
    import pytest
    with pytest.raises(capi.DugksError):
        capi.partition(28, 1, 2, 0)                   # one velocity row cannot be split
    with pytest.raises(capi.DugksError):
        capi.partition(28, 3, 4, 4)


WORKER = textwrap.dedent('''
    import os, sys
    sys.path.insert(0, {root!r}); sys.path.insert(0, {here!r})
    import numpy as np, torch, torch.distributed as dist
    from dugksfoam_b200 import capi, case as cs
    from oracle import oracle as orc
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    case = cs.cavity2d_case(8, 8, perturb=0.01)
    o = orc.Oracle(case)
    dt = case.courant_dt(0.5)
    for _ in range(2): o.step(dt)
    g, h = o.state(); xi, w, *_ = o.dvs()
    ids = capi.partition(case.nXiPerDim, case.geom.nSolutionD, world, rank)
    # local moment sums (fvDVM.C:612-622) then the all-reduce (fvDVM.C:626-628)
    part = np.stack([(w[ids, None] * g[ids]).sum(0)] + [(w[ids, None] * g[ids] * xi[ids, d:d+1]).sum(0) for d in range(3)]
                    + [0.5 * (w[ids, None] * (g[ids] * (xi[ids] ** 2).sum(1)[:, None] + h[ids])).sum(0)])
    t = torch.from_numpy(part.copy()); dist.all_reduce(t)
    tot = t.numpy(); m = o.cell_macros()
    rho = tot[0]; U = tot[1:4].T / rho[:, None]
    T = (tot[4] - 0.5 * rho * (U ** 2).sum(1)) / (1.5 * case.gas["R"] * rho)
    ok = (np.abs(rho - m["rho"]).max() / m["rho"].max() < 1e-13 and np.abs(T - m["T"]).max() / m["T"].max() < 1e-13
          and np.abs(U - m["U"]).max() < 1e-10)
    print("GLOO_RANK", rank, "PASS" if ok else "FAIL", flush=True)
    dist.destroy_process_group()
''')


def test_gloo_two_rank_moment_allreduce(tmp_path, oracle_lib):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=os.path.dirname(HERE), here=HERE))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29633", str(script)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.stdout.count("PASS") == 2, out.stdout[-2000:] + out.stderr[-2000:]
