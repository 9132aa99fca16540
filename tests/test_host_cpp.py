"""The C++ host mirror of Foam::fvDVM (host/fvDVM.hpp) and the stand-alone driver
(host/dugks_run.cpp, the role of dugksFoam.C's time loop) over the C-ABI."""
import os
import subprocess

import numpy as np
import pytest

from dugksfoam_b200 import capi, dump_case
from dugksfoam_b200 import case as cs
import parity_util as util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def runner(tmp_path_factory):
    capi.build_library()
    exe = str(tmp_path_factory.mktemp("host") / "dugks_run")
    libdir = os.path.join(ROOT, "dugksfoam_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-o", exe, os.path.join(ROOT, "host", "dugks_run.cpp"),
                           "-L" + libdir, "-ldugks", "-Wl,-rpath," + libdir])
    return exe


def test_host_driver_fails_loudly_without_gpu(runner, tmp_path):
    """No CPU fallback: without a device the driver stops with the library's message."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    casef, outf = str(tmp_path / "c.bin"), str(tmp_path / "o.bin")
    dump_case.dump(cs.cavity2d_case(6, 8), casef)
    p = subprocess.run([runner, casef, "1", "0", outf], capture_output=True, text=True)
    assert p.returncode == 2
    assert "no CUDA device" in p.stderr and "no CPU fallback" in p.stderr
    assert not os.path.exists(outf)


@pytest.mark.gpu
def test_host_driver_matches_oracle(runner, tmp_path, oracle_lib):
    case = cs.cavity3d_case(5, 8, perturb=0.01)
    casef, outf = str(tmp_path / "c.bin"), str(tmp_path / "o.bin")
    dump_case.dump(case, casef)
    nsteps = 3
    p = subprocess.run([runner, casef, str(nsteps), "0.5", outf], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    raw = np.fromfile(outf, dtype=np.float64)
    nc = case.nCells
    rho, U, T, q = raw[:nc], raw[nc:4 * nc].reshape(nc, 3), raw[4 * nc:5 * nc], raw[5 * nc:8 * nc].reshape(nc, 3)
    # the same adaptive-dt loop on the oracle (setDeltaTvar.H:34-47: cuts immediate, growth damped to x1.2)
    orc = oracle_lib.Oracle(case)
    dt = case.deltaT or case.courant_dt(0.5)
    for _ in range(nsteps):
        maxCo, _mean = orc.courant(dt)
        fact = 0.5 / (maxCo + 1e-15)
        dt = min(min(fact, 1.0 + 0.1 * fact), 1.2) * dt
        orc.step(dt)
    m = orc.cell_macros()
    sc = util.macro_scales(case)
    tol = util.TOL_STEP * nsteps * 10     # dt itself carries the Courant-number round-off
    assert util.rel_err(rho, m["rho"]) <= tol
    assert util.rel_err(T, m["T"]) <= tol
    assert util.rel_err(U, m["U"], sc["U"]) <= tol
    assert util.rel_err(q, m["q"], sc["q"]) <= tol
    orc.close()
