"""Writes a Case as the flat file host/dugks_run.cpp reads: per array one text line
`name dtype nbytes` followed by the raw little-endian bytes.

    python -m dugksfoam_b200.dump_case <reference case dir | cavity2d:N:nDV | cavity3d:N:nDV> out.bin
"""
from __future__ import annotations

import sys

import numpy as np

from . import case as cs


def dump(case: cs.Case, path: str) -> None:
    g = case.geom
    gas = case.gas
    arrays = [
        ("sizes", np.array([g.nCells, g.nInternalFaces, g.nBoundaryFaces, g.nSolutionD, int(gas.get("KInner", 0))], np.int32)),
        ("owner", np.asarray(g.owner, np.int32)), ("neighbour", np.asarray(g.neighbour, np.int32)),
        ("C", g.C), ("V", g.V), ("Cf", g.Cf), ("Sf", g.Sf), ("ownLs", g.ownLs), ("neiLs", g.neiLs),
        ("patchLs", g.patchLs), ("deltaCoeffs", g.deltaCoeffs), ("Xis", case.Xis), ("weights", case.weights),
        ("scalars", np.array([case.xiMax, case.xiMin, gas["R"], gas["omega"], gas["Tref"], gas["muRef"], gas["Pr"],
                              case.deltaT or case.courant_dt(0.5)], np.float64)),
        ("patches", np.array([[p.kind, p.start, p.size, p.U_bc, p.T_bc] for p in case.patches], np.int32).reshape(-1)),
        ("patch_pressure", np.array([p.pressure for p in case.patches], np.float64)),
        ("rho", case.rho), ("U", case.U), ("T", case.T), ("rho_b", case.rho_b), ("U_b", case.U_b), ("T_b", case.T_b),
    ]
    with open(path, "wb") as f:
        for name, a in arrays:
            a = np.ascontiguousarray(a, dtype=np.int32 if a.dtype.kind == "i" else np.float64)
            f.write(f"{name} {a.dtype.name} {a.nbytes}\n".encode())
            f.write(a.tobytes())
            f.write(b"\n")


def main(argv):
    spec, out = argv[1], argv[2]
    if spec.startswith("cavity2d:"):
        _, n, ndv = spec.split(":")
        c = cs.cavity2d_case(int(n), int(ndv))
    elif spec.startswith("cavity3d:"):
        _, n, ndv = spec.split(":")
        c = cs.cavity3d_case(int(n), int(ndv))
    else:
        c = cs.read_case(spec)
    dump(c, out)


if __name__ == "__main__":
    main(sys.argv)
