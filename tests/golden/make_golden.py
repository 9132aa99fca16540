"""Generates the golden fixtures under tests/golden/ from the reference tree.

Run in the build container (where /root/reference exists):
    python tests/golden/make_golden.py

demo_cavity_dvset.json : the reference's only golden vector on the hot path —
    demo/cavity/constant/{Xis,weights} (28-point half-range Gauss-Hermite set, also
    printed in doc/usage.tex:114-148) — plus the mesh / case facts of demo/cavity
    (polyMesh/owner header note, polyMesh/boundary, DVMProperties, 0/*).
demo_cavity_case.npz : the shipped demo/cavity case as data (mesh with its 4-block numbering, quadrature,
    gas, fields), loaded by tests/parity_util.demo_cavity_case() where the reference tree is not mounted.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

import numpy as np  # noqa: E402

from dugksfoam_b200 import foam  # noqa: E402
from dugksfoam_b200.case import read_case  # noqa: E402

REF = "/root/reference/demo/cavity"


def main():
    Xis = foam.read_scalar_list(os.path.join(REF, "constant", "Xis"))
    w = foam.read_scalar_list(os.path.join(REF, "constant", "weights"))
    c = read_case(REF)
    g = c.geom
    out = dict(
        source="zhulianhua/dugksFoam demo/cavity/constant/{Xis,weights}, polyMesh/*, DVMProperties, 0/*",
        Xis=[float(v) for v in Xis], weights=[float(v) for v in w],
        nCells=g.nCells, nInternalFaces=g.nInternalFaces, nBoundaryFaces=g.nBoundaryFaces,
        nSolutionD=g.nSolutionD, patch_names=g.patch_names, patch_sizes=g.patch_size,
        nPoints=int(len(c.mesh.points)), nFacesAll=int(c.mesh.nFaces),
        gas=c.gas, rho0=float(c.rho[0]), T0=float(c.T[0]),
        lid_U=[float(v) for v in c.U_b[0]], deltaT=c.deltaT, maxCo=c.maxCo,
        xiMax=c.xiMax, xiMin=c.xiMin,
        # geometry checksums of the shipped mesh under the [OF-lib] formulas
        V_sum=float(g.V.sum()), V_min=float(g.V.min()), V_max=float(g.V.max()),
        magSf_sum=float(np.sqrt((g.Sf ** 2).sum(axis=1)).sum()),
    )
    with open(os.path.join(HERE, "demo_cavity.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote demo_cavity.json")
    # The shipped case itself, as data (the GPU box has no /root/reference): polyMesh exactly as shipped (point
    # coordinates, face vertex lists, owner / neighbour with the 4-block blockMesh numbering, patch table), the
    # quadrature files, DVMProperties values and the initial / boundary fields as read_case() parses them.
    m = c.mesh
    np.savez_compressed(
        os.path.join(HERE, "demo_cavity_case.npz"),
        points=m.points, face_verts=m.face_verts.astype(np.int32), face_offsets=m.face_offsets.astype(np.int32),
        owner=m.owner, neighbour=m.neighbour,
        patch_table=json.dumps([[p.name, p.type, int(p.nFaces), int(p.startFace)] for p in m.patches]),
        case_patches=json.dumps([[p.name, p.kind, p.start, p.size, p.U_bc, p.T_bc, p.pressure] for p in c.patches]),
        Xis=c.Xis, weights=c.weights, xiMax=c.xiMax, xiMin=c.xiMin, gas=json.dumps(c.gas),
        rho=c.rho, U=c.U, T=c.T, rho_b=c.rho_b, U_b=c.U_b, T_b=c.T_b, deltaT=c.deltaT, maxCo=c.maxCo)
    print("wrote demo_cavity_case.npz", os.path.getsize(os.path.join(HERE, "demo_cavity_case.npz")), "bytes")


if __name__ == "__main__":
    main()
