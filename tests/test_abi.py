"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports every
symbol include/dugks.h declares, and fails loudly (no CPU fallback) when asked to compute."""
import ctypes
import os
import re

import numpy as np
import pytest

from dugksfoam_b200 import capi
from dugksfoam_b200 import case as cs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "dugks.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dugks_[a-z_0-9]+)\s*\(", text)) - {"dugks_allreduce_fn"})


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(capi.LIB_PATH)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"libdugks.so does not export {n}"
    assert sorted(capi.EXPORTS) == names
    assert lib.dugks_abi_version() == 2


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(capi.DugksError, match="no CUDA device"):
        capi.fvDVM(cs.cavity2d_case(4, 8))


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "dugksfoam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "libdugks_oracle" not in src, f


def test_struct_layouts_match_header():
    from dugksfoam_b200 import abi
    assert ctypes.sizeof(abi.PatchT) == 32
    assert ctypes.sizeof(abi.GasT) == 48
    assert ctypes.sizeof(abi.DvsetT) == 40
    assert ctypes.sizeof(abi.MeshT) == 16 + 10 * 8
    assert ctypes.sizeof(abi.StatsT) == 48


def test_row_layout_covers_every_velocity_once():
    """Host logic of the velocity-row layout (rows, slabs, short-row tail slab): over all ranks every
    discrete velocity is in exactly one row, rows fit the slab rules, short rows only in the last slab."""
    from dugksfoam_b200 import capi
    import numpy as np
    for D in (1, 2, 3):
        for n in (2, 3, 8, 9, 16, 28, 32, 33, 101):
            if D == 3 and n > 33:
                continue
            ny, nz = (n if D >= 2 else 1), (n if D == 3 else 1)
            for nranks in (1, 2, 3, 4, 8):
                if ny * nz < nranks:
                    continue
                seen = np.zeros((nz, ny, n), dtype=np.int32)
                for rank in range(nranks):
                    lay = capi.row_layout(n, D, nranks, rank)
                    L, Lt, nch = lay["L"], lay["Lt"], lay["nch"]
                    assert 1 <= L <= 32 and 0 <= Lt < max(L, 1) + (Lt == 0)
                    assert nch * L >= n and (nch - 1) * L < n
                    nrows = len(lay["iy"])
                    nslab = (nrows + 31) // 32
                    for k in range(nrows):
                        ln, first = int(lay["len"][k]), int(lay["first"][k])
                        if ln == 0:          # filler row: keeps a slab to two adjacent ix-chunks, owns no velocity
                            assert nch > 1
                            continue
                        in_last = k // 32 == nslab - 1
                        assert ln == (Lt if (Lt > 0 and in_last) else L), (D, n, nranks, rank, k)
                        hi = min(first + ln, n)
                        seen[lay["iz"][k], lay["iy"][k], first:hi] += 1
                    # rows longer than 32 points are cut into ix-chunks; a slab (one warp's 32 rows) spans at most two
                    # ADJACENT chunks, whatever the number of rows a rank owns (a warp's tables hold 64 entries)
                    for s0 in range(0, nrows, 32):
                        fl = [(int(f), int(l)) for f, l in zip(lay["first"][s0:s0 + 32], lay["len"][s0:s0 + 32]) if l > 0]
                        firsts = {f for f, _ in fl}
                        assert max(f + l for f, l in fl) - min(firsts) <= 64, (D, n, nranks, rank, s0, firsts)
                        if nch > 1:
                            assert len(firsts) <= 2 and max(firsts) - min(firsts) <= L, (D, n, nranks, rank, s0, firsts)
                    # the DVs of this rank are exactly its partition
                    ids = capi.partition(n, D, nranks, rank)
                    assert ids.size == sum(min(int(f) + int(l), n) - int(f) for f, l in zip(lay["first"], lay["len"]))
                assert (seen == 1).all(), (D, n, nranks)
    # BASELINE configs 2 and 5 over 8 GPUs: 101 (81) points per direction, 12-13 (10-11) rows per rank
    for n in (101, 81):
        lay = capi.row_layout(n, 2, 8, 3)
        assert lay["nch"] == 4 and (np.asarray(lay["len"]) > 0).sum() == lay["nch"] * (len(capi.partition(n, 2, 8, 3)) // n)
        assert len(lay["iy"]) <= 64          # two slabs: chunks (0, 1) and (2, 3)
    # 1-D with more than 32 points: one row per chunk
    lay = capi.row_layout(41, 1, 1, 0)
    assert lay["nch"] == 2 and list(lay["len"]) == [21, 21]
    # the case that motivated the tail slab: 28^3 over 8 ranks = 98 rows = 3 slabs + 2 rows -> short rows of 2
    lay = capi.row_layout(28, 3, 8, 0)
    assert lay["L"] == 28 and lay["Lt"] == 2 and len(lay["iy"]) == 96 + 28


def test_cell_order_is_a_permutation_with_the_promised_structure():
    """dugks_cell_order (host only): every kind is a permutation; the first class comes first; the wave
    order runs through nWarps lines of cells x position by x position."""
    nx, ny, nz = 12, 10, 8
    iz, iy, ix = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    centres = np.stack([(ix.ravel() + 0.5) / nx, (iy.ravel() + 0.5) / ny * 0.9, (iz.ravel() + 0.5) / nz * 0.7], axis=1)
    n = nx * ny * nz
    interior = ((ix > 0) & (ix < nx - 1) & (iy > 0) & (iy < ny - 1) & (iz > 0) & (iz < nz - 1)).ravel().astype(np.uint8)
    for kind in ("tiled", "wave", "morton", "natural"):
        for first in (None, interior):
            o = capi.cell_order(centres, 3, kind, nWarps=16, first_class=first)
            assert np.array_equal(np.sort(o), np.arange(n)), kind
            if first is not None:
                k = int(interior.sum())
                assert interior[o[:k]].all() and not interior[o[k:]].any(), kind
    assert np.array_equal(capi.cell_order(centres, 3, "natural"), np.arange(n))
    # wave, whole mesh, 16 warps: ny * nz = 80 lines = 5 rounds of 16; inside a round the items
    # [s * 16, (s + 1) * 16) are the 16 lines' cells at x position s, line by line
    o = capi.cell_order(centres, 3, "wave", nWarps=16)
    X, Y, Z = ix.ravel()[o], iy.ravel()[o], iz.ravel()[o]
    for r in range(5):
        blk = slice(r * 16 * nx, (r + 1) * 16 * nx)
        xs = X[blk].reshape(nx, 16)
        assert (xs == np.arange(nx)[:, None]).all()
        lines = (Z[blk] * ny + Y[blk]).reshape(nx, 16)
        assert (lines == lines[0]).all() and len(set(lines[0])) == 16
    # a warp (item w, w + 16, ...) walks ONE line along x
    assert len(set((Z * ny + Y)[3:16 * nx:16])) == 1


def test_pencil_plan_structure():
    """dugks_pencil_plan (host only): x-lines of axis-aligned interior cells in 2 x 2 bundles, cut into work items.
    Every pencil cell appears once; the four cells of a step are the corners of a 2 x 2 square in (y, z) at one x;
    consecutive steps of an item are x+ neighbours; lines that find no partner stay out (the axis-only launch
    takes them); orthogonal meshes whose coordinates are not binary fractions are recognised as well."""
    for n, length, nctas in ((10, 1.0, 296), (11, 1.0, 296), (13, 13 / 16, 7), (16, 1.0, 5)):
        case = cs.cavity3d_case(n, 8, length=length)
        plan = capi.pencil_plan(case, nctas)
        m = n - 2                                    # interior cells per direction
        assert plan["n_axis"] == m ** 3
        nb = (m // 2) ** 2                           # complete 2 x 2 bundles of the m x m lines
        assert len(plan["cells"]) == nb * 4 * m and len(set(plan["cells"].tolist())) == len(plan["cells"])
        assert int(plan["item_steps"].sum()) * 4 == len(plan["cells"])
        C = case.geom.C
        h = length / n
        for first, steps in zip(plan["item_first"], plan["item_steps"]):
            blk = plan["cells"][first:first + 4 * steps].reshape(steps, 4)
            xyz = np.rint(C[blk] / h - 0.5).astype(int)         # integer cell coordinates, [step, line, 3]
            assert (xyz[:, :, 0] == xyz[:, :1, 0]).all() and (np.diff(xyz[:, 0, 0]) == 1).all()
            assert (xyz[:, 1] - xyz[:, 0] == [0, 1, 0]).all() and (xyz[:, 2] - xyz[:, 0] == [0, 0, 1]).all()
            assert (xyz[:, 3] - xyz[:, 0] == [0, 1, 1]).all()
            assert xyz.min() >= 1 and xyz.max() <= n - 2        # interior cells only
        # the schedule: whole lines while they fill the grid, the rest cut so that the last rounds fill it too
        rounds = -(-len(plan["item_steps"]) // nctas)
        work = [int(plan["item_steps"][j::nctas].sum()) for j in range(min(nctas, len(plan["item_steps"])))]
        assert max(work) <= -(-nb * m // min(nctas, nb)) + m // 2 + rounds
    # distorted and triangular meshes have no axis-aligned cells: no pencils, nothing breaks
    for case in (cs.cavity3d_case(5, 8, distort=0.15), cs.tri_cavity_case(6, 8)):
        plan = capi.pencil_plan(case)
        assert plan["n_axis"] == 0 and len(plan["cells"]) == 0
    # 2-D: axis-aligned cells are recognised (the plate thickness 0.1 is not a binary fraction), pencils are 3-D only
    plan = capi.pencil_plan(cs.cavity2d_case(60, 8))
    assert plan["n_axis"] == 58 * 58 and len(plan["cells"]) == 0
