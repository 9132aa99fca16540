"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports every
symbol include/dugks.h declares, and fails loudly (no CPU fallback) when asked to compute."""
import ctypes
import os
import re

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
    assert lib.dugks_abi_version() == 1


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
    assert ctypes.sizeof(abi.StatsT) == 40
