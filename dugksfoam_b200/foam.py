"""Minimal ASCII OpenFOAM reader for the files the reference's cases use
(SURVEY.md Appendix B): polyMesh/{points,faces,owner,neighbour,boundary},
constant/{Xis,weights,DVMProperties}, 0/{rho,U,T}, system/controlDict.

Only what dugksFoam's demo cases need: uniform and nonuniform List fields,
flat dictionaries with one level of sub-dictionaries, dimensioned scalars.
"""
from __future__ import annotations

import os
import re
from typing import Any, Dict, List, Tuple

import numpy as np

from .polymesh import Patch, PolyMesh

_COMMENT_BLOCK = re.compile(r"/\*.*?\*/", re.S)
_COMMENT_LINE = re.compile(r"//[^\n]*")


def _strip(text: str) -> str:
    return _COMMENT_LINE.sub("", _COMMENT_BLOCK.sub("", text))


def _body(text: str) -> str:
    """Text after the FoamFile header dictionary."""
    text = _strip(text)
    m = re.search(r"FoamFile\s*\{.*?\}", text, re.S)
    return text[m.end():] if m else text


def _read(path: str) -> str:
    with open(path, "r") as f:
        return f.read()


def read_scalar_list(path: str) -> np.ndarray:
    """scalarIOList: ``N ( v ... )`` (constant/Xis, constant/weights; fvDVM.C:66-88)."""
    body = _body(_read(path))
    m = re.search(r"(\d+)\s*\(", body)
    n = int(m.group(1))
    end = body.index(")", m.end())
    vals = np.array(body[m.end():end].split(), dtype=np.float64)
    if len(vals) != n:
        raise ValueError(f"{path}: expected {n} scalars, found {len(vals)}")
    return vals


def read_label_list(path: str) -> np.ndarray:
    body = _body(_read(path))
    m = re.search(r"(\d+)\s*\(", body)
    n = int(m.group(1))
    end = body.index(")", m.end())
    vals = np.array(body[m.end():end].split(), dtype=np.int64)
    if len(vals) != n:
        raise ValueError(f"{path}: expected {n} labels, found {len(vals)}")
    return vals


def read_points(path: str) -> np.ndarray:
    body = _body(_read(path))
    m = re.search(r"(\d+)\s*\(", body)
    n = int(m.group(1))
    nums = re.findall(r"[-+]?(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?", body[m.end():])
    pts = np.array(nums[: 3 * n], dtype=np.float64).reshape(n, 3)
    return pts


def read_faces(path: str) -> Tuple[np.ndarray, np.ndarray]:
    body = _body(_read(path))
    m = re.search(r"(\d+)\s*\(", body)
    n = int(m.group(1))
    verts: List[int] = []
    offs = [0]
    for fm in re.finditer(r"(\d+)\s*\(([^()]*)\)", body[m.end():]):
        k = int(fm.group(1))
        vs = fm.group(2).split()
        if len(vs) != k:
            raise ValueError(f"{path}: face with {k} vertices lists {len(vs)}")
        verts.extend(int(v) for v in vs)
        offs.append(len(verts))
        if len(offs) - 1 == n:
            break
    if len(offs) - 1 != n:
        raise ValueError(f"{path}: expected {n} faces, found {len(offs) - 1}")
    return np.array(verts, dtype=np.int64), np.array(offs, dtype=np.int64)


def _tokenize(text: str) -> List[str]:
    return re.findall(r"[{}();]|\"[^\"]*\"|[^\s{}();]+", text)


def _parse_dict(tokens: List[str], pos: int) -> Tuple[Dict[str, Any], int]:
    """Parses ``key value...;`` and ``key { ... }`` entries until '}' or the end."""
    out: Dict[str, Any] = {}
    while pos < len(tokens):
        t = tokens[pos]
        if t == "}":
            return out, pos + 1
        key = t
        pos += 1
        if pos < len(tokens) and tokens[pos] == "{":
            sub, pos = _parse_dict(tokens, pos + 1)
            out[key] = sub
            continue
        vals: List[str] = []
        depth = 0
        while pos < len(tokens):
            t = tokens[pos]
            if t == ";" and depth == 0:
                pos += 1
                break
            if t == "(":
                depth += 1
            elif t == ")":
                depth -= 1
            vals.append(t)
            pos += 1
        out[key] = vals
    return out, pos


def read_dict(path: str) -> Dict[str, Any]:
    tokens = _tokenize(_body(_read(path)))
    d, _ = _parse_dict(tokens, 0)
    return d


def dimensioned_value(vals: List[str]) -> float:
    """``name [dims] value`` or plain ``value`` -> float (fvDVM.C:916-931)."""
    return float(vals[-1])


def read_boundary(path: str) -> List[Patch]:
    body = _body(_read(path))
    m = re.search(r"(\d+)\s*\(", body)
    tokens = _tokenize(body[m.end():])
    patches: List[Patch] = []
    pos = 0
    while pos < len(tokens) and tokens[pos] != ")":
        name = tokens[pos]
        assert tokens[pos + 1] == "{", f"{path}: malformed patch {name}"
        d, pos = _parse_dict(tokens, pos + 2)
        patches.append(Patch(name, d["type"][0], int(d["nFaces"][0]), int(d["startFace"][0])))
    return patches


def read_polymesh(case_dir: str) -> PolyMesh:
    pm = os.path.join(case_dir, "constant", "polyMesh")
    points = read_points(os.path.join(pm, "points"))
    verts, offs = read_faces(os.path.join(pm, "faces"))
    owner = read_label_list(os.path.join(pm, "owner"))
    neighbour = read_label_list(os.path.join(pm, "neighbour"))
    patches = read_boundary(os.path.join(pm, "boundary"))
    return PolyMesh(points=points, face_verts=verts, face_offsets=offs, owner=owner,
                    neighbour=neighbour, patches=patches)


def _field_values(vals: List[str], n: int, ncomp: int) -> np.ndarray:
    """``uniform v`` / ``uniform (x y z)`` / ``nonuniform List<..> N ( ... )``."""
    if vals[0] == "uniform":
        nums = [float(v) for v in vals[1:] if v not in "()"]
        return np.tile(np.array(nums, dtype=np.float64).reshape(1, ncomp), (n, 1)).reshape(n, ncomp)
    if vals[0] == "nonuniform":
        nums = [v for v in vals[3:] if v not in "()"]
        arr = np.array(nums, dtype=np.float64).reshape(-1, ncomp)
        if len(arr) != n:
            raise ValueError(f"nonuniform field has {len(arr)} entries, expected {n}")
        return arr
    raise ValueError(f"unsupported field entry {vals[:3]}")


def read_field(path: str, n_cells: int, ncomp: int):
    """Returns (internal [n_cells, ncomp], {patch: dict-of-token-lists})."""
    d = read_dict(path)
    internal = _field_values(d["internalField"], n_cells, ncomp)
    return internal, d.get("boundaryField", {})


def patch_field_values(entry: Dict[str, Any], n: int, ncomp: int):
    """Value of a patch entry or None when it has no ``value`` keyword."""
    if "value" not in entry:
        return None
    return _field_values(entry["value"], n, ncomp)
