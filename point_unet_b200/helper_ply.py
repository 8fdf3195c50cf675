"""Binary PLY point-cloud reader / writer with the reference's calling convention.

Mirrors ``PointSegment/helper_ply.py``: ``read_ply(filename)`` (``:116-196``) returns a numpy structured array whose
fields are the PLY properties (``data['x']``, ``data['class']`` ...); ``write_ply(filename, field_list, field_names)``
(``:217-329``) writes a ``binary_<byteorder>_endian 1.0`` file with one ``element vertex`` whose properties follow
``field_names`` -- the layout PointSegment's data preparation emits (``x y z value class`` for Pancreas,
``x y z t1ce t1 flair t2 class`` for BraTS; ``utils/dataPreparePancreas.py:168``, ``dataPrepareBraTS.py:98``).
Only vertex clouds are handled (the hot path never stores meshes).  Written from the PLY format definition, not
from the reference source.
"""
from __future__ import annotations

import sys

import numpy as np

# PLY scalar type names <-> numpy type codes (both spellings are accepted on input)
_PLY2NP = {"char": "i1", "int8": "i1", "uchar": "u1", "uint8": "u1", "short": "i2", "int16": "i2", "ushort": "u2",
           "uint16": "u2", "int": "i4", "int32": "i4", "uint": "u4", "uint32": "u4", "float": "f4", "float32": "f4",
           "double": "f8", "float64": "f8"}
# the reference writer spells property types the numpy way (float32, uint8, ...); files stay byte-identical
_NP2PLY = {"i1": "int8", "u1": "uint8", "i2": "int16", "u2": "uint16", "i4": "int32", "u4": "uint32", "f4": "float32",
           "f8": "float64"}
_ENDIAN = {"binary_little_endian": "<", "binary_big_endian": ">", "ascii": ""}


def read_ply(filename: str) -> np.ndarray:
    """Read the vertex element of a ``.ply`` file into a structured array (one field per property)."""
    with open(filename, "rb") as f:
        if f.readline().strip() != b"ply":
            raise ValueError(f"{filename}: not a PLY file")
        fmt, n_vertex, props, in_vertex = None, None, [], False
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f"{filename}: truncated PLY header")
            tok = line.decode("ascii", "replace").split()
            if not tok:
                continue
            if tok[0] == "format":
                if tok[1] not in _ENDIAN:
                    raise ValueError(f"{filename}: unsupported PLY format {tok[1]}")
                fmt = tok[1]
            elif tok[0] == "element":
                in_vertex = tok[1] == "vertex"
                if in_vertex:
                    n_vertex = int(tok[2])
            elif tok[0] == "property" and in_vertex:
                if tok[1] == "list":
                    raise ValueError(f"{filename}: list properties are not supported on vertices")
                props.append((tok[2], _PLY2NP[tok[1]]))
            elif tok[0] == "end_header":
                break
        if fmt is None or n_vertex is None:
            raise ValueError(f"{filename}: PLY header lacks format / vertex element")
        if fmt == "ascii":
            rows = np.loadtxt(f, max_rows=n_vertex, ndmin=2)
            out = np.empty(n_vertex, dtype=[(n, t) for n, t in props])
            for i, (n, _) in enumerate(props):
                out[n] = rows[:, i]
            return out
        dtype = np.dtype([(n, _ENDIAN[fmt] + t) for n, t in props])
        return np.fromfile(f, dtype=dtype, count=n_vertex)


def write_ply(filename: str, field_list, field_names) -> bool:
    """Write 1-D / 2-D arrays (each column one property) as a binary PLY vertex cloud.  Returns True on success and
    False (after printing the reason, like the reference) when the fields are inconsistent."""
    fields = list(field_list) if isinstance(field_list, (list, tuple)) else [field_list]
    cols = []
    for a in fields:
        a = np.asarray(a)
        if a.ndim == 1:
            a = a.reshape(-1, 1)
        if a.ndim != 2:
            print("fields have more than 2 dimensions")
            return False
        cols.extend(a[:, j] for j in range(a.shape[1]))
    if len({c.shape[0] for c in cols}) > 1:
        print("wrong field dimensions")
        return False
    if len(cols) != len(field_names):
        print("wrong number of field names")
        return False
    if not filename.endswith(".ply"):
        filename += ".ply"
    order = "<" if sys.byteorder == "little" else ">"
    rec = np.empty(cols[0].shape[0] if cols else 0, dtype=[(n, order + c.dtype.str[1:]) for n, c in zip(field_names, cols)])
    header = ["ply", f"format binary_{sys.byteorder}_endian 1.0", f"element vertex {rec.shape[0]}"]
    for n, c in zip(field_names, cols):
        header.append(f"property {_NP2PLY[c.dtype.str[1:]]} {n}")
        rec[n] = c
    header.append("end_header")
    with open(filename, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        rec.tofile(f)
    return True
