"""ORACLE: reader for the legacy ASCII VTK files write_VTK produces (reference src/mesh/unstructured_mesh/5_VTK.jl:123-157)."""
import numpy as np


def read_vtk(path):
    toks = open(path).read().split("\n")
    out, i = {}, 0
    while i < len(toks):
        t = toks[i].split()
        if t and t[0] == "POINTS":
            n = int(t[1])
            out["points"] = np.array([[float(s) for s in toks[i + 1 + r].split()] for r in range(n)])
            i += n + 1
            continue
        if t and t[0] == "CELLS":
            n = int(t[1])
            out["cells"] = [np.array([int(s) for s in toks[i + 1 + r].split()][1:]) for r in range(n)]
            i += n + 1
            continue
        if t and t[0] == "POINT_DATA":
            npts = int(t[1])
        if t and t[0] == "SCALARS":
            name = t[1]
            out[name] = np.array([float(toks[i + 2 + r]) for r in range(npts)])
            i += npts + 2
            continue
        i += 1
    return out
