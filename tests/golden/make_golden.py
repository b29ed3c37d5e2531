"""Generates the golden fixtures of tests/golden/ from the reference's own committed example inputs and results.

Run in the build container (needs /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
Fixtures (inputs + the reference's published result for the same inputs):
  thermal3d.npz  examples/thermal_conduction/3D_COMSOL_Mesh.mphtxt  ->  3D_MetaFEM_Result.vtk (scalar T, cell lines)
  stress3d.npz   examples/linear_elasticity/stress_concentration/3D_Mesh.inp -> 3D_MetaFEM.vtk (d1, d2, d3)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import refgeom as rg, vtk  # noqa: E402

REF = "/root/reference/examples/"


def main():
    vert, conn = rg.read_MPHTXT(REF + "thermal_conduction/3D_COMSOL_Mesh.mphtxt")
    g = vtk.read_vtk(REF + "thermal_conduction/3D_MetaFEM_Result.vtk")
    np.savez_compressed(os.path.join(HERE, "thermal3d.npz"), vert=vert, conn=conn.astype(np.int32),
                        points=g["points"].astype(np.float32), T=g["T"], cells=np.array(g["cells"], dtype=np.int32))
    vert, conn = rg.read_INP(REF + "linear_elasticity/stress_concentration/3D_Mesh.inp")
    g = vtk.read_vtk(REF + "linear_elasticity/stress_concentration/3D_MetaFEM.vtk")
    np.savez_compressed(os.path.join(HERE, "stress3d.npz"), vert=vert, conn=conn.astype(np.int32),
                        points=g["points"].astype(np.float32), d1=g["d1"], d2=g["d2"], d3=g["d3"])
    for f in ("thermal3d.npz", "stress3d.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
