"""Writes the committed .msh fixtures: a tiny concentric annulus (R1 = 5/8, R2 = 13/8 and the physical groups of
/root/reference/examples/meshes/2.6/co-annulus_unstructured.geo:12-48) in Gmsh format 2.2 (what the .geo requests:
Mesh.MshFileVersion = 2.0) and 4.1 (the structured variants' default).  The reference's own .msh files are Git-LFS pointers
and gmsh is not in the image, so these are written by gemotion.jl_b200/gmsh.py:write_msh from fe.annulus_model.
    python tests/golden/make_msh_fixtures.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from _pkg import gm  # noqa: E402

model = gm.fe.annulus_model(nr=3, nt=12, jitter=0.25, seed=11)
gm.gmsh.write_msh(model, os.path.join(HERE, "annulus_tiny_v22.msh"), "2.2")
gm.gmsh.write_msh(model, os.path.join(HERE, "annulus_tiny_v41.msh"), "4.1")
print("wrote", model.ncells, "triangles")
