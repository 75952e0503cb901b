"""Gmsh .msh reader (gemotion.jl_b200/gmsh.py): GmshDiscreteModel of /root/reference/examples/validation-kuehn/kuehn.jl:13-22
with the physical names of examples/meshes/2.6/co-annulus_unstructured.geo:25-48 (BASELINE config 3)."""
import os

import numpy as np
import pytest

import cases
from cases import fe

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _same_problem(a, b):
    """Two models describe the same FE problem: same cells, same Dirichlet sets for the kuehn BCs -> same dof tables."""
    sa, sb = cases.spaces(a, cases.KUEHN), cases.spaces(b, cases.KUEHN)
    assert np.array_equal(a.cell_vertices, b.cell_vertices) and np.allclose(a.vertex_coords, b.vertex_coords, rtol=0, atol=0)
    assert np.array_equal(sa.U.cell_dofs, sb.U.cell_dofs) and np.array_equal(sa.T.cell_dofs, sb.T.cell_dofs)
    assert np.array_equal(sa.U.diri_values, sb.U.diri_values) and np.array_equal(sa.T.diri_values, sb.T.diri_values)
    assert np.array_equal(sa.n_free, sb.n_free)
    return sa


@pytest.mark.parametrize("fn", ["annulus_tiny_v22.msh", "annulus_tiny_v41.msh"])
def test_committed_fixture_reads_back_as_the_generated_annulus(gm, fn):
    gen = fe.annulus_model(nr=3, nt=12, jitter=0.25, seed=11)  # tests/golden/make_msh_fixtures.py
    rd = gm.gmsh.read_msh(os.path.join(GOLD, fn))
    assert rd.ncells == 72 and rd.nverts == 48
    for name in ("inner", "outer", "all", "heatfunction_zero_points", "mesh"):
        assert name in rd.tag_to_entities
    # "all" = inner + outer points and curves; the four arc end points are point entities (geo:25-28)
    assert sorted(rd.tag_to_entities["all"]) == sorted(rd.tag_to_entities["inner"] + rd.tag_to_entities["outer"])
    pts = np.isin(rd.vertex_entity, rd.tag_to_entities["heatfunction_zero_points"])
    assert pts.sum() == 4 and np.allclose(np.abs(rd.vertex_coords[pts][:, 0]), 0.0, atol=1e-15)
    sp = _same_problem(gen, rd)
    assert sp.T.diri_values.size == 48 and set(np.unique(sp.T.diri_values)) == {0.0, 1.0}  # inner = 1, outer = 0 (kuehn.jl:48)


@pytest.mark.parametrize("version", ["2.2", "4.1"])
def test_roundtrip_of_an_irregular_mesh_and_orientation(gm, tmp_path, version):
    """Triangles are written with the generator's (counter-clockwise) vertex order; the reader sorts each simplex's
    vertex ids like GridapGmsh, so det J may be negative: the assembled operator must not care."""
    from oracle.oracle import Oracle
    gen = fe.annulus_model(nr=4, nt=16, jitter=0.3, seed=5)
    ccw = []
    for t in gen.cell_vertices:  # a file in which every triangle is counter-clockwise but not sorted
        p = gen.vertex_coords[t]
        area = (p[1, 0] - p[0, 0]) * (p[2, 1] - p[0, 1]) - (p[1, 1] - p[0, 1]) * (p[2, 0] - p[0, 0])
        t = t if area > 0 else t[[0, 2, 1]]
        ccw.append(np.roll(t, 1))
    path = str(tmp_path / "a.msh")
    gm.gmsh.write_msh(gen, path, version, tri_vertices=np.array(ccw))
    rd = gm.gmsh.read_msh(path)
    sp = _same_problem(gen, rd)
    x = fe.splitmix_iterate(sp.n_total, 1234)
    prm = dict(Pr=0.706, Ra=4.7e4, n=1.0)
    nz_a, b_a = Oracle(cases.spaces(gen, cases.KUEHN), **prm).assemble(x)
    nz_b, b_b = Oracle(sp, **prm).assemble(x)
    assert np.array_equal(nz_a, nz_b) and np.array_equal(b_a, b_b)


def test_unused_nodes_and_multi_group_duplicates(gm, tmp_path):
    """A format-2 file as gmsh writes it for the .geo: the circle centre is a geometry point without mesh use, elements of
    entities in several physical groups appear once per group."""
    txt = """$MeshFormat
2.2 0 8
$EndMeshFormat
$PhysicalNames
4
0 1 "inner"
1 2 "inner"
1 3 "all"
2 4 "mesh"
$EndPhysicalNames
$Nodes
5
1 0 0 0
2 1 0 0
3 0 1 0
4 5 5 0
5 1 1 0
$EndNodes
$Elements
6
1 15 2 1 1000 1
2 1 2 2 2000 1 2
3 1 2 3 2000 1 2
4 1 2 3 2001 2 5
5 2 2 4 4000 1 2 3
6 2 2 4 4000 2 5 3
$EndElements
"""
    p = tmp_path / "t.msh"
    p.write_text(txt)
    m = gm.gmsh.read_msh(str(p))
    assert m.nverts == 4 and m.ncells == 2  # node 4 is not used by a triangle
    assert len(m.tag_to_entities["inner"]) == 2 and len(m.tag_to_entities["all"]) == 2
    e_pt = [e for e in m.tag_to_entities["inner"] if e not in m.tag_to_entities["all"]][0]
    assert m.vertex_entity[0] == e_pt                       # the point element wins over the curve through it
    assert m.vertex_entity[1] in m.tag_to_entities["all"]   # vertex on a labelled curve inherits the curve's entity
    assert m.vertex_entity[2] == m.cell_entity[0]
