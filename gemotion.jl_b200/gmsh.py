"""Gmsh `.msh` reader (ASCII, format 2.x and 4.1) for triangle meshes with named physical groups.

Host-side counterpart of `GmshDiscreteModel("...co-annulus_unstructured_4.msh")`
(/root/reference/examples/validation-kuehn/kuehn.jl:13-22; GridapGmsh 0.7.2) for BASELINE config 3: the tags
"inner" / "outer" / "all" / "heatfunction_zero_points" / "mesh" of
/root/reference/examples/meshes/2.6/co-annulus_unstructured.geo:25-48 select the Dirichlet sets (kuehn.jl:47-50).
The `.msh` files of the reference are Git-LFS pointers and gmsh is not in the image, so tests write their own files
with :func:`write_msh` from generated models; the reader itself makes no assumption about who wrote the file.

Conventions restated from GridapGmsh (unverified against its source: Julia is absent):
  * every geometric (elementary) entity that carries elements becomes one entity id; a tag name maps to the entities of
    all physical groups of that name, whatever their dimension (the .geo gives the name "inner" to a Physical Point group
    AND a Physical Curve group, and puts the same points / curves into "all" as well: in a format-2 file such an element
    is written once per group, with the same elementary tag);
  * a mesh vertex takes the entity of the lowest-dimensional element it belongs to (point element, else a line element
    containing it, else the surface); an edge takes the entity of its line element, else the surface's;
  * mesh nodes keep file order (nodes not used by any triangle are dropped); each triangle's vertex ids are sorted
    ascending (fe.TriangleModel does that), so det J may be negative and |det J| is used.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

from . import fe


def _sections(text: str) -> Dict[str, List[str]]:
    out, name, buf = {}, None, []
    for line in text.splitlines():
        s = line.strip()
        if s.startswith("$End"):
            out[name] = buf
            name, buf = None, []
        elif s.startswith("$"):
            name, buf = s[1:], []
        elif name is not None and s:
            buf.append(s)
    return out


def _parse_v2(sec):
    nodes = {}
    for ln in sec["Nodes"][1:]:
        p = ln.split()
        nodes[int(p[0])] = (float(p[1]), float(p[2]))
    elems = []  # (etype, [physical ids], elementary entity, node ids)
    for ln in sec["Elements"][1:]:
        p = [int(v) for v in ln.split()]
        etype, ntags = p[1], p[2]
        phys = [p[3]] if ntags > 0 and p[3] else []
        elem = p[4] if ntags > 1 else (p[3] if ntags > 0 else 0)
        elems.append((etype, phys, elem, p[3 + ntags:]))
    return nodes, elems


def _parse_v41(sec):
    ent_phys = {0: {}, 1: {}, 2: {}, 3: {}}  # geometric entity -> physical ids
    if "Entities" in sec:
        L = sec["Entities"]
        cnt = [int(v) for v in L[0].split()]
        k = 1
        for dim in range(4):
            for _ in range(cnt[dim]):
                p = L[k].split()
                k += 1
                off = 4 if dim == 0 else 7  # point: tag x y z ; others: tag + bounding box
                nphys = int(p[off])
                ent_phys[dim][int(p[0])] = [abs(int(v)) for v in p[off + 1: off + 1 + nphys]]
    nodes = {}
    L = sec["Nodes"]
    nblocks = int(L[0].split()[0])
    k = 1
    for _ in range(nblocks):
        nb = int(L[k].split()[3])
        k += 1
        ids = [int(L[k + i]) for i in range(nb)]
        k += nb
        for i in range(nb):
            p = L[k + i].split()
            nodes[ids[i]] = (float(p[0]), float(p[1]))
        k += nb
    elems = []
    L = sec["Elements"]
    nblocks = int(L[0].split()[0])
    k = 1
    for _ in range(nblocks):
        dim, etag, etype, nb = [int(v) for v in L[k].split()]
        k += 1
        for i in range(nb):
            p = [int(v) for v in L[k + i].split()]
            elems.append((etype, ent_phys.get(dim, {}).get(etag, []), etag, p[1:]))
        k += nb
    return nodes, elems


_DIM = {15: 0, 1: 1, 2: 2}


def read_msh(path: str) -> fe.TriangleModel:
    sec = _sections(open(path).read())
    head = sec["MeshFormat"][0].split()
    version = float(head[0])
    if int(head[1]) != 0:
        raise ValueError("binary .msh files are not supported")
    names: Dict[Tuple[int, int], str] = {}
    for ln in sec.get("PhysicalNames", [])[1:]:
        p = ln.split(None, 2)
        names[(int(p[0]), int(p[1]))] = p[2].strip().strip('"')
    nodes, elems = _parse_v41(sec) if version >= 4.0 else _parse_v2(sec)
    elems = [e for e in elems if e[0] in _DIM]
    tris = [e for e in elems if e[0] == 2]
    if not tris:
        raise ValueError("no triangles in %s (quad-recombined meshes are not supported)" % path)
    # one entity per geometric (elementary) entity that carries elements, numbered in (dim, tag) order; an element that
    # belongs to several physical groups appears once per group in a format-2 file: same elementary tag, so same entity
    geo = sorted({(_DIM[et], el) for et, _, el, _ in elems})
    ent_of = {g: i + 1 for i, g in enumerate(geo)}
    tag_to_entities: Dict[str, List[int]] = {}
    for et, phys, el, _ in elems:
        for ph in phys:
            lst = tag_to_entities.setdefault(names.get((_DIM[et], ph), "physical_%d_%d" % (_DIM[et], ph)), [])
            if ent_of[(_DIM[et], el)] not in lst:
                lst.append(ent_of[(_DIM[et], el)])
    # nodes used by triangles, file order; triangles once (duplicates of multi-group surfaces dropped)
    seen, tri_nodes = set(), []
    for _, _, el, n in tris:
        key = tuple(n)
        if key not in seen:
            seen.add(key)
            tri_nodes.append((el, n))
    used = sorted({v for _, n in tri_nodes for v in n})
    new_id = {old: k for k, old in enumerate(used)}
    coords = np.array([nodes[o] for o in used], dtype=np.float64)
    tri = np.array([[new_id[v] for v in n] for _, n in tri_nodes], dtype=np.int64)
    interior = ent_of[(2, tri_nodes[0][0])]
    vent = np.full(len(used), interior, dtype=np.int32)
    bedges = {}
    for et, _, el, n in elems:  # curves first, then points override (lowest dimension wins)
        if et == 1 and all(v in new_id for v in n):
            a, b = new_id[n[0]], new_id[n[1]]
            bedges[(min(a, b), max(a, b))] = ent_of[(1, el)]
            vent[a] = vent[b] = ent_of[(1, el)]
    for et, _, el, n in elems:
        if et == 15 and n[0] in new_id:
            vent[new_id[n[0]]] = ent_of[(0, el)]
    return fe.TriangleModel(coords, tri, vent, bedges, tag_to_entities, interior_entity=interior)


def write_msh(model: fe.TriangleModel, path: str, version: str = "2.2", tri_vertices=None) -> None:
    """Writes `model` the way gmsh saves a mesh with physical groups (fixtures for the reader): one geometric entity per model
    entity, one physical group per (dimension, tag name); in format 2 an element that belongs to several groups is written
    once per group.  `tri_vertices` (unsorted) may be given to keep a generator's orientation in the file."""
    interior = int(model.cell_entity[0])
    dim_of = {e: 0 for e in _dim_entities(model, 0)}
    dim_of.update({e: 1 for e in _dim_entities(model, 1)})
    dim_of[interior] = 2
    phys, pid = [], {}  # physical groups (dim, id, name)
    for name, ents in model.tag_to_entities.items():
        for d in sorted({dim_of[e] for e in ents if e in dim_of}):
            pid[(d, str(name))] = len(phys) + 1
            phys.append((d, len(phys) + 1, str(name)))
    groups_of = {e: [pid[(dim_of[e], str(nm))] for nm, ents in model.tag_to_entities.items() if e in ents] for e in dim_of}
    vent, eent = model.vertex_entity, model.edge_entity
    tris = model.cell_vertices if tri_vertices is None else np.asarray(tri_vertices)
    el = []  # (etype, physical ids, elementary, nodes 1-based)
    for v in range(model.nverts):
        if dim_of.get(int(vent[v])) == 0:
            el.append((15, groups_of[int(vent[v])], int(vent[v]), [v + 1]))
    for k, (a, b) in enumerate(model.edge_vertices):
        if dim_of.get(int(eent[k])) == 1:
            el.append((1, groups_of[int(eent[k])], int(eent[k]), [int(a) + 1, int(b) + 1]))
    for t in tris:
        el.append((2, groups_of[interior], interior, [int(v) + 1 for v in t]))
    with open(path, "w") as f:
        f.write("$MeshFormat\n%s 0 8\n$EndMeshFormat\n$PhysicalNames\n%d\n" % ("2.2" if version.startswith("2") else "4.1", len(phys)))
        for d, i, nm in phys:
            f.write('%d %d "%s"\n' % (d, i, nm))
        f.write("$EndPhysicalNames\n")
        if version.startswith("2"):
            f.write("$Nodes\n%d\n" % model.nverts)
            for k, (x, y) in enumerate(model.vertex_coords):
                f.write("%d %.17g %.17g 0\n" % (k + 1, x, y))
            rows = [(et, ph, elem, n) for et, phs, elem, n in el for ph in (phs or [0])]
            f.write("$EndNodes\n$Elements\n%d\n" % len(rows))
            for k, (et, ph, elem, n) in enumerate(rows):
                f.write("%d %d 2 %d %d %s\n" % (k + 1, et, ph, elem, " ".join(str(v) for v in n)))
            f.write("$EndElements\n")
        else:
            ents = {d: sorted(e for e, dd in dim_of.items() if dd == d) for d in range(3)}
            f.write("$Entities\n%d %d %d 0\n" % (len(ents[0]), len(ents[1]), len(ents[2])))
            for e in ents[0]:
                f.write("%d 0 0 0 %d %s\n" % (e, len(groups_of[e]), " ".join(str(g) for g in groups_of[e])))
            for d in (1, 2):
                for e in ents[d]:
                    f.write("%d -2 -2 0 2 2 0 %d %s 0\n" % (e, len(groups_of[e]), " ".join(str(g) for g in groups_of[e])))
            f.write("$EndEntities\n$Nodes\n1 %d 1 %d\n2 %d 0 %d\n" % (model.nverts, model.nverts, interior, model.nverts))
            for k in range(model.nverts):
                f.write("%d\n" % (k + 1))
            for x, y in model.vertex_coords:
                f.write("%.17g %.17g 0\n" % (x, y))
            f.write("$EndNodes\n")
            blocks = {}
            for et, _, elem, n in el:
                blocks.setdefault((et, elem), []).append(n)
            f.write("$Elements\n%d %d 1 %d\n" % (len(blocks), len(el), len(el)))
            k = 0
            for (et, elem), lst in blocks.items():
                f.write("%d %d %d %d\n" % (_DIM[et], elem, et, len(lst)))
                for n in lst:
                    k += 1
                    f.write("%d %s\n" % (k, " ".join(str(v) for v in n)))
            f.write("$EndElements\n")


def _dim_entities(model, dim):
    """Entities that label vertices only (dim 0) / boundary edges (dim 1) in a fe.TriangleModel."""
    interior = int(model.cell_entity[0])
    edge_ents = sorted({int(e) for e in model.edge_entity if e != interior})
    if dim == 1:
        return edge_ents
    return sorted({int(e) for e in model.vertex_entity if e != interior and int(e) not in edge_ents})
