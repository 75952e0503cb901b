"""Host-side discrete models, FE spaces and reference-element tables.

In production these arrays come from Gridap (`get_cell_dof_ids`,
`get_dirichlet_dof_values`, tabulated shape functions, cell quadrature; see
INTEGRATION.md).  Julia/Gridap are not available in this image, so this module
restates the parts of Gridap 0.18.9 that the reference drives at
/root/reference/src/Solver.jl:80-109:

* `CartesianDiscreteModel((x0,x1,y0,y1),(nx,ny))`  -> :class:`CartesianModel`
  (examples/simple.jl:12) - vertex/cell ids x-fastest, tensor-order cell
  vertices, first-encounter edge ids, face entities 1..9.
* `GmshDiscreteModel(...)` triangles -> :class:`TriangleModel`
  (examples/validation-kuehn/kuehn.jl:13-22) - named tags.
* `TestFESpace(model, ReferenceFE(lagrangian, ..., 2), conformity=:H1,
  dirichlet_tags=...)` (Solver.jl:81-96), the discontinuous P1 `space=:P`
  zero-mean pressure space (Solver.jl:89-90), `MultiFieldFESpace` (Solver.jl:104-105)
  -> :func:`build_spaces`.
* `Measure(Ω, 2)` (Solver.jl:107-109) -> :func:`reference_tables`.

Everything here is integer / table generation done ONCE per mesh; nothing in this
file runs per Newton iteration.
"""
from __future__ import annotations

import dataclasses
from typing import Callable, Dict, List, Optional, Sequence, Union

import numpy as np

QUAD = 0  # cell type ids used across the C ABI (include/gemotion_b200.h)
TRI = 1


# --------------------------------------------------------------------------------------
# reference elements + quadrature
# --------------------------------------------------------------------------------------
@dataclasses.dataclass
class RefTables:
    """Tabulated reference-element data at the quadrature points.

    w      [nq]           quadrature weights (reference cell)
    qp     [nq,2]         quadrature points
    phi    [nq,nn]        scalar P2 shape values        (velocity comps and temperature)
    dphi   [nq,nn,2]      reference gradients of phi
    psi    [nq,3]         pressure shape values
    gphi   [nq,nv]        geometry (vertex) shape values
    dgphi  [nq,nv,2]      reference gradients of the geometry shape functions
    nodes  [nn,2]         P2 node coordinates on the reference cell
    """
    cell_type: int
    w: np.ndarray
    qp: np.ndarray
    phi: np.ndarray
    dphi: np.ndarray
    psi: np.ndarray
    gphi: np.ndarray
    dgphi: np.ndarray
    nodes: np.ndarray

    @property
    def nq(self):
        return self.w.shape[0]

    @property
    def nn(self):
        return self.phi.shape[1]

    @property
    def nv(self):
        return self.gphi.shape[1]


def _lag3(t):
    """1-D quadratic Lagrange basis on nodes {0, 1/2, 1}: values and derivatives."""
    v = np.array([2.0 * (t - 0.5) * (t - 1.0), -4.0 * t * (t - 1.0), 2.0 * t * (t - 0.5)])
    d = np.array([4.0 * t - 3.0, -8.0 * t + 4.0, 4.0 * t - 1.0])
    return v, d


# QUAD P2 node order: 4 vertices in tensor order, 4 edge midpoints
# (bottom, top, left, right), centre.  (ix,iy) index the 1-D nodes {0,1/2,1}.
_QUAD_NODE_IXIY = [(0, 0), (2, 0), (0, 2), (2, 2), (1, 0), (1, 2), (0, 1), (2, 1), (1, 1)]
# TRI P2 node order: vertices (0,0),(1,0),(0,1); edge midpoints of (v1v2),(v1v3),(v2v3)
_TRI_NODES = np.array([[0, 0], [1, 0], [0, 1], [.5, 0], [0, .5], [.5, .5]], dtype=np.float64)


def quad_p2(x, y):
    vx, dx = _lag3(x)
    vy, dy = _lag3(y)
    phi = np.array([vx[i] * vy[j] for i, j in _QUAD_NODE_IXIY])
    dphi = np.array([[dx[i] * vy[j], vx[i] * dy[j]] for i, j in _QUAD_NODE_IXIY])
    return phi, dphi


def tri_p2(x, y):
    l0, l1, l2 = 1.0 - x - y, x, y
    phi = np.array([l0 * (2 * l0 - 1), l1 * (2 * l1 - 1), l2 * (2 * l2 - 1),
                    4 * l0 * l1, 4 * l0 * l2, 4 * l1 * l2])
    g0, g1, g2 = np.array([-1.0, -1.0]), np.array([1.0, 0.0]), np.array([0.0, 1.0])
    dphi = np.array([(4 * l0 - 1) * g0, (4 * l1 - 1) * g1, (4 * l2 - 1) * g2,
                     4 * (l0 * g1 + l1 * g0), 4 * (l0 * g2 + l2 * g0), 4 * (l1 * g2 + l2 * g1)])
    return phi, dphi


def p1_simplex(x, y):
    """Pressure basis `space=:P`, order 1 (Solver.jl:89): {1-ξ-η, ξ, η} on either cell type."""
    return np.array([1.0 - x - y, x, y])


def reference_tables(cell_type: int, degree: int = 2) -> RefTables:
    """`Measure(Ωₕ, degree)` with degree=2 (Solver.jl:107-109) + tabulated bases."""
    if cell_type == QUAD:
        n1 = (degree + 2) // 2  # Gauss-Legendre points per direction: ceil((degree+1)/2)
        xg, wg = np.polynomial.legendre.leggauss(n1)
        xg = 0.5 * (xg + 1.0)
        wg = 0.5 * wg
        qp = np.array([[xg[i], xg[j]] for j in range(n1) for i in range(n1)])  # x fastest
        w = np.array([wg[i] * wg[j] for j in range(n1) for i in range(n1)])
        shp = quad_p2
        nodes = np.array([[0.5 * i, 0.5 * j] for i, j in _QUAD_NODE_IXIY])

        def geo(x, y):
            g = np.array([(1 - x) * (1 - y), x * (1 - y), (1 - x) * y, x * y])
            dg = np.array([[-(1 - y), -(1 - x)], [(1 - y), -x], [-y, (1 - x)], [y, x]])
            return g, dg
    elif cell_type == TRI:
        if degree == 2:
            qp = np.array([[2.0 / 3, 1.0 / 6], [1.0 / 6, 2.0 / 3], [1.0 / 6, 1.0 / 6]])
            w = np.full(3, 1.0 / 6)
        elif degree == 1:
            qp = np.array([[1.0 / 3, 1.0 / 3]])
            w = np.array([0.5])
        else:
            raise ValueError("TRI quadrature degree %d not tabulated" % degree)
        shp = tri_p2
        nodes = _TRI_NODES.copy()

        def geo(x, y):
            return (np.array([1 - x - y, x, y]),
                    np.array([[-1.0, -1.0], [1.0, 0.0], [0.0, 1.0]]))
    else:
        raise ValueError("unknown cell type")
    phi = np.array([shp(x, y)[0] for x, y in qp])
    dphi = np.array([shp(x, y)[1] for x, y in qp])
    psi = np.array([p1_simplex(x, y) for x, y in qp])
    gphi = np.array([geo(x, y)[0] for x, y in qp])
    dgphi = np.array([geo(x, y)[1] for x, y in qp])
    c = np.ascontiguousarray
    return RefTables(cell_type, c(w), c(qp), c(phi), c(dphi), c(psi), c(gphi), c(dgphi), c(nodes))


def shape_at(cell_type: int, x: float, y: float):
    """Scalar P2 shape values / reference gradients at one reference point."""
    return quad_p2(x, y) if cell_type == QUAD else tri_p2(x, y)


# --------------------------------------------------------------------------------------
# discrete models
# --------------------------------------------------------------------------------------
def _first_encounter_edges(cell_vertices: np.ndarray, local_edges: Sequence[Sequence[int]]):
    """Edge ids in first-encounter order: cells in id order, local edges in order."""
    nc = cell_vertices.shape[0]
    ne = len(local_edges)
    a = np.stack([cell_vertices[:, e[0]] for e in local_edges], axis=1).astype(np.int64)
    b = np.stack([cell_vertices[:, e[1]] for e in local_edges], axis=1).astype(np.int64)
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    nvert = int(cell_vertices.max()) + 1
    key = (lo * nvert + hi).reshape(-1)  # (cell, local edge) order == encounter order
    uniq, first, inv = np.unique(key, return_index=True, return_inverse=True)
    order = np.argsort(first, kind="stable")  # unique keys sorted by first occurrence
    rank = np.empty_like(order)
    rank[order] = np.arange(order.size)
    cell_edges = rank[inv].reshape(nc, ne).astype(np.int32)
    edge_vertices = np.stack([uniq[order] // nvert, uniq[order] % nvert], axis=1).astype(np.int32)
    return cell_edges, edge_vertices


class DiscreteModel:
    """Common face/label container.

    cell_vertices [nc,nv] int32 (0-based), vertex_coords [nvert,2],
    cell_edges [nc,ne], edge_vertices [nedge,2],
    vertex_entity / edge_entity / cell_entity: int32 entity ids,
    tag_to_entities: name-or-int -> list of entity ids.
    """
    cell_type: int
    cartesian: bool = False

    def tag_entities(self, tag) -> List[int]:
        if tag in self.tag_to_entities:
            return list(self.tag_to_entities[tag])
        raise KeyError("unknown tag %r" % (tag,))

    @property
    def ncells(self):
        return self.cell_vertices.shape[0]

    @property
    def nverts(self):
        return self.vertex_coords.shape[0]

    @property
    def nedges(self):
        return self.edge_vertices.shape[0]


class CartesianModel(DiscreteModel):
    """`CartesianDiscreteModel((x0,x1,y0,y1),(nx,ny))` (examples/simple.jl:12).

    Entities: corners 1:(x0,y0) 2:(x1,y0) 3:(x0,y1) 4:(x1,y1); open edges 5:y=y0 6:y=y1 7:x=x0
    8:x=x1; interior 9 (the sketch at examples/simple.jl:6-11).
    """
    cell_type = QUAD
    cartesian = True
    LOCAL_EDGES = [(0, 1), (2, 3), (0, 2), (1, 3)]  # bottom, top, left, right

    def __init__(self, domain=(0.0, 1.0, 0.0, 1.0), partition=(10, 10)):
        x0, x1, y0, y1 = [float(v) for v in domain]
        nx, ny = int(partition[0]), int(partition[1])
        self.domain = (x0, x1, y0, y1)
        self.partition = (nx, ny)
        self.origin = (x0, y0)
        self.h = ((x1 - x0) / nx, (y1 - y0) / ny)
        i = np.arange(nx + 1)
        j = np.arange(ny + 1)
        X = x0 + i * self.h[0]
        Y = y0 + j * self.h[1]
        self.vertex_coords = np.stack([np.tile(X, ny + 1), np.repeat(Y, nx + 1)], axis=1)
        cx = np.tile(np.arange(nx), ny)
        cy = np.repeat(np.arange(ny), nx)
        v0 = cx + cy * (nx + 1)
        self.cell_vertices = np.stack([v0, v0 + 1, v0 + nx + 1, v0 + nx + 2], axis=1).astype(np.int32)
        self.cell_edges, self.edge_vertices = self._closed_form_edges(nx, ny, self.cell_vertices)
        # entities
        vi = np.tile(i, ny + 1)
        vj = np.repeat(j, nx + 1)
        self.vertex_entity = self._entity(vi == 0, vi == nx, vj == 0, vj == ny)
        ea, eb = self.edge_vertices[:, 0], self.edge_vertices[:, 1]
        ei0, ei1, ej0, ej1 = vi[ea], vi[eb], vj[ea], vj[eb]
        self.edge_entity = self._entity((ei0 == 0) & (ei1 == 0), (ei0 == nx) & (ei1 == nx),
                                        (ej0 == 0) & (ej1 == 0), (ej0 == ny) & (ej1 == ny))
        self.cell_entity = np.full(nx * ny, 9, dtype=np.int32)
        self.tag_to_entities: Dict[Union[int, str], List[int]] = {k: [k] for k in range(1, 10)}
        for k in range(1, 9):
            self.tag_to_entities["tag_%d" % k] = [k]
        self.tag_to_entities["interior"] = [9]
        self.tag_to_entities["boundary"] = list(range(1, 9))

    @staticmethod
    def _closed_form_edges(nx, ny, cell_vertices):
        """First-encounter edge ids in closed form (identical to `_first_encounter_edges`, checked in
        tests/test_fe.py): row 0 numbers bottom, top, (left of the first cell), right per cell; every
        later row numbers top, (left of the first cell), right per cell."""
        cx = np.tile(np.arange(nx, dtype=np.int64), ny)
        cy = np.repeat(np.arange(ny, dtype=np.int64), nx)
        row0 = cy == 0
        # ids inside row 0: cell 0 -> bottom 0, top 1, left 2, right 3 ; cell cx>0 -> base 4+3(cx-1): bottom, top, right
        b0 = np.where(cx == 0, 0, 4 + 3 * (cx - 1))
        bottom0 = b0
        top0 = b0 + 1
        right0 = np.where(cx == 0, 3, b0 + 2)
        # rows >= 1: base(cy) = 3nx+1 + (cy-1)(2nx+1) ; cell 0 -> top, left, right ; cell cx>0 -> top, right at base+3+2(cx-1)
        base = 3 * nx + 1 + (cy - 1) * (2 * nx + 1)
        b1 = base + np.where(cx == 0, 0, 3 + 2 * (cx - 1))
        top1 = b1
        right1 = np.where(cx == 0, b1 + 2, b1 + 1)
        top = np.where(row0, top0, top1)
        right = np.where(row0, right0, right1)
        nc = nx * ny
        topg = top.reshape(ny, nx)
        rightg = right.reshape(ny, nx)
        bottom = np.empty((ny, nx), dtype=np.int64)
        bottom[0] = bottom0[:nx]
        if ny > 1:
            bottom[1:] = topg[:-1]
        left = np.empty((ny, nx), dtype=np.int64)
        left[:, 0] = np.where(np.arange(ny) == 0, 2, 3 * nx + 1 + (np.arange(ny) - 1) * (2 * nx + 1) + 1)
        if nx > 1:
            left[:, 1:] = rightg[:, :-1]
        cell_edges = np.stack([bottom.reshape(-1), top, left.reshape(-1), right], axis=1).astype(np.int32)
        nedges = 3 * nx + 1 + (ny - 1) * (2 * nx + 1)
        ev = np.zeros((nedges, 2), dtype=np.int32)
        cv = cell_vertices
        for k, (a, b) in enumerate(CartesianModel.LOCAL_EDGES):
            ev[cell_edges[:, k], 0] = np.minimum(cv[:, a], cv[:, b])
            ev[cell_edges[:, k], 1] = np.maximum(cv[:, a], cv[:, b])
        return cell_edges, ev

    @staticmethod
    def _entity(left, right, bot, top):
        ent = np.full(left.shape, 9, dtype=np.int32)
        ent[bot] = 5
        ent[top] = 6
        ent[left] = 7
        ent[right] = 8
        ent[left & bot] = 1
        ent[right & bot] = 2
        ent[left & top] = 3
        ent[right & top] = 4
        return ent


class JitteredQuadModel(CartesianModel):
    """An UNSTRUCTURED quadrilateral mesh (what Gridap holds for a recombined Gmsh mesh): the topology, numbering and labels of
    the Cartesian model, interior vertices displaced by up to `jitter` of the cell size, `cartesian = False` -- so the bilinear
    geometry map (a different Jacobian at every quadrature point) of the generic kernels is exercised; P2 edge / cell nodes sit
    at the images of the reference nodes (edge midpoints, mean of the four vertices)."""
    cartesian = False

    def __init__(self, domain=(0.0, 1.0, 0.0, 1.0), partition=(6, 5), jitter=0.25, seed=0):
        super().__init__(domain, partition)
        nx, ny = self.partition
        rng = np.random.default_rng(seed)
        vi = np.tile(np.arange(nx + 1), ny + 1)
        vj = np.repeat(np.arange(ny + 1), nx + 1)
        interior = (vi > 0) & (vi < nx) & (vj > 0) & (vj < ny)
        d = (rng.random((interior.sum(), 2)) - 0.5) * 2.0 * jitter * np.array(self.h)
        self.vertex_coords = self.vertex_coords.copy()
        self.vertex_coords[interior] += d


class TriangleModel(DiscreteModel):
    """Triangle model with named tags, the shape GridapGmsh hands to the reference
    (examples/validation-kuehn/kuehn.jl:13-22).

    `vertex_entity`/`edge_entity` carry entity ids; `tag_to_entities` maps the physical names
    ("inner","outer","all",...) to entity id lists.  Each triangle's vertex ids are sorted
    ascending like GridapGmsh does (so det J may be negative; |det J| is used).
    """
    cell_type = TRI
    cartesian = False
    LOCAL_EDGES = [(0, 1), (0, 2), (1, 2)]

    def __init__(self, vertex_coords, triangles, vertex_entity, boundary_edge_entity,
                 tag_to_entities, interior_entity=0):
        self.vertex_coords = np.ascontiguousarray(vertex_coords, dtype=np.float64)
        tri = np.sort(np.asarray(triangles, dtype=np.int64), axis=1)
        self.cell_vertices = tri.astype(np.int32)
        self.cell_edges, self.edge_vertices = _first_encounter_edges(self.cell_vertices, self.LOCAL_EDGES)
        self.vertex_entity = np.asarray(vertex_entity, dtype=np.int32)
        # boundary_edge_entity: dict {(lo,hi): entity}
        ent = np.full(self.nedges, interior_entity, dtype=np.int32)
        if boundary_edge_entity:
            nvert = self.nverts
            keys = self.edge_vertices[:, 0].astype(np.int64) * nvert + self.edge_vertices[:, 1]
            srt = np.argsort(keys)
            bk = np.array([min(a, b) * nvert + max(a, b) for (a, b) in boundary_edge_entity.keys()],
                          dtype=np.int64)
            bv = np.array(list(boundary_edge_entity.values()), dtype=np.int32)
            pos = np.searchsorted(keys[srt], bk)
            assert np.all(keys[srt][pos] == bk), "boundary edge not present in the triangulation"
            ent[srt[pos]] = bv
        self.edge_entity = ent
        self.cell_entity = np.full(self.ncells, interior_entity, dtype=np.int32)
        self.tag_to_entities = dict(tag_to_entities)


# --------------------------------------------------------------------------------------
# FE spaces
# --------------------------------------------------------------------------------------
@dataclasses.dataclass
class H1P2Space:
    """P2 Lagrangian H1 space, scalar (ncomp=1) or vector (ncomp=2) (Solver.jl:81-86,92-96)."""
    ncomp: int
    n_free: int
    n_diri: int
    cell_dofs: np.ndarray          # [nc, ncomp*nn] int32, signed, 1-based, field-local; local dof = node + nn*comp
    diri_values: np.ndarray        # [n_diri] f64 ; value of Dirichlet dof -id-1
    node_face_dim: np.ndarray      # bookkeeping for post-processing: per global node (0 vertex, 1 edge, 2 cell)
    cell_nodes: np.ndarray         # [nc, nn] global node ids (0-based, vertices|edges|cells)
    node_coords: np.ndarray        # [nnodes,2]
    node_signed: np.ndarray        # [nnodes] signed scalar node id (free k>0 / diri -k)


def _face_tag_index(entity: np.ndarray, model: DiscreteModel, tags: Sequence) -> np.ndarray:
    """Gridap `get_face_tag_index`: position (0-based) of the LAST tag in `tags` whose entity
    set contains the face's entity, or -1 ('later tags overwrite earlier')."""
    out = np.full(entity.shape, -1, dtype=np.int32)
    for i, tag in enumerate(tags):
        ents = np.asarray(model.tag_entities(tag), dtype=np.int32)
        out[np.isin(entity, ents)] = i
    return out


def build_h1_p2(model: DiscreteModel, ncomp: int, diri_tags: Sequence,
                diri_exprs: Optional[Sequence[Union[float, Callable]]] = None,
                cells: Optional[slice] = None) -> H1P2Space:
    """`cells`: materialise the cell tables only for this contiguous cell range (multi-GPU row
    blocks); dof ids stay GLOBAL."""
    nv, ne, nc = model.nverts, model.nedges, model.ncells
    has_cell_node = model.cell_type == QUAD
    ent = [model.vertex_entity, model.edge_entity]
    if has_cell_node:
        ent.append(model.cell_entity)
    ent = np.concatenate(ent)
    tagidx = _face_tag_index(ent, model, diri_tags)
    diri = tagidx >= 0
    nfree_nodes = int((~diri).sum())
    ndiri_nodes = int(diri.sum())
    signed = np.where(diri, -np.cumsum(diri), np.cumsum(~diri)).astype(np.int64)  # 1-based scalar ids
    # node coordinates (vertices, edge midpoints, cell centres)
    vc = model.vertex_coords
    coords = [vc, 0.5 * (vc[model.edge_vertices[:, 0]] + vc[model.edge_vertices[:, 1]])]
    if has_cell_node:
        coords.append(vc[model.cell_vertices].mean(axis=1))
    coords = np.concatenate(coords)
    sl = cells if cells is not None else slice(0, nc)
    cn = [model.cell_vertices[sl].astype(np.int64), nv + model.cell_edges[sl].astype(np.int64)]
    if has_cell_node:
        cn.append((nv + ne + np.arange(nc, dtype=np.int64)[sl])[:, None])
    cell_nodes = np.concatenate(cn, axis=1)
    s = signed[cell_nodes]  # [nc, nn]
    comps = []
    for c in range(ncomp):
        comps.append(np.where(s > 0, ncomp * (s - 1) + c + 1, -(ncomp * (-s - 1) + c + 1)))
    cell_dofs = np.concatenate(comps, axis=1).astype(np.int32)
    # Dirichlet values
    dv = np.zeros(ndiri_nodes * ncomp)
    if diri_exprs is not None and ndiri_nodes:
        dnode = np.nonzero(diri)[0]
        for i, ex in enumerate(diri_exprs):
            sel = dnode[tagidx[dnode] == i]
            if sel.size == 0:
                continue
            if callable(ex):
                vals = np.array([ex(coords[k]) for k in sel], dtype=np.float64)
            else:
                vals = np.full(sel.size, float(ex))
            vals = vals.reshape(sel.size, -1)
            for c in range(ncomp):
                dv[ncomp * (-signed[sel] - 1) + c] = vals[:, c if vals.shape[1] > 1 else 0]
    dim = np.concatenate([np.zeros(nv, np.int8), np.ones(ne, np.int8)] +
                         ([np.full(nc, 2, np.int8)] if has_cell_node else []))
    return H1P2Space(ncomp, nfree_nodes * ncomp, ndiri_nodes * ncomp, np.ascontiguousarray(cell_dofs),
                     dv, dim, cell_nodes, coords, signed)


@dataclasses.dataclass
class Spaces:
    """The multi-field space X = [U, P, T] (Solver.jl:98-105) in the form the C ABI consumes."""
    model: DiscreteModel
    ref: RefTables
    U: H1P2Space
    T: H1P2Space
    cell_dofs_p: np.ndarray     # [nc,3] signed 1-based; last dof of the last cell is fixed (-1)
    n_free: np.ndarray          # [3] int64 (U,P,T)
    diri_p: np.ndarray          # [1] = 0.0
    cells: Optional[slice] = None   # local cell range of a multi-GPU row block, or local cell ids (own + ghost) of an unstructured partition (None: all cells)
    rows: Optional[tuple] = None    # (r0, r1, ghost_lo, ghost_hi) cell rows owned by this rank
    part: Optional[tuple] = None    # (c0, c1, ghost cell ids) of a contiguous-cell-range partition of an unstructured mesh

    @property
    def n_total(self):
        return int(self.n_free.sum())

    @property
    def offsets(self):
        return np.array([0, self.n_free[0], self.n_free[0] + self.n_free[1]], dtype=np.int64)


def build_spaces(model: DiscreteModel, V_diri_tags, T_diri_tags, T_diri_expressions,
                 cells: Optional[slice] = None) -> Spaces:
    ref = reference_tables(model.cell_type, 2)
    U = build_h1_p2(model, 2, V_diri_tags, None, cells)  # u_noslip = (0,0)  (Solver.jl:98)
    T = build_h1_p2(model, 1, T_diri_tags if len(T_diri_tags) else [], T_diri_expressions, cells)
    nc = model.ncells
    cdp = (3 * np.arange(nc, dtype=np.int64)[:, None] + np.arange(1, 4)[None, :]).astype(np.int32)
    cdp[nc - 1, 2] = -1  # constraint=:zeromean fixes the last free dof
    if cells is not None:
        cdp = cdp[cells]
    n_free = np.array([U.n_free, 3 * nc - 1, T.n_free], dtype=np.int64)
    sp = Spaces(model, ref, U, T, np.ascontiguousarray(cdp), n_free, np.zeros(1))
    sp.cells = cells
    return sp


def row_block(ny: int, rank: int, nranks: int):
    """Owner-computes row blocks: rank owns cell rows [r0, r1); its tables also carry one
    ghost cell row on each interior side (needed for the symbolic pattern of the interface rows).
    Returns (r0, r1, ghost_lo, ghost_hi)."""
    r0 = (ny * rank) // nranks
    r1 = (ny * (rank + 1)) // nranks
    return r0, r1, (1 if rank > 0 else 0), (1 if rank < nranks - 1 else 0)


def build_rank_spaces(model: "CartesianModel", bc_V, bc_Ttags, bc_Texpr, rank: int, nranks: int) -> Spaces:
    """Spaces of one rank of a Cartesian row-block partition: global dof ids, local cell tables
    (owned rows + ghost rows)."""
    nx, ny = model.partition
    r0, r1, glo, ghi = row_block(ny, rank, nranks)
    sl = slice((r0 - glo) * nx, (r1 + ghi) * nx)
    sp = build_spaces(model, bc_V, bc_Ttags, bc_Texpr, cells=sl)
    sp.rows = (r0, r1, glo, ghi)
    return sp


def cell_block(model: DiscreteModel, rank: int, nranks: int):
    """Unstructured meshes ("contiguous ranges"): rank owns the cells [c0, c1); its GHOSTS are the other ranks'
    cells that share a vertex (hence possibly a P2 node) with one of them, ascending.  Returns (c0, c1, ghosts)."""
    nc = model.ncells
    c0, c1 = (nc * rank) // nranks, (nc * (rank + 1)) // nranks
    mark = np.zeros(model.nverts, dtype=bool)
    mark[model.cell_vertices[c0:c1].reshape(-1)] = True
    touch = mark[model.cell_vertices].any(axis=1)
    touch[c0:c1] = False
    return c0, c1, np.nonzero(touch)[0].astype(np.int64)


def build_rank_spaces_cells(model: DiscreteModel, bc_V, bc_Ttags, bc_Texpr, rank: int, nranks: int) -> Spaces:
    """Spaces of one rank of a contiguous-cell-range partition of an unstructured mesh: global dof ids, local cell tables
    (own cells, then ghost cells).  `sp.part` = (c0, c1, ghosts)."""
    c0, c1, ghosts = cell_block(model, rank, nranks)
    cells = np.concatenate([np.arange(c0, c1, dtype=np.int64), ghosts])
    sp = build_spaces(model, bc_V, bc_Ttags, bc_Texpr, cells=cells)
    sp.part = (c0, c1, ghosts)
    return sp


def owned_dofs_cells(sp: Spaces) -> np.ndarray:
    """0-based global free-dof ids owned by the rank of `sp` (build_rank_spaces_cells): a dof belongs to the rank of its
    lowest-numbered incident cell.  The local tables hold every incident cell of every dof of an own cell."""
    c0, c1, ghosts = sp.part
    gid = np.concatenate([np.arange(c0, c1, dtype=np.int64), ghosts])
    off = sp.offsets
    out = []
    for tab, o in ((sp.U.cell_dofs, off[0]), (sp.cell_dofs_p, off[1]), (sp.T.cell_dofs, off[2])):
        d = tab.astype(np.int64)
        g = np.broadcast_to(gid[:, None], d.shape)[d > 0]
        ids = d[d > 0] - 1 + o
        order = np.lexsort((g, ids))
        ids, g = ids[order], g[order]
        first = np.ones(ids.size, dtype=bool)
        first[1:] = ids[1:] != ids[:-1]
        sel = first & (g >= c0) & (g < c1)
        out.append(ids[sel])
    return np.concatenate(out)


# --------------------------------------------------------------------------------------
# iterates used by tests and the bench
# --------------------------------------------------------------------------------------
def splitmix_iterate(n: int, seed: int = 1234) -> np.ndarray:
    """x[k] = (splitmix64(seed + k) >> 11) * 2^-53, k = 0-based global free-dof id."""
    with np.errstate(over="ignore"):
        z = (np.arange(n, dtype=np.uint64) + np.uint64(seed))
        z = z + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (2.0 ** -53)


def smooth_iterate(sp: Spaces) -> np.ndarray:
    """Nodal interpolation of u = (∂yψ, -∂xψ), ψ = x²(1-x)²y²(1-y)², T = 1-x, p = 0."""
    x = np.zeros(sp.n_total)
    off = sp.offsets

    def fill(space, fun, o):
        s = space.node_signed
        free = s > 0
        xy = space.node_coords[free]
        vals = fun(xy[:, 0], xy[:, 1])
        for c in range(space.ncomp):
            x[o + space.ncomp * (s[free] - 1) + c] = vals[c]

    def uv(X, Y):
        f = lambda t: t * t * (1 - t) * (1 - t)
        df = lambda t: 2 * t * (1 - t) * (1 - 2 * t)
        return [f(X) * df(Y), -df(X) * f(Y)]

    fill(sp.U, uv, off[0])
    fill(sp.T, lambda X, Y: [1.0 - X], off[2])
    return x


# --------------------------------------------------------------------------------------
# self-generated triangle models (the reference's .msh files are Git-LFS pointers)
# --------------------------------------------------------------------------------------
def annulus_model(R1=5.0 / 8, R2=13.0 / 8, nr=8, nt=32, jitter=0.0, seed=0) -> TriangleModel:
    """Concentric annulus following examples/meshes/2.6/co-annulus_unstructured.geo:12-48
    (R1=5/8, R2=13/8; tags "inner","outer","all","heatfunction_zero_points").

    Rings of nt vertices, each ring quad split by a FIXED diagonal (never union-jack, no triangle
    with two boundary edges: SURVEY 0.4 shows the P2/P1disc pair is singular otherwise).  `jitter`
    perturbs interior vertices to make the triangulation irregular.  nt must be divisible by 4
    so that (0,+-R1), (0,+-R2) are vertices (the four Physical Points of the .geo, lines 25-28).
    Entities: 1 inner curve, 2 outer curve, 3 inner points, 4 outer points, 5 interior.
    """
    assert nt % 4 == 0 and nr >= 2
    rng = np.random.default_rng(seed)
    r = np.linspace(R1, R2, nr + 1)
    th = 2 * np.pi * np.arange(nt) / nt
    R, TH = np.meshgrid(r, th, indexing="ij")  # [nr+1, nt]
    if jitter > 0:
        dr = (R2 - R1) / nr
        R[1:-1] += jitter * dr * (rng.random((nr - 1, nt)) - 0.5)
        TH[1:-1] += jitter * (2 * np.pi / nt) * (rng.random((nr - 1, nt)) - 0.5)
    X, Y = R * np.cos(TH), R * np.sin(TH)
    coords = np.stack([X.reshape(-1), Y.reshape(-1)], axis=1)
    vid = lambda i, j: i * nt + (j % nt)
    tris = []
    for i in range(nr):
        for j in range(nt):
            a, b, c, d = vid(i, j), vid(i + 1, j), vid(i, j + 1), vid(i + 1, j + 1)
            tris.append((a, b, d))
            tris.append((a, d, c))
    vent = np.full((nr + 1) * nt, 5, dtype=np.int32)
    vent[:nt] = 1
    vent[nr * nt:] = 2
    for j in (nt // 4, 3 * nt // 4):  # theta = pi/2, 3pi/2 -> (0,+-R)
        vent[vid(0, j)] = 3
        vent[vid(nr, j)] = 4
    bedges = {}
    for j in range(nt):
        bedges[(min(vid(0, j), vid(0, j + 1)), max(vid(0, j), vid(0, j + 1)))] = 1
        bedges[(min(vid(nr, j), vid(nr, j + 1)), max(vid(nr, j), vid(nr, j + 1)))] = 2
    tags = {"inner": [1, 3], "outer": [2, 4], "all": [1, 2, 3, 4], "heatfunction_zero_points": [3, 4],
            "mesh": [5]}
    return TriangleModel(coords, np.array(tris), vent, bedges, tags, interior_entity=5)


def square_tri_model(n=8, jitter=0.0, seed=0) -> TriangleModel:
    """Unit square split into 2 n^2 triangles with a fixed diagonal; corner cells get the diagonal
    that avoids a triangle with two boundary edges.  Entities as CartesianModel (1..9)."""
    rng = np.random.default_rng(seed)
    xs = np.linspace(0, 1, n + 1)
    X, Y = np.meshgrid(xs, xs, indexing="xy")
    if jitter > 0:
        X[1:-1, 1:-1] += jitter / n * (rng.random((n - 1, n - 1)) - 0.5)
        Y[1:-1, 1:-1] += jitter / n * (rng.random((n - 1, n - 1)) - 0.5)
    coords = np.stack([X.reshape(-1), Y.reshape(-1)], axis=1)
    vid = lambda i, j: i + j * (n + 1)
    tris = []
    for j in range(n):
        for i in range(n):
            a, b, c, d = vid(i, j), vid(i + 1, j), vid(i, j + 1), vid(i + 1, j + 1)
            if n > 1 and ((i == 0 and j == n - 1) or (i == n - 1 and j == 0)):
                tris += [(a, b, c), (b, d, c)]  # diagonal b-c: with a-d the corner triangle would have two boundary edges
            else:
                tris += [(a, b, d), (a, d, c)]  # diagonal a-d
    vi, vj = np.tile(np.arange(n + 1), n + 1), np.repeat(np.arange(n + 1), n + 1)
    vent = CartesianModel._entity(vi == 0, vi == n, vj == 0, vj == n)
    bedges = {}
    for k in range(n):
        bedges[(vid(k, 0), vid(k + 1, 0))] = 5
        bedges[(vid(k, n), vid(k + 1, n))] = 6
        bedges[(vid(0, k), vid(0, k + 1))] = 7
        bedges[(vid(n, k), vid(n, k + 1))] = 8
    tags = {k: [k] for k in range(1, 10)}
    return TriangleModel(coords, np.array(tris), vent, bedges, tags, interior_entity=9)


def _global_ids_of_rows(sp: Spaces, local_rows: slice, nodes=None) -> np.ndarray:
    """Sorted unique 0-based global free-dof ids of the cells in the given LOCAL cell rows
    (optionally only of the local P2 nodes `nodes`, U both comps + T)."""
    nx = sp.model.partition[0]
    sl = slice(local_rows.start * nx, local_rows.stop * nx)
    off = sp.offsets
    u, t, p = sp.U.cell_dofs[sl], sp.T.cell_dofs[sl], sp.cell_dofs_p[sl]
    if nodes is not None:
        nn = sp.ref.nn
        u = u[:, list(nodes) + [nn + k for k in nodes]]
        t = t[:, list(nodes)]
        p = p[:, :0]
    ids = np.concatenate([u[u > 0].astype(np.int64) - 1, p[p > 0].astype(np.int64) - 1 + off[1],
                          t[t > 0].astype(np.int64) - 1 + off[2]])
    return np.unique(ids)


def interface_dofs(sp: Spaces, side: int) -> np.ndarray:
    """Global dof ids on the rank's lower (side 0) / upper (side 1) interface line, ascending --
    the order both neighbours pack the exchange buffer in (csrc/gm_multi.cuh)."""
    r0, r1, glo, ghi = sp.rows
    if side == 0:
        return _global_ids_of_rows(sp, slice(glo, glo + 1), nodes=(0, 1, 4)) if glo else np.zeros(0, np.int64)
    last = glo + (r1 - r0) - 1
    return _global_ids_of_rows(sp, slice(last, last + 1), nodes=(2, 3, 5)) if ghi else np.zeros(0, np.int64)


def owned_dofs(sp: Spaces) -> np.ndarray:
    """Owner-computes rule: a dof belongs to the lowest row block (cell range) touching it."""
    if getattr(sp, "part", None) is not None:
        return np.sort(owned_dofs_cells(sp))
    if getattr(sp, "rows", None) is None:
        return np.arange(sp.n_total, dtype=np.int64)
    r0, r1, glo, ghi = sp.rows
    mine = _global_ids_of_rows(sp, slice(glo, glo + (r1 - r0)))
    return np.setdiff1d(mine, interface_dofs(sp, 0), assume_unique=True)


# --------------------------------------------------------------------------------------
# post-processing spaces and tables (/root/reference/src/Solver.jl:197-292): P1 stream / heat function spaces,
# Measure(Ω, 4) for the heat function, nodal gradients for the entropy interpolation
# --------------------------------------------------------------------------------------
@dataclasses.dataclass
class P1Space:
    """`TestFESpace(model, ReferenceFE(lagrangian, Float64, 1), conformity=:H1, dirichlet_tags=...)` (Solver.jl:200-204,
    239-245): one dof per vertex; on quads the Q1 basis."""
    n_free: int
    n_diri: int
    cell_dofs: np.ndarray     # [nc, nv] int32 signed 1-based
    diri_values: np.ndarray   # [n_diri]


def build_p1(model: DiscreteModel, diri_tags: Sequence, diri_value: float = 0.0) -> P1Space:
    tagidx = _face_tag_index(model.vertex_entity, model, diri_tags) if len(diri_tags) else np.full(model.nverts, -1, np.int32)
    diri = tagidx >= 0
    signed = np.where(diri, -np.cumsum(diri), np.cumsum(~diri)).astype(np.int32)
    return P1Space(int((~diri).sum()), int(diri.sum()), np.ascontiguousarray(signed[model.cell_vertices]),
                   np.full(int(diri.sum()), float(diri_value)))


def quadrature(cell_type: int, degree: int):
    """(points [nq,2], weights [nq]) of Gridap's `Measure(Ω, degree)` on the reference cell for degree 2 and 4.
    QUAD: tensor Gauss-Legendre with ceil((degree+1)/2) points per direction (x fastest).  TRI degree 4: the 6-point
    Strang-Fix / Dunavant rule [Gridap ◐: unverified which of its tabulated degree-4 simplex rules is picked]."""
    if cell_type == QUAD or degree == 2:
        t = reference_tables(cell_type, degree)
        return t.qp, t.w
    if degree == 4:
        a1, a2 = 0.445948490915965, 0.091576213509771
        w1, w2 = 0.223381589678011 / 2, 0.109951743655322 / 2
        qp = np.array([[a1, a1], [1 - 2 * a1, a1], [a1, 1 - 2 * a1], [a2, a2], [1 - 2 * a2, a2], [a2, 1 - 2 * a2]])
        return qp, np.array([w1, w1, w1, w2, w2, w2])
    raise ValueError("quadrature degree %d not tabulated" % degree)


def vertex_basis(cell_type: int, x: float, y: float):
    """P1 / Q1 (= geometry) basis values [nv] and reference gradients [nv,2] at one reference point."""
    if cell_type == QUAD:
        g = np.array([(1 - x) * (1 - y), x * (1 - y), (1 - x) * y, x * y])
        dg = np.array([[-(1 - y), -(1 - x)], [(1 - y), -x], [-y, (1 - x)], [y, x]])
    else:
        g = np.array([1 - x - y, x, y])
        dg = np.array([[-1.0, -1.0], [1.0, 0.0], [0.0, 1.0]])
    return g, dg


@dataclasses.dataclass
class PostTables:
    """Tabulated bases for the post-processing assemblies: per rule r in (degree 2, degree 4): w[r] [nq], g[r] [nq,nv],
    dg[r] [nq,nv,2] (vertex basis = geometry basis), phi[r] [nq,nn], dphi[r] [nq,nn,2] (P2 basis); at the nn P2 nodes:
    dphi_n [nn,nn,2] (P2 gradients), dg_n [nn,nv,2] (geometry gradients)."""
    w: list
    g: list
    dg: list
    phi: list
    dphi: list
    dphi_n: np.ndarray
    dg_n: np.ndarray


def post_tables(cell_type: int) -> PostTables:
    w, g, dg, phi, dphi = [], [], [], [], []
    for degree in (2, 4):
        qp, ww = quadrature(cell_type, degree)
        w.append(np.ascontiguousarray(ww))
        g.append(np.array([vertex_basis(cell_type, x, y)[0] for x, y in qp]))
        dg.append(np.array([vertex_basis(cell_type, x, y)[1] for x, y in qp]))
        phi.append(np.array([shape_at(cell_type, x, y)[0] for x, y in qp]))
        dphi.append(np.array([shape_at(cell_type, x, y)[1] for x, y in qp]))
    nodes = reference_tables(cell_type, 2).nodes
    dphi_n = np.array([shape_at(cell_type, x, y)[1] for x, y in nodes])
    dg_n = np.array([vertex_basis(cell_type, x, y)[1] for x, y in nodes])
    return PostTables(w, g, dg, phi, dphi, dphi_n, dg_n)
