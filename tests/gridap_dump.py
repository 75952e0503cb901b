"""Reader / writer of the dump format of tools/dump_gridap_golden.jl -- TEST INFRASTRUCTURE.

A dump directory holds Gridap's own assembly (colptr, rowval, nzval, b at the splitmix iterate x) and every table the
C ABI consumes (signed cell dof ids, Dirichlet values, quadrature weights, tabulated bases, geometry), all as Gridap
produced them.  `write_self_dump` writes the SAME format from this repo's restatement (fe.py + the oracle): it pins
nothing (source = "restatement"), it only lets the loader test run where Julia is absent.
"""
import json
import os

import numpy as np

FILES = {"colptr": ("colptr.i64", np.int64), "rowval": ("rowval.i64", np.int64), "nzval": ("nzval.f64", np.float64),
         "b": ("b.f64", np.float64), "x": ("x.f64", np.float64),
         "du": ("cell_dofs_U.i32", np.int32), "dp": ("cell_dofs_P.i32", np.int32), "dt": ("cell_dofs_T.i32", np.int32),
         "vu": ("diri_U.f64", np.float64), "vp": ("diri_P.f64", np.float64), "vt": ("diri_T.f64", np.float64),
         "w": ("w.f64", np.float64), "phi": ("phi.f64", np.float64), "dphi": ("dphi.f64", np.float64),
         "psi": ("psi.f64", np.float64), "dgphi": ("dgphi.f64", np.float64),
         "coords": ("coords.f64", np.float64), "cellv": ("cell_vertices.i32", np.int32)}


def load(d):
    meta = json.load(open(os.path.join(d, "meta.json")))
    out = {"meta": meta}
    for k, (fn, dt) in FILES.items():
        p = os.path.join(d, fn)
        out[k] = np.fromfile(p, dtype=dt) if os.path.exists(p) else None
    return out


def write_self_dump(d, sp, prm, x, orc):
    """Same files from fe.py tables + the oracle's assembly (1-based indices like Julia)."""
    os.makedirs(d, exist_ok=True)
    m, r = sp.model, sp.ref
    ptr, idx = orc.pattern("csc")
    nz, b = orc.assemble(x, "csc")
    arrs = {"colptr": ptr + 1, "rowval": idx + 1, "nzval": nz, "b": b, "x": x,
            "du": sp.U.cell_dofs, "dp": sp.cell_dofs_p, "dt": sp.T.cell_dofs,
            "vu": sp.U.diri_values, "vp": sp.diri_p, "vt": sp.T.diri_values,
            "w": r.w, "phi": r.phi, "dphi": r.dphi, "psi": r.psi, "dgphi": r.dgphi}
    if not m.cartesian:
        arrs["coords"] = m.vertex_coords
        arrs["cellv"] = m.cell_vertices + 1
    for k, v in arrs.items():
        fn, dt = FILES[k]
        np.ascontiguousarray(v, dtype=dt).tofile(os.path.join(d, fn))
    p = dict(Pr=1.0, Ra=1.0, n=1.0, reg=1e-3, g=(0.0, 1.0), jac_scaling=1.0)
    p.update(prm)
    meta = {"source": "restatement", "cell_type": "QUAD" if m.cell_type == 0 else "TRI", "cartesian": bool(m.cartesian),
            "nx": int(m.partition[0]) if m.cartesian else 0, "ny": int(m.partition[1]) if m.cartesian else 0,
            "origin": list(m.origin) if m.cartesian else [0.0, 0.0], "h": list(m.h) if m.cartesian else [0.0, 0.0],
            "ncells": int(m.ncells), "nverts": int(m.nverts), "n_free": [int(v) for v in sp.n_free],
            "n_diri": [int(sp.U.diri_values.size), 1, int(sp.T.diri_values.size)], "nq": int(r.nq), "nn": int(r.nn), "nv": int(r.nv),
            "index_base": 1, "vertex_index_base": 1, "nnz": int(idx.size),
            "params": {"Pr": p["Pr"], "Ra": p["Ra"], "n": p["n"], "reg": p["reg"], "g": list(p["g"]), "jac_scaling": p["jac_scaling"]}}
    json.dump(meta, open(os.path.join(d, "meta.json"), "w"))


class DumpSpaces:
    """Duck-typed stand-in for fe.Spaces built from a dump: exactly the attributes capi.Handle / oracle.Oracle read, filled
    with GRIDAP'S tables (so DOF maps are Gridap's by construction)."""

    def __init__(self, D):
        from _pkg import gm
        fe = gm.fe
        me = D["meta"]
        nn, nv, nq = me["nn"], me["nv"], me["nq"]
        nc = me["ncells"]

        class M:
            pass

        class S:
            pass

        m = M()
        m.cell_type = 0 if me["cell_type"] == "QUAD" else 1
        m.cartesian = bool(me["cartesian"])
        m.ncells, m.nverts = nc, me["nverts"]
        if m.cartesian:
            m.partition, m.origin, m.h = (me["nx"], me["ny"]), tuple(me["origin"]), tuple(me["h"])
            m.vertex_coords = np.zeros((1, 2))
            m.cell_vertices = np.zeros((1, nv), dtype=np.int32)
        else:
            m.vertex_coords = D["coords"].reshape(-1, 2)
            m.cell_vertices = (D["cellv"].reshape(nc, nv) - me.get("vertex_index_base", 1)).astype(np.int32)
        self.model = m
        self.ref = fe.RefTables(m.cell_type, D["w"], np.zeros((nq, 2)), D["phi"].reshape(nq, nn), D["dphi"].reshape(nq, nn, 2),
                                D["psi"].reshape(nq, 3), np.zeros((nq, nv)), D["dgphi"].reshape(nq, nv, 2), np.zeros((nn, 2)))
        self.U, self.T = S(), S()
        self.U.cell_dofs = D["du"].reshape(nc, 2 * nn)
        self.T.cell_dofs = D["dt"].reshape(nc, nn)
        self.U.diri_values = D["vu"] if D["vu"] is not None else np.zeros(0)
        self.T.diri_values = D["vt"] if D["vt"] is not None else np.zeros(0)
        self.cell_dofs_p = D["dp"].reshape(nc, 3)
        self.diri_p = D["vp"] if D["vp"] is not None and D["vp"].size else np.zeros(1)
        self.n_free = np.array(me["n_free"], dtype=np.int64)
        self.rows = None
        self.cells = None

    @property
    def n_total(self):
        return int(self.n_free.sum())

    @property
    def offsets(self):
        return np.array([0, self.n_free[0], self.n_free[0] + self.n_free[1]], dtype=np.int64)
