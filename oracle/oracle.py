"""ctypes wrapper around oracle/libnsf_oracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module (see the header of nsf_oracle.c).  It follows /root/reference/src/Solver.jl:111-152
as executed by Gridap 0.18.9; parity status is documented in nsf_oracle.c and DESIGN.md.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "libnsf_oracle.so")
    src = os.path.join(_HERE, "nsf_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libnsf_oracle.so"])
    return so


class _Problem(C.Structure):
    _fields_ = [
        ("cell_type", C.c_int32), ("cartesian", C.c_int32),
        ("ncells", C.c_int64), ("nverts", C.c_int64),
        ("hx", C.c_double), ("hy", C.c_double),
        ("coords", C.c_void_p), ("cell_vertices", C.c_void_p),
        ("cell_dofs_u", C.c_void_p), ("cell_dofs_p", C.c_void_p), ("cell_dofs_t", C.c_void_p),
        ("n_free", C.c_int64 * 3),
        ("diri_u", C.c_void_p), ("diri_p", C.c_void_p), ("diri_t", C.c_void_p),
        ("nq", C.c_int32), ("nn", C.c_int32), ("nv", C.c_int32),
        ("w", C.c_void_p), ("phi", C.c_void_p), ("dphi", C.c_void_p), ("psi", C.c_void_p),
        ("dgphi", C.c_void_p),
        ("Pr", C.c_double), ("Ra", C.c_double), ("n", C.c_double), ("reg", C.c_double),
        ("gx", C.c_double), ("gy", C.c_double), ("jac_scaling", C.c_double),
    ]


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_pattern.restype = C.c_int64
        _LIB.orc_assemble.restype = C.c_int
        _LIB.orc_assemble_colored.restype = C.c_int
        _LIB.orc_num_threads.restype = C.c_int
        _LIB.orc_rows.restype = C.c_int64
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Oracle:
    """CPU oracle for one (spaces, params) problem."""

    def __init__(self, sp, Pr=1.0, Ra=1.0, n=1.0, reg=1e-3, g=(0.0, 1.0), jac_scaling=1.0):
        self.sp = sp
        m = sp.model
        r = sp.ref
        c = np.ascontiguousarray
        self._keep = dict(
            coords=c(m.vertex_coords, dtype=np.float64),
            cv=c(m.cell_vertices if (m.cartesian or getattr(sp, "cells", None) is None) else m.cell_vertices[sp.cells], dtype=np.int32),
            du=c(sp.U.cell_dofs, dtype=np.int32), dp=c(sp.cell_dofs_p, dtype=np.int32),
            dt=c(sp.T.cell_dofs, dtype=np.int32),
            vu=c(sp.U.diri_values if sp.U.diri_values.size else np.zeros(1)),
            vp=c(sp.diri_p), vt=c(sp.T.diri_values if sp.T.diri_values.size else np.zeros(1)),
            w=c(r.w), phi=c(r.phi), dphi=c(r.dphi), psi=c(r.psi), dgphi=c(r.dgphi))
        k = self._keep
        P = _Problem()
        P.cell_type = m.cell_type
        P.cartesian = 1 if m.cartesian else 0
        P.ncells, P.nverts = int(k["du"].shape[0]), m.nverts  # (a row-block rank holds the cell tables of its rows only)
        P.hx, P.hy = (m.h if m.cartesian else (0.0, 0.0))
        P.coords, P.cell_vertices = _p(k["coords"]), _p(k["cv"])
        P.cell_dofs_u, P.cell_dofs_p, P.cell_dofs_t = _p(k["du"]), _p(k["dp"]), _p(k["dt"])
        for i in range(3):
            P.n_free[i] = int(sp.n_free[i])
        P.diri_u, P.diri_p, P.diri_t = _p(k["vu"]), _p(k["vp"]), _p(k["vt"])
        P.nq, P.nn, P.nv = r.nq, r.nn, r.nv
        P.w, P.phi, P.dphi, P.psi, P.dgphi = _p(k["w"]), _p(k["phi"]), _p(k["dphi"]), _p(k["psi"]), _p(k["dgphi"])
        P.Pr, P.Ra, P.n, P.reg, P.gx, P.gy, P.jac_scaling = Pr, Ra, n, reg, g[0], g[1], jac_scaling
        self.P = P
        self.n = sp.n_total
        self.ld = 3 * r.nn + 3
        self._pat = {}

    def set_params(self, **kw):
        for k, v in kw.items():
            setattr(self.P, k, v)

    def cell(self, c, x, jac=True):
        r = np.zeros(self.ld)
        J = np.zeros((self.ld, self.ld)) if jac else None
        _lib().orc_cell(C.byref(self.P), C.c_int64(c), _p(np.ascontiguousarray(x)), _p(r), _p(J))
        return r, J

    def pattern(self, layout="csc"):
        """(ptr[n+1], idx[nnz]) 0-based int64.  layout 'csc' (Gridap default) or 'csr'."""
        if layout not in self._pat:
            tr = 1 if layout == "csr" else 0
            ptr = np.zeros(self.n + 1, dtype=np.int64)
            nnz = _lib().orc_pattern(C.byref(self.P), tr, _p(ptr), None)
            idx = np.zeros(nnz, dtype=np.int64)
            _lib().orc_pattern(C.byref(self.P), tr, _p(ptr), _p(idx))
            self._pat[layout] = (ptr, idx)
        return self._pat[layout]

    def assemble(self, x, layout="csc", jac=True, res=True, colors=None):
        ptr, idx = self.pattern(layout)
        tr = 1 if layout == "csr" else 0
        x = np.ascontiguousarray(x, dtype=np.float64)
        nz = np.empty(idx.size) if jac else None
        b = np.empty(self.n) if res else None
        if colors is None:
            rc = _lib().orc_assemble(C.byref(self.P), tr, _p(ptr), _p(idx), _p(x), _p(nz), _p(b))
        else:
            cptr, ccells = colors
            rc = _lib().orc_assemble_colored(C.byref(self.P), tr, _p(ptr), _p(idx), _p(x), _p(nz), _p(b),
                                             len(cptr) - 1, _p(cptr), _p(ccells))
        if rc != 0:
            raise RuntimeError("oracle: entry missing from pattern")
        return nz, b

    def rows(self, rows, x, layout="csc"):
        """Row oracle: complete lists (sorted-unique minor ids, values) and residual entries of the
        sampled major dofs `rows` (ascending, 0-based global), rebuilt from orc_cell over the incident cells.
        Returns (ptr[nrows+1], idx, val, b)."""
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        assert rows.size == 0 or np.all(np.diff(rows) > 0)
        tr = 1 if layout == "csr" else 0
        cap = int(rows.size) * 96 + 64
        ptr = np.zeros(rows.size + 1, dtype=np.int64)
        idx = np.zeros(cap, dtype=np.int64)
        val = np.zeros(cap)
        b = np.zeros(rows.size)
        x = np.ascontiguousarray(x, dtype=np.float64)
        n = _lib().orc_rows(C.byref(self.P), tr, C.c_int64(rows.size), _p(rows), _p(x), _p(ptr), _p(idx), _p(val), _p(b),
                            C.c_int64(cap))
        if n < 0:
            raise RuntimeError("oracle: row buffer too small")
        return ptr, idx[:n], val[:n], b

    @staticmethod
    def set_num_threads(n):
        _lib().orc_set_num_threads(C.c_int(int(n)))

    @staticmethod
    def num_threads():
        return _lib().orc_num_threads()
