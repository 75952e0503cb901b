"""ctypes binding of libgemotion_b200.so (include/gemotion_b200.h).

This is the Python twin of the Julia `ccall` stub in INTEGRATION.md: same entry points, same
argument meaning, rc != 0 -> exception (Gridap style).  There is no fallback: if the shared
library or a B200 is missing every call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgemotion_b200.so")

GM_QUAD, GM_TRI = 0, 1
GM_LAYOUT_CSC, GM_LAYOUT_CSR = 0, 1
GM_JACOBIAN, GM_RESIDUAL, GM_CHECK_FINITE, GM_NO_EXCHANGE = 1, 2, 4, 8
GM_PATH_AUTO, GM_PATH_SCATTER, GM_PATH_CART, GM_PATH_GATHER = 0, 1, 2, 3
GM_KERNEL_AUTO, GM_KERNEL_MMA, GM_KERNEL_GATH, GM_KERNEL_GATH1, GM_KERNEL_SWEEP = 0, 1, 2, 3, 4

EXPORTS = ["gm_version", "gm_last_error", "gm_create", "gm_destroy", "gm_set_params", "gm_pattern",
           "gm_get_pattern", "gm_get_rows", "gm_assemble", "gm_assemble_device", "gm_host_register", "gm_host_unregister",
           "gm_set_stream", "gm_timing", "gm_info", "gm_nccl_unique_id", "gm_comm_init",
           "gm_iface_count", "gm_iface_pack", "gm_iface_unpack_add",
           "gm_post_create", "gm_post_sizes", "gm_post_get_pattern", "gm_post_assemble", "gm_post_entropy"]


class GmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("gemotion_b200 error %d: %s" % (code, msg))
        self.code = code


class gm_desc(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("cell_type", C.c_int32),
        ("ncells", C.c_int64), ("nverts", C.c_int64),
        ("cartesian", C.c_int32), ("nx", C.c_int32), ("ny", C.c_int32),
        ("origin", C.c_double * 2), ("h", C.c_double * 2),
        ("coords", C.c_void_p), ("cell_vertices", C.c_void_p),
        ("vertex_index_base", C.c_int32), ("layout", C.c_int32),
        ("cell_dofs_u", C.c_void_p), ("cell_dofs_p", C.c_void_p), ("cell_dofs_t", C.c_void_p),
        ("n_free", C.c_int64 * 3), ("n_diri", C.c_int64 * 3),
        ("diri_u", C.c_void_p), ("diri_p", C.c_void_p), ("diri_t", C.c_void_p),
        ("nq", C.c_int32), ("nn", C.c_int32), ("nv", C.c_int32),
        ("w", C.c_void_p), ("phi", C.c_void_p), ("dphi", C.c_void_p), ("psi", C.c_void_p), ("dgphi", C.c_void_p),
        ("device", C.c_int32), ("path", C.c_int32),
        ("rank", C.c_int32), ("nranks", C.c_int32),
        ("cell_begin", C.c_int64), ("ncells_global", C.c_int64),
        ("ghost_lo", C.c_int32), ("ghost_hi", C.c_int32),
        ("kernel", C.c_int32), ("xchunk", C.c_int32),
        ("n_own_cells", C.c_int64), ("ghost_cell_ids", C.c_void_p),
    ]


class gm_post_desc(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32), ("nq2", C.c_int32), ("nq4", C.c_int32),
        ("cell_dofs_psi", C.c_void_p), ("n_free_psi", C.c_int64), ("n_diri_psi", C.c_int64), ("diri_psi", C.c_void_p),
        ("cell_dofs_pi", C.c_void_p), ("n_free_pi", C.c_int64), ("n_diri_pi", C.c_int64), ("diri_pi", C.c_void_p),
        ("w2", C.c_void_p), ("g2", C.c_void_p), ("dg2", C.c_void_p), ("phi2", C.c_void_p), ("dphi2", C.c_void_p),
        ("w4", C.c_void_p), ("g4", C.c_void_p), ("dg4", C.c_void_p), ("phi4", C.c_void_p), ("dphi4", C.c_void_p),
        ("dphi_nodes", C.c_void_p), ("dg_nodes", C.c_void_p),
        ("cell_nodes", C.c_void_p), ("n_nodes", C.c_int64),
    ]


class gm_times(C.Structure):
    _fields_ = [("h2d_ms", C.c_double), ("kernel_ms", C.c_double), ("d2h_ms", C.c_double),
                ("exchange_ms", C.c_double), ("launches", C.c_int64),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64)]


_lib = None


def build(force=False, verbose=False):
    """Compile csrc/ for sm_100a into libgemotion_b200.so (in-tree)."""
    src = os.path.join(HERE, "csrc")
    newest = max(os.path.getmtime(os.path.join(src, f)) for f in os.listdir(src))
    newest = max(newest, os.path.getmtime(os.path.join(HERE, "..", "include", "gemotion_b200.h")))
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < newest:
        subprocess.check_call(["make", "-C", src] + ([] if verbose else ["-s"]))
    return LIB_PATH


def _preload_nccl():
    """libgemotion_b200.so needs libnccl.so.2.  PyTorch bundles its own (newer) copy; whichever
    copy is mapped first serves both, and torch fails to import against the older system one.  So
    map torch's copy first when it exists (no torch import needed)."""
    import glob
    import sysconfig
    for base in {sysconfig.get_paths()["purelib"], sysconfig.get_paths()["platlib"]}:
        for cand in glob.glob(os.path.join(base, "nvidia", "nccl", "lib", "libnccl.so.2")):
            try:
                C.CDLL(cand, mode=C.RTLD_GLOBAL)
                return
            except OSError:
                pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GmError(-2, "libgemotion_b200.so not built (run __graft_entry__.build()); there is no CPU path")
        _preload_nccl()
        L = C.CDLL(LIB_PATH)
        L.gm_last_error.restype = C.c_char_p
        L.gm_destroy.restype = None
        for name in EXPORTS:
            getattr(L, name)  # raises if a symbol is missing
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise GmError(rc, lib().gm_last_error().decode())


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Handle:
    """Owns one gm_handle."""

    def __init__(self, sp, layout="csc", device=0, path=GM_PATH_AUTO, rank=0, nranks=1, kernel=GM_KERNEL_AUTO, xchunk=0):
        """`sp` from fe.build_spaces (whole mesh) or fe.build_rank_spaces (one row block of `nranks`)."""
        m, r = sp.model, sp.ref
        c = np.ascontiguousarray
        d = gm_desc()
        d.struct_size = C.sizeof(gm_desc)
        d.cell_type = m.cell_type
        d.ncells, d.nverts = m.ncells, m.nverts
        keep = {}
        rows = getattr(sp, "rows", None)
        if m.cartesian:
            d.cartesian = 1
            d.nx, d.ny = m.partition
            if rows is not None:
                r0, r1, glo, ghi = rows
                d.ny = r1 - r0 + glo + ghi
                d.ncells = d.nx * d.ny
                d.ghost_lo, d.ghost_hi = glo, ghi
                d.cell_begin = (r0 - glo) * d.nx
            d.origin[0], d.origin[1] = m.origin
            d.h[0], d.h[1] = m.h
        else:
            keep["coords"] = c(m.vertex_coords, dtype=np.float64)
            part = getattr(sp, "part", None)   # (c0, c1, ghost cell ids): one rank of a contiguous-cell-range partition
            keep["cv"] = c(m.cell_vertices if sp.cells is None else m.cell_vertices[sp.cells], dtype=np.int32)
            if part is not None:
                c0, c1, ghosts = part
                keep["ghosts"] = c(ghosts, dtype=np.int64)
                d.ncells = int(c1 - c0 + len(ghosts))
                d.n_own_cells = int(c1 - c0)
                d.cell_begin = int(c0)
                d.ghost_cell_ids = _p(keep["ghosts"]) if len(ghosts) else None
            d.coords, d.cell_vertices = _p(keep["coords"]), _p(keep["cv"])
        d.vertex_index_base = 0
        d.layout = GM_LAYOUT_CSR if layout == "csr" else GM_LAYOUT_CSC
        keep["du"], keep["dp"], keep["dt"] = c(sp.U.cell_dofs, dtype=np.int32), c(sp.cell_dofs_p, dtype=np.int32), c(sp.T.cell_dofs, dtype=np.int32)
        d.cell_dofs_u, d.cell_dofs_p, d.cell_dofs_t = _p(keep["du"]), _p(keep["dp"]), _p(keep["dt"])
        nd = [sp.U.diri_values.size, 1, sp.T.diri_values.size]
        for i in range(3):
            d.n_free[i] = int(sp.n_free[i])
            d.n_diri[i] = int(nd[i])
        keep["vu"], keep["vp"], keep["vt"] = c(sp.U.diri_values), c(sp.diri_p), c(sp.T.diri_values)
        d.diri_u = _p(keep["vu"]) if nd[0] else None
        d.diri_p = _p(keep["vp"])
        d.diri_t = _p(keep["vt"]) if nd[2] else None
        d.nq, d.nn, d.nv = r.nq, r.nn, r.nv
        keep["w"], keep["phi"], keep["dphi"], keep["psi"], keep["dgphi"] = c(r.w), c(r.phi), c(r.dphi), c(r.psi), c(r.dgphi)
        d.w, d.phi, d.dphi, d.psi, d.dgphi = _p(keep["w"]), _p(keep["phi"]), _p(keep["dphi"]), _p(keep["psi"]), _p(keep["dgphi"])
        d.device, d.path = device, path
        d.kernel, d.xchunk = kernel, xchunk
        d.rank, d.nranks, d.ncells_global = rank, nranks, m.ncells
        if rows is None and getattr(sp, "part", None) is None:
            d.cell_begin = 0
        self._h = C.c_void_p()
        self.layout = layout
        self.n = sp.n_total
        self.nnz = None
        _check(lib().gm_create(C.byref(d), C.byref(self._h)))
        del keep  # the library copied everything

    def close(self):
        if self._h:
            lib().gm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, Pr=1.0, Ra=1.0, n=1.0, reg=1e-3, g=(0.0, 1.0), jac_scaling=1.0):
        _check(lib().gm_set_params(self._h, C.c_double(Pr), C.c_double(Ra), C.c_double(n), C.c_double(reg),
                                   C.c_double(g[0]), C.c_double(g[1]), C.c_double(jac_scaling)))

    def pattern(self):
        nnz = C.c_int64()
        _check(lib().gm_pattern(self._h, C.byref(nnz)))
        self.nnz = nnz.value
        return self.nnz

    def get_pattern(self, index_base=0, with_idx=True):
        if self.nnz is None:
            self.pattern()
        ptr = np.empty(self.n + 1, dtype=np.int64)
        idx = np.empty(self.nnz, dtype=np.int64) if with_idx else None
        _check(lib().gm_get_pattern(self._h, _p(ptr), _p(idx), C.c_int32(index_base)))
        return ptr, idx

    def get_rows(self, rows, d_nz=0, d_b=0):
        """Lists of the selected major dofs (0-based global ids): (ptr[nrows+1], idx, val | None, b | None);
        `d_nz` / `d_b` are DEVICE pointers of the value array / residual to read (0: skipped)."""
        rows = np.ascontiguousarray(rows, dtype=np.int64)
        cap = int(rows.size) * 96 + 64
        ptr = np.zeros(rows.size + 1, dtype=np.int64)
        idx = np.zeros(cap, dtype=np.int64)
        val = np.zeros(cap) if d_nz else None
        b = np.zeros(rows.size) if d_b else None
        _check(lib().gm_get_rows(self._h, C.c_int64(rows.size), _p(rows), _p(ptr), _p(idx), _p(val), _p(b), C.c_int64(cap),
                                 C.c_void_p(d_nz or 0), C.c_void_p(d_b or 0)))
        n = int(ptr[-1])
        return ptr, idx[:n], (val[:n] if val is not None else None), b

    def assemble(self, x, nzval=None, resid=None, flags=None):
        """Host buffers in/out (numpy float64, C-contiguous)."""
        if flags is None:
            flags = (GM_JACOBIAN if nzval is not None else 0) | (GM_RESIDUAL if resid is not None else 0)
        assert x.dtype == np.float64 and x.flags.c_contiguous and x.size == self.n
        _check(lib().gm_assemble(self._h, _p(x), _p(nzval), _p(resid), C.c_uint32(flags)))

    def assemble_device(self, d_x, d_nz, d_b, flags, sync=True):
        """Device pointers (ints)."""
        _check(lib().gm_assemble_device(self._h, C.c_void_p(d_x), C.c_void_p(d_nz or 0), C.c_void_p(d_b or 0),
                                        C.c_uint32(flags), C.c_int32(1 if sync else 0)))

    # ---- post-processing assemblies (Solver.jl:197-292)
    def post_create(self, sp, psi_space, pi_space, tabs):
        """`psi_space` / `pi_space`: fe.P1Space of the stream / heat function; `tabs`: fe.PostTables."""
        c = np.ascontiguousarray
        k = dict(dpsi=c(psi_space.cell_dofs, dtype=np.int32), dpi=c(pi_space.cell_dofs, dtype=np.int32),
                 vpsi=c(psi_space.diri_values, dtype=np.float64), vpi=c(pi_space.diri_values, dtype=np.float64),
                 cn=c(sp.T.cell_nodes, dtype=np.int32), dpn=c(tabs.dphi_n), dgn=c(tabs.dg_n))
        for r in (0, 1):
            for nm, arr in (("w", tabs.w), ("g", tabs.g), ("dg", tabs.dg), ("phi", tabs.phi), ("dphi", tabs.dphi)):
                k["%s%d" % (nm, r)] = c(arr[r], dtype=np.float64)
        d = gm_post_desc()
        d.struct_size = C.sizeof(gm_post_desc)
        d.nq2, d.nq4 = tabs.w[0].size, tabs.w[1].size
        d.cell_dofs_psi, d.n_free_psi, d.n_diri_psi, d.diri_psi = _p(k["dpsi"]), psi_space.n_free, psi_space.n_diri, (_p(k["vpsi"]) if psi_space.n_diri else None)
        d.cell_dofs_pi, d.n_free_pi, d.n_diri_pi, d.diri_pi = _p(k["dpi"]), pi_space.n_free, pi_space.n_diri, (_p(k["vpi"]) if pi_space.n_diri else None)
        d.w2, d.g2, d.dg2, d.phi2, d.dphi2 = _p(k["w0"]), _p(k["g0"]), _p(k["dg0"]), _p(k["phi0"]), _p(k["dphi0"])
        d.w4, d.g4, d.dg4, d.phi4, d.dphi4 = _p(k["w1"]), _p(k["g1"]), _p(k["dg1"]), _p(k["phi1"]), _p(k["dphi1"])
        d.dphi_nodes, d.dg_nodes = _p(k["dpn"]), _p(k["dgn"])
        d.cell_nodes, d.n_nodes = _p(k["cn"]), int(sp.T.node_coords.shape[0])
        _check(lib().gm_post_create(self._h, C.byref(d)))
        self._post_nodes = d.n_nodes

    def post_assemble(self, which, x):
        """(A scipy csc, b) of the stream (0) / heat (1) function problem at the solution x."""
        import scipy.sparse as sps
        n, nnz = C.c_int64(), C.c_int64()
        _check(lib().gm_post_sizes(self._h, C.c_int32(which), C.byref(n), C.byref(nnz)))
        ptr, idx = np.empty(n.value + 1, dtype=np.int64), np.empty(nnz.value, dtype=np.int64)
        _check(lib().gm_post_get_pattern(self._h, C.c_int32(which), _p(ptr), _p(idx), C.c_int32(0)))
        nz, b = np.empty(nnz.value), np.empty(n.value)
        _check(lib().gm_post_assemble(self._h, C.c_int32(which), _p(np.ascontiguousarray(x, dtype=np.float64)), _p(nz), _p(b)))
        return sps.csc_matrix((nz, idx, ptr), shape=(n.value, n.value)), b

    def post_entropy(self, x):
        """(Sth nodal, Sfl nodal, Sth_tot, Sfl_tot, Be_avg)."""
        sth, sfl, tot = np.empty(self._post_nodes), np.empty(self._post_nodes), np.empty(3)
        _check(lib().gm_post_entropy(self._h, _p(np.ascontiguousarray(x, dtype=np.float64)), _p(sth), _p(sfl), _p(tot)))
        return sth, sfl, float(tot[0]), float(tot[1]), float(tot[2])

    # ---- multi-GPU row blocks
    def comm_init(self, id128: bytes):
        buf = C.create_string_buffer(id128, 128)
        _check(lib().gm_comm_init(self._h, buf))

    def iface_count(self, side):
        n = C.c_int64()
        _check(lib().gm_iface_count(self._h, C.c_int32(side), C.byref(n)))
        return n.value

    def iface_pack(self, d_nz, d_b, d_buf, flags):
        _check(lib().gm_iface_pack(self._h, C.c_void_p(d_nz or 0), C.c_void_p(d_b or 0), C.c_void_p(d_buf), C.c_uint32(flags)))

    def iface_unpack_add(self, d_nz, d_b, d_buf, flags):
        _check(lib().gm_iface_unpack_add(self._h, C.c_void_p(d_nz or 0), C.c_void_p(d_b or 0), C.c_void_p(d_buf), C.c_uint32(flags)))

    def set_stream(self, stream_ptr):
        _check(lib().gm_set_stream(self._h, C.c_void_p(stream_ptr)))

    def timing(self):
        t = gm_times()
        _check(lib().gm_timing(self._h, C.byref(t)))
        return {k: getattr(t, k) for k, _ in gm_times._fields_}

    def info(self):
        n, p, c, b = C.c_int64(), C.c_int32(), C.c_int32(), C.c_int64()
        _check(lib().gm_info(self._h, C.byref(n), C.byref(p), C.byref(c), C.byref(b)))
        return dict(n_owned=n.value, active_path=p.value, ncolors=c.value, device_bytes=b.value)


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _check(lib().gm_nccl_unique_id(buf))
    return buf.raw


def host_register(a: np.ndarray):
    _check(lib().gm_host_register(_p(a), C.c_int64(a.nbytes)))


def host_unregister(a: np.ndarray):
    _check(lib().gm_host_unregister(_p(a)))
