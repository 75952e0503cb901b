"""CPU check of the generic gather path's kernels (gm_gather_kernel.cuh: k_cellmat + k_lists): the kernel source is
compiled as host C++ (tests/emu/cuda_emu.h: one pthread per CUDA thread, shared memory and workspaces poisoned) and
its output compared with the oracle on triangle and quadrilateral meshes.  The adjacency and the compact cell->nnz map
are built here from the oracle's pattern the way gm_pattern.cuh builds them on the device.  Test infrastructure only --
the product runs the same source on sm_100a; the GPU parity proper is tests/test_gpu_parity.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import cases
from cases import fe
from oracle.oracle import Oracle

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "emu")
ROOT = os.path.dirname(HERE)
TOL = 1e-12


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(EMU_DIR, "_emu_gather.so")
    srcs = [os.path.join(EMU_DIR, "emu_gather.cpp"), os.path.join(EMU_DIR, "cuda_emu.h"),
            os.path.join(ROOT, "gemotion.jl_b200", "csrc", "gm_gather_kernel.cuh"),
            os.path.join(ROOT, "gemotion.jl_b200", "csrc", "gm_const.h")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-pthread", "-Wno-unknown-pragmas", "-o", so, srcs[0]])
    lib = C.CDLL(so)
    lib.emu_gather_run.restype = C.c_int
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def global_dofs(sp):
    """Signed global ids exactly as k_combine_dofs builds them: > 0 free (1-based), < 0 -> dvals[-id-1]."""
    off = sp.offsets
    nd = [sp.U.diri_values.size, 1, sp.T.diri_values.size]
    du, dp, dt = sp.U.cell_dofs.astype(np.int64), sp.cell_dofs_p.astype(np.int64), sp.T.cell_dofs.astype(np.int64)
    gp = np.where(dp > 0, dp + off[1], dp - nd[0])
    gt = np.where(dt > 0, dt + off[2], dt - (nd[0] + nd[1]))
    g = np.concatenate([du, gp, gt], axis=1).astype(np.int32)
    dv = np.concatenate([sp.U.diri_values, sp.diri_p, sp.T.diri_values]).astype(np.float64)
    return np.ascontiguousarray(g), np.ascontiguousarray(dv)


def gather_tables(g, nn, ptr, idx, n):
    """adjacency dof -> cell*ld + local (ascending), grank in the compact [U majors][T majors] layout, grankp of the P majors."""
    nc, ld = g.shape
    field = np.array([0] * (2 * nn) + [1] * 3 + [2] * nn)
    touch = np.array([[1, 1, 1], [1, 0, 0], [1, 0, 1]], dtype=bool)
    su, st = ld, 3 * nn   # segment lengths of a velocity / temperature major
    off_t = 2 * nn * su
    nec = off_t + nn * st
    grank = np.full((nc, nec), 0xFFFF, dtype=np.uint16)
    grankp = np.full((nc, 6 * nn), 0xFFFF, dtype=np.uint16)
    adj = [[] for _ in range(n)]
    for c in range(nc):
        for M in range(ld):
            gM = g[c, M]
            if gM <= 0:
                continue
            adj[gM - 1].append(c * ld + M)
            lst = idx[ptr[gM - 1]:ptr[gM]]
            fM = field[M]
            tab, off = (grankp, (M - 2 * nn) * 2 * nn) if fM == 1 else (grank, M * su if fM == 0 else off_t + (M - 2 * nn - 3) * st)
            for l in range(ld):
                if not touch[fM, field[l]]:
                    continue
                s = l - 3 if (fM == 2 and l >= 2 * nn) else l
                if g[c, l] > 0:
                    k = np.searchsorted(lst, g[c, l] - 1)
                    assert k < lst.size and lst[k] == g[c, l] - 1
                    tab[c, off + s] = k
    adjptr = np.zeros(n + 1, dtype=np.int64)
    adjptr[1:] = np.cumsum([len(a) for a in adj])
    adjv = np.array([v for a in adj for v in sorted(a)], dtype=np.int32)
    return np.ascontiguousarray(grank), np.ascontiguousarray(grankp), adjptr, adjv


def run_emu(emu, sp, prm, layout, x, flags=3, grid=3):
    m = sp.model
    nn = 9 if m.cell_type == fe.QUAD else 6
    g, dv = global_dofs(sp)
    orc = Oracle(sp, **prm)
    ptr, idx = orc.pattern(layout)   # (a rank of a partition: the pattern of its local tables)
    n = sp.n_total
    grank, grankp, adjptr, adj = gather_tables(g, nn, ptr, idx, n)
    nz, b = np.full(int(ptr[-1]), 7e77), np.full(n, 7e77)
    p = dict(Pr=1.0, Ra=1.0, n=1.0, reg=1e-3, g=(0.0, 1.0), jac_scaling=1.0)
    p.update(prm)
    pv = np.array([p["Pr"], p["Ra"], p["n"], p["reg"], p["g"][0], p["g"][1], p["jac_scaling"]])
    ref = sp.ref
    cart = 1 if m.cartesian else 0
    if cart:
        hx, hy = m.h
        coords = cellv = None
    else:
        hx = hy = 0.0
        coords = np.ascontiguousarray(m.vertex_coords, dtype=np.float64)
        cellv = np.ascontiguousarray(m.cell_vertices if sp.cells is None else m.cell_vertices[sp.cells], dtype=np.int32)
    part = getattr(sp, "part", None)
    nown, cbeg, ghosts = (part[1] - part[0], part[0], np.ascontiguousarray(part[2], dtype=np.int64)) if part is not None else (0, 0, None)
    off = sp.offsets
    rc = emu.emu_gather_run(C.c_int(nn), C.c_int64(g.shape[0]), C.c_int64(off[1]), C.c_int64(off[2]), C.c_int64(n), _p(g), _p(dv),
                            _p(coords), _p(cellv), _p(ptr), _p(grank), _p(grankp), C.c_int(grid), _p(adjptr), _p(adj), _p(x), _p(nz), _p(b), C.c_uint(flags),
                            C.c_int(1 if layout == "csr" else 0), _p(pv), C.c_int(cart), C.c_double(hx), C.c_double(hy),
                            _p(np.ascontiguousarray(ref.w)), _p(np.ascontiguousarray(ref.phi)), _p(np.ascontiguousarray(ref.dphi)),
                            _p(np.ascontiguousarray(ref.psi)), _p(np.ascontiguousarray(ref.dgphi)),
                            C.c_int64(nown), C.c_int64(cbeg), _p(ghosts))
    assert rc == 0
    return orc, ptr, idx, nz, b


MODELS = {
    "sqtri": lambda: cases.spaces(fe.square_tri_model(5, jitter=0.3, seed=2), cases.TURAN),
    "annulus": lambda: cases.spaces(fe.annulus_model(nr=3, nt=12, jitter=0.3, seed=1), cases.KUEHN),
    "cart7x5": lambda: cases.cart(7, 5, bc=cases.TURAN),
    "cart1x1": lambda: cases.cart(1, 1, bc=cases.SIMPLE),
    "sqtri1": lambda: cases.spaces(fe.square_tri_model(1), cases.TURAN),   # two triangles: two free velocity dofs, one ragged batch, one group per field
    "sqtri16": lambda: cases.spaces(fe.square_tri_model(16, jitter=0.2, seed=5), cases.TURAN),   # 512 cells: 32 batches, 11 per CTA
    "fan14": lambda: cases.spaces(cases.fan_model(14), cases.FAN),                                # valence 14: list > 96, cnt > 10
    "jquad6x5": lambda: cases.spaces(fe.JitteredQuadModel((0, 2, 0, 1), (6, 5), jitter=0.25, seed=2), cases.TURAN),   # bilinear geometry
    "cart12x8": lambda: cases.cart(12, 8, bc=cases.WAVE),                                       # 96 cells: 6 full batches
}
PARAMS = [dict(Pr=100.0, Ra=1e5, n=0.6), dict(Pr=0.706, Ra=4.7e4, n=1.0), dict(Pr=10.0, Ra=1e4, n=1.4, g=(0.3, -0.8), jac_scaling=1e-6)]


@pytest.mark.parametrize("name", list(MODELS))
@pytest.mark.parametrize("pi", range(len(PARAMS)))
@pytest.mark.parametrize("layout", ["csc", "csr"])
def test_gather_kernels_match_oracle(emu, name, pi, layout):
    sp = MODELS[name]()
    x = fe.splitmix_iterate(sp.n_total, 77 + pi)
    orc, ptr, idx, nz, b = run_emu(emu, sp, PARAMS[pi], layout, x)
    onz, ob = orc.assemble(x, layout)
    e1 = cases.block_scaled_err(nz, onz, cases.nz_blocks(sp, ptr, idx))
    e2 = cases.block_scaled_err(b, ob, cases.res_blocks(sp))
    assert e1 <= TOL and e2 <= TOL, (e1, e2)


def test_gather_kernels_jacobian_only_and_residual_only(emu):
    sp = MODELS["sqtri"]()
    x = fe.splitmix_iterate(sp.n_total, 5)
    orc, ptr, idx, nz, b = run_emu(emu, sp, PARAMS[0], "csc", x, flags=3)
    _, _, _, nz1, b1 = run_emu(emu, sp, PARAMS[0], "csc", x, flags=1)
    _, _, _, nz2, b2 = run_emu(emu, sp, PARAMS[0], "csc", x, flags=2)
    assert np.array_equal(nz1, nz) and np.all(b1 == 7e77)
    assert np.array_equal(b2, b) and np.all(nz2 == 7e77)


@pytest.mark.parametrize("nranks", [2, 3])
def test_gather_kernels_on_cell_range_ranks(emu, nranks):
    """gm_desc.n_own_cells / ghost_cell_ids: every rank assembles, from its own + ghost cells, exactly the global lists and
    residual entries of the dofs it owns (pressure rows only from own cells, records of unowned majors without incident cells)."""
    model = fe.square_tri_model(7, jitter=0.25, seed=3)
    spg = cases.spaces(model, cases.TURAN)
    prm = PARAMS[2]
    x = fe.splitmix_iterate(spg.n_total, 21)
    orc = Oracle(spg, **prm)
    optr, oidx = orc.pattern("csc")
    onz, ob = orc.assemble(x, "csc")
    got_nz, got_b = np.full(onz.size, np.nan), np.full(ob.size, np.nan)
    for r in range(nranks):
        sp = fe.build_rank_spaces_cells(model, cases.TURAN["V"], cases.TURAN["Ttags"], cases.TURAN["Texpr"], r, nranks)
        _, ptr, idx, nz, b = run_emu(emu, sp, prm, "csc", x)
        for dof in fe.owned_dofs(sp):
            assert np.array_equal(idx[ptr[dof]:ptr[dof + 1]], oidx[optr[dof]:optr[dof + 1]])
            got_nz[optr[dof]:optr[dof + 1]] = nz[ptr[dof]:ptr[dof + 1]]
            got_b[dof] = b[dof]
    assert cases.block_scaled_err(got_nz, onz, cases.nz_blocks(spg, optr, oidx)) <= TOL
    assert cases.block_scaled_err(got_b, ob, cases.res_blocks(spg)) <= TOL
