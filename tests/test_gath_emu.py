"""CPU check of the index logic of k_gath (gm_gath_kernel.cuh): the kernel source is compiled as host C++
(tests/emu/cuda_emu.h: one pthread per CUDA thread, asynchronous copies delayed to their waits, shared memory
poisoned) and its output compared with the oracle.  Test infrastructure only -- the product runs the same
source on sm_100a; the GPU parity proper is tests/test_gpu_parity.py.
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


def _build():
    so = os.path.join(EMU_DIR, "_emu_gath.so")
    srcs = [os.path.join(EMU_DIR, "emu_gath.cpp"), os.path.join(EMU_DIR, "cuda_emu.h"),
            os.path.join(ROOT, "gemotion.jl_b200", "csrc", "gm_gath_kernel.cuh"),
            os.path.join(ROOT, "gemotion.jl_b200", "csrc", "gm_mma_kernel.cuh"),
            os.path.join(ROOT, "gemotion.jl_b200", "csrc", "gm_mma_tiles.h"),
            os.path.join(ROOT, "gemotion.jl_b200", "csrc", "gm_cart_tables.h")]
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-pthread", "-Wno-unknown-pragmas",
                               "-o", so, srcs[0]])
    lib = C.CDLL(so)
    lib.emu_gath_run.restype = C.c_int
    return lib


@pytest.fixture(scope="module")
def emu():
    return _build()


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


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


FIELD = np.array([0] * 18 + [1] * 3 + [2] * 9)
TOUCH = np.array([[1, 1, 1], [1, 0, 0], [1, 0, 1]], dtype=bool)  # [major field][minor field]: PP, PT, TP never touched


def rank_classes(g, ptr, idx):
    """rank[c][minor][major] = position of the minor dof inside the major dof's list (0xFF: none), and the
    classes of identical tables -- what gm_pattern.cuh / gm_cart.cuh build on the device."""
    nc = g.shape[0]
    rank = np.full((nc, 30, 30), 0xFF, dtype=np.uint8)
    for c in range(nc):
        for M in range(30):
            gM = g[c, M]
            if gM <= 0:
                continue
            lst = idx[ptr[gM - 1]:ptr[gM]]
            for l in range(30):
                if g[c, l] > 0 and TOUCH[FIELD[M], FIELD[l]]:
                    k = np.searchsorted(lst, g[c, l] - 1)
                    assert k < lst.size and lst[k] == g[c, l] - 1
                    rank[c, l, M] = k
    tabs, cls = np.unique(rank.reshape(nc, 900), axis=0, return_inverse=True)
    return np.ascontiguousarray(tabs), np.ascontiguousarray(cls.astype(np.int32).ravel())


def local_pattern(g, n):
    """Sorted-unique lists over the touched blocks of the cells in `g` (the pattern a rank builds from its
    local table, ghost rows included) -- the same for CSC and CSR because the touched-block set is symmetric."""
    lists = [set() for _ in range(n)]
    for c in range(g.shape[0]):
        for M in range(30):
            if g[c, M] <= 0:
                continue
            lists[g[c, M] - 1].update(int(g[c, l]) - 1 for l in range(30) if g[c, l] > 0 and TOUCH[FIELD[M], FIELD[l]])
    ptr = np.zeros(n + 1, dtype=np.int64)
    ptr[1:] = np.cumsum([len(l) for l in lists])
    idx = np.array([v for l in lists for v in sorted(l)], dtype=np.int64)
    return ptr, idx


def run_emu(emu, sp, prm, layout, x, rows=None, ng=2, flags=3, xch=256):
    m = sp.model
    nx, ny_glob = m.partition
    g, dv = global_dofs(sp)
    if rows is None:
        orc = Oracle(sp, **prm)
        ptr, idx = orc.pattern(layout)
    else:
        orc = None
        ptr, idx = local_pattern(g, sp.n_total)
    ctab, cls = rank_classes(g, ptr, idx)
    if rows is None:
        ny, rb, re = ny_glob, 0, ny_glob
    else:
        r0, r1, glo, ghi = rows
        ny, rb, re = r1 - r0 + glo + ghi, glo, glo + (r1 - r0)
    assert g.shape[0] == nx * ny
    nz = np.full(idx.size, np.nan)
    b = np.full(sp.n_total, np.nan)
    p = dict(Pr=1.0, Ra=1.0, n=1.0, reg=1e-3, g=(0.0, 1.0), jac_scaling=1.0)
    p.update(prm)
    pv = np.array([p["Pr"], p["Ra"], p["n"], p["reg"], p["g"][0], p["g"][1], p["jac_scaling"]])
    r = sp.ref
    c = np.ascontiguousarray
    w, phi, dphi, psi = c(r.w, dtype=np.float64), c(r.phi, dtype=np.float64), c(r.dphi, dtype=np.float64), c(r.psi, dtype=np.float64)
    xx = c(x, dtype=np.float64)
    diag = np.zeros(4, dtype=np.int64)
    rc = emu.emu_gath_run(C.c_int(nx), C.c_int(ny), C.c_int(rb), C.c_int(re), _p(g), _p(dv), _p(ptr), _p(cls),
                          C.c_int(ctab.shape[0]), _p(ctab), _p(xx), _p(nz), _p(b), C.c_uint(flags),
                          C.c_int(1 if layout == "csr" else 0), _p(pv), C.c_double(m.h[0]), C.c_double(m.h[1]),
                          _p(w), _p(phi), _p(dphi), _p(psi), C.c_int(ng), C.c_int(xch), _p(diag))
    assert rc == 0
    assert diag[0] == 0, "misaligned TMA bulk copy"
    return orc, ptr, idx, nz, b, g


PARAMS = [
    dict(Pr=0.7, Ra=1e3, n=1.0),
    dict(Pr=100.0, Ra=1e5, n=0.6),
    dict(Pr=100.0, Ra=1e5, n=1.4, jac_scaling=1e-6),
]
CART_CASES = [
    ("1x1", 1, 1, cases.TURAN), ("2x3_wave", 2, 3, cases.WAVE), ("8x8_simple", 8, 8, cases.SIMPLE),
    ("13x7_turan", 13, 7, cases.TURAN), ("5x19_turan", 5, 19, cases.TURAN), ("130x3_wave", 130, 3, cases.WAVE), ("259x2_turan", 259, 2, cases.TURAN), ("515x1_simple", 515, 1, cases.SIMPLE),
]


@pytest.mark.parametrize("kernel", ["mma", "gath"])
@pytest.mark.parametrize("case", CART_CASES, ids=[c[0] for c in CART_CASES])
@pytest.mark.parametrize("pi", range(len(PARAMS)))
@pytest.mark.parametrize("layout", ["csc", "csr"])
def test_gather_kernel_logic_matches_oracle(emu, kernel, case, pi, layout):
    """`ng` = 0 selects k_mma (gm_mma_kernel.cuh: the tensor instruction, the mbarrier and the warp roles are emulated
    by cuda_emu.h), 1 / 2 the trios of k_gath."""
    _, nx, ny, bc = case
    sp = cases.cart(nx, ny, bc=bc, domain=(0, 1.5, 0, 1))
    x = fe.splitmix_iterate(sp.n_total, 1234)
    orc, ptr, idx, nz, b, _ = run_emu(emu, sp, PARAMS[pi], layout, x, ng=0 if kernel == "mma" else 1 + (pi % 2), xch=(256, 128, 32)[pi])
    onz, ob = orc.assemble(x, layout)
    assert not np.isnan(nz).any() and not np.isnan(b).any(), "every list position / residual entry must be written"
    assert cases.block_scaled_err(nz, onz, cases.nz_blocks(sp, ptr, idx)) <= TOL
    assert cases.block_scaled_err(b, ob, cases.res_blocks(sp)) <= TOL


@pytest.mark.parametrize("ng", [0, 2], ids=["mma", "gath"])
def test_gather_kernel_jacobian_only_and_residual_only(emu, ng):
    sp = cases.cart(6, 5, bc=cases.TURAN)
    x = fe.splitmix_iterate(sp.n_total, 7)
    _, _, _, nz, b, _ = run_emu(emu, sp, PARAMS[1], "csc", x, ng=ng)
    _, _, _, nz1, b1, _ = run_emu(emu, sp, PARAMS[1], "csc", x, flags=1, ng=ng)
    _, _, _, nz2, b2, _ = run_emu(emu, sp, PARAMS[1], "csc", x, flags=2, ng=ng)
    assert np.array_equal(nz1, nz) and np.isnan(b1).all()
    assert np.array_equal(b2, b) and np.isnan(nz2).all()


@pytest.mark.parametrize("ng", [0, 2], ids=["mma", "gath"])
@pytest.mark.parametrize("nranks,nx,ny", [(2, 9, 8), (3, 5, 19)])
@pytest.mark.parametrize("layout", ["csc", "csr"])
def test_gather_kernel_row_blocks(emu, ng, nranks, nx, ny, layout):
    """Each rank assembles its owned cells only; lists of all nodes touching them (interface lines partial)."""
    prm = PARAMS[1]
    model = fe.CartesianModel((0, 1.5, 0, 1), (nx, ny))
    full = cases.spaces(model, cases.TURAN)
    x = fe.splitmix_iterate(full.n_total, 1234)
    forc = Oracle(full, **prm)
    for r in range(nranks):
        sp = fe.build_rank_spaces(model, cases.TURAN["V"], cases.TURAN["Ttags"], cases.TURAN["Texpr"], r, nranks)
        _, ptr, idx, nz, b, g = run_emu(emu, sp, prm, layout, x, rows=sp.rows, ng=ng)
        r0, r1, glo, ghi = sp.rows
        # reference: scatter of the owned cells through the local pattern
        rnz, rb = np.zeros(idx.size), np.zeros(full.n_total)
        touched = np.zeros(full.n_total, dtype=bool)
        for c in range(glo * nx, (glo + r1 - r0) * nx):
            rc, J = forc.cell((r0 - glo) * nx + c, x)
            for i in range(30):
                gi = g[c, i]
                if gi <= 0:
                    continue
                touched[gi - 1] = True
                rb[gi - 1] += rc[i]
                for j in range(30):
                    gj = g[c, j]
                    if gj <= 0 or not TOUCH[FIELD[i], FIELD[j]]:
                        continue
                    M, mnr = (gi, gj) if layout == "csr" else (gj, gi)
                    lst = idx[ptr[M - 1]:ptr[M]]
                    rnz[ptr[M - 1] + np.searchsorted(lst, mnr - 1)] += J[i, j]
        rows_t = np.flatnonzero(touched)
        sel = np.concatenate([np.arange(ptr[d], ptr[d + 1]) for d in rows_t])
        assert not np.isnan(nz[sel]).any() and not np.isnan(b[rows_t]).any()
        scale = np.abs(rnz).max()
        assert np.max(np.abs(nz[sel] - rnz[sel])) <= 1e-12 * scale
        assert np.max(np.abs(b[rows_t] - rb[rows_t])) <= 1e-12 * np.abs(rb).max()
