"""CPU tests of the oracle itself (no GPU): golden anchor, finite differences, size regressions."""
import os

import numpy as np
import pytest

import cases
from cases import fe
from oracle.oracle import Oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# /root/reference/examples/study1/conv_study_square.csv:2-3 (Nu_bot_avg)
NU_GOLD = {62: 6.670431887977048, 86: 6.646034152207981}


def test_golden_nu_62(gm):
    """The committed iterate is the oracle's own Newton solution (tests/golden/make_square_golden.py);
    it must (i) have a ~zero oracle residual and (ii) reproduce the reference's CSV value."""
    sp = cases.cart(62, bc=cases.WAVE)
    x = np.load(os.path.join(GOLD, "square62_wave_n06_x.npy"))
    assert x.size == sp.n_total
    _, b = Oracle(sp, Pr=100.0, Ra=1e4, n=0.6).assemble(x, jac=False)
    assert np.max(np.abs(b)) <= 1e-8  # ftol of study1.jl:82
    nu = gm.solver.nu_bot_avg(sp, x)
    assert abs(nu - NU_GOLD[62]) / NU_GOLD[62] < 1e-13, nu
    # the same CSV row also pins the VELOCITY field: Sth_max (|grad T|^2) and Sfl_max (viscous dissipation),
    # src/Solver.jl:281-289 sampled as in examples/study1/study1.jl:251-271
    sth, sfl = gm.solver.entropy_fields_max(sp, x)
    assert abs(sth - 104.37198885839547) / 104.37198885839547 < 1e-12, sth
    assert abs(sfl - 2406.6836086605012) / 2406.6836086605012 < 5e-12, sfl


@pytest.mark.skipif(os.environ.get("GM_SLOW") != "1", reason="2 min Newton run; GM_SLOW=1")
def test_golden_newton_62(gm):
    from orc_operator import OracleOperator
    sp = cases.cart(62, bc=cases.WAVE)
    op = OracleOperator(sp, Pr=100.0, Ra=1e4, n=0.6)
    res = gm.solver.newton(op, np.zeros(op.n), ftol=1e-8, xtol=1e-50, iterations=50)
    assert res.f_converged and res.iterations == 20
    assert abs(gm.solver.nu_bot_avg(sp, res.x) - NU_GOLD[62]) / NU_GOLD[62] < 1e-13


def test_triangle_discretisation_converges_to_the_reference_nusselt_number(gm):
    """'TRI path pinned only by self-consistency (..., quad-vs-tri convergence to the same Nu)'.  The study-1 problem
    (wave BCs, n = 0.6, Ra = 1e4, Pr = 100) on 2 x 32 x 32 jittered triangles: Newton on the oracle converges and Nu_bot_avg lies
    within 1 % of the finest value of the reference's quadrilateral convergence study (conv_study_square.csv:6: 6.6189; the
    study itself still moves by 0.06 % between its last two meshes).  A wrong triangle quadrature rule or node order changes
    Nu at O(1)."""
    from orc_operator import OracleOperator
    prm = dict(Pr=100.0, Ra=1e4, n=0.6)
    spq = cases.cart(12, bc=cases.WAVE)
    rq = gm.solver.newton(OracleOperator(spq, **prm), np.zeros(spq.n_total), ftol=1e-8, xtol=1e-50, iterations=80)
    assert abs(gm.solver.nu_bot_avg_generic(spq, rq.x) - gm.solver.nu_bot_avg(spq, rq.x)) < 1e-12   # same functional on quads
    spt = cases.spaces(fe.square_tri_model(32, jitter=0.25, seed=1), cases.WAVE)
    rt = gm.solver.newton(OracleOperator(spt, **prm), np.zeros(spt.n_total), ftol=1e-8, xtol=1e-50, iterations=80)
    assert rt.residual_norm < 1e-6
    nu = gm.solver.nu_bot_avg_generic(spt, rt.x)
    assert abs(nu - 6.618933375624088) / 6.618933375624088 < 0.01, nu


@pytest.mark.parametrize("n", [0.6, 1.0, 1.4])
@pytest.mark.parametrize("kind", ["quad", "jquad", "tri"])
def test_jacobian_is_derivative_of_residual(kind, n):
    sp = (cases.cart(5, 4, bc=cases.WAVE) if kind == "quad" else
          cases.spaces(fe.JitteredQuadModel((0, 1, 0, 1), (5, 4), jitter=0.25, seed=1), cases.WAVE) if kind == "jquad" else
          cases.spaces(fe.annulus_model(nr=3, nt=8, jitter=0.2), cases.KUEHN))
    orc = Oracle(sp, Pr=7.0, Ra=300.0, n=n)
    rng = np.random.default_rng(0)
    x = rng.random(sp.n_total)
    ptr, idx = orc.pattern("csr")
    nz, _ = orc.assemble(x, "csr")
    import scipy.sparse as sps
    A = sps.csr_matrix((nz, idx, ptr), shape=(sp.n_total,) * 2)
    v = rng.standard_normal(sp.n_total)
    eps = 1e-6
    _, bp = orc.assemble(x + eps * v, "csr", jac=False)
    _, bm = orc.assemble(x - eps * v, "csr", jac=False)
    fd = (bp - bm) / (2 * eps)
    assert np.max(np.abs(A @ v - fd)) <= 1e-6 * max(1.0, np.max(np.abs(fd)))


def test_csc_is_transpose_of_csr():
    sp = cases.cart(6, bc=cases.TURAN)
    orc = Oracle(sp, Pr=100.0, Ra=1e5, n=0.6)
    x = fe.splitmix_iterate(sp.n_total)
    import scipy.sparse as sps
    p1, i1 = orc.pattern("csc")
    p2, i2 = orc.pattern("csr")
    assert np.array_equal(p1, p2) and np.array_equal(i1, i2)  # structurally symmetric
    a, _ = orc.assemble(x, "csc")
    b, _ = orc.assemble(x, "csr")
    A = sps.csc_matrix((a, i1, p1), shape=(sp.n_total,) * 2)
    B = sps.csr_matrix((b, i2, p2), shape=(sp.n_total,) * 2)
    assert abs(A - B).max() == 0.0


@pytest.mark.parametrize("N", [2, 3, 10, 32, 50])
def test_size_regressions(N):
    """SURVEY Appendix A.5/B closed forms."""
    sp = cases.cart(N, bc=cases.SIMPLE)
    assert sp.n_free[0] == 2 * (2 * N - 1) ** 2 and sp.n_free[1] == 3 * N * N - 1
    assert sp.n_free[2] == (2 * N + 1) ** 2 - (6 * N - 1)
    ptr, idx = Oracle(sp).pattern()
    assert idx.size == 684 * N * N - 1336 * N + 702
    sp = cases.cart(N, bc=cases.TURAN)
    assert sp.n_free[2] == 4 * N * N - 1
    ptr, idx = Oracle(sp).pattern()
    assert idx.size == 684 * N * N - 1232 * N + 527


def test_cartesian_numbering_conventions():
    """tensor-order vertices, first-encounter edges, entities 1..9."""
    m = fe.CartesianModel((0, 1, 0, 1), (3, 2))
    assert m.cell_vertices[0].tolist() == [0, 1, 4, 5]
    assert m.cell_edges[0].tolist() == [0, 1, 2, 3]      # bottom, top, left, right of the first cell
    assert m.cell_edges[1].tolist() == [4, 5, 3, 6]      # left edge shared with cell 0's right
    assert m.cell_edges[3].tolist() == [1, 10, 11, 12]   # second row: bottom = top of the cell below
    assert m.vertex_entity.reshape(3, 4).tolist() == [[1, 5, 5, 2], [7, 9, 9, 8], [3, 6, 6, 4]]


def test_colored_oracle_equals_serial():
    sp = cases.cart(9, 8, bc=cases.TURAN)
    orc = Oracle(sp, Pr=100.0, Ra=1e5, n=0.6)
    x = fe.splitmix_iterate(sp.n_total)
    nx, ny = sp.model.partition
    cx, cy = np.tile(np.arange(nx), ny), np.repeat(np.arange(ny), nx)
    col = (cx & 1) + 2 * (cy & 1)
    order = np.argsort(col, kind="stable").astype(np.int64)
    cptr = np.concatenate([[0], np.cumsum(np.bincount(col, minlength=4))]).astype(np.int64)
    a, b = orc.assemble(x)
    a2, b2 = orc.assemble(x, colors=(cptr, order))
    ptr, idx = orc.pattern()
    assert cases.block_scaled_err(a2, a, cases.nz_blocks(sp, ptr, idx)) < 1e-13
    assert cases.block_scaled_err(b2, b, cases.res_blocks(sp)) < 1e-13


@pytest.mark.parametrize("layout", ["csc", "csr"])
def test_row_oracle_equals_full_assembly(layout):
    """orc_rows (the sampled row oracle of bench.py --verify / tests at the BASELINE sizes) rebuilds exactly the
    lists of orc_pattern / orc_assemble: same minor ids, bit-equal values (same cell order of the sums)."""
    import rowcheck
    for sp in (cases.cart(13, 7, bc=cases.TURAN), cases.spaces(fe.annulus_model(nr=3, nt=12, jitter=0.2, seed=3), cases.KUEHN)):
        prm = dict(Pr=100.0, Ra=1e5, n=0.6)
        orc = Oracle(sp, **prm)
        x = fe.splitmix_iterate(sp.n_total, 1234)
        ptr, idx = orc.pattern(layout)
        nz, b = orc.assemble(x, layout)
        for rows in (np.arange(sp.n_total), rowcheck.sample_rows(sp, nrandom=50)):
            rp, ri, rv, rb = orc.rows(rows, x, layout)
            assert np.array_equal(np.diff(rp), ptr[rows + 1] - ptr[rows])
            sel = np.concatenate([np.arange(ptr[r], ptr[r + 1]) for r in rows])
            assert np.array_equal(ri, idx[sel]) and np.array_equal(rv, nz[sel]) and np.array_equal(rb, b[rows])


def test_row_oracle_on_row_block_ranks():
    """bench.py --verify on N GPUs: every rank checks a sample of its OWNED dofs against the row oracle built from its
    local cell tables (owned rows + ghost rows).  Those lists must be the global matrix's: every incident cell of an
    owned dof is in the rank's table (owner = lowest row block touching the dof)."""
    import rowcheck
    model = fe.CartesianModel((0, 1, 0, 1), (24, 24))
    full = cases.spaces(model, cases.TURAN)
    prm = dict(Pr=100.0, Ra=1e5, n=1.4)
    x = fe.splitmix_iterate(full.n_total, 1234)
    forc = Oracle(full, **prm)
    fp, fi = forc.pattern("csc")
    fnz, fb = forc.assemble(x, "csc")
    seen = np.zeros(full.n_total, dtype=np.int32)
    for r in range(3):
        sp = fe.build_rank_spaces(model, cases.TURAN["V"], cases.TURAN["Ttags"], cases.TURAN["Texpr"], r, 3)
        seen[fe.owned_dofs(sp)] += 1
        rows = rowcheck.sample_rows(sp, nrandom=100)
        assert np.all(np.isin(rows, fe.owned_dofs(sp)))
        p, i, v, b = Oracle(sp, **prm).rows(rows, x, "csc")
        sel = np.concatenate([np.arange(fp[q], fp[q + 1]) for q in rows])
        assert np.array_equal(i, fi[sel]) and np.array_equal(v, fnz[sel]) and np.array_equal(b, fb[rows])
    assert np.all(seen == 1)
