"""CPU tests of the restated nonlinear drivers (solver.py) on the oracle operator."""
import numpy as np

import cases
from orc_operator import OracleOperator


def test_trust_region_reaches_the_newton_solution(gm):
    """NLsolve method=:trust_region (validation-basak/basak.jl:22) and method=:newton + BackTracking (examples/simple.jl:23-29)
    solve the same discrete problem: fields agree to the solver tolerance."""
    sp = cases.cart(12, bc=cases.SIMPLE)
    prm = dict(Pr=0.7, Ra=1e3, n=1.0)
    a = gm.solver.newton(OracleOperator(sp, **prm), np.zeros(sp.n_total), ftol=1e-10, xtol=1e-50, iterations=60)
    b = gm.solver.trust_region(OracleOperator(sp, **prm), np.zeros(sp.n_total), ftol=1e-10, xtol=1e-50, iterations=300)
    assert a.f_converged and b.f_converged
    assert b.g_calls <= b.f_calls and b.iterations + 1 == b.f_calls
    assert np.max(np.abs(a.x - b.x)) <= 1e-7


def test_dogleg_step_stays_in_the_region(gm):
    import scipy.sparse as sps
    import scipy.sparse.linalg as spla
    rng = np.random.default_rng(0)
    A = sps.csc_matrix(rng.standard_normal((6, 6)) + 4 * np.eye(6))
    r = rng.standard_normal(6)
    d = np.sqrt(np.asarray(A.multiply(A).sum(axis=0)).ravel())
    lu = spla.splu(A)
    gn = -lu.solve(r)
    for delta in (1e-3, 0.3, 1.0, 1e3):
        p = gm.solver._dogleg(lu.solve, A, r, d, delta)
        assert gm.solver._wnorm(d, p) <= delta * (1 + 1e-12)
        if gm.solver._wnorm(d, gn) <= delta:
            assert np.allclose(p, gn)
        # the model decreases along the step
        assert np.linalg.norm(A @ p + r) <= np.linalg.norm(r) + 1e-12
