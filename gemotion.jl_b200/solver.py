"""Newton / trust-region drivers -- the host loop that CALLS the hot path.

Mirrors the caller side of the reference: `NLSolver(LUSolver(); nlsolver_opts...)`,
`FESolver`, `solve` (/root/reference/src/Solver.jl:162-195), i.e. NLsolve 4.5.1 `newton_`
with LineSearches 7.3.0 `BackTracking(order=3, c_1=1e-4, ρ_hi=0.5, ρ_lo=0.1)` (the options
every example passes, e.g. examples/simple.jl:23-29) and a sparse LU per iteration.  In
production this loop stays in Julia (NLsolve + UMFPACK); it is restated here only because
Julia is not in the image, so that end-to-end Newton parity (iteration counts, fields) can
be tested through the same operator interface the Julia plug-in implements
(INTEGRATION.md).  The operator `op` must provide

    op.n                                   number of free dofs
    op.residual_(b, x)                     residual!(b, op, x)
    op.jacobian_(A, x)                     jacobian!(A, op, x)            A: scipy CSC/CSR
    op.residual_and_jacobian_(b, A, x)     residual_and_jacobian!(b, A, op, x)
    op.allocate_jacobian()                 allocate_jacobian(op, x)       (symbolic pattern)
"""
from __future__ import annotations

import dataclasses
import math
import time
from typing import Optional

import numpy as np
import scipy.sparse.linalg as spla


@dataclasses.dataclass
class NLResult:
    x: np.ndarray
    iterations: int
    f_converged: bool
    x_converged: bool
    residual_norm: float
    f_calls: int
    g_calls: int
    trace: list
    t_assembly: float = 0.0
    t_linsolve: float = 0.0


def _nanmin(a, b):
    return b if math.isnan(a) else (a if math.isnan(b) else min(a, b))


def _nanmax(a, b):
    return b if math.isnan(a) else (a if math.isnan(b) else max(a, b))


def backtracking(phi, alpha0, phi0, dphi0, c_1=1e-4, rho_hi=0.5, rho_lo=0.1, iterations=1000, order=3):
    """LineSearches.BackTracking (v7.3.0) restated: cubic/quadratic interpolation backtracking."""
    eps = np.finfo(np.float64).eps
    iterfinitemax = -math.log2(eps)
    phix_0, phix_1 = phi0, phi0
    a1, a2 = alpha0, alpha0
    phix_1 = phi(a1)
    iterfinite = 0
    while not math.isfinite(phix_1) and iterfinite < iterfinitemax:
        iterfinite += 1
        a1 = a2
        a2 = a1 / 2
        phix_1 = phi(a2)
    it = 0
    while phix_1 > phi0 + c_1 * a2 * dphi0:
        it += 1
        if it > iterations:
            raise RuntimeError("Linesearch failed to converge, reached maximum iterations %d" % iterations)
        if order == 2 or it == 1:
            a_tmp = -(dphi0 * a2 ** 2) / (2 * (phix_1 - phi0 - dphi0 * a2))
        else:
            div = 1.0 / (a1 ** 2 * a2 ** 2 * (a2 - a1))
            a = (a1 ** 2 * (phix_1 - phi0 - dphi0 * a2) - a2 ** 2 * (phix_0 - phi0 - dphi0 * a1)) * div
            b = (-a1 ** 3 * (phix_1 - phi0 - dphi0 * a2) + a2 ** 3 * (phix_0 - phi0 - dphi0 * a1)) * div
            if abs(a) <= eps:
                a_tmp = dphi0 / (2 * b)
            else:
                d = max(b * b - 3 * a * dphi0, 0.0)
                a_tmp = (-b + math.sqrt(d)) / (3 * a)
        a1 = a2
        a_tmp = _nanmin(a_tmp, a2 * rho_hi)  # avoid too small reductions
        a2 = _nanmax(a_tmp, a2 * rho_lo)     # avoid too big reductions
        phix_0, phix_1 = phix_1, phi(a2)
    return a2, phix_1


def newton(op, x0, ftol=1e-8, xtol=0.0, iterations=1000, linesearch="backtracking",
           show_trace=False) -> NLResult:
    """NLsolve.nlsolve(df, x0; method=:newton, linesearch=BackTracking(), ftol, xtol, iterations)."""
    n = op.n
    x = np.array(x0, dtype=np.float64, copy=True)
    f = np.zeros(n)
    A = op.allocate_jacobian()
    t_asm = t_lin = 0.0
    t0 = time.perf_counter()
    op.residual_and_jacobian_(f, A, x)
    t_asm += time.perf_counter() - t0
    f_calls, g_calls = 1, 1
    if not np.all(np.isfinite(f)):
        raise FloatingPointError("non-finite residual at the initial guess")
    xold = np.full(n, np.nan)
    trace = []
    it = 0
    fnorm = float(np.max(np.abs(f))) if n else 0.0
    x_conv, f_conv = False, fnorm <= ftol
    trace.append((0, fnorm, float("nan")))
    if show_trace:
        print("Iter     f(x) inf-norm    Step 2-norm")
        print("%6d   %14e   %14e" % trace[-1])
    stopped = False
    while not stopped and not (x_conv or f_conv) and it < iterations:
        it += 1
        if it > 1:
            t0 = time.perf_counter()
            op.jacobian_(A, x)
            t_asm += time.perf_counter() - t0
            g_calls += 1
        t0 = time.perf_counter()
        lu = spla.splu(A.tocsc() if A.format != "csc" else A)
        p = -lu.solve(f)
        t_lin += time.perf_counter() - t0
        xold[:] = x
        if linesearch == "static":
            x = xold + p
            t0 = time.perf_counter()
            op.residual_(f, x)
            t_asm += time.perf_counter() - t0
            f_calls += 1
        else:
            g = A.T @ f
            phi0 = 0.5 * float(f @ f)
            dphi0 = float(g @ p)
            xbase = xold.copy()

            def phi(alpha):
                nonlocal f_calls, t_asm
                t1 = time.perf_counter()
                op.residual_(f, xbase + alpha * p)
                t_asm += time.perf_counter() - t1
                f_calls += 1
                return 0.5 * float(f @ f)

            alpha, _ = backtracking(phi, 1.0, phi0, dphi0)
            x = xbase + alpha * p
        fnorm = float(np.max(np.abs(f)))
        x_conv = float(np.max(np.abs(x - xold))) <= xtol
        f_conv = fnorm <= ftol
        stopped = bool(np.any(np.isnan(x)) or np.any(np.isnan(f)))
        trace.append((it, fnorm, float(np.linalg.norm(x - xold))))
        if show_trace:
            print("%6d   %14e   %14e" % trace[-1])
    return NLResult(x, it, f_conv, x_conv, fnorm, f_calls, g_calls, trace, t_asm, t_lin)


def _wnorm(d, v):
    return float(np.sqrt(np.sum((d * v) ** 2)))


def _dogleg(lu_solve, A, r, d, delta):
    """NLsolve `dogleg!` (Nocedal & Wright, eq. 4.16) with the column scaling d: Gauss-Newton step if it lies inside the
    region, else the Cauchy point along -D^-2 J'r, else the point of the dogleg path on the boundary."""
    p_i = -lu_solve(r)
    if _wnorm(d, p_i) <= delta:
        return p_i
    g = -(A.T @ r) / d ** 2
    Jg = A @ g
    p_c = g * (_wnorm(d, g) ** 2 / float(Jg @ Jg))
    if _wnorm(d, p_c) >= delta:
        return p_c * (delta / _wnorm(d, p_c))
    p_i = p_i - p_c
    b = 2.0 * float(np.sum(d * p_c * d * p_i))
    a = _wnorm(d, p_i) ** 2
    tau = (-b + math.sqrt(b * b - 4.0 * a * (_wnorm(d, p_c) ** 2 - delta ** 2))) / (2.0 * a)
    return p_c + tau * p_i


def trust_region(op, x0, ftol=1e-8, xtol=0.0, iterations=1000, factor=1.0, autoscale=True, show_trace=False) -> NLResult:
    """NLsolve.nlsolve(df, x0; method=:trust_region, factor=1.0, autoscale=true) -- the driver of
    /root/reference/examples/validation-basak/basak.jl:20-26 (`linesearch` is ignored by this method).  Restated from the
    algorithm NLsolve documents (Powell dogleg, N&W ch. 4 / 11; trust radius halved for rho < 0.1, doubled to
    2 |D p| for rho >= 0.9, scaling d_j = max(0.1 d_j, |J[:, j]|)); NLsolve's source is not in the image, so details such as
    tie handling are unverified against it."""
    n = op.n
    x = np.array(x0, dtype=np.float64, copy=True)
    r = np.zeros(n)
    f = np.zeros(n)
    A = op.allocate_jacobian()
    t_asm = t_lin = 0.0
    t0 = time.perf_counter()
    op.residual_and_jacobian_(r, A, x)
    t_asm += time.perf_counter() - t0
    f_calls, g_calls = 1, 1
    if not np.all(np.isfinite(r)):
        raise FloatingPointError("non-finite residual at the initial guess")

    def colnorms(M):
        return np.sqrt(np.asarray(M.multiply(M).sum(axis=0)).ravel())

    d = np.ones(n)
    if autoscale:
        d = colnorms(A)
        d[d == 0.0] = 1.0
    delta = factor * _wnorm(d, x)
    if delta == 0.0:
        delta = factor
    eta = 1e-4
    it = 0
    fnorm = float(np.max(np.abs(r))) if n else 0.0
    f_conv, x_conv = fnorm <= ftol, False
    trace = [(0, fnorm, float("nan"))]
    stopped = False
    lu = None
    while not stopped and not (f_conv or x_conv) and it < iterations:
        it += 1
        if lu is None:
            t0 = time.perf_counter()
            lu = spla.splu(A.tocsc() if A.format != "csc" else A)
            t_lin += time.perf_counter() - t0
        t0 = time.perf_counter()
        p = _dogleg(lu.solve, A, r, d, delta)
        t_lin += time.perf_counter() - t0
        xold = x.copy()
        x = x + p
        t0 = time.perf_counter()
        op.residual_(f, x)
        t_asm += time.perf_counter() - t0
        f_calls += 1
        r_predict = A @ p + r
        rho = (float(r @ r) - float(f @ f)) / (float(r @ r) - float(r_predict @ r_predict))
        if rho > eta:  # successful iteration
            r[:] = f
            t0 = time.perf_counter()
            op.jacobian_(A, x)
            t_asm += time.perf_counter() - t0
            g_calls += 1
            lu = None
            if autoscale:
                d = np.maximum(0.1 * d, colnorms(A))
            x_conv = float(np.max(np.abs(x - xold))) <= xtol
            f_conv = float(np.max(np.abs(r))) <= ftol
        else:
            x = xold
            x_conv = f_conv = False
        if rho < 0.1:
            delta = delta / 2.0
        elif rho >= 0.9:
            delta = 2.0 * _wnorm(d, p)
        elif rho >= 0.5:
            delta = max(delta, 2.0 * _wnorm(d, p))
        stopped = bool(np.any(np.isnan(x)) or np.any(np.isnan(f)))
        fnorm = float(np.max(np.abs(r)))
        trace.append((it, fnorm, float(np.linalg.norm(p))))
        if show_trace:
            print("%6d   %14e   %14e   rho %10.3e  delta %10.3e" % (it, fnorm, trace[-1][2], rho, delta))
    return NLResult(x, it, f_conv, x_conv, fnorm, f_calls, g_calls, trace, t_asm, t_lin)


# --------------------------------------------------------------------------------------
# field evaluation helpers used by the parity tests (Nu_bot_avg of examples/study1/study1.jl:223-249)
# --------------------------------------------------------------------------------------
def temperature_cell_values(sp, x):
    """Per-cell nodal T values (free + Dirichlet) [nc, nn]."""
    oT = sp.offsets[2]
    d = sp.T.cell_dofs
    vals = np.where(d > 0, x[oT + np.maximum(d, 1) - 1], sp.T.diri_values[np.maximum(-d, 1) - 1])
    return vals


def nu_bot_avg(sp, x, n_plot=100):
    """mean over the first n_plot-1 of n_plot equispaced x in [0,1] of -dT/dy(x, 0)
    (examples/study1/study1.jl:223-249) on a CartesianModel."""
    from . import fe
    m = sp.model
    nx, ny = m.partition
    hx, hy = m.h
    Tc = temperature_cell_values(sp, x)
    xs = np.linspace(0.0, 1.0, n_plot)
    out = []
    for X in xs[:-1]:
        cx = min(int(math.floor((X - m.origin[0]) / hx)), nx - 1)
        xi = (X - (m.origin[0] + cx * hx)) / hx
        _, dphi = fe.shape_at(fe.QUAD, xi, 0.0)
        out.append(-(dphi[:, 1] @ Tc[cx]) / hy)
    return float(np.sum(out) / len(out))


def nu_bot_avg_generic(sp, x, n_plot=100):
    """The same functional on any mesh of the unit square (triangles: the cell with an edge on y = 0 that contains the sample
    point, physical gradient through the affine map) -- used to compare the triangle discretisation with the reference's
    quadrilateral convergence study (examples/study1/conv_study_square.csv)."""
    from . import fe
    m = sp.model
    d = sp.T.cell_dofs
    off = sp.offsets[2]
    Tc = np.where(d > 0, x[off + np.maximum(d, 1) - 1], sp.T.diri_values[np.maximum(-d, 1) - 1])
    vc, cv = m.vertex_coords, m.cell_vertices
    cand = np.nonzero(np.isclose(vc[cv][:, :, 1], 0.0).sum(axis=1) >= 2)[0]
    out = []
    for X in np.linspace(0.0, 1.0, n_plot)[:-1]:
        for c in cand:
            P = vc[cv[c]]
            if m.cell_type == fe.TRI:
                A = np.array([P[1] - P[0], P[2] - P[0]]).T          # x = P0 + A (xi, eta)
                r = np.linalg.solve(A, np.array([X, 0.0]) - P[0])
                if r[0] < -1e-12 or r[1] < -1e-12 or r[0] + r[1] > 1 + 1e-12:
                    continue
                _, dphi = fe.shape_at(fe.TRI, r[0], r[1])
                out.append(-np.linalg.solve(A.T, dphi.T @ Tc[c])[1])
            else:
                x0, x1 = P[:, 0].min(), P[:, 0].max()
                if X < x0 - 1e-12 or X > x1 + 1e-12:
                    continue
                _, dphi = fe.shape_at(fe.QUAD, (X - x0) / (x1 - x0), 0.0)
                out.append(-(dphi[:, 1] @ Tc[c]) / (P[:, 1].max() - P[:, 1].min()))
            break
    assert len(out) == n_plot - 1
    return float(np.mean(out))


# --------------------------------------------------------------------------------------
# entropy fields of /root/reference/src/Solver.jl:281-289, sampled like examples/study1/study1.jl:251-271
# (used only to pin the oracle's VELOCITY field against conv_study_square.csv: Sth_max, Sfl_max)
# --------------------------------------------------------------------------------------
def _cell_values(space, x, off):
    d = space.cell_dofs
    return np.where(d > 0, x[off + np.maximum(d, 1) - 1], space.diri_values[np.maximum(-d, 1) - 1])


def entropy_fields_max(sp, x, n_plot=100):
    """(Sth_max, Sfl_max): nodal interpolation of |grad T|^2 and 1e-4*(2(dx ux^2 + dy uy^2) + (dy ux + dx uy)^2)
    into a Q2 H1 space (cell by cell, the last cell writes shared nodes), evaluated on the n_plot x n_plot
    grid of LinRange(0,1) points and maximised."""
    from . import fe
    m = sp.model
    nx, ny = m.partition
    hx, hy = m.h
    nn = sp.ref.nn
    Uc = _cell_values(sp.U, x, sp.offsets[0])          # [nc, 18]
    Tc = _cell_values(sp.T, x, sp.offsets[2])          # [nc, 9]
    nodes = sp.ref.nodes
    dph = np.array([fe.shape_at(fe.QUAD, a, b)[1] for a, b in nodes])   # [node p][shape i][2] reference gradients
    dph = dph / np.array([hx, hy])
    gT = np.einsum("pik,ci->cpk", dph, Tc)                               # grad T at the 9 nodes of each cell
    sth = (gT ** 2).sum(axis=2)
    G = np.stack([np.einsum("pik,ci->cpk", dph, Uc[:, :nn]), np.einsum("pik,ci->cpk", dph, Uc[:, nn:])], axis=3)  # [c,p,k,a] = d_k u_a
    sfl = 1e-4 * (2.0 * (G[:, :, 0, 0] ** 2 + G[:, :, 1, 1] ** 2) + (G[:, :, 0, 1] + G[:, :, 1, 0]) ** 2)
    nnode = sp.T.node_coords.shape[0]
    out = []
    for cellvals in (sth, sfl):
        nodal = np.zeros(nnode)
        nodal[sp.T.cell_nodes.reshape(-1)] = cellvals.reshape(-1)       # later cells overwrite earlier ones
        best = -np.inf
        xs = np.linspace(0.0, 1.0, n_plot)
        for Y in xs:
            cy = min(int(math.floor((Y - m.origin[1]) / hy)), ny - 1)
            eta = (Y - (m.origin[1] + cy * hy)) / hy
            for X in xs:
                cx = min(int(math.floor((X - m.origin[0]) / hx)), nx - 1)
                xi = (X - (m.origin[0] + cx * hx)) / hx
                phi, _ = fe.shape_at(fe.QUAD, xi, eta)
                v = float(phi @ nodal[sp.T.cell_nodes[cx + cy * nx]])
                best = max(best, v)
        out.append(best)
    return tuple(out)
