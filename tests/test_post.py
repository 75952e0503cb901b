"""Post-processing assemblies (/root/reference/src/Solver.jl:197-292): stream function, heat function, entropies.

CPU: the numpy oracle (oracle/post_oracle.py) against the reference's golden Sth_max / Sfl_max
(examples/study1/conv_study_square.csv:2) and a manufactured stream function.  GPU: the library against the oracle."""
import math
import os

import numpy as np
import pytest
import scipy.sparse.linalg as spla

import cases
from cases import fe
from oracle import post_oracle as po

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
STH_MAX, SFL_MAX = 104.37198885839547, 2406.6836086605012   # conv_study_square.csv:2 (62x62)


def _grid_max(sp, nodal, n_plot=100):
    """max over the n_plot x n_plot LinRange(0,1) grid of the P2 interpolant (examples/study1/study1.jl:251-271)."""
    m = sp.model
    nx, ny = m.partition
    hx, hy = m.h
    xs = np.linspace(0.0, 1.0, n_plot)
    best = -np.inf
    for Y in xs:
        cy = min(int(math.floor(Y / hy)), ny - 1)
        eta = (Y - cy * hy) / hy
        for X in xs:
            cx = min(int(math.floor(X / hx)), nx - 1)
            phi, _ = fe.shape_at(fe.QUAD, (X - cx * hx) / hx, eta)
            best = max(best, float(phi @ nodal[sp.T.cell_nodes[cx + cy * nx]]))
    return best


def test_entropy_oracle_reproduces_the_reference_csv():
    sp = cases.cart(62, bc=cases.WAVE)
    x = np.load(os.path.join(GOLD, "square62_wave_n06_x.npy"))
    sth, sfl, a, b, be = po.entropy(sp, fe.post_tables(fe.QUAD), x)
    assert abs(_grid_max(sp, sth) - STH_MAX) <= 1e-12 * STH_MAX
    assert abs(_grid_max(sp, sfl) - SFL_MAX) <= 1e-11 * SFL_MAX
    assert 0.0 < be < 1.0 and abs(be - a / (a + b)) < 1e-15


@pytest.mark.parametrize("kind", ["quad", "tri"])
def test_stream_function_of_a_manufactured_velocity(kind):
    """u = (∂yΨ, -∂xΨ) with Ψ = sin(πx) sin(πy) (zero on the walls): -ΔΨ = ∂x u_y - ∂y u_x ... the sign the reference's
    weak form implies is  ∫∇v·∇ψ = ∫ v (∂x u_y - ∂y u_x) = ∫ v (-ΔΨ)  ->  ψ = Ψ... with (∂yΨ, -∂xΨ): ∂x u_y - ∂y u_x = -ΔΨ."""
    N = 24
    model = fe.CartesianModel((0, 1, 0, 1), (N, N)) if kind == "quad" else fe.square_tri_model(N, jitter=0.1, seed=1)
    sp = cases.spaces(model, cases.TURAN)
    x = np.zeros(sp.n_total)
    s = sp.U.node_signed
    free = s > 0
    X, Y = sp.U.node_coords[free, 0], sp.U.node_coords[free, 1]
    x[2 * (s[free] - 1)] = math.pi * np.sin(math.pi * X) * np.cos(math.pi * Y)
    x[2 * (s[free] - 1) + 1] = -math.pi * np.cos(math.pi * X) * np.sin(math.pi * Y)
    psi_sp = fe.build_p1(model, cases.TURAN["V"])
    A, b = po.assemble_p1(sp, psi_sp, fe.post_tables(model.cell_type), x, 0)
    psi = spla.spsolve(A, b)
    vfree = np.zeros(model.nverts, dtype=bool)
    vfree[model.cell_vertices[psi_sp.cell_dofs > 0]] = True
    ids = np.zeros(model.nverts, dtype=np.int64)
    ids[model.cell_vertices.ravel()] = psi_sp.cell_dofs.ravel()
    vc = model.vertex_coords[vfree]
    exact = np.sin(math.pi * vc[:, 0]) * np.sin(math.pi * vc[:, 1])
    assert np.max(np.abs(psi[ids[vfree] - 1] - exact)) < 6e-2  # (O(h): the no-slip wall nodes carry u = 0, the manufactured field slips)


GPU_CASES = [("cart20_wave", lambda: cases.cart(20, 14, bc=cases.WAVE, domain=(0, 1.5, 0, 1)), cases.WAVE["V"], [3, 4, 6]),
             ("annulus", lambda: cases.spaces(fe.annulus_model(nr=5, nt=20, jitter=0.3, seed=2), cases.KUEHN), ["all"], ["heatfunction_zero_points"]),
             ("sqtri_nodiri", lambda: cases.spaces(fe.square_tri_model(9, jitter=0.2, seed=5), cases.TURAN), cases.TURAN["V"], [])]


@pytest.mark.gpu
@pytest.mark.parametrize("case", GPU_CASES, ids=[c[0] for c in GPU_CASES])
def test_library_post_processing_matches_oracle(gm, gpu_lib, case):
    _, mk, vtags, pitags = case
    sp = mk()
    model = sp.model
    tabs = fe.post_tables(model.cell_type)
    x = fe.splitmix_iterate(sp.n_total, 99) - 0.5
    spaces = [fe.build_p1(model, vtags), fe.build_p1(model, pitags)]
    h = gm.capi.Handle(sp)
    try:
        h.post_create(sp, spaces[0], spaces[1], tabs)
        for which in (0, 1):
            Ao, bo = po.assemble_p1(sp, spaces[which], tabs, x, which)
            A, b = h.post_assemble(which, x)
            assert np.array_equal(A.indptr, Ao.indptr) and np.array_equal(A.indices, Ao.indices)
            assert np.max(np.abs(A.data - Ao.data)) <= 1e-12 * np.max(np.abs(Ao.data))
            assert np.max(np.abs(b - bo)) <= 1e-12 * max(np.max(np.abs(bo)), 1e-300)
        sth, sfl, a, bb, be = h.post_entropy(x)
        osth, osfl, oa, ob, obe = po.entropy(sp, tabs, x)
        assert np.max(np.abs(sth - osth)) <= 1e-12 * np.max(np.abs(osth)) and np.max(np.abs(sfl - osfl)) <= 1e-12 * np.max(np.abs(osfl))
        assert abs(a - oa) <= 1e-12 * oa and abs(bb - ob) <= 1e-12 * ob and abs(be - obe) <= 1e-12
    finally:
        h.close()


@pytest.mark.gpu
def test_library_entropy_reproduces_the_reference_csv(gm, gpu_lib):
    """Sth_max / Sfl_max of examples/study1/conv_study_square.csv:2 through the GPU path (the iterate is the committed Newton solution)."""
    sp = cases.cart(62, bc=cases.WAVE)
    x = np.load(os.path.join(GOLD, "square62_wave_n06_x.npy"))
    h = gm.capi.Handle(sp)
    try:
        h.post_create(sp, fe.build_p1(sp.model, cases.WAVE["V"]), fe.build_p1(sp.model, [3, 4, 6]), fe.post_tables(fe.QUAD))
        sth, sfl, *_ = h.post_entropy(x)
    finally:
        h.close()
    assert abs(_grid_max(sp, sth) - STH_MAX) <= 1e-12 * STH_MAX and abs(_grid_max(sp, sfl) - SFL_MAX) <= 1e-11 * SFL_MAX
