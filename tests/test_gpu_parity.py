"""GPU parity: CUDA path (through the C ABI) vs the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): pattern bit-exact; nzval and residual within 1e-12 in the
block-scaled sense of SURVEY 8a-note.
"""
import os

import numpy as np
import pytest

import cases
from cases import fe
from oracle.oracle import Oracle

pytestmark = pytest.mark.gpu

TOL = 1e-12  # north_star: "CSR values and residual within 1e-12 relative in fp64"

PARAMS = [
    dict(Pr=0.7, Ra=1e3, n=1.0),                       # examples/simple.jl
    dict(Pr=100.0, Ra=1e5, n=0.6),                     # config 2 (turan-style)
    dict(Pr=100.0, Ra=1e5, n=1.4, jac_scaling=1e-6),   # config 5 exponent; jac_scaling of validate-matin
]


def _models():
    yield "cart8_simple", cases.cart(8, bc=cases.SIMPLE)
    yield "cart13x7_turan", cases.cart(13, 7, bc=cases.TURAN, domain=(0, 2, 0, 1))
    yield "cart1x1", cases.cart(1, 1, bc=cases.TURAN)
    yield "cart2x3_wave", cases.cart(2, 3, bc=cases.WAVE)
    yield "annulus", cases.spaces(fe.annulus_model(nr=4, nt=16, jitter=0.3, seed=1), cases.KUEHN)
    yield "sqtri", cases.spaces(fe.square_tri_model(6, jitter=0.3, seed=2), cases.TURAN)
    yield "fan14", cases.spaces(cases.fan_model(14), cases.FAN)
    yield "sqtri1", cases.spaces(fe.square_tri_model(1), cases.TURAN)   # two triangles
    yield "jquad7x5", cases.spaces(fe.JitteredQuadModel((0, 2, 0, 1), (7, 5), jitter=0.25, seed=2), cases.WAVE)   # unstructured quads: bilinear geometry


MODELS = dict(_models())


def _check(gm, sp, prm, layout, path, x, floor=0.0, kernel=0, xchunk=0):
    orc = Oracle(sp, **prm)
    optr, oidx = orc.pattern(layout)
    onz, ob = orc.assemble(x, layout)
    h = gm.capi.Handle(sp, layout=layout, path=path, kernel=kernel, xchunk=xchunk)
    try:
        h.set_params(**prm)
        nnz = h.pattern()
        ptr, idx = h.get_pattern()
        assert nnz == oidx.size
        assert np.array_equal(ptr, optr) and np.array_equal(idx, oidx), "pattern must be bit-exact"
        nz = np.full(nnz, np.nan)
        b = np.full(sp.n_total, np.nan)
        h.assemble(x, nz, b)
        e_nz = cases.block_scaled_err(nz, onz, cases.nz_blocks(sp, optr, oidx))
        e_b = cases.block_scaled_err(b, ob, cases.res_blocks(sp), floor)
        # x-momentum x T entries are stored exact zeros (g_x = 0) on both sides
        f = cases.field_of_dof(sp)
        major = np.repeat(np.arange(sp.n_total), np.diff(optr))
        tu = (f[major] == 2) & (f[oidx] == 0) if layout == "csc" else (f[major] == 0) & (f[oidx] == 2)
        xrow = tu & ((oidx if layout == "csc" else major) % 2 == 0)
        assert np.all(onz[xrow] == 0.0) and np.all(nz[xrow] == 0.0)
        assert e_nz <= TOL and e_b <= TOL, (e_nz, e_b)
        # residual-only and jacobian-only calls give the same numbers
        b2 = np.full(sp.n_total, np.nan)
        h.assemble(x, None, b2)
        nz2 = np.full(nnz, np.nan)
        h.assemble(x, nz2, None)
        assert np.array_equal(b2, b) and np.array_equal(nz2, nz)
        info = h.info()
    finally:
        h.close()
    return info


@pytest.mark.parametrize("name", list(MODELS))
@pytest.mark.parametrize("pi", range(len(PARAMS)))
@pytest.mark.parametrize("layout", ["csc", "csr"])
def test_scatter_path_matches_oracle(gm, gpu_lib, name, pi, layout):
    sp = MODELS[name]
    x = fe.splitmix_iterate(sp.n_total, 1234)
    info = _check(gm, sp, PARAMS[pi], layout, gm.capi.GM_PATH_SCATTER, x)
    assert info["active_path"] == gm.capi.GM_PATH_SCATTER


@pytest.mark.parametrize("name", list(MODELS))
@pytest.mark.parametrize("pi", range(len(PARAMS)))
@pytest.mark.parametrize("layout", ["csc", "csr"])
def test_gather_path_matches_oracle(gm, gpu_lib, name, pi, layout):
    """Generic owner-computes path (gm_gather.cuh: cell matrices by node pairs + list gather): triangles and quads."""
    sp = MODELS[name]
    x = fe.splitmix_iterate(sp.n_total, 1234)
    info = _check(gm, sp, PARAMS[pi], layout, gm.capi.GM_PATH_GATHER, x)
    assert info["active_path"] == gm.capi.GM_PATH_GATHER


def test_gather_path_on_a_gmsh_file(gm, gpu_lib, tmp_path):
    """A triangle mesh read back from a .msh file (format 4.1, physical names -> tags) through the default path."""
    from _pkg import gm as pkg
    model = fe.annulus_model(nr=6, nt=40, jitter=0.3, seed=11)
    path = str(tmp_path / "annulus.msh")
    pkg.gmsh.write_msh(model, path, version="4.1")
    sp = cases.spaces(pkg.gmsh.read_msh(path), cases.KUEHN)
    x = fe.splitmix_iterate(sp.n_total, 99)
    info = _check(gm, sp, dict(Pr=0.706, Ra=4.7e4, n=0.8), "csc", gm.capi.GM_PATH_AUTO, x)
    assert info["active_path"] == gm.capi.GM_PATH_GATHER


def test_smooth_iterate_and_zero(gm, gpu_lib):
    sp = cases.cart(16, bc=cases.TURAN)
    for x in (fe.smooth_iterate(sp), np.zeros(sp.n_total)):
        _check(gm, sp, PARAMS[1], "csc", gm.capi.GM_PATH_SCATTER, x, floor=1e-6)


def test_config1_sizes(gm, gpu_lib):
    """examples/simple.jl: 50x50, nnz = 1,643,902."""
    sp = cases.cart(50, bc=cases.SIMPLE)
    assert sp.n_total == 37003
    info = _check(gm, sp, PARAMS[0], "csc", gm.capi.GM_PATH_AUTO, fe.splitmix_iterate(sp.n_total))
    assert info["n_owned"] == 37003


def test_error_paths(gm, gpu_lib):
    sp = cases.cart(4)
    h = gm.capi.Handle(sp)
    with pytest.raises(gm.capi.GmError):  # assemble before pattern
        h.assemble(np.zeros(sp.n_total), None, np.zeros(sp.n_total))
    h.pattern()
    with pytest.raises(gm.capi.GmError):  # nothing requested
        h.assemble(np.zeros(sp.n_total), None, None, flags=0)
    h.close()


@pytest.mark.parametrize("path", ["scatter", "cart", "gather"])
def test_check_finite_flag_scans_on_the_device(gm, gpu_lib, path):
    """GM_CHECK_FINITE: a NaN in the iterate must surface as GM_ERR_NONFINITE (-6) from both entry points, for nzval and for
    the residual; a clean iterate passes."""
    import torch
    capi = gm.capi
    sp = cases.cart(9, 6, bc=cases.TURAN)
    h = capi.Handle(sp, path={"scatter": capi.GM_PATH_SCATTER, "cart": capi.GM_PATH_CART, "gather": capi.GM_PATH_GATHER}[path])
    try:
        h.set_params(Pr=100.0, Ra=1e5, n=0.6)
        nnz = h.pattern()
        x = fe.splitmix_iterate(sp.n_total, 1234)
        nz, b = np.empty(nnz), np.empty(sp.n_total)
        h.assemble(x, nz, b, flags=capi.GM_JACOBIAN | capi.GM_RESIDUAL | capi.GM_CHECK_FINITE)
        bad = x.copy()
        bad[sp.n_total // 3] = np.nan
        for fl in (capi.GM_JACOBIAN, capi.GM_RESIDUAL):
            with pytest.raises(capi.GmError) as ei:
                h.assemble(bad, nz if fl == capi.GM_JACOBIAN else None, b if fl == capi.GM_RESIDUAL else None, flags=fl | capi.GM_CHECK_FINITE)
            assert ei.value.code == -6
        dx, dnz, db = torch.from_numpy(bad).cuda(), torch.empty(nnz, dtype=torch.float64, device="cuda"), torch.empty(sp.n_total, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        with pytest.raises(capi.GmError) as ei:
            h.assemble_device(dx.data_ptr(), dnz.data_ptr(), db.data_ptr(), capi.GM_JACOBIAN | capi.GM_RESIDUAL | capi.GM_CHECK_FINITE, sync=False)
        assert ei.value.code == -6
    finally:
        h.close()


def test_two_handles_one_process_interleaved(gm, gpu_lib):
    """'distinct handles are independent' (include/gemotion_b200.h): two handles with different parameters and paths in one
    process, assemblies issued back to back without synchronisation in between, each on its own stream."""
    import torch
    capi = gm.capi
    spa, spb = cases.cart(24, 16, bc=cases.TURAN), cases.spaces(fe.square_tri_model(10, jitter=0.2, seed=4), cases.TURAN)
    pa, pb = dict(Pr=100.0, Ra=1e5, n=0.6), dict(Pr=0.7, Ra=1e3, n=1.0)
    items = []
    for sp, prm, path in ((spa, pa, capi.GM_PATH_SCATTER), (spb, pb, capi.GM_PATH_SCATTER), (spa, pb, capi.GM_PATH_CART),
                          (spb, pa, capi.GM_PATH_GATHER), (spa, pb, capi.GM_PATH_GATHER)):
        h = capi.Handle(sp, path=path)
        h.set_params(**prm)
        nnz = h.pattern()
        x = fe.splitmix_iterate(sp.n_total, 1234)
        items.append((sp, prm, h, torch.from_numpy(x).cuda(), torch.empty(nnz, dtype=torch.float64, device="cuda"),
                      torch.empty(sp.n_total, dtype=torch.float64, device="cuda"), x))
    torch.cuda.synchronize()
    fl = capi.GM_JACOBIAN | capi.GM_RESIDUAL
    for _ in range(3):
        for sp, prm, h, dx, dnz, db, x in items:
            h.assemble_device(dx.data_ptr(), dnz.data_ptr(), db.data_ptr(), fl, sync=False)
    for sp, prm, h, dx, dnz, db, x in items:
        h.timing()  # waits for the handle's last event
        orc = Oracle(sp, **prm)
        optr, oidx = orc.pattern("csc")
        onz, ob = orc.assemble(x, "csc")
        assert cases.block_scaled_err(dnz.cpu().numpy(), onz, cases.nz_blocks(sp, optr, oidx)) <= TOL
        assert cases.block_scaled_err(db.cpu().numpy(), ob, cases.res_blocks(sp)) <= TOL
        h.close()


CART_CASES = [
    ("8x8_simple", 8, 8, cases.SIMPLE), ("13x7_turan", 13, 7, cases.TURAN), ("1x1", 1, 1, cases.TURAN),
    ("2x3_wave", 2, 3, cases.WAVE), ("20x19_turan", 20, 19, cases.TURAN), ("70x17_simple", 70, 17, cases.SIMPLE),
    ("130x9_wave", 130, 9, cases.WAVE), ("3x40_turan", 3, 40, cases.TURAN), ("65x24_turan", 65, 24, cases.TURAN),
    ("300x5_simple", 300, 5, cases.SIMPLE),
]


@pytest.mark.parametrize("case", CART_CASES, ids=[c[0] for c in CART_CASES])
@pytest.mark.parametrize("pi", range(len(PARAMS)))
@pytest.mark.parametrize("layout", ["csc", "csr"])
def test_structured_path_matches_oracle(gm, gpu_lib, case, pi, layout):
    """The owner-computes row-gather kernels (gm_cart.cuh): strips, halo rows, x-chunks, ragged sizes."""
    _, nx, ny, bc = case
    sp = cases.cart(nx, ny, bc=bc, domain=(0, 1.5, 0, 1))
    x = fe.splitmix_iterate(sp.n_total, 1234)
    info = _check(gm, sp, PARAMS[pi], layout, gm.capi.GM_PATH_CART, x)
    assert info["active_path"] == gm.capi.GM_PATH_CART


@pytest.mark.parametrize("kernel", ["gath", "gath1", "sweep"])
@pytest.mark.parametrize("case", [CART_CASES[1], CART_CASES[6], CART_CASES[9]], ids=lambda c: c[0])
@pytest.mark.parametrize("layout", ["csc", "csr"])
def test_structured_path_other_kernels(gm, gpu_lib, kernel, case, layout):
    """The older generations of the structured kernel (gm_desc.kernel: register gather k_gath with two / one trio,
    first-generation sweep k_cart) stay selectable for A/B measurements and stay green."""
    kid = {"gath": gm.capi.GM_KERNEL_GATH, "gath1": gm.capi.GM_KERNEL_GATH1, "sweep": gm.capi.GM_KERNEL_SWEEP}[kernel]
    _, nx, ny, bc = case
    sp = cases.cart(nx, ny, bc=bc, domain=(0, 1.5, 0, 1))
    x = fe.splitmix_iterate(sp.n_total, 1234)
    for pi in (0, 1):
        _check(gm, sp, PARAMS[pi], layout, gm.capi.GM_PATH_CART, x, kernel=kid)


@pytest.mark.parametrize("kernel", ["mma", "gath"])
@pytest.mark.parametrize("xchunk", [32, 128, 256])
@pytest.mark.parametrize("case", [CART_CASES[5], CART_CASES[6], CART_CASES[9]], ids=lambda c: c[0])
def test_structured_path_chunk_lengths(gm, gpu_lib, kernel, xchunk, case):
    """Every x-chunk length the library instantiates (gm_desc.xchunk; bench.py's grids pick 256, multi-GPU ranks 128)
    against the oracle, with chunk seams inside the mesh where the mesh is long enough."""
    kid = {"mma": gm.capi.GM_KERNEL_MMA, "gath": gm.capi.GM_KERNEL_GATH}[kernel]
    if kernel == "gath" and xchunk == 32:
        pytest.skip("k_gath is instantiated for 128 / 256 columns only")
    _, nx, ny, bc = case
    sp = cases.cart(nx, ny, bc=bc, domain=(0, 1.5, 0, 1))
    x = fe.splitmix_iterate(sp.n_total, 1234)
    _check(gm, sp, PARAMS[2], "csc", gm.capi.GM_PATH_CART, x, kernel=kid, xchunk=xchunk)


def test_config2_256_matches_oracle(gm, gpu_lib):
    """Config 2 (256x256 quads, turan BCs, n=0.6, Ra=1e5, Pr=100; nnz = 44,511,759): the structured
    kernel against the oracle's full serial assembly -- pattern bit-exact, every entry within 1e-12 block-scaled --
    and the generic scatter path against the same numbers."""
    sp = cases.cart(256, bc=cases.TURAN)
    x = fe.splitmix_iterate(sp.n_total, 1234)
    orc = Oracle(sp, **PARAMS[1])
    optr, oidx = orc.pattern("csc")
    onz, ob = orc.assemble(x, "csc")
    assert oidx.size == 44511759
    blocks, rblocks = cases.nz_blocks(sp, optr, oidx), cases.res_blocks(sp)
    for path in (gm.capi.GM_PATH_CART, gm.capi.GM_PATH_SCATTER):
        h = gm.capi.Handle(sp, layout="csc", path=path)
        h.set_params(**PARAMS[1])
        nnz = h.pattern()
        assert nnz == 44511759
        nz, b = np.empty(nnz), np.empty(sp.n_total)
        h.assemble(x, nz, b)
        ptr, idx = h.get_pattern()
        h.close()
        assert np.array_equal(ptr, optr) and np.array_equal(idx, oidx)
        assert cases.block_scaled_err(nz, onz, blocks) <= TOL
        assert cases.block_scaled_err(b, ob, rblocks) <= TOL


@pytest.mark.parametrize("N,pi,layout", [(2048, 1, "csc"), (1100, 2, "csr")])
def test_bench_sizes_row_oracle(gm, gpu_lib, N, pi, layout):
    """Config 4 (2048x2048, n=0.6: nnz = 2,866,381,327 > 2^31, 9 x 257 CTAs of the 256-column k-chunk build that
    bench.py times) and a ragged 1100x1100 CSR case: complete lists of ~10^4 sampled major dofs (chunk / strip
    boundaries, field tails with the largest 64-bit list bases, random) against the row oracle.  bench.py --verify runs
    the same check at 4096^2 and on every rank of the multi-GPU runs."""
    import torch
    import rowcheck
    sp = cases.cart(N, bc=cases.TURAN)
    prm = PARAMS[pi]
    x = fe.splitmix_iterate(sp.n_total, 1234)
    h = gm.capi.Handle(sp, layout=layout, path=gm.capi.GM_PATH_CART)
    try:
        h.set_params(**prm)
        nnz = h.pattern()
        assert nnz == 684 * N * N - 1232 * N + 527  # SURVEY A.5 closed form (turan BCs)
        dx = torch.from_numpy(x).cuda()
        dnz = torch.full((nnz,), float("nan"), dtype=torch.float64, device="cuda")
        db = torch.full((sp.n_total,), float("nan"), dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()  # (the handle runs on its own stream: the NaN fill must have finished)
        h.assemble_device(dx.data_ptr(), dnz.data_ptr(), db.data_ptr(), gm.capi.GM_JACOBIAN | gm.capi.GM_RESIDUAL, sync=True)
        assert not bool(torch.isnan(dnz).any()) and not bool(torch.isnan(db).any()), "every entry must be written"
        rows = rowcheck.sample_rows(sp, nrandom=3000)
        out = rowcheck.check_rows(sp, h, Oracle(sp, **prm), x, dnz.data_ptr(), db.data_ptr(), layout, rows)
        ptr_tail = h.get_rows(rows[-1:])[0]
        assert out["pattern_ok"], out
        assert out["max_err_nz"] <= TOL and out["max_err_b"] <= TOL, out
        assert out["rows"] >= 5000
        if N == 2048:
            assert nnz > 2 ** 31 and ptr_tail[-1] > 0
    finally:
        h.close()


@pytest.mark.parametrize("nranks,nx,ny", [(2, 9, 8), (3, 70, 19), (4, 6, 4)])
@pytest.mark.parametrize("layout", ["csc", "csr"])
def test_row_blocks_emulated_on_one_gpu(gm, gpu_lib, nranks, nx, ny, layout):
    """owner-computes row blocks.  All ranks live in this process on one GPU (NCCL cannot
    put two ranks on one device), so the packed interface buffers are handed over directly
    (GM_NO_EXCHANGE + gm_iface_pack / gm_iface_unpack_add); bench.py --gpus N runs the same
    kernels with ncclSend/ncclRecv in between.  Every owned row must equal the oracle's."""
    import torch
    capi = gm.capi
    prm = PARAMS[1]
    model = fe.CartesianModel((0, 1.5, 0, 1), (nx, ny))
    full = cases.spaces(model, cases.TURAN)
    x = fe.splitmix_iterate(full.n_total, 1234)
    orc = Oracle(full, **prm)
    optr, oidx = orc.pattern(layout)
    onz, ob = orc.assemble(x, layout)
    blocks = cases.nz_blocks(full, optr, oidx)
    rblocks = cases.res_blocks(full)
    dx = torch.from_numpy(x).cuda()
    flags = capi.GM_JACOBIAN | capi.GM_RESIDUAL
    ranks = []
    for r in range(nranks):
        sp = fe.build_rank_spaces(model, cases.TURAN["V"], cases.TURAN["Ttags"], cases.TURAN["Texpr"], r, nranks)
        h = capi.Handle(sp, layout=layout, path=capi.GM_PATH_CART, rank=r, nranks=nranks)
        h.set_params(**prm)
        nnz = h.pattern()
        dnz = torch.full((nnz,), float("nan"), dtype=torch.float64, device="cuda")
        db = torch.full((full.n_total,), float("nan"), dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()  # (the handle runs on its own stream)
        h.assemble_device(dx.data_ptr(), dnz.data_ptr(), db.data_ptr(), flags | capi.GM_NO_EXCHANGE, sync=True)
        ranks.append((sp, h, dnz, db))
    for r in range(1, nranks):  # the "NCCL" step: lower interface of r -> upper interface of r-1
        sp, h, dnz, db = ranks[r]
        n = h.iface_count(0)
        assert n == ranks[r - 1][1].iface_count(1) and n > 0
        buf = torch.empty(n, dtype=torch.float64, device="cuda")
        h.iface_pack(dnz.data_ptr(), db.data_ptr(), buf.data_ptr(), flags)
        _, h2, dnz2, db2 = ranks[r - 1]
        h2.iface_unpack_add(dnz2.data_ptr(), db2.data_ptr(), buf.data_ptr(), flags)
    seen = np.zeros(full.n_total, dtype=np.int32)
    for sp, h, dnz, db in ranks:
        own = fe.owned_dofs(sp)
        seen[own] += 1
        ptr, idx = h.get_pattern()
        nz, b = dnz.cpu().numpy(), db.cpu().numpy()
        # owned rows: same column indices and values as the global oracle matrix
        lens = ptr[own + 1] - ptr[own]
        assert np.array_equal(lens, optr[own + 1] - optr[own])
        loc = np.concatenate([np.arange(ptr[g], ptr[g + 1]) for g in own])
        glb = np.concatenate([np.arange(optr[g], optr[g + 1]) for g in own])
        assert np.array_equal(idx[loc], oidx[glb])
        full_nz = onz.copy()
        full_nz[glb] = nz[loc]
        assert cases.block_scaled_err(full_nz, onz, blocks) <= TOL
        full_b = ob.copy()
        full_b[own] = b[own]
        assert cases.block_scaled_err(full_b, ob, rblocks) <= TOL
        h.close()
    assert np.all(seen == 1)


@pytest.mark.parametrize("n,bc,N", [(0.6, "WAVE", 16), (1.0, "SIMPLE", 20)])
def test_newton_through_the_plugin_matches_oracle_newton(gm, gpu_lib, n, bc, N):
    """north_star: 'Newton converges in the same iteration count with fields within 1e-10'.
    Same driver (NLsolve newton_ + BackTracking restated in solver.py), operator = B200 plug-in vs oracle."""
    from orc_operator import OracleOperator
    sp = cases.cart(N, bc=getattr(cases, bc))
    prm = dict(Pr=100.0, Ra=1e4, n=n) if bc == "WAVE" else dict(Pr=0.7, Ra=1e3, n=n)
    ref = gm.solver.newton(OracleOperator(sp, **prm), np.zeros(sp.n_total), ftol=1e-8, xtol=1e-50, iterations=60)
    op = gm.operator.B200FEOperator(sp, **prm)
    try:
        got = gm.solver.newton(op, np.zeros(sp.n_total), ftol=1e-8, xtol=1e-50, iterations=60)
    finally:
        op.close()
    assert ref.f_converged and got.f_converged
    assert got.iterations == ref.iterations and got.f_calls == ref.f_calls
    assert np.max(np.abs(got.x - ref.x)) <= 1e-10 * max(1.0, np.max(np.abs(ref.x)))


def test_newton_config3_annulus_through_the_gather_path(gm, gpu_lib):
    """Config 3 (validation-kuehn/kuehn.jl:30-52: concentric annulus, tags inner / outer / all, Pr=0.706, Ra=4.7e4, n=1) on a triangle
    mesh read from a .msh file: Newton + BackTracking through the plug-in (generic gather path) takes the same iterations / residual
    evaluations as through the oracle operator and ends at the same fields."""
    import tempfile
    from orc_operator import OracleOperator
    from _pkg import gm as pkg
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "co-annulus.msh")
        pkg.gmsh.write_msh(fe.annulus_model(nr=8, nt=48, jitter=0.2, seed=5), path, version="2.2")
        sp = cases.spaces(pkg.gmsh.read_msh(path), cases.KUEHN)
    prm = dict(Pr=0.706, Ra=4.7e4, n=1.0)
    ref = gm.solver.newton(OracleOperator(sp, **prm), np.zeros(sp.n_total), ftol=1e-8, xtol=1e-50, iterations=60)
    op = gm.operator.B200FEOperator(sp, **prm)
    try:
        assert op.h.info()["active_path"] == gm.capi.GM_PATH_GATHER
        got = gm.solver.newton(op, np.zeros(sp.n_total), ftol=1e-8, xtol=1e-50, iterations=60)
    finally:
        op.close()
    assert ref.f_converged and got.f_converged
    assert (got.iterations, got.f_calls, got.g_calls) == (ref.iterations, ref.f_calls, ref.g_calls)
    assert np.max(np.abs(got.x - ref.x)) <= 1e-10 * max(1.0, np.max(np.abs(ref.x)))


def test_newton_configs_1_and_study1_through_the_plugin(gm, gpu_lib):
    """north_star: 'Newton converges in the same iteration count with fields within 1e-10' on the reference's own runs:
    config 1 (examples/simple.jl: 50x50, Pr=0.7, Ra=1e3, n=1, Newton + BackTracking) and the study-1 mesh whose Nusselt
    number the reference ships (examples/study1/conv_study_square.csv:2: 62x62 wave BCs, n=0.6, Nu_bot_avg =
    6.670431887977048) -- reproduced through the GPU operator, not only through the oracle."""
    from orc_operator import OracleOperator
    for N, bc, prm, nu in ((50, cases.SIMPLE, dict(Pr=0.7, Ra=1e3, n=1.0), None),
                           (62, cases.WAVE, dict(Pr=100.0, Ra=1e4, n=0.6), 6.670431887977048)):
        sp = cases.cart(N, bc=bc)
        ref = gm.solver.newton(OracleOperator(sp, **prm), np.zeros(sp.n_total), ftol=1e-8, xtol=1e-50, iterations=60)
        op = gm.operator.B200FEOperator(sp, **prm)
        try:
            got = gm.solver.newton(op, np.zeros(sp.n_total), ftol=1e-8, xtol=1e-50, iterations=60)
        finally:
            op.close()
        assert ref.f_converged and got.f_converged
        assert got.iterations == ref.iterations and got.f_calls == ref.f_calls and got.g_calls == ref.g_calls
        assert np.max(np.abs(got.x - ref.x)) <= 1e-10 * max(1.0, np.max(np.abs(ref.x)))
        if nu is not None:
            assert abs(gm.solver.nu_bot_avg(sp, got.x) - nu) <= 1e-12 * nu


def test_trust_region_through_the_plugin(gm, gpu_lib):
    """examples/validation-basak/basak.jl:20-26 drives NLsolve's trust-region method: same driver on the GPU operator and on
    the oracle -> same accepted / rejected steps, same fields."""
    from orc_operator import OracleOperator
    sp = cases.cart(20, bc=cases.SIMPLE)
    prm = dict(Pr=0.7, Ra=1e3, n=1.0)
    ref = gm.solver.trust_region(OracleOperator(sp, **prm), np.zeros(sp.n_total), ftol=1e-8, xtol=1e-50, iterations=200)
    op = gm.operator.B200FEOperator(sp, **prm)
    try:
        got = gm.solver.trust_region(op, np.zeros(sp.n_total), ftol=1e-8, xtol=1e-50, iterations=200)
    finally:
        op.close()
    assert ref.f_converged and got.f_converged
    assert (got.iterations, got.f_calls, got.g_calls) == (ref.iterations, ref.f_calls, ref.g_calls)
    assert np.max(np.abs(got.x - ref.x)) <= 1e-10 * max(1.0, np.max(np.abs(ref.x)))


def test_config3_annulus_kuehn_size(gm, gpu_lib):
    """BASELINE config 3 (validation-kuehn): concentric annulus R1=5/8, R2=13/8, Pr=0.706, Ra=4.7e4, n=1, tags
    inner/outer/all; ~4.6 k irregular triangles (the .msh files are LFS pointers, so the mesh is generated)."""
    sp = cases.spaces(fe.annulus_model(nr=18, nt=128, jitter=0.35, seed=7), cases.KUEHN)
    assert sp.model.ncells == 2 * 18 * 128
    x = fe.splitmix_iterate(sp.n_total, 1234)
    info = _check(gm, sp, dict(Pr=0.706, Ra=4.7e4, n=1.0), "csc", gm.capi.GM_PATH_AUTO, x)
    assert info["active_path"] == gm.capi.GM_PATH_GATHER
    info = _check(gm, sp, dict(Pr=0.706, Ra=4.7e4, n=1.0), "csc", gm.capi.GM_PATH_SCATTER, x)
    assert info["active_path"] == gm.capi.GM_PATH_SCATTER and info["ncolors"] >= 6


@pytest.mark.gpu
@pytest.mark.parametrize("nranks", [2, 3])
@pytest.mark.parametrize("layout", ["csc", "csr"])
def test_cell_range_partition_emulated_on_one_gpu(gm, gpu_lib, nranks, layout):
    """Unstructured meshes on several GPUs (gather path): contiguous cell ranges + ghost cells, no exchange step.  Every rank's
    handle is created on this GPU in turn; the lists and residual entries of the dofs it owns must equal the global oracle's,
    the owned sets must partition the dofs, and gm_info.n_owned must agree with the host-side ownership rule."""
    import torch
    capi = gm.capi
    model = fe.square_tri_model(9, jitter=0.25, seed=3)
    spg = cases.spaces(model, cases.TURAN)
    prm = dict(Pr=100.0, Ra=1e5, n=1.4, jac_scaling=1e-3)
    x = fe.splitmix_iterate(spg.n_total, 4321)
    orc = Oracle(spg, **prm)
    optr, oidx = orc.pattern(layout)
    onz, ob = orc.assemble(x, layout)
    blocks_nz, blocks_b = cases.nz_blocks(spg, optr, oidx), cases.res_blocks(spg)
    got_nz, got_b = np.full(onz.size, np.nan), np.full(ob.size, np.nan)
    seen = np.zeros(spg.n_total, dtype=np.int32)
    for r in range(nranks):
        sp = fe.build_rank_spaces_cells(model, cases.TURAN["V"], cases.TURAN["Ttags"], cases.TURAN["Texpr"], r, nranks)
        assert sp.part[2].size > 0  # (ghost cells)
        h = capi.Handle(sp, layout=layout, path=capi.GM_PATH_GATHER, rank=r, nranks=nranks)
        try:
            h.set_params(**prm)
            nnz_local = h.pattern()
            dx = torch.from_numpy(x).cuda()
            dnz = torch.full((nnz_local,), float("nan"), dtype=torch.float64, device="cuda")
            db = torch.full((spg.n_total,), float("nan"), dtype=torch.float64, device="cuda")
            torch.cuda.synchronize()
            h.assemble_device(dx.data_ptr(), dnz.data_ptr(), db.data_ptr(), capi.GM_JACOBIAN | capi.GM_RESIDUAL, sync=True)
            own = fe.owned_dofs(sp)
            assert h.info()["n_owned"] == own.size
            # GM_CHECK_FINITE on a rank: entries of dofs it does not own are never assembled and must not trip the scan
            h.assemble_device(dx.data_ptr(), dnz.data_ptr(), db.data_ptr(), capi.GM_JACOBIAN | capi.GM_RESIDUAL | capi.GM_CHECK_FINITE, sync=True)
            ptr, idx, val, b = h.get_rows(own, dnz.data_ptr(), db.data_ptr())
            for k, dof in enumerate(own):
                lo, hi = optr[dof], optr[dof + 1]
                assert np.array_equal(idx[ptr[k]:ptr[k + 1]], oidx[lo:hi]), "owned list must be the global list"
                got_nz[lo:hi] = val[ptr[k]:ptr[k + 1]]
            got_b[own] = b
            seen[own] += 1
        finally:
            h.close()
    assert np.all(seen == 1)
    assert cases.block_scaled_err(got_nz, onz, blocks_nz) <= TOL
    assert cases.block_scaled_err(got_b, ob, blocks_b) <= TOL


def test_gather_path_rejects_a_shared_pressure_dof(gm, gpu_lib):
    """The gather path writes pressure lists straight from their one cell (P1-disc, Solver.jl:89-90).  Tables in which two cells
    share a pressure dof must be refused by gm_pattern (GM_ERR_ARG), not assembled wrongly; the scatter path takes them."""
    import copy
    sp = copy.copy(cases.spaces(fe.square_tri_model(4, jitter=0.2, seed=9), cases.TURAN))
    sp.cell_dofs_p = sp.cell_dofs_p.copy()
    sp.cell_dofs_p[1, 0] = sp.cell_dofs_p[0, 0]   # cells 0 and 1 now share a pressure dof (the id of cell 1's own dof stays unused)
    h = gm.capi.Handle(sp, path=gm.capi.GM_PATH_GATHER)
    try:
        with pytest.raises(gm.capi.GmError) as ei:
            h.pattern()
        assert ei.value.code == -1 and "discontinuous pressure" in str(ei.value)
    finally:
        h.close()
    x = fe.splitmix_iterate(sp.n_total, 3)
    _check(gm, sp, PARAMS[0], "csc", gm.capi.GM_PATH_SCATTER, x)


def test_host_api_on_a_partition_rank_copies_only_its_dof_ranges(gm, gpu_lib):
    """gm_assemble (host arrays) on one rank of a cell-range partition: the iterate is read and the residual written only on the
    dof ranges the rank's cells touch (gm_times.h2d_bytes < 8 n); owned entries equal the global oracle, entries outside the
    touched ranges keep what the caller's array held."""
    model = fe.square_tri_model(24, jitter=0.2, seed=6)
    spg = cases.spaces(model, cases.TURAN)
    prm = dict(Pr=100.0, Ra=1e5, n=0.6)
    x = fe.splitmix_iterate(spg.n_total, 77)
    orc = Oracle(spg, **prm)
    optr, oidx = orc.pattern("csc")
    onz, ob = orc.assemble(x, "csc")
    nranks, r = 4, 2
    sp = fe.build_rank_spaces_cells(model, cases.TURAN["V"], cases.TURAN["Ttags"], cases.TURAN["Texpr"], r, nranks)
    h = gm.capi.Handle(sp, path=gm.capi.GM_PATH_GATHER, rank=r, nranks=nranks)
    try:
        h.set_params(**prm)
        nnz_local = h.pattern()
        nz, b = np.full(nnz_local, np.nan), np.full(spg.n_total, -7.0)
        h.assemble(x, nz, b)
        t = h.timing()
        assert 0 < t["h2d_bytes"] < 8 * spg.n_total * 0.6          # a quarter of the mesh (+ ghosts), not the whole iterate
        own = fe.owned_dofs(sp)
        assert np.max(np.abs(b[own] - ob[own])) <= TOL * np.max(np.abs(ob))
        touched = np.zeros(spg.n_total, dtype=bool)
        for tab, o in ((sp.U.cell_dofs, sp.offsets[0]), (sp.cell_dofs_p, sp.offsets[1]), (sp.T.cell_dofs, sp.offsets[2])):
            d = tab.astype(np.int64).reshape(-1)
            touched[d[d > 0] - 1 + o] = True
        gap = max(16, min(65536, spg.n_total // 256))   # (runs closer than this are merged into one copy)
        far = ~np.convolve(touched.astype(np.int32), np.ones(2 * gap + 1, dtype=np.int32), mode="same").astype(bool)
        assert far.sum() > spg.n_total // 4 and np.all(b[far] == -7.0)
        # nzval through the host path: lists of a sample of owned dofs
        lp, _ = h.get_pattern(with_idx=False)
        for k, dof in enumerate(own[::7]):
            got = nz[lp[dof]:lp[dof + 1]]
            ref = onz[optr[dof]:optr[dof + 1]]
            assert got.size == ref.size and np.max(np.abs(got - ref)) <= 1e-12 * max(1.0, np.max(np.abs(onz)))
    finally:
        h.close()
