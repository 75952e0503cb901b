"""Parity against GRIDAP'S OWN assembly, from dumps written by tools/dump_gridap_golden.jl on a machine with Julia.

Set GM_GRIDAP_DUMP to a directory of dumps (one sub-directory per case).  Every dump is checked three ways:
  * oracle vs Gridap (CPU): pattern bit-equal, nzval / b within 1e-12 block-scaled, fed with GRIDAP'S tables
    -> turns "parity unpinned" (oracle/nsf_oracle.c header) into pinned;
  * fe.py vs Gridap (CPU): the restated numbering (cell dof ids, Dirichlet values, quadrature, bases) equals Gridap's;
  * CUDA library vs Gridap (GPU): gm_create fed with Gridap's tables, pattern bit-equal, values 1e-12.
Without GM_GRIDAP_DUMP the Gridap checks are skipped and only the loader is exercised on a dump written in the same format
from this repo's own restatement (source = "restatement": a format test, it pins nothing).
"""
import glob
import os

import numpy as np
import pytest

import cases
import gridap_dump
from cases import fe
from oracle.oracle import Oracle

TOL = 1e-12
ROOT = os.environ.get("GM_GRIDAP_DUMP", "")
DUMPS = sorted(d for d in glob.glob(os.path.join(ROOT, "*")) if os.path.exists(os.path.join(d, "meta.json"))) if ROOT else []


def _check_against_dump(D, assemble):
    """assemble(sp, prm, x) -> (ptr0, idx0, nz, b) with 0-based pattern."""
    me = D["meta"]
    sp = gridap_dump.DumpSpaces(D)
    prm = dict(Pr=me["params"]["Pr"], Ra=me["params"]["Ra"], n=me["params"]["n"], reg=me["params"]["reg"],
               g=tuple(me["params"]["g"]), jac_scaling=me["params"]["jac_scaling"])
    ptr, idx, nz, b = assemble(sp, prm, D["x"])
    base = me["index_base"]
    assert np.array_equal(ptr + base, D["colptr"]) and np.array_equal(idx + base, D["rowval"]), "pattern must be bit-exact"
    assert cases.block_scaled_err(nz, D["nzval"], cases.nz_blocks(sp, ptr, idx)) <= TOL
    assert cases.block_scaled_err(b, D["b"], cases.res_blocks(sp)) <= TOL
    return sp


def _oracle_assemble(sp, prm, x):
    orc = Oracle(sp, **prm)
    ptr, idx = orc.pattern("csc")
    nz, b = orc.assemble(x, "csc")
    return ptr, idx, nz, b


@pytest.fixture(scope="module")
def self_dump(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("selfdump"))
    out = []
    for name, sp, prm in (("cart9x7", cases.cart(9, 7, bc=cases.WAVE), dict(Pr=100.0, Ra=1e4, n=0.6)),
                          ("annulus", cases.spaces(fe.annulus_model(nr=3, nt=12, jitter=0.2, seed=3), cases.KUEHN), dict(Pr=0.706, Ra=4.7e4, n=1.0))):
        x = fe.splitmix_iterate(sp.n_total, 1234)
        gridap_dump.write_self_dump(os.path.join(d, name), sp, prm, x, Oracle(sp, **prm))
        out.append(os.path.join(d, name))
    return out


def test_loader_roundtrip_on_self_dump(self_dump):
    """Format test only (source = restatement): the loader rebuilds spaces from the files and the oracle reproduces them."""
    for d in self_dump:
        D = gridap_dump.load(d)
        assert D["meta"]["source"] == "restatement"
        _check_against_dump(D, _oracle_assemble)


@pytest.mark.gpu
def test_library_on_self_dump(gm, gpu_lib, self_dump):
    """Same through gm_create (1-based vertex ids, tables from files): the path a Gridap dump takes."""
    for d in self_dump:
        D = gridap_dump.load(d)

        def lib_assemble(sp, prm, x):
            h = gm.capi.Handle(sp, layout="csc")
            try:
                h.set_params(**prm)
                nnz = h.pattern()
                ptr, idx = h.get_pattern()
                nz, b = np.empty(nnz), np.empty(sp.n_total)
                h.assemble(np.ascontiguousarray(x), nz, b)
            finally:
                h.close()
            return ptr, idx, nz, b
        _check_against_dump(D, lib_assemble)


@pytest.mark.skipif(not DUMPS, reason="GM_GRIDAP_DUMP not set: no Gridap dump available (Julia is absent from the build image)")
@pytest.mark.parametrize("d", DUMPS, ids=[os.path.basename(d) for d in DUMPS])
def test_oracle_equals_gridap(d):
    D = gridap_dump.load(d)
    assert D["meta"]["source"] == "Gridap"
    _check_against_dump(D, _oracle_assemble)


@pytest.mark.skipif(not DUMPS, reason="GM_GRIDAP_DUMP not set")
@pytest.mark.parametrize("d", DUMPS, ids=[os.path.basename(d) for d in DUMPS])
def test_restated_numbering_equals_gridap(d):
    """fe.py (SURVEY Appendix A restatement of Gridap's Cartesian model / P2 spaces / quadrature) vs Gridap's tables."""
    D = gridap_dump.load(d)
    me = D["meta"]
    if not me["cartesian"]:
        pytest.skip("numbering restatement is checked on Cartesian models (Gmsh models are read, not restated)")
    case = me.get("case", "").split("_")[0]
    bc = {"simple": cases.SIMPLE, "turan": cases.TURAN, "wave": cases.WAVE}[case]
    sp = cases.cart(me["nx"], me["ny"], bc=bc, domain=(me["origin"][0], me["origin"][0] + me["nx"] * me["h"][0],
                                                        me["origin"][1], me["origin"][1] + me["ny"] * me["h"][1]))
    assert np.array_equal(sp.U.cell_dofs.ravel(), D["du"]) and np.array_equal(sp.cell_dofs_p.ravel(), D["dp"]) and np.array_equal(sp.T.cell_dofs.ravel(), D["dt"])
    assert np.allclose(sp.T.diri_values, D["vt"], rtol=0, atol=1e-15)
    r = sp.ref
    for a, b_ in ((r.w, D["w"]), (r.phi.ravel(), D["phi"]), (r.dphi.ravel(), D["dphi"]), (r.psi.ravel(), D["psi"])):
        assert np.allclose(a, b_, rtol=0, atol=1e-13)


@pytest.mark.gpu
@pytest.mark.skipif(not DUMPS, reason="GM_GRIDAP_DUMP not set")
@pytest.mark.parametrize("d", DUMPS, ids=[os.path.basename(d) for d in DUMPS])
def test_library_equals_gridap(gm, gpu_lib, d):
    D = gridap_dump.load(d)

    def lib_assemble(sp, prm, x):
        h = gm.capi.Handle(sp, layout="csc")
        try:
            h.set_params(**prm)
            nnz = h.pattern()
            ptr, idx = h.get_pattern()
            nz, b = np.empty(nnz), np.empty(sp.n_total)
            h.assemble(np.ascontiguousarray(x), nz, b)
        finally:
            h.close()
        return ptr, idx, nz, b
    _check_against_dump(D, lib_assemble)
