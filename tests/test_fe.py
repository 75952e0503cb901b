"""Host FE tables: closed forms vs generic algorithms, row-block partition bookkeeping (no GPU)."""
import numpy as np
import pytest

import cases
from cases import fe


@pytest.mark.parametrize("nx,ny", [(1, 1), (1, 4), (5, 1), (3, 2), (7, 5), (16, 9)])
def test_closed_form_edges_equal_first_encounter(nx, ny):
    m = fe.CartesianModel((0, 1, 0, 1), (nx, ny))
    ce, ev = fe._first_encounter_edges(m.cell_vertices, m.LOCAL_EDGES)
    assert np.array_equal(ce, m.cell_edges) and np.array_equal(ev, m.edge_vertices)


@pytest.mark.parametrize("nranks", [2, 3, 4])
def test_row_blocks_slice_the_global_tables(nranks):
    m = fe.CartesianModel((0, 1, 0, 1), (6, 13))
    full = fe.build_spaces(m, cases.TURAN["V"], cases.TURAN["Ttags"], cases.TURAN["Texpr"])
    owned_rows = []
    for r in range(nranks):
        sp = fe.build_rank_spaces(m, cases.TURAN["V"], cases.TURAN["Ttags"], cases.TURAN["Texpr"], r, nranks)
        r0, r1, glo, ghi = sp.rows
        sl = slice((r0 - glo) * 6, (r1 + ghi) * 6)
        assert np.array_equal(sp.U.cell_dofs, full.U.cell_dofs[sl])
        assert np.array_equal(sp.T.cell_dofs, full.T.cell_dofs[sl])
        assert np.array_equal(sp.cell_dofs_p, full.cell_dofs_p[sl])
        assert np.array_equal(sp.n_free, full.n_free)
        owned_rows += list(range(r0, r1))
    assert owned_rows == list(range(13))
