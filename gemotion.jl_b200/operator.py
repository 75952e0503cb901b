"""B200FEOperator -- the Gridap `FEOperator` plug-in, host side.

Mirrors the methods Gridap 0.18.9 dispatches on for the operator built at
/root/reference/src/Solver.jl:157 (`op = FEOperator(res, jac, X, Y)`):
`allocate_jacobian`, `jacobian!`, `residual!`, `residual_and_jacobian!` (the trailing `!` is
spelled `_` here).  Every numeric call goes through the C ABI (capi.Handle -> gm_assemble);
nothing is computed on the host.  The Julia twin is julia/B200FEOperator.jl.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sps

from . import capi


class B200FEOperator:
    def __init__(self, sp, Pr=1.0, Ra=1.0, n=1.0, jac_scaling=1.0, reg=1e-3, g=(0.0, 1.0),
                 layout="csc", device=0, path=capi.GM_PATH_AUTO):
        self.sp = sp
        self.n = sp.n_total
        self.layout = layout
        self.h = capi.Handle(sp, layout=layout, device=device, path=path)
        self.h.set_params(Pr=Pr, Ra=Ra, n=n, reg=reg, g=g, jac_scaling=jac_scaling)
        self.nnz = self.h.pattern()

    # allocate_jacobian(op, uh): symbolic pattern, Gridap's SparseMatrixCSC{Float64,Int64}
    def allocate_jacobian(self):
        ptr, idx = self.h.get_pattern(index_base=0)
        cls = sps.csr_matrix if self.layout == "csr" else sps.csc_matrix
        return cls((np.zeros(self.nnz), idx, ptr), shape=(self.n, self.n))

    def allocate_residual(self):
        return np.zeros(self.n)

    def residual_(self, b, x):
        self.h.assemble(np.ascontiguousarray(x, dtype=np.float64), None, b)
        return b

    def jacobian_(self, A, x):
        self.h.assemble(np.ascontiguousarray(x, dtype=np.float64), A.data, None)
        return A

    def residual_and_jacobian_(self, b, A, x):
        self.h.assemble(np.ascontiguousarray(x, dtype=np.float64), A.data, b)
        return b, A

    def close(self):
        self.h.close()
