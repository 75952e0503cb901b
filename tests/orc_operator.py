"""Oracle-backed operator with the same interface as the B200 plug-in (tests only)."""
import numpy as np
import scipy.sparse as sps

from oracle.oracle import Oracle


class OracleOperator:
    def __init__(self, sp, layout="csc", **params):
        self.sp = sp
        self.orc = Oracle(sp, **params)
        self.n = sp.n_total
        self.layout = layout

    def allocate_jacobian(self):
        ptr, idx = self.orc.pattern(self.layout)
        cls = sps.csc_matrix if self.layout == "csc" else sps.csr_matrix
        return cls((np.zeros(idx.size), idx.copy(), ptr.copy()), shape=(self.n, self.n))

    def residual_(self, b, x):
        _, r = self.orc.assemble(x, self.layout, jac=False, res=True)
        b[:] = r

    def jacobian_(self, A, x):
        nz, _ = self.orc.assemble(x, self.layout, jac=True, res=False)
        A.data[:] = nz

    def residual_and_jacobian_(self, b, A, x):
        nz, r = self.orc.assemble(x, self.layout, jac=True, res=True)
        A.data[:] = nz
        b[:] = r
