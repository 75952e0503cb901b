"""Sampled row oracle -- TEST INFRASTRUCTURE (used by tests/ and by bench.py --verify, outside the timed region).

At the BASELINE sizes (2048^2: nnz 2.9 G, 4096^2: nnz 11.5 G > 2^33) the whole matrix cannot be compared with the
serial oracle, so a sample of major dofs (CSC columns / CSR rows of the operator of /root/reference/src/Solver.jl:157)
is checked completely: the sorted-unique list of each sampled dof and its values are rebuilt from `orc_cell`
over the incident cells (oracle/nsf_oracle.c: orc_rows, semantics of Solver.jl:142-157) and compared with
`idx[ptr[r]:ptr[r+1]]`, `nzval[ptr[r]:ptr[r+1]]`, `b[r]` fetched from the device through gm_get_rows.
Bars: pattern rows bit-equal; values / residual within 1e-12 block-scaled (s_B over the sample).
"""
import numpy as np


def sample_rows(sp, nrandom=2000, seed=7, strip=8, chunks=(128, 256), extra_rows=()):
    """Owned major dofs to check: `nrandom` random ones + every dof of the cells at the places where an
    indexing bug of the structured kernels would show: first / last columns, both sides of every x-chunk
    boundary (multiples of 128 / 256 columns), first / last strip, strip boundaries, both interface rows of a
    row-block rank, and the tails of the three field ranges (largest list bases: > 2^31 at 2048^2, > 2^33 at 4096^2)."""
    m = sp.model
    off = sp.offsets
    n = sp.n_total
    rng = np.random.default_rng(seed)
    rows_info = getattr(sp, "rows", None)
    picked = []
    if m.cartesian:
        nx, ny_glob = m.partition
        if rows_info is not None:
            r0, r1, glo, ghi = rows_info
        else:
            r0, r1, glo, ghi = 0, ny_glob, 0, 0
        nyl = r1 - r0  # owned cell rows; local row index = glo + (row - r0)
        cols = {0, 1, 2, nx - 3, nx - 2, nx - 1, nx // 2}
        for c in chunks:
            for k in range(c, nx, c):
                if k <= 4 * c or k >= nx - 2 * c:  # the first few and the last chunk boundaries
                    cols.update({k - 2, k - 1, k, k + 1})
        rws = {0, 1, 2, nyl - 3, nyl - 2, nyl - 1, nyl // 2}
        for k in range(strip, nyl, strip):
            if k <= 2 * strip or k >= nyl - 2 * strip:
                rws.update({k - 2, k - 1, k, k + 1})
        rws.update(extra_rows)
        cols = np.array(sorted(c for c in cols if 0 <= c < nx), dtype=np.int64)
        rws = np.array(sorted(r for r in rws if 0 <= r < nyl), dtype=np.int64)
        cells = ((glo + rws)[:, None] * nx + cols[None, :]).reshape(-1)
        u = sp.U.cell_dofs[cells].astype(np.int64).reshape(-1)
        p = sp.cell_dofs_p[cells].astype(np.int64).reshape(-1)
        t = sp.T.cell_dofs[cells].astype(np.int64).reshape(-1)
        picked += [u[u > 0] - 1, p[p > 0] - 1 + off[1], t[t > 0] - 1 + off[2]]
        # random owned cells -> one random dof of each
        rc = (glo + rng.integers(0, nyl, nrandom)) * nx + rng.integers(0, nx, nrandom)
        allids = np.concatenate([sp.U.cell_dofs[rc].astype(np.int64), np.where(sp.cell_dofs_p[rc] > 0, sp.cell_dofs_p[rc].astype(np.int64) + off[1], -1),
                                 np.where(sp.T.cell_dofs[rc] > 0, sp.T.cell_dofs[rc].astype(np.int64) + off[2], -1)], axis=1)
        pick = allids[np.arange(nrandom), rng.integers(0, allids.shape[1], nrandom)]
        picked.append(pick[pick > 0] - 1)
        if rows_info is None:  # tails of the field ranges (largest list bases)
            for f in range(3):
                hi = off[f] + sp.n_free[f]
                picked.append(np.arange(max(off[f], hi - 40), hi, dtype=np.int64))
    elif getattr(sp, "part", None) is not None:  # one rank of a cell-range partition: owned dofs only, plus those next to the range ends
        from _pkg import gm
        own = gm.fe.owned_dofs(sp)
        picked.append(own[rng.integers(0, own.size, nrandom)])
        c0, c1, _ = sp.part
        nloc = c1 - c0
        edge_cells = np.unique(np.clip(np.concatenate([np.arange(0, 40), np.arange(nloc - 40, nloc)]), 0, nloc - 1))
        ids = np.concatenate([sp.U.cell_dofs[edge_cells].astype(np.int64).reshape(-1), (sp.cell_dofs_p[edge_cells].astype(np.int64) + off[1] * (sp.cell_dofs_p[edge_cells] > 0)).reshape(-1),
                              (sp.T.cell_dofs[edge_cells].astype(np.int64) + off[2] * (sp.T.cell_dofs[edge_cells] > 0)).reshape(-1)])
        picked.append(np.intersect1d(ids[ids > 0] - 1, own))
    else:
        picked.append(rng.integers(0, n, nrandom))
    rows = np.unique(np.concatenate(picked))
    if rows_info is not None:  # keep owned dofs only: drop the lower interface line
        from _pkg import gm
        rows = np.setdiff1d(rows, gm.fe.interface_dofs(sp, 0), assume_unique=True)
    return rows


def field_of(sp, ids):
    off = sp.offsets
    return (ids >= off[1]).astype(np.int8) + (ids >= off[2]).astype(np.int8)


def check_rows(sp, handle, orc, x, d_nz, d_b, layout, rows, floor=0.0):
    """Compares the device lists of `rows` with the row oracle.  d_nz / d_b: device pointers (0: skipped).
    Returns dict(rows, entries, pattern_ok, max_err_nz, max_err_b, max_base)."""
    optr, oidx, oval, ob = orc.rows(rows, x, layout)
    ptr, idx, val, b = handle.get_rows(rows, d_nz, d_b)
    pattern_ok = bool(np.array_equal(ptr, optr) and np.array_equal(idx, oidx))
    out = {"rows": int(rows.size), "entries": int(oidx.size), "pattern_ok": pattern_ok, "max_err_nz": None, "max_err_b": None}
    if pattern_ok and val is not None:
        fm = field_of(sp, np.repeat(rows, np.diff(optr)))
        fi = field_of(sp, oidx)
        err = 0.0
        for a in range(3):
            for c in range(3):
                sel = (fm == a) & (fi == c)
                if not np.any(sel):
                    continue
                r = oval[sel]
                s = float(np.max(np.abs(r)))
                if s == 0.0:
                    err = max(err, float(np.max(np.abs(val[sel]))))
                else:
                    err = max(err, float(np.max(np.abs(val[sel] - r) / np.maximum(np.abs(r), s))))
        out["max_err_nz"] = err
    if b is not None:
        fr = field_of(sp, rows)
        err = 0.0
        gmax = float(np.max(np.abs(ob))) if ob.size else 0.0
        for a in range(3):
            sel = fr == a
            if not np.any(sel):
                continue
            r = ob[sel]
            s = max(float(np.max(np.abs(r))), floor * gmax)
            if s == 0.0:
                err = max(err, float(np.max(np.abs(b[sel]))))
            else:
                err = max(err, float(np.max(np.abs(b[sel] - r) / np.maximum(np.abs(r), s))))
        out["max_err_b"] = err
    return out
