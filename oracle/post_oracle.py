"""CPU oracle of the post-processing assemblies -- TEST INFRASTRUCTURE ONLY (numpy, vectorised over cells).

Restates /root/reference/src/Solver.jl:197-292 as Gridap executes it:
  * stream function  (Solver.jl:200-210):  a(u,v) = ∫ ∇v·∇u dΩ,  b(v) = ∫ v (∇uh ⊙ TensorValue(0,-1,1,0)) dΩ   [degree 2]
    with Gridap's column-major TensorValue and ∇u[i,j] = ∂_i u_j:  ∇uh ⊙ R = ∂_x u_y - ∂_y u_x;
  * heat function    (Solver.jl:254-273):  a = ∫ ∇v·∇u dΩ2,  b(v) = ∫ -∇v·(R·(Th uh - ∇Th)) dΩ2   [Measure(Ω, 4)],  R·w = (w_y, -w_x);
  * entropies        (Solver.jl:281-292):  Sth = interpolate(|∇Th|^2), Sfl = interpolate(1e-4 (2(∂x ux^2 + ∂y uy^2) + (∂y ux + ∂x uy)^2))
    into a P2 H1 space (nodal values cell by cell, the LAST cell that contains a node wins), Sth_tot = Σ∫Sth dΩ, Be_avg.
AffineFEOperator: Dirichlet columns are moved to the right-hand side (b -= A[:, D] g_D); both spaces use g_D = 0.
PARITY STATUS: the entropy fields are pinned by Sth_max / Sfl_max of examples/study1/conv_study_square.csv:2 (tests/test_post.py);
the stream / heat function systems have no golden in the reference ("unpinned"; checked by a manufactured solution).
"""
import numpy as np
import scipy.sparse as sps


def _cell_values(cell_dofs, x, off, diri):
    d = cell_dofs
    dv = diri if diri.size else np.zeros(1)
    return np.where(d > 0, x[off + np.maximum(d, 1) - 1], dv[np.maximum(-d, 1) - 1])


def _geometry(sp, dg):
    """Per cell and point: Jinv^T-mapped factor.  Returns (Ji [nc,nq,2,2] with physical d_k = sum_l Ji[l,k] dhat_l, |det| [nc,nq])."""
    m = sp.model
    nq = dg.shape[0]
    nc = m.ncells
    if m.cartesian:
        J = np.zeros((nc, nq, 2, 2))
        J[:, :, 0, 0] = m.h[0]
        J[:, :, 1, 1] = m.h[1]
    else:
        X = m.vertex_coords[m.cell_vertices]           # [nc, nv, 2]
        J = np.einsum("cvk,qvl->cqkl", X, dg)          # J[k,l] = d x_k / d xi_l
    det = J[..., 0, 0] * J[..., 1, 1] - J[..., 0, 1] * J[..., 1, 0]
    Ji = np.empty_like(J)
    Ji[..., 0, 0] = J[..., 1, 1] / det
    Ji[..., 0, 1] = -J[..., 0, 1] / det
    Ji[..., 1, 0] = -J[..., 1, 0] / det
    Ji[..., 1, 1] = J[..., 0, 0] / det
    return Ji, np.abs(det)


def _phys(Ji, dref):
    """dref [nq, n, 2] reference gradients -> physical gradients [nc, nq, n, 2]: d_k = sum_l Ji[l,k] dhat_l."""
    return np.einsum("cqlk,qnl->cqnk", Ji, dref)


def assemble_p1(sp, space, tabs, x, which):
    """(A csc [n,n], b [n]) of the stream function (which = 0, degree 2) or heat function (which = 1, degree 4) problem."""
    r = which
    w, g, dg, phi, dphi = tabs.w[r], tabs.g[r], tabs.dg[r], tabs.phi[r], tabs.dphi[r]
    nn = sp.ref.nn
    Ji, adet = _geometry(sp, dg)
    G = _phys(Ji, dg)                                   # vertex basis gradients [nc,nq,nv,2]
    P = _phys(Ji, dphi)                                 # P2 gradients          [nc,nq,nn,2]
    ww = w[None, :] * adet                              # [nc,nq]
    Uc = _cell_values(sp.U.cell_dofs, x, sp.offsets[0], sp.U.diri_values)
    Tc = _cell_values(sp.T.cell_dofs, x, sp.offsets[2], sp.T.diri_values)
    gu = np.stack([np.einsum("cqnk,cn->cqk", P, Uc[:, :nn]), np.einsum("cqnk,cn->cqk", P, Uc[:, nn:])], axis=3)  # [c,q,k,a] = d_k u_a
    A = np.einsum("cq,cqik,cqjk->cij", ww, G, G)
    if which == 0:
        om = gu[:, :, 0, 1] - gu[:, :, 1, 0]
        bl = np.einsum("cq,qi,cq->ci", ww, g, om)
    else:
        u = np.stack([phi @ Uc[:, :nn].T, phi @ Uc[:, nn:].T], axis=2).transpose(1, 0, 2)   # [c,q,a]
        T = (phi @ Tc.T).T                                                                  # [c,q]
        gT = np.einsum("cqnk,cn->cqk", P, Tc)
        F = T[..., None] * u - gT                                                           # Th uh - grad Th
        RF = np.stack([F[..., 1], -F[..., 0]], axis=2)
        bl = -np.einsum("cq,cqik,cqk->ci", ww, G, RF)
    d = space.cell_dofs
    nv = d.shape[1]
    n = space.n_free
    dv = space.diri_values if space.diri_values.size else np.zeros(1)
    gD = np.where(d < 0, dv[np.maximum(-d, 1) - 1], 0.0)                # Dirichlet values per local dof
    bl = bl - np.einsum("cij,cj->ci", A, gD)
    rows = np.repeat(d[:, :, None], nv, axis=2)
    cols = np.repeat(d[:, None, :], nv, axis=1)
    sel = (rows > 0) & (cols > 0)
    Amat = sps.coo_matrix((A[sel], (rows[sel] - 1, cols[sel] - 1)), shape=(n, n)).tocsc()
    Amat.sum_duplicates()
    Amat.sort_indices()
    b = np.zeros(n)
    fr = d > 0
    np.add.at(b, d[fr] - 1, bl[fr])
    return Amat, b


def entropy(sp, tabs, x):
    """(Sth_nodal, Sfl_nodal [n_nodes], Sth_tot, Sfl_tot, Be_avg); nodes = sp.T.cell_nodes numbering."""
    nn = sp.ref.nn
    Ji, _ = _geometry(sp, tabs.dg_n)
    P = _phys(Ji, tabs.dphi_n)                                       # [nc, node p, shape n, 2]
    Uc = _cell_values(sp.U.cell_dofs, x, sp.offsets[0], sp.U.diri_values)
    Tc = _cell_values(sp.T.cell_dofs, x, sp.offsets[2], sp.T.diri_values)
    gT = np.einsum("cpnk,cn->cpk", P, Tc)
    gu = np.stack([np.einsum("cpnk,cn->cpk", P, Uc[:, :nn]), np.einsum("cpnk,cn->cpk", P, Uc[:, nn:])], axis=3)
    sth = (gT ** 2).sum(axis=2)
    sfl = 1e-4 * (2.0 * (gu[:, :, 0, 0] ** 2 + gu[:, :, 1, 1] ** 2) + (gu[:, :, 0, 1] + gu[:, :, 1, 0]) ** 2)
    nnode = sp.T.node_coords.shape[0]
    cn = sp.T.cell_nodes
    out = []
    for cellvals in (sth, sfl):
        nodal = np.zeros(nnode)
        nodal[cn.reshape(-1)] = cellvals.reshape(-1)                 # later cells overwrite earlier ones
        out.append(nodal)
    # totals with Measure(Ω, 2)
    Ji2, adet2 = _geometry(sp, tabs.dg[0])
    ww = tabs.w[0][None, :] * adet2
    tot = [float(np.einsum("cq,qn,cn->", ww, tabs.phi[0], nodal[cn])) for nodal in out]
    return out[0], out[1], tot[0], tot[1], tot[0] / (tot[0] + tot[1])
