// gm_post.cuh -- post-processing assemblies that follow the Newton solve in the reference
// (/root/reference/src/Solver.jl:197-292), on the device behind the same handle:
//   * stream function  (Solver.jl:200-210): P1 Laplacian  a(u,v) = ∫ ∇v·∇u dΩ  and  b(v) = ∫ v (∂x u_y - ∂y u_x) dΩ   [Measure(Ω,2)]
//   * heat function    (Solver.jl:254-273): P1 Laplacian on Measure(Ω,4)  and  b(v) = ∫ -∇v·(R·(T u - ∇T)) dΩ2,  R·w = (w_y, -w_x)
//   * entropies        (Solver.jl:281-292): nodal interpolation of |∇T|^2 and 1e-4 (2(∂x ux^2 + ∂y uy^2) + (∂y ux + ∂x uy)^2) into a
//     P2 H1 space (the last cell that contains a node wins, like Gridap's cell loop), Sth_tot, Sfl_tot, Be_avg.
// These run once per simulation, not per Newton iteration: the design goal is determinism and parity, not the last GB/s.
// Two phases without atomics or colours: (1) one thread per cell writes its local matrix / vector to a cell-major buffer;
// (2) one thread per dof walks its incident cells in ascending cell id -- the order of Gridap's serial assembly loop -- and sums
// into its own CSC column / vector entry.  Symbolic data (dof -> cells adjacency, sorted-unique pattern) is built on the host once.
#pragma once
#include <cub/cub.cuh>

#include <algorithm>
#include <vector>

#include "gm_internal.h"

struct gm_post_space {
  int64_t n = 0, ndiri = 0, nnz = 0;
  int32_t *dofs = nullptr;   // [ncells*nv] signed 1-based
  double *diri = nullptr;
  int64_t *ptr = nullptr;    // [n+1]
  int32_t *idx = nullptr;    // [nnz] sorted rows per column
  int64_t *adjptr = nullptr; // [n+1]
  int32_t *adj = nullptr;    // [.] cell*nv + local, ascending
  std::vector<int64_t> hptr;
  std::vector<int32_t> hidx;
};

struct gm_post {
  int nq[2] = {0, 0};
  double *tab[2] = {nullptr, nullptr};  // per rule: w[nq] | g[nq*nv] | dg[nq*nv*2] | phi[nq*nn] | dphi[nq*nn*2]
  double *ntab = nullptr;               // dphi_nodes[nn*nn*2] | dg_nodes[nn*nv*2]
  gm_post_space sp[2];
  int32_t *cell_nodes = nullptr;        // [ncells*nn] 0-based
  int32_t *node_last = nullptr;         // [n_nodes] (last cell containing the node)*nn + local node
  int64_t n_nodes = 0;
  double *cellbuf = nullptr;            // [ncells*(nv*nv+nv)] local matrices + vectors; reused as [ncells*nn*2] by the entropy pass
};

struct GmPostArgs {
  int64_t ncells;
  int nn, nv, nq, cartesian;
  double hx, hy;
  const int32_t *gdofs;  // combined signed ids [ncells*ld]
  const double *dvals, *coords, *x;
  const int32_t *cellv;
  const double *tab;
  const int32_t *pdofs;
  const double *pdiri;
  double *cellbuf;
};

__device__ __forceinline__ void gp_jacobian(const GmPostArgs &a, int64_t c, const double *dg /*[nv][2] at this point*/, double Ji[2][2], double &adet) {
  double J00, J01, J10, J11;
  if (a.cartesian) { J00 = a.hx; J01 = 0; J10 = 0; J11 = a.hy; }
  else {
    J00 = J01 = J10 = J11 = 0;
    for (int v = 0; v < a.nv; ++v) {
      const double *X = a.coords + 2 * (int64_t)a.cellv[c * a.nv + v];
      J00 += X[0] * dg[2 * v]; J01 += X[0] * dg[2 * v + 1]; J10 += X[1] * dg[2 * v]; J11 += X[1] * dg[2 * v + 1];
    }
  }
  const double det = J00 * J11 - J01 * J10;
  Ji[0][0] = J11 / det; Ji[0][1] = -J01 / det; Ji[1][0] = -J10 / det; Ji[1][1] = J00 / det;  // d_k = sum_l Ji[l][k] dhat_l
  adet = fabs(det);
}

// phase 1 of the stream (WHICH = 0) / heat (WHICH = 1) function problem: local P1 Laplacian and load vector of every cell
template <int WHICH>
__global__ void k_post_cells(GmPostArgs a) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= a.ncells) return;
  const int nn = a.nn, nv = a.nv, nq = a.nq, ld = 3 * nn + 3;
  const double *w = a.tab, *g = w + nq, *dg = g + nq * nv, *phi = dg + nq * nv * 2, *dphi = phi + nq * nn;
  double U[2][GM_MAXNN], T[GM_MAXNN];
  for (int i = 0; i < nn; ++i) {
    const int32_t d0 = a.gdofs[c * ld + i], d1 = a.gdofs[c * ld + nn + i], dt = a.gdofs[c * ld + 2 * nn + 3 + i];
    U[0][i] = d0 > 0 ? a.x[d0 - 1] : a.dvals[-d0 - 1];
    U[1][i] = d1 > 0 ? a.x[d1 - 1] : a.dvals[-d1 - 1];
    T[i] = dt > 0 ? a.x[dt - 1] : a.dvals[-dt - 1];
  }
  double A[GM_MAXNV][GM_MAXNV], bl[GM_MAXNV];
  for (int i = 0; i < nv; ++i) { bl[i] = 0; for (int j = 0; j < nv; ++j) A[i][j] = 0; }
  for (int q = 0; q < nq; ++q) {
    double Ji[2][2], adet;
    gp_jacobian(a, c, dg + q * nv * 2, Ji, adet);
    const double ww = w[q] * adet;
    double G[GM_MAXNV][2];
    for (int i = 0; i < nv; ++i) {
      const double d0 = dg[(q * nv + i) * 2], d1 = dg[(q * nv + i) * 2 + 1];
      G[i][0] = Ji[0][0] * d0 + Ji[1][0] * d1; G[i][1] = Ji[0][1] * d0 + Ji[1][1] * d1;
    }
    double u[2] = {0, 0}, gu[2][2] = {{0, 0}, {0, 0}}, Tq = 0, gT[2] = {0, 0};  // gu[k][a] = d_k u_a
    for (int i = 0; i < nn; ++i) {
      const double ph = phi[q * nn + i], d0 = dphi[(q * nn + i) * 2], d1 = dphi[(q * nn + i) * 2 + 1];
      const double px = Ji[0][0] * d0 + Ji[1][0] * d1, py = Ji[0][1] * d0 + Ji[1][1] * d1;
      u[0] += ph * U[0][i]; u[1] += ph * U[1][i]; Tq += ph * T[i];
      gu[0][0] += px * U[0][i]; gu[1][0] += py * U[0][i]; gu[0][1] += px * U[1][i]; gu[1][1] += py * U[1][i];
      gT[0] += px * T[i]; gT[1] += py * T[i];
    }
    for (int i = 0; i < nv; ++i)
      for (int j = 0; j < nv; ++j) A[i][j] += ww * (G[i][0] * G[j][0] + G[i][1] * G[j][1]);
    if (WHICH == 0) {
      const double om = gu[0][1] - gu[1][0];  // grad(u) (.) TensorValue(0,-1,1,0)
      for (int i = 0; i < nv; ++i) bl[i] += ww * g[q * nv + i] * om;
    } else {
      const double F0 = Tq * u[0] - gT[0], F1 = Tq * u[1] - gT[1];
      for (int i = 0; i < nv; ++i) bl[i] -= ww * (G[i][0] * F1 - G[i][1] * F0);  // -grad(v) . (R . F),  R . F = (F1, -F0)
    }
  }
  // AffineFEOperator: Dirichlet columns go to the right-hand side
  for (int j = 0; j < nv; ++j) {
    const int32_t dj = a.pdofs[c * nv + j];
    if (dj < 0) {
      const double gD = a.pdiri[-dj - 1];
      for (int i = 0; i < nv; ++i) bl[i] -= A[i][j] * gD;
    }
  }
  double *o = a.cellbuf + c * (nv * nv + nv);
  for (int i = 0; i < nv; ++i) { for (int j = 0; j < nv; ++j) o[i * nv + j] = A[i][j]; o[nv * nv + i] = bl[i]; }
}

// phase 2: dof j sums its incident cells (ascending cell id) into CSC column j and b[j]
__global__ void k_post_gather(int64_t n, int nv, const int64_t *__restrict__ adjptr, const int32_t *__restrict__ adj, const int32_t *__restrict__ pdofs,
                              const int64_t *__restrict__ ptr, const int32_t *__restrict__ idx, const double *__restrict__ cellbuf,
                              double *__restrict__ nz, double *__restrict__ b) {
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int64_t p0 = ptr[j], p1 = ptr[j + 1];
  for (int64_t p = p0; p < p1; ++p) nz[p] = 0.0;
  double bj = 0.0;
  for (int64_t e = adjptr[j]; e < adjptr[j + 1]; ++e) {
    const int64_t c = adj[e] / nv;
    const int lj = adj[e] - (int)(c * nv);
    const double *o = cellbuf + c * (nv * nv + nv);
    bj += o[nv * nv + lj];
    for (int i = 0; i < nv; ++i) {
      const int32_t di = pdofs[c * nv + i];
      if (di <= 0) continue;
      int64_t p = p0;
      while (idx[p] != di - 1) ++p;   // (columns have at most nine rows)
      nz[p] += o[i * nv + lj];
    }
  }
  b[j] = bj;
}

// entropy pass: nodal values of both fields for every (cell, local node)
__global__ void k_post_entropy_cells(GmPostArgs a, const double *__restrict__ ntab) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= a.ncells) return;
  const int nn = a.nn, nv = a.nv, ld = 3 * nn + 3;
  const double *dphn = ntab, *dgn = ntab + nn * nn * 2;
  double U[2][GM_MAXNN], T[GM_MAXNN];
  for (int i = 0; i < nn; ++i) {
    const int32_t d0 = a.gdofs[c * ld + i], d1 = a.gdofs[c * ld + nn + i], dt = a.gdofs[c * ld + 2 * nn + 3 + i];
    U[0][i] = d0 > 0 ? a.x[d0 - 1] : a.dvals[-d0 - 1];
    U[1][i] = d1 > 0 ? a.x[d1 - 1] : a.dvals[-d1 - 1];
    T[i] = dt > 0 ? a.x[dt - 1] : a.dvals[-dt - 1];
  }
  for (int p = 0; p < nn; ++p) {
    double Ji[2][2], adet;
    gp_jacobian(a, c, dgn + p * nv * 2, Ji, adet);
    double gu[2][2] = {{0, 0}, {0, 0}}, gT[2] = {0, 0};
    for (int i = 0; i < nn; ++i) {
      const double d0 = dphn[(p * nn + i) * 2], d1 = dphn[(p * nn + i) * 2 + 1];
      const double px = Ji[0][0] * d0 + Ji[1][0] * d1, py = Ji[0][1] * d0 + Ji[1][1] * d1;
      gu[0][0] += px * U[0][i]; gu[1][0] += py * U[0][i]; gu[0][1] += px * U[1][i]; gu[1][1] += py * U[1][i];
      gT[0] += px * T[i]; gT[1] += py * T[i];
    }
    const double s = gu[0][1] + gu[1][0];
    a.cellbuf[(c * nn + p) * 2] = gT[0] * gT[0] + gT[1] * gT[1];
    a.cellbuf[(c * nn + p) * 2 + 1] = 1e-4 * (2.0 * (gu[0][0] * gu[0][0] + gu[1][1] * gu[1][1]) + s * s);
  }
}
__global__ void k_post_entropy_nodes(int64_t n, const int32_t *__restrict__ node_last, const double *__restrict__ cellbuf, double *__restrict__ sth,
                                     double *__restrict__ sfl) {
  const int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (k >= n) return;
  sth[k] = cellbuf[(int64_t)node_last[k] * 2];
  sfl[k] = cellbuf[(int64_t)node_last[k] * 2 + 1];
}
// per-cell integrals of the interpolated fields with Measure(Ω, 2)
__global__ void k_post_entropy_int(GmPostArgs a, const int32_t *__restrict__ cell_nodes, const double *__restrict__ sth, const double *__restrict__ sfl,
                                   double *__restrict__ out /*[2][ncells]*/) {
  const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= a.ncells) return;
  const int nn = a.nn, nv = a.nv, nq = a.nq;
  const double *w = a.tab, *dg = w + nq + nq * nv, *phi = dg + nq * nv * 2;
  double s0 = 0, s1 = 0;
  for (int q = 0; q < nq; ++q) {
    double Ji[2][2], adet;
    gp_jacobian(a, c, dg + q * nv * 2, Ji, adet);
    double v0 = 0, v1 = 0;
    for (int i = 0; i < nn; ++i) {
      const int32_t nd = cell_nodes[c * nn + i];
      v0 += phi[q * nn + i] * sth[nd]; v1 += phi[q * nn + i] * sfl[nd];
    }
    s0 += w[q] * adet * v0; s1 += w[q] * adet * v1;
  }
  out[c] = s0; out[a.ncells + c] = s1;
}

static void gm_post_destroy(gm_handle *h) {
  gm_post *p = (gm_post *)h->post;
  if (!p) return;
  for (int r = 0; r < 2; ++r) {
    cudaFree(p->tab[r]);
    gm_post_space &s = p->sp[r];
    cudaFree(s.dofs); cudaFree(s.diri); cudaFree(s.ptr); cudaFree(s.idx); cudaFree(s.adjptr); cudaFree(s.adj);
  }
  cudaFree(p->ntab); cudaFree(p->cell_nodes); cudaFree(p->node_last); cudaFree(p->cellbuf);
  delete p;
  h->post = nullptr;
}

template <typename T>
static int gm_post_up(T **dst, const T *src, size_t count) {
  GM_CUDA(cudaMalloc((void **)dst, sizeof(T) * (count ? count : 1)));
  if (count) GM_CUDA(cudaMemcpy(*dst, src, sizeof(T) * count, cudaMemcpyHostToDevice));
  return GM_OK;
}

static int gm_post_space_build(gm_handle *h, gm_post_space &s, const int32_t *cell_dofs, int64_t n_free, int64_t n_diri, const double *diri) {
  const int nv = h->nv;
  const int64_t nc = h->ncells;
  s.n = n_free; s.ndiri = n_diri;
  std::vector<int64_t> ap((size_t)n_free + 1, 0);
  for (int64_t k = 0; k < nc * nv; ++k)
    if (cell_dofs[k] > 0) ap[cell_dofs[k]]++;
  for (int64_t k = 0; k < n_free; ++k) ap[k + 1] += ap[k];
  std::vector<int32_t> adj((size_t)ap[n_free]);
  {
    std::vector<int64_t> cur(ap.begin(), ap.end() - 1);
    for (int64_t k = 0; k < nc * nv; ++k)
      if (cell_dofs[k] > 0) adj[cur[cell_dofs[k] - 1]++] = (int32_t)k;  // ascending (cell, local)
  }
  s.hptr.assign((size_t)n_free + 1, 0);
  s.hidx.clear();
  std::vector<int32_t> buf;
  for (int64_t j = 0; j < n_free; ++j) {
    buf.clear();
    for (int64_t e = ap[j]; e < ap[j + 1]; ++e) {
      const int64_t c = adj[e] / nv;
      for (int i = 0; i < nv; ++i)
        if (cell_dofs[c * nv + i] > 0) buf.push_back(cell_dofs[c * nv + i] - 1);
    }
    std::sort(buf.begin(), buf.end());
    buf.erase(std::unique(buf.begin(), buf.end()), buf.end());
    s.hidx.insert(s.hidx.end(), buf.begin(), buf.end());
    s.hptr[j + 1] = (int64_t)s.hidx.size();
  }
  s.nnz = (int64_t)s.hidx.size();
  int rc;
  if ((rc = gm_post_up(&s.dofs, cell_dofs, (size_t)(nc * nv)))) return rc;
  if ((rc = gm_post_up(&s.diri, diri, (size_t)n_diri))) return rc;
  if ((rc = gm_post_up(&s.ptr, s.hptr.data(), s.hptr.size()))) return rc;
  if ((rc = gm_post_up(&s.idx, s.hidx.data(), s.hidx.size()))) return rc;
  if ((rc = gm_post_up(&s.adjptr, ap.data(), ap.size()))) return rc;
  if ((rc = gm_post_up(&s.adj, adj.data(), adj.size()))) return rc;
  return GM_OK;
}

static GmPostArgs gm_post_args(gm_handle *h, gm_post *p, int rule, const double *d_x) {
  GmPostArgs a;
  a.ncells = h->ncells; a.nn = h->nn; a.nv = h->nv; a.nq = p->nq[rule]; a.cartesian = h->d.cartesian;
  a.hx = h->d.h[0]; a.hy = h->d.h[1];
  a.gdofs = h->gdofs; a.dvals = h->dvals; a.coords = h->coords; a.cellv = h->cellv; a.x = d_x;
  a.tab = p->tab[rule]; a.pdofs = nullptr; a.pdiri = nullptr; a.cellbuf = p->cellbuf;
  return a;
}
