/*
 * nsf_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU oracle).  NOT part of the product path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may build, load or call this file.  The shipped path is the CUDA library in
 * gemotion.jl_b200/csrc behind include/gemotion_b200.h and never links against this.
 *
 * What it is: a plain-C restatement of the arithmetic that the reference declares at
 * /root/reference/src/Solver.jl:111-152 (weak-form closures res/jac) and that
 * Gridap.jl 0.18.9 (Manifest.toml:637-641; source NOT vendored under /root/reference)
 * executes in FEOperatorFromWeakForm + SparseMatrixAssembler: per-cell evaluation at the
 * quadrature points of Measure(Ω,2) (Solver.jl:107-109), symbolic CSC pattern
 * (sorted-unique rows per column over the touched blocks), serial scatter in cell order.
 *
 * PARITY STATUS: "parity unpinned" at the assembly boundary -- the reference ships no
 * matrix/residual golden and Julia/Gridap cannot run in this image.  It IS pinned at the
 * solution level: Newton on this oracle reproduces examples/study1/conv_study_square.csv:2-3
 * (Nu_bot_avg, 62x62 and 86x86) to ~1e-14 (tests/test_oracle_golden.py), which
 * discriminates quadrature degree, pressure basis, reg, gradient convention and BCs.
 *
 * Conventions:
 *   local dofs of a cell: U (node + nn*comp, comp-major), then P (3), then T (nn)
 *   cell dof ids: signed, 1-based, field-local; id>0 free, id<0 Dirichlet value diri[-id-1]
 *   global free vector x = [U | P | T]
 *   grad convention (Gridap): G[k][a] = d_k u_a
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXNN 9
#define MAXLD (3 * MAXNN + 3)
#define MAXQ 9

typedef struct {
  int32_t cell_type;      /* 0 QUAD, 1 TRI */
  int32_t cartesian;      /* 1: J = diag(hx,hy) exactly (Gridap CartesianDiscreteModel) */
  int64_t ncells;
  int64_t nverts;
  double hx, hy;
  const double *coords;         /* [nverts*2] */
  const int32_t *cell_vertices; /* [ncells*nv], 0-based */
  const int32_t *cell_dofs_u;   /* [ncells*2nn] */
  const int32_t *cell_dofs_p;   /* [ncells*3] */
  const int32_t *cell_dofs_t;   /* [ncells*nn] */
  int64_t n_free[3];
  const double *diri_u, *diri_p, *diri_t;
  int32_t nq, nn, nv;
  const double *w;     /* [nq] */
  const double *phi;   /* [nq*nn] */
  const double *dphi;  /* [nq*nn*2] */
  const double *psi;   /* [nq*3] */
  const double *dgphi; /* [nq*nv*2] */
  double Pr, Ra, n, reg, gx, gy, jac_scaling;
} orc_problem;

static inline int ld_of(const orc_problem *P) { return 3 * P->nn + 3; }

/* gather signed dof ids of one cell, with field offsets applied to the free ones */
static void cell_ids(const orc_problem *P, int64_t c, int64_t *ids) {
  const int nn = P->nn;
  const int64_t oP = P->n_free[0], oT = P->n_free[0] + P->n_free[1];
  for (int l = 0; l < 2 * nn; ++l) ids[l] = P->cell_dofs_u[c * 2 * nn + l];
  for (int m = 0; m < 3; ++m) {
    int64_t d = P->cell_dofs_p[c * 3 + m];
    ids[2 * nn + m] = d > 0 ? d + oP : d;
  }
  for (int l = 0; l < nn; ++l) {
    int64_t d = P->cell_dofs_t[c * nn + l];
    ids[2 * nn + 3 + l] = d > 0 ? d + oT : d;
  }
}

/* 1 if block (test field fr, trial field fc) is touched by the integrands (Solver.jl:125-140):
   UU UP UT PU TU TT; never PP PT TP */
static inline int touched(int fr, int fc) {
  static const int t[3][3] = {{1, 1, 1}, {1, 0, 0}, {1, 0, 1}};
  return t[fr][fc];
}
static inline int field_of(const orc_problem *P, int l) {
  return l < 2 * P->nn ? 0 : (l < 2 * P->nn + 3 ? 1 : 2);
}

/*
 * Residual (r[ld]) and Jacobian (J[ld*ld], row-major [test][trial]) of one cell.
 * Follows SURVEY A.4; terms are integrated separately and summed as ((a+b)+c)+d like
 * Gridap's lazy DomainContribution sum (Solver.jl:142-152).
 */
void orc_cell(const orc_problem *P, int64_t c, const double *x, double *r, double *J) {
  const int nn = P->nn, nq = P->nq, nv = P->nv, ld = ld_of(P);
  const int oPl = 2 * nn, oTl = 2 * nn + 3;
  double U[2][MAXNN], Pp[3], T[MAXNN];
  {
    const int64_t oP = P->n_free[0], oT = P->n_free[0] + P->n_free[1];
    for (int a = 0; a < 2; ++a)
      for (int i = 0; i < nn; ++i) {
        int32_t d = P->cell_dofs_u[c * 2 * nn + a * nn + i];
        U[a][i] = d > 0 ? x[d - 1] : P->diri_u[-d - 1];
      }
    for (int m = 0; m < 3; ++m) {
      int32_t d = P->cell_dofs_p[c * 3 + m];
      Pp[m] = d > 0 ? x[oP + d - 1] : P->diri_p[-d - 1];
    }
    for (int i = 0; i < nn; ++i) {
      int32_t d = P->cell_dofs_t[c * nn + i];
      T[i] = d > 0 ? x[oT + d - 1] : P->diri_t[-d - 1];
    }
  }
  /* per-term accumulators */
  static _Thread_local double ra[MAXLD], rb[MAXLD], rc[MAXLD], rd[MAXLD];
  static _Thread_local double Ja[MAXLD * MAXLD], Jb[MAXLD * MAXLD], Jc[MAXLD * MAXLD], Jd[MAXLD * MAXLD];
  memset(ra, 0, sizeof ra); memset(rb, 0, sizeof rb); memset(rc, 0, sizeof rc); memset(rd, 0, sizeof rd);
  if (J) { memset(Ja, 0, sizeof Ja); memset(Jb, 0, sizeof Jb); memset(Jc, 0, sizeof Jc); memset(Jd, 0, sizeof Jd); }
  const double Pr = P->Pr, Ra = P->Ra, n = P->n, reg = P->reg;
  const double g[2] = {P->gx, P->gy};

  for (int q = 0; q < nq; ++q) {
    /* geometry: Jm[k][l] = d x_k / d xi_l */
    double Jm[2][2];
    if (P->cartesian) {
      Jm[0][0] = P->hx; Jm[0][1] = 0; Jm[1][0] = 0; Jm[1][1] = P->hy;
    } else {
      Jm[0][0] = Jm[0][1] = Jm[1][0] = Jm[1][1] = 0;
      for (int v = 0; v < nv; ++v) {
        const double *X = P->coords + 2 * (int64_t)P->cell_vertices[c * nv + v];
        for (int k = 0; k < 2; ++k)
          for (int l = 0; l < 2; ++l) Jm[k][l] += X[k] * P->dgphi[(q * nv + v) * 2 + l];
      }
    }
    const double det = Jm[0][0] * Jm[1][1] - Jm[0][1] * Jm[1][0];
    /* Ji = J^{-1}: Ji[l][k] = d xi_l / d x_k ; physical d_k phi = sum_l Ji[l][k] dhat_l phi */
    double Ji[2][2] = {{Jm[1][1] / det, -Jm[0][1] / det}, {-Jm[1][0] / det, Jm[0][0] / det}};
    const double w = P->w[q] * fabs(det);
    double ph[MAXNN], dph[MAXNN][2];
    for (int i = 0; i < nn; ++i) {
      ph[i] = P->phi[q * nn + i];
      const double d0 = P->dphi[(q * nn + i) * 2 + 0], d1 = P->dphi[(q * nn + i) * 2 + 1];
      dph[i][0] = Ji[0][0] * d0 + Ji[1][0] * d1;
      dph[i][1] = Ji[0][1] * d0 + Ji[1][1] * d1;
    }
    const double *ps = P->psi + q * 3;
    /* fields at the point */
    double u[2] = {0, 0}, G[2][2] = {{0, 0}, {0, 0}}, Tq = 0, dT[2] = {0, 0}, pq = 0;
    for (int i = 0; i < nn; ++i) {
      for (int a = 0; a < 2; ++a) {
        u[a] += ph[i] * U[a][i];
        for (int k = 0; k < 2; ++k) G[k][a] += dph[i][k] * U[a][i];
      }
      Tq += ph[i] * T[i];
      for (int k = 0; k < 2; ++k) dT[k] += dph[i][k] * T[i];
    }
    for (int m = 0; m < 3; ++m) pq += ps[m] * Pp[m];
    double D[2][2];
    for (int k = 0; k < 2; ++k)
      for (int a = 0; a < 2; ++a) D[k][a] = 0.5 * (G[k][a] + G[a][k]);
    const double gam = reg + 2.0 * (D[0][0] * D[0][0] + D[0][1] * D[0][1] + D[1][0] * D[1][0] + D[1][1] * D[1][1]);
    const double mu = pow(gam, (n - 1.0) / 2.0);                          /* Solver.jl:116 */
    const double kap = ((n - 1.0) / 2.0) * pow(gam, (n - 3.0) / 2.0) * 4.0; /* Solver.jl:117-120 */
    double conv[2];
    for (int a = 0; a < 2; ++a) conv[a] = u[0] * G[0][a] + u[1] * G[1][a]; /* Solver.jl:122 */
    const double divu = G[0][0] + G[1][1];
    const double udT = u[0] * dT[0] + u[1] * dT[1];
    double S[MAXNN][2], ugrad[MAXNN];
    for (int i = 0; i < nn; ++i) {
      for (int a = 0; a < 2; ++a) S[i][a] = D[0][a] * dph[i][0] + D[1][a] * dph[i][1];
      ugrad[i] = u[0] * dph[i][0] + u[1] * dph[i][1];
    }
    /* residual */
    for (int i = 0; i < nn; ++i) {
      for (int a = 0; a < 2; ++a) {
        const int l = a * nn + i;
        ra[l] += w * (-dph[i][a] * pq - Ra * Pr * Tq * g[a] * ph[i]);   /* a: Solver.jl:125-129 */
        rb[l] += w * (2.0 * Pr * mu * S[i][a]);                          /* b: Solver.jl:131 */
        rc[l] += w * (ph[i] * conv[a]);                                  /* c: Solver.jl:136 */
      }
      ra[oTl + i] += w * (dT[0] * dph[i][0] + dT[1] * dph[i][1]);
      rd[oTl + i] += w * (udT * ph[i]);                                  /* d: Solver.jl:139 */
    }
    for (int m = 0; m < 3; ++m) ra[oPl + m] += w * ps[m] * divu;
    if (!J) continue;
    /* Jacobian */
    for (int i = 0; i < nn; ++i) {
      for (int a = 0; a < 2; ++a) {
        const int li = a * nn + i;
        for (int j = 0; j < nn; ++j) {
          const double gg = dph[i][0] * dph[j][0] + dph[i][1] * dph[j][1];
          for (int b = 0; b < 2; ++b) {
            const int lj = b * nn + j;
            const double dab = (a == b) ? 1.0 : 0.0;
            /* db: Solver.jl:132-134 */
            Jb[li * ld + lj] += w * (2.0 * Pr * (kap * S[i][a] * S[j][b] +
                                                 mu * 0.5 * (gg * dab + dph[i][b] * dph[j][a])));
            /* dc: Solver.jl:137 */
            Jc[li * ld + lj] += w * (ph[i] * (ugrad[j] * dab + ph[j] * G[b][a]));
          }
          /* a: UT buoyancy */
          Ja[li * ld + oTl + j] += w * (-Ra * Pr * g[a] * ph[i] * ph[j]);
        }
        for (int m = 0; m < 3; ++m) {
          Ja[li * ld + oPl + m] += w * (-dph[i][a] * ps[m]);
          Ja[(oPl + m) * ld + li] += w * (ps[m] * dph[i][a]);
        }
      }
      for (int j = 0; j < nn; ++j) {
        Ja[(oTl + i) * ld + oTl + j] += w * (dph[i][0] * dph[j][0] + dph[i][1] * dph[j][1]);
        /* dd: Solver.jl:140 */
        Jd[(oTl + i) * ld + oTl + j] += w * (ph[i] * ugrad[j]);
        for (int b = 0; b < 2; ++b) Jd[(oTl + i) * ld + b * nn + j] += w * (ph[i] * ph[j] * dT[b]);
      }
    }
  }
  for (int l = 0; l < ld; ++l) r[l] = ((ra[l] + rb[l]) + rc[l]) + rd[l];
  if (J)
    for (int k = 0; k < ld * ld; ++k) J[k] = P->jac_scaling * (((Ja[k] + Jb[k]) + Jc[k]) + Jd[k]);
}

/* ---------------- pattern: per column sorted-unique rows over touched blocks ---------------- */
static int cmp_i64(const void *a, const void *b) {
  int64_t x = *(const int64_t *)a, y = *(const int64_t *)b;
  return (x > y) - (x < y);
}

/*
 * Symbolic pass.  colptr[n+1] (0-based offsets) is always filled; rowval may be NULL for a
 * count-only call.  Returns nnz.  Mirrors SparseMatrixAssembler's count->allocate->insert->
 * compress sequence restated in SURVEY A.5 (structural zeros kept).
 * `transpose`=0: CSC of A (column j lists test rows i). 1: CSR of A (row i lists trial cols j).
 */
int64_t orc_pattern(const orc_problem *P, int transpose, int64_t *ptr, int64_t *idx) {
  const int ld = ld_of(P);
  const int64_t n = P->n_free[0] + P->n_free[1] + P->n_free[2];
  /* dof -> cells adjacency (CSR) */
  int64_t *cnt = (int64_t *)calloc((size_t)n + 1, sizeof(int64_t));
  int64_t ids[MAXLD];
  for (int64_t c = 0; c < P->ncells; ++c) {
    cell_ids(P, c, ids);
    for (int l = 0; l < ld; ++l)
      if (ids[l] > 0) cnt[ids[l]]++;
  }
  for (int64_t k = 0; k < n; ++k) cnt[k + 1] += cnt[k];
  int64_t *adj = (int64_t *)malloc(sizeof(int64_t) * (size_t)(cnt[n] ? cnt[n] : 1));
  int32_t *adjl = (int32_t *)malloc(sizeof(int32_t) * (size_t)(cnt[n] ? cnt[n] : 1));
  int64_t *cur = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + 1));
  memcpy(cur, cnt, sizeof(int64_t) * (size_t)(n + 1));
  for (int64_t c = 0; c < P->ncells; ++c) {
    cell_ids(P, c, ids);
    for (int l = 0; l < ld; ++l)
      if (ids[l] > 0) { adj[cur[ids[l] - 1]] = c; adjl[cur[ids[l] - 1]++] = l; }
  }
  int64_t nnz = 0;
  int64_t cap = 4096;
  int64_t *buf = (int64_t *)malloc(sizeof(int64_t) * (size_t)cap);
  ptr[0] = 0;
  for (int64_t k = 0; k < n; ++k) {
    int64_t m = 0;
    for (int64_t p = cnt[k]; p < cnt[k + 1]; ++p) {
      const int64_t c = adj[p];
      const int fk = field_of(P, adjl[p]);
      cell_ids(P, c, ids);
      for (int l = 0; l < ld; ++l) {
        if (ids[l] <= 0) continue;
        const int fl = field_of(P, l);
        const int ok = transpose ? touched(fk, fl) : touched(fl, fk);
        if (!ok) continue;
        if (m == cap) { cap *= 2; buf = (int64_t *)realloc(buf, sizeof(int64_t) * (size_t)cap); }
        buf[m++] = ids[l] - 1;
      }
    }
    qsort(buf, (size_t)m, sizeof(int64_t), cmp_i64);
    int64_t u = 0;
    for (int64_t t = 0; t < m; ++t)
      if (t == 0 || buf[t] != buf[t - 1]) {
        if (idx) idx[nnz + u] = buf[t];
        ++u;
      }
    nnz += u;
    ptr[k + 1] = nnz;
  }
  free(buf); free(cur); free(adjl); free(adj); free(cnt);
  return nnz;
}

static inline int64_t nzindex(const int64_t *ptr, const int64_t *idx, int64_t major, int64_t minor) {
  int64_t lo = ptr[major], hi = ptr[major + 1] - 1;
  while (lo <= hi) {
    int64_t mid = (lo + hi) >> 1;
    if (idx[mid] < minor) lo = mid + 1;
    else if (idx[mid] > minor) hi = mid - 1;
    else return mid;
  }
  return -1;
}

/*
 * Numeric pass: fillstored!(A,0); serial loop over cells in id order; per entry a binary
 * search (nzindex) and +=.  nzval / resid may be NULL.  Returns 0, or -1 if an
 * entry is missing from the pattern.
 */
int orc_assemble(const orc_problem *P, int transpose, const int64_t *ptr, const int64_t *idx,
                 const double *x, double *nzval, double *resid) {
  const int ld = ld_of(P);
  const int64_t n = P->n_free[0] + P->n_free[1] + P->n_free[2];
  if (nzval) memset(nzval, 0, sizeof(double) * (size_t)ptr[n]);
  if (resid) memset(resid, 0, sizeof(double) * (size_t)n);
  double r[MAXLD], J[MAXLD * MAXLD];
  int64_t ids[MAXLD];
  for (int64_t c = 0; c < P->ncells; ++c) {
    orc_cell(P, c, x, r, nzval ? J : NULL);
    cell_ids(P, c, ids);
    if (resid)
      for (int l = 0; l < ld; ++l)
        if (ids[l] > 0) resid[ids[l] - 1] += r[l];
    if (!nzval) continue;
    for (int i = 0; i < ld; ++i) {
      if (ids[i] <= 0) continue;
      const int fi = field_of(P, i);
      for (int j = 0; j < ld; ++j) {
        if (ids[j] <= 0 || !touched(fi, field_of(P, j))) continue;
        const int64_t k = transpose ? nzindex(ptr, idx, ids[i] - 1, ids[j] - 1)
                                    : nzindex(ptr, idx, ids[j] - 1, ids[i] - 1);
        if (k < 0) return -1;
        nzval[k] += J[i * ld + j];
      }
    }
  }
  return 0;
}

/*
 * Multi-threaded variant used ONLY as the reported CPU baseline (bench.py cpu_baseline):
 * cells are visited colour by colour (caller supplies a colouring in which no two cells of a
 * colour share a dof), so threads never touch the same entry.  Values equal orc_assemble up
 * to the cell summation order.
 */
int orc_assemble_colored(const orc_problem *P, int transpose, const int64_t *ptr, const int64_t *idx,
                         const double *x, double *nzval, double *resid, int ncolors,
                         const int64_t *color_ptr, const int64_t *color_cells) {
  const int ld = ld_of(P);
  const int64_t n = P->n_free[0] + P->n_free[1] + P->n_free[2];
  if (nzval) memset(nzval, 0, sizeof(double) * (size_t)ptr[n]);
  if (resid) memset(resid, 0, sizeof(double) * (size_t)n);
  int bad = 0;
  for (int col = 0; col < ncolors; ++col) {
#pragma omp parallel for schedule(static)
    for (int64_t t = color_ptr[col]; t < color_ptr[col + 1]; ++t) {
      const int64_t c = color_cells[t];
      double r[MAXLD], J[MAXLD * MAXLD];
      int64_t ids[MAXLD];
      orc_cell(P, c, x, r, nzval ? J : NULL);
      cell_ids(P, c, ids);
      if (resid)
        for (int l = 0; l < ld; ++l)
          if (ids[l] > 0) resid[ids[l] - 1] += r[l];
      if (!nzval) continue;
      for (int i = 0; i < ld; ++i) {
        if (ids[i] <= 0) continue;
        const int fi = field_of(P, i);
        for (int j = 0; j < ld; ++j) {
          if (ids[j] <= 0 || !touched(fi, field_of(P, j))) continue;
          const int64_t k = transpose ? nzindex(ptr, idx, ids[i] - 1, ids[j] - 1)
                                      : nzindex(ptr, idx, ids[j] - 1, ids[i] - 1);
          if (k < 0) { bad = 1; continue; }
          nzval[k] += J[i * ld + j];
        }
      }
    }
  }
  return bad ? -1 : 0;
}

/*
 * Row oracle (bench.py --verify, tests at the BASELINE sizes): the complete lists of a SAMPLE of
 * major dofs (CSC columns / CSR rows) -- sorted-unique minor ids, values summed over the incident
 * cells in cell-id order like the serial scatter above -- and their residual entries, rebuilt
 * directly from orc_cell.  One pass over the cell table (a byte map marks the sampled dofs), so
 * it is affordable on the 4096^2 mesh where orc_assemble is not.  `rows` ascending, 0-based global.
 * out_ptr[nrows+1] are offsets into out_idx/out_val (capacity cap).  Returns the number of
 * entries, or -1 when cap is too small.  Cells of a row-block rank: every incident cell of an
 * OWNED dof is in the local table (owned rows + upper ghost row), so the lists are complete.
 */
typedef struct { int64_t slot; int64_t minor; int64_t seq; double v; } orc_trip;
static int cmp_trip(const void *a, const void *b) {
  const orc_trip *x = (const orc_trip *)a, *y = (const orc_trip *)b;
  if (x->slot != y->slot) return x->slot < y->slot ? -1 : 1;
  if (x->minor != y->minor) return x->minor < y->minor ? -1 : 1;
  return x->seq < y->seq ? -1 : (x->seq > y->seq ? 1 : 0);
}
int64_t orc_rows(const orc_problem *P, int transpose, int64_t nrows, const int64_t *rows, const double *x,
                 int64_t *out_ptr, int64_t *out_idx, double *out_val, double *out_b, int64_t cap) {
  const int ld = ld_of(P);
  const int64_t n = P->n_free[0] + P->n_free[1] + P->n_free[2];
  unsigned char *mark = (unsigned char *)calloc((size_t)n + 1, 1);
  for (int64_t k = 0; k < nrows; ++k) mark[rows[k]] = 1;
  int64_t tcap = nrows * 4 * ld + 64, nt = 0, seq = 0;
  orc_trip *t = (orc_trip *)malloc(sizeof(orc_trip) * (size_t)tcap);
  if (out_b) for (int64_t k = 0; k < nrows; ++k) out_b[k] = 0.0;
  double r[MAXLD], J[MAXLD * MAXLD];
  int64_t ids[MAXLD];
  for (int64_t c = 0; c < P->ncells; ++c) {
    cell_ids(P, c, ids);
    int hit = 0;
    for (int l = 0; l < ld; ++l)
      if (ids[l] > 0 && mark[ids[l] - 1]) { hit = 1; break; }
    if (!hit) continue;
    orc_cell(P, c, x, r, J);
    for (int M = 0; M < ld; ++M) {
      if (ids[M] <= 0 || !mark[ids[M] - 1]) continue;
      /* slot of this major in rows[] (binary search) */
      int64_t lo = 0, hi = nrows - 1, slot = -1;
      while (lo <= hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (rows[mid] < ids[M] - 1) lo = mid + 1; else if (rows[mid] > ids[M] - 1) hi = mid - 1; else { slot = mid; break; }
      }
      if (out_b) out_b[slot] += r[M];
      const int fM = field_of(P, M);
      for (int m = 0; m < ld; ++m) {
        if (ids[m] <= 0) continue;
        const int fm = field_of(P, m);
        /* CSC: major = trial column j, minor = test row i ; CSR: major = test row, minor = trial col */
        const int ok = transpose ? touched(fM, fm) : touched(fm, fM);
        if (!ok) continue;
        if (nt == tcap) { tcap *= 2; t = (orc_trip *)realloc(t, sizeof(orc_trip) * (size_t)tcap); }
        t[nt].slot = slot; t[nt].minor = ids[m] - 1; t[nt].seq = seq++;
        t[nt].v = transpose ? J[M * ld + m] : J[m * ld + M];
        ++nt;
      }
    }
  }
  qsort(t, (size_t)nt, sizeof(orc_trip), cmp_trip);
  int64_t nout = 0, k = 0;
  for (int64_t s = 0; s < nrows; ++s) {
    out_ptr[s] = nout;
    while (k < nt && t[k].slot == s) {
      const int64_t minor = t[k].minor;
      double v = 0.0;
      while (k < nt && t[k].slot == s && t[k].minor == minor) v += t[k++].v;
      if (nout >= cap) { free(t); free(mark); return -1; }
      out_idx[nout] = minor; out_val[nout] = v; ++nout;
    }
  }
  out_ptr[nrows] = nout;
  free(t); free(mark);
  return nout;
}

/* torchrun pins OMP_NUM_THREADS=1 in its workers; bench.py's cpu_baseline leg asks for every host core */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
