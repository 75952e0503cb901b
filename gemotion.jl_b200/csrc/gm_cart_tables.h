// gm_cart_tables.h -- constant tables of the structured path (plain C++, no CUDA dependency: also
// compiled by the host emulation build of tests/emu).
#pragma once
#include <stdint.h>
#include <string.h>

struct CartPairTab {  // per (test node, trial node): 27 doubles, all scaled by jac_scaling*w_q*|J|
  double TA[4], TB[4], TC[4], TD[4], TM[4], TP[4];
  double TK, TUT0, TUT1;
};

struct CartCell {  // constants of the P blocks (cell independent on a Cartesian mesh)
  double UP[18][3];   // jac_scaling * sum_q w (-d_a phi_i psi_p), index a*9+i
  double wpsi[4][3];  // w_q psi_p
  double psi[4][3];
  double sphi[4][9], sdx[4][9], sdy[4][9];  // physical gradients
  double w[4];
};

struct CartArgs {
  int nx, ny;
  int row_begin, row_end;  // owned cell rows of the local table (ghost rows excluded)
  const int32_t *gdofs;
  const double *dvals;
  const int64_t *ptr;
  const int32_t *cls;
  const uint8_t *ctab;
  const uint8_t *ptab;  // [nclass][81][16]  ranks of the 9 entries of a node pair, see gm_cart_pair_ranks
  const double *visc;
  const CartPairTab *tab;
  const CartCell *cc;
  const double *x;
  double *nz;
  double *b;
  unsigned flags;
  double Pr, RaPr, gx, gy, reg, n, js;
};

// pair / cell constant tables from the reference tables (w[4], phi[4*9], dphi[4*9*2], psi[4*3]), the cell
// size and the current parameters (Solver.jl:111-152)
static inline void gm_cart_fill_tables(const double *rw, const double *rphi, const double *rdphi, const double *rpsi,
                                       double hx, double hy, double Pr, double Ra, double gx, double gy, double jac_scaling,
                                       CartPairTab *tab, CartCell *cellp) {
  CartCell &cell = *cellp;
  const double detJ = hx * hy;
  for (int q = 0; q < 4; ++q) {
    cell.w[q] = rw[q] * detJ;
    for (int i = 0; i < 9; ++i) {
      cell.sphi[q][i] = rphi[q * 9 + i];
      cell.sdx[q][i] = rdphi[(q * 9 + i) * 2] / hx;
      cell.sdy[q][i] = rdphi[(q * 9 + i) * 2 + 1] / hy;
    }
    for (int m = 0; m < 3; ++m) { cell.wpsi[q][m] = cell.w[q] * rpsi[q * 3 + m]; cell.psi[q][m] = rpsi[q * 3 + m]; }
  }
  for (int a = 0; a < 2; ++a)
    for (int i = 0; i < 9; ++i)
      for (int m = 0; m < 3; ++m) {
        double v = 0;
        for (int q = 0; q < 4; ++q) v += cell.w[q] * (-(a == 0 ? cell.sdx[q][i] : cell.sdy[q][i]) * rpsi[q * 3 + m]);
        cell.UP[a * 9 + i][m] = jac_scaling * v;
      }
  for (int i = 0; i < 9; ++i)
    for (int j = 0; j < 9; ++j) {
      CartPairTab &t = tab[i * 9 + j];
      double kbar = 0, mbar = 0;
      for (int q = 0; q < 4; ++q) {
        const double w = jac_scaling * cell.w[q];
        const double xi = cell.sdx[q][i], yi = cell.sdy[q][i], xj = cell.sdx[q][j], yj = cell.sdy[q][j];
        const double K = xi * xj + yi * yj;
        t.TA[q] = w * Pr * (K + xi * xj);
        t.TB[q] = w * Pr * (K + yi * yj);
        t.TC[q] = w * Pr * (yi * xj);  // (a=0,b=1): d_b phi_i d_a phi_j
        t.TD[q] = w * Pr * (xi * yj);  // (a=1,b=0)
        t.TM[q] = w * cell.sphi[q][i] * cell.sphi[q][j];
        t.TP[q] = w * cell.sphi[q][i];
        kbar += w * K;
        mbar += w * cell.sphi[q][i] * cell.sphi[q][j];
      }
      t.TK = kbar;
      t.TUT0 = -Ra * Pr * gx * mbar;
      t.TUT1 = -Ra * Pr * gy * mbar;
    }
}

// x-chunk length of k_gath (a template parameter: 256 or 128 quad columns).  Long chunks amortise the 4-step
// pipeline warm-up (1.6 % at 256), but a rank of a multi-GPU run has few strips: the short chunk is taken when the
// long one would leave fewer than ~12 waves of CTAs (one CTA per SM), so that the last, partially filled wave stays small.
static inline int gm_gath_chunk(int nx, int nstrips, int nsm) {
  return (long long)((nx + 1 + 255) / 256) * nstrips < 12LL * nsm ? 128 : 256;
}

// x-chunk length of k_mma (a run-time argument).  A CTA sweeps `warm` warm-up columns + one chunk; `slots` CTAs are
// resident at a time (2 per SM), so a launch of nstrips x nchunks CTAs takes ceil(nstrips * nchunks / slots) waves of
// (chunk + warm) column steps.  Picks the chunk in [96, 768] that minimises that estimate among the chunkings with at
// least 12 waves (else 8, else any): measured on 4096^2 (profiles/r02_chunk_sweep.txt), few long waves lose more than the
// warm-up of short chunks costs -- the CTAs of a wave start in lock-step and their write-out bursts collide, and the tail
// of the last wave is a larger share -- e.g. 8 GPUs (65 strips): 228 columns / 4 waves 5.14 ms, 114 columns / 8 waves
// 4.83 ms; 4 GPUs: 456 / 4 waves 10.44 ms, 152 / 12 waves 9.48 ms; 1 GPU (513 strips, >= 30 waves): 96 -> 37.6, 216 -> 37.0 ms.
static inline int gm_mma_chunk(int nx, int nstrips, int slots, int warm) {
  const int min_waves[3] = {12, 8, 1};
  for (int pass = 0; pass < 3; ++pass) {
    long long best = -1;
    int best_ch = 256;
    for (int ch = 96; ch <= 768; ch += 4) {
      const long long nchunks = (nx + 1 + ch - 1) / ch;
      const long long waves = (nchunks * nstrips + slots - 1) / slots;
      if (waves < min_waves[pass]) continue;
      const long long cost = waves * (ch + warm);
      if (best < 0 || cost * 200 < best * 199) { best = cost; best_ch = ch; }  // (a longer chunk must win by 0.5 %)
    }
    if (best >= 0) return best_ch;
  }
  return 256;
}

// Pair-major copy of the class tables for k_gath: record [class][major node M][minor node m] holds the 9
// ranks (list k of M) x (dof l of m) at byte k*3+l (bytes 9..15 unused), so a lane fetches them with one
// 16-byte load.  ctab is [nclass][900] = rank[minor dof][major dof].  Entries without a position
// (Dirichlet minor or major) point at the trash element of the list's staging slot (slot size - 2), so
// the kernel stores all nine values without a branch.
static inline int gm_gath_slot_size(int M, int k) {  // staging slot (doubles) of list k of local node M
  return M < 4 ? (k < 2 ? 90 : 78) : (M < 8 ? (k < 2 ? 54 : 48) : (k < 2 ? 34 : 30));
}
static inline void gm_cart_pair_ranks(int nclass, const uint8_t *ctab, uint8_t *ptab) {
  static const int off[3] = {0, 9, 21};
  memset(ptab, 0xFF, (size_t)nclass * 81 * 16);
  for (int c = 0; c < nclass; ++c)
    for (int M = 0; M < 9; ++M)
      for (int m = 0; m < 9; ++m)
        for (int k = 0; k < 3; ++k)
          for (int l = 0; l < 3; ++l) {
            const uint8_t r = ctab[(size_t)c * 900 + (off[l] + m) * 30 + off[k] + M];
            ptab[((size_t)c * 81 + M * 9 + m) * 16 + k * 3 + l] = r == 0xFF ? (uint8_t)(gm_gath_slot_size(M, k) - 2) : r;
          }
}
