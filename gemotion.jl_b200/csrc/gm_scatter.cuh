// gm_scatter.cuh -- generic numeric path: fused fp64 cell kernel + colour-ordered, atomic-free
// scatter.  Works for any conforming mesh of QUAD (Q2/P1disc/Q2) or TRI (P2/P1disc/P2) cells.
//
// Replaces jacobian!/residual!/residual_and_jacobian! of Gridap's FEOperatorFromWeakForm for the
// integrands declared at /root/reference/src/Solver.jl:111-152 (per-point arithmetic restated in
// SURVEY.md A.4).  One 32-lane group per cell: lanes first stage the cell's dof values, the
// physical shape gradients and the per-quadrature-point state (u, grad u, T, grad T, p, D, mu,
// kappa) in shared memory, then lane M owns major local dof M (a CSC column or CSR row of the
// cell matrix) and walks the minors, summing over quadrature points.  Cells of one colour share
// no dof, so the read-modify-write into nzval / b needs no atomics and is deterministic.
#pragma once
#include "gm_internal.h"
#include "gm_pattern.cuh"

__constant__ GmConst c_gm;

struct GmScatterArgs {
  int64_t ncells;
  const int32_t *gdofs;
  const double *dvals;
  const double *coords;
  const int32_t *cellv;
  const int64_t *ptr;
  const uint16_t *rank;
  const double *x;
  double *nz;
  double *b;
  // colour selection: explicit list, or Cartesian parity class
  const int32_t *cells;
  int64_t nsel;
  int cart_color;  // -1: use `cells`; else colour = (cx&1) + 2*(cy&1)
  unsigned flags;
};

struct GmQP {
  double w, u[2], G[2][2], T, dT[2], p, D[2][2], mu, kap, conv[2], divu, udT;
};

template <int NN, int NQ, int NV>
__global__ void __launch_bounds__(128) k_scatter(GmScatterArgs a) {
  constexpr int LD = 3 * NN + 3;
  constexpr int CPB = 4;
  __shared__ double sval[CPB][LD];
  __shared__ double sdph[CPB][NQ][NN][2];
  __shared__ double sS[CPB][NQ][NN][2];
  __shared__ double sug[CPB][NQ][NN];
  __shared__ GmQP sq[CPB][NQ];
  __shared__ int32_t sg[CPB][LD];

  const int lane = threadIdx.x & 31, cw = threadIdx.x >> 5;
  const int64_t t = (int64_t)blockIdx.x * CPB + cw;
  const bool active = t < a.nsel;
  int64_t c = 0;
  if (active) {
    if (a.cart_color >= 0) {
      const int px = a.cart_color & 1, py = a.cart_color >> 1;
      const int ncx = (c_gm.nx - px + 1) >> 1;
      const int64_t iy = t / ncx;
      const int ix = (int)(t - iy * ncx);
      c = (int64_t)(2 * iy + py) * c_gm.nx + (2 * ix + px);
    } else {
      c = a.cells[t];
    }
  }
  // ---- stage dof ids and values
  if (active && lane < LD) {
    const int32_t d = a.gdofs[c * LD + lane];
    sg[cw][lane] = d;
    sval[cw][lane] = d > 0 ? a.x[d - 1] : a.dvals[-d - 1];
  }
  __syncwarp();
  // ---- physical gradients (every lane recomputes the tiny Jacobian; lanes < NN store)
  if (active) {
    for (int q = 0; q < NQ; ++q) {
      double J00, J01, J10, J11;
      if (c_gm.cartesian) {
        J00 = c_gm.hx; J01 = 0; J10 = 0; J11 = c_gm.hy;
      } else {
        J00 = J01 = J10 = J11 = 0;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const int64_t vid = a.cellv[c * NV + v];
          const double X = a.coords[2 * vid], Y = a.coords[2 * vid + 1];
          const double g0 = c_gm.dgphi[(q * NV + v) * 2], g1 = c_gm.dgphi[(q * NV + v) * 2 + 1];
          J00 += X * g0; J01 += X * g1; J10 += Y * g0; J11 += Y * g1;
        }
      }
      const double det = J00 * J11 - J01 * J10;
      const double i00 = J11 / det, i01 = -J01 / det, i10 = -J10 / det, i11 = J00 / det;
      if (lane < NN) {
        const double d0 = c_gm.dphi[(q * NN + lane) * 2], d1 = c_gm.dphi[(q * NN + lane) * 2 + 1];
        sdph[cw][q][lane][0] = i00 * d0 + i10 * d1;
        sdph[cw][q][lane][1] = i01 * d0 + i11 * d1;
      }
      if (lane == 0) sq[cw][q].w = c_gm.w[q] * fabs(det);
    }
  }
  __syncwarp();
  // ---- per-point state: lane q
  if (active && lane < NQ) {
    const int q = lane;
    double u0 = 0, u1 = 0, G00 = 0, G01 = 0, G10 = 0, G11 = 0, T = 0, dT0 = 0, dT1 = 0, p = 0;
#pragma unroll
    for (int i = 0; i < NN; ++i) {
      const double ph = c_gm.phi[q * NN + i], dx = sdph[cw][q][i][0], dy = sdph[cw][q][i][1];
      const double ux = sval[cw][i], uy = sval[cw][NN + i], tt = sval[cw][2 * NN + 3 + i];
      u0 += ph * ux; u1 += ph * uy;
      G00 += dx * ux; G01 += dx * uy; G10 += dy * ux; G11 += dy * uy;
      T += ph * tt; dT0 += dx * tt; dT1 += dy * tt;
    }
#pragma unroll
    for (int m = 0; m < 3; ++m) p += c_gm.psi[q * 3 + m] * sval[cw][2 * NN + m];
    GmQP &s = sq[cw][q];
    s.u[0] = u0; s.u[1] = u1;
    s.G[0][0] = G00; s.G[0][1] = G01; s.G[1][0] = G10; s.G[1][1] = G11;
    s.T = T; s.dT[0] = dT0; s.dT[1] = dT1; s.p = p;
    const double D01 = 0.5 * (G01 + G10);
    s.D[0][0] = G00; s.D[0][1] = D01; s.D[1][0] = D01; s.D[1][1] = G11;
    const double gam = c_gm.reg + 2.0 * (G00 * G00 + 2.0 * D01 * D01 + G11 * G11);
    if (c_gm.n == 1.0) {
      s.mu = 1.0; s.kap = 0.0;
    } else {
      s.mu = pow(gam, 0.5 * (c_gm.n - 1.0));
      s.kap = 0.5 * (c_gm.n - 1.0) * pow(gam, 0.5 * (c_gm.n - 3.0)) * 4.0;
    }
    s.conv[0] = u0 * G00 + u1 * G10;
    s.conv[1] = u0 * G01 + u1 * G11;
    s.divu = G00 + G11;
    s.udT = u0 * dT0 + u1 * dT1;
  }
  __syncwarp();
  if (active && lane < NN) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const GmQP &s = sq[cw][q];
      const double dx = sdph[cw][q][lane][0], dy = sdph[cw][q][lane][1];
      sS[cw][q][lane][0] = s.D[0][0] * dx + s.D[1][0] * dy;
      sS[cw][q][lane][1] = s.D[0][1] * dx + s.D[1][1] * dy;
      sug[cw][q][lane] = s.u[0] * dx + s.u[1] * dy;
    }
  }
  __syncwarp();
  if (!active || lane >= LD) return;
  const int M = lane;
  const int32_t gM = sg[cw][M];
  const int fM = M < 2 * NN ? 0 : (M < 2 * NN + 3 ? 1 : 2);
  const int iM = fM == 0 ? (M % NN) : (fM == 1 ? M - 2 * NN : M - 2 * NN - 3);
  const int aM = fM == 0 ? M / NN : 0;
  const double Pr = c_gm.Pr, RaPr = c_gm.Ra * c_gm.Pr;
  const double gv[2] = {c_gm.gx, c_gm.gy};

  // ---- residual (M is the test dof)
  if ((a.flags & GM_RESIDUAL) && gM > 0) {
    double r = 0;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const GmQP &s = sq[cw][q];
      double v;
      if (fM == 0) {
        const double ph = c_gm.phi[q * NN + iM];
        v = 2.0 * Pr * s.mu * sS[cw][q][iM][aM] + ph * s.conv[aM] - sdph[cw][q][iM][aM] * s.p - RaPr * s.T * gv[aM] * ph;
      } else if (fM == 1) {
        v = c_gm.psi[q * 3 + iM] * s.divu;
      } else {
        v = s.dT[0] * sdph[cw][q][iM][0] + s.dT[1] * sdph[cw][q][iM][1] + s.udT * c_gm.phi[q * NN + iM];
      }
      r += s.w * v;
    }
    a.b[gM - 1] += r;
  }
  if (!(a.flags & GM_JACOBIAN) || gM <= 0) return;
  const bool csr = c_gm.layout == GM_LAYOUT_CSR;
  const int64_t base = a.ptr[gM - 1];
  const uint16_t *rk = a.rank + c * LD * LD + M;
  for (int m = 0; m < LD; ++m) {
    const uint16_t pos = rk[m * LD];
    if (pos == 0xFFFF) continue;
    const int fm = m < 2 * NN ? 0 : (m < 2 * NN + 3 ? 1 : 2);
    const int im = fm == 0 ? (m % NN) : (fm == 1 ? m - 2 * NN : m - 2 * NN - 3);
    const int am = fm == 0 ? m / NN : 0;
    // (test, trial) = (M, m) for CSR rows, (m, M) for CSC columns
    const int ft = csr ? fM : fm, it = csr ? iM : im, at = csr ? aM : am;
    const int fr = csr ? fm : fM, ir = csr ? im : iM, ar = csr ? am : aM;
    double val = 0;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const GmQP &s = sq[cw][q];
      double v;
      if (ft == 0) {
        if (fr == 0) {  // UU: db + dc  (Solver.jl:132-137)
          const double gg = sdph[cw][q][it][0] * sdph[cw][q][ir][0] + sdph[cw][q][it][1] * sdph[cw][q][ir][1];
          const double dab = at == ar ? 1.0 : 0.0;
          v = 2.0 * Pr * (s.kap * sS[cw][q][it][at] * sS[cw][q][ir][ar] +
                          s.mu * 0.5 * (gg * dab + sdph[cw][q][it][ar] * sdph[cw][q][ir][at])) +
              c_gm.phi[q * NN + it] * (sug[cw][q][ir] * dab + c_gm.phi[q * NN + ir] * s.G[ar][at]);
        } else if (fr == 1) {  // UP
          v = -sdph[cw][q][it][at] * c_gm.psi[q * 3 + ir];
        } else {  // UT buoyancy
          v = -RaPr * gv[at] * c_gm.phi[q * NN + it] * c_gm.phi[q * NN + ir];
        }
      } else if (ft == 1) {  // PU
        v = c_gm.psi[q * 3 + it] * sdph[cw][q][ir][ar];
      } else {
        if (fr == 0) {  // TU
          v = c_gm.phi[q * NN + it] * c_gm.phi[q * NN + ir] * s.dT[ar];
        } else {  // TT
          v = sdph[cw][q][it][0] * sdph[cw][q][ir][0] + sdph[cw][q][it][1] * sdph[cw][q][ir][1] +
              c_gm.phi[q * NN + it] * sug[cw][q][ir];
        }
      }
      val += s.w * v;
    }
    a.nz[base + pos] += c_gm.jac_scaling * val;
  }
}

static int gm_launch_scatter(gm_handle *h, const double *d_x, double *d_nz, double *d_b, unsigned flags,
                             int64_t *launches) {
  GmScatterArgs a;
  a.ncells = h->ncells;
  a.gdofs = h->gdofs; a.dvals = h->dvals; a.coords = h->coords; a.cellv = h->cellv;
  a.ptr = h->ptr; a.rank = h->rank; a.x = d_x; a.nz = d_nz; a.b = d_b; a.flags = flags;
  if ((flags & GM_JACOBIAN) && d_nz) GM_CUDA(cudaMemsetAsync(d_nz, 0, sizeof(double) * (size_t)h->nnz, h->stream));
  if ((flags & GM_RESIDUAL) && d_b) GM_CUDA(cudaMemsetAsync(d_b, 0, sizeof(double) * (size_t)h->n_total, h->stream));
  for (int col = 0; col < h->ncolors; ++col) {
    if (h->d.cartesian) {
      const int px = col & 1, py = col >> 1;
      const int64_t ncx = (h->d.nx - px + 1) / 2, ncy = (h->d.ny - py + 1) / 2;
      a.cart_color = col; a.cells = nullptr; a.nsel = ncx * ncy;
    } else {
      a.cart_color = -1;
      a.cells = h->color_cells + h->color_ptr[col];
      a.nsel = h->color_ptr[col + 1] - h->color_ptr[col];
    }
    if (a.nsel <= 0) continue;
    const unsigned nb = (unsigned)((a.nsel + 3) / 4);
    if (h->d.cell_type == GM_QUAD) k_scatter<9, 4, 4><<<nb, 128, 0, h->stream>>>(a);
    else k_scatter<6, 3, 3><<<nb, 128, 0, h->stream>>>(a);
    ++*launches;
  }
  GM_CUDA(cudaGetLastError());
  return GM_OK;
}
