// gm_gather_kernel.cuh -- generic numeric path, second generation: cell matrices by node pairs + owner-computes list
// gather.  Any conforming QUAD (Q2/P1disc/Q2) or TRI (P2/P1disc/P2) mesh -- the Gmsh annulus of
// /root/reference/examples/validation-kuehn/kuehn.jl:13-22 -- for the integrands of /root/reference/src/Solver.jl:111-152
// (per-point arithmetic restated in SURVEY.md A.4).  Replaces the colour-ordered read-modify-write scatter of gm_scatter.cuh
// (kept as GM_PATH_SCATTER): no memset, no RMW, no colours, every list written exactly once with coalesced stores.
//
//   k_cellmat : NN threads per cell (thread = test node `it`).  The cell's threads stage dof values, inverse Jacobians, physical
//               gradients and the per-point state in shared memory; thread `it` then walks the trial nodes `ir`.  A node pair
//               (it, ir) owns NINE entries -- UU x4, UT x2, TU x2, TT x1 -- that share gradient products and phi_t phi_r, so a
//               pair costs 21 FMA per quadrature point (k_scatter: 13 per ENTRY).  UP / PU come from six per-node integrals.
//               Entries go to a cell-major intermediate `cm` ([cell][major][minor], only the touched blocks: 360 / 783 doubles
//               per TRI / QUAD cell); the pressure space is discontinuous, so the lists of pressure majors are complete after
//               one cell and are written straight into nzval (and the pressure residual into b).
//   k_lists   : one warp per group of four consecutive velocity / temperature major dofs (their lists are adjacent in nzval).  The
//               incident (cell, local index) pairs of a major are walked in ascending cell id -- Gridap's own accumulation order
//               (SparseMatrixAssembler's cell loop) -- and the segments of `cm` are added into the list in shared memory at the
//               positions of the compressed cell->nnz map (`grank`); the four lists are then stored once, as one contiguous
//               range.  No atomics: the result is bit-reproducible.  The residual entry is the same ordered sum over the cell
//               residuals `cr`.  Everything about a major sits in one 64-byte record (GgDesc); the groups are processed in an
//               order built at pattern time (gm_gather.cuh): locality windows of cells, long groups first.
//
// Layout of a cell's NEC = 2NN*LD + 3NN*NN intermediate values (cm) and map entries (grank):
//   [U majors: 2NN x LD minors][T majors: NN x (2NN velocity + NN temperature minors)]; cell residuals: cr[cell][LD]
//   (a compact array of their own: a residual-only call must not leave partially written 32-byte sectors all over cm);
// the map of the pressure majors (3 x 2NN velocity minors per cell) is a separate array (grankp).
// "major" = CSC column (trial dof) or CSR row (test dof) by the template flag.
//
// k_cellmat is persistent (a few CTAs per SM looping over batches of CPB cells) and software-pipelined with cp.async
// (LDGSTS): while batch k is computed, the dof values / vertex coordinates / pressure list bases of batch k+1 and the dof
// ids / vertex ids / pressure map of batch k+2 are in flight -- the first version (one batch per CTA, plain loads) spent
// half of its samples waiting for the x gather at 15 warps per SM.
//
// The file compiles for the device and -- with GM_EMU -- as host C++ under tests/emu/cuda_emu.h (tests/test_gather_emu.py).
#pragma once
#include <stdint.h>

#include "gm_const.h"

#define GG_LWARPS 8     // warps per CTA of k_lists

struct GgArgs {
  int64_t ncells;
  int64_t nown;       // own cells (the first nown of the tables); the others are ghosts: no pressure rows from them
  int64_t nU, oT, n;  // free velocity dofs, first temperature dof, all free dofs
  const GmConst *gc;  // reference tables + parameters (device copy owned by the handle)
  const int32_t *gdofs;
  const double *dvals;
  const double *coords;
  const int32_t *cellv;
  const int64_t *ptr;
  const uint16_t *grank;   // [ncells][NEC]
  const uint16_t *grankp;  // [ncells][6NN]
  const int64_t *adjptr;
  const int32_t *adj;
  const double *x;
  double *cm, *cr;
  double *nz, *b;
  unsigned flags;
  int nbatch;   // ceil(ncells / CPB)
  int lstride;  // doubles of shared memory per major of k_lists (>= the longest list; a warp holds GG_MPW lists)
  const struct GgDesc *desc;  // [ngroup][GG_MPW] records of the velocity / temperature majors; groups with work only, in processing order
  int64_t ngroup;
};

template <int NN, int NQ>
struct GgCellBody {
  static constexpr int LD = 3 * NN + 3;
  static constexpr int NGP = (6 * NN + 3) & ~3;  // pressure-map entries, padded to 8 bytes
  double val[LD];
  double ji[NQ][4];  // inverse Jacobian (i00 i01 i10 i11 as in k_scatter)
  double w[NQ];      // w_q |det J|
  double d[NQ][NN][2];
  double S[NQ][NN][2];
  double ug[NQ][NN];
  double G[NQ][4], dT[NQ][2], u[NQ][2];
  double mu[NQ], kap[NQ];  // Pr w mu, 2 Pr w kappa
  double rs[NQ][6];        // residual scalars: conv0 conv1 p divu udT T
  int64_t pb[3];           // list bases of the three pressure majors
  int32_t g[LD + (LD & 1)];
  uint16_t gp[NGP];
};
template <int NN, int NQ>
struct GgCell : GgCellBody<NN, NQ> {  // per-cell staging area; the size in doubles is odd so that the cells of a warp sit in different banks
  double pad[((sizeof(GgCellBody<NN, NQ>) / 8) & 1) ? 2 : 1];
};

// cp.async rings of a CTA: ids two batches ahead, values one batch ahead
template <int NN, int NV, int CPB>
struct GgRings {
  static constexpr int LD = 3 * NN + 3;
  int32_t g[3][CPB * LD];
  int32_t cv[3][CPB * NV];
  uint16_t gp[3][CPB * 6 * NN];
  double val[2][CPB * LD];
  double xy[2][CPB * NV * 2];
  int64_t pb[2][CPB * 3];
};
template <int NN, int NQ, int NV, int CPB>
struct GgSmem {
  GgRings<NN, NV, CPB> r;  // (first: 16-byte aligned destinations)
  GmConst k;               // reference tables + parameters: read in every inner loop, so not left in global memory
  GgCell<NN, NQ> cells[CPB];
};

template <int NN>
struct GgLayout {
  static constexpr int LD = 3 * NN + 3;
  static constexpr int SU = LD, ST = 3 * NN;  // segment lengths of a velocity / temperature major
  static constexpr int OFF_T = 2 * NN * SU;
  static constexpr int NEC = OFF_T + NN * ST;  // intermediate values / map entries per cell (360 / 783)
};

#ifdef GM_EMU
#define GG_SMEM(T, name) T *name = reinterpret_cast<T *>(emu_smem())
#else
#define GG_SMEM(T, name)                                     \
  extern __shared__ __align__(16) unsigned char gg_smem_raw[]; \
  T *name = reinterpret_cast<T *>(gg_smem_raw)
__device__ __forceinline__ void gg_syncwarp() { __syncwarp(); }
__device__ __forceinline__ void gg_cp8(void *sdst, const void *gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(sdst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void gg_cp16(void *sdst, const void *gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(sdst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void gg_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void gg_cp_wait_prev() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }  // all but the newest group
__device__ __forceinline__ double gg_shfl_xor(double v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ int32_t gg_shfl_i(int32_t v, int src) { return __shfl_sync(0xffffffffu, v, src); }
// word `src` of the record, for a lane-dependent `src` (every lane of the warp calls it)
__device__ __forceinline__ int32_t gg_lane_word(int32_t v, int src) { return __shfl_sync(0xffffffffu, v, src & 31); }
#endif

template <int NN, int NQ, int NV, int CPB, bool CSR>
#ifndef GG_CM_MINB
#define GG_CM_MINB 4   // resident CTAs per SM the register allocation of the triangle instance aims at (5: 128 registers with a
                       // spilled loop variable whose reload cost 16 % of the samples; measured 7.08 vs 6.85 ms)
#endif
#ifndef GG_CM_UNROLL
#define GG_CM_UNROLL 1  // unrolling of the trial-node loop
#endif
__global__ void __launch_bounds__(NN *CPB, NN == 6 ? GG_CM_MINB : 2) k_cellmat(GgArgs a) {
  using L = GgLayout<NN>;
  using Cell = GgCell<NN, NQ>;
  using Smem = GgSmem<NN, NQ, NV, CPB>;
  static_assert(sizeof(Cell) % 16 == 8, "odd number of doubles per cell");
  static_assert(NQ <= NN, "one thread per quadrature point");
  static_assert((CPB * (3 * NN + 3) * 4) % 16 == 0 && (CPB * NV * 4) % 16 == 0 && (CPB * 6 * NN * 2) % 16 == 0, "16-byte batch copies");
  constexpr int LD = L::LD, NT = NN * CPB, UNR = GG_CM_UNROLL;
  GG_SMEM(Smem, smp);
  Smem &sm = *smp;
  const int tid = threadIdx.x, lc = tid / NN, it = tid - lc * NN;
  static_assert(sizeof(GmConst) % 8 == 0, "copied by doubles");
  for (int t = tid; t < (int)(sizeof(GmConst) / 8); t += NT) reinterpret_cast<double *>(&sm.k)[t] = reinterpret_cast<const double *>(a.gc)[t];
  __syncthreads();
  const GmConst &k = sm.k;
  Cell &s = sm.cells[lc];
  const int nk = ((int)a.nbatch - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;  // batches of this CTA

  // ids of batch kk (contiguous in the tables: 16-byte copies; the tail of the last batch is clamped)
  auto issue_idx = [&](int kk) {
    if (kk < nk) {
      const int64_t c0 = ((int64_t)blockIdx.x + (int64_t)kk * gridDim.x) * CPB;
      const int64_t left = a.ncells - c0;
      const int nc = left < CPB ? (int)left : CPB, slot = kk % 3;
      if (nc == CPB) {
        for (int t = tid; t < CPB * LD / 4; t += NT) gg_cp16(&sm.r.g[slot][4 * t], a.gdofs + c0 * LD + 4 * t);
        if (!k.cartesian)
          for (int t = tid; t < CPB * NV / 4; t += NT) gg_cp16(&sm.r.cv[slot][4 * t], a.cellv + c0 * NV + 4 * t);
        for (int t = tid; t < CPB * 6 * NN / 8; t += NT) gg_cp16(&sm.r.gp[slot][8 * t], a.grankp + c0 * 6 * NN + 8 * t);
      } else {  // ragged last batch: plain copies (visible after the barrier that follows the wait)
        for (int t = tid; t < nc * LD; t += NT) sm.r.g[slot][t] = a.gdofs[c0 * LD + t];
        if (!k.cartesian)
          for (int t = tid; t < nc * NV; t += NT) sm.r.cv[slot][t] = a.cellv[c0 * NV + t];
        for (int t = tid; t < nc * 6 * NN; t += NT) sm.r.gp[slot][t] = a.grankp[c0 * 6 * NN + t];
      }
    }
    gg_cp_commit();
  };
  // values of batch kk, addressed by its ids (which have landed and are visible)
  auto issue_val = [&](int kk) {
    if (kk < nk) {
      const int64_t c0 = ((int64_t)blockIdx.x + (int64_t)kk * gridDim.x) * CPB;
      const int64_t left = a.ncells - c0;
      const int nc = left < CPB ? (int)left : CPB, slot = kk % 3, vs = kk & 1;
      for (int t = tid; t < nc * LD; t += NT) {
        const int32_t d = sm.r.g[slot][t];
        gg_cp8(&sm.r.val[vs][t], d > 0 ? a.x + (d - 1) : a.dvals + (-d - 1));
      }
      if (!k.cartesian)
        for (int t = tid; t < nc * NV; t += NT) gg_cp16(&sm.r.xy[vs][2 * t], a.coords + 2 * (int64_t)sm.r.cv[slot][t]);
      for (int t = tid; t < nc * 3; t += NT) {
        const int cl = t / 3, m = t - cl * 3;
        const int32_t gP = sm.r.g[slot][cl * LD + 2 * NN + m];
        if (gP > 0) gg_cp8(&sm.r.pb[vs][t], a.ptr + (gP - 1));
      }
    }
    gg_cp_commit();
  };
  issue_idx(0);
  issue_idx(1);
  gg_cp_wait_prev();
  __syncthreads();
  issue_val(0);

#pragma unroll 1
  for (int kb = 0; kb < nk; ++kb) {
  issue_idx(kb + 2);
  gg_cp_wait_prev();  // ids of batch kb+1 and values of batch kb have landed
  __syncthreads();
  issue_val(kb + 1);
  const int64_t c = ((int64_t)blockIdx.x + (int64_t)kb * gridDim.x) * CPB + lc;
  const bool active = c < a.ncells;
  const int slot = kb % 3, vs = kb & 1;

  // ---- (A) dof ids and values out of the rings; inverse Jacobian of point `it`
  if (active) {
    for (int l = it; l < LD; l += NN) {
      s.g[l] = sm.r.g[slot][lc * LD + l];
      s.val[l] = sm.r.val[vs][lc * LD + l];
    }
    for (int l = it; l < 6 * NN; l += NN) s.gp[l] = sm.r.gp[slot][lc * 6 * NN + l];
    if (it < 3) s.pb[it] = sm.r.pb[vs][lc * 3 + it];
    if (it < NQ) {
      const int q = it;
      double J00, J01, J10, J11;
      if (k.cartesian) {
        J00 = k.hx; J01 = 0; J10 = 0; J11 = k.hy;
      } else {
        J00 = J01 = J10 = J11 = 0;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const double X = sm.r.xy[vs][(lc * NV + v) * 2], Y = sm.r.xy[vs][(lc * NV + v) * 2 + 1];
          const double g0 = k.dgphi[(q * NV + v) * 2], g1 = k.dgphi[(q * NV + v) * 2 + 1];
          J00 += X * g0; J01 += X * g1; J10 += Y * g0; J11 += Y * g1;
        }
      }
      const double det = J00 * J11 - J01 * J10, rdet = 1.0 / det;
      s.ji[q][0] = J11 * rdet; s.ji[q][1] = -J01 * rdet; s.ji[q][2] = -J10 * rdet; s.ji[q][3] = J00 * rdet;
      s.w[q] = k.w[q] * fabs(det);
    }
  }
  __syncthreads();
  // ---- (B) physical gradients of node `it`
  if (active) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const double d0 = k.dphi[(q * NN + it) * 2], d1 = k.dphi[(q * NN + it) * 2 + 1];
      s.d[q][it][0] = s.ji[q][0] * d0 + s.ji[q][2] * d1;
      s.d[q][it][1] = s.ji[q][1] * d0 + s.ji[q][3] * d1;
    }
  }
  __syncthreads();
  // ---- (C) state at point `it` (same arithmetic as k_scatter, Solver.jl:114-123)
  if (active && it < NQ) {
    const int q = it;
    double u0 = 0, u1 = 0, G00 = 0, G01 = 0, G10 = 0, G11 = 0, T = 0, dT0 = 0, dT1 = 0, p = 0;
#pragma unroll
    for (int i = 0; i < NN; ++i) {
      const double ph = k.phi[q * NN + i], dx = s.d[q][i][0], dy = s.d[q][i][1];
      const double ux = s.val[i], uy = s.val[NN + i], tt = s.val[2 * NN + 3 + i];
      u0 += ph * ux; u1 += ph * uy;
      G00 += dx * ux; G01 += dx * uy; G10 += dy * ux; G11 += dy * uy;
      T += ph * tt; dT0 += dx * tt; dT1 += dy * tt;
    }
#pragma unroll
    for (int m = 0; m < 3; ++m) p += k.psi[q * 3 + m] * s.val[2 * NN + m];
    const double D01 = 0.5 * (G01 + G10);
    const double gam = k.reg + 2.0 * (G00 * G00 + 2.0 * D01 * D01 + G11 * G11);
    double mu = 1.0, kap = 0.0;
    if (k.n != 1.0) {
      mu = pow(gam, 0.5 * (k.n - 1.0));
      kap = 0.5 * (k.n - 1.0) * pow(gam, 0.5 * (k.n - 3.0)) * 4.0;
    }
    s.u[q][0] = u0; s.u[q][1] = u1;
    s.G[q][0] = G00; s.G[q][1] = G01; s.G[q][2] = G10; s.G[q][3] = G11;
    s.dT[q][0] = dT0; s.dT[q][1] = dT1;
    s.mu[q] = k.Pr * s.w[q] * mu;
    s.kap[q] = 2.0 * k.Pr * s.w[q] * kap;
    s.rs[q][0] = u0 * G00 + u1 * G10;
    s.rs[q][1] = u0 * G01 + u1 * G11;
    s.rs[q][2] = p;
    s.rs[q][3] = G00 + G11;
    s.rs[q][4] = u0 * dT0 + u1 * dT1;
    s.rs[q][5] = T;
  }
  __syncthreads();
  // ---- (D) S = D . grad phi and u . grad phi of node `it`
  if (active) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const double dx = s.d[q][it][0], dy = s.d[q][it][1];
      const double D01 = 0.5 * (s.G[q][1] + s.G[q][2]);
      s.S[q][it][0] = s.G[q][0] * dx + D01 * dy;
      s.S[q][it][1] = D01 * dx + s.G[q][3] * dy;
      s.ug[q][it] = s.u[q][0] * dx + s.u[q][1] * dy;
    }
  }
  __syncthreads();
  if (active) {
  const double gv[2] = {k.gx, k.gy};
  const double RaPr = k.Ra * k.Pr;

  // ---- residual (Solver.jl:142-146): test node `it`
  if (a.flags & GM_RESIDUAL) {
    double r0 = 0, r1 = 0, rT = 0, rP = 0;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const double ph = k.phi[q * NN + it], dx = s.d[q][it][0], dy = s.d[q][it][1];
      const double w = s.w[q], mu2 = 2.0 * s.mu[q];  // 2 Pr w mu
      r0 += mu2 * s.S[q][it][0] + w * (ph * s.rs[q][0] - dx * s.rs[q][2] - RaPr * s.rs[q][5] * gv[0] * ph);
      r1 += mu2 * s.S[q][it][1] + w * (ph * s.rs[q][1] - dy * s.rs[q][2] - RaPr * s.rs[q][5] * gv[1] * ph);
      rT += w * (s.dT[q][0] * dx + s.dT[q][1] * dy + s.rs[q][4] * ph);
      if (it < 3) rP += w * k.psi[q * 3 + it] * s.rs[q][3];
    }
    double *cr = a.cr + c * LD;
    cr[it] = r0;
    cr[NN + it] = r1;
    cr[2 * NN + 3 + it] = rT;
    if (it < 3) {
      cr[2 * NN + it] = rP;  // (never read: keeps the 32-byte sectors of cr completely written)
      const int32_t gP = s.g[2 * NN + it];
      if (gP > 0 && c < a.nown) a.b[gP - 1] = rP;  // discontinuous pressure: the cell's contribution is the whole entry
    }
  }
  if (a.flags & GM_JACOBIAN) {
  const double js = k.jac_scaling;
  double *cm = a.cm + c * L::NEC;
  // test-side quantities of node `it`
  double md[NQ][2], kS[NQ][2], wd[NQ][2], wph[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const double dx = s.d[q][it][0], dy = s.d[q][it][1];
    md[q][0] = s.mu[q] * dx; md[q][1] = s.mu[q] * dy;
    kS[q][0] = s.kap[q] * s.S[q][it][0]; kS[q][1] = s.kap[q] * s.S[q][it][1];
    wd[q][0] = s.w[q] * dx; wd[q][1] = s.w[q] * dy;
    wph[q] = s.w[q] * k.phi[q * NN + it];
  }
  // ---- UP / PU (Solver.jl:125-127): V[a][m] = sum_q w d_a phi_it psi_m;  UP = -V (test U, trial P), PU = +V (test P, trial U)
  {
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      double V0 = 0, V1 = 0;
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        V0 += wd[q][0] * k.psi[q * 3 + m];
        V1 += wd[q][1] * k.psi[q * 3 + m];
      }
      V0 *= js; V1 *= js;
      // the entry whose major is the velocity dof (it, a) and whose minor is pressure dof m goes to the intermediate ...
      cm[it * L::SU + 2 * NN + m] = CSR ? -V0 : V0;
      cm[(NN + it) * L::SU + 2 * NN + m] = CSR ? -V1 : V1;
      // ... the one whose major is the pressure dof straight into its (complete) list
      const int32_t gP = s.g[2 * NN + m];
      if (gP > 0 && c < a.nown) {
        const int64_t base = s.pb[m];
        const uint16_t p0 = s.gp[m * 2 * NN + it], p1 = s.gp[m * 2 * NN + NN + it];
        if (p0 != 0xFFFF) a.nz[base + p0] = CSR ? V0 : -V0;
        if (p1 != 0xFFFF) a.nz[base + p1] = CSR ? V1 : -V1;
      }
    }
  }
  // ---- node pairs (it, ir): UU x4, UT x2, TU x2, TT x1 (Solver.jl:128-140)
#pragma unroll UNR
  for (int ir = 0; ir < NN; ++ir) {
    double uu[2][2] = {{0, 0}, {0, 0}}, tu[2] = {0, 0}, ggm = 0, ggw = 0, pu = 0, ww = 0;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const double dr0 = s.d[q][ir][0], dr1 = s.d[q][ir][1], Sr0 = s.S[q][ir][0], Sr1 = s.S[q][ir][1];
      const double w2 = wph[q] * k.phi[q * NN + ir];
      ggm += md[q][0] * dr0 + md[q][1] * dr1;
      ggw += wd[q][0] * dr0 + wd[q][1] * dr1;
      pu += wph[q] * s.ug[q][ir];
      ww += w2;
      // uu[at][ar] = 2 Pr kappa S_t[at] S_r[ar] + Pr mu d_t[ar] d_r[at] + phi_t phi_r G[ar][at]
      uu[0][0] += kS[q][0] * Sr0 + md[q][0] * dr0 + w2 * s.G[q][0];
      uu[0][1] += kS[q][0] * Sr1 + md[q][1] * dr0 + w2 * s.G[q][2];
      uu[1][0] += kS[q][1] * Sr0 + md[q][0] * dr1 + w2 * s.G[q][1];
      uu[1][1] += kS[q][1] * Sr1 + md[q][1] * dr1 + w2 * s.G[q][3];
      tu[0] += w2 * s.dT[q][0];
      tu[1] += w2 * s.dT[q][1];
    }
    const double diag = ggm + pu;
    uu[0][0] += diag; uu[1][1] += diag;
    const double tt = ggw + pu;
    // (test, trial) -> (minor, major) for CSC columns, (major, minor) for CSR rows
    if (!CSR) {
      double *mU0 = cm + ir * L::SU, *mU1 = cm + (NN + ir) * L::SU, *mT = cm + L::OFF_T + ir * L::ST;
      mU0[it] = js * uu[0][0];      mU0[NN + it] = js * uu[1][0];  mU0[2 * NN + 3 + it] = js * tu[0];
      mU1[it] = js * uu[0][1];      mU1[NN + it] = js * uu[1][1];  mU1[2 * NN + 3 + it] = js * tu[1];
      mT[it] = js * (-RaPr * gv[0] * ww); mT[NN + it] = js * (-RaPr * gv[1] * ww); mT[2 * NN + it] = js * tt;
    } else {
      double *mU0 = cm + it * L::SU, *mU1 = cm + (NN + it) * L::SU, *mT = cm + L::OFF_T + it * L::ST;
      mU0[ir] = js * uu[0][0];      mU0[NN + ir] = js * uu[0][1];  mU0[2 * NN + 3 + ir] = js * (-RaPr * gv[0] * ww);
      mU1[ir] = js * uu[1][0];      mU1[NN + ir] = js * uu[1][1];  mU1[2 * NN + 3 + ir] = js * (-RaPr * gv[1] * ww);
      mT[ir] = js * tu[0];          mT[NN + ir] = js * tu[1];      mT[2 * NN + ir] = js * tt;
    }
  }
  }  // Jacobian
  }  // active
  }  // batches
}

// Everything a warp of k_lists needs to know about a major dof in ONE 64-byte record (the first version chased
// adjptr -> adj -> data: three dependent rounds of global loads per warp, and it was bound by exactly that latency).
struct GgDesc {
  int64_t p0;        // list base in nzval
  int32_t j, isU;    // the dof (-1: padding record); velocity (1) or temperature (0) major
  int32_t len, cnt;  // list length; incident cells
  uint32_t e[10];    // segment offsets (cell * NEC + offset of the major's segment) of the first ten incident cells, ascending
                     // (more: from adj[adjptr[j] + t] = cell * LD + local)
};

// One warp per GROUP of four consecutive velocity / temperature major dofs (the record array is padded so that a group
// never spans the velocity -> temperature boundary): their lists are adjacent in nzval, so the warp builds one contiguous
// range in shared memory and stores it with fully coalesced accesses.  Loads are batched by depth -- the d-th and (d+1)-th
// incident cell of all four majors are fetched before any of them is added -- so a warp pays 1 + ceil(max cnt / 2) rounds
// of global-memory latency per four majors (the one-major-per-warp version paid 2 + cnt / 2 per major and was bound by
// it).  Within a major the cells are added in ascending order.  The GROUPS are stored sorted by their first incident cell
// (gm_gather.cuh), not by dof id: warps that run at the same time then read the same neighbourhood of `cm`, which turns
// scattered 168-byte DRAM reads into L2 hits (Gridap numbers all vertex dofs before all edge dofs, so in dof order the
// 18 majors of a triangle are visited at three far-apart times).  Lane l < 16 holds word l of records 0 / 2, lane
// l >= 16 of records 1 / 3; entries are handed round by shuffles; the residual is a fixed-order butterfly over 8 lanes.
#define GG_MPW 4  // majors per warp
// segment offset of adjacency entry e = cell * LD + local (velocity or temperature major)
template <int NN>
__device__ __forceinline__ uint32_t gg_seg(int32_t e) {
  using L = GgLayout<NN>;
  const int32_t cc = e / L::LD;
  const int M = e - cc * L::LD;
  return (uint32_t)cc * L::NEC + (M < 2 * NN ? M * L::SU : L::OFF_T + (M - 2 * NN - 3) * L::ST);
}
// ... and back: cell * LD + local (index of the cell residual)
template <int NN>
__device__ __forceinline__ int64_t gg_seg_to_adj(uint32_t sg) {
  using L = GgLayout<NN>;
  const uint32_t cc = sg / L::NEC, r = sg - cc * L::NEC;
  return (int64_t)cc * L::LD + (r < L::OFF_T ? r / L::SU : 2 * NN + 3 + (r - L::OFF_T) / L::ST);
}
template <int NN>
__global__ void __launch_bounds__(GG_LWARPS * 32) k_lists(GgArgs a) {
  using L = GgLayout<NN>;
  constexpr int LD = L::LD;
  GG_SMEM(double, lists);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t grp = (int64_t)blockIdx.x * GG_LWARPS + warp;
  if (grp >= a.ngroup) return;  // (whole warp)
  int32_t w[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) w[h] = reinterpret_cast<const int32_t *>(a.desc + grp * GG_MPW + 2 * h + (lane >> 4))[lane & 15];
  // word k of record g: held by lane (g & 1) * 16 + k in w[g >> 1]
  const bool isU = gg_shfl_i(w[0], 3) != 0;
  int len[GG_MPW], cnt[GG_MPW], off[GG_MPW];
  int maxcnt = 0, total = 0;
#pragma unroll
  for (int g = 0; g < GG_MPW; ++g) {
    len[g] = gg_shfl_i(w[g >> 1], (g & 1) * 16 + 4);
    cnt[g] = gg_shfl_i(w[g >> 1], (g & 1) * 16 + 5);
    off[g] = total;
    total += len[g];
    maxcnt = cnt[g] > maxcnt ? cnt[g] : maxcnt;
  }
  const int64_t p0 = (int64_t)(((uint64_t)(uint32_t)gg_shfl_i(w[0], 1) << 32) | (uint32_t)gg_shfl_i(w[0], 0));
  // first adjacency entry of record g (only read when it has more than ten incident cells)
  auto a0_of = [&](int g) { return a.adjptr[gg_shfl_i(w[g >> 1], (g & 1) * 16 + 2)]; };
  if (a.flags & GM_RESIDUAL) {
    // lane = 8 g + t: cell t of record g; sums over 8-lane groups; further cells are added one by one
    const int g = lane >> 3, t = lane & 7;
    const int32_t ew0 = gg_shfl_i(w[0], (g & 1) * 16 + 6 + t), ew1 = gg_shfl_i(w[1], (g & 1) * 16 + 6 + t);  // (every lane contributes both words)
    const int32_t jw0 = gg_shfl_i(w[0], (g & 1) * 16 + 2), jw1 = gg_shfl_i(w[1], (g & 1) * 16 + 2);
    const int32_t ew = g < 2 ? ew0 : ew1, j = g < 2 ? jw0 : jw1;
    int cg = cnt[0];
#pragma unroll
    for (int q = 1; q < GG_MPW; ++q) cg = g == q ? cnt[q] : cg;
    double v = t < cg ? a.cr[gg_seg_to_adj<NN>((uint32_t)ew)] : 0.0;
#pragma unroll
    for (int m = 1; m < 8; m <<= 1) v += gg_shfl_xor(v, m);
#pragma unroll
    for (int q = 0; q < GG_MPW; ++q) {
      if (cnt[q] > 8) {  // (uniform)
        const int64_t a0 = a0_of(q);
        for (int tt = 8; tt < cnt[q]; ++tt) {
          const int64_t e = tt < 10 ? gg_seg_to_adj<NN>((uint32_t)gg_shfl_i(w[q >> 1], (q & 1) * 16 + 6 + tt)) : (int64_t)a.adj[a0 + tt];
          if (lane == 8 * q) v += a.cr[e];
        }
      }
    }
    if (t == 0 && j >= 0) a.b[j] = v;
  }
  if (!(a.flags & GM_JACOBIAN)) return;
  double *Lw = lists + warp * (GG_MPW * a.lstride);
  for (int t = lane; t < total; t += 32) Lw[t] = 0.0;
  gg_syncwarp();
  const bool slot = lane < (isU ? LD : 3 * NN);
  for (int d = 0; d < maxcnt; d += 2) {
    uint16_t pos[GG_MPW][2];
    double v[GG_MPW][2];
#pragma unroll
    for (int g = 0; g < GG_MPW; ++g) {
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int tt = d + t;
        uint32_t sg = (uint32_t)gg_shfl_i(w[g >> 1], (g & 1) * 16 + (tt < 10 ? 6 + tt : 0));
        if (tt >= 10 && tt < cnt[g]) sg = gg_seg<NN>(a.adj[a0_of(g) + tt]);  // (uniform branch)
        pos[g][t] = 0xFFFF;
        v[g][t] = 0;
        if (tt < cnt[g] && slot) {
          pos[g][t] = a.grank[sg + lane];
          v[g][t] = a.cm[sg + lane];
        }
      }
    }
#pragma unroll
    for (int t = 0; t < 2; ++t) {
#pragma unroll
      for (int g = 0; g < GG_MPW; ++g)
        if (pos[g][t] != 0xFFFF) Lw[off[g] + pos[g][t]] += v[g][t];
      gg_syncwarp();
    }
  }
  double *out = a.nz + p0;
  for (int t = lane; t < total; t += 32) out[t] = Lw[t];
}
