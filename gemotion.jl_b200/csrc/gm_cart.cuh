// gm_cart.cuh -- structured numeric path for Cartesian quad meshes: owner-computes row gather.
//
// Replaces the same reference interface as gm_scatter.cuh (jacobian!/residual!/
// residual_and_jacobian! of the operator built at /root/reference/src/Solver.jl:157, integrands
// Solver.jl:111-152, arithmetic SURVEY.md A.4) but never read-modify-writes nzval in HBM:
//
//  * A CTA owns a strip of TH=8 cell rows and sweeps it column by column in x.  The value lists
//    (CSC columns / CSR rows) of the P2 nodes of the current column are accumulated in SHARED
//    memory; a list is complete when the sweep has passed its node, and is then streamed to HBM
//    exactly once, contiguously.  Cells of the strip's upper neighbour row (and of the column left
//    of an x-chunk) are evaluated as halo, only for the lists this CTA owns, so no entry of any
//    cell matrix is ever computed twice and no inter-CTA communication or atomics are needed.
//  * Inside a column, a warp evaluates one third of a cell matrix: lane = (major node, minor
//    node) pair.  The pair's reference-element products (phi_i phi_j, d phi_i d phi_j, ... times
//    w|J|) are loop-invariant on a Cartesian mesh and live in REGISTERS; the per-cell state
//    (mu, grad u, grad T, S, u.grad phi) is broadcast from shared memory.  (Measured on B200:
//    DFMA with one operand streamed from the constant bank or shared memory tops out at 2.8-5.6
//    TDFMA/s, with register operands at 15.5 - profiles/r01_ubench_dfma_operand_source.txt.)
//  * The cell->nnz map is table driven (rank classes): cells whose 30x30 rank table is identical
//    share one 900-byte class table, so whatever numbering Gridap produced is honoured while the
//    map costs 2 bytes per cell instead of 837 x 8.
//
// Summation order differs from Gridap's (cells of a node are added left column first, even rows
// before odd rows); values agree with the oracle to ~1e-15 relative (tests/test_gpu_parity.py).
#pragma once
#include <cub/cub.cuh>

#include <cmath>

#include "gm_internal.h"
#include "gm_scatter.cuh"

#define CT_TH 8          // cell rows per strip
#define CT_XCH 64        // columns per x-chunk
#define CT_WARPS 12      // 4 cells x 3 thirds per phase
#define CT_THREADS (CT_WARPS * 32)
#define CT_MAXCLASS 4096

// shared-memory list slots (doubles).  Node blocks: [list0][list1][list2][3 residual slots]
#define CT_VTX_L1 88
#define CT_VTX_L2 176
#define CT_VTX_RES 252
#define CT_VTX_BLK 256
#define CT_EDG_L1 52
#define CT_EDG_L2 104
#define CT_EDG_RES 150
#define CT_EDG_BLK 154
#define CT_CEN_L1 30
#define CT_CEN_L2 60
#define CT_CEN_RES 88
#define CT_CEN_BLK 92
#define CT_P_RES 54
#define CT_P_BLK 58
#define CT_EVEN_STRIDE (CT_VTX_BLK + CT_EDG_BLK)                  // 410 per cell row
#define CT_EVEN_SZ (CT_TH * CT_EVEN_STRIDE + CT_VTX_BLK)          // 3536
#define CT_ODD_STRIDE (CT_EDG_BLK + CT_CEN_BLK)                   // 246
#define CT_ODD_P0 (CT_TH * CT_ODD_STRIDE + CT_EDG_BLK)            // 2122
#define CT_ODD_SZ (CT_ODD_P0 + CT_TH * CT_P_BLK)                  // 2586
#define CT_NCELL (CT_TH + 1)                                      // strip cells + halo
#define CT_UNI 16                                                 // per (cell,q) uniform state
#define CT_NOD 6                                                  // per (cell,q,node) state

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

struct gm_cart {
  int nclass = 0;
  uint16_t *cls = nullptr;    // [ncells]
  uint8_t *ctab = nullptr;    // [nclass][900]   rank[minor][major]
  double *visc = nullptr;     // [ncells][4][2]  (mu, kappa)
  CartPairTab *d_tab = nullptr;
  CartCell *d_cell = nullptr;
  gm_params last_prm;
  bool tab_valid = false;
};

struct CartArgs {
  int nx, ny;
  const int32_t *gdofs;
  const double *dvals;
  const int64_t *ptr;
  const uint16_t *cls;
  const uint8_t *ctab;
  const double *visc;
  const CartPairTab *tab;
  const CartCell *cc;
  const double *x;
  double *nz;
  double *b;
  unsigned flags;
  double Pr, RaPr, gx, gy, reg, n, js;
};

// ------------------------------------------------------------------------------------------
// set-up kernels: rank classes + verification of the structural assumptions
// ------------------------------------------------------------------------------------------
__global__ void k_ct_hash(int64_t ncells, const uint16_t *__restrict__ rank, uint64_t *__restrict__ hash,
                          int32_t *__restrict__ cellid) {
  int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  const uint16_t *r = rank + c * 900;
  uint64_t h = 0x9E3779B97F4A7C15ull;
  for (int k = 0; k < 900; ++k) {
    h ^= (uint64_t)r[k] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h *= 0xBF58476D1CE4E5B9ull;
    h ^= h >> 29;
  }
  hash[c] = h;
  cellid[c] = (int32_t)c;
}

__global__ void k_ct_heads(int64_t n, const uint64_t *__restrict__ hs, int32_t *__restrict__ head) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) head[i] = (i == 0 || hs[i] != hs[i - 1]) ? 1 : 0;
}

__global__ void k_ct_assign(int64_t n, const int32_t *__restrict__ head, const int32_t *__restrict__ cid_incl,
                            const int32_t *__restrict__ cell_sorted, uint16_t *__restrict__ cls, int32_t *__restrict__ rep,
                            int maxclass) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int id = cid_incl[i] - 1;
  if (id >= maxclass) return;
  cls[cell_sorted[i]] = (uint16_t)id;
  if (head[i]) rep[id] = cell_sorted[i];
}

__global__ void k_ct_table(int nclass, const int32_t *__restrict__ rep, const uint16_t *__restrict__ rank,
                           uint8_t *__restrict__ ctab, int *__restrict__ err) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nclass * 900) return;
  const int k = t / 900, e = t - k * 900;
  const uint16_t r = rank[(int64_t)rep[k] * 900 + e];
  if (r != 0xFFFF && r >= 255) atomicExch(err, 2);
  ctab[t] = r == 0xFFFF ? 0xFF : (uint8_t)r;
}

// slot capacity of the list of local dof l (node type decides)
__device__ __forceinline__ int ct_cap(int l) {
  if (l >= 18 && l < 21) return 18;
  const bool isT = l >= 21;
  const int node = isT ? l - 21 : l % 9;
  if (node < 4) return isT ? 75 : 87;
  if (node < 8) return isT ? 45 : 51;
  return isT ? 27 : 30;
}

__global__ void k_ct_verify(int nx, int ny, const int32_t *__restrict__ g, const int64_t *__restrict__ ptr,
                            const uint16_t *__restrict__ rank, const uint16_t *__restrict__ cls,
                            const uint8_t *__restrict__ ctab, int *__restrict__ err) {
  int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= (int64_t)nx * ny) return;
  const int x = (int)(c % nx), y = (int)(c / nx);
  const int32_t *gc = g + c * 30;
  // (1) rank table equals the class table
  const uint16_t *r = rank + c * 900;
  const uint8_t *t = ctab + (int64_t)cls[c] * 900;
  for (int k = 0; k < 900; ++k) {
    const uint16_t a = r[k];
    if ((a == 0xFFFF ? 0xFF : (uint8_t)a) != t[k]) { atomicExch(err, 3); return; }
  }
  // (2) list lengths fit the slots
  for (int l = 0; l < 30; ++l) {
    const int32_t d = gc[l];
    if (d > 0 && ptr[d] - ptr[d - 1] > ct_cap(l)) { atomicExch(err, 4); return; }
  }
  // (3) conforming tensor-order sharing with the right and upper neighbours, U (both comps) and T
  const int offs[3] = {0, 9, 21};
  if (x + 1 < nx) {
    const int32_t *gr = gc + 30;
    for (int f = 0; f < 3; ++f)
      if (gc[offs[f] + 1] != gr[offs[f] + 0] || gc[offs[f] + 3] != gr[offs[f] + 2] || gc[offs[f] + 7] != gr[offs[f] + 6]) {
        atomicExch(err, 5); return;
      }
  }
  if (y + 1 < ny) {
    const int32_t *gu = gc + (int64_t)nx * 30;
    for (int f = 0; f < 3; ++f)
      if (gc[offs[f] + 2] != gu[offs[f] + 0] || gc[offs[f] + 3] != gu[offs[f] + 1] || gc[offs[f] + 5] != gu[offs[f] + 4]) {
        atomicExch(err, 5); return;
      }
  }
}

// ------------------------------------------------------------------------------------------
// viscosity pre-pass (n != 1): mu = gamma^((n-1)/2), kappa = ((n-1)/2) gamma^((n-3)/2) 4
// (Solver.jl:113-120).  One thread per (cell, quadrature point); 64 B/cell each way.
// ------------------------------------------------------------------------------------------
__global__ void k_ct_visc(int64_t ncells, const int32_t *__restrict__ g, const double *__restrict__ dvals,
                          const double *__restrict__ x, const CartCell *__restrict__ cc, double reg, double n,
                          double *__restrict__ visc) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= ncells * 4) return;
  const int64_t c = t >> 2;
  const int q = (int)(t & 3);
  double G00 = 0, G01 = 0, G10 = 0, G11 = 0;
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const int32_t d0 = g[c * 30 + i], d1 = g[c * 30 + 9 + i];
    const double ux = d0 > 0 ? x[d0 - 1] : dvals[-d0 - 1];
    const double uy = d1 > 0 ? x[d1 - 1] : dvals[-d1 - 1];
    const double dx = cc->sdx[q][i], dy = cc->sdy[q][i];
    G00 += dx * ux; G01 += dx * uy; G10 += dy * ux; G11 += dy * uy;
  }
  const double D01 = 0.5 * (G01 + G10);
  const double gam = reg + 2.0 * (G00 * G00 + 2.0 * D01 * D01 + G11 * G11);
  const double mu = pow(gam, 0.5 * (n - 1.0));
  visc[t * 2] = mu;
  visc[t * 2 + 1] = 0.5 * (n - 1.0) * pow(gam, 0.5 * (n - 3.0)) * 4.0;
}

// ------------------------------------------------------------------------------------------
// main kernel
// ------------------------------------------------------------------------------------------
__device__ __constant__ int8_t c_ct_node[3][3] = {{0, 4, 1}, {6, 8, 7}, {2, 5, 3}};  // [iy][ix] -> local node

struct CtSmem {
  double even[2][CT_EVEN_SZ];
  double odd[CT_ODD_SZ];
  double uni[CT_NCELL][4][CT_UNI];
  double nod[CT_NCELL][4][9][CT_NOD];
  double val[CT_NCELL][30];
  int32_t gd[CT_NCELL][30];
  uint16_t cls[CT_NCELL];
};

// uniform state slots
#define U_MU 0
#define U_G00 1
#define U_G01 2
#define U_G10 3
#define U_G11 4
#define U_DT0 5
#define U_DT1 6
#define U_KAP 7
#define U_CONV0 8
#define U_CONV1 9
#define U_P 10
#define U_T 11
#define U_UDT 12
#define U_DIVU 13
#define U_U0 14
#define U_U1 15

template <bool CSR, bool NEWT>
__global__ void __launch_bounds__(CT_THREADS, 1) k_cart(CartArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  CtSmem &s = *reinterpret_cast<CtSmem *>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int y0 = blockIdx.y * CT_TH;
  const int th = min(CT_TH, a.ny - y0);
  const bool has_halo = y0 + th < a.ny;
  const int ncol_cells = th + (has_halo ? 1 : 0);
  const int x0 = blockIdx.x * CT_XCH, x1 = min(a.nx, x0 + CT_XCH);
  const int jlo = y0 == 0 ? 0 : 1;
  const bool jac = (a.flags & GM_JACOBIAN) != 0, res = (a.flags & GM_RESIDUAL) != 0;
  const CartCell &cc = *a.cc;

  for (int k = tid; k < 2 * CT_EVEN_SZ + CT_ODD_SZ; k += CT_THREADS) s.even[0][k] = 0.0;  // even[2] and odd are contiguous

  // ---- lane role: task type (bottom/mid/top third), cell slot, (major node, minor node)
  const int tt = warp % 3, slot = warp / 3;
  const int Mi = lane / 9, mi = lane - Mi * 9;  // major index within the third (= ix), minor node
  const bool pair_lane = lane < 27;
  const int ixM = pair_lane ? Mi : 0, iyM = tt;
  const int M = c_ct_node[iyM][ixM];
  const int it = CSR ? M : mi, jt = CSR ? mi : M;
  // register tables of the pair
  double TA[4], TB[4], TC[4], TD[4], TM[4], TP[4], TK, TUT0, TUT1;
  {
    const CartPairTab &t = a.tab[it * 9 + jt];
#pragma unroll
    for (int q = 0; q < 4; ++q) { TA[q] = t.TA[q]; TB[q] = t.TB[q]; TC[q] = t.TC[q]; TD[q] = t.TD[q]; TM[q] = t.TM[q]; TP[q] = t.TP[q]; }
    TK = t.TK; TUT0 = t.TUT0; TUT1 = t.TUT1;
  }
  // local dof ids of the three major lists and three minor dofs
  const int Ld[3] = {M, 9 + M, 21 + M};
  const int md[3] = {mi, 9 + mi, 21 + mi};
  // node-type dependent list offsets of the major node
  const bool mvtx = !(ixM & 1) && !(iyM & 1), mcen = (ixM & 1) && (iyM & 1);
  const int lo1 = mvtx ? CT_VTX_L1 : (mcen ? CT_CEN_L1 : CT_EDG_L1);
  const int lo2 = mvtx ? CT_VTX_L2 : (mcen ? CT_CEN_L2 : CT_EDG_L2);
  // extras: lanes mi<6 of every major: the (U major list a, P minor p) entry [constant]
  const int ex_a = mi / 3, ex_p = mi % 3;
  const double ex_val = (pair_lane && mi < 6) ? (CSR ? cc.UP[ex_a * 9 + M][ex_p] : -cc.UP[ex_a * 9 + M][ex_p]) : 0.0;
  // extras of the mid third: P major lists, two (P list, U minor) entries per lane
  double pv[2];
  int pl[2], pd[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int e = 2 * lane + k;  // 0..53 for lane<27
    pl[k] = (e / 18) % 3;
    pd[k] = e % 18;
    const double up = cc.UP[pd[k]][pl[k]];
    pv[k] = CSR ? -up : up;  // PU[p,d] = -UP[d,p]
  }
  // cached ranks of the current class
  int cur_cls = -1;
  uint8_t rk[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  uint8_t rk_ex = 0xFF, rk_p[2] = {0xFF, 0xFF};

  int eL = 0;  // index of the even buffer holding the LEFT line (right line = 1-eL)
  __syncthreads();

  const int xs = x0 > 0 ? x0 - 1 : x0;
  for (int x = xs; x < x1; ++x) {
    const bool pre = x < x0;  // halo column left of the chunk: only right-line contributions count
    // =============================== state phase ===============================
    // S1: dof ids + values of the column's cells
    for (int k = tid; k < ncol_cells * 30; k += CT_THREADS) {
      const int cy = k / 30, l = k - cy * 30;
      const int64_t c = (int64_t)(y0 + cy) * a.nx + x;
      const int32_t d = a.gdofs[c * 30 + l];
      s.gd[cy][l] = d;
      s.val[cy][l] = d > 0 ? a.x[d - 1] : a.dvals[-d - 1];
      if (l == 0) s.cls[cy] = a.cls[c];
    }
    __syncthreads();
    // S2: fields at the quadrature points: items (cell, q, k<10)
    for (int k = tid; k < ncol_cells * 40; k += CT_THREADS) {
      const int cy = k / 40, r = k - cy * 40, q = r / 10, f = r - q * 10;
      double v = 0;
      if (f == 9) {
#pragma unroll
        for (int m = 0; m < 3; ++m) v += cc.psi[q][m] * s.val[cy][18 + m];
        s.uni[cy][q][U_P] = v;
      } else {
        // f: 0 u0, 1 u1, 2 G00, 3 G01, 4 G10, 5 G11, 6 T, 7 dT0, 8 dT1
        const double *tab = (f < 2 || f == 6) ? cc.sphi[q] : ((f == 2 || f == 3 || f == 7) ? cc.sdx[q] : cc.sdy[q]);
        const int vo = (f == 0 || f == 2 || f == 4) ? 0 : ((f == 1 || f == 3 || f == 5) ? 9 : 21);
#pragma unroll
        for (int i = 0; i < 9; ++i) v += tab[i] * s.val[cy][vo + i];
        const int dst = f == 0 ? U_U0 : f == 1 ? U_U1 : f == 2 ? U_G00 : f == 3 ? U_G01 : f == 4 ? U_G10 : f == 5 ? U_G11
                      : f == 6 ? U_T : f == 7 ? U_DT0 : U_DT1;
        s.uni[cy][q][dst] = v;
      }
    }
    __syncthreads();
    // S3: derived uniform state: items (cell, q)
    for (int k = tid; k < ncol_cells * 4; k += CT_THREADS) {
      const int cy = k >> 2, q = k & 3;
      double *u = s.uni[cy][q];
      double mu = 1.0, kap = 0.0;
      if (!NEWT) {
        const int64_t c = (int64_t)(y0 + cy) * a.nx + x;
        mu = a.visc[(c * 4 + q) * 2];
        kap = a.visc[(c * 4 + q) * 2 + 1];
      }
      u[U_MU] = mu;
      u[U_KAP] = kap;
      u[U_CONV0] = u[U_U0] * u[U_G00] + u[U_U1] * u[U_G10];
      u[U_CONV1] = u[U_U0] * u[U_G01] + u[U_U1] * u[U_G11];
      u[U_UDT] = u[U_U0] * u[U_DT0] + u[U_U1] * u[U_DT1];
      u[U_DIVU] = u[U_G00] + u[U_G11];
    }
    __syncthreads();
    // S4: per-node state: items (cell, q, node): S0,S1 raw; S'0,S'1 = js w 2Pr kappa S; c = u.grad phi
    for (int k = tid; k < ncol_cells * 36; k += CT_THREADS) {
      const int cy = k / 36, r = k - cy * 36, q = r / 9, i = r - q * 9;
      const double *u = s.uni[cy][q];
      const double dx = cc.sdx[q][i], dy = cc.sdy[q][i];
      const double D01 = 0.5 * (u[U_G01] + u[U_G10]);
      const double S0 = u[U_G00] * dx + D01 * dy;
      const double S1 = D01 * dx + u[U_G11] * dy;
      const double sc = a.js * cc.w[q] * 2.0 * a.Pr * u[U_KAP];
      double *o = s.nod[cy][q][i];
      o[0] = S0; o[1] = S1; o[2] = sc * S0; o[3] = sc * S1;
      o[4] = u[U_U0] * dx + u[U_U1] * dy;
    }
    __syncthreads();
    // S5: residual gather: items (ixn, j', comp); sums the <=2 cells of this column sharing the node
    if (res) {
      for (int k = tid; k < 3 * 17 * 3; k += CT_THREADS) {
        const int ixn = k / 51, r = k - ixn * 51, jp = r / 3, comp = r - jp * 3;
        if (jp < jlo || jp > 2 * th || (pre && ixn != 2)) continue;
        double acc = 0;
        // contributing cells: jp odd -> cell (jp-1)/2 with iy=1; jp even -> cell jp/2 (iy=0) and jp/2-1 (iy=2)
        for (int side = 0; side < 2; ++side) {
          int cy, iy;
          if (jp & 1) { if (side) break; cy = jp >> 1; iy = 1; }
          else { cy = (jp >> 1) - side; iy = side ? 2 : 0; }
          if (cy < 0 || cy >= ncol_cells) continue;
          if (cy == th && iy != 0) continue;  // halo cell: bottom nodes only
          const int i = c_ct_node[iy][ixn];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const double *u = s.uni[cy][q];
            const double *nd = s.nod[cy][q][i];
            const double ph = cc.sphi[q][i], dx = cc.sdx[q][i], dy = cc.sdy[q][i];
            double v;
            if (comp == 0) v = 2.0 * a.Pr * u[U_MU] * nd[0] + ph * u[U_CONV0] - dx * u[U_P] - a.RaPr * u[U_T] * a.gx * ph;
            else if (comp == 1) v = 2.0 * a.Pr * u[U_MU] * nd[1] + ph * u[U_CONV1] - dy * u[U_P] - a.RaPr * u[U_T] * a.gy * ph;
            else v = u[U_DT0] * dx + u[U_DT1] * dy + u[U_UDT] * ph;
            acc += cc.w[q] * v;
          }
        }
        double *line = (ixn == 1) ? s.odd : s.even[ixn == 0 ? eL : 1 - eL];
        const int blk = (ixn == 1) ? (jp >> 1) * CT_ODD_STRIDE + ((jp & 1) ? CT_EDG_BLK + CT_CEN_RES : CT_EDG_RES)
                                   : (jp >> 1) * CT_EVEN_STRIDE + ((jp & 1) ? CT_VTX_BLK + CT_EDG_RES : CT_VTX_RES);
        line[blk + comp] += acc;
      }
      if (!pre)
        for (int k = tid; k < th * 3; k += CT_THREADS) {
          const int cy = k / 3, p = k - cy * 3;
          double acc = 0;
#pragma unroll
          for (int q = 0; q < 4; ++q) acc += cc.wpsi[q][p] * s.uni[cy][q][U_DIVU];
          s.odd[CT_ODD_P0 + cy * CT_P_BLK + CT_P_RES + p] = acc;
        }
    }
    // =============================== matrix phases ===============================
    if (jac) {
      for (int ph = 0; ph < 2; ++ph) {
        // tasks of this warp in this phase: its regular cell, and (bottom third, slot 0, even phase) the halo cell
        for (int rep = 0; rep < 2; ++rep) {
          int cy;
          if (rep == 0) cy = 2 * slot + ph;
          else { if (!(ph == 0 && tt == 0 && slot == 0 && has_halo)) break; cy = th; }
          if (rep == 0 && cy >= th) continue;
          const int jp = 2 * cy + iyM;  // row index of the major node inside the strip
          if (jp < jlo || jp > 2 * th) continue;          // list not owned by this strip
          // class / ranks (warp uniform)
          const int cl = s.cls[cy];
          if (cl != cur_cls) {
            cur_cls = cl;
            const uint8_t *t = a.ctab + (int64_t)cl * 900;
            if (pair_lane) {
#pragma unroll
              for (int kk = 0; kk < 3; ++kk)
#pragma unroll
                for (int ll = 0; ll < 3; ++ll) rk[kk * 3 + ll] = t[md[ll] * 30 + Ld[kk]];
              rk_ex = mi < 6 ? t[(18 + ex_p) * 30 + Ld[ex_a]] : 0xFF;
              if (tt == 1) {
                rk_p[0] = t[pd[0] * 30 + 18 + pl[0]];
                rk_p[1] = t[pd[1] * 30 + 18 + pl[1]];
              }
            }
          }
          if (!pair_lane) continue;
          if (pre && ixM != 2) continue;  // halo column: only the right line is owned
          // ---- 9 values of the pair
          double v00 = 0, v01 = 0, v10 = 0, v11 = 0, e = 0, tu0 = 0, tu1 = 0;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const double2 *u2 = reinterpret_cast<const double2 *>(s.uni[cy][q]);
            const double2 a0 = u2[0], a1 = u2[1], a2 = u2[2], a3 = u2[3];  // (mu,G00) (G01,G10) (G11,dT0) (dT1,kap)
            const double2 si = *reinterpret_cast<const double2 *>(&s.nod[cy][q][it][0]);
            const double2 sj = *reinterpret_cast<const double2 *>(&s.nod[cy][q][jt][2]);
            const double cj = s.nod[cy][q][jt][4];
            const double mu = a0.x;
            v00 = fma(mu, TA[q], v00); v11 = fma(mu, TB[q], v11); v01 = fma(mu, TC[q], v01); v10 = fma(mu, TD[q], v10);
            if (!NEWT) {
              v00 = fma(si.x, sj.x, v00); v01 = fma(si.x, sj.y, v01); v10 = fma(si.y, sj.x, v10); v11 = fma(si.y, sj.y, v11);
            }
            v00 = fma(TM[q], a0.y, v00);  // G00
            v01 = fma(TM[q], a1.y, v01);  // (a=0,b=1): G[1][0]
            v10 = fma(TM[q], a1.x, v10);  // (a=1,b=0): G[0][1]
            v11 = fma(TM[q], a2.x, v11);  // G11
            e = fma(TP[q], cj, e);
            tu0 = fma(TM[q], a2.y, tu0);
            tu1 = fma(TM[q], a3.x, tu1);
          }
          v00 += e; v11 += e;
          const double tt_v = TK + e;
          // ---- accumulate into the lists of the major node
          double *line = (ixM == 1) ? s.odd : s.even[ixM == 0 ? eL : 1 - eL];
          const int row = cy + (iyM >> 1);
          double *blk = line + ((ixM == 1) ? row * CT_ODD_STRIDE + ((iyM & 1) ? CT_EDG_BLK : 0)
                                           : row * CT_EVEN_STRIDE + ((iyM & 1) ? CT_VTX_BLK : 0));
          // value of (major list k, minor dof l) in terms of test/trial
          //   CSR: major=test (it=M): [k][l] = J[(M,k),(m,l)] ; CSC: major=trial: [k][l] = J[(m,l),(M,k)]
          double w9[9];
          if (CSR) {
            w9[0] = v00; w9[1] = v01; w9[2] = TUT0; w9[3] = v10; w9[4] = v11; w9[5] = TUT1; w9[6] = tu0; w9[7] = tu1; w9[8] = tt_v;
          } else {
            w9[0] = v00; w9[1] = v10; w9[2] = tu0; w9[3] = v01; w9[4] = v11; w9[5] = tu1; w9[6] = TUT0; w9[7] = TUT1; w9[8] = tt_v;
          }
#pragma unroll
          for (int kk = 0; kk < 3; ++kk) {
            double *lst = blk + (kk == 0 ? 0 : (kk == 1 ? lo1 : lo2));
#pragma unroll
            for (int ll = 0; ll < 3; ++ll) {
              const uint8_t r = rk[kk * 3 + ll];
              if (r != 0xFF) lst[r] += w9[kk * 3 + ll];
            }
          }
          if (rk_ex != 0xFF) (blk + (ex_a == 0 ? 0 : lo1))[rk_ex] += ex_val;
          if (tt == 1 && cy < th && !pre) {  // P major lists of this cell (complete within the cell)
            double *pb = s.odd + CT_ODD_P0 + cy * CT_P_BLK;
            if (rk_p[0] != 0xFF) pb[pl[0] * 18 + rk_p[0]] += pv[0];
            if (rk_p[1] != 0xFF) pb[pl[1] * 18 + rk_p[1]] += pv[1];
          }
        }
        __syncthreads();
      }
    } else {
      __syncthreads();
    }
    // =============================== write-out ===============================
    if (!pre) {
      // lists: line 0 (left, even buffer eL), line 1 (mid, odd buffer), line 2 (right) only at the last column
      const int nlines = (x == a.nx - 1) ? 3 : 2;
      const int per_line = 17 * 3;
      const int ntask = nlines * per_line + th * 3;
      for (int tk = warp; tk < ntask; tk += CT_WARPS) {
        double *src;
        int cap;
        int32_t gdof;
        int resoff;  // slot holding the residual of this dof
        double *blk;
        if (tk < nlines * per_line) {
          const int ln = tk / per_line, r = tk - ln * per_line, jp = r / 3, k = r - jp * 3;
          if (jp < jlo || jp > 2 * th) continue;
          // owning cell / local node of the node (ln, jp)
          const int cy = jp == 2 * th ? th - 1 : (jp >> 1);
          const int iy = jp == 2 * th ? 2 : (jp & 1);
          const int node = c_ct_node[iy][ln];
          gdof = s.gd[cy][k == 0 ? node : (k == 1 ? 9 + node : 21 + node)];
          double *line = (ln == 1) ? s.odd : s.even[ln == 0 ? eL : 1 - eL];
          const bool odd_j = jp & 1;
          if (ln == 1) {
            blk = line + (jp >> 1) * CT_ODD_STRIDE + (odd_j ? CT_EDG_BLK : 0);
            src = blk + (k == 0 ? 0 : (odd_j ? (k == 1 ? CT_CEN_L1 : CT_CEN_L2) : (k == 1 ? CT_EDG_L1 : CT_EDG_L2)));
            cap = odd_j ? (k == 2 ? 28 : 30) : (k == 2 ? 46 : 52);
            resoff = odd_j ? CT_CEN_RES : CT_EDG_RES;
          } else {
            blk = line + (jp >> 1) * CT_EVEN_STRIDE + (odd_j ? CT_VTX_BLK : 0);
            src = blk + (k == 0 ? 0 : (odd_j ? (k == 1 ? CT_EDG_L1 : CT_EDG_L2) : (k == 1 ? CT_VTX_L1 : CT_VTX_L2)));
            cap = odd_j ? (k == 2 ? 46 : 52) : (k == 2 ? 76 : 88);
            resoff = odd_j ? CT_EDG_RES : CT_VTX_RES;
          }
          resoff += k;
        } else {
          const int r = tk - nlines * per_line, cy = r / 3, p = r - cy * 3;
          gdof = s.gd[cy][18 + p];
          blk = s.odd + CT_ODD_P0 + cy * CT_P_BLK;
          src = blk + p * 18;
          cap = 18;
          resoff = CT_P_RES + p;
        }
        if (gdof > 0) {
          if (jac) {
            const int64_t base = a.ptr[gdof - 1];
            const int len = (int)(a.ptr[gdof] - base);
            for (int k2 = lane; k2 < len; k2 += 32) a.nz[base + k2] = src[k2];
          }
          if (res && lane == 0) a.b[gdof - 1] = blk[resoff];
        }
        if (jac)
          for (int k2 = lane; k2 < cap; k2 += 32) src[k2] = 0.0;
        if (lane == 0) blk[resoff] = 0.0;
      }
    }
    __syncthreads();
    eL = 1 - eL;  // the right line becomes the next column's left line
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static void gm_cart_destroy(gm_handle *h) {
  gm_cart *c = (gm_cart *)h->cart;
  if (!c) return;
  cudaFree(c->cls); cudaFree(c->ctab); cudaFree(c->visc); cudaFree(c->d_tab); cudaFree(c->d_cell);
  delete c;
  h->cart = nullptr;
}

static int gm_cart_create(gm_handle *h) {
  if (!h->d.cartesian || h->d.cell_type != GM_QUAD) { gm_set_error("structured path needs Cartesian quads"); return GM_ERR_ARG; }
  gm_cart *c = new (std::nothrow) gm_cart();
  if (!c) return GM_ERR_OOM;
  h->cart = c;
  h->active_path = GM_PATH_CART;  // confirmed (or revoked) by gm_cart_after_pattern
  return GM_OK;
}

// Builds rank classes from the per-cell rank table and verifies the structural assumptions.
// On any mismatch the handle silently stays on the scatter path (GM_PATH_AUTO) or fails (GM_PATH_CART).
static int gm_cart_after_pattern(gm_handle *h) {
  gm_cart *c = (gm_cart *)h->cart;
  const int64_t nc = h->ncells;
  uint64_t *hash = nullptr, *hash2 = nullptr;
  int32_t *cell = nullptr, *cell2 = nullptr, *head = nullptr, *cid = nullptr, *rep = nullptr;
  int *err = nullptr;
  void *tmp = nullptr;
  size_t tmpb = 0, t2 = 0;
  int rc = GM_OK, herr = 0, nclass = 0;
  const char *why = "";
#define CB(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { gm_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); rc = e__ == cudaErrorMemoryAllocation ? GM_ERR_OOM : GM_ERR_CUDA; goto done; } } while (0)
  CB(cudaMalloc(&hash, 8 * nc)); CB(cudaMalloc(&hash2, 8 * nc));
  CB(cudaMalloc(&cell, 4 * nc)); CB(cudaMalloc(&cell2, 4 * nc));
  CB(cudaMalloc(&head, 4 * nc)); CB(cudaMalloc(&cid, 4 * nc));
  CB(cudaMalloc(&rep, 4 * CT_MAXCLASS)); CB(cudaMalloc(&err, 4));
  CB(cudaMemsetAsync(err, 0, 4, h->stream));
  k_ct_hash<<<gm_blocks(nc, 128), 128, 0, h->stream>>>(nc, h->rank, hash, cell);
  cub::DeviceRadixSort::SortPairs(nullptr, tmpb, hash, hash2, cell, cell2, nc, 0, 64, h->stream);
  cub::DeviceScan::InclusiveSum(nullptr, t2, head, cid, nc, h->stream);
  if (t2 > tmpb) tmpb = t2;
  CB(cudaMalloc(&tmp, tmpb));
  CB(cub::DeviceRadixSort::SortPairs(tmp, tmpb, hash, hash2, cell, cell2, nc, 0, 64, h->stream));
  k_ct_heads<<<gm_blocks(nc, 256), 256, 0, h->stream>>>(nc, hash2, head);
  CB(cub::DeviceScan::InclusiveSum(tmp, tmpb, head, cid, nc, h->stream));
  CB(cudaMemcpyAsync(&nclass, cid + nc - 1, 4, cudaMemcpyDeviceToHost, h->stream));
  CB(cudaStreamSynchronize(h->stream));
  if (nclass > CT_MAXCLASS) { why = "too many distinct cell->nnz rank tables (irregular numbering)"; herr = 1; goto decided; }
  CB(cudaMalloc(&c->cls, 2 * nc));
  CB(cudaMalloc(&c->ctab, (size_t)nclass * 900));
  k_ct_assign<<<gm_blocks(nc, 256), 256, 0, h->stream>>>(nc, head, cid, cell2, c->cls, rep, CT_MAXCLASS);
  k_ct_table<<<gm_blocks((int64_t)nclass * 900, 256), 256, 0, h->stream>>>(nclass, rep, h->rank, c->ctab, err);
  k_ct_verify<<<gm_blocks(nc, 128), 128, 0, h->stream>>>(h->d.nx, h->d.ny, h->gdofs, h->ptr, h->rank, c->cls, c->ctab, err);
  CB(cudaMemcpyAsync(&herr, err, 4, cudaMemcpyDeviceToHost, h->stream));
  CB(cudaStreamSynchronize(h->stream));
  if (herr) why = herr == 2 ? "a list is longer than 254 entries" : herr == 3 ? "rank-class hash collision"
                : herr == 4 ? "a list exceeds its shared-memory slot" : "dof tables are not a conforming tensor-order Cartesian numbering";
decided:
  if (herr) {
    if (h->d.path == GM_PATH_CART) {
      gm_set_error("structured path unavailable: %s", why);
      rc = GM_ERR_ARG;
    } else {
      gm_cart_destroy(h);
      h->active_path = GM_PATH_SCATTER;
    }
    goto done;
  }
  c->nclass = nclass;
  h->device_bytes += 2 * nc + (int64_t)nclass * 900;
  CB(cudaMalloc(&c->visc, sizeof(double) * 8 * nc));
  CB(cudaMalloc(&c->d_tab, sizeof(CartPairTab) * 81));
  CB(cudaMalloc(&c->d_cell, sizeof(CartCell)));
  h->device_bytes += 64 * nc;
  // the per-cell rank table is no longer needed (the class table replaces it)
  cudaFree(h->rank);
  h->rank = nullptr;
  h->device_bytes -= (int64_t)sizeof(uint16_t) * nc * 900;
  CB(cudaFuncSetAttribute(k_cart<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CtSmem)));
  CB(cudaFuncSetAttribute(k_cart<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CtSmem)));
  CB(cudaFuncSetAttribute(k_cart<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CtSmem)));
  CB(cudaFuncSetAttribute(k_cart<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(CtSmem)));
done:
  cudaFree(hash); cudaFree(hash2); cudaFree(cell); cudaFree(cell2); cudaFree(head); cudaFree(cid); cudaFree(rep);
  cudaFree(err); cudaFree(tmp);
#undef CB
  return rc;
}

// pair / cell constant tables from the reference tables and the current parameters
static int gm_cart_tables(gm_handle *h) {
  gm_cart *c = (gm_cart *)h->cart;
  if (c->tab_valid && memcmp(&c->last_prm, &h->prm, sizeof(gm_params)) == 0) return GM_OK;
  const GmConst &k = h->hc;
  const gm_params &p = h->prm;
  static thread_local CartPairTab tab[81];
  static thread_local CartCell cell;
  const double detJ = k.hx * k.hy;
  for (int q = 0; q < 4; ++q) {
    cell.w[q] = k.w[q] * detJ;
    for (int i = 0; i < 9; ++i) {
      cell.sphi[q][i] = k.phi[q * 9 + i];
      cell.sdx[q][i] = k.dphi[(q * 9 + i) * 2] / k.hx;
      cell.sdy[q][i] = k.dphi[(q * 9 + i) * 2 + 1] / k.hy;
    }
    for (int m = 0; m < 3; ++m) { cell.wpsi[q][m] = cell.w[q] * k.psi[q * 3 + m]; cell.psi[q][m] = k.psi[q * 3 + m]; }
  }
  for (int a = 0; a < 2; ++a)
    for (int i = 0; i < 9; ++i)
      for (int m = 0; m < 3; ++m) {
        double v = 0;
        for (int q = 0; q < 4; ++q) v += cell.w[q] * (-(a == 0 ? cell.sdx[q][i] : cell.sdy[q][i]) * k.psi[q * 3 + m]);
        cell.UP[a * 9 + i][m] = p.jac_scaling * v;
      }
  for (int i = 0; i < 9; ++i)
    for (int j = 0; j < 9; ++j) {
      CartPairTab &t = tab[i * 9 + j];
      double kbar = 0, mbar = 0;
      for (int q = 0; q < 4; ++q) {
        const double w = p.jac_scaling * cell.w[q];
        const double xi = cell.sdx[q][i], yi = cell.sdy[q][i], xj = cell.sdx[q][j], yj = cell.sdy[q][j];
        const double K = xi * xj + yi * yj;
        t.TA[q] = w * p.Pr * (K + xi * xj);
        t.TB[q] = w * p.Pr * (K + yi * yj);
        t.TC[q] = w * p.Pr * (yi * xj);  // (a=0,b=1): d_b phi_i d_a phi_j
        t.TD[q] = w * p.Pr * (xi * yj);  // (a=1,b=0)
        t.TM[q] = w * cell.sphi[q][i] * cell.sphi[q][j];
        t.TP[q] = w * cell.sphi[q][i];
        kbar += w * K;
        mbar += w * cell.sphi[q][i] * cell.sphi[q][j];
      }
      t.TK = kbar;
      t.TUT0 = -p.Ra * p.Pr * p.gx * mbar;
      t.TUT1 = -p.Ra * p.Pr * p.gy * mbar;
    }
  GM_CUDA(cudaMemcpyAsync(c->d_tab, tab, sizeof tab, cudaMemcpyHostToDevice, h->stream));
  GM_CUDA(cudaMemcpyAsync(c->d_cell, &cell, sizeof cell, cudaMemcpyHostToDevice, h->stream));
  GM_CUDA(cudaStreamSynchronize(h->stream));  // the staging arrays are reused by the next call
  c->last_prm = h->prm;
  c->tab_valid = true;
  return GM_OK;
}

static int gm_cart_run(gm_handle *h, const double *d_x, double *d_nz, double *d_b, unsigned flags, int64_t *launches) {
  gm_cart *c = (gm_cart *)h->cart;
  int rc = gm_cart_tables(h);
  if (rc) return rc;
  const gm_params &p = h->prm;
  CartArgs a;
  a.nx = h->d.nx; a.ny = h->d.ny;
  a.gdofs = h->gdofs; a.dvals = h->dvals; a.ptr = h->ptr; a.cls = c->cls; a.ctab = c->ctab; a.visc = c->visc;
  a.tab = c->d_tab; a.cc = c->d_cell; a.x = d_x; a.nz = d_nz; a.b = d_b; a.flags = flags;
  a.Pr = p.Pr; a.RaPr = p.Ra * p.Pr; a.gx = p.gx; a.gy = p.gy; a.reg = p.reg; a.n = p.n; a.js = p.jac_scaling;
  const bool newt = p.n == 1.0;
  if (!newt) {
    k_ct_visc<<<gm_blocks(h->ncells * 4, 256), 256, 0, h->stream>>>(h->ncells, h->gdofs, h->dvals, d_x, c->d_cell, p.reg, p.n, c->visc);
    ++*launches;
  }
  dim3 grid((a.nx + CT_XCH - 1) / CT_XCH, (a.ny + CT_TH - 1) / CT_TH);
  const size_t sm = sizeof(CtSmem);
  const bool csr = h->d.layout == GM_LAYOUT_CSR;
  if (csr && newt) k_cart<true, true><<<grid, CT_THREADS, sm, h->stream>>>(a);
  else if (csr) k_cart<true, false><<<grid, CT_THREADS, sm, h->stream>>>(a);
  else if (newt) k_cart<false, true><<<grid, CT_THREADS, sm, h->stream>>>(a);
  else k_cart<false, false><<<grid, CT_THREADS, sm, h->stream>>>(a);
  ++*launches;
  GM_CUDA(cudaGetLastError());
  return GM_OK;
}
