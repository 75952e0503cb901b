// gm_cart_kernel.cuh -- the sweep kernel of the structured path (see gm_cart.cuh for the design).
//
// CTA = 3*TH/2 matrix warps + 2 aux warps, software-pipelined over the columns x of the strip:
//   matrix : two colour phases (even / odd cell rows) of column x; lane = (major node, minor node)
//            pair with its reference-element tables in registers; 9 values (+ constant P-block
//            entries) are accumulated into the shared-memory lists without branches (entries of
//            Dirichlet minors go to a trash element at the end of each slot); then the state of
//            column x+1 (fields at the quadrature points, S, u.grad(phi)) into the other state
//            buffer.  One CTA barrier per column.
//   aux    : concurrently: write-out of column x-1 -- one TMA bulk copy (cp.async.bulk shared->global) per
//            completed list; residual gather of column x; recycling (zeroing) of the drained
//            line buffers.  (Giving the aux warps the next column's state as well made them the
//            critical path: 33.5 ms vs 24.1 ms at 2048^2.)
// Three even-line and two odd-line buffers rotate; the per-cell state is double buffered.  The
// global loads of column x+2 (iterate, list bases, viscosity, rank class; dof ids of x+4) are
// cp.async (LDGSTS) copies issued by the matrix threads, so they have a full column to land.
#pragma once

__device__ __constant__ int8_t c_ct_node[3][3] = {{0, 4, 1}, {6, 8, 7}, {2, 5, 3}};  // [iy][ix] -> local node

struct CtDesc {  // one output list (static part)
  uint16_t soff;   // offset of the list slot inside its line buffer
  uint16_t roff;   // offset of the residual slot of this dof inside its line buffer
  uint16_t gdidx;  // cy*30 + local dof : where the global dof id / list base of this list lives
  uint8_t line;    // 0 left (even), 1 mid (odd), 2 right (even)
  uint8_t jp;      // node row inside the strip (ownership test); 255 for P lists
};

#define CT_UNI_Q 18                  // doubles per quadrature point (16 used)
#define CT_UNI_C (4 * CT_UNI_Q + 2)  // doubles per cell

template <int TH>
struct CtSmem {
  static constexpr int NC = TH + 1;
  static constexpr int EVEN_SZ = TH * CT_EVEN_STRIDE + CT_VTX_BLK;
  static constexpr int ODD_P0 = TH * CT_ODD_STRIDE + CT_EDG_BLK;
  static constexpr int ODD_SZ = ODD_P0 + TH * CT_P_BLK;
  static constexpr int NNODE = 3 * (2 * TH + 1);
  static constexpr int NDESC = NNODE * 3 + TH * 3;
  double even[3][EVEN_SZ];
  double odd[2][ODD_SZ];
  // per-cell state, padded so that lanes working on different cells / points hit different banks
  double uni[2][NC * CT_UNI_C];  // [cell][q][CT_UNI_Q]
  double val[4][NC][30];
  double visc[4][NC][4][2];
  double tphi[4][9], tdx[4][9], tdy[4][9], tw[4], twpsi[4][3], tpsi[4][3];
  double tdxy[4][9][2];  // (d_x phi_i, d_y phi_i) pairs for 128-bit loads
  long long lbase[4][NC][30];
  long long lnext[4][NC][30];
  int32_t gd[6][NC][30];
  int32_t cls[4][NC];
  CtDesc desc[NDESC];
};

// uniform state slots of uni[cell][q][*]
#define U_MU 0
#define U_G00 1
#define U_G01 2
#define U_G10 3
#define U_G11 4
#define U_DT0 5
#define U_DT1 6
#define U_SC 7   // jac_scaling * w_q|J| * 2 Pr * kappa_q
#define U_U0 8
#define U_U1 9
#define U_P 10
#define U_T 11

__device__ __forceinline__ void ct_bulk_store(double *gdst, const double *ssrc, int bytes) {
  const uint32_t sa = (uint32_t)__cvta_generic_to_shared(ssrc);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(sa), "r"(bytes) : "memory");
}

// three read-modify-writes to DISTINCT elements (or trash): loads first, so the three latencies overlap
__device__ __forceinline__ void ct_rmw3(char *base, int o0, int o1, int o2, double v0, double v1, double v2) {
  double *p0 = reinterpret_cast<double *>(base + o0), *p1 = reinterpret_cast<double *>(base + o1),
         *p2 = reinterpret_cast<double *>(base + o2);
  const double a0 = *p0, a1 = *p1, a2 = *p2;
  *p0 = a0 + v0; *p1 = a1 + v1; *p2 = a2 + v2;
}

__device__ __forceinline__ void ct_rmw(char *base, int off, double v) {
  double *p = reinterpret_cast<double *>(base + off);
  *p += v;
}

template <int TH, bool CSR, bool NEWT>
__global__ void __launch_bounds__((3 * TH / 2 + 2) * 32, (TH == 8 ? 1 : 2)) k_cart(CartArgs a) {
  using S = CtSmem<TH>;
  constexpr int NMW = 3 * TH / 2;        // matrix warps
  constexpr int NMT = NMW * 32;          // matrix threads
  constexpr int NT = NMT + 64;           // threads
  extern __shared__ __align__(16) unsigned char smem_raw[];
  S &s = *reinterpret_cast<S *>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool is_aux = warp >= NMW;
  const int at = tid - NMT;  // aux thread index 0..63
  const int y0 = a.row_begin + blockIdx.y * TH;
  const int th = min(TH, a.row_end - y0);
  const bool has_halo = y0 + th < a.row_end;  // (a rank's top interface line has no halo: NCCL adds the neighbour's part)
  const int ncol_cells = th + (has_halo ? 1 : 0);
  const int x0 = blockIdx.x * CT_XCH, x1 = min(a.nx, x0 + CT_XCH);
  const int jlo = y0 == a.row_begin ? 0 : 1;  // the first strip also assembles its bottom line (partial on rank > 0)
  const bool jac = (a.flags & GM_JACOBIAN) != 0, res = (a.flags & GM_RESIDUAL) != 0;
  const CartCell &cc = *a.cc;

  {
    double *z = &s.even[0][0];
    for (int k = tid; k < 3 * S::EVEN_SZ + 2 * S::ODD_SZ; k += NT) z[k] = 0.0;  // even[3], odd[2] are contiguous
    for (int k = tid; k < 36; k += NT) {
      s.tphi[0][k] = cc.sphi[0][k]; s.tdx[0][k] = cc.sdx[0][k]; s.tdy[0][k] = cc.sdy[0][k];
      s.tdxy[0][k][0] = cc.sdx[0][k]; s.tdxy[0][k][1] = cc.sdy[0][k];
    }
    if (tid < 4) s.tw[tid] = cc.w[tid];
    if (tid < 12) { s.twpsi[0][tid] = cc.wpsi[0][tid]; s.tpsi[0][tid] = cc.psi[0][tid]; }
  }
  // ---- static list descriptors
  for (int tk = tid; tk < S::NDESC; tk += NT) {
    CtDesc d;
    if (tk < S::NNODE * 3) {
      const int per_line = (2 * TH + 1) * 3;
      const int ln = tk / per_line, r = tk - ln * per_line, jp = r / 3, k = r - jp * 3;
      const bool odd_j = jp & 1;
      // the strip-top row (jp = 2*th) is resolved at run time (th may be < TH): gdidx keeps the
      // "bottom/mid node of cell jp>>1" form and is patched in the write-out
      const int node = c_ct_node[odd_j ? 1 : 0][ln];
      d.gdidx = (uint16_t)((jp >> 1) * 30 + (k == 0 ? node : (k == 1 ? 9 + node : 21 + node)));
      int blk, l1, l2, ro;
      if (ln == 1) {
        blk = (jp >> 1) * CT_ODD_STRIDE + (odd_j ? CT_EDG_BLK : 0);
        l1 = odd_j ? CT_CEN_L1 : CT_EDG_L1; l2 = odd_j ? CT_CEN_L2 : CT_EDG_L2; ro = odd_j ? CT_CEN_RES : CT_EDG_RES;
      } else {
        blk = (jp >> 1) * CT_EVEN_STRIDE + (odd_j ? CT_VTX_BLK : 0);
        l1 = odd_j ? CT_EDG_L1 : CT_VTX_L1; l2 = odd_j ? CT_EDG_L2 : CT_VTX_L2; ro = odd_j ? CT_EDG_RES : CT_VTX_RES;
      }
      d.soff = (uint16_t)(blk + (k == 0 ? 0 : (k == 1 ? l1 : l2)));
      d.roff = (uint16_t)(blk + ro + k);
      d.line = (uint8_t)ln;
      d.jp = (uint8_t)jp;
    } else {
      const int r = tk - S::NNODE * 3, cy = r / 3, p = r - cy * 3;
      d.gdidx = (uint16_t)(cy * 30 + 18 + p);
      d.soff = (uint16_t)(S::ODD_P0 + cy * CT_P_BLK + p * CT_P_L);
      d.roff = (uint16_t)(S::ODD_P0 + cy * CT_P_BLK + CT_P_RES + p);
      d.line = 1;
      d.jp = 255;
    }
    s.desc[tk] = d;
  }
  // ---- residual item of this aux lane: lane pair = node (ixn, jp) of the column's 3 x (2TH+1) node grid,
  //      lane parity = which of the node's (up to) two cells of this column it sums
  const int r_n = at >= 0 ? (at >> 1) : 0, r_side = at & 1;
  const int r_ixn = r_n / (2 * TH + 1) < 3 ? r_n / (2 * TH + 1) : 0, r_jp = r_n % (2 * TH + 1);
  const bool r_valid = is_aux && r_n < S::NNODE && r_jp >= jlo && r_jp <= 2 * th;
  int r_cy = -1, r_nd = 0;
  if (r_jp & 1) {
    if (r_side == 0) { r_cy = r_jp >> 1; r_nd = c_ct_node[1][r_ixn]; }
  } else if (r_side == 0) {  // node is the bottom of cell jp/2 ...
    if ((r_jp >> 1) < ncol_cells) { r_cy = r_jp >> 1; r_nd = c_ct_node[0][r_ixn]; }
  } else {                   // ... and the top of cell jp/2 - 1
    if ((r_jp >> 1) - 1 >= 0 && (r_jp >> 1) - 1 < th) { r_cy = (r_jp >> 1) - 1; r_nd = c_ct_node[2][r_ixn]; }
  }
  if (r_cy >= ncol_cells) r_cy = -1;
  const int r_dst = (r_ixn == 1) ? (r_jp >> 1) * CT_ODD_STRIDE + ((r_jp & 1) ? CT_EDG_BLK + CT_CEN_RES : CT_EDG_RES)
                                 : (r_jp >> 1) * CT_EVEN_STRIDE + ((r_jp & 1) ? CT_VTX_BLK + CT_EDG_RES : CT_VTX_RES);
  static_assert(2 * S::NNODE <= 64, "one residual lane pair per node of the column needs TH <= 4");

  // ---- lane role (matrix warps): task type (bottom/mid/top third), cell slot, (major node, minor node)
  const int tt = warp % 3, slot = warp / 3;
  const bool lane_on = lane < 27;
  const int Mi = lane_on ? lane / 9 : 0, mi = lane_on ? lane - Mi * 9 : 0;
  const int ixM = Mi, iyM = tt;
  const int M = c_ct_node[iyM][ixM];
  const int it = CSR ? M : mi, jt = CSR ? mi : M;
  double TA[4], TB[4], TC[4], TD[4], TM[4], TP[4], TK, TUT0, TUT1;
  {
    const CartPairTab &t = a.tab[it * 9 + jt];
#pragma unroll
    for (int q = 0; q < 4; ++q) { TA[q] = t.TA[q]; TB[q] = t.TB[q]; TC[q] = t.TC[q]; TD[q] = t.TD[q]; TM[q] = t.TM[q]; TP[q] = t.TP[q]; }
    TK = t.TK; TUT0 = t.TUT0; TUT1 = t.TUT1;
  }
  const bool mvtx = !(ixM & 1) && !(iyM & 1), mcen = (ixM & 1) && (iyM & 1);
  // list offsets / capacities of the major node's three lists (doubles)
  const int lo[3] = {0, mvtx ? CT_VTX_L1 : (mcen ? CT_CEN_L1 : CT_EDG_L1), mvtx ? CT_VTX_L2 : (mcen ? CT_CEN_L2 : CT_EDG_L2)};
  const int capU = lo[1], capT = (mvtx ? CT_VTX_RES : (mcen ? CT_CEN_RES : CT_EDG_RES)) - lo[2];
  const int ex_a = mi < 6 ? mi / 3 : 0, ex_p = mi % 3;  // lanes mi>=6 have no extra entry (trash of list 0)
  const double ex_val = (lane_on && mi < 6) ? (CSR ? cc.UP[ex_a * 9 + M][ex_p] : -cc.UP[ex_a * 9 + M][ex_p]) : 0.0;
  double pv[2];
  int pl[2], pd[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int e = 2 * (lane_on ? lane : 0) + k;
    pl[k] = (e / 18) % 3;
    pd[k] = e % 18;
    const double up = cc.UP[pd[k]][pl[k]];
    pv[k] = CSR ? -up : up;
  }
  // byte offsets (relative to the major node's block + list parity) of the lane's 12 entries; entries
  // without a slot (Dirichlet minor, idle lane) point at the trash element cap-2 of their list
  int o[12];
  int cur_cls = -1;
  // per-lane invariant part of the block address: ((iyM>>1)*stride + sub-block) in bytes
  const int stride_b = (ixM == 1 ? CT_ODD_STRIDE : CT_EVEN_STRIDE) * 8;
  const int blk0_b = ((iyM >> 1) * (ixM == 1 ? CT_ODD_STRIDE : CT_EVEN_STRIDE) +
                      ((iyM & 1) ? (ixM == 1 ? CT_EDG_BLK : CT_VTX_BLK) : 0)) * 8;
  // tasks of this warp (kernel invariant)
  const int cyA = 2 * slot, cyB = 2 * slot + 1;
  const bool doA = cyA < th && (2 * cyA + iyM) >= jlo;  // (<= 2*th always holds for cy < th)
  const bool doB = cyB < th;
  const bool doH = tt == 0 && slot == 0 && has_halo;

  // ---- global-load helpers: item = one (cell, local dof) of a column; PF_N items per thread
  const int xs = x0 > 0 ? x0 - 1 : x0;
  auto gd_addr = [&](int x, int item) -> const int32_t * {
    const int cy = item / 30, l = item - cy * 30;
    return a.gdofs + ((int64_t)(y0 + cy) * a.nx + x) * 30 + l;
  };
  auto cp_async8 = [&](void *sdst, const void *gsrc) {
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(sdst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gsrc) : "memory");
  };
  auto cp_async4 = [&](void *sdst, const void *gsrc) {
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(sdst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gsrc) : "memory");
  };
  auto cp_async16 = [&](void *sdst, const void *gsrc) {
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(sdst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
  };
  // asynchronous copies (LDGSTS, no registers) of one item of column x: iterate value, list base, next list base
  auto issue_item = [&](int x, int item) {
    const int32_t d = (&s.gd[x % 6][0][0])[item];
    const int r = x & 3;
    double *pv = &(&s.val[r][0][0])[item];
    long long *pb = &(&s.lbase[r][0][0])[item], *pn = &(&s.lnext[r][0][0])[item];
    if (d > 0) {
      cp_async8(pv, a.x + (d - 1));
      cp_async8(pb, a.ptr + (d - 1));
      cp_async8(pn, a.ptr + d);
    } else {
      cp_async8(pv, a.dvals + (-d - 1));
      *pb = 0; *pn = 0;
    }
  };
  auto issue_cell = [&](int x, int k) {  // k = cy*4 + q : viscosity pair and (q == 0) the rank class
    const int64_t c = (int64_t)(y0 + (k >> 2)) * a.nx + x;
    const int r = x & 3;
    if (!NEWT) cp_async16(&s.visc[r][k >> 2][k & 3][0], a.visc + (c * 4 + (k & 3)) * 2);
    if ((k & 3) == 0) cp_async4(&s.cls[r][k >> 2], a.cls + c);
  };
  // state of one column: S2 (fields at the quadrature points) and S4 (per-node state); `nthr` threads, index `t`
  auto state_S2 = [&](int x, int t, int nthr) {
    const int sb = x & 1, rb = x & 3;
    for (int k = t; k < ncol_cells * 16; k += nthr) {
      const int cy = k >> 4, q = (k >> 2) & 3, g = k & 3;
      double *u = &s.uni[sb][cy * CT_UNI_C + q * CT_UNI_Q];
      if (g == 3) {
        double v = 0;
#pragma unroll
        for (int m = 0; m < 3; ++m) v += s.tpsi[q][m] * s.val[rb][cy][18 + m];
        u[U_P] = v;
        u[U_MU] = s.visc[rb][cy][q][0];
        u[U_SC] = a.js * s.tw[q] * 2.0 * a.Pr * s.visc[rb][cy][q][1];
      } else {
        const double *vv = &s.val[rb][cy][g == 0 ? 0 : (g == 1 ? 9 : 21)];
        double f0 = 0, f1 = 0, f2 = 0;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
          const double tv = vv[i];
          f0 = fma(s.tphi[q][i], tv, f0); f1 = fma(s.tdx[q][i], tv, f1); f2 = fma(s.tdy[q][i], tv, f2);
        }
        if (g == 0) { u[U_U0] = f0; u[U_G00] = f1; u[U_G10] = f2; }
        else if (g == 1) { u[U_U1] = f0; u[U_G01] = f1; u[U_G11] = f2; }
        else { u[U_T] = f0; u[U_DT0] = f1; u[U_DT1] = f2; }
      }
    }
  };
  // ---- prologue (all threads): dof ids of columns xs..xs+3, data of xs, xs+1, state of xs
  for (int k = tid; k < 4 * S::NC * 4; k += NT) { (&s.visc[0][0][0][0])[2 * k] = 1.0; (&s.visc[0][0][0][0])[2 * k + 1] = 0.0; }
  for (int item = tid; item < ncol_cells * 30; item += NT)
    for (int dxc = 0; dxc < 4; ++dxc)
      if (xs + dxc < x1) (&s.gd[(xs + dxc) % 6][0][0])[item] = *gd_addr(xs + dxc, item);
  __syncthreads();
  for (int dxc = 0; dxc < 2; ++dxc) {
    if (xs + dxc >= x1) break;
    for (int item = tid; item < ncol_cells * 30; item += NT) issue_item(xs + dxc, item);
    for (int k = tid; k < ncol_cells * 4; k += NT) issue_cell(xs + dxc, k);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  state_S2(xs, tid, NT);
  __syncthreads();

  int eL = 0, eR = 1, eI = 2;  // even buffers: left line, right line, idle (being written out / zeroed)
  bool prev_out = false;       // the previous column has completed lists waiting to be written out

  // ---- one third of a cell matrix: 9 pair values + constant P-block entries -> shared lists
  auto matrix_task = [&](int cy, int sb, int rb, char *lineb, char *pblk, bool with_p) {
    // class / offsets (warp uniform branch, taken only when the class changes)
    const int cl = s.cls[rb][cy];
    if (cl != cur_cls) {
      cur_cls = cl;
      const uint8_t *t = a.ctab + (int64_t)cl * 900;
      const int Ld[3] = {M, 9 + M, 21 + M};
      const int md[3] = {mi, 9 + mi, 21 + mi};
#pragma unroll
      for (int kk = 0; kk < 3; ++kk)
#pragma unroll
        for (int ll = 0; ll < 3; ++ll) {
          const int r = lane_on ? t[md[ll] * 30 + Ld[kk]] : 0xFF;
          o[kk * 3 + ((ll < 2 && (mi & 1)) ? 1 - ll : ll)] = (lo[kk] + (r == 0xFF ? (kk == 2 ? capT : capU) - 2 : r)) * 8;
        }
      {
        const int r = (lane_on && mi < 6) ? t[(18 + ex_p) * 30 + Ld[ex_a]] : 0xFF;
        o[9] = (lo[ex_a] + (r == 0xFF ? capU - 2 : r)) * 8;
      }
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int r = (lane_on && tt == 1) ? t[pd[k] * 30 + 18 + pl[k]] : 0xFF;
        o[10 + k] = (pl[k] * CT_P_L + (r == 0xFF ? CT_P_L - 2 : r)) * 8;
      }
    }
    double v00 = 0, v01 = 0, v10 = 0, v11 = 0, e = 0, tu0 = 0, tu1 = 0;
    const double2 *u2 = reinterpret_cast<const double2 *>(&s.uni[sb][cy * CT_UNI_C]);
    const double2 *dxy = reinterpret_cast<const double2 *>(&s.tdxy[0][0][0]);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const double2 a0 = u2[q * (CT_UNI_Q / 2) + 0], a1 = u2[q * (CT_UNI_Q / 2) + 1], a2 = u2[q * (CT_UNI_Q / 2) + 2],
                    a3 = u2[q * (CT_UNI_Q / 2) + 3], a4 = u2[q * (CT_UNI_Q / 2) + 4];  // (mu,G00) (G01,G10) (G11,dT0) (dT1,sc) (u0,u1)
      const double2 di = dxy[q * 9 + it], dj = dxy[q * 9 + jt];  // physical gradients of the test / trial shape function
      const double mu = a0.x;
      v00 = fma(mu, TA[q], v00); v11 = fma(mu, TB[q], v11); v01 = fma(mu, TC[q], v01); v10 = fma(mu, TD[q], v10);
      if (!NEWT) {
        // S_ia = sum_k D[k][a] d_k phi_i (test), S'_jb = sc * S_jb (trial)   
        const double D01 = 0.5 * (a1.x + a1.y);
        const double si0 = fma(a0.y, di.x, D01 * di.y), si1 = fma(D01, di.x, a2.x * di.y);
        const double sj0 = a3.y * fma(a0.y, dj.x, D01 * dj.y), sj1 = a3.y * fma(D01, dj.x, a2.x * dj.y);
        v00 = fma(si0, sj0, v00); v01 = fma(si0, sj1, v01); v10 = fma(si1, sj0, v10); v11 = fma(si1, sj1, v11);
      }
      const double cj = fma(a4.x, dj.x, a4.y * dj.y);  // u . grad phi_j
      v00 = fma(TM[q], a0.y, v00);
      v01 = fma(TM[q], a1.y, v01);  // (a=0,b=1): G[1][0]
      v10 = fma(TM[q], a1.x, v10);  // (a=1,b=0): G[0][1]
      v11 = fma(TM[q], a2.x, v11);
      e = fma(TP[q], cj, e);
      tu0 = fma(TM[q], a2.y, tu0);
      tu1 = fma(TM[q], a3.x, tu1);
    }
    v00 += e; v11 += e;
    const double tt_v = TK + e;
    // block of the major node, list bases shifted by the parity of their global base (TMA alignment)
    char *blk = lineb + blk0_b + cy * stride_b;
    const int *lbw = reinterpret_cast<const int *>(&s.lbase[rb][cy][0]);
    char *b0 = blk + ((lbw[2 * M] & 1) << 3);
    char *b1 = blk + ((lbw[2 * (9 + M)] & 1) << 3);
    char *b2 = blk + ((lbw[2 * (21 + M)] & 1) << 3);
    // entries (list k, minor ux) and (list k, minor uy) sit at even / odd positions of the list: lanes
    // with an odd minor node handle them in swapped order so one warp access covers both bank parities
    const bool sw = mi & 1;
    double e00, e01, e02, e10, e11, e12, e20, e21, e22;  // [list k][minor dof l]
    if (CSR) { e00 = v00; e01 = v01; e02 = TUT0; e10 = v10; e11 = v11; e12 = TUT1; e20 = tu0; e21 = tu1; e22 = tt_v; }
    else     { e00 = v00; e01 = v10; e02 = tu0;  e10 = v01; e11 = v11; e12 = tu1;  e20 = TUT0; e21 = TUT1; e22 = tt_v; }
    ct_rmw3(b0, o[0], o[1], o[2], sw ? e01 : e00, sw ? e00 : e01, e02);
    ct_rmw3(b1, o[3], o[4], o[5], sw ? e11 : e10, sw ? e10 : e11, e12);
    ct_rmw3(b2, o[6], o[7], o[8], sw ? e21 : e20, sw ? e20 : e21, e22);
    ct_rmw(ex_a == 0 ? b0 : b1, o[9], ex_val);
    if (with_p) {  // warp uniform: P major lists of this cell (mid third only)
      char *pb = pblk + cy * (CT_P_BLK * 8);
      ct_rmw(pb + ((lbw[2 * (18 + pl[0])] & 1) << 3), o[10], pv[0]);
      ct_rmw(pb + ((lbw[2 * (18 + pl[1])] & 1) << 3), o[11], pv[1]);
    }
  };

  // ---- write-out of column xw (aux warps): one TMA bulk copy per completed list
  auto write_out = [&](int xw, int bL, int bR) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const int per_line = (2 * TH + 1) * 3;
    const int nline_tasks = ((xw == a.nx - 1) ? 3 : 2) * per_line;
    const int wb = xw & 1, wr = xw & 3;
    const int32_t *gdw = &s.gd[xw % 6][0][0];
    for (int tk = at; tk < 3 * per_line + th * 3; tk += 64) {
      if (tk >= nline_tasks && tk < 3 * per_line) continue;
      const CtDesc d = s.desc[tk];
      int gi = d.gdidx;
      if (d.jp != 255) {
        if (d.jp < jlo || d.jp > 2 * th) continue;
        if (d.jp == 2 * th) gi += -30 + (c_ct_node[2][d.line] - c_ct_node[0][d.line]);  // top nodes of cell th-1
      }
      const int32_t g = gdw[gi];
      double *line = d.line == 1 ? s.odd[wb] : s.even[d.line == 0 ? bL : bR];
      if (g > 0) {
        if (jac) {
          const long long base = (&s.lbase[wr][0][0])[gi];
          const int len = (int)((&s.lnext[wr][0][0])[gi] - base);
          const int par = (int)(base & 1);
          const double *src = line + d.soff + par;  // element k of the list lives at src[k]
          double *dst = a.nz + base;
          int k0 = 0, n = len;
          if (par) { dst[0] = src[0]; k0 = 1; n -= 1; }
          if (n & 1) { dst[k0 + n - 1] = src[k0 + n - 1]; n -= 1; }
          if (n > 0) ct_bulk_store(dst + k0, src + k0, n * 8);
        }
        if (res) {
          a.b[g - 1] = line[d.roff];
        }
      }
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  };
  auto zero_lines = [&](int be, int bo) {  // aux warps: zero an even and an odd line buffer
    double2 *z0 = reinterpret_cast<double2 *>(s.even[be]);
    for (int k = at; k < S::EVEN_SZ / 2; k += 64) z0[k] = make_double2(0.0, 0.0);
    double2 *z1 = reinterpret_cast<double2 *>(s.odd[bo]);
    for (int k = at; k < S::ODD_SZ / 2; k += 64) z1[k] = make_double2(0.0, 0.0);
  };

  for (int x = xs; x < x1; ++x) {
    const bool pre = x < x0;
    const int sb = x & 1;        // state buffer of this column
    const int rb = x & 3;        // value / list-base ring slot of this column
    const int ob = x & 1;        // odd (mid) line buffer of this column
    if (!is_aux) {
      // ---- asynchronous global->shared copies of column x+2 (dof ids of x+4): one item per thread, no registers
      if (x + 2 < x1) {
        if (tid < ncol_cells * 30) issue_item(x + 2, tid);
        if (tid < ncol_cells * 4) issue_cell(x + 2, tid);
      }
      if (x + 4 < x1 && tid < ncol_cells * 30) cp_async4(&(&s.gd[(x + 4) % 6][0][0])[tid], gd_addr(x + 4, tid));
      asm volatile("cp.async.commit_group;" ::: "memory");
      if (jac) {
        // =============================== matrix phases ===============================
        char *lineb = reinterpret_cast<char *>(ixM == 1 ? s.odd[ob] : s.even[ixM == 0 ? eL : eR]);
        char *pblk = reinterpret_cast<char *>(s.odd[ob] + S::ODD_P0);
        const bool lane_go = !(pre && ixM != 2);  // halo column: only the right line is owned
        if (lane_go) {
          if (doA) matrix_task(cyA, sb, rb, lineb, pblk, tt == 1 && !pre);
          if (doH) matrix_task(th, sb, rb, lineb, pblk, false);
        }
        asm volatile("cp.async.wait_group 1;" ::: "memory");  // the copies issued ONE column ago (column x+1) have landed
        __syncwarp();
        asm volatile("bar.sync 1, %0;" ::"n"(NMT) : "memory");  // matrix warps only: even rows done, column x+1 data visible
        if (lane_go && doB) matrix_task(cyB, sb, rb, lineb, pblk, tt == 1 && !pre);
      } else {
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncwarp();
        asm volatile("bar.sync 1, %0;" ::"n"(NMT) : "memory");
      }
      // ---- state of the next column (its values were parked in shared memory one column ago)
      if (x + 1 < x1) {
        state_S2(x + 1, tid, NMT);
      }
    } else {
      // ---- (2) write-out of the previous column
      if (prev_out) write_out(x - 1, eI, eR);  // previous left line = this column's idle buffer
      // ---- (4) residual gather of this column: lane pair = (node, side); partial sums combined by shuffle
      if (res) {
        double r0 = 0, r1 = 0, r2 = 0;
        if (r_valid && !(pre && r_ixn != 2) && r_cy >= 0) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const double *u = &s.uni[sb][r_cy * CT_UNI_C + q * CT_UNI_Q];
            const double w = s.tw[q];
            const double gx_ = s.tdx[q][r_nd], gy_ = s.tdy[q][r_nd];
            const double ph = w * s.tphi[q][r_nd], dx = w * gx_, dy = w * gy_;
            const double u0 = u[U_U0], u1 = u[U_U1], G00 = u[U_G00], G01 = u[U_G01], G10 = u[U_G10], G11 = u[U_G11];
            const double D01 = 0.5 * (G01 + G10);
            const double m2 = 2.0 * a.Pr * w * u[U_MU], bu = a.RaPr * u[U_T];
            r0 = fma(m2, fma(G00, gx_, D01 * gy_), r0); r0 = fma(ph, fma(u0, G00, u1 * G10) - bu * a.gx, r0); r0 = fma(-dx, u[U_P], r0);
            r1 = fma(m2, fma(D01, gx_, G11 * gy_), r1); r1 = fma(ph, fma(u0, G01, u1 * G11) - bu * a.gy, r1); r1 = fma(-dy, u[U_P], r1);
            r2 = fma(u[U_DT0], dx, r2); r2 = fma(u[U_DT1], dy, r2); r2 = fma(fma(u0, u[U_DT0], u1 * u[U_DT1]), ph, r2);
          }
        }
        __syncwarp();
        r0 += __shfl_xor_sync(0xffffffffu, r0, 1);
        r1 += __shfl_xor_sync(0xffffffffu, r1, 1);
        r2 += __shfl_xor_sync(0xffffffffu, r2, 1);
        if (r_valid && !(pre && r_ixn != 2) && (at & 1) == 0) {
          double *dst = ((r_ixn == 1) ? s.odd[ob] : s.even[r_ixn == 0 ? eL : eR]) + r_dst;
          dst[0] += r0; dst[1] += r1; dst[2] += r2;
        }
        if (!pre && at < th * 3) {
          const int cy = at / 3, p = at - cy * 3;
          double acc = 0;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const double *u = &s.uni[sb][cy * CT_UNI_C + q * CT_UNI_Q];
            acc += s.twpsi[q][p] * (u[U_G00] + u[U_G11]);
          }
          s.odd[ob][S::ODD_P0 + cy * CT_P_BLK + CT_P_RES + p] = acc;
        }
      }
      // ---- (5) recycle the buffers whose copies (issued in (2)) have drained
      // (wait_group is per thread: every aux thread waits for its own copies, then all 64 meet)
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
      asm volatile("bar.sync 3, 64;" ::: "memory");
      if (prev_out) zero_lines(eI, ob ^ 1);
    }
    __syncwarp();
    __syncthreads();  // A: lists of column x complete, state of x+1 ready, recycled buffers zero
    prev_out = !pre;
    if (pre) {
      const int t = eL; eL = eR; eR = t;  // nothing to write out: the old left line is still all zero
    } else {
      const int t = eL; eL = eR; eR = eI; eI = t;
    }
  }
  // last column of the chunk: write it out (including the right line at the mesh boundary)
  if (is_aux && prev_out) {
    write_out(x1 - 1, eI, eL);  // after the final rotation: previous left = eI, previous right = eL
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}
