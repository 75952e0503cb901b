// gm_cart_kernel.cuh -- the sweep kernel of the structured path (see gm_cart.cuh for the design).
//
// CTA = 3*TH/2 matrix warps + 2 aux warps.  Per column x of the strip:
//   all warps : state phase (fields at the quadrature points, S, u.grad(phi)) for the TH(+1 halo) cells
//   matrix    : two colour phases (even / odd cell rows); lane = (major node, minor node) pair,
//               pair tables in registers, 9 values accumulated into the shared-memory lists
//   aux       : residual gather for the column (concurrently with the matrix phases), then the
//               write-out: one TMA bulk copy (cp.async.bulk shared->global) per completed list
//   all warps : zero the line buffers whose copies have drained (three even / two odd line
//               buffers rotate so a copy has a full column of time to drain)
// Global loads of the next column (dof ids two columns ahead, iterate / list bases / viscosity
// one column ahead) are issued before the matrix phases and parked in registers.
#pragma once

__device__ __constant__ int8_t c_ct_node[3][3] = {{0, 4, 1}, {6, 8, 7}, {2, 5, 3}};  // [iy][ix] -> local node

struct CtDesc {  // one output list (static part)
  uint16_t soff;   // offset of the list slot inside its line buffer
  uint16_t roff;   // offset of the residual slot of this dof inside its line buffer
  uint16_t gdidx;  // cy*30 + local dof : where the global dof id / list base of this list lives
  uint8_t line;    // 0 left (even), 1 mid (odd), 2 right (even)
  uint8_t jp;      // node row inside the strip (ownership test); 255 for P lists
};

template <int TH>
struct CtSmem {
  static constexpr int NC = TH + 1;
  static constexpr int EVEN_SZ = TH * CT_EVEN_STRIDE + CT_VTX_BLK;
  static constexpr int ODD_P0 = TH * CT_ODD_STRIDE + CT_EDG_BLK;
  static constexpr int ODD_SZ = ODD_P0 + TH * CT_P_BLK;
  static constexpr int NDESC = 3 * (2 * TH + 1) * 3 + TH * 3;
  double even[3][EVEN_SZ];
  double odd[2][ODD_SZ];
  double uni[NC][4][CT_UNI];
  double nod[NC][4][9][CT_NOD];
  double val[2][NC][30];
  double visc[2][NC][4][2];
  long long lbase[2][NC][30];
  int32_t llen[2][NC][30];
  int32_t gd[3][NC][30];
  int32_t cls[2][NC];
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
#define U_KAP 7
#define U_CONV0 8
#define U_CONV1 9
#define U_P 10
#define U_T 11
#define U_UDT 12
#define U_DIVU 13
#define U_U0 14
#define U_U1 15

__device__ __forceinline__ void ct_bulk_store(double *gdst, const double *ssrc, int bytes) {
  const uint32_t sa = (uint32_t)__cvta_generic_to_shared(ssrc);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(sa), "r"(bytes) : "memory");
}

template <int TH, bool CSR, bool NEWT>
__global__ void __launch_bounds__((3 * TH / 2 + 2) * 32, (TH == 8 ? 1 : 2)) k_cart(CartArgs a) {
  using S = CtSmem<TH>;
  constexpr int NMW = 3 * TH / 2;        // matrix warps
  constexpr int NT = (NMW + 2) * 32;     // threads
  constexpr int NC = TH + 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  S &s = *reinterpret_cast<S *>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool is_aux = warp >= NMW;
  const int y0 = blockIdx.y * TH;
  const int th = min(TH, a.ny - y0);
  const bool has_halo = y0 + th < a.ny;
  const int ncol_cells = th + (has_halo ? 1 : 0);
  const int x0 = blockIdx.x * CT_XCH, x1 = min(a.nx, x0 + CT_XCH);
  const int jlo = y0 == 0 ? 0 : 1;
  const bool jac = (a.flags & GM_JACOBIAN) != 0, res = (a.flags & GM_RESIDUAL) != 0;
  const CartCell &cc = *a.cc;

  {
    double *z = &s.even[0][0];
    for (int k = tid; k < 3 * S::EVEN_SZ + 2 * S::ODD_SZ; k += NT) z[k] = 0.0;  // even[3], odd[2] are contiguous
  }
  // ---- static list descriptors
  for (int tk = tid; tk < S::NDESC; tk += NT) {
    CtDesc d;
    if (tk < 3 * (2 * TH + 1) * 3) {
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
      const int r = tk - 3 * (2 * TH + 1) * 3, cy = r / 3, p = r - cy * 3;
      d.gdidx = (uint16_t)(cy * 30 + 18 + p);
      d.soff = (uint16_t)(S::ODD_P0 + cy * CT_P_BLK + p * CT_P_L);
      d.roff = (uint16_t)(S::ODD_P0 + cy * CT_P_BLK + CT_P_RES + p);
      d.line = 1;
      d.jp = 255;
    }
    s.desc[tk] = d;
  }

  // ---- lane role (matrix warps): task type (bottom/mid/top third), cell slot, (major node, minor node)
  const int tt = warp % 3, slot = warp / 3;
  const int Mi = lane / 9, mi = lane - Mi * 9;
  const bool pair_lane = lane < 27 && !is_aux;
  const int ixM = lane < 27 ? Mi : 0, iyM = tt;
  const int M = c_ct_node[iyM][ixM];
  const int it = CSR ? M : mi, jt = CSR ? mi : M;
  double TA[4], TB[4], TC[4], TD[4], TM[4], TP[4], TK, TUT0, TUT1;
  {
    const CartPairTab &t = a.tab[it * 9 + jt];
#pragma unroll
    for (int q = 0; q < 4; ++q) { TA[q] = t.TA[q]; TB[q] = t.TB[q]; TC[q] = t.TC[q]; TD[q] = t.TD[q]; TM[q] = t.TM[q]; TP[q] = t.TP[q]; }
    TK = t.TK; TUT0 = t.TUT0; TUT1 = t.TUT1;
  }
  const int Ld[3] = {M, 9 + M, 21 + M};
  const int md[3] = {mi, 9 + mi, 21 + mi};
  const bool mvtx = !(ixM & 1) && !(iyM & 1), mcen = (ixM & 1) && (iyM & 1);
  const int lo1 = mvtx ? CT_VTX_L1 : (mcen ? CT_CEN_L1 : CT_EDG_L1);
  const int lo2 = mvtx ? CT_VTX_L2 : (mcen ? CT_CEN_L2 : CT_EDG_L2);
  const int ex_a = mi / 3, ex_p = mi % 3;
  const double ex_val = (lane < 27 && mi < 6) ? (CSR ? cc.UP[ex_a * 9 + M][ex_p] : -cc.UP[ex_a * 9 + M][ex_p]) : 0.0;
  double pv[2];
  int pl[2], pd[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int e = 2 * lane + k;
    pl[k] = (e / 18) % 3;
    pd[k] = e % 18;
    const double up = cc.UP[pd[k]][pl[k]];
    pv[k] = CSR ? -up : up;
  }
  int cur_cls = -1;
  uint8_t rk[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  uint8_t rk_ex = 0xFF, rk_p[2] = {0xFF, 0xFF};

  // ---- global-load helpers (one item = one (cell, local dof) of a column)
  const int pf_cy = tid / 30, pf_l = tid - pf_cy * 30;
  const bool pf_on = tid < ncol_cells * 30;
  auto load_gd = [&](int x) -> int32_t {
    return a.gdofs[((int64_t)(y0 + pf_cy) * a.nx + x) * 30 + pf_l];
  };
  const int xs = x0 > 0 ? x0 - 1 : x0;
  // prologue: dof ids of the first two columns, values of the first
  if (pf_on) {
    s.gd[xs % 3][pf_cy][pf_l] = load_gd(xs);
    if (xs + 1 < x1) s.gd[(xs + 1) % 3][pf_cy][pf_l] = load_gd(xs + 1);
  }
  __syncthreads();
  auto fetch_col = [&](int x, double &v, long long &lb, int &ll, double &vm, double &vk, int &cl) {
    // dependent loads of column x (its dof ids are already in shared memory)
    if (pf_on) {
      const int32_t d = s.gd[x % 3][pf_cy][pf_l];
      if (d > 0) {
        v = a.x[d - 1];
        lb = a.ptr[d - 1];
        ll = (int)(a.ptr[d] - lb);
      } else {
        v = a.dvals[-d - 1]; lb = 0; ll = 0;
      }
    }
    if (tid < ncol_cells * 4) {
      const int64_t c = (int64_t)(y0 + (tid >> 2)) * a.nx + x;
      if (!NEWT) { vm = a.visc[(c * 4 + (tid & 3)) * 2]; vk = a.visc[(c * 4 + (tid & 3)) * 2 + 1]; }
      if ((tid & 3) == 0) cl = a.cls[c];
    }
  };
  auto store_col = [&](int x, double v, long long lb, int ll, double vm, double vk, int cl) {
    const int b = x & 1;
    if (pf_on) { s.val[b][pf_cy][pf_l] = v; s.lbase[b][pf_cy][pf_l] = lb; s.llen[b][pf_cy][pf_l] = ll; }
    if (tid < ncol_cells * 4) {
      s.visc[b][tid >> 2][tid & 3][0] = NEWT ? 1.0 : vm;
      s.visc[b][tid >> 2][tid & 3][1] = NEWT ? 0.0 : vk;
      if ((tid & 3) == 0) s.cls[b][tid >> 2] = cl;
    }
  };
  {
    double v = 0, vm = 1, vk = 0; long long lb = 0; int ll = 0, cl = 0;
    fetch_col(xs, v, lb, ll, vm, vk, cl);
    store_col(xs, v, lb, ll, vm, vk, cl);
  }
  __syncthreads();

  int eL = 0, eR = 1, eI = 2;  // even buffers: left line, right line, idle (being written out / zeroed)
  bool prev_out = false;       // the previous column has completed lists waiting to be written out
  // write-out of column xw (aux warps): one TMA bulk copy per completed list
  auto write_out = [&](int xw, int bL, int bR) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const int at = tid - NMW * 32;
    const int per_line = (2 * TH + 1) * 3;
    const int nline_tasks = ((xw == a.nx - 1) ? 3 : 2) * per_line;
    const int wb = xw & 1;
    const int32_t *gdw = &s.gd[xw % 3][0][0];
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
          const long long base = (&s.lbase[wb][0][0])[gi];
          const int len = (&s.llen[wb][0][0])[gi];
          const int par = (int)(base & 1);
          const double *src = line + d.soff + par;  // element k of the list lives at src[k]
          double *dst = a.nz + base;
          int k0 = 0, n = len;
          if (par) { dst[0] = src[0]; k0 = 1; n -= 1; }
          if (n & 1) { dst[k0 + n - 1] = src[k0 + n - 1]; n -= 1; }
          if (n > 0) {
            if (a.flags & 0x100u) {
            } else if (a.flags & 0x200u) {
              for (int k2 = 0; k2 < n; k2 += 2) *reinterpret_cast<double2 *>(dst + k0 + k2) = *reinterpret_cast<const double2 *>(src + k0 + k2);
            } else {
              ct_bulk_store(dst + k0, src + k0, n * 8);
            }
          }
        }
        if (res) a.b[g - 1] = line[d.roff];
      }
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  };
  auto zero_lines = [&](int be, int bo) {  // aux warps: zero an even and an odd line buffer
    const int at = tid - NMW * 32;
    double2 *z0 = reinterpret_cast<double2 *>(s.even[be]);
    for (int k = at; k < S::EVEN_SZ / 2; k += 64) z0[k] = make_double2(0.0, 0.0);
    double2 *z1 = reinterpret_cast<double2 *>(s.odd[bo]);
    for (int k = at; k < S::ODD_SZ / 2; k += 64) z1[k] = make_double2(0.0, 0.0);
  };

  for (int x = xs; x < x1; ++x) {
    const bool pre = x < x0;
    const int sb = x & 1;        // state / value buffer of this column
    const int ob = x & 1;        // odd (mid) line buffer of this column
    double nv = 0, nvm = 1, nvk = 0; long long nlb = 0; int nll = 0, ncl = 0;
    int32_t ngd = 0;
    const bool have_next = x + 1 < x1;
    if (!is_aux) {
      // =============================== state phase (matrix warps) ===============================
    // S2: fields at the quadrature points.  item = (cell, q, group): group 0: u0,G00,G10 | 1: u1,G01,G11 | 2: T,dT0,dT1 | 3: p
    for (int k = tid; k < ncol_cells * 16; k += NMW * 32) {
      const int cy = k >> 4, q = (k >> 2) & 3, g = k & 3;
      double *u = s.uni[cy][q];
      if (g == 3) {
        double v = 0;
#pragma unroll
        for (int m = 0; m < 3; ++m) v += cc.psi[q][m] * s.val[sb][cy][18 + m];
        u[U_P] = v;
        u[U_MU] = s.visc[sb][cy][q][0];
        u[U_KAP] = s.visc[sb][cy][q][1];
      } else {
        const double *vv = &s.val[sb][cy][g == 0 ? 0 : (g == 1 ? 9 : 21)];
        double f0 = 0, f1 = 0, f2 = 0;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
          const double t = vv[i];
          f0 = fma(cc.sphi[q][i], t, f0); f1 = fma(cc.sdx[q][i], t, f1); f2 = fma(cc.sdy[q][i], t, f2);
        }
        if (g == 0) { u[U_U0] = f0; u[U_G00] = f1; u[U_G10] = f2; }
        else if (g == 1) { u[U_U1] = f0; u[U_G01] = f1; u[U_G11] = f2; }
        else { u[U_T] = f0; u[U_DT0] = f1; u[U_DT1] = f2; }
      }
    }
    asm volatile("bar.sync 2, %0;" ::"n"(NMW * 32) : "memory");  // matrix warps only
    // S4: per-node state; the node-0 item also derives the remaining uniform values
    for (int k = tid; k < ncol_cells * 36; k += NMW * 32) {
      const int cy = k / 36, r = k - cy * 36, q = r / 9, i = r - q * 9;
      double *u = s.uni[cy][q];
      const double u0 = u[U_U0], u1 = u[U_U1], G00 = u[U_G00], G01 = u[U_G01], G10 = u[U_G10], G11 = u[U_G11];
      const double dx = cc.sdx[q][i], dy = cc.sdy[q][i];
      const double D01 = 0.5 * (G01 + G10);
      const double S0 = G00 * dx + D01 * dy;
      const double S1 = D01 * dx + G11 * dy;
      const double sc = a.js * cc.w[q] * 2.0 * a.Pr * u[U_KAP];
      double *o = s.nod[cy][q][i];
      o[0] = S0; o[1] = S1; o[2] = sc * S0; o[3] = sc * S1;
      o[4] = u0 * dx + u1 * dy;
      if (i == 0) {
        u[U_CONV0] = u0 * G00 + u1 * G10;
        u[U_CONV1] = u0 * G01 + u1 * G11;
        u[U_UDT] = u0 * u[U_DT0] + u1 * u[U_DT1];
        u[U_DIVU] = G00 + G11;
      }
    }
    } else if (prev_out) {
      // ============ write-out of the previous column (aux warps), overlapped with the state phase ============
      write_out(x - 1, eI, eR);   // previous left line is this column's idle buffer; its right line (last column only) never gets here
    }
    __syncthreads();  // B: state of column x ready
    if (!is_aux) {
      // ---- issue the global loads of the next columns (parked in registers over the matrix phases)
            if (have_next) {
        fetch_col(x + 1, nv, nlb, nll, nvm, nvk, ncl);
        if (x + 2 < x1 && pf_on) ngd = load_gd(x + 2);
      }
      if (jac) {
        // =============================== matrix phases ===============================
        for (int ph = 0; ph < 2; ++ph) {
          for (int rep = 0; rep < 2; ++rep) {
            int cy;
            if (rep == 0) cy = 2 * slot + ph;
            else { if (!(ph == 0 && tt == 0 && slot == 0 && has_halo)) break; cy = th; }
            if (rep == 0 && cy >= th) continue;
            const int jp = 2 * cy + iyM;
            if (jp < jlo || jp > 2 * th) continue;
            const int cl = s.cls[sb][cy];
            if (cl != cur_cls) {
              cur_cls = cl;
              const uint8_t *t = a.ctab + (int64_t)cl * 900;
              if (lane < 27) {
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
            if (pre && ixM != 2) continue;
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
            double *line = (ixM == 1) ? s.odd[ob] : s.even[ixM == 0 ? eL : eR];
            const int row = cy + (iyM >> 1);
            double *blk = line + ((ixM == 1) ? row * CT_ODD_STRIDE + ((iyM & 1) ? CT_EDG_BLK : 0)
                                             : row * CT_EVEN_STRIDE + ((iyM & 1) ? CT_VTX_BLK : 0));
            double w9[9];
            if (CSR) {
              w9[0] = v00; w9[1] = v01; w9[2] = TUT0; w9[3] = v10; w9[4] = v11; w9[5] = TUT1; w9[6] = tu0; w9[7] = tu1; w9[8] = tt_v;
            } else {
              w9[0] = v00; w9[1] = v10; w9[2] = tu0; w9[3] = v01; w9[4] = v11; w9[5] = tu1; w9[6] = TUT0; w9[7] = TUT1; w9[8] = tt_v;
            }
            // list k starts at slot + parity of its global base, so that shared and global addresses
            // are 16-byte aligned together (TMA bulk copy)
            int par[3];
  #pragma unroll
            for (int kk = 0; kk < 3; ++kk) par[kk] = (int)(s.lbase[sb][cy][Ld[kk]] & 1);
  #pragma unroll
            for (int kk = 0; kk < 3; ++kk) {
              double *lst = blk + (kk == 0 ? 0 : (kk == 1 ? lo1 : lo2)) + par[kk];
  #pragma unroll
              for (int ll = 0; ll < 3; ++ll) {
                const uint8_t r = rk[kk * 3 + ll];
                if (r != 0xFF) lst[r] += w9[kk * 3 + ll];
              }
            }
            if (rk_ex != 0xFF) (blk + (ex_a == 0 ? par[0] : lo1 + par[1]))[rk_ex] += ex_val;
            if (tt == 1 && cy < th && !pre) {
              double *pb = s.odd[ob] + S::ODD_P0 + cy * CT_P_BLK;
  #pragma unroll
              for (int k = 0; k < 2; ++k)
                if (rk_p[k] != 0xFF) pb[pl[k] * CT_P_L + (int)(s.lbase[sb][cy][18 + pl[k]] & 1) + rk_p[k]] += pv[k];
            }
          }
          if (ph == 0) asm volatile("bar.sync 1, %0;" ::"n"(NMW * 32) : "memory");  // matrix warps only
        }
      }
      // park the prefetched column in shared memory
      if (have_next) {
        store_col(x + 1, nv, nlb, nll, nvm, nvk, ncl);
        if (x + 2 < x1 && pf_on) s.gd[(x + 2) % 3][pf_cy][pf_l] = ngd;
      }
    } else {
      // =========================== residual gather (aux warps) ===========================
      if (res) {
        const int at = tid - NMW * 32;
        for (int k = at; k < 3 * (2 * TH + 1) * 3; k += 64) {
          const int ixn = k / ((2 * TH + 1) * 3), r = k - ixn * ((2 * TH + 1) * 3), jp = r / 3, comp = r - jp * 3;
          if (jp < jlo || jp > 2 * th || (pre && ixn != 2)) continue;
          double acc = 0;
          for (int side = 0; side < 2; ++side) {
            int cy, iy;
            if (jp & 1) { if (side) break; cy = jp >> 1; iy = 1; }
            else { cy = (jp >> 1) - side; iy = side ? 2 : 0; }
            if (cy < 0 || cy >= ncol_cells) continue;
            if (cy == th && iy != 0) continue;
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
          double *line = (ixn == 1) ? s.odd[ob] : s.even[ixn == 0 ? eL : eR];
          const int blk = (ixn == 1) ? (jp >> 1) * CT_ODD_STRIDE + ((jp & 1) ? CT_EDG_BLK + CT_CEN_RES : CT_EDG_RES)
                                     : (jp >> 1) * CT_EVEN_STRIDE + ((jp & 1) ? CT_VTX_BLK + CT_EDG_RES : CT_VTX_RES);
          line[blk + comp] += acc;
        }
        if (!pre)
          for (int k = at; k < th * 3; k += 64) {
            const int cy = k / 3, p = k - cy * 3;
            double acc = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) acc += cc.wpsi[q][p] * s.uni[cy][q][U_DIVU];
            s.odd[ob][S::ODD_P0 + cy * CT_P_BLK + CT_P_RES + p] = acc;
          }
      }
      // the copies issued at the top of this iteration have drained by now: recycle their buffers
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      if (prev_out) zero_lines(eI, ob ^ 1);
    }
    __syncthreads();  // A: lists of column x complete, recycled buffers zero
    prev_out = !pre;
    if (pre) {
      const int t = eL; eL = eR; eR = t;  // nothing to write out: the old left line is still all zero
    } else {
      const int t = eL; eL = eR; eR = eI; eI = t;
    }
  }
  // last column of the chunk: write it out (including the right line at the mesh boundary)
  if (is_aux && prev_out) {
    // after the final rotation: previous left = eI, previous right = eL
    write_out(x1 - 1, eI, eL);
  }
  if (is_aux) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
