// gm_cart_kernel.cuh -- the sweep kernel of the structured path (see gm_cart.cuh for the design).
//
// CTA = 3*TH/2 matrix warps + 2 aux warps.  Per column x of the strip:
//   matrix : state phase (fields at the quadrature points, S, u.grad(phi)) for the TH(+1 halo)
//            cells, then two colour phases (even / odd cell rows); lane = (major node, minor node)
//            pair with its reference-element tables in registers; 9 values (+ constant P-block
//            entries) are accumulated into the shared-memory lists without branches (entries of
//            Dirichlet minors go to a trash element at the end of each slot)
//   aux    : write-out of the PREVIOUS column -- one TMA bulk copy (cp.async.bulk shared->global)
//            per completed list -- overlapped with the matrix warps' state phase, then the
//            residual gather of this column overlapped with the matrix phases, then recycling
//            (zeroing) of the line buffers whose copies have drained
// Three even-line and two odd-line buffers rotate.  Global loads of the next column (dof ids two
// columns ahead, iterate / list bases / viscosity one column ahead) are issued before the matrix
// phases and parked in registers.
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
  static constexpr int NNODE = 3 * (2 * TH + 1);
  static constexpr int NDESC = NNODE * 3 + TH * 3;
  double even[3][EVEN_SZ];
  double odd[2][ODD_SZ];
  double uni[NC][4][CT_UNI];
  double nod[NC][4][9][CT_NOD];
  double val[2][NC][30];
  double visc[2][NC][4][2];
  double tphi[4][9], tdx[4][9], tdy[4][9], tw[4], twpsi[4][3], tpsi[4][3];
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
  // ---- residual item of this aux lane: node (ixn, jp) of the column's 3 x (2TH+1) node grid (registers)
  const int r_ixn = at >= 0 ? at / (2 * TH + 1) : 0, r_jp = at >= 0 ? at - r_ixn * (2 * TH + 1) : 0;
  const bool r_valid = is_aux && at < S::NNODE && r_jp >= jlo && r_jp <= 2 * th;
  int r_cy0 = -1, r_cy1 = -1, r_nd0 = 0, r_nd1 = 0;
  if (r_jp & 1) {
    r_cy0 = r_jp >> 1; r_nd0 = c_ct_node[1][r_ixn < 3 ? r_ixn : 0];
  } else {
    const int c0 = r_jp >> 1, c1 = (r_jp >> 1) - 1;  // node is the bottom of c0 and the top of c1
    if (c0 < ncol_cells) { r_cy0 = c0; r_nd0 = c_ct_node[0][r_ixn < 3 ? r_ixn : 0]; }
    if (c1 >= 0 && c1 < th) { r_cy1 = c1; r_nd1 = c_ct_node[2][r_ixn < 3 ? r_ixn : 0]; }
  }
  if (r_cy0 >= ncol_cells) r_cy0 = -1;
  const int r_dst = (r_ixn == 1) ? (r_jp >> 1) * CT_ODD_STRIDE + ((r_jp & 1) ? CT_EDG_BLK + CT_CEN_RES : CT_EDG_RES)
                                 : (r_jp >> 1) * CT_EVEN_STRIDE + ((r_jp & 1) ? CT_VTX_BLK + CT_EDG_RES : CT_VTX_RES);

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

  // ---- global-load helpers (one item = one (cell, local dof) of a column)
  const int pf_cy = tid / 30, pf_l = tid - pf_cy * 30;
  const bool pf_on = tid < ncol_cells * 30;
  auto load_gd = [&](int x) -> int32_t {
    return a.gdofs[((int64_t)(y0 + pf_cy) * a.nx + x) * 30 + pf_l];
  };
  const int xs = x0 > 0 ? x0 - 1 : x0;
  if (pf_on) {
    s.gd[xs % 3][pf_cy][pf_l] = load_gd(xs);
    if (xs + 1 < x1) s.gd[(xs + 1) % 3][pf_cy][pf_l] = load_gd(xs + 1);
  }
  __syncthreads();
  auto fetch_col = [&](int x, double &v, long long &lb, int &ll, double &vm, double &vk, int &cl) {
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

  // ---- one third of a cell matrix: 9 pair values + constant P-block entries -> shared lists
  auto matrix_task = [&](int cy, int sb, char *lineb, char *pblk, bool with_p) {
    // class / offsets (warp uniform branch, taken only when the class changes)
    const int cl = s.cls[sb][cy];
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
          o[kk * 3 + ll] = (lo[kk] + (r == 0xFF ? (kk == 2 ? capT : capU) - 2 : r)) * 8;
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
    const double2 *u2 = reinterpret_cast<const double2 *>(&s.uni[cy][0][0]);
    const double *ni = &s.nod[cy][0][it][0];
    const double *nj = &s.nod[cy][0][jt][0];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const double2 a0 = u2[q * (CT_UNI / 2) + 0], a1 = u2[q * (CT_UNI / 2) + 1], a2 = u2[q * (CT_UNI / 2) + 2],
                    a3 = u2[q * (CT_UNI / 2) + 3];  // (mu,G00) (G01,G10) (G11,dT0) (dT1,kap)
      const double2 si = *reinterpret_cast<const double2 *>(ni + q * 9 * CT_NOD);
      const double2 sj = *reinterpret_cast<const double2 *>(nj + q * 9 * CT_NOD + 2);
      const double cj = nj[q * 9 * CT_NOD + 4];
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
    // block of the major node, list bases shifted by the parity of their global base (TMA alignment)
    char *blk = lineb + blk0_b + cy * stride_b;
    const int *lbw = reinterpret_cast<const int *>(&s.lbase[sb][cy][0]);
    char *b0 = blk + ((lbw[2 * M] & 1) << 3);
    char *b1 = blk + ((lbw[2 * (9 + M)] & 1) << 3);
    char *b2 = blk + ((lbw[2 * (21 + M)] & 1) << 3);
    if (CSR) {
      ct_rmw(b0, o[0], v00); ct_rmw(b0, o[1], v01); ct_rmw(b0, o[2], TUT0);
      ct_rmw(b1, o[3], v10); ct_rmw(b1, o[4], v11); ct_rmw(b1, o[5], TUT1);
      ct_rmw(b2, o[6], tu0); ct_rmw(b2, o[7], tu1); ct_rmw(b2, o[8], tt_v);
    } else {
      ct_rmw(b0, o[0], v00); ct_rmw(b0, o[1], v10); ct_rmw(b0, o[2], tu0);
      ct_rmw(b1, o[3], v01); ct_rmw(b1, o[4], v11); ct_rmw(b1, o[5], tu1);
      ct_rmw(b2, o[6], TUT0); ct_rmw(b2, o[7], TUT1); ct_rmw(b2, o[8], tt_v);
    }
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
    const int sb = x & 1;        // state / value buffer of this column
    const int ob = x & 1;        // odd (mid) line buffer of this column
    double nv = 0, nvm = 1, nvk = 0; long long nlb = 0; int nll = 0, ncl = 0;
    int32_t ngd = 0;
    const bool have_next = x + 1 < x1;
    if (!is_aux) {
      // =============================== state phase (matrix warps) ===============================
      // S2: fields at the quadrature points.  item = (cell, q, group): 0: u0,G00,G10 | 1: u1,G01,G11 | 2: T,dT0,dT1 | 3: p, mu, kappa
      for (int k = tid; k < ncol_cells * 16; k += NMT) {
        const int cy = k >> 4, q = (k >> 2) & 3, g = k & 3;
        double *u = s.uni[cy][q];
        if (g == 3) {
          double v = 0;
#pragma unroll
          for (int m = 0; m < 3; ++m) v += s.tpsi[q][m] * s.val[sb][cy][18 + m];
          u[U_P] = v;
          u[U_MU] = s.visc[sb][cy][q][0];
          u[U_KAP] = s.visc[sb][cy][q][1];
        } else {
          const double *vv = &s.val[sb][cy][g == 0 ? 0 : (g == 1 ? 9 : 21)];
          double f0 = 0, f1 = 0, f2 = 0;
#pragma unroll
          for (int i = 0; i < 9; ++i) {
            const double t = vv[i];
            f0 = fma(s.tphi[q][i], t, f0); f1 = fma(s.tdx[q][i], t, f1); f2 = fma(s.tdy[q][i], t, f2);
          }
          if (g == 0) { u[U_U0] = f0; u[U_G00] = f1; u[U_G10] = f2; }
          else if (g == 1) { u[U_U1] = f0; u[U_G01] = f1; u[U_G11] = f2; }
          else { u[U_T] = f0; u[U_DT0] = f1; u[U_DT1] = f2; }
        }
      }
      __syncwarp();
      asm volatile("bar.sync 2, %0;" ::"n"(NMT) : "memory");  // matrix warps only
      // S4: per-node state; the node-0 item also derives the remaining uniform values
      for (int k = tid; k < ncol_cells * 36; k += NMT) {
        const int cy = k / 36, r = k - cy * 36, q = r / 9, i = r - q * 9;
        double *u = s.uni[cy][q];
        const double u0 = u[U_U0], u1 = u[U_U1], G00 = u[U_G00], G01 = u[U_G01], G10 = u[U_G10], G11 = u[U_G11];
        const double dx = s.tdx[q][i], dy = s.tdy[q][i];
        const double D01 = 0.5 * (G01 + G10);
        const double S0 = G00 * dx + D01 * dy;
        const double S1 = D01 * dx + G11 * dy;
        const double sc = a.js * s.tw[q] * 2.0 * a.Pr * u[U_KAP];
        double *op = s.nod[cy][q][i];
        op[0] = S0; op[1] = S1; op[2] = sc * S0; op[3] = sc * S1;
        op[4] = u0 * dx + u1 * dy;
        if (i == 0) {
          u[U_CONV0] = u0 * G00 + u1 * G10;
          u[U_CONV1] = u0 * G01 + u1 * G11;
          u[U_UDT] = u0 * u[U_DT0] + u1 * u[U_DT1];
          u[U_DIVU] = G00 + G11;
        }
      }
    } else if (prev_out) {
      // ====== write-out of the previous column (aux warps), overlapped with the state phase ======
      write_out(x - 1, eI, eR);  // previous left line = this column's idle buffer
    }
    __syncwarp();     // (lanes of a warp have different trip counts above: reconverge before the barrier)
    __syncthreads();  // B: state of column x ready
    if (!is_aux) {
      // ---- issue the global loads of the next columns (parked in registers over the matrix phases)
      if (have_next) {
        fetch_col(x + 1, nv, nlb, nll, nvm, nvk, ncl);
        if (x + 2 < x1 && pf_on) ngd = load_gd(x + 2);
      }
      if (jac) {
        // =============================== matrix phases ===============================
        char *lineb = reinterpret_cast<char *>(ixM == 1 ? s.odd[ob] : s.even[ixM == 0 ? eL : eR]);
        char *pblk = reinterpret_cast<char *>(s.odd[ob] + S::ODD_P0);
        const bool lane_go = !(pre && ixM != 2);  // halo column: only the right line is owned
        if (lane_go) {
          if (doA) matrix_task(cyA, sb, lineb, pblk, tt == 1 && !pre);
          if (doH) matrix_task(th, sb, lineb, pblk, false);
        }
        __syncwarp();
        asm volatile("bar.sync 1, %0;" ::"n"(NMT) : "memory");  // matrix warps only: even rows done
        if (lane_go && doB) matrix_task(cyB, sb, lineb, pblk, tt == 1 && !pre);
      }
      // park the prefetched column in shared memory
      if (have_next) {
        store_col(x + 1, nv, nlb, nll, nvm, nvk, ncl);
        if (x + 2 < x1 && pf_on) s.gd[(x + 2) % 3][pf_cy][pf_l] = ngd;
      }
    } else {
      // =========================== residual gather (aux warps) ===========================
      if (res) {
        if (r_valid && !(pre && r_ixn != 2)) {
          double r0 = 0, r1 = 0, r2 = 0;
#pragma unroll
          for (int side = 0; side < 2; ++side) {
            const int cy = side ? r_cy1 : r_cy0, i = side ? r_nd1 : r_nd0;
            if (cy >= 0) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const double *u = s.uni[cy][q];
                const double *nd = s.nod[cy][q][i];
                const double w = s.tw[q];
                const double ph = w * s.tphi[q][i], dx = w * s.tdx[q][i], dy = w * s.tdy[q][i];
                const double m2 = 2.0 * a.Pr * w * u[U_MU], bu = a.RaPr * u[U_T];
                r0 = fma(m2, nd[0], r0); r0 = fma(ph, u[U_CONV0] - bu * a.gx, r0); r0 = fma(-dx, u[U_P], r0);
                r1 = fma(m2, nd[1], r1); r1 = fma(ph, u[U_CONV1] - bu * a.gy, r1); r1 = fma(-dy, u[U_P], r1);
                r2 = fma(u[U_DT0], dx, r2); r2 = fma(u[U_DT1], dy, r2); r2 = fma(u[U_UDT], ph, r2);
              }
            }
          }
          double *dst = ((r_ixn == 1) ? s.odd[ob] : s.even[r_ixn == 0 ? eL : eR]) + r_dst;
          dst[0] += r0; dst[1] += r1; dst[2] += r2;
        }
        if (!pre)
          for (int k = at; k < th * 3; k += 64) {
            const int cy = k / 3, p = k - cy * 3;
            double acc = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) acc += s.twpsi[q][p] * s.uni[cy][q][U_DIVU];
            s.odd[ob][S::ODD_P0 + cy * CT_P_BLK + CT_P_RES + p] = acc;
          }
      }
      // the copies issued at the top of this iteration have drained by now: recycle their buffers
      // (wait_group is per thread: every aux thread waits for its own copies, then all 64 meet)
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
      asm volatile("bar.sync 3, 64;" ::: "memory");
      if (prev_out) zero_lines(eI, ob ^ 1);
    }
    __syncwarp();
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
    write_out(x1 - 1, eI, eL);  // after the final rotation: previous left = eI, previous right = eL
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}
