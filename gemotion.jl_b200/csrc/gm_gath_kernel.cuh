// gm_gath_kernel.cuh -- structured path, second generation: register gather ("owner lane computes its
// nnz entries completely").  Same reference interface as gm_cart_kernel.cuh (jacobian!/residual! of the
// operator built at /root/reference/src/Solver.jl:157, integrands Solver.jl:111-152, arithmetic SURVEY A.4).
//
// k_cart (gm_cart_kernel.cuh) accumulates the 837 entries of every cell matrix into shared-memory lists by
// read-modify-write; ncu shows it bound by shared-memory wavefronts (855 per cell).  Here a lane owns an
// OUTPUT node pair (A = list owner, B = neighbour) and sums the contributions of the 1, 2 or 4 cells that
// contain both nodes in registers; every list position is then stored to shared memory exactly once (no
// read-modify-write, no zeroing) and each completed list leaves with one TMA bulk copy.
//
//  * Mirror trick: on a Cartesian mesh the reference-element products of the local pair (i,j) of the cell
//    left of / below a node equal those of the mirrored pair of the cell right of / above it, with the
//    quadrature points permuted and sign flips on odd x/y derivative counts.  The flips are moved into the
//    per-cell STATE, of which four variants (I, X, Y, XY) are kept in shared memory: variant m stores at
//    slot q the fields of the actual point q^m, with G01, G10 and the coefficients C1, C2 multiplied by
//    sx*sy and (u0,u1) by (sx,sy).  A lane therefore keeps ONE set of 28 table values in registers and
//    evaluates any of its cells with the same instruction stream; only v01, v10 need the sign sx*sy when
//    they are merged.
//  * Six coefficients: the UU block is sum_q sum_kl C^ab_kl(q) (d_k phi_i d_l phi_j)(q) and the C^ab_kl reduce to
//    six numbers per quadrature point (see the state record below), computed once per cell with the state.
//  * Work unit: a "quad" = the four nodes (V vertex, H bottom-edge, E left-edge, C centre) at the lower left
//    of a cell.  Four vertically adjacent quads = 64 output node pairs x 9 entries = 81 cell-pair
//    evaluations per cell = 81 lanes x 4 rounds: V lanes (36) evaluate their 4 cells, E / H lanes (18 + 18)
//    two quads x two cells, C lanes (9) four quads x one cell.  A trio of warps (27 lanes each) runs 4
//    identical rounds per column; consecutive rounds that hit the same output are merged before the store.
//  * CTA = NG trios + GA_NAUX aux warps (2 + 6 = 12 warps at 168 registers, 1 CTA/SM), sweeping a strip of
//    4*NG quad rows column by column in two role-specific loops that meet at one named barrier per column.
//    Aux warps: dof ids / iterate values / rank classes of the coming columns and list bases of column X+1
//    (cp.async), state of column X+1 (four variants), the constant P-block entries, the residual (written
//    straight to b), TMA write-out of column X-1.  Measurements and the experiments that did not help are in
//    DESIGN.md section 4.
//
// The file compiles for the device (nvcc, sm_100a) and -- with GM_EMU -- as host C++ under
// tests/emu/cuda_emu.h, which runs every CUDA thread as a pthread; that build exists only so that the
// index logic can be checked against the oracle without a GPU (tests/test_gath_emu.py).
#pragma once

#define GA_NAUX 6                  // aux warps
#define GA_MAXREG 168               // 65536 / (12 warps * 32): warps are allocated in fours
#define GA_QS 16                   // doubles per (variant, point) state record
#define GA_VS 72                   // doubles per variant: 4 records + 8 (p, T of the 4 points; variant 0 only); 64 B bank shift
#define GA_CS (4 * GA_VS + 2)      // doubles per cell (290: consecutive cell rows are shifted by 16 B in the banks)
#define GA_QSTRIDE 732             // staged doubles per quad: 15 list slots (728, gm_gath_slot_size) + bank shift

// State record of one quadrature point (doubles).  The UU block of a node pair is
//   K_ab = sum_q sum_kl C^ab_kl(q) (d_k phi_i d_l phi_j)(q) + ...   (viscous part + power-law tangent, Solver.jl:131-134)
// and the 16 coefficients C^ab_kl = mu' V^ab_kl + sc D_ak D_bl reduce to six per point (mu' = js w Pr mu, sc = js w 2 Pr kappa):
//   A0 = 2 mu' + sc D00^2, A1 = 2 mu' + sc D11^2, BM = mu' + sc D01^2, AA = sc D00 D11, C1 = sc D00 D01, C2 = sc D01 D11
// so a lane needs only its four gradient products per point (registers) and no per-lane S vectors.
#define GS_A0 0
#define GS_A1 1
#define GS_BM 2
#define GS_AA 3
#define GS_C1 4
#define GS_C2 5
#define GS_G00 6
#define GS_G11 7
#define GS_G01 8
#define GS_G10 9
#define GS_DT0 10
#define GS_DT1 11
#define GS_U0 12
#define GS_U1 13
#define GS_MUR 14  // plain mu (residual, variant 0)

#ifdef GM_EMU
#define GA_LDG(p) (*(p))
#else
#define GA_LDG(p) __ldg(p)
__device__ __forceinline__ void ga_cp_async(void *sdst, const void *gsrc, int bytes);
#endif

// ---- asynchronous copies (device: LDGSTS / UBLKCP; emulation: deferred memcpy, see cuda_emu.h)
#ifndef GM_EMU
__device__ __forceinline__ void ga_cp4(void *sdst, const void *gsrc) {
  const uint32_t sa = (uint32_t)__cvta_generic_to_shared(sdst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void ga_cp8(void *sdst, const void *gsrc) {
  const uint32_t sa = (uint32_t)__cvta_generic_to_shared(sdst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void ga_cp16(void *sdst, const void *gsrc) {
  const uint32_t sa = (uint32_t)__cvta_generic_to_shared(sdst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void ga_cp_wait_all() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void ga_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void ga_cp_wait_prev() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }  // all but the newest group
__device__ __forceinline__ void ga_bulk_store(double *gdst, const double *ssrc, int bytes) {
  const uint32_t sa = (uint32_t)__cvta_generic_to_shared(ssrc);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(sa), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ga_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void ga_bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void ga_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ double ga_shfl_xor(double v, int mask) { return __shfl_xor_sync(0xffffffffu, v, mask); }
template <int NT>
__device__ __forceinline__ void ga_bar_step() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }  // both role loops meet here
#endif

// local node index of the tensor position (ix, iy) of a Q2 cell (same table as c_ct_node)
__host__ __device__ __forceinline__ int ga_node(int ix, int iy) {
  return (0x352786140ull >> (4 * (iy * 3 + ix))) & 15;  // {0,4,1, 6,8,7, 2,5,3}
}

// Role of trio lane t (0..80): list-owner node type A (0 V, 1 H, 2 E, 3 C), neighbour position (bx, by) in
// the canonical (upper right) cell, and four rounds: quad of the group, state variant m = mx + 2 my (the
// evaluated cell is (X - mx, Y - my)), flush flag (the next round belongs to another output).
struct GaRole {
  int A, bx, by;
  int gq[4], m[4], flush[4];
};
__host__ __device__ inline GaRole ga_role(int t) {
  GaRole r;
  int j;
  if (t < 36) {            // V: one quad, its four cells
    r.A = 0; j = t % 9;
    const int g = t / 9;
    r.bx = j % 3; r.by = j / 3;
    const bool yfirst = r.bx > 0 && r.by == 0;  // outputs (I,Y) and (X,XY) coincide: keep them adjacent
    for (int k = 0; k < 4; ++k) { r.gq[k] = g; r.m[k] = yfirst ? ((k & 1) * 2 + (k >> 1)) : k; }
  } else if (t < 72) {     // E (left edge): cells I, X;  H (bottom edge): cells I, Y;  two quads each
    const bool isE = t < 54;
    const int u = isE ? t - 36 : t - 54, h = u / 9;
    j = u % 9;
    r.A = isE ? 2 : 1;
    r.bx = j % 3; r.by = j / 3;
    for (int k = 0; k < 4; ++k) { r.gq[k] = 2 * h + (k >> 1); r.m[k] = (k & 1) ? (isE ? 1 : 2) : 0; }
  } else {                 // C: four quads, own cell
    j = t - 72;
    r.A = 3; r.bx = j % 3; r.by = j / 3;
    for (int k = 0; k < 4; ++k) { r.gq[k] = k; r.m[k] = 0; }
  }
  for (int k = 0; k < 4; ++k) {
    if (k == 3 || r.gq[k + 1] != r.gq[k]) { r.flush[k] = 1; continue; }
    const int m0 = r.m[k], m1 = r.m[k + 1];
    const int ox0 = (m0 & 1) ? -r.bx : r.bx, oy0 = (m0 & 2) ? -r.by : r.by;
    const int ox1 = (m1 & 1) ? -r.bx : r.bx, oy1 = (m1 & 2) ? -r.by : r.by;
    r.flush[k] = (ox0 != ox1 || oy0 != oy1) ? 1 : 0;
  }
  return r;
}
__host__ __device__ __forceinline__ int ga_ax(int A) { return (A == 1 || A == 3) ? 1 : 0; }
__host__ __device__ __forceinline__ int ga_ay(int A) { return (A >= 2) ? 1 : 0; }
// cells (variants) that contain node type A:  V: I X Y XY,  H: I Y,  E: I X,  C: I
__host__ __device__ __forceinline__ int ga_nvar(int A) { return A == 0 ? 4 : (A == 3 ? 1 : 2); }
__host__ __device__ __forceinline__ int ga_var(int A, int s) { return A == 0 ? s : (A == 1 ? 2 * s : (A == 2 ? s : 0)); }
// local node of node type A inside the cell of variant m
__host__ __device__ __forceinline__ int ga_anode(int A, int m) {
  const int ax = ga_ax(A), ay = ga_ay(A);
  return ga_node((m & 1) ? 2 - ax : ax, (m & 2) ? 2 - ay : ay);
}

template <int NG>
struct GaSmem {
  static constexpr int QH = 4 * NG, NR = QH + 1;
  double st[3][NR * GA_CS];          // state ring (cell column mod 3)
  double stage[2][QH * GA_QSTRIDE];  // staged lists of quad column X (X & 1)
  double val[4][NR][30];             // iterate values of cell column c (c & 3)
  double visc[4][NR][4][2];
  long long lbase[4][QH][16];        // list bases / ends of quad column X (X & 3); 15 used
  long long lnext[4][QH][16];
  double ltk[3 * NG * 32];           // per matrix lane: TK of its node pair
  double tphi[4][9], tdx[4][9], tdy[4][9], tw[4], twpsi[4][3], tpsi[4][3];
  double UP[18][3];
  int32_t gd[8][NR][30];             // dof ids of cell column c (c & 7)
  int32_t lg[4][QH][16];             // dof id of each list of quad column X (0: no list / not assembled here)
  int32_t cls[4][NR];
  int16_t slot[16];
  uint8_t par[4][QH][8];             // parities of the list bases of quad column X (X & 3): [node type | 4 = P] bits k
  uint8_t pu[54][4];                 // (A, m, a, p) of the 54 P-row entries of a quad's U lists
};

template <int NG, bool CSR, bool NEWT, int XCH>
__global__ void __maxnreg__(GA_MAXREG) k_gath(CartArgs a) {
  using S = GaSmem<NG>;
  constexpr int QH = S::QH, NR = S::NR;
  constexpr int NMW = 3 * NG, NMT = NMW * 32, NAT = GA_NAUX * 32, NT = NMT + NAT;
#ifdef GM_EMU
  unsigned char *smem_raw = emu_smem();
#else
  extern __shared__ __align__(16) unsigned char smem_raw[];
#endif
  S &s = *reinterpret_cast<S *>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool is_aux = warp >= NMW;
  const int at = tid - NMT;  // aux thread index
  const int Yq0 = a.row_begin + (int)blockIdx.y * QH;     // first quad row of the strip
  const int nqr = min(QH, a.row_end + 1 - Yq0);           // quad rows Yq0 .. Yq0+nqr-1 (row_end = the top mesh line)
  const int Yc0 = Yq0 - 1;                                // cell row of state row 0
  const int X0 = (int)blockIdx.x * XCH, X1 = min(a.nx + 1, X0 + XCH);  // quad columns of this x-chunk
  const bool jac = (a.flags & GM_JACOBIAN) != 0, res = (a.flags & GM_RESIDUAL) != 0;
  const CartCell &cc = *a.cc;

  // state rows (cell rows Yc0 + rr, rr = 0..nqr) that exist in the local table / are owned by this rank
  const int tab_lo = max(0, -Yc0), tab_n = min(nqr, a.ny - 1 - Yc0) - tab_lo;
  const int own_lo = max(0, a.row_begin - Yc0), own_n = min(nqr, a.row_end - 1 - Yc0) - own_lo;
  auto row_in_table = [&](int rr) { return (unsigned)(rr - tab_lo) <= (unsigned)tab_n && tab_n >= 0; };
  auto row_owned = [&](int rr) { return (unsigned)(rr - own_lo) <= (unsigned)own_n && own_n >= 0; };
  auto col_ok = [&](int c) { return c >= 0 && c >= X0 - 1 && c < a.nx; };  // cell columns this chunk works with

  // ---- constant tables
  for (int k = tid; k < 36; k += NT) { s.tphi[0][k] = cc.sphi[0][k]; s.tdx[0][k] = cc.sdx[0][k]; s.tdy[0][k] = cc.sdy[0][k]; }
  if (tid < 4) s.tw[tid] = cc.w[tid];
  if (tid < 12) { s.twpsi[0][tid] = cc.wpsi[0][tid]; s.tpsi[0][tid] = cc.psi[0][tid]; }
  for (int k = tid; k < 54; k += NT) s.UP[0][k] = cc.UP[0][k];
  if (tid == 0) {
    const int16_t sl[16] = {0, 90, 180, 258, 312, 366, 414, 468, 522, 570, 604, 638, 668, 688, 708, 728};
    for (int k = 0; k < 16; ++k) s.slot[k] = sl[k];
    int n = 0;
    for (int A = 0; A < 4; ++A)
      for (int sv = 0; sv < ga_nvar(A); ++sv)
        for (int c2 = 0; c2 < 2; ++c2)
          for (int p = 0; p < 3; ++p) {
            s.pu[n][0] = (uint8_t)A; s.pu[n][1] = (uint8_t)ga_var(A, sv); s.pu[n][2] = (uint8_t)c2; s.pu[n][3] = (uint8_t)p;
            ++n;
          }
  }

  __syncthreads();

  // =====================================================================================
  // column sweep (two role-specific loops, so that the registers of one role are not live in the other).  Step X (matrix: quad column X): aux prepares dof ids of X+3, values of X+2, state and
  // list records of X+1.  Four warm-up steps fill the pipeline.
  // =====================================================================================
  if (is_aux) {
  // =====================================================================================
  // aux tasks
  // =====================================================================================
  // Work split over the six aux warps (w0..w5; per-task warp-instruction counts from the ncu source view):
  //   dof ids / values : second pass on w3..w5      list records : w0, w1       state : velocity w2, w3; T w0, w1; p w4, w5
  //   P entries, write-out : w0, w1, w4, w5         residual : w2..w5
  static_assert(NAT == 192, "the aux task split below assumes six aux warps");
  const int at_hi = NAT - 1 - at;                              // item order for the two-pass loops
  const int at_o = at < 64 ? at : (at >= 128 ? at - 64 : -1);  // 0..127 on w0, w1, w4, w5
  // (A1/A2) every aux thread owns up to two fixed (state row, local dof) items of a cell column: the row / dof
  //         split and the row part of the addresses are loop invariant, a column only adds its offset
  static_assert(NR * 30 <= 2 * NAT, "two load items per aux thread");
  int li_off[2];          // item index rr*30 + l (-1: none)
  int64_t li_goff[2];     // offset of the item in gdofs for column 0
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int k = at_hi + j * NAT, rr = k / 30;
    const bool ok = k < NR * 30 && row_in_table(rr);
    li_off[j] = ok ? k : -1;
    li_goff[j] = ok ? (int64_t)(Yc0 + rr) * a.nx * 30 + (k - rr * 30) : 0;
  }
  // (A1) dof ids of cell column c
  auto issue_gd = [&](int c) {
    if (!col_ok(c)) return;
    int32_t *dst = &s.gd[c & 7][0][0];
    const int32_t *src = a.gdofs + (int64_t)c * 30;
#pragma unroll
    for (int j = 0; j < 2; ++j)
      if (li_off[j] >= 0) ga_cp4(dst + li_off[j], src + li_goff[j]);
  };
  // (A2) iterate values, viscosity, rank class of cell column c (its dof ids have landed)
  auto issue_val = [&](int c) {
    if (!col_ok(c)) return;
    const int32_t *gdc = &s.gd[c & 7][0][0];
    double *dst = &s.val[c & 3][0][0];
#pragma unroll
    for (int j = 0; j < 2; ++j)
      if (li_off[j] >= 0) {
        const int32_t d = gdc[li_off[j]];
        ga_cp8(dst + li_off[j], d > 0 ? a.x + (d - 1) : a.dvals + (-d - 1));
      }
    for (int k = at; k < NR * 4; k += NAT) {
      const int rr = k >> 2, q = k & 3;
      if (!row_in_table(rr)) continue;
      const int64_t cell = (int64_t)(Yc0 + rr) * a.nx + c;
      if (!NEWT) ga_cp16(&s.visc[c & 3][rr][q][0], a.visc + (cell * 4 + q) * 2);
      if (q == 0) ga_cp4(&s.cls[c & 3][rr], a.cls + cell);
    }
  };
  // (A3) state of cell column c: fields at the quadrature points, four mirror variants
  auto state = [&](int c) {
    if (!col_ok(c)) return;
    double *sc0 = s.st[c % 3];
    for (int k0 = at; k0 < 3 * 64; k0 += NAT) {  // g-major item order: a warp runs one kind of item;
      const int k = k0 < 64 ? k0 + 64 : (k0 < 128 ? k0 - 64 : k0);  // velocity items (0..63) on w2, w3
      const int g = k >> 6, e = k & 63, rr = e >> 2, q = e & 3;
      if (rr >= NR || !row_owned(rr)) continue;
      double *cs = sc0 + rr * GA_CS;
      const double *vv = &s.val[c & 3][rr][0];
      if (g == 0) {         // velocity
        double u0 = 0, u1 = 0, G00 = 0, G01 = 0, G10 = 0, G11 = 0;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
          const double ph = s.tphi[q][i], dx = s.tdx[q][i], dy = s.tdy[q][i], ux = vv[i], uy = vv[9 + i];
          u0 = fma(ph, ux, u0); G00 = fma(dx, ux, G00); G10 = fma(dy, ux, G10);
          u1 = fma(ph, uy, u1); G01 = fma(dx, uy, G01); G11 = fma(dy, uy, G11);
        }
        const double D01 = 0.5 * (G01 + G10);
        const double mu = NEWT ? 1.0 : s.visc[c & 3][rr][q][0];
        const double wj = a.js * s.tw[q] * a.Pr, mup = wj * mu, sc = NEWT ? 0.0 : 2.0 * wj * s.visc[c & 3][rr][q][1];
        const double s00 = sc * G00, s01 = sc * D01, s11 = sc * G11;
        const double A0 = fma(s00, G00, 2.0 * mup), A1 = fma(s11, G11, 2.0 * mup), BM = fma(s01, D01, mup);
        const double AA = s00 * G11, C1 = s00 * D01, C2 = s01 * G11;
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const double sx = (m & 1) ? -1.0 : 1.0, sy = (m & 2) ? -1.0 : 1.0, sxy = sx * sy;
          double2 *o = reinterpret_cast<double2 *>(cs + m * GA_VS + (q ^ m) * GA_QS);  // 16-byte stores: pairs as the lanes load them
          o[GS_A0 / 2] = make_double2(A0, A1); o[GS_BM / 2] = make_double2(BM, AA); o[GS_C1 / 2] = make_double2(sxy * C1, sxy * C2);
          o[GS_G00 / 2] = make_double2(G00, G11); o[GS_G01 / 2] = make_double2(sxy * G01, sxy * G10);
          o[GS_U0 / 2] = make_double2(sx * u0, sy * u1);
        }
        cs[q * GA_QS + GS_MUR] = mu;
      } else if (g == 1) {  // temperature
        double T = 0, d0 = 0, d1 = 0;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
          const double tv = vv[21 + i];
          T = fma(s.tphi[q][i], tv, T); d0 = fma(s.tdx[q][i], tv, d0); d1 = fma(s.tdy[q][i], tv, d1);
        }
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          reinterpret_cast<double2 *>(cs + m * GA_VS + (q ^ m) * GA_QS)[GS_DT0 / 2] = make_double2(d0, d1);
        }
        cs[4 * GA_QS + 2 * q + 1] = T;
      } else {              // pressure
        double p = 0;
#pragma unroll
        for (int m = 0; m < 3; ++m) p += s.tpsi[q][m] * vv[18 + m];
        cs[4 * GA_QS + 2 * q] = p;
      }
    }
  };
  // (A4) list records of quad column c: one thread per (quad row, node type | P lists) = three lists: dof ids,
  //      then (cp.async) list bases and ends; lists_parity() packs the base parities once they have landed
  auto issue_lists = [&](int c) {
    if (at >= QH * 5) return;
    const int rowq = at / 5, grp = at - rowq * 5;
    int32_t g[3] = {0, 0, 0};
    bool active = false;
    if (rowq < nqr) {
      if (grp < 4) {
        bool found = false;
        for (int sv = 0; sv < ga_nvar(grp); ++sv) {
          const int m = ga_var(grp, sv), cx = c - (m & 1), rr = rowq + 1 - (m >> 1);
          if (!col_ok(cx) || !row_in_table(rr)) continue;
          if (row_owned(rr)) active = true;
          if (!found) {
            found = true;
            const int nd = ga_anode(grp, m);
            const int32_t *gp = &s.gd[cx & 7][rr][0];
            g[0] = gp[nd]; g[1] = gp[9 + nd]; g[2] = gp[21 + nd];
          }
        }
      } else if (col_ok(c) && row_owned(rowq + 1)) {
        active = true;
        const int32_t *gp = &s.gd[c & 7][rowq + 1][18];
        g[0] = gp[0]; g[1] = gp[1]; g[2] = gp[2];
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int li = grp * 3 + k;
      const int32_t gk = (active && g[k] > 0) ? g[k] : 0;
      s.lg[c & 3][rowq][li] = gk;
      if (gk > 0) {
        ga_cp8(&s.lbase[c & 3][rowq][li], a.ptr + (gk - 1));
        ga_cp8(&s.lnext[c & 3][rowq][li], a.ptr + gk);
      } else {
        s.lbase[c & 3][rowq][li] = 0; s.lnext[c & 3][rowq][li] = 0;
      }
    }
  };
  auto lists_parity = [&](int c) {  // after this thread's cp.async copies have landed
    if (at >= QH * 5) return;
    const int rowq = at / 5, grp = at - rowq * 5;
    const long long *lb = &s.lbase[c & 3][rowq][grp * 3];
    s.par[c & 3][rowq][grp] = (uint8_t)((lb[0] & 1) | ((lb[1] & 1) << 1) | ((lb[2] & 1) << 2));
  };
  // (A5) constant P-block entries of quad column X -> staged lists.  Aux thread e < 108 owns entry type e
  //      (everything but the rank class and the list parity is loop invariant) and walks the quad rows.
  int pe_li = 0, pe_rk = 0, pe_dx = 0, pe_dy = 0;
  double pe_val = 0.0;
  const int pe = at_o;  // entry type of this thread
  const bool pe_on = pe >= 0 && pe < 108, pe_plist = pe >= 54;
  if (pe_on) {
    if (pe < 54) {      // P rows (CSC) / P columns (CSR) inside the U lists of node A
      const int An = s.pu[pe][0], m = s.pu[pe][1], c2 = s.pu[pe][2], p = s.pu[pe][3];
      const int nd = ga_anode(An, m);
      pe_dx = m & 1; pe_dy = m >> 1;
      pe_li = An * 3 + c2;
      pe_rk = (18 + p) * 30 + c2 * 9 + nd;
      pe_val = (CSR ? 1.0 : -1.0) * s.UP[c2 * 9 + nd][p];
    } else {            // the three P lists of the cell of this quad
      const int p = (pe - 54) / 18, ud = (pe - 54) - p * 18;
      pe_li = 12 + p;
      pe_rk = ud * 30 + 18 + p;
      pe_val = (CSR ? -1.0 : 1.0) * s.UP[ud][p];
    }
  }
  // rank of this thread's entry type in one (interior) class, kept in a register: the table is only read for other classes
  const int pe_cls0 = pe_on ? a.cls[(int64_t)min(max(Yc0 + 3, 0), a.ny - 1) * a.nx + min(X0 + 3, a.nx - 1)] : 0;
  const int pe_rk0 = pe_on ? (int)GA_LDG(a.ctab + (int64_t)pe_cls0 * 900 + pe_rk) : 0xFF;
  auto p_entries = [&](int X) {
    if (!pe_on) return;
    const int cx = X - pe_dx;
    if (!col_ok(cx)) return;
    double *stg0 = s.stage[X & 1] + s.slot[pe_li];
    int rk[QH];
    bool own[QH];
#pragma unroll
    for (int rowq = 0; rowq < QH; ++rowq) {  // all rank loads in flight together
      const int rr = rowq + 1 - pe_dy;
      own[rowq] = row_owned(rr);
      const bool ok = rowq < nqr && row_in_table(rr) && (own[rowq] || !pe_plist);
      const int cl = ok ? s.cls[cx & 3][rr] : pe_cls0;
      rk[rowq] = !ok ? 0xFF : (cl == pe_cls0 ? pe_rk0 : (int)GA_LDG(a.ctab + (int64_t)cl * 900 + pe_rk));
    }
#pragma unroll
    for (int rowq = 0; rowq < QH; ++rowq) {
      if (rk[rowq] == 0xFF) continue;
      const int par = (s.par[X & 3][rowq][pe_li / 3] >> (pe_li % 3)) & 1;
      stg0[rowq * GA_QSTRIDE + par + rk[rowq]] = own[rowq] ? pe_val : 0.0;
    }
  };
  // (A6) residual of the dofs of quad column X: lane quartet = the (up to) four cells of one node; partial sums
  //      are combined by two butterfly shuffles, lane 0 of the quartet stores the three dofs of the node
  auto residual = [&](int X) {
    for (int k0 = 0; k0 < QH * 16; k0 += NAT) {  // (NAT is a multiple of 32: whole warps enter together)
      const int k = k0 + at - 64;  // (w2..w5; negative on w0, w1: no item)
      const int rowq = k < 0 ? QH : (k >> 4), An = (k >> 2) & 3, sv = k & 3;
      double r0 = 0, r1 = 0, r2 = 0;
      if (rowq < nqr && sv < ga_nvar(An)) {
        const int m = ga_var(An, sv), cx = X - (m & 1), rr = rowq + 1 - (m >> 1);
        if (col_ok(cx) && row_owned(rr)) {
          const int nd = ga_anode(An, m);
          const double *cs = s.st[cx % 3] + rr * GA_CS;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const double *u = cs + q * GA_QS;
            const double w = s.tw[q], gx_ = s.tdx[q][nd], gy_ = s.tdy[q][nd];
            const double ph = w * s.tphi[q][nd], dx = w * gx_, dy = w * gy_;
            const double u0 = u[GS_U0], u1 = u[GS_U1], G00 = u[GS_G00], G01 = u[GS_G01], G10 = u[GS_G10], G11 = u[GS_G11];
            const double D01 = 0.5 * (G01 + G10), pq = cs[4 * GA_QS + 2 * q], Tq = cs[4 * GA_QS + 2 * q + 1];
            const double m2 = 2.0 * a.Pr * w * u[GS_MUR], bu = a.RaPr * Tq;
            r0 = fma(m2, fma(G00, gx_, D01 * gy_), r0); r0 = fma(ph, fma(u0, G00, u1 * G10) - bu * a.gx, r0); r0 = fma(-dx, pq, r0);
            r1 = fma(m2, fma(D01, gx_, G11 * gy_), r1); r1 = fma(ph, fma(u0, G01, u1 * G11) - bu * a.gy, r1); r1 = fma(-dy, pq, r1);
            r2 = fma(u[GS_DT0], dx, r2); r2 = fma(u[GS_DT1], dy, r2); r2 = fma(fma(u0, u[GS_DT0], u1 * u[GS_DT1]), ph, r2);
          }
        }
      }
      r0 += ga_shfl_xor(r0, 1); r1 += ga_shfl_xor(r1, 1); r2 += ga_shfl_xor(r2, 1);
      r0 += ga_shfl_xor(r0, 2); r1 += ga_shfl_xor(r1, 2); r2 += ga_shfl_xor(r2, 2);
      if (sv == 0 && rowq < nqr) {
        const int32_t *lg = &s.lg[X & 3][rowq][An * 3];
        if (lg[0] > 0) a.b[lg[0] - 1] = r0;
        if (lg[1] > 0) a.b[lg[1] - 1] = r1;
        if (lg[2] > 0) a.b[lg[2] - 1] = r2;
      }
    }
    for (int k = at - 64; k >= 0 && k < QH * 3; k += NAT) {  // pressure rows: cell-local
      const int rowq = k / 3, p = k - rowq * 3;
      if (rowq >= nqr) break;
      const int32_t g = s.lg[X & 3][rowq][12 + p];
      if (g <= 0) continue;
      const double *cs = s.st[X % 3] + (rowq + 1) * GA_CS;
      double v = 0;
#pragma unroll
      for (int q = 0; q < 4; ++q) v += s.twpsi[q][p] * (cs[q * GA_QS + GS_G00] + cs[q * GA_QS + GS_G11]);
      a.b[g - 1] = v;
    }
  };
  // (A7) write-out of quad column X: one TMA bulk copy per completed list
  auto write_out = [&](int X) {
    ga_fence_async();
    const double *stg0 = s.stage[X & 1];
    for (int k = at_o; k >= 0 && k < QH * 15; k += NAT) {
      const int rowq = k / 15, li = k - rowq * 15;
      if (rowq >= nqr) break;
      if (s.lg[X & 3][rowq][li] <= 0) continue;
      const long long base = s.lbase[X & 3][rowq][li];
      const int len = (int)(s.lnext[X & 3][rowq][li] - base);
      const int par = (int)(base & 1);
      const double *src = stg0 + rowq * GA_QSTRIDE + s.slot[li] + par;  // element k of the list lives at src[k]
      double *dst = a.nz + base;
      int k0 = 0, n = len;
      if (par) { dst[0] = src[0]; k0 = 1; n -= 1; }
      if (n & 1) { dst[k0 + n - 1] = src[k0 + n - 1]; n -= 1; }
      if (n > 0) ga_bulk_store(dst + k0, src + k0, n * 8);
    }
    ga_bulk_commit();
  };

    for (int X = X0 - 4; X < X1; ++X) {
      const bool live = X >= X0;
      if (live && X > X0 && jac) write_out(X - 1);
      issue_gd(X + 3);
      issue_val(X + 2);
      if (X + 1 >= X0 && X + 1 < X1) issue_lists(X + 1);
      if (live && jac) p_entries(X);
      if (live && res) residual(X);
      state(X + 1);
      ga_cp_wait_all();
      if (X + 1 >= X0 && X + 1 < X1) lists_parity(X + 1);
      ga_bulk_wait_read();
      ga_fence_async();
      ga_bar_step<NT>();  // lists of column X staged, state of X+1 ready, asynchronous copies landed
    }
    if (jac) {  // last column of the chunk
      write_out(X1 - 1);
      ga_bulk_wait_read();
    }
  } else {
  // ---- role of a matrix lane
  const int grp = is_aux ? 0 : warp / 3, wl = warp % 3;
  const bool lane_on = !is_aux && lane < 27;
  const GaRole role = ga_role(lane_on ? wl * 27 + lane : 0);
  const int A = role.A;
  uint32_t R[4];  // per round: [0:3] state row, [4:5] variant, [6] flush, [7:10] quad row, [11:17] pair index, [18] in table, [19] owned
  {
    const int ax = ga_ax(A), ay = ga_ay(A);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int m = role.m[r], mx = m & 1, my = m >> 1;
      const int rowq = grp * 4 + role.gq[r], rr = rowq + 1 - my;
      const bool rq_ok = lane_on && rowq < nqr;
      const int Mact = ga_node(mx ? 2 - ax : ax, my ? 2 - ay : ay);
      const int mact = ga_node(mx ? 2 - role.bx : role.bx, my ? 2 - role.by : role.by);
      R[r] = (uint32_t)rr | ((uint32_t)m << 4) | ((uint32_t)role.flush[r] << 6) | ((uint32_t)rowq << 7) |
             ((uint32_t)(Mact * 9 + mact) << 11) | ((rq_ok && row_in_table(rr)) ? (1u << 18) : 0u) |
             ((rq_ok && row_owned(rr)) ? (1u << 19) : 0u);
    }
  }
  // reference-element products of the canonical pair (test node it, trial node jt)
  const int nA = ga_node(ga_ax(A), ga_ay(A)), nB = ga_node(role.bx, role.by);
  const int it = CSR ? nA : nB, jt = CSR ? nB : nA;
  // 28 doubles in registers: the four gradient products, w phi_i phi_j, w phi_i grad phi_j, and the constant TK
  double PXX[4], PYY[4], PXY[4], PYX[4], TM[4], EX[4], EY[4];  // (TK is used once per flush: parked in shared memory)
  {
    const CartPairTab &t = a.tab[it * 9 + jt];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const double xi = cc.sdx[q][it], yi = cc.sdy[q][it], xj = cc.sdx[q][jt], yj = cc.sdy[q][jt];
      PXX[q] = xi * xj; PYY[q] = yi * yj; PXY[q] = xi * yj; PYX[q] = yi * xj;
      TM[q] = t.TM[q]; EX[q] = t.TP[q] * xj; EY[q] = t.TP[q] * yj;
    }
    s.ltk[tid] = t.TK;
  }
    const int slotpk = (A == 0 ? 0 : (A == 1 ? 258 : (A == 2 ? 414 : 570))) | ((A == 0 ? 90 : (A == 3 ? 34 : 54)) << 16);
    int xm3 = X0 % 3;
    for (int X = X0 - 4; X < X1; ++X) {
      if (X >= X0 && jac) {
      const double *stc0 = s.st[xm3], *stc1 = s.st[xm3 == 0 ? 2 : xm3 - 1];  // cell columns X and X-1
      double *stg0 = s.stage[X & 1];
      double v00 = 0, v01 = 0, v10 = 0, v11 = 0, tu0 = 0, tu1 = 0, tt = 0;
        int nev = 0;  // cells summed into the current output (constant parts are added at the flush)
      uint32_t rk0 = 0, rk1 = 0, rk2 = 0;
      bool have = false;
      // rank record of a round: loaded one round ahead (the first one before round 0)
      auto load_rec = [&](int r) -> uint4 {
        const uint32_t info = R[r];
        const int cx = X - (int)((info >> 4) & 1);
        uint4 v = {0, 0, 0, 0};
        if (cx >= 0 && cx < a.nx && (info & (1u << 18)))
          v = GA_LDG(reinterpret_cast<const uint4 *>(a.ptab) + ((int64_t)s.cls[cx & 3][info & 15] * 81 + ((info >> 11) & 127)));
        return v;
      };
      uint4 rec_nxt = load_rec(0);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const uint32_t info = R[r];
        const int rr = info & 15, m = (info >> 4) & 3, mx = m & 1;
        const int cx = X - mx;
        const bool cok = cx >= 0 && cx < a.nx;
        const uint4 rec_cur = rec_nxt;
        if (r < 3) rec_nxt = load_rec(r + 1);
        if (cok && (info & (1u << 18))) {
          rk0 = rec_cur.x; rk1 = rec_cur.y; rk2 = rec_cur.z;
          have = true;
        }
        if (cok && (info & (1u << 19))) {
          const double2 *sp = reinterpret_cast<const double2 *>((mx ? stc1 : stc0) + rr * GA_CS + m * GA_VS);
          const double sg = (m == 1 || m == 2) ? -1.0 : 1.0;
          // two independent chains per output (points 0,1 and 2,3): 10 chains per evaluation
          double t00[2] = {0, 0}, t11[2] = {0, 0}, t01[2] = {0, 0}, t10[2] = {0, 0}, eq[2] = {0, 0};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int h = q >> 1;
            const double2 a0 = sp[q * 8 + 0], a1 = sp[q * 8 + 1], a2 = sp[q * 8 + 2], a3 = sp[q * 8 + 3], a4 = sp[q * 8 + 4],
                          a5 = sp[q * 8 + 5], a6 = sp[q * 8 + 6];  // (A0,A1) (BM,AA) (C1,C2) (G00,G11) (G01,G10) (dT0,dT1) (u0,u1)
            const double ps = PXY[q] + PYX[q], c1x = a2.x * PXX[q];
            t00[h] = fma(TM[q], a3.x, fma(a1.x, PYY[q], fma(a2.x, ps, fma(a0.x, PXX[q], t00[h]))));
            t11[h] = fma(TM[q], a3.y, fma(a0.y, PYY[q], fma(a2.y, ps, fma(a1.x, PXX[q], t11[h]))));
            t01[h] = fma(TM[q], a4.y, fma(a2.y, PYY[q], fma(a1.x, PYX[q], fma(a1.y, PXY[q], c1x + t01[h]))));  // (a=0,b=1): + w phi phi G[1][0]
            t10[h] = fma(TM[q], a4.x, fma(a2.y, PYY[q], fma(a1.y, PYX[q], fma(a1.x, PXY[q], c1x + t10[h]))));  // (a=1,b=0): + w phi phi G[0][1]
            eq[h] = fma(a6.y, EY[q], fma(a6.x, EX[q], eq[h]));  // w phi_i (u . grad phi_j)
            tu0 = fma(TM[q], a5.x, tu0);
            tu1 = fma(TM[q], a5.y, tu1);
          }
          const double w00 = t00[0] + t00[1], w11 = t11[0] + t11[1], w01 = t01[0] + t01[1], w10 = t10[0] + t10[1];
          const double e = eq[0] + eq[1];
          v00 += w00 + e; v11 += w11 + e;
          v01 = fma(sg, w01, v01); v10 = fma(sg, w10, v10);
          tt += e;
          ++nev;
        }
        if (info & (1u << 6)) {
          if (have) {
            const int rowq = (info >> 7) & 15;
            double *q0 = stg0 + rowq * GA_QSTRIDE;
            const uint32_t pm = s.par[X & 3][rowq][A];
            const int slot0 = slotpk & 0xFFFF, slen = slotpk >> 16;
            double *b0 = q0 + slot0 + (pm & 1), *b1 = q0 + slot0 + slen + ((pm >> 1) & 1), *b2 = q0 + slot0 + 2 * slen + ((pm >> 2) & 1);
            // constant parts: nev * (TK ; buoyancy -Ra Pr g_a sum_q w phi_i phi_j)
            const double cn = (double)nev, mb = cn * ((TM[0] + TM[1]) + (TM[2] + TM[3]));
            const double ut0 = -a.RaPr * a.gx * mb, ut1 = -a.RaPr * a.gy * mb;
            tt = fma(cn, s.ltk[tid], tt);
            double e00, e01, e02, e10, e11, e12, e20, e21, e22;  // [list k][minor dof l]
            if (CSR) { e00 = v00; e01 = v01; e02 = ut0; e10 = v10; e11 = v11; e12 = ut1; e20 = tu0; e21 = tu1; e22 = tt; }
            else     { e00 = v00; e01 = v10; e02 = tu0; e10 = v01; e11 = v11; e12 = tu1; e20 = ut0; e21 = ut1; e22 = tt; }
            int rk;
            rk = rk0 & 255; b0[rk] = e00;
            rk = (rk0 >> 8) & 255; b0[rk] = e01;
            rk = (rk0 >> 16) & 255; b0[rk] = e02;
            rk = rk0 >> 24; b1[rk] = e10;
            rk = rk1 & 255; b1[rk] = e11;
            rk = (rk1 >> 8) & 255; b1[rk] = e12;
            rk = (rk1 >> 16) & 255; b2[rk] = e20;
            rk = rk1 >> 24; b2[rk] = e21;
            rk = rk2 & 255; b2[rk] = e22;
          }
          v00 = v01 = v10 = v11 = tu0 = tu1 = tt = 0;
          nev = 0;
          have = false;
        }
      }
      ga_fence_async();  // staged entries -> visible to the TMA (async proxy) reads of the next step
        xm3 = xm3 == 2 ? 0 : xm3 + 1;
      }
      ga_bar_step<NT>();
    }
  }
}
