// gm_mma_kernel.cuh -- structured path, third generation: the quadrature sums run on the fp64 tensor pipe.
// Same reference interface as gm_gath_kernel.cuh (jacobian!/residual! of the operator built at
// /root/reference/src/Solver.jl:157, integrands Solver.jl:111-152, arithmetic SURVEY A.4).
//
// k_gath gives every output node pair to one lane that walks the four quadrature points of its 1, 2 or 4 cells with
// ~100 DFMA and 28 shared-memory loads per cell: ncu shows it bound by shared-memory wavefronts and issue slots
// (801 wavefronts, 1,541 warp instructions per cell), the fp64 pipe is 20 % busy.  Every Jacobian entry is LINEAR in the
// per-(cell, point) coefficients (A0, A1, BM, AA, C1, C2, grad u, grad T, u: the state record of gm_gath_kernel.cuh), so
//     D[quad row r (8), output n (8)] += A_m[r, q (4)] * B[q, n]
// with A_m = one coefficient plane of the 8 cells (X - mx, Y0 + r - my) of a quad column and B = the loop-invariant
// reference-element products of the outputs' canonical pairs is exactly one mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4;
// measured on B200, tools/ubench_dmma.cu: 4.0 cycles per instruction per SM = 64 FMA/clk/SM, the plain-DFMA peak, for
// 1/8 of the issue slots and one shared-memory load per 8 cells x 4 points instead of one per lane).  The mirror variants
// (cells left of / below the quad) accumulate into the SAME accumulator fragment: the point permutation q ^ m is an
// address bit of the A load and the sign flips of odd x / y derivative counts are folded into B.
//
//  * Outputs: the 64 node pairs owned by a quad (gm_mma_tiles.h) are packed into eight 8-column tiles so that the 81
//    cell-pair evaluations need 14 (tile, variant) passes of 22 DMMA (14 for n = 1): 308 DMMA per 8 cells.
//  * State: ONE copy per cell (k_gath: four variants), plane-major [plane][cell row][point], computed with DMMA too
//    (fields = iterate values x tabulated basis), 20 planes: 14 Jacobian coefficients + 6 residual fluxes.
//  * Residual: b[node] = sum_m F_m . B, 38 DMMA per quad column, written straight to b.
//  * CTA = 4 main warps (one per SM sub-partition: the tensor pipe is per sub-partition) + 4 aux warps, 8 quad rows,
//    2 CTAs per SM; setmaxnreg moves registers from the aux to the main warpgroup.  Main: tiles, state of column X+1,
//    residual.  Aux: cp.async gathers of dof ids / iterate values / list bases of the coming columns, the constant P-block
//    entries, write-out of the completed lists with 16-byte vector stores.  One staging buffer: the write-out of column
//    X-1 runs under the first DMMA passes of column X; an mbarrier ("stage free") releases the first epilogue store.
//
// The file compiles for the device and -- with GM_EMU -- as host C++ under tests/emu/cuda_emu.h (tests/test_gath_emu.py).
#pragma once
#include <type_traits>

#include "gm_mma_tiles.h"

#define MM_QH 8                      // quad rows per CTA (the M dimension of the tensor instruction)
#define MM_NR (MM_QH + 1)            // cell rows of state (one halo row below)
#define MM_PST (MM_NR * 4)           // doubles per state plane: [cell row][point]
#define MM_VST 36                    // row stride of the staged iterate values (30 used; k-steps read up to index 32; bank spread)
#define MM_QSTRIDE 732               // staged doubles per quad row: 15 list slots (728) + pad (738 = 2 mod 16 was tried: more bank conflicts in the epilogue, 113 vs 93 wavefronts per cell)
#define MM_NMW 4                     // main warps
#define MM_NAW 4                     // aux warps
#define MM_NT ((MM_NMW + MM_NAW) * 32)
#define MM_WARM 6                    // warm-up steps of the column pipeline
#define MM_REG_MAIN 160              // setmaxnreg targets: main + aux = 256 (x 128 threads = half the register file: 2 CTAs / SM)
#define MM_REG_AUX 96

enum { PL_A0, PL_A1, PL_BM, PL_AA, PL_C1, PL_C2, PL_G00, PL_G11, PL_G01, PL_G10, PL_DT0, PL_DT1, PL_U0, PL_U1,
       PL_F0X, PL_F0Y, PL_F1Y, PL_F0C, PL_F1C, PL_FTC, MM_NPL };
enum { MB_PXX, MB_PYY, MB_PXY, MB_PYX, MB_PS, MB_TM, MB_EX, MB_EY };

struct MmArgs {
  CartArgs a;
  MmOut out[8][8];
  int xch;     // quad columns per x-chunk (a CTA sweeps one chunk of one strip)
  int strip0;  // first strip of this launch (the strip next to the lower rank interface runs first, on its own stream)
};

// ---- tensor instruction, mbarrier, register reallocation (device) ; the emulation versions live in cuda_emu.h
#ifndef GM_EMU
__device__ __forceinline__ void mm_dmma(double (&d)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mm_mbar_init(unsigned long long *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mm_mbar_arrive(unsigned long long *bar) {
  asm volatile("{ .reg .b64 t; mbarrier.arrive.shared::cta.b64 t, [%0]; }" ::"r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mm_mbar_wait(unsigned long long *bar, unsigned parity) {
  const uint32_t sa = (uint32_t)__cvta_generic_to_shared(bar);
  asm volatile(
      "{ .reg .pred p;\n"
      "W_%=: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra D_%=;\n"
      "bra W_%=;\n"
      "D_%=: }" ::"r"(sa), "r"(parity) : "memory");
}
__device__ __forceinline__ void mm_syncwarp() { __syncwarp(); }
template <int N> __device__ __forceinline__ void mm_reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void mm_reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }
__device__ __forceinline__ double mm_mask(double v, uint32_t mw, uint32_t sign) {  // (v masked by mw) with the sign bit flipped by `sign`
  return __hiloint2double((int)(((uint32_t)__double2hiint(v) & mw) ^ sign), (int)((uint32_t)__double2loint(v) & mw));
}
#endif

// ---- optional phase timing (development builds only: make EXTRA=-DMM_PROF): cycles per warp and phase of one CTA
#if defined(MM_PROF) && !defined(GM_EMU)
__device__ long long mm_prof[8][8];
#define MM_T(k) do { if (prof_on) { const long long t__ = clock64(); pacc[k] += t__ - ptl; ptl = t__; } } while (0)
#else
#define MM_T(k) do { } while (0)
#endif

struct MmSmem {
  double stage[MM_QH * MM_QSTRIDE];    // staged lists of the quad column being assembled
  double st[3][MM_NPL * MM_PST];       // state ring (cell column mod 3), plane-major
  double val[4][MM_NR][MM_VST];        // iterate values of cell column c (c & 3)
  double visc[4][MM_NR][4][2];
  long long lbase[4][MM_QH][16];       // list bases / ends of quad column X (X & 3); 15 used
  long long lnext[4][MM_QH][16];
  double ek[64][2];                    // per output: TK, sum_q w phi_i phi_j (constant parts, added in the epilogue)
  double sB[7][32];                    // B fragments of the state products (3 k-steps x {phi | dx, dy | 0}, pressure)
  double rB[4][3][32];                 // B fragments of the residual per variant: w dx phi, w dy phi, w phi of the node type (signed, masked)
  double rBp[32];                      // pressure rows: w psi
  double tphi[4][9], tdx[4][9], tdy[4][9], tw[4], twpsi[4][3], tpsi[4][3];
  double UP[18][3];
  unsigned long long mbar_free;        // "stage free": the write-out of the previous column has read its lists
  int32_t gd[8][MM_NR][30];            // dof ids of cell column c (c & 7)
  int32_t lg[4][MM_QH][16];            // dof id of each list of quad column X (0: no list / not assembled here)
  int32_t cls[8][MM_NR];               // rank class of the cells of column c (c & 7)
  int32_t pdiff[4];                    // quad column X (X & 3): its constant P entries differ from those of column X-1
  uint4 wd[2][MM_QH * 15 + 8];         // write-out descriptors of the lists of quad column X (X & 1), see list_desc
  uint4 prec[81];                      // rank records (a.ptab) of the strip's reference class: most cells belong to it
  uint32_t out[64];                    // MmOut of output t*8+n (A | bx<<8 | by<<16 | vmask<<24)
  int16_t slot[16];
  uint8_t par[4][MM_QH][8];            // parities of the list bases of quad column X (X & 3): [node type | 4 = P] bits k
  uint8_t pu[54][4];                   // (A, m, a, p) of the 54 P-row entries of a quad's U lists
};

// the two tiles of main warp w (slot order = processing order).  DMMA per column and warp (one warp per SM sub-partition,
// whose tensor pipe issues one DMMA per 16 cycles): w0 110, w1 88, w2 66 + residual 38, w3 44 + state 38.
__host__ __device__ constexpr int mm_warp_tile(int w, int s) {
  return w == 0 ? (s == 0 ? 6 : 0) : w == 1 ? (s == 0 ? 7 : 1) : w == 2 ? (s == 0 ? 5 : 2) : (s == 0 ? 3 : 4);
}
#define MM_WARP_RESIDUAL 2
#define MM_WARP_STATE 3

template <bool CSR, bool NEWT>
__global__ void __launch_bounds__(MM_NT, 2) k_mma(MmArgs args) {
  using S = MmSmem;
  constexpr int QH = MM_QH, NR = MM_NR, NAT = MM_NAW * 32, NT = MM_NT;
  const CartArgs &a = args.a;
#ifdef GM_EMU
  unsigned char *smem_raw = emu_smem();
#else
  extern __shared__ __align__(16) unsigned char smem_raw[];
#endif
  S &s = *reinterpret_cast<S *>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool is_aux = warp >= MM_NMW;
  const int at = tid - MM_NMW * 32;  // aux thread index
  const int Yq0 = a.row_begin + ((int)blockIdx.y + args.strip0) * QH;  // first quad row of the strip
  const int nqr = min(QH, a.row_end + 1 - Yq0);           // quad rows Yq0 .. Yq0+nqr-1 (row_end = the top mesh line)
  const int Yc0 = Yq0 - 1;                                // cell row of state row 0
  const int X0 = (int)blockIdx.x * args.xch, X1 = min(a.nx + 1, X0 + args.xch);  // quad columns of this x-chunk
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
  for (int k = tid; k < 4 * NR * (MM_VST - 30); k += NT) {  // the pad of every value row is read by the last k-steps (times a zero B row)
    const int row = k / (MM_VST - 30);
    (&s.val[0][0][0])[row * MM_VST + 30 + (k - row * (MM_VST - 30))] = 0.0;
  }
  if (tid < 64) {
    const MmOut &o = args.out[tid >> 3][tid & 7];
    s.out[tid] = (uint32_t)o.A | ((uint32_t)o.bx << 8) | ((uint32_t)o.by << 16) | ((uint32_t)o.vmask << 24);
    const int nA = ga_node(ga_ax(o.A), ga_ay(o.A)), nB = ga_node(o.bx, o.by);
    const CartPairTab &t = a.tab[(CSR ? nA : nB) * 9 + (CSR ? nB : nA)];
    s.ek[tid][0] = t.TK;
    s.ek[tid][1] = (t.TM[0] + t.TM[1]) + (t.TM[2] + t.TM[3]);
  }
  if (tid < 32) {  // B fragments: lane = 4 n + k holds B[k][n]
    const int n = tid >> 2, k = tid & 3;
    {  // state: D column n = 2 q + sel
      const int q = n >> 1, sel = n & 1;
      for (int ks = 0; ks < 3; ++ks) {
        const int dof = 4 * ks + k;
        s.sB[ks][tid] = dof < 9 ? (sel ? cc.sdx[q][dof] : cc.sphi[q][dof]) : 0.0;
        s.sB[3 + ks][tid] = (dof < 9 && !sel) ? cc.sdy[q][dof] : 0.0;
      }
      s.sB[6][tid] = (k < 3 && !sel) ? cc.psi[q][k] : 0.0;
    }
    {  // residual: D column n < 4 = node type n of the quad, k = canonical point
      const int q = k;
      for (int m = 0; m < 4; ++m) {
        const bool part = n < 4 && (n == 0 || (n == 1 && !(m & 1)) || (n == 2 && !(m & 2)) || (n == 3 && m == 0));
        const int nd = n < 4 ? ga_node(ga_ax(n), ga_ay(n)) : 0;
        const double sx = (m & 1) ? -1.0 : 1.0, sy = (m & 2) ? -1.0 : 1.0;
        s.rB[m][0][tid] = part ? sx * cc.w[q] * cc.sdx[q][nd] : 0.0;
        s.rB[m][1][tid] = part ? sy * cc.w[q] * cc.sdy[q][nd] : 0.0;
        s.rB[m][2][tid] = part ? cc.w[q] * cc.sphi[q][nd] : 0.0;
      }
      s.rBp[tid] = n < 3 ? cc.wpsi[q][n] : 0.0;
    }
  }
  // reference rank class of this CTA (the class of an interior cell of the strip): its records are served from shared memory
  const int cls0 = a.cls[(int64_t)min(max(Yc0 + 3, 0), a.ny - 1) * a.nx + min(X0 + 3, a.nx - 1)];
  if (tid < 81) s.prec[tid] = GA_LDG(reinterpret_cast<const uint4 *>(a.ptab) + ((int64_t)cls0 * 81 + tid));
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
    mm_mbar_init(&s.mbar_free, NAT);
  }
  __syncthreads();

  // ---- write-out of quad column X (the four aux warps, two quad rows each -- 16 lanes per row --, right after the step
  //      barrier): plain 16-byte vector stores.  (A TMA bulk copy per list -- k_gath's way -- has a read-completion latency
  //      of several thousand cycles, profiles/r02_ubench_bulk_copy_rate.txt: with ONE staging buffer the main warps would
  //      wait for it every column; a store has released its staging bytes as soon as it has issued.)  Lists are staged at
  //      slot + (base & 1), so source and destination are 16-byte aligned together.  Descriptors (staging offset, chunks,
  //      offset in nzval) are left by list_desc.  The 15 lists go in five groups of three with all descriptor loads, then all
  //      data loads, then all stores of a group issued back to back (one list after the other is a chain of three
  //      shared-memory latencies per list: 3,000 cycles per column); lanes 0 / 1 of a half warp store the odd head / tail.
  auto write_out = [&](int X) {
    const int rowq = 2 * (warp - MM_NMW) + (lane >> 4), sub = lane & 15;
    const double2 *st2 = reinterpret_cast<const double2 *>(s.stage);
    const uint4 *wd = &s.wd[X & 1][rowq * 15];
    auto group = [&](auto L0, auto RR) {
      constexpr int li0 = decltype(L0)::value, R = decltype(RR)::value;  // first list of the group, rounds of 16 chunks per list
      uint4 e[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) e[k] = wd[li0 + k];
      double2 v[3][R];
      double vh[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int nch = (int)(e[k].y & 255);
        const double2 *s2 = st2 + e[k].x;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          v[k][r] = make_double2(0, 0);
          if (sub + 16 * r < nch) v[k][r] = s2[sub + 16 * r];
        }
        vh[k] = 0.0;
        if (sub < 2 && ((e[k].y >> (8 + sub)) & 1)) vh[k] = reinterpret_cast<const double *>(s2)[sub ? 2 * nch : -1];
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const int nch = (int)(e[k].y & 255);
        double *gp = a.nz + (((unsigned long long)e[k].w << 32) | e[k].z);
        double2 *d2 = reinterpret_cast<double2 *>(gp);
#pragma unroll
        for (int r = 0; r < R; ++r)
          if (sub + 16 * r < nch) d2[sub + 16 * r] = v[k][r];
        if (sub < 2 && ((e[k].y >> (8 + sub)) & 1)) gp[sub ? 2 * nch : -1] = vh[k];
      }
    };
    using std::integral_constant;
    group(integral_constant<int, 0>{}, integral_constant<int, 3>{});   // vertex lists: at most 45 chunks
    group(integral_constant<int, 3>{}, integral_constant<int, 2>{});   // bottom-edge lists: at most 27
    group(integral_constant<int, 6>{}, integral_constant<int, 2>{});   // left-edge lists
    group(integral_constant<int, 9>{}, integral_constant<int, 1>{});   // centre lists: at most 15
    group(integral_constant<int, 12>{}, integral_constant<int, 1>{});  // pressure lists: at most 9
  };

#if defined(MM_PROF) && !defined(GM_EMU)
  const bool prof_on = lane == 0 && blockIdx.x == gridDim.x / 2 && blockIdx.y == gridDim.y / 2;
  long long pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ptl = clock64();
#endif
  // =====================================================================================
  // column sweep.  Step X (main: quad column X): aux prepares dof ids of X+3, values of X+2, list records of X+1 and
  // writes out column X-1; main computes the state of X+1 after its tiles.  Four warm-up steps fill the pipeline.
  // =====================================================================================
  if (is_aux) {
#ifndef GM_EMU
    mm_reg_dec<MM_REG_AUX>();
#endif
    // (A1/A2) every aux thread owns up to three fixed (state row, local dof) items of a cell column
    static_assert(NR * 30 <= 3 * NAT, "three load items per aux thread");
    static_assert(2 * MM_NAW == MM_QH, "write_out: two quad rows per aux warp");
    int li_src[3], li_dst[3];   // item index rr*30 + l in gd (-1: none), rr*MM_VST + l in val
    int64_t li_goff[3];         // offset of the item in gdofs for column 0
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int k = at + j * NAT, rr = k / 30;
      const bool ok = k < NR * 30 && row_in_table(rr);
      li_src[j] = ok ? k : -1;
      li_dst[j] = rr * MM_VST + (k - rr * 30);
      li_goff[j] = ok ? (int64_t)(Yc0 + rr) * a.nx * 30 + (k - rr * 30) : 0;
    }
    auto issue_gd = [&](int c) {
      if (!col_ok(c)) return;
      int32_t *dst = &s.gd[c & 7][0][0];
      const int32_t *src = a.gdofs + (int64_t)c * 30;
#pragma unroll
      for (int j = 0; j < 3; ++j)
        if (li_src[j] >= 0) ga_cp4(dst + li_src[j], src + li_goff[j]);
    };
    auto issue_val = [&](int c) {
      if (!col_ok(c)) return;
      const int32_t *gdc = &s.gd[c & 7][0][0];
      double *dst = &s.val[c & 3][0][0];
#pragma unroll
      for (int j = 0; j < 3; ++j)
        if (li_src[j] >= 0) {
          const int32_t d = gdc[li_src[j]];
          ga_cp8(dst + li_dst[j], d > 0 ? a.x + (d - 1) : a.dvals + (-d - 1));
        }
      for (int k = at - 32; k >= 0 && k < NR * 4; k += NAT) {  // (warp 5 and a few threads of warp 6)
        const int rr = k >> 2, q = k & 3;
        if (!row_in_table(rr)) continue;
        const int64_t cell = (int64_t)(Yc0 + rr) * a.nx + c;
        if (!NEWT) ga_cp16(&s.visc[c & 3][rr][q][0], a.visc + (cell * 4 + q) * 2);
        if (q == 0) ga_cp4(&s.cls[c & 7][rr], a.cls + cell);
      }
    };
    // (A4) list records of quad column c: one thread per (quad row, node type | P lists) = three lists
    const int lt = at - 64;  // list-record tasks live on warps 6 and 7 (the value gathers keep warps 4 and 5 busier)
    auto issue_lists = [&](int c) {
      if (lt < 0 || lt >= QH * 5) return;
      const int rowq = lt / 5, grp = lt - rowq * 5;
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
    // ... and the "same P entries as the previous column" test: the constant P-block entries keep their staging positions
    // from column to column as long as rank classes, base parities of the U and P lists and the column validity do not
    // change (the staging buffer is never cleared and nobody else writes those positions), so p_entries runs only when
    // this thread-wide vote finds a difference (the interior of a mesh: once per x-chunk).
    auto lists_parity = [&](int c) {  // after this thread's cp.async copies have landed (end of step c-1)
      if (at == 0) s.pdiff[(c + 2) & 3] = 0;  // (read at step c+2, voted at the end of step c+1)
      if (lt < 0 || lt >= QH * 5) return;
      const int rowq = lt / 5, grp = lt - rowq * 5;
      const long long *lb = &s.lbase[c & 3][rowq][grp * 3];
      const uint8_t pb = (uint8_t)((lb[0] & 1) | ((lb[1] & 1) << 1) | ((lb[2] & 1) << 2));
      s.par[c & 3][rowq][grp] = pb;
      bool diff = ((pb ^ s.par[(c - 1) & 3][rowq][grp]) & (grp < 4 ? 3 : 7)) != 0;
      if (grp == 0) {  // classes of the cells of this quad row (and of the halo row by the thread of row 0) in columns c, c-1, c-2
        for (int rr = rowq == 0 ? 0 : rowq + 1; rr <= rowq + 1; ++rr)
          if (row_in_table(rr)) diff = diff || s.cls[c & 7][rr] != s.cls[(c - 1) & 7][rr] || s.cls[(c - 1) & 7][rr] != s.cls[(c - 2) & 7][rr];
      }
      if (diff) s.pdiff[c & 3] = 1;
    };
    // write-out descriptor of list `at` of quad column c (one step after its base / end have landed, off the critical path)
    auto list_desc = [&](int c) {
      if (at >= QH * 15) return;
      const int rowq = at / 15, li = at - rowq * 15;
      uint4 d = uint4{0, 0, 0, 0};
      if (s.lg[c & 3][rowq][li] > 0) {
        const long long base = s.lbase[c & 3][rowq][li];
        const int len = (int)(s.lnext[c & 3][rowq][li] - base), par = (int)(base & 1);
        const unsigned long long off = (unsigned long long)(base + par);  // first 16-byte chunk in nzval (doubles)
        // x: first chunk in the staging buffer (16-byte units); y: chunks | odd head << 8 | odd tail << 9; (z, w): off
        d = uint4{(uint32_t)((rowq * MM_QSTRIDE + s.slot[li]) / 2 + par), (uint32_t)((len - par) >> 1) | ((uint32_t)par << 8) | ((uint32_t)((len - par) & 1) << 9),
                  (uint32_t)off, (uint32_t)(off >> 32)};
      }
      s.wd[c & 1][at] = d;
    };
    // (A5) constant P-block entries of quad column X -> staged lists.  Aux thread e < 108 owns entry type e
    int pe_li = 0, pe_rk = 0, pe_dx = 0, pe_dy = 0;
    double pe_val = 0.0;
    const int pe = at;
    const bool pe_on = pe < 108, pe_plist = pe >= 54;
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
    const int pe_cls0 = cls0;
    const int pe_rk0 = pe_on ? (int)GA_LDG(a.ctab + (int64_t)pe_cls0 * 900 + pe_rk) : 0xFF;
    auto p_entries = [&](int X) {
      if (!pe_on) return;
      const int cx = X - pe_dx;
      if (!col_ok(cx)) return;
      double *stg0 = s.stage + s.slot[pe_li];
      int rk[QH];
      bool own[QH];
#pragma unroll
      for (int rowq = 0; rowq < QH; ++rowq) {  // all rank loads in flight together
        const int rr = rowq + 1 - pe_dy;
        own[rowq] = row_owned(rr);
        const bool ok = rowq < nqr && row_in_table(rr) && (own[rowq] || !pe_plist);
        const int cl = ok ? s.cls[cx & 7][rr] : pe_cls0;
        rk[rowq] = !ok ? 0xFF : (cl == pe_cls0 ? pe_rk0 : (int)GA_LDG(a.ctab + (int64_t)cl * 900 + pe_rk));
      }
#pragma unroll
      for (int rowq = 0; rowq < QH; ++rowq) {
        if (rk[rowq] == 0xFF) continue;
        const int par = (s.par[X & 3][rowq][pe_li / 3] >> (pe_li % 3)) & 1;
        stg0[rowq * MM_QSTRIDE + par + rk[rowq]] = own[rowq] ? pe_val : 0.0;
      }
    };
    // Step X: the gathers issued here (dof ids of column X+5, iterate values / viscosity / rank class of X+3, list bases of
    // X+2) form one cp.async group that only has to land by the END of the NEXT step (wait_group 1): no memory latency on
    // the step's critical path.  Consumers: issue_val(c) reads the dof ids of c two steps after their issue, state(c) the
    // values of c two steps after theirs, lists_parity(c) the bases of c one step after theirs.
    unsigned phase = 0;
    for (int X = X0 - MM_WARM; X < X1; ++X) {
      const bool live = X >= X0;
      MM_T(7);
      if (live && X > X0 && jac) write_out(X - 1);
      mm_mbar_arrive(&s.mbar_free);        // my stores of column X-1 have issued: their staging bytes are released
      MM_T(0);
      if (live && jac) list_desc(X);
      issue_gd(X + 5);
      issue_val(X + 3);
      if (X + 2 >= X0 && X + 2 < X1) issue_lists(X + 2);
      ga_cp_commit();
      MM_T(1);
      if (live && jac && (X < X0 + 3 || X + 1 >= a.nx || s.pdiff[X & 3] != 0)) {  // (uniform over the aux warps)
        mm_mbar_wait(&s.mbar_free, phase);  // ... and everybody else's
        p_entries(X);
      }
      phase ^= 1;
      MM_T(2);
      ga_cp_wait_prev();                   // the group of step X-1 has landed
      MM_T(3);
      if (X + 1 >= X0 && X + 1 < X1) lists_parity(X + 1);
      MM_T(4);
      ga_bar_step<NT>();  // lists of column X staged, state of X+1 ready
    }
    if (jac) write_out(X1 - 1);  // last column of the chunk
    ga_cp_wait_all();            // (gathers issued for columns beyond the chunk)
  } else {
#ifndef GM_EMU
    mm_reg_inc<MM_REG_MAIN>();
#endif
    // =====================================================================================
    // main warps: lane = 4 g + j.  As A / accumulator row: quad row g.  As B column: output n = g of the warp's tiles,
    // k = point j.  As accumulator columns: outputs 2 j, 2 j + 1.
    // =====================================================================================
    const int g = lane >> 2, j = lane & 3;
    // role of this warp (which tiles / tasks): rotated from CTA to CTA, so that the two CTAs of an SM do not put their heaviest
    // warps (110 tensor instructions per column) on the same sub-partition's tensor pipe
    const int wrole = (warp + (int)blockIdx.x + (int)blockIdx.y) & 3;
    double Bt[2][8];      // per tile slot: reference-element products of B column g at point j
    uint32_t bmask[2];    // per tile slot: variant mask of B column g
    int tile8[2];         // per tile slot: 8 * tile id
    uint32_t tvar[2];     // per tile slot: variants of the tile (union over its columns)
    uint32_t ebase[2][2]; // per tile slot and accumulator column: first staging slot | slot stride << 10 | owner type << 20
    uint32_t selfast[2];  // per tile slot: epilogue selection of the two accumulator columns when both cell columns exist
    // epilogue selection of output `ow` (packed MmOut) for accumulator row g: which cell supplies the rank record
    // (any cell in the table that contains both nodes), how many owned cells are summed.
    // [0:6] pair index, [7:10] state row of the record cell, [11] its mx, [12:14] owned cells, [15] have
    auto select = [&](uint32_t ow, bool okL, bool okR) -> uint32_t {
      const int A = ow & 255, bx = (ow >> 8) & 255, by = (ow >> 16) & 255, vm = ow >> 24;
      const int ax = ga_ax(A), ay = ga_ay(A);
      uint32_t r = 0, nev = 0;
      if (g < nqr) {
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          if (!((vm >> m) & 1)) continue;
          const int mx = m & 1, my = m >> 1, rr = g + 1 - my;
          if (!(mx ? okL : okR)) continue;
          if (row_in_table(rr)) {
            const int Mact = ga_node(mx ? 2 - ax : ax, my ? 2 - ay : ay), mact = ga_node(mx ? 2 - bx : bx, my ? 2 - by : by);
            r = (uint32_t)(Mact * 9 + mact) | ((uint32_t)rr << 7) | ((uint32_t)mx << 11) | (1u << 15);
          }
          if (row_owned(rr)) ++nev;
        }
      }
      return r | (nev << 12);
    };
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
      int t = -1;
#pragma unroll
      for (int w = 0; w < 4; ++w)
        if (wrole == w) t = mm_warp_tile(w, sl);
      bmask[sl] = 0; selfast[sl] = 0; tile8[sl] = 8 * t; tvar[sl] = 0; ebase[sl][0] = ebase[sl][1] = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) Bt[sl][k] = 0.0;
      if (t < 0) continue;
      const uint32_t ob = s.out[t * 8 + g];
      const int A = ob & 255, bx = (ob >> 8) & 255, by = (ob >> 16) & 255;
      const int nA = ga_node(ga_ax(A), ga_ay(A)), nB = ga_node(bx, by);
      const int it = CSR ? nA : nB, jt = CSR ? nB : nA;
      const CartPairTab &tb = a.tab[it * 9 + jt];
      const double xi = cc.sdx[j][it], yi = cc.sdy[j][it], xj = cc.sdx[j][jt], yj = cc.sdy[j][jt];
      Bt[sl][MB_PXX] = xi * xj; Bt[sl][MB_PYY] = yi * yj; Bt[sl][MB_PXY] = xi * yj; Bt[sl][MB_PYX] = yi * xj;
      Bt[sl][MB_PS] = xi * yj + yi * xj;
      Bt[sl][MB_TM] = tb.TM[j]; Bt[sl][MB_EX] = tb.TP[j] * xj; Bt[sl][MB_EY] = tb.TP[j] * yj;
      bmask[sl] = ob >> 24;
#pragma unroll
      for (int tt = 0; tt < 8; ++tt)
        if (t == tt) tvar[sl] = (uint32_t)mm_tvar(tt);
      if ((tvar[sl] & (tvar[sl] - 1)) == 0) {  // single-variant tile (every column takes part): fold the variant's signs into B once
        const uint32_t tvv = tvar[sl];
        const int m = tvv == 1 ? 0 : (tvv == 2 ? 1 : (tvv == 4 ? 2 : 3));
        const double sxy = (m == 1 || m == 2) ? -1.0 : 1.0, sx = (m & 1) ? -1.0 : 1.0, sy = (m & 2) ? -1.0 : 1.0;
        Bt[sl][MB_PXY] *= sxy; Bt[sl][MB_PYX] *= sxy; Bt[sl][MB_PS] *= sxy; Bt[sl][MB_EX] *= sx; Bt[sl][MB_EY] *= sy;
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {  // staging offsets of the three lists of the owner of accumulator column 2 j + c
        const int Ao = s.out[t * 8 + 2 * j + c] & 255;
        const int slot0 = Ao == 0 ? 0 : (Ao == 1 ? 258 : (Ao == 2 ? 414 : 570)), slen = Ao == 0 ? 90 : (Ao == 3 ? 34 : 54);
        ebase[sl][c] = (uint32_t)slot0 | ((uint32_t)slen << 10) | ((uint32_t)Ao << 20);
      }
      selfast[sl] = select(s.out[t * 8 + 2 * j], true, true) | (select(s.out[t * 8 + 2 * j + 1], true, true) << 16);
    }
    int xm3 = (((X0 - MM_WARM) % 3) + 3) % 3;
    unsigned phase = 0;
    bool stage_ok = false;  // this step's "stage free" wait is done

    // ---- rank records of the two accumulator columns of both tiles: needed in the epilogues only, fetched first thing in a
    //      step so that their latency (a class lookup, then the record: shared memory for the strip's reference class, L2 else)
    //      is hidden behind the passes.
    uint32_t selx[2][2];
    uint4 recx[2][2];
    auto fetch_records = [&](int X) {
      const bool okL = X > 0, okR = X < a.nx;
      int cl[4];
#pragma unroll
      for (int sl = 0; sl < 2; ++sl) {
        const uint32_t sf = (okL && okR) ? selfast[sl] : (select(s.out[tile8[sl] + 2 * j], okL, okR) | (select(s.out[tile8[sl] + 2 * j + 1], okL, okR) << 16));
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const uint32_t se = (sf >> (16 * c)) & 0xFFFFu;
          selx[sl][c] = se;
          const int cx = X - (int)((se >> 11) & 1);
          cl[2 * sl + c] = (se >> 15) ? s.cls[cx & 7][(se >> 7) & 15] : cls0;  // (four independent class lookups ...)
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) recx[k >> 1][k & 1] = s.prec[selx[k >> 1][k & 1] & 127];  // (... four independent record loads)
      if (cl[0] != cls0 || cl[1] != cls0 || cl[2] != cls0 || cl[3] != cls0) {  // rare: a cell of another rank class (mesh boundary)
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (cl[k] != cls0)
            recx[k >> 1][k & 1] = GA_LDG(reinterpret_cast<const uint4 *>(a.ptab) + ((int64_t)cl[k] * 81 + (selx[k >> 1][k & 1] & 127)));
      }
    };

    // ---- one accumulator tile of quad column X (the tile in slot SL of this warp).  The variant loop is a run-time loop
    //      and all four main warps run the same code (tile id, variant set and column mask are per-warp registers): the
    //      straight-line version (every tile x variant unrolled: 130 KB of SASS) lost 20 % of the issue slots to instruction fetch.
    auto tile = [&](auto SS, int X) {
      constexpr int SL = decltype(SS)::value;
      const int T8 = tile8[SL];  // 8 * tile id
      const uint32_t tv = tvar[SL];
      const bool okL = X > 0, okR = X < a.nx;
      const uint32_t *sel = selx[SL];  // epilogue selection and rank records: fetched at the start of the step (fetch_records)
      const uint4 *rec = recx[SL];
      double v00[2] = {0, 0}, v11[2] = {0, 0}, v01[2] = {0, 0}, v10[2] = {0, 0}, ee[2] = {0, 0}, tu0[2] = {0, 0}, tu1[2] = {0, 0};
#pragma unroll 1
      for (int m = 0; m < 4; ++m) {
        if (!((tv >> m) & 1)) continue;
        const int mx = m & 1, my = m >> 1;
        if (!(mx ? okL : okR)) continue;
        const double *P = s.st[mx ? (xm3 == 0 ? 2 : xm3 - 1) : xm3] + (g + 1 - my) * 4 + (j ^ m);
        // slot 1 always holds a single-variant tile (signs already folded into B, every column takes part); slot 0 masks
        // the B column of a lane that does not take part in this variant and applies the variant's signs
        const uint32_t mw = (SL == 1 || ((bmask[SL] >> m) & 1)) ? 0xFFFFFFFFu : 0u;
        const bool fl = SL == 0 && (tv & (tv - 1)) != 0;
        const uint32_t sxy = fl ? (uint32_t)((m ^ (m >> 1)) & 1) << 31 : 0u, sx = fl ? (uint32_t)mx << 31 : 0u, sy = fl ? (uint32_t)my << 31 : 0u;
        const double bPXX = SL ? Bt[SL][MB_PXX] : mm_mask(Bt[SL][MB_PXX], mw, 0), bPYY = SL ? Bt[SL][MB_PYY] : mm_mask(Bt[SL][MB_PYY], mw, 0),
                     bTM = SL ? Bt[SL][MB_TM] : mm_mask(Bt[SL][MB_TM], mw, 0);
        const double bPXY = SL ? Bt[SL][MB_PXY] : mm_mask(Bt[SL][MB_PXY], mw, sxy), bPYX = SL ? Bt[SL][MB_PYX] : mm_mask(Bt[SL][MB_PYX], mw, sxy),
                     bPS = SL ? Bt[SL][MB_PS] : mm_mask(Bt[SL][MB_PS], mw, sxy);
        const double bEX = SL ? Bt[SL][MB_EX] : mm_mask(Bt[SL][MB_EX], mw, sx), bEY = SL ? Bt[SL][MB_EY] : mm_mask(Bt[SL][MB_EY], mw, sy);
        // all coefficient planes of the pass first (the loads are independent of everything before them)
        double pa[14];
#pragma unroll
        for (int k = 0; k < 14; ++k)
          if (!(NEWT && (k == PL_AA || k == PL_C1 || k == PL_C2))) pa[k] = P[k * MM_PST];
        mm_dmma(v00, pa[PL_A0], bPXX);
        mm_dmma(v11, pa[PL_A1], bPYY);
        mm_dmma(v01, pa[PL_BM], bPYX); mm_dmma(v10, pa[PL_BM], bPXY); mm_dmma(v00, pa[PL_BM], bPYY); mm_dmma(v11, pa[PL_BM], bPXX);
        mm_dmma(tu0, pa[PL_DT0], bTM);
        mm_dmma(tu1, pa[PL_DT1], bTM);
        mm_dmma(ee, pa[PL_U0], bEX);      // w phi_i (u . grad phi_j)
        mm_dmma(v01, pa[PL_G10], bTM);    // (a=0,b=1): + w phi phi G[1][0]
        mm_dmma(v10, pa[PL_G01], bTM);    // (a=1,b=0): + w phi phi G[0][1]
        mm_dmma(v00, pa[PL_G00], bTM);
        mm_dmma(v11, pa[PL_G11], bTM);
        mm_dmma(ee, pa[PL_U1], bEY);
        if (!NEWT) {
          mm_dmma(v01, pa[PL_AA], bPXY); mm_dmma(v10, pa[PL_AA], bPYX);
          mm_dmma(v00, pa[PL_C1], bPS); mm_dmma(v01, pa[PL_C1], bPXX); mm_dmma(v10, pa[PL_C1], bPXX);
          mm_dmma(v11, pa[PL_C2], bPS); mm_dmma(v01, pa[PL_C2], bPYY); mm_dmma(v10, pa[PL_C2], bPYY);
        }
      }
      // ---- epilogue: nine entries per output -> staged lists (positions from the rank record; entries without a
      //      position go to the trash element of their slot)
      MM_T(2);
      if (!stage_ok) { mm_mbar_wait(&s.mbar_free, phase); stage_ok = true; }
      MM_T(3);
      double *q0 = s.stage + g * MM_QSTRIDE;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (!(sel[c] >> 15)) continue;
        const uint32_t eb = ebase[SL][c];
        const uint32_t pm = s.par[X & 3][g][eb >> 20];
        const int slot0 = eb & 1023, slen = (eb >> 10) & 1023;
        double *b0 = q0 + slot0 + (pm & 1), *b1 = q0 + slot0 + slen + ((pm >> 1) & 1), *b2 = q0 + slot0 + 2 * slen + ((pm >> 2) & 1);
        const double cn = (double)((sel[c] >> 12) & 7);
        const double2 ekv = *reinterpret_cast<const double2 *>(&s.ek[T8 + 2 * j + c][0]);  // TK, sum_q w phi_i phi_j
        const double mb = cn * ekv.y;
        const double ut0 = -a.RaPr * a.gx * mb, ut1 = -a.RaPr * a.gy * mb;
        const double e = ee[c], tt = fma(cn, ekv.x, e);
        const double w00 = v00[c] + e, w11 = v11[c] + e;
        double e00, e01, e02, e10, e11, e12, e20, e21, e22;  // [list k][minor dof l]
        if (CSR) { e00 = w00; e01 = v01[c]; e02 = ut0; e10 = v10[c]; e11 = w11; e12 = ut1; e20 = tu0[c]; e21 = tu1[c]; e22 = tt; }
        else     { e00 = w00; e01 = v10[c]; e02 = tu0[c]; e10 = v01[c]; e11 = w11; e12 = tu1[c]; e20 = ut0; e21 = ut1; e22 = tt; }
        const uint32_t rk0 = rec[c].x, rk1 = rec[c].y, rk2 = rec[c].z;
        b0[rk0 & 255] = e00;
        b0[(rk0 >> 8) & 255] = e01;
        b0[(rk0 >> 16) & 255] = e02;
        b1[rk0 >> 24] = e10;
        b1[rk1 & 255] = e11;
        b1[(rk1 >> 8) & 255] = e12;
        b2[(rk1 >> 16) & 255] = e20;
        b2[rk1 >> 24] = e21;
        b2[rk2 & 255] = e22;
      }
    };

    // ---- state of cell column c (ring slot cm3): fields at the quadrature points by DMMA (D[cell row][2 q + sel] =
    //      values[cell row][dof] * table[dof][2 q + sel]), then the 20 coefficient planes.  Two row tiles: rows 1..8, halo row 0.
    auto state_task = [&](int c, int cm3) {
      if (!col_ok(c)) return;
      double sb[7];
#pragma unroll
      for (int k = 0; k < 7; ++k) sb[k] = s.sB[k][lane];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int rr = mt == 0 ? g + 1 : 0;
        const double *vv = &s.val[c & 3][rr][0];
        double d1[3][2] = {{0, 0}, {0, 0}, {0, 0}}, d2[3][2] = {{0, 0}, {0, 0}, {0, 0}}, d3[2] = {0, 0};
#pragma unroll
        for (int f = 0; f < 3; ++f) {
          const int base = f == 0 ? 0 : (f == 1 ? 9 : 21);
#pragma unroll
          for (int ks = 0; ks < 3; ++ks) {
            const double av = vv[base + 4 * ks + j];
            mm_dmma(d1[f], av, sb[ks]);
            mm_dmma(d2[f], av, sb[3 + ks]);
          }
        }
        mm_dmma(d3, vv[18 + j], sb[6]);
        if (mt == 1 && g != 0) continue;
        const int q = j;
        const bool own = row_owned(rr);
        const double u0 = d1[0][0], G00 = d1[0][1], G10 = d2[0][0], u1 = d1[1][0], G01 = d1[1][1], G11 = d2[1][0];
        const double Tq = d1[2][0], dT0 = d1[2][1], dT1 = d2[2][0], pq = d3[0];
        const double D01 = 0.5 * (G01 + G10);
        const double mu = NEWT ? 1.0 : s.visc[c & 3][rr][q][0];
        const double wj = a.js * s.tw[q] * a.Pr, mup = wj * mu, sc = NEWT ? 0.0 : 2.0 * wj * s.visc[c & 3][rr][q][1];
        const double s00 = sc * G00, s01 = sc * D01, s11 = sc * G11;
        const double m2 = 2.0 * a.Pr * mu, bu = a.RaPr * Tq;
        double pl[MM_NPL];
        pl[PL_A0] = fma(s00, G00, 2.0 * mup); pl[PL_A1] = fma(s11, G11, 2.0 * mup); pl[PL_BM] = fma(s01, D01, mup);
        pl[PL_AA] = s00 * G11; pl[PL_C1] = s00 * D01; pl[PL_C2] = s01 * G11;
        pl[PL_G00] = G00; pl[PL_G11] = G11; pl[PL_G01] = G01; pl[PL_G10] = G10;
        pl[PL_DT0] = dT0; pl[PL_DT1] = dT1; pl[PL_U0] = u0; pl[PL_U1] = u1;
        // residual fluxes (Solver.jl:125-140): r_ux = sum w (F0X dx phi + F0Y dy phi + F0C phi), r_uy = (F0Y, F1Y, F1C), r_T = (dT0, dT1, FTC)
        pl[PL_F0X] = fma(m2, G00, -pq); pl[PL_F0Y] = m2 * D01; pl[PL_F1Y] = fma(m2, G11, -pq);
        pl[PL_F0C] = fma(u0, G00, u1 * G10) - bu * a.gx;
        pl[PL_F1C] = fma(u0, G01, u1 * G11) - bu * a.gy;
        pl[PL_FTC] = fma(u0, dT0, u1 * dT1);
        double *o = s.st[cm3] + rr * 4 + q;
#pragma unroll
        for (int k = 0; k < MM_NPL; ++k) o[k * MM_PST] = own ? pl[k] : 0.0;
      }
    };

    // ---- residual of the dofs of quad column X: D[quad row][node type] over the four variants, straight to b
    auto residual_task = [&](int X) {
      const bool okL = X > 0, okR = X < a.nx;
      double ru[2] = {0, 0}, rv[2] = {0, 0}, rt[2] = {0, 0}, rp[2] = {0, 0};
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int mx = m & 1, my = m >> 1;
        if (!(mx ? okL : okR)) continue;
        const double *P = s.st[mx ? (xm3 == 0 ? 2 : xm3 - 1) : xm3] + (g + 1 - my) * 4 + (j ^ m);
        const double bx_ = s.rB[m][0][lane], by_ = s.rB[m][1][lane], bc_ = s.rB[m][2][lane];
        double av;
        av = P[PL_F0X * MM_PST]; mm_dmma(ru, av, bx_);
        av = P[PL_F0Y * MM_PST]; mm_dmma(ru, av, by_); mm_dmma(rv, av, bx_);
        av = P[PL_F1Y * MM_PST]; mm_dmma(rv, av, by_);
        av = P[PL_F0C * MM_PST]; mm_dmma(ru, av, bc_);
        av = P[PL_F1C * MM_PST]; mm_dmma(rv, av, bc_);
        av = P[PL_DT0 * MM_PST]; mm_dmma(rt, av, bx_);
        av = P[PL_DT1 * MM_PST]; mm_dmma(rt, av, by_);
        av = P[PL_FTC * MM_PST]; mm_dmma(rt, av, bc_);
        if (m == 0) {
          const double bp = s.rBp[lane];
          av = P[PL_G00 * MM_PST]; mm_dmma(rp, av, bp);
          av = P[PL_G11 * MM_PST]; mm_dmma(rp, av, bp);
        }
      }
      if (g < nqr && j < 2) {
        const int32_t *lg = &s.lg[X & 3][g][0];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int An = 2 * j + c;
          if (lg[An * 3 + 0] > 0) a.b[lg[An * 3 + 0] - 1] = ru[c];
          if (lg[An * 3 + 1] > 0) a.b[lg[An * 3 + 1] - 1] = rv[c];
          if (lg[An * 3 + 2] > 0) a.b[lg[An * 3 + 2] - 1] = rt[c];
          const int p = 2 * j + c;
          if (p < 3 && lg[12 + p] > 0) a.b[lg[12 + p] - 1] = rp[c];
        }
      }
    };

    using std::integral_constant;
    for (int X = X0 - MM_WARM; X < X1; ++X) {
      const bool live = X >= X0;
      stage_ok = false;
      MM_T(7);
      if (live && jac) fetch_records(X);
      // (the tasks that do not touch the staging buffer come first: the write-out of column X-1 is still reading it)
      if (wrole == MM_WARP_RESIDUAL && live && res) residual_task(X);
      if (wrole == MM_WARP_STATE) state_task(X + 1, xm3 == 2 ? 0 : xm3 + 1);
      MM_T(1);
      if (live && jac) {
        tile(integral_constant<int, 0>{}, X);
        MM_T(5);
        tile(integral_constant<int, 1>{}, X);
        MM_T(6);
      }
      phase ^= 1;
      xm3 = xm3 == 2 ? 0 : xm3 + 1;
      ga_bar_step<NT>();
    }
  }
#if defined(MM_PROF) && !defined(GM_EMU)
  if (prof_on)
    for (int k = 0; k < 8; ++k) mm_prof[warp][k] = pacc[k];
#endif
}
