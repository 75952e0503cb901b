// emu_gather.cpp -- TEST INFRASTRUCTURE ONLY: host emulation build of the generic gather path's kernels
// (gm_gather_kernel.cuh; see cuda_emu.h).  Built by tests/test_gather_emu.py:
//   g++ -O1 -std=c++17 -shared -fPIC -pthread -o _emu_gather.so emu_gather.cpp
#include "cuda_emu.h"

static inline void gg_syncwarp() { mm_syncwarp(); }
static inline void gg_cp8(void *d, const void *s) { ga_cp8(d, s); }
static inline void gg_cp16(void *d, const void *s) { ga_cp16(d, s); }
static inline void gg_cp_commit() { ga_cp_commit(); }
static inline void gg_cp_wait_prev() { ga_cp_wait_prev(); }
static inline double gg_shfl_xor(double v, int m) { return ga_shfl_xor(v, m); }
static inline int32_t gg_shfl_i(int32_t v, int src);
static inline int32_t gg_lane_word(int32_t v, int src) { return gg_shfl_i(v, src & 31); }
static int32_t g_emu_ibuf[64][32];
static inline int32_t gg_shfl_i(int32_t v, int src) {  // all 32 threads of the warp must call it
  const unsigned w = threadIdx.x >> 5, l = threadIdx.x & 31;
  g_emu_ibuf[w][l] = v;
  pthread_barrier_wait(&g_emu_bar_warp[w]);
  const int32_t r = g_emu_ibuf[w][src];
  pthread_barrier_wait(&g_emu_bar_warp[w]);
  return r;
}
#include "../../gemotion.jl_b200/csrc/gm_gather_kernel.cuh"

// partitioned mesh (gm_desc.n_own_cells / ghost_cell_ids): global ids of the local cells; a major is assembled iff its
// lowest-numbered incident cell is an own cell (k_gg_desc's rule)
static int64_t g_cell_begin = 0;
static const int64_t *g_ghost_ids = nullptr;

template <int NN, int NQ, int NV>
static int run(GgArgs a, bool csr, int grid) {
  constexpr int CPB = 16;
  a.nbatch = (int)((a.ncells + CPB - 1) / CPB);
  const unsigned nb = (unsigned)grid < (unsigned)a.nbatch ? (unsigned)grid : (unsigned)a.nbatch;  // persistent CTAs: several batches each
  const size_t sm = sizeof(GgSmem<NN, NQ, NV, CPB>);
  // (k_cellmat uses block barriers only; the per-warp barriers of the emulator are created for whole warps)
  int rc = csr ? emu_launch(k_cellmat<NN, NQ, NV, CPB, true>, a, nb, 1, NN * CPB, 0, sm)
               : emu_launch(k_cellmat<NN, NQ, NV, CPB, false>, a, nb, 1, NN * CPB, 0, sm);
  if (rc) return rc;
  // records of the majors as k_gg_desc / k_gg_permute (gm_gather.cuh) build them: groups of GG_MPW, sorted by first incident cell
  const int64_t nUp = (a.nU + GG_MPW - 1) / GG_MPW * GG_MPW, nrec = (nUp + (a.n - a.oT) + GG_MPW - 1) / GG_MPW * GG_MPW;
  const int64_t ngroup = nrec / GG_MPW;
  std::vector<GgDesc> d0((size_t)nrec), desc((size_t)nrec);
  std::vector<std::pair<int32_t, int32_t>> order((size_t)ngroup);
  int maxlen = 0;
  const int ld = 3 * NN + 3;
  for (int64_t wg = 0; wg < nrec; ++wg) {
    GgDesc &d = d0[wg];
    memset(&d, 0, sizeof d);
    const bool real = wg < a.nU || (wg >= nUp && wg - nUp + a.oT < a.n);
    const int64_t j = wg < a.nU ? wg : wg - nUp + a.oT;
    d.isU = wg < nUp ? 1 : 0;
    if (!real) { d.p0 = a.ptr[wg < nUp ? a.nU : a.n]; d.j = -1; }
    else {
      const int64_t a0 = a.adjptr[j];
      d.p0 = a.ptr[j]; d.j = (int32_t)j;
      d.len = (int32_t)(a.ptr[j + 1] - d.p0); d.cnt = (int32_t)(a.adjptr[j + 1] - a0);
      if (g_ghost_ids) {
        int64_t lowest = INT64_MAX;
        for (int t = 0; t < d.cnt; ++t) {
          const int64_t lc = a.adj[a0 + t] / ld, gc = lc < a.nown ? g_cell_begin + lc : g_ghost_ids[lc - a.nown];
          lowest = std::min(lowest, gc);
        }
        if (!(lowest >= g_cell_begin && lowest < g_cell_begin + a.nown)) { d.cnt = 0; d.j = -1; }
      }
      for (int t = 0; t < 10; ++t) d.e[t] = t < d.cnt ? gg_seg<NN>(a.adj[a0 + t]) : 0;
      maxlen = std::max(maxlen, (int)d.len);
    }
    if (wg % GG_MPW == 0) order[wg / GG_MPW] = {real && d.cnt > 0 ? ((a.adj[a.adjptr[j]] / ld) / 8) * 4 + (wg < nUp ? 0 : 2) + (d.cnt >= 4 ? 0 : 1) : 0x7fffffff, (int32_t)(wg / GG_MPW)};  // (windows of 8 cells here)
  }
  std::stable_sort(order.begin(), order.end(), [](auto &x, auto &y) { return x.first < y.first; });
  for (int64_t g = 0; g < ngroup; ++g)
    for (int r = 0; r < GG_MPW; ++r) desc[g * GG_MPW + r] = d0[(int64_t)order[g].second * GG_MPW + r];
  a.desc = desc.data();
  a.ngroup = ngroup;
  a.lstride = (maxlen + 3) & ~3;
  const unsigned need = (unsigned)((ngroup + GG_LWARPS - 1) / GG_LWARPS);
  (void)grid;
  return emu_launch(k_lists<NN>, a, need, 1, GG_LWARPS * 32, 0, sizeof(double) * GG_LWARPS * GG_MPW * a.lstride);
}

extern "C" int emu_gather_run(int nn, int64_t ncells, int64_t nU, int64_t oT, int64_t n, const int32_t *gdofs, const double *dvals,
                              const double *coords, const int32_t *cellv, const int64_t *ptr, const uint16_t *grank, const uint16_t *grankp, int grid,
                              const int64_t *adjptr, const int32_t *adj, const double *x, double *nz, double *b, unsigned flags, int csr,
                              const double *prm /*Pr Ra n reg gx gy js*/, int cartesian, double hx, double hy, const double *w,
                              const double *phi, const double *dphi, const double *psi, const double *dgphi,
                              int64_t nown, int64_t cell_begin, const int64_t *ghost_ids) {
  const int nq = nn == 9 ? 4 : 3, nv = nn == 9 ? 4 : 3, ld = 3 * nn + 3;
  static GmConst c;
  memset(&c, 0, sizeof c);
  memcpy(c.w, w, 8 * nq);
  memcpy(c.phi, phi, 8 * nq * nn);
  memcpy(c.dphi, dphi, 8 * nq * nn * 2);
  memcpy(c.psi, psi, 8 * nq * 3);
  if (dgphi) memcpy(c.dgphi, dgphi, 8 * nq * nv * 2);
  c.Pr = prm[0]; c.Ra = prm[1]; c.n = prm[2]; c.reg = prm[3]; c.gx = prm[4]; c.gy = prm[5]; c.jac_scaling = prm[6];
  c.cartesian = cartesian; c.hx = hx; c.hy = hy;
  const int64_t nec = (int64_t)2 * nn * ld + 3 * nn * nn;
  // poisoned workspaces: an entry k_cellmat does not write, or k_lists reads from the wrong place, shows up as a huge number
  std::vector<double> cm((size_t)(ncells * nec), 1e300), cr((size_t)(ncells * ld), 1e300);
  GgArgs a;
  a.ncells = ncells; a.nown = nown > 0 ? nown : ncells; a.nU = nU;
  g_cell_begin = cell_begin; g_ghost_ids = (nown > 0 && nown < ncells) ? ghost_ids : nullptr; a.oT = oT; a.n = n; a.gc = &c;
  a.gdofs = gdofs; a.dvals = dvals; a.coords = coords; a.cellv = cellv; a.ptr = ptr; a.grank = grank; a.grankp = grankp; a.adjptr = adjptr; a.adj = adj;
  a.x = x; a.cm = cm.data(); a.cr = cr.data(); a.nz = nz; a.b = b; a.flags = flags;
  return nn == 9 ? run<9, 4, 4>(a, csr != 0, grid) : run<6, 3, 3>(a, csr != 0, grid);
}
