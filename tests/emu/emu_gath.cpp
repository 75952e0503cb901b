// emu_gath.cpp -- TEST INFRASTRUCTURE ONLY: host emulation build of k_gath (see cuda_emu.h).
// Built by tests/test_gath_emu.py:  g++ -O1 -std=c++17 -shared -fPIC -pthread -o _emu_gath.so emu_gath.cpp
#include "cuda_emu.h"

#include "../../gemotion.jl_b200/csrc/gm_cart_tables.h"
#include "../../gemotion.jl_b200/csrc/gm_gath_kernel.cuh"
#include "../../gemotion.jl_b200/csrc/gm_mma_kernel.cuh"

// viscosity pre-pass, same arithmetic as k_ct_visc (gm_cart.cuh)
static void emu_visc(int64_t ncells, const int32_t *g, const double *dvals, const double *x, const CartCell *cc, double reg,
                     double n, double *visc) {
  for (int64_t t = 0; t < ncells * 4; ++t) {
    const int64_t c = t >> 2;
    const int q = (int)(t & 3);
    double G00 = 0, G01 = 0, G10 = 0, G11 = 0;
    for (int i = 0; i < 9; ++i) {
      const int32_t d0 = g[c * 30 + i], d1 = g[c * 30 + 9 + i];
      const double ux = d0 > 0 ? x[d0 - 1] : dvals[-d0 - 1];
      const double uy = d1 > 0 ? x[d1 - 1] : dvals[-d1 - 1];
      const double dx = cc->sdx[q][i], dy = cc->sdy[q][i];
      G00 += dx * ux; G01 += dx * uy; G10 += dy * ux; G11 += dy * uy;
    }
    const double D01 = 0.5 * (G01 + G10);
    const double gam = reg + 2.0 * (G00 * G00 + 2.0 * D01 * D01 + G11 * G11);
    visc[t * 2] = pow(gam, 0.5 * (n - 1.0));
    visc[t * 2 + 1] = 0.5 * (n - 1.0) * pow(gam, 0.5 * (n - 3.0)) * 4.0;
  }
}

template <int NG, int XCH>
static int run(CartArgs a, bool csr, bool newt) {
  const unsigned nt = (3 * NG + GA_NAUX) * 32;
  const unsigned gx = (a.nx + 1 + XCH - 1) / XCH, gy = (a.row_end - a.row_begin + 1 + 4 * NG - 1) / (4 * NG);
  const size_t sm = sizeof(GaSmem<NG>);
  if (csr) return newt ? emu_launch(k_gath<NG, true, true, XCH>, a, gx, gy, nt, GA_NAUX * 32, sm) : emu_launch(k_gath<NG, true, false, XCH>, a, gx, gy, nt, GA_NAUX * 32, sm);
  return newt ? emu_launch(k_gath<NG, false, true, XCH>, a, gx, gy, nt, GA_NAUX * 32, sm) : emu_launch(k_gath<NG, false, false, XCH>, a, gx, gy, nt, GA_NAUX * 32, sm);
}

static int run_mma(const CartArgs &a, bool csr, bool newt, int xch) {
  MmArgs ma;
  ma.a = a;
  memcpy(ma.out, MM_OUT, sizeof ma.out);
  ma.xch = xch;
  ma.strip0 = 0;
  const unsigned gx = (a.nx + 1 + xch - 1) / xch, gy = (a.row_end - a.row_begin + 1 + MM_QH - 1) / MM_QH;
  const size_t sm = sizeof(MmSmem);
  if (csr) return newt ? emu_launch(k_mma<true, true>, ma, gx, gy, MM_NT, MM_NAW * 32, sm) : emu_launch(k_mma<true, false>, ma, gx, gy, MM_NT, MM_NAW * 32, sm);
  return newt ? emu_launch(k_mma<false, true>, ma, gx, gy, MM_NT, MM_NAW * 32, sm) : emu_launch(k_mma<false, false>, ma, gx, gy, MM_NT, MM_NAW * 32, sm);
}

extern "C" int emu_gath_run(int nx, int ny, int row_begin, int row_end, const int32_t *gdofs, const double *dvals,
                            const int64_t *ptr, const int32_t *cls, int nclass, const uint8_t *ctab, const double *x, double *nz,
                            double *b, unsigned flags, int csr, const double *prm /*Pr Ra n reg gx gy js*/, double hx, double hy,
                            const double *w, const double *phi, const double *dphi, const double *psi, int ng, int xch, int64_t *diag) {
  static CartPairTab tab[81];
  static CartCell cell;
  gm_cart_fill_tables(w, phi, dphi, psi, hx, hy, prm[0], prm[1], prm[4], prm[5], prm[6], tab, &cell);
  std::vector<uint8_t> ptab((size_t)nclass * 81 * 16);
  gm_cart_pair_ranks(nclass, ctab, ptab.data());
  const int64_t ncells = (int64_t)nx * ny;
  std::vector<double> visc((size_t)ncells * 8, 0.0);
  const bool newt = prm[2] == 1.0;
  if (!newt) emu_visc(ncells, gdofs, dvals, x, &cell, prm[3], prm[2], visc.data());
  CartArgs a;
  a.nx = nx; a.ny = ny; a.row_begin = row_begin; a.row_end = row_end;
  a.gdofs = gdofs; a.dvals = dvals; a.ptr = ptr; a.cls = cls; a.ctab = ctab; a.ptab = ptab.data(); a.visc = visc.data();
  a.tab = tab; a.cc = &cell; a.x = x; a.nz = nz; a.b = b; a.flags = flags;
  a.Pr = prm[0]; a.RaPr = prm[1] * prm[0]; a.gx = prm[4]; a.gy = prm[5]; a.reg = prm[3]; a.n = prm[2]; a.js = prm[6];
  g_emu_bulk_misaligned = 0;
  if (ng == 0) {  // k_mma (gm_mma_kernel.cuh)
    int rc = run_mma(a, csr != 0, newt, xch);
    if (diag) diag[0] = g_emu_bulk_misaligned;
    return rc;
  }
  // the configurations the library instantiates (x-chunks of 256 / 128 columns), plus 32-column chunks to exercise chunk seams on small meshes
  int rc = ng == 1 ? (xch == 256 ? run<1, 256>(a, csr != 0, newt) : (xch == 128 ? run<1, 128>(a, csr != 0, newt) : run<1, 32>(a, csr != 0, newt)))
                   : (xch == 256 ? run<2, 256>(a, csr != 0, newt) : (xch == 128 ? run<2, 128>(a, csr != 0, newt) : run<2, 32>(a, csr != 0, newt)));
  if (diag) diag[0] = g_emu_bulk_misaligned;
  return rc;
}
