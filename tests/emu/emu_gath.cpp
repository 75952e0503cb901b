// emu_gath.cpp -- TEST INFRASTRUCTURE ONLY: host emulation build of k_gath (see cuda_emu.h).
// Built by tests/test_gath_emu.py:  g++ -O1 -std=c++17 -shared -fPIC -pthread -o _emu_gath.so emu_gath.cpp
#include "cuda_emu.h"

#include "../../gemotion.jl_b200/csrc/gm_cart_tables.h"
#include "../../gemotion.jl_b200/csrc/gm_gath_kernel.cuh"

template <int NG>
static int run(CartArgs a, bool csr, bool newt) {
  const unsigned nt = (3 * NG + GA_NAUX) * 32;
  const unsigned gx = (a.nx + 1 + GA_XCH - 1) / GA_XCH, gy = (a.row_end - a.row_begin + 1 + 4 * NG - 1) / (4 * NG);
  const size_t sm = sizeof(GaSmem<NG>);
  if (csr) return newt ? emu_launch(k_gath<NG, true, true>, a, gx, gy, nt, GA_NAUX * 32, sm) : emu_launch(k_gath<NG, true, false>, a, gx, gy, nt, GA_NAUX * 32, sm);
  return newt ? emu_launch(k_gath<NG, false, true>, a, gx, gy, nt, GA_NAUX * 32, sm) : emu_launch(k_gath<NG, false, false>, a, gx, gy, nt, GA_NAUX * 32, sm);
}

extern "C" int emu_gath_run(int nx, int ny, int row_begin, int row_end, const int32_t *gdofs, const double *dvals,
                            const int64_t *ptr, const int32_t *cls, int nclass, const uint8_t *ctab, const double *x, double *nz,
                            double *b, unsigned flags, int csr, const double *prm /*Pr Ra n reg gx gy js*/, double hx, double hy,
                            const double *w, const double *phi, const double *dphi, const double *psi, int ng, int64_t *diag) {
  static CartPairTab tab[81];
  static CartCell cell;
  gm_cart_fill_tables(w, phi, dphi, psi, hx, hy, prm[0], prm[1], prm[4], prm[5], prm[6], tab, &cell);
  std::vector<uint8_t> ptab((size_t)nclass * 81 * 16);
  gm_cart_pair_ranks(nclass, ctab, ptab.data());
  const bool newt = prm[2] == 1.0;
  CartArgs a;
  a.nx = nx; a.ny = ny; a.row_begin = row_begin; a.row_end = row_end;
  a.gdofs = gdofs; a.dvals = dvals; a.ptr = ptr; a.cls = cls; a.ctab = ctab; a.ptab = ptab.data(); a.visc = nullptr;
  a.tab = tab; a.cc = &cell; a.x = x; a.nz = nz; a.b = b; a.flags = flags;
  a.Pr = prm[0]; a.RaPr = prm[1] * prm[0]; a.gx = prm[4]; a.gy = prm[5]; a.reg = prm[3]; a.n = prm[2]; a.js = prm[6];
  g_emu_bulk_misaligned = 0;
  int rc = ng == 1 ? run<1>(a, csr != 0, newt) : (ng == 2 ? run<2>(a, csr != 0, newt) : run<3>(a, csr != 0, newt));
  if (diag) diag[0] = g_emu_bulk_misaligned;
  return rc;
}
