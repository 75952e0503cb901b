// gm_const.h -- reference tables + parameters as seen by the generic-path kernels (plain C++: also compiled by the
// host emulation build of tests/emu).
#pragma once
#include <stdint.h>

#define GM_MAXQ 4
#define GM_MAXNN 9
#define GM_MAXNV 4
#define GM_MAXLD (3 * GM_MAXNN + 3)

// k_scatter reads it from __constant__ memory (c_gm), the gather kernels from a per-handle device copy.
struct GmConst {
  double w[GM_MAXQ];
  double phi[GM_MAXQ * GM_MAXNN];
  double dphi[GM_MAXQ * GM_MAXNN * 2];
  double psi[GM_MAXQ * 3];
  double dgphi[GM_MAXQ * GM_MAXNV * 2];
  double Pr, Ra, n, reg, gx, gy, jac_scaling;
  double hx, hy;
  int cartesian, nx, ny, layout;
};
