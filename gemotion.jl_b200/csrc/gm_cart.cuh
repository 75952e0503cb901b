// gm_cart.cuh -- structured (Cartesian) numeric path.  Placeholder: filled in by the
// owner-computes row-gather kernels; until then the scatter path serves Cartesian meshes too.
#pragma once
#include "gm_internal.h"

static int gm_cart_create(gm_handle *h) { (void)h; gm_set_error("structured path not built"); return GM_ERR_STATE; }
static int gm_cart_after_pattern(gm_handle *h) { (void)h; return GM_OK; }
static void gm_cart_destroy(gm_handle *h) { (void)h; }
static int gm_cart_run(gm_handle *h, const double *d_x, double *d_nz, double *d_b, unsigned flags, int64_t *launches) {
  (void)h; (void)d_x; (void)d_nz; (void)d_b; (void)flags; (void)launches;
  return GM_ERR_STATE;
}
