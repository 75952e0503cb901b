// gm_lib.cu -- C ABI of libgemotion_b200.so (see include/gemotion_b200.h).
// Host orchestration only; the kernels live in gm_pattern.cuh / gm_scatter.cuh / gm_cart.cuh.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <new>
#include <vector>

#include "gm_internal.h"
#include "gm_pattern.cuh"
#include "gm_scatter.cuh"
#include "gm_gather.cuh"
#include "gm_multi.cuh"
#include "gm_cart.cuh"
#include "gm_post.cuh"

static thread_local char g_err[1024] = "";

void gm_set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}

extern "C" int gm_version(void) { return GM_VERSION; }
extern "C" const char *gm_last_error(void) { return g_err; }

// ------------------------------------------------------------------------------------------
// greedy colouring of a generic mesh (host, one-time): no two cells of a colour share a dof
// ------------------------------------------------------------------------------------------
static int gm_color_generic(gm_handle *h, const gm_desc *d) {
  const int nn = h->nn;
  const int64_t nc = h->ncells;
  // Cells conflict iff they share a P2 node; the T table numbers every node (free >0, Dirichlet <0),
  // the U table may have different Dirichlet sets, so use both.
  auto build = [&](const int32_t *tab, int per, int64_t nfree, int64_t ndiri, std::vector<int64_t> &ptr,
                   std::vector<int32_t> &adj) {
    const int64_t nnode = nfree + ndiri;
    ptr.assign(nnode + 1, 0);
    auto key = [&](int32_t v) -> int64_t { return v > 0 ? v - 1 : nfree + (-(int64_t)v - 1); };
    for (int64_t c = 0; c < nc; ++c)
      for (int l = 0; l < per; ++l) ptr[key(tab[c * per + l]) + 1]++;
    for (int64_t k = 0; k < nnode; ++k) ptr[k + 1] += ptr[k];
    adj.resize(ptr[nnode]);
    std::vector<int64_t> cur(ptr.begin(), ptr.end() - 1);
    for (int64_t c = 0; c < nc; ++c)
      for (int l = 0; l < per; ++l) adj[cur[key(tab[c * per + l])]++] = (int32_t)c;
  };
  std::vector<int64_t> pt;
  std::vector<int32_t> at;
  build(d->cell_dofs_t, nn, d->n_free[2], d->n_diri[2], pt, at);
  std::vector<int> color(nc, -1);
  int ncol = 0;
  std::vector<char> used;
  for (int64_t c = 0; c < nc; ++c) {
    used.assign(ncol + 1, 0);
    for (int l = 0; l < nn; ++l) {
      const int32_t v = d->cell_dofs_t[c * nn + l];
      const int64_t k = v > 0 ? v - 1 : d->n_free[2] + (-(int64_t)v - 1);
      for (int64_t p = pt[k]; p < pt[k + 1]; ++p) {
        const int cc = color[at[p]];
        if (cc >= 0) used[cc] = 1;
      }
    }
    int k = 0;
    while (k < ncol && used[k]) ++k;
    if (k == ncol) ++ncol;
    color[c] = k;
  }
  h->ncolors = ncol;
  h->color_ptr.assign(ncol + 1, 0);
  for (int64_t c = 0; c < nc; ++c) h->color_ptr[color[c] + 1]++;
  for (int k = 0; k < ncol; ++k) h->color_ptr[k + 1] += h->color_ptr[k];
  std::vector<int32_t> cells(nc);
  std::vector<int64_t> cur(h->color_ptr.begin(), h->color_ptr.end() - 1);
  for (int64_t c = 0; c < nc; ++c) cells[cur[color[c]]++] = (int32_t)c;
  int rc = gm_dalloc(h, &h->color_cells, nc);
  if (rc) return rc;
  GM_CUDA(cudaMemcpy(h->color_cells, cells.data(), sizeof(int32_t) * nc, cudaMemcpyHostToDevice));
  return GM_OK;
}

template <typename T>
static int gm_upload(gm_handle *h, T **dst, const T *src, int64_t count) {
  int rc = gm_dalloc(h, dst, count);
  if (rc) return rc;
  if (count > 0 && src) GM_CUDA(cudaMemcpy(*dst, src, sizeof(T) * (size_t)count, cudaMemcpyHostToDevice));
  return GM_OK;
}

extern "C" int gm_create(const gm_desc *d, gm_handle **out) {
  if (!d || !out) { gm_set_error("gm_create: null argument"); return GM_ERR_ARG; }
  *out = nullptr;
  if (d->struct_size != (int32_t)sizeof(gm_desc)) {
    gm_set_error("gm_create: gm_desc size mismatch (%d vs %d)", d->struct_size, (int)sizeof(gm_desc));
    return GM_ERR_ARG;
  }
  const bool quad = d->cell_type == GM_QUAD;
  if (!(quad || d->cell_type == GM_TRI)) { gm_set_error("gm_create: unknown cell type %d", d->cell_type); return GM_ERR_ARG; }
  const int nn = quad ? 9 : 6, nv = quad ? 4 : 3, nq = quad ? 4 : 3;
  if (d->nn != nn || d->nv != nv || d->nq != nq) {
    gm_set_error("gm_create: expected nn=%d nv=%d nq=%d (degree-2 measure) for this cell type, got %d %d %d", nn, nv, nq,
                 d->nn, d->nv, d->nq);
    return GM_ERR_ARG;
  }
  if (d->ncells <= 0 || !d->cell_dofs_u || !d->cell_dofs_p || !d->cell_dofs_t || !d->w || !d->phi || !d->dphi || !d->psi) {
    gm_set_error("gm_create: missing table");
    return GM_ERR_ARG;
  }
  if (d->cartesian) {
    if (!quad || (int64_t)d->nx * d->ny != d->ncells || !(d->h[0] > 0) || !(d->h[1] > 0) || d->ghost_lo < 0 || d->ghost_lo > 1 ||
        d->ghost_hi < 0 || d->ghost_hi > 1 || d->ny - d->ghost_lo - d->ghost_hi < 1) {
      gm_set_error("gm_create: inconsistent Cartesian description");
      return GM_ERR_ARG;
    }
    if ((d->ghost_lo || d->ghost_hi) && (d->path == GM_PATH_SCATTER || d->path == GM_PATH_GATHER)) {
      gm_set_error("gm_create: row blocks with ghost rows need the structured path");
      return GM_ERR_ARG;
    }
  } else if (d->ghost_lo || d->ghost_hi) {
    gm_set_error("gm_create: ghost ROWS describe Cartesian row blocks; unstructured meshes are partitioned by cell ranges (n_own_cells, ghost_cell_ids)");
    return GM_ERR_ARG;
  } else if (!d->coords || !d->cell_vertices || !d->dgphi || d->nverts <= 0) {
    gm_set_error("gm_create: unstructured mesh needs coords, cell_vertices and dgphi");
    return GM_ERR_ARG;
  }
  if (d->n_own_cells < 0 || d->n_own_cells > d->ncells || (d->n_own_cells > 0 && d->n_own_cells < d->ncells && !d->ghost_cell_ids)) {
    gm_set_error("gm_create: inconsistent cell partition (n_own_cells / ghost_cell_ids)");
    return GM_ERR_ARG;
  }
  if (d->n_own_cells > 0 && d->n_own_cells < d->ncells && (d->cartesian || d->path == GM_PATH_SCATTER || d->path == GM_PATH_CART)) {
    gm_set_error("gm_create: ghost cells need the gather path on an unstructured description");
    return GM_ERR_ARG;
  }
  if (d->layout != GM_LAYOUT_CSC && d->layout != GM_LAYOUT_CSR) { gm_set_error("gm_create: bad layout"); return GM_ERR_ARG; }
  const int64_t ntot = d->n_free[0] + d->n_free[1] + d->n_free[2];
  if (ntot <= 0 || ntot >= (int64_t)INT32_MAX) { gm_set_error("gm_create: n_free out of range"); return GM_ERR_ARG; }
  for (int f = 0; f < 3; ++f)
    if (d->n_diri[f] > 0 && !(f == 0 ? d->diri_u : f == 1 ? d->diri_p : d->diri_t)) {
      gm_set_error("gm_create: Dirichlet values missing for field %d", f);
      return GM_ERR_ARG;
    }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || d->device < 0 || d->device >= ndev) {
    gm_set_error("gm_create: no usable CUDA device (count=%d, requested %d); this library has no CPU path", ndev, d->device);
    return GM_ERR_CUDA;
  }
  GM_CUDA(cudaSetDevice(d->device));
  cudaDeviceProp prop;
  GM_CUDA(cudaGetDeviceProperties(&prop, d->device));
  if (prop.major < 10) {
    gm_set_error("gm_create: device %d is sm_%d%d; this library is built for sm_100a only", d->device, prop.major, prop.minor);
    return GM_ERR_CUDA;
  }
  gm_handle *h = new (std::nothrow) gm_handle();
  if (!h) return GM_ERR_OOM;
  h->d = *d;
  h->dev = d->device;
  h->nn = nn; h->nv = nv; h->nq = nq; h->ld = 3 * nn + 3;
  h->ncells = d->ncells;
  h->n_total = ntot;
  h->off[0] = 0; h->off[1] = d->n_free[0]; h->off[2] = d->n_free[0] + d->n_free[1];
  h->n_diri_total = d->n_diri[0] + d->n_diri[1] + d->n_diri[2];
  memset(&h->times, 0, sizeof h->times);
  int rc = GM_OK;
  auto fail = [&](int code) { gm_destroy(h); return code; };
#define TRY(x) do { rc = (x); if (rc != GM_OK) return fail(rc); } while (0)
#define TRYC(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { gm_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); return fail(e__ == cudaErrorMemoryAllocation ? GM_ERR_OOM : GM_ERR_CUDA); } } while (0)
  TRYC(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
  {
    int nsm = 0;
    if (cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, d->device) == cudaSuccess && nsm > 0) h->num_sms = nsm;
  }
  h->own_stream = true;
  for (int k = 0; k < 4; ++k) TRYC(cudaEventCreate(&h->ev[k]));
  // constants
  GmConst &c = h->hc;
  memset(&c, 0, sizeof c);
  memcpy(c.w, d->w, sizeof(double) * nq);
  memcpy(c.phi, d->phi, sizeof(double) * nq * nn);
  memcpy(c.dphi, d->dphi, sizeof(double) * nq * nn * 2);
  memcpy(c.psi, d->psi, sizeof(double) * nq * 3);
  if (d->dgphi) memcpy(c.dgphi, d->dgphi, sizeof(double) * nq * nv * 2);
  c.cartesian = d->cartesian; c.nx = d->nx; c.ny = d->ny; c.hx = d->h[0]; c.hy = d->h[1]; c.layout = d->layout;
  h->prm = gm_params{1.0, 1.0, 1.0, 1e-3, 0.0, 1.0, 1.0};
  // tables -> device, combined signed global ids
  {
    int32_t *du = nullptr, *dp = nullptr, *dt = nullptr;
    const int64_t keep = h->device_bytes;
    TRY(gm_upload(h, &du, d->cell_dofs_u, h->ncells * 2 * nn));
    TRY(gm_upload(h, &dp, d->cell_dofs_p, h->ncells * 3));
    TRY(gm_upload(h, &dt, d->cell_dofs_t, h->ncells * nn));
    TRY(gm_dalloc(h, &h->gdofs, h->ncells * h->ld));
    k_combine_dofs<<<gm_blocks(h->ncells * h->ld, 256), 256, 0, h->stream>>>(h->ncells, nn, du, dp, dt, h->off[1], h->off[2],
                                                                             d->n_diri[0], d->n_diri[0] + d->n_diri[1],
                                                                             h->gdofs);
    TRYC(cudaStreamSynchronize(h->stream));
    cudaFree(du); cudaFree(dp); cudaFree(dt);
    h->device_bytes = keep + (int64_t)sizeof(int32_t) * h->ncells * h->ld;
  }
  {
    std::vector<double> dv((size_t)std::max<int64_t>(h->n_diri_total, 1), 0.0);
    int64_t o = 0;
    const double *src[3] = {d->diri_u, d->diri_p, d->diri_t};
    for (int f = 0; f < 3; ++f) {
      if (d->n_diri[f] > 0) memcpy(dv.data() + o, src[f], sizeof(double) * d->n_diri[f]);
      o += d->n_diri[f];
    }
    TRY(gm_upload(h, &h->dvals, dv.data(), (int64_t)dv.size()));
  }
  if (!d->cartesian) {
    TRY(gm_upload(h, &h->coords, d->coords, d->nverts * 2));
    if (d->vertex_index_base == 0) {
      TRY(gm_upload(h, &h->cellv, d->cell_vertices, h->ncells * nv));
    } else {
      std::vector<int32_t> cv((size_t)(h->ncells * nv));
      for (size_t k = 0; k < cv.size(); ++k) cv[k] = d->cell_vertices[k] - d->vertex_index_base;
      TRY(gm_upload(h, &h->cellv, cv.data(), (int64_t)cv.size()));
    }
    if (d->path == GM_PATH_SCATTER) TRY(gm_color_generic(h, d));
    h->n_own_cells = d->n_own_cells > 0 ? d->n_own_cells : h->ncells;
    if (h->n_own_cells < h->ncells) TRY(gm_upload(h, &h->ghost_ids, d->ghost_cell_ids, h->ncells - h->n_own_cells));
  } else {
    h->ncolors = 4;
  }
  TRY(gm_dalloc(h, &h->d_x, h->n_total));
  TRY(gm_dalloc(h, &h->d_b, h->n_total));
  TRYC(cudaMemset(h->d_x, 0, sizeof(double) * (size_t)h->n_total));  // (entries no kernel writes / no copy fills stay finite)
  TRYC(cudaMemset(h->d_b, 0, sizeof(double) * (size_t)h->n_total));
  // generic meshes: owner-computes gather unless the caller asks for the colour-ordered scatter
  h->active_path = d->path == GM_PATH_SCATTER ? GM_PATH_SCATTER : GM_PATH_GATHER;
  if (d->path < GM_PATH_AUTO || d->path > GM_PATH_GATHER) { gm_set_error("gm_create: bad path"); return fail(GM_ERR_ARG); }
  if (d->cartesian && (d->path == GM_PATH_AUTO || d->path == GM_PATH_CART)) {
    rc = gm_cart_create(h);  // verifies the tables; falls back to the generic gather if they do not fit
    if (rc != GM_OK && d->path == GM_PATH_CART) return fail(rc);
  } else if (d->path == GM_PATH_CART) {
    gm_set_error("gm_create: GM_PATH_CART needs a Cartesian description");
    return fail(GM_ERR_ARG);
  }
#undef TRY
#undef TRYC
  // scrub host pointers: the library keeps none
  h->d.coords = nullptr; h->d.cell_vertices = nullptr; h->d.cell_dofs_u = h->d.cell_dofs_p = h->d.cell_dofs_t = nullptr;
  h->d.diri_u = h->d.diri_p = h->d.diri_t = nullptr;
  h->d.w = h->d.phi = h->d.dphi = h->d.psi = h->d.dgphi = nullptr;
  h->d.ghost_cell_ids = nullptr;
  if (h->n_own_cells == 0) h->n_own_cells = h->ncells;
  *out = h;
  return GM_OK;
}

extern "C" void gm_destroy(gm_handle *h) {
  if (!h) return;
  cudaSetDevice(h->dev);
  gm_multi_destroy(h);
  gm_post_destroy(h);
  gm_cart_destroy(h);
  gm_gather_destroy(h);
  cudaFree(h->gdofs); cudaFree(h->dvals); cudaFree(h->coords); cudaFree(h->cellv);
  cudaFree(h->ptr); cudaFree(h->idx); cudaFree(h->rank); cudaFree(h->color_cells);
  cudaFree(h->d_x); cudaFree(h->d_b); cudaFree(h->d_nz); cudaFree(h->ghost_ids);
  for (int k = 0; k < 4; ++k)
    if (h->ev[k]) cudaEventDestroy(h->ev[k]);
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

extern "C" int gm_set_params(gm_handle *h, double Pr, double Ra, double n, double reg, double gx, double gy,
                             double jac_scaling) {
  if (!h) { gm_set_error("gm_set_params: null handle"); return GM_ERR_ARG; }
  if (!(reg >= 0)) { gm_set_error("gm_set_params: reg must be >= 0"); return GM_ERR_ARG; }
  h->prm = gm_params{Pr, Ra, n, reg, gx, gy, jac_scaling};
  return GM_OK;
}

extern "C" int gm_set_stream(gm_handle *h, void *stream) {
  if (!h) { gm_set_error("gm_set_stream: null handle"); return GM_ERR_ARG; }
  GM_CUDA(cudaSetDevice(h->dev));
  if (h->stream) GM_CUDA(cudaStreamSynchronize(h->stream));  // work and table uploads ordered on the old stream are complete before the switch
  if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
  h->stream = (cudaStream_t)stream;
  h->own_stream = false;
  return GM_OK;
}

// dof ranges touched by the local cells (a rank of a partition sees a fraction of the global numbering: Gridap numbers seven
// groups -- U vertices / edges / cells, P, T vertices / edges / cells -- one after another, so a row block touches about seven
// runs).  Runs closer than min(64 Ki, n / 256) dofs are merged; more than 64 runs -> one full copy.
__global__ void k_mark_touched(int64_t nent, const int32_t *__restrict__ g, uint8_t *__restrict__ flag) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t < nent && g[t] > 0) flag[g[t] - 1] = 1;
}
__global__ void k_run_bounds(int64_t n, const uint8_t *__restrict__ flag, int cap, int *__restrict__ cnt, int64_t *__restrict__ out) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i > n) return;
  const uint8_t cur = i < n ? flag[i] : 0, prev = i > 0 ? flag[i - 1] : 0;
  if (cur != prev) {
    const int k = atomicAdd(cnt, 1);
    if (k < cap) out[k] = i;
  }
}
static int gm_touched_runs(gm_handle *h) {
  h->xruns.clear();
  if (h->d.nranks <= 1) return GM_OK;
  const int64_t n = h->n_total, nent = h->ncells * h->ld;
  const int CAP = 8192;
  uint8_t *flag = nullptr;
  int *cnt = nullptr, hc = 0;
  int64_t *out = nullptr;
  std::vector<int64_t> hb;
  cudaError_t e = cudaMalloc(&flag, (size_t)n + 1);
  if (e == cudaSuccess) e = cudaMalloc(&cnt, sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc(&out, sizeof(int64_t) * CAP);
  if (e == cudaSuccess) e = cudaMemsetAsync(flag, 0, (size_t)n + 1, h->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(cnt, 0, sizeof(int), h->stream);
  if (e == cudaSuccess) {
    k_mark_touched<<<gm_blocks(nent, 256), 256, 0, h->stream>>>(nent, h->gdofs, flag);
    k_run_bounds<<<gm_blocks(n + 1, 256), 256, 0, h->stream>>>(n, flag, CAP, cnt, out);
    e = cudaMemcpyAsync(&hc, cnt, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  if (e == cudaSuccess && hc > 0 && hc <= CAP && (hc & 1) == 0) {
    hb.resize((size_t)hc);
    e = cudaMemcpy(hb.data(), out, sizeof(int64_t) * hc, cudaMemcpyDeviceToHost);
  }
  cudaFree(flag); cudaFree(cnt); cudaFree(out);
  if (e != cudaSuccess) { gm_set_error("gm_pattern: %s", cudaGetErrorString(e)); return GM_ERR_CUDA; }
  if (hb.empty()) return GM_OK;  // (nothing touched / too fragmented: full copies)
  std::sort(hb.begin(), hb.end());
  std::vector<std::pair<int64_t, int64_t>> runs;
  const int64_t gap = std::max<int64_t>(16, std::min<int64_t>(65536, n / 256));
  for (size_t k = 0; k + 1 < hb.size(); k += 2) {
    if (!runs.empty() && hb[k] - runs.back().second < gap) runs.back().second = hb[k + 1];
    else runs.emplace_back(hb[k], hb[k + 1]);
  }
  if (runs.size() <= 64) h->xruns = runs;
  return GM_OK;
}

extern "C" int gm_pattern(gm_handle *h, int64_t *nnz) {
  if (!h) { gm_set_error("gm_pattern: null handle"); return GM_ERR_ARG; }
  GM_CUDA(cudaSetDevice(h->dev));
  if (!h->has_pattern) {
    int rc = gm_build_pattern(h);
    if (rc) return rc;
    if (h->active_path == GM_PATH_CART) {
      rc = gm_cart_after_pattern(h);
      if (rc) return rc;
    }
    if (h->active_path == GM_PATH_GATHER) {
      rc = gm_gather_after_pattern(h);
      if (rc) {  // (e.g. the cell-matrix workspace does not fit): no half-built state; gm_assemble then reports the missing workspaces
        gm_gather *g = (gm_gather *)h->gather;
        if (g) { cudaFree(g->d_const); cudaFree(g->cm); cudaFree(g->cr); cudaFree(g->desc); delete g; }
        h->gather = nullptr;
        return rc;
      }
    }
    rc = gm_touched_runs(h);
    if (rc) return rc;
    if (h->d.ghost_lo || h->d.ghost_hi) {
      if (h->active_path != GM_PATH_CART) { gm_set_error("gm_pattern: row blocks need the structured path"); return GM_ERR_ARG; }
      rc = gm_multi_after_pattern(h);
      if (rc) return rc;
    }
  }
  if (nnz) *nnz = h->nnz;
  return GM_OK;
}

__global__ void k_widen(const int32_t *__restrict__ s, int64_t *__restrict__ d, int64_t n, int64_t add) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t < n) d[t] = (int64_t)s[t] + add;
}
__global__ void k_addbase(const int64_t *__restrict__ s, int64_t *__restrict__ d, int64_t n, int64_t add) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t < n) d[t] = s[t] + add;
}

extern "C" int gm_get_pattern(gm_handle *h, int64_t *ptr, int64_t *idx, int32_t index_base) {
  if (!h || !ptr) { gm_set_error("gm_get_pattern: null argument"); return GM_ERR_ARG; }
  if (!h->has_pattern) { gm_set_error("gm_get_pattern: call gm_pattern first"); return GM_ERR_STATE; }
  if (idx && !h->idx) { gm_set_error("gm_get_pattern: the index array was released (huge problem)"); return GM_ERR_STATE; }
  GM_CUDA(cudaSetDevice(h->dev));
  const int64_t n = h->n_total;
  const int64_t CH = 1 << 24;
  int64_t *tmp = nullptr;
  GM_CUDA(cudaMalloc(&tmp, sizeof(int64_t) * CH));
  auto chunked = [&](auto launch, int64_t total, int64_t *dst) -> int {
    for (int64_t o = 0; o < total; o += CH) {
      const int64_t m = std::min(CH, total - o);
      launch(o, m);
      cudaError_t e = cudaMemcpyAsync(dst + o, tmp, sizeof(int64_t) * m, cudaMemcpyDeviceToHost, h->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
      if (e != cudaSuccess) { gm_set_error("gm_get_pattern: %s", cudaGetErrorString(e)); return GM_ERR_CUDA; }
    }
    return GM_OK;
  };
  int rc = chunked([&](int64_t o, int64_t m) { k_addbase<<<gm_blocks(m, 256), 256, 0, h->stream>>>(h->ptr + o, tmp, m, index_base); }, n + 1, ptr);
  if (rc == GM_OK && idx)
    rc = chunked([&](int64_t o, int64_t m) { k_widen<<<gm_blocks(m, 256), 256, 0, h->stream>>>(h->idx + o, tmp, m, index_base); }, h->nnz, idx);
  cudaFree(tmp);
  return rc;
}

// gathers the lists of selected major dofs into dense buffers (one thread block per row)
__global__ void k_get_rows(int64_t nrows, const int64_t *__restrict__ rows, const int64_t *__restrict__ off,
                           const int64_t *__restrict__ ptr, const int32_t *__restrict__ idx, const double *__restrict__ nz,
                           const double *__restrict__ b, int64_t *__restrict__ idx_out, double *__restrict__ val_out,
                           double *__restrict__ b_out) {
  const int64_t s = blockIdx.x;
  if (s >= nrows) return;
  const int64_t r = rows[s], p0 = ptr[r], len = ptr[r + 1] - p0, o = off[s];
  for (int64_t k = threadIdx.x; k < len; k += blockDim.x) {
    if (idx_out) idx_out[o + k] = idx[p0 + k];
    if (val_out) val_out[o + k] = nz[p0 + k];
  }
  if (b_out && threadIdx.x == 0) b_out[s] = b[r];
}
__global__ void k_row_lens(int64_t nrows, const int64_t *__restrict__ rows, const int64_t *__restrict__ ptr, int64_t *__restrict__ len) {
  const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (s < nrows) len[s] = ptr[rows[s] + 1] - ptr[rows[s]];
}

extern "C" int gm_get_rows(gm_handle *h, int64_t nrows, const int64_t *rows, int64_t *row_ptr, int64_t *idx_out, double *val_out,
                           double *b_out, int64_t cap, const double *d_nzval, const double *d_resid) {
  if (!h || !rows || !row_ptr || nrows < 0) { gm_set_error("gm_get_rows: bad argument"); return GM_ERR_ARG; }
  if (!h->has_pattern) { gm_set_error("gm_get_rows: call gm_pattern first"); return GM_ERR_STATE; }
  if (idx_out && !h->idx) { gm_set_error("gm_get_rows: the index array was released"); return GM_ERR_STATE; }
  for (int64_t k = 0; k < nrows; ++k)
    if (rows[k] < 0 || rows[k] >= h->n_total) { gm_set_error("gm_get_rows: row %lld out of range", (long long)rows[k]); return GM_ERR_ARG; }
  GM_CUDA(cudaSetDevice(h->dev));
  row_ptr[0] = 0;
  if (nrows == 0) return GM_OK;
  int64_t *d_rows = nullptr, *d_off = nullptr, *d_idx = nullptr;
  double *d_val = nullptr, *d_b = nullptr;
  int rc = GM_OK;
  std::vector<int64_t> len((size_t)nrows);
#define GR(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { gm_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); rc = e__ == cudaErrorMemoryAllocation ? GM_ERR_OOM : GM_ERR_CUDA; goto done; } } while (0)
  GR(cudaMalloc(&d_rows, 8 * nrows));
  GR(cudaMalloc(&d_off, 8 * nrows));
  GR(cudaMemcpyAsync(d_rows, rows, 8 * nrows, cudaMemcpyHostToDevice, h->stream));
  k_row_lens<<<gm_blocks(nrows, 256), 256, 0, h->stream>>>(nrows, d_rows, h->ptr, d_off);
  GR(cudaMemcpyAsync(len.data(), d_off, 8 * nrows, cudaMemcpyDeviceToHost, h->stream));
  GR(cudaStreamSynchronize(h->stream));
  for (int64_t k = 0; k < nrows; ++k) row_ptr[k + 1] = row_ptr[k] + len[k];
  if (row_ptr[nrows] > cap) { gm_set_error("gm_get_rows: %lld entries, capacity %lld", (long long)row_ptr[nrows], (long long)cap); rc = GM_ERR_ARG; goto done; }
  {
    const int64_t tot = row_ptr[nrows] > 0 ? row_ptr[nrows] : 1;
    GR(cudaMemcpyAsync(d_off, row_ptr, 8 * nrows, cudaMemcpyHostToDevice, h->stream));
    if (idx_out) GR(cudaMalloc(&d_idx, 8 * tot));
    if (val_out && d_nzval) GR(cudaMalloc(&d_val, 8 * tot));
    if (b_out && d_resid) GR(cudaMalloc(&d_b, 8 * nrows));
    k_get_rows<<<(unsigned)nrows, 32, 0, h->stream>>>(nrows, d_rows, d_off, h->ptr, h->idx, d_nzval, d_resid, d_idx, d_val, d_b);
    GR(cudaGetLastError());
    if (d_idx) GR(cudaMemcpyAsync(idx_out, d_idx, 8 * row_ptr[nrows], cudaMemcpyDeviceToHost, h->stream));
    if (d_val) GR(cudaMemcpyAsync(val_out, d_val, 8 * row_ptr[nrows], cudaMemcpyDeviceToHost, h->stream));
    if (d_b) GR(cudaMemcpyAsync(b_out, d_b, 8 * nrows, cudaMemcpyDeviceToHost, h->stream));
    GR(cudaStreamSynchronize(h->stream));
  }
done:
#undef GR
  cudaFree(d_rows); cudaFree(d_off); cudaFree(d_idx); cudaFree(d_val); cudaFree(d_b);
  return rc;
}

// GM_CHECK_FINITE: device scan of the outputs (both entry points); err[0] = 1 + kind (0 nzval, 1 residual), err64[1] = first index seen
__global__ void k_check_finite(const double *__restrict__ v, int64_t n, int kind, unsigned long long *__restrict__ err) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (!isfinite(v[i]) && atomicCAS(&err[0], 0ull, (unsigned long long)(1 + kind)) == 0ull) err[1] = (unsigned long long)i;
}
static int gm_check_finite(gm_handle *h, const double *d_nz, const double *d_b) {
  unsigned long long *err = nullptr, herr[2] = {0, 0};
  GM_CUDA(cudaMalloc(&err, 16));
  GM_CUDA(cudaMemsetAsync(err, 0, 16, h->stream));
  if (d_nz) k_check_finite<<<h->num_sms * 8, 256, 0, h->stream>>>(d_nz, h->nnz, 0, err);
  if (d_b) k_check_finite<<<h->num_sms * 8, 256, 0, h->stream>>>(d_b, h->n_total, 1, err);
  cudaError_t e = cudaMemcpyAsync(herr, err, 16, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(err);
  if (e != cudaSuccess) { gm_set_error("gm_check_finite: %s", cudaGetErrorString(e)); return GM_ERR_CUDA; }
  if (herr[0]) {
    gm_set_error("non-finite %s entry at index %llu", herr[0] == 1 ? "nzval" : "residual", herr[1]);
    return GM_ERR_NONFINITE;
  }
  return GM_OK;
}

// The scatter kernels read reference tables and parameters from ONE __constant__ block per device (c_gm).  Handles on the
// same device run on their own streams, so an upload must not overtake the kernels of the handle that used the block last:
// every upload first makes its stream wait for an event recorded after that handle's last scatter launch.
struct GmConstUse { cudaEvent_t ev = nullptr; bool used = false; };
static GmConstUse g_const_use[64];
static std::mutex g_const_mutex;

static int gm_upload_const(gm_handle *h) {
  GmConst &c = h->hc;
  c.Pr = h->prm.Pr; c.Ra = h->prm.Ra; c.n = h->prm.n; c.reg = h->prm.reg;
  c.gx = h->prm.gx; c.gy = h->prm.gy; c.jac_scaling = h->prm.jac_scaling;
  std::lock_guard<std::mutex> lock(g_const_mutex);
  GmConstUse &u = g_const_use[h->dev & 63];
  if (!u.ev) GM_CUDA(cudaEventCreateWithFlags(&u.ev, cudaEventDisableTiming));
  if (u.used) GM_CUDA(cudaStreamWaitEvent(h->stream, u.ev, 0));
  GM_CUDA(cudaMemcpyToSymbolAsync(c_gm, &c, sizeof(GmConst), 0, cudaMemcpyHostToDevice, h->stream));
  return GM_OK;
}
static int gm_const_release(gm_handle *h) {  // after the last kernel that reads c_gm has been launched
  std::lock_guard<std::mutex> lock(g_const_mutex);
  GmConstUse &u = g_const_use[h->dev & 63];
  GM_CUDA(cudaEventRecord(u.ev, h->stream));
  u.used = true;
  return GM_OK;
}

static int gm_run(gm_handle *h, const double *d_x, double *d_nz, double *d_b, uint32_t flags, int64_t *launches) {
  if (!h->has_pattern) { gm_set_error("gm_assemble: call gm_pattern first"); return GM_ERR_STATE; }
  if ((flags & GM_JACOBIAN) && !d_nz) { gm_set_error("gm_assemble: GM_JACOBIAN without nzval"); return GM_ERR_ARG; }
  if ((flags & GM_RESIDUAL) && !d_b) { gm_set_error("gm_assemble: GM_RESIDUAL without resid"); return GM_ERR_ARG; }
  if (!(flags & (GM_JACOBIAN | GM_RESIDUAL))) { gm_set_error("gm_assemble: nothing to do"); return GM_ERR_ARG; }
  if (h->active_path == GM_PATH_CART) return gm_cart_run(h, d_x, d_nz, d_b, flags, launches);  // (parameters travel as kernel arguments)
  if (h->active_path == GM_PATH_GATHER) return gm_gather_run(h, d_x, d_nz, d_b, flags, launches);
  int rc = gm_upload_const(h);
  if (rc) return rc;
  rc = gm_launch_scatter(h, d_x, d_nz, d_b, flags, launches);
  if (rc) return rc;
  return gm_const_release(h);
}

extern "C" int gm_assemble_device(gm_handle *h, const double *d_x, double *d_nzval, double *d_resid, uint32_t flags,
                                  int32_t sync) {
  if (!h || !d_x) { gm_set_error("gm_assemble_device: null argument"); return GM_ERR_ARG; }
  GM_CUDA(cudaSetDevice(h->dev));
  int64_t launches = 0;
  if (h->multi) ((gm_multi *)h->multi)->timed = false;
  GM_CUDA(cudaEventRecord(h->ev[1], h->stream));
  int rc = gm_run(h, d_x, d_nzval, d_resid, flags, &launches);
  if (rc) return rc;
  GM_CUDA(cudaEventRecord(h->ev[2], h->stream));
  h->times.launches = launches;
  h->times.h2d_ms = h->times.d2h_ms = 0;
  h->times.h2d_bytes = h->times.d2h_bytes = 0;
  h->times.kernel_ms = h->times.exchange_ms = -1.0;  // filled by gm_timing once the events have completed
  h->times_pending = true;
  if (flags & GM_CHECK_FINITE) return gm_check_finite(h, (flags & GM_JACOBIAN) ? d_nzval : nullptr, (flags & GM_RESIDUAL) ? d_resid : nullptr);  // (synchronises)
  if (sync) GM_CUDA(cudaStreamSynchronize(h->stream));
  return GM_OK;
}

extern "C" int gm_assemble(gm_handle *h, const double *x, double *nzval, double *resid, uint32_t flags) {
  if (!h || !x) { gm_set_error("gm_assemble: null argument"); return GM_ERR_ARG; }
  GM_CUDA(cudaSetDevice(h->dev));
  if (!h->has_pattern) { gm_set_error("gm_assemble: call gm_pattern first"); return GM_ERR_STATE; }
  const bool jac = (flags & GM_JACOBIAN) != 0, res = (flags & GM_RESIDUAL) != 0;
  if (jac && !nzval) { gm_set_error("gm_assemble: GM_JACOBIAN without nzval"); return GM_ERR_ARG; }
  if (res && !resid) { gm_set_error("gm_assemble: GM_RESIDUAL without resid"); return GM_ERR_ARG; }
  if (jac && h->d_nz_cap < h->nnz) {
    cudaFree(h->d_nz);
    h->d_nz = nullptr;
    h->d_nz_cap = 0;
    int rc = gm_dalloc(h, &h->d_nz, h->nnz);
    if (rc) return rc;
    h->d_nz_cap = h->nnz;
    GM_CUDA(cudaMemsetAsync(h->d_nz, 0, sizeof(double) * (size_t)h->nnz, h->stream));
  }
  int64_t launches = 0;
  GM_CUDA(cudaEventRecord(h->ev[0], h->stream));
  int64_t xfer = 0;  // dofs copied each way (a rank of a partition: only the ranges its cells touch)
  if (h->xruns.empty()) {
    GM_CUDA(cudaMemcpyAsync(h->d_x, x, sizeof(double) * (size_t)h->n_total, cudaMemcpyHostToDevice, h->stream));
    xfer = h->n_total;
  } else {
    for (const auto &r : h->xruns) {
      GM_CUDA(cudaMemcpyAsync(h->d_x + r.first, x + r.first, sizeof(double) * (size_t)(r.second - r.first), cudaMemcpyHostToDevice, h->stream));
      xfer += r.second - r.first;
    }
  }
  GM_CUDA(cudaEventRecord(h->ev[1], h->stream));
  int rc = gm_run(h, h->d_x, jac ? h->d_nz : nullptr, res ? h->d_b : nullptr, flags, &launches);
  if (rc) return rc;
  GM_CUDA(cudaEventRecord(h->ev[2], h->stream));
  if (flags & GM_CHECK_FINITE) {  // scanned on the device, before the copies
    rc = gm_check_finite(h, jac ? h->d_nz : nullptr, res ? h->d_b : nullptr);
    if (rc) return rc;
  }
  if (jac) GM_CUDA(cudaMemcpyAsync(nzval, h->d_nz, sizeof(double) * (size_t)h->nnz, cudaMemcpyDeviceToHost, h->stream));
  if (res) {
    if (h->xruns.empty()) {
      GM_CUDA(cudaMemcpyAsync(resid, h->d_b, sizeof(double) * (size_t)h->n_total, cudaMemcpyDeviceToHost, h->stream));
    } else {  // entries outside the touched ranges are not this rank's: the caller's array keeps what it held there
      for (const auto &r : h->xruns)
        GM_CUDA(cudaMemcpyAsync(resid + r.first, h->d_b + r.first, sizeof(double) * (size_t)(r.second - r.first), cudaMemcpyDeviceToHost, h->stream));
    }
  }
  GM_CUDA(cudaEventRecord(h->ev[3], h->stream));
  GM_CUDA(cudaStreamSynchronize(h->stream));
  float a = 0, b = 0, c = 0;
  GM_CUDA(cudaEventElapsedTime(&a, h->ev[0], h->ev[1]));
  GM_CUDA(cudaEventElapsedTime(&b, h->ev[1], h->ev[2]));
  GM_CUDA(cudaEventElapsedTime(&c, h->ev[2], h->ev[3]));
  h->times.h2d_ms = a; h->times.kernel_ms = b; h->times.d2h_ms = c;
  h->times.launches = launches;
  h->times.h2d_bytes = 8 * xfer;
  h->times.d2h_bytes = (jac ? 8 * h->nnz : 0) + (res ? 8 * xfer : 0);
  return GM_OK;
}

extern "C" int gm_timing(gm_handle *h, gm_times *t) {
  if (!h || !t) { gm_set_error("gm_timing: null argument"); return GM_ERR_ARG; }
  if (h->times_pending) {  // device entry point: the step may have been asynchronous; wait for its last event now
    GM_CUDA(cudaSetDevice(h->dev));
    GM_CUDA(cudaEventSynchronize(h->ev[2]));
    float ms = 0;
    GM_CUDA(cudaEventElapsedTime(&ms, h->ev[1], h->ev[2]));
    h->times.kernel_ms = ms;
    h->times.exchange_ms = 0;
    gm_multi *m = (gm_multi *)h->multi;
    if (m && m->timed) {
      GM_CUDA(cudaEventSynchronize(m->ev1));
      GM_CUDA(cudaEventElapsedTime(&ms, m->ev0, m->ev1));
      h->times.exchange_ms = ms;  // pack + NCCL send / receive + unpack-add (overlapped with the interior strips by k_mma)
    }
    h->times_pending = false;
  }
  *t = h->times;
  return GM_OK;
}

extern "C" int gm_info(gm_handle *h, int64_t *n_owned, int32_t *active_path, int32_t *ncolors, int64_t *device_bytes) {
  if (!h) { gm_set_error("gm_info: null handle"); return GM_ERR_ARG; }
  if (n_owned) *n_owned = h->n_owned > 0 ? h->n_owned : h->n_total;
  if (active_path) *active_path = h->active_path;
  if (ncolors) *ncolors = h->ncolors;
  if (device_bytes) *device_bytes = h->device_bytes;
  return GM_OK;
}

extern "C" int gm_host_register(void *ptr, int64_t bytes) {
  if (!ptr || bytes <= 0) { gm_set_error("gm_host_register: bad argument"); return GM_ERR_ARG; }
  GM_CUDA(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterDefault));
  return GM_OK;
}
extern "C" int gm_host_unregister(void *ptr) {
  if (!ptr) { gm_set_error("gm_host_unregister: bad argument"); return GM_ERR_ARG; }
  GM_CUDA(cudaHostUnregister(ptr));
  return GM_OK;
}

extern "C" int gm_nccl_unique_id(void *id128) {
  if (!id128) { gm_set_error("gm_nccl_unique_id: null argument"); return GM_ERR_ARG; }
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  GM_NCCL(ncclGetUniqueId(&id));
  memcpy(id128, &id, 128);
  return GM_OK;
}

extern "C" int gm_comm_init(gm_handle *h, const void *id128) {
  if (!h || !id128) { gm_set_error("gm_comm_init: null argument"); return GM_ERR_ARG; }
  if (h->d.nranks <= 1) return GM_OK;
  GM_CUDA(cudaSetDevice(h->dev));
  gm_multi *m = nullptr;
  {
    const int rc0 = gm_multi_get(h, &m);
    if (rc0) return rc0;
  }
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  GM_NCCL(ncclCommInitRank(&m->comm, h->d.nranks, id, h->d.rank));
  return GM_OK;
}

extern "C" int gm_iface_count(gm_handle *h, int32_t side, int64_t *ndoubles) {
  if (!h || !ndoubles) { gm_set_error("gm_iface_count: null argument"); return GM_ERR_ARG; }
  gm_multi *m = (gm_multi *)h->multi;
  const gm_iface *f = m ? (side == 0 ? &m->lo : &m->hi) : nullptr;
  *ndoubles = f ? f->nlist + f->ndof : 0;
  return GM_OK;
}

extern "C" int gm_iface_pack(gm_handle *h, const double *d_nzval, const double *d_resid, double *d_buf, uint32_t flags) {
  if (!h || !d_buf) { gm_set_error("gm_iface_pack: null argument"); return GM_ERR_ARG; }
  gm_multi *m = (gm_multi *)h->multi;
  if (!m) { gm_set_error("gm_iface_pack: handle has no interface"); return GM_ERR_STATE; }
  GM_CUDA(cudaSetDevice(h->dev));
  int rc = gm_iface_run(h, m->lo, false, const_cast<double *>(d_nzval), const_cast<double *>(d_resid), d_buf, flags);
  if (rc) return rc;
  GM_CUDA(cudaStreamSynchronize(h->stream));
  return GM_OK;
}

extern "C" int gm_iface_unpack_add(gm_handle *h, double *d_nzval, double *d_resid, const double *d_buf, uint32_t flags) {
  if (!h || !d_buf) { gm_set_error("gm_iface_unpack_add: null argument"); return GM_ERR_ARG; }
  gm_multi *m = (gm_multi *)h->multi;
  if (!m) { gm_set_error("gm_iface_unpack_add: handle has no interface"); return GM_ERR_STATE; }
  GM_CUDA(cudaSetDevice(h->dev));
  int rc = gm_iface_run(h, m->hi, true, d_nzval, d_resid, const_cast<double *>(d_buf), flags);
  if (rc) return rc;
  GM_CUDA(cudaStreamSynchronize(h->stream));
  return GM_OK;
}

// ------------------------------------------------------------------------------------------
// post-processing assemblies (gm_post.cuh)
// ------------------------------------------------------------------------------------------
extern "C" int gm_post_create(gm_handle *h, const gm_post_desc *d) {
  if (!h || !d) { gm_set_error("gm_post_create: null argument"); return GM_ERR_ARG; }
  if (d->struct_size != (int32_t)sizeof(gm_post_desc)) { gm_set_error("gm_post_create: gm_post_desc size mismatch"); return GM_ERR_ARG; }
  if (h->d.ghost_lo || h->d.ghost_hi) { gm_set_error("gm_post_create: post-processing runs on a single-GPU handle"); return GM_ERR_ARG; }
  if (!d->cell_dofs_psi || !d->cell_dofs_pi || !d->w2 || !d->g2 || !d->dg2 || !d->phi2 || !d->dphi2 || !d->w4 || !d->g4 || !d->dg4 || !d->phi4 ||
      !d->dphi4 || !d->dphi_nodes || !d->dg_nodes || !d->cell_nodes || d->nq2 <= 0 || d->nq4 <= 0 || d->n_nodes <= 0) {
    gm_set_error("gm_post_create: missing table");
    return GM_ERR_ARG;
  }
  GM_CUDA(cudaSetDevice(h->dev));
  gm_post_destroy(h);
  gm_post *p = new (std::nothrow) gm_post();
  if (!p) return GM_ERR_OOM;
  h->post = p;
  const int nn = h->nn, nv = h->nv;
  int rc = GM_OK;
  for (int r = 0; r < 2 && rc == GM_OK; ++r) {
    const int nq = r == 0 ? d->nq2 : d->nq4;
    p->nq[r] = nq;
    const double *src[5] = {r == 0 ? d->w2 : d->w4, r == 0 ? d->g2 : d->g4, r == 0 ? d->dg2 : d->dg4, r == 0 ? d->phi2 : d->phi4, r == 0 ? d->dphi2 : d->dphi4};
    const size_t cnt[5] = {(size_t)nq, (size_t)nq * nv, (size_t)nq * nv * 2, (size_t)nq * nn, (size_t)nq * nn * 2};
    std::vector<double> t;
    for (int k = 0; k < 5; ++k) t.insert(t.end(), src[k], src[k] + cnt[k]);
    rc = gm_post_up(&p->tab[r], t.data(), t.size());
  }
  if (rc == GM_OK) {
    std::vector<double> t(d->dphi_nodes, d->dphi_nodes + (size_t)nn * nn * 2);
    t.insert(t.end(), d->dg_nodes, d->dg_nodes + (size_t)nn * nv * 2);
    rc = gm_post_up(&p->ntab, t.data(), t.size());
  }
  if (rc == GM_OK) rc = gm_post_space_build(h, p->sp[0], d->cell_dofs_psi, d->n_free_psi, d->n_diri_psi, d->diri_psi);
  if (rc == GM_OK) rc = gm_post_space_build(h, p->sp[1], d->cell_dofs_pi, d->n_free_pi, d->n_diri_pi, d->diri_pi);
  if (rc == GM_OK) {
    p->n_nodes = d->n_nodes;
    rc = gm_post_up(&p->cell_nodes, d->cell_nodes, (size_t)(h->ncells * nn));
    std::vector<int32_t> last((size_t)d->n_nodes, 0);
    for (int64_t k = 0; k < h->ncells * nn; ++k) last[d->cell_nodes[k]] = (int32_t)k;  // ascending: the last cell wins
    if (rc == GM_OK) rc = gm_post_up(&p->node_last, last.data(), last.size());
  }
  if (rc == GM_OK) {
    const size_t per = (size_t)std::max(nv * nv + nv, 2 * nn);
    cudaError_t e = cudaMalloc(&p->cellbuf, sizeof(double) * per * (size_t)h->ncells);
    if (e != cudaSuccess) { gm_set_error("gm_post_create: %s", cudaGetErrorString(e)); rc = GM_ERR_OOM; }
  }
  if (rc != GM_OK) gm_post_destroy(h);
  return rc;
}

extern "C" int gm_post_sizes(gm_handle *h, int32_t which, int64_t *n, int64_t *nnz) {
  gm_post *p = h ? (gm_post *)h->post : nullptr;
  if (!p || which < 0 || which > 1) { gm_set_error("gm_post_sizes: call gm_post_create first / bad problem id"); return GM_ERR_STATE; }
  if (n) *n = p->sp[which].n;
  if (nnz) *nnz = p->sp[which].nnz;
  return GM_OK;
}

extern "C" int gm_post_get_pattern(gm_handle *h, int32_t which, int64_t *colptr, int64_t *rowval, int32_t index_base) {
  gm_post *p = h ? (gm_post *)h->post : nullptr;
  if (!p || which < 0 || which > 1 || !colptr || !rowval) { gm_set_error("gm_post_get_pattern: bad argument"); return GM_ERR_ARG; }
  const gm_post_space &s = p->sp[which];
  for (int64_t k = 0; k <= s.n; ++k) colptr[k] = s.hptr[k] + index_base;
  for (int64_t k = 0; k < s.nnz; ++k) rowval[k] = (int64_t)s.hidx[k] + index_base;
  return GM_OK;
}

extern "C" int gm_post_assemble(gm_handle *h, int32_t which, const double *x, double *nzval, double *b) {
  gm_post *p = h ? (gm_post *)h->post : nullptr;
  if (!p || which < 0 || which > 1 || !x || !nzval || !b) { gm_set_error("gm_post_assemble: bad argument"); return GM_ERR_ARG; }
  GM_CUDA(cudaSetDevice(h->dev));
  gm_post_space &s = p->sp[which];
  double *d_nz = nullptr, *d_b = nullptr;
  GM_CUDA(cudaMalloc(&d_nz, sizeof(double) * (size_t)std::max<int64_t>(s.nnz, 1)));
  GM_CUDA(cudaMalloc(&d_b, sizeof(double) * (size_t)std::max<int64_t>(s.n, 1)));
  GM_CUDA(cudaMemcpyAsync(h->d_x, x, sizeof(double) * (size_t)h->n_total, cudaMemcpyHostToDevice, h->stream));
  GmPostArgs a = gm_post_args(h, p, which, h->d_x);
  a.pdofs = s.dofs; a.pdiri = s.diri;
  if (which == 0) k_post_cells<0><<<gm_blocks(h->ncells, 128), 128, 0, h->stream>>>(a);
  else k_post_cells<1><<<gm_blocks(h->ncells, 128), 128, 0, h->stream>>>(a);
  k_post_gather<<<gm_blocks(s.n, 128), 128, 0, h->stream>>>(s.n, h->nv, s.adjptr, s.adj, s.dofs, s.ptr, s.idx, p->cellbuf, d_nz, d_b);
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(nzval, d_nz, sizeof(double) * (size_t)s.nnz, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(b, d_b, sizeof(double) * (size_t)s.n, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(d_nz); cudaFree(d_b);
  if (e != cudaSuccess) { gm_set_error("gm_post_assemble: %s", cudaGetErrorString(e)); return GM_ERR_CUDA; }
  return GM_OK;
}

extern "C" int gm_post_entropy(gm_handle *h, const double *x, double *sth, double *sfl, double *totals) {
  gm_post *p = h ? (gm_post *)h->post : nullptr;
  if (!p || !x || !sth || !sfl || !totals) { gm_set_error("gm_post_entropy: bad argument"); return GM_ERR_ARG; }
  GM_CUDA(cudaSetDevice(h->dev));
  double *d_s = nullptr, *d_part = nullptr, *d_tot = nullptr;
  void *tmp = nullptr;
  size_t tb = 0;
  int rc = GM_OK;
  cudaError_t e = cudaMalloc(&d_s, sizeof(double) * 2 * (size_t)p->n_nodes);
  if (e == cudaSuccess) e = cudaMalloc(&d_part, sizeof(double) * 2 * (size_t)h->ncells);
  if (e == cudaSuccess) e = cudaMalloc(&d_tot, 16);
  if (e == cudaSuccess) e = cudaMemcpyAsync(h->d_x, x, sizeof(double) * (size_t)h->n_total, cudaMemcpyHostToDevice, h->stream);
  if (e == cudaSuccess) {
    GmPostArgs a = gm_post_args(h, p, 0, h->d_x);
    k_post_entropy_cells<<<gm_blocks(h->ncells, 128), 128, 0, h->stream>>>(a, p->ntab);
    k_post_entropy_nodes<<<gm_blocks(p->n_nodes, 256), 256, 0, h->stream>>>(p->n_nodes, p->node_last, p->cellbuf, d_s, d_s + p->n_nodes);
    k_post_entropy_int<<<gm_blocks(h->ncells, 128), 128, 0, h->stream>>>(a, p->cell_nodes, d_s, d_s + p->n_nodes, d_part);
    cub::DeviceReduce::Sum(nullptr, tb, d_part, d_tot, h->ncells, h->stream);
    e = cudaMalloc(&tmp, tb ? tb : 16);
    if (e == cudaSuccess) {
      cub::DeviceReduce::Sum(tmp, tb, d_part, d_tot, h->ncells, h->stream);
      cub::DeviceReduce::Sum(tmp, tb, d_part + h->ncells, d_tot + 1, h->ncells, h->stream);
      e = cudaGetLastError();
    }
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(sth, d_s, sizeof(double) * (size_t)p->n_nodes, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(sfl, d_s + p->n_nodes, sizeof(double) * (size_t)p->n_nodes, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(totals, d_tot, 16, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(d_s); cudaFree(d_part); cudaFree(d_tot); cudaFree(tmp);
  if (e != cudaSuccess) { gm_set_error("gm_post_entropy: %s", cudaGetErrorString(e)); rc = GM_ERR_CUDA; }
  else totals[2] = totals[0] / (totals[0] + totals[1]);  // Be_avg (Solver.jl:292)
  return rc;
}
