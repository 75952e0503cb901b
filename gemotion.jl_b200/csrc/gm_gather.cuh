// gm_gather.cuh -- host side of the generic owner-computes path (kernels: gm_gather_kernel.cuh).
// GM_PATH_GATHER: the default for every mesh that is not a verified Cartesian strip mesh (Gmsh triangles, kuehn.jl:13-22).
#pragma once
#include "gm_internal.h"
#include "gm_gather_kernel.cuh"

#define GG_WINDOW 4096  // cells per locality window of the group order

struct gm_gather {
  GmConst *d_const = nullptr;  // per handle: uploads are ordered on the handle's stream, handles never share it
  double *cm = nullptr;        // [ncells][NEC] cell matrices (touched blocks, major-major)
  double *cr = nullptr;        // [ncells][ld]  cell residuals
  GgDesc *desc = nullptr;      // [ngroup][GG_MPW] records of the velocity / temperature majors, groups sorted by first incident cell
  int64_t ngroup = 0;
  int lstride = 96;
  // launch set-up done once per handle (a handle lives on one device; the shared-memory opt-in is a per-device function attribute)
  bool prepared = false;
  int cm_per_sm = 1;
  // parameters of the constant block on the device (re-uploaded only when gm_set_params changed them)
  bool const_valid = false;
  gm_params const_prm;
};

// one 64-byte record per velocity / temperature major (list base, length, incident cells) + the longest list;
// key[group] = first incident cell of the group (its processing order in k_lists)
// Partitioned meshes: a major is assembled by this rank iff its lowest-numbered incident cell (GLOBAL id) is an own cell; the
// others get cnt = 0 (their lists are zero-filled, never summed) and are counted out of n_owned.
__global__ void k_gg_desc(int64_t nU, int64_t oT, int64_t n, int ld, const int64_t *__restrict__ adjptr, const int32_t *__restrict__ adj,
                          const int64_t *__restrict__ ptr, int64_t nown, int64_t cell_begin, const int64_t *__restrict__ ghost_ids,
                          GgDesc *__restrict__ desc, int *__restrict__ maxlen, unsigned long long *__restrict__ nowned) {
  const int64_t wg = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  const int64_t nUp = (nU + GG_MPW - 1) / GG_MPW * GG_MPW;
  if (wg >= (nUp + (n - oT) + GG_MPW - 1) / GG_MPW * GG_MPW) return;
  GgDesc d;
  const bool real = wg < nU || (wg >= nUp && wg - nUp + oT < n);
  const int64_t j = wg < nU ? wg : wg - nUp + oT;
  d.isU = wg < nUp ? 1 : 0;
  if (!real) {  // padding: the lists of a group are adjacent in nzval
    d.p0 = ptr[wg < nUp ? nU : n]; d.j = -1; d.len = 0; d.cnt = 0;
    for (int t = 0; t < 10; ++t) d.e[t] = 0;
  } else {
    const int64_t a0 = adjptr[j];
    d.p0 = ptr[j];
    d.j = (int32_t)j;
    d.len = (int32_t)(ptr[j + 1] - d.p0);
    d.cnt = (int32_t)(adjptr[j + 1] - a0);
    if (ghost_ids) {
      int64_t lowest = INT64_MAX;
      for (int t = 0; t < d.cnt; ++t) {
        const int64_t lc = adj[a0 + t] / ld, gc = lc < nown ? cell_begin + lc : ghost_ids[lc - nown];
        lowest = gc < lowest ? gc : lowest;
      }
      if (!(lowest >= cell_begin && lowest < cell_begin + nown)) { d.cnt = 0; d.j = -1; }
    }
    if (d.j >= 0) atomicAdd(nowned, 1ull);
    for (int t = 0; t < 10; ++t) {
      uint32_t sg = 0;
      if (t < d.cnt) {
        const int32_t e = adj[a0 + t], cc = e / ld;
        const int M = e - cc * ld, nn = (ld - 3) / 3;
        sg = (uint32_t)cc * (uint32_t)(2 * nn * ld + 3 * nn * nn) + (M < 2 * nn ? M * ld : 2 * nn * ld + (M - 2 * nn - 3) * 3 * nn);
      }
      d.e[t] = sg;
    }
    atomicMax(maxlen, d.len);
  }
  desc[wg] = d;
}
// Processing order of the groups: windows of GG_WINDOW cells keep the reads of `cm` local; inside a window the long groups
// (vertex dofs) come first, so that the warps of a CTA finish together (a CTA's registers and shared memory are held until its
// slowest warp ends).  Groups without any work -- padding, dofs this rank does not see or does not own: most of the global dof
// range on a rank of a partitioned mesh -- get the largest key and are cut off after the sort (`nlive`).
__global__ void k_gg_key(int64_t ngroup, int nec, const GgDesc *__restrict__ desc, int32_t *__restrict__ key, int32_t *__restrict__ gid,
                         unsigned long long *__restrict__ nlive) {
  const int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (g >= ngroup) return;
  uint32_t first = 0xFFFFFFFFu;
  int maxcnt = 0;
  for (int r = 0; r < GG_MPW; ++r) {
    const GgDesc &d = desc[g * GG_MPW + r];
    if (d.cnt > 0) {
      const uint32_t cell = d.e[0] / (uint32_t)nec;
      first = cell < first ? cell : first;
      maxcnt = d.cnt > maxcnt ? d.cnt : maxcnt;
    }
  }
  key[g] = maxcnt > 0 ? (int32_t)((first / GG_WINDOW) * 4 + (desc[g * GG_MPW].isU ? 0 : 2) + (maxcnt >= 4 ? 0 : 1)) : 0x7fffffff;
  gid[g] = (int32_t)g;
  if (maxcnt > 0) atomicAdd(nlive, 1ull);
}
// The gather path writes the lists of the pressure majors straight from the one cell that owns them (P1-disc, `space=:P`,
// Solver.jl:89-90): a pressure dof shared by several cells (a continuous pressure space) is not supported here.
__global__ void k_gg_check_p(int64_t nP, int64_t oP, const int64_t *__restrict__ adjptr, int *__restrict__ bad) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t < nP && adjptr[oP + t + 1] - adjptr[oP + t] > 1) atomicExch(bad, 1);
}
__global__ void k_gg_permute(int64_t ngroup, const int32_t *__restrict__ gid, const GgDesc *__restrict__ src, GgDesc *__restrict__ dst) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;  // one thread per 16-byte quarter record
  if (t >= ngroup * GG_MPW * 4) return;
  const int64_t g = t / (GG_MPW * 4), r = t - g * (GG_MPW * 4);
  reinterpret_cast<int4 *>(dst)[t] = reinterpret_cast<const int4 *>(src)[(int64_t)gid[g] * (GG_MPW * 4) + r];
}

static void gm_gather_destroy(gm_handle *h) {
  gm_gather *g = (gm_gather *)h->gather;
  if (g) {
    cudaFree(g->d_const); cudaFree(g->cm); cudaFree(g->cr); cudaFree(g->desc);
    delete g;
  }
  h->gather = nullptr;
  cudaFree(h->g_adjptr); cudaFree(h->g_adj); cudaFree(h->g_rank); cudaFree(h->g_rankp);
  h->g_adjptr = nullptr; h->g_adj = nullptr; h->g_rank = nullptr; h->g_rankp = nullptr;
}

// workspaces, after the symbolic pass
static int gm_gather_after_pattern(gm_handle *h) {
  gm_gather *g = new (std::nothrow) gm_gather();
  if (!g) return GM_ERR_OOM;
  h->gather = g;
  const int64_t nec = (int64_t)2 * h->nn * h->ld + 3 * h->nn * h->nn;
  if (h->ncells * nec >= ((int64_t)1 << 32)) {
    gm_set_error("gm_pattern: the gather path indexes its cell-matrix workspace with 32 bits (%lld cells x %lld values do not fit)", (long long)h->ncells, (long long)nec);
    return GM_ERR_ARG;
  }
  int rc = gm_dalloc(h, &g->d_const, 1);
  if (rc == GM_OK) rc = gm_dalloc(h, &g->cm, h->ncells * nec);
  if (rc == GM_OK) rc = gm_dalloc(h, &g->cr, h->ncells * h->ld);
  if (rc != GM_OK) {
    gm_set_error("gm_pattern: the gather path needs %lld bytes of cell-matrix workspace (use GM_PATH_SCATTER for meshes that do not fit)",
                 (long long)(8 * h->ncells * nec));
    return rc;
  }
  static_assert(sizeof(GgDesc) == 64, "one 64-byte record per major");
  {
    int *d_bad = nullptr, hbad = 0;
    const int64_t nP = h->off[2] - h->off[1];
    GM_CUDA(cudaMalloc(&d_bad, sizeof(int)));
    cudaError_t eb = cudaMemsetAsync(d_bad, 0, sizeof(int), h->stream);
    if (eb == cudaSuccess && nP > 0) k_gg_check_p<<<gm_blocks(nP, 256), 256, 0, h->stream>>>(nP, h->off[1], h->g_adjptr, d_bad);
    if (eb == cudaSuccess) eb = cudaMemcpyAsync(&hbad, d_bad, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
    if (eb == cudaSuccess) eb = cudaStreamSynchronize(h->stream);
    cudaFree(d_bad);
    if (eb != cudaSuccess) { gm_set_error("gm_pattern: %s", cudaGetErrorString(eb)); return GM_ERR_CUDA; }
    if (hbad) {
      gm_set_error("gm_pattern: the gather path needs a discontinuous pressure space (every pressure dof in exactly one cell, Solver.jl:89-90); use GM_PATH_SCATTER");
      return GM_ERR_ARG;
    }
  }
  const int64_t nrec = ((h->off[1] + GG_MPW - 1) / GG_MPW * GG_MPW + (h->n_total - h->off[2]) + GG_MPW - 1) / GG_MPW * GG_MPW;
  g->ngroup = nrec / GG_MPW;
  if (g->ngroup == 0) return GM_OK;
  GgDesc *tmpd = nullptr;
  int32_t *key = nullptr, *gid = nullptr, *key2 = nullptr, *gid2 = nullptr;
  int *d_max = nullptr, hmax = 0;
  unsigned long long *d_own = nullptr, hown[2] = {0, 0};  // owned velocity / temperature dofs, groups with work
  void *tmp = nullptr;
  size_t tb = 0;
  rc = gm_dalloc(h, &g->desc, nrec);
  if (rc) return rc;
  cudaError_t e = cudaMalloc(&tmpd, sizeof(GgDesc) * nrec);
  if (e == cudaSuccess) e = cudaMalloc(&key, 4 * g->ngroup);
  if (e == cudaSuccess) e = cudaMalloc(&gid, 4 * g->ngroup);
  if (e == cudaSuccess) e = cudaMalloc(&key2, 4 * g->ngroup);
  if (e == cudaSuccess) e = cudaMalloc(&gid2, 4 * g->ngroup);
  if (e == cudaSuccess) e = cudaMalloc(&d_max, sizeof(int));
  if (e == cudaSuccess) e = cudaMemsetAsync(d_max, 0, sizeof(int), h->stream);
  if (e == cudaSuccess) e = cudaMalloc(&d_own, 16);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_own, 0, 16, h->stream);
  if (e == cudaSuccess) {
    k_gg_desc<<<gm_blocks(nrec, 256), 256, 0, h->stream>>>(h->off[1], h->off[2], h->n_total, h->ld, h->g_adjptr, h->g_adj, h->ptr, h->n_own_cells,
                                                          h->d.cell_begin, h->ghost_ids, tmpd, d_max, d_own);
    k_gg_key<<<gm_blocks(g->ngroup, 256), 256, 0, h->stream>>>(g->ngroup, (int)nec, tmpd, key, gid, d_own + 1);
    cub::DeviceRadixSort::SortPairs(nullptr, tb, key, key2, gid, gid2, (int)g->ngroup, 0, 32, h->stream);
    e = cudaMalloc(&tmp, tb ? tb : 16);
  }
  if (e == cudaSuccess) e = cub::DeviceRadixSort::SortPairs(tmp, tb, key, key2, gid, gid2, (int)g->ngroup, 0, 32, h->stream);  // (stable)
  if (e == cudaSuccess) {
    k_gg_permute<<<gm_blocks(nrec * 4, 256), 256, 0, h->stream>>>(g->ngroup, gid2, tmpd, g->desc);
    e = cudaMemcpyAsync(&hmax, d_max, sizeof(int), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(hown, d_own, 16, cudaMemcpyDeviceToHost, h->stream);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(tmpd); cudaFree(key); cudaFree(gid); cudaFree(key2); cudaFree(gid2); cudaFree(d_max); cudaFree(d_own); cudaFree(tmp);
  if (e != cudaSuccess) { gm_set_error("gm_pattern: %s", cudaGetErrorString(e)); return e == cudaErrorMemoryAllocation ? GM_ERR_OOM : GM_ERR_CUDA; }
  h->max_row_len = hmax;
  if (h->ghost_ids) {  // owned dofs of a partitioned mesh: velocity / temperature majors counted above + the pressure dofs of the own cells
    int64_t np_own = 3 * h->n_own_cells;
    if (h->d.cell_begin + h->n_own_cells == h->d.ncells_global) np_own -= 1;  // (the fixed last pressure dof, Solver.jl:90)
    h->n_owned = (int64_t)hown[0] + np_own;
  }
  g->ngroup = (int64_t)hown[1];  // the groups with work come first in the sorted record array
  g->lstride = (hmax + 3) & ~3;
  return GM_OK;
}

template <int NN, int NQ, int NV, int CPB, bool CSR>
static int gm_gather_cellmat(gm_handle *h, GgArgs &a) {
  const size_t sm = sizeof(GgSmem<NN, NQ, NV, CPB>);
  gm_gather *g = (gm_gather *)h->gather;
  if (!g->prepared) {  // per device, not per process: every handle does it for its own device
    GM_CUDA(cudaFuncSetAttribute(k_cellmat<NN, NQ, NV, CPB, CSR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    int per_sm = 0;
    GM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_cellmat<NN, NQ, NV, CPB, CSR>, NN * CPB, sm));
    g->cm_per_sm = per_sm < 1 ? 1 : per_sm;
    GM_CUDA(cudaFuncSetAttribute(k_lists<NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * GG_LWARPS * GG_MPW * a.lstride)));
    g->prepared = true;
  }
  const int64_t nbatch = (h->ncells + CPB - 1) / CPB;
  a.nbatch = (int)nbatch;
  const int64_t grid = std::min<int64_t>(nbatch, (int64_t)h->num_sms * g->cm_per_sm);  // persistent: every CTA loops over batches
  k_cellmat<NN, NQ, NV, CPB, CSR><<<(unsigned)grid, NN * CPB, sm, h->stream>>>(a);
  return GM_OK;
}

template <int NN, int NQ, int NV, int CPB>
static int gm_gather_launch(gm_handle *h, GgArgs &a, int64_t *launches) {
  const int rc = h->d.layout == GM_LAYOUT_CSR ? gm_gather_cellmat<NN, NQ, NV, CPB, true>(h, a) : gm_gather_cellmat<NN, NQ, NV, CPB, false>(h, a);
  if (rc) return rc;
  ++*launches;
  const int64_t nw = a.ngroup;  // warps
  if (nw > 0) {
    const size_t sml = sizeof(double) * GG_LWARPS * GG_MPW * a.lstride;
    k_lists<NN><<<(unsigned)((nw + GG_LWARPS - 1) / GG_LWARPS), GG_LWARPS * 32, sml, h->stream>>>(a);
    ++*launches;
  }
  GM_CUDA(cudaGetLastError());
  return GM_OK;
}

static int gm_gather_run(gm_handle *h, const double *d_x, double *d_nz, double *d_b, unsigned flags, int64_t *launches) {
  gm_gather *g = (gm_gather *)h->gather;
  if (!g) { gm_set_error("gm_assemble: gather path without workspaces"); return GM_ERR_STATE; }
  if (!g->const_valid || memcmp(&g->const_prm, &h->prm, sizeof(gm_params)) != 0) {
    GmConst &c = h->hc;
    c.Pr = h->prm.Pr; c.Ra = h->prm.Ra; c.n = h->prm.n; c.reg = h->prm.reg;
    c.gx = h->prm.gx; c.gy = h->prm.gy; c.jac_scaling = h->prm.jac_scaling;
    GM_CUDA(cudaMemcpyAsync(g->d_const, &c, sizeof(GmConst), cudaMemcpyHostToDevice, h->stream));  // (pageable source: staged before the call returns; ordered on the handle's stream)
    g->const_prm = h->prm;
    g->const_valid = true;
  }
  if ((flags & GM_CHECK_FINITE) && h->ghost_ids) {
    // partitioned mesh: the entries of dofs this rank does not own are never written; give the scan defined values
    if ((flags & GM_JACOBIAN) && d_nz) GM_CUDA(cudaMemsetAsync(d_nz, 0, sizeof(double) * (size_t)h->nnz, h->stream));
    if ((flags & GM_RESIDUAL) && d_b) GM_CUDA(cudaMemsetAsync(d_b, 0, sizeof(double) * (size_t)h->n_total, h->stream));
  }
  GgArgs a;
  a.ncells = h->ncells;
  a.nown = h->n_own_cells;
  a.nU = h->off[1]; a.oT = h->off[2]; a.n = h->n_total;
  a.gc = g->d_const;
  a.gdofs = h->gdofs; a.dvals = h->dvals; a.coords = h->coords; a.cellv = h->cellv;
  a.ptr = h->ptr; a.grank = h->g_rank; a.grankp = h->g_rankp; a.adjptr = h->g_adjptr; a.adj = h->g_adj;
  a.x = d_x; a.cm = g->cm; a.cr = g->cr; a.nz = d_nz; a.b = d_b;
  a.flags = flags & (GM_JACOBIAN | GM_RESIDUAL);
  a.desc = g->desc; a.lstride = g->lstride; a.ngroup = g->ngroup;
  if (h->d.cell_type == GM_QUAD) return gm_gather_launch<9, 4, 4, 16>(h, a, launches);
  return gm_gather_launch<6, 3, 3, 16>(h, a, launches);
}
