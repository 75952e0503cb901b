// gm_pattern.cuh -- one-time symbolic pass on the device.
//
// Replaces allocate_jacobian(op,uh) of Gridap's SparseMatrixAssembler as entered from
// /root/reference/src/Solver.jl:157 (symbolic count -> allocate -> insert -> compress; restated
// in SURVEY.md A.5): for every free dof the sorted-unique union, over its incident cells, of the
// free dofs of the touched blocks (UU UP UT PU TU TT).  The pattern is structurally symmetric, so
// the same (ptr, idx) is Gridap's (colptr,rowval) and the CSR (rowptr,colind).
//
// Besides (ptr, idx) it emits the cell->nnz map in compressed form: rank[c][m][M] = position of
// minor dof m inside the list of major dof M (uint16), so nnz index = ptr[M] + rank.  837 x 8-byte
// explicit indices per cell (112 GB at 4096^2) are never materialised.
#pragma once
#include <cub/cub.cuh>

#include "gm_internal.h"

#define GM_MAXCAND 320  // max distinct entries of one row (valence-14 P2 vertex: ~170)

__device__ __forceinline__ int gm_field_of(int l, int nn) { return l < 2 * nn ? 0 : (l < 2 * nn + 3 ? 1 : 2); }
__device__ __forceinline__ bool gm_touched(int f1, int f2) { return !(f1 == 1 && f2 != 0) && !(f2 == 1 && f1 != 0); }

// combine the three field-local signed tables into one table of global signed ids
__global__ void k_combine_dofs(int64_t ncells, int nn, const int32_t *__restrict__ du, const int32_t *__restrict__ dp,
                               const int32_t *__restrict__ dt, int64_t oP, int64_t oT, int64_t dP, int64_t dT,
                               int32_t *__restrict__ g) {
  const int ld = 3 * nn + 3;
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= ncells * ld) return;
  int64_t c = t / ld;
  int l = (int)(t - c * ld);
  int32_t v;
  if (l < 2 * nn) {
    v = du[c * 2 * nn + l];
  } else if (l < 2 * nn + 3) {
    int32_t d = dp[c * 3 + (l - 2 * nn)];
    v = d > 0 ? (int32_t)(d + oP) : (int32_t)(d - dP);
  } else {
    int32_t d = dt[c * nn + (l - 2 * nn - 3)];
    v = d > 0 ? (int32_t)(d + oT) : (int32_t)(d - dT);
  }
  g[t] = v;
}

__global__ void k_count_adj(int64_t nent, const int32_t *__restrict__ g, int32_t *__restrict__ cnt) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= nent) return;
  int32_t d = g[t];
  if (d > 0) atomicAdd(&cnt[d - 1], 1);
}

__global__ void k_fill_adj(int64_t nent, const int32_t *__restrict__ g, const int64_t *__restrict__ adjptr,
                           int32_t *__restrict__ cursor, int32_t *__restrict__ adj) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= nent) return;
  int32_t d = g[t];
  if (d > 0) {
    int p = atomicAdd(&cursor[d - 1], 1);
    adj[adjptr[d - 1] + p] = (int32_t)t;  // t = cell*ld + local
  }
}

// ascending (cell, local) order inside every dof's adjacency segment: the gather path sums a list's contributions in
// Gridap's cell-loop order, independent of the atomics of k_fill_adj
__global__ void k_sort_adj(int64_t n, const int64_t *__restrict__ adjptr, int32_t *__restrict__ adj) {
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n) return;
  const int64_t a0 = adjptr[r], a1 = adjptr[r + 1];
  for (int64_t i = a0 + 1; i < a1; ++i) {
    const int32_t v = adj[i];
    int64_t k = i;
    while (k > a0 && adj[k - 1] > v) { adj[k] = adj[k - 1]; --k; }
    adj[k] = v;
  }
}

// Thread per major dof.  PASS 0: row length.  PASS 1: write idx and the rank table.
template <int PASS>
__global__ void k_rows(int64_t n, int nn, const int32_t *__restrict__ g, const int64_t *__restrict__ adjptr,
                       const int32_t *__restrict__ adj, int64_t *__restrict__ len_or_ptr, int32_t *__restrict__ idx,
                       uint16_t *__restrict__ rank, uint16_t *__restrict__ grank, uint16_t *__restrict__ grankp, int *__restrict__ err) {
  const int ld = 3 * nn + 3;
  int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= n) return;
  int32_t cols[GM_MAXCAND];
  int m = 0;
  const int64_t a0 = adjptr[r], a1 = adjptr[r + 1];
  for (int64_t p = a0; p < a1; ++p) {
    const int32_t e = adj[p];
    const int64_t c = e / ld;
    const int fM = gm_field_of(e - (int32_t)(c * ld), nn);
    for (int l = 0; l < ld; ++l) {
      if (!gm_touched(fM, gm_field_of(l, nn))) continue;
      const int32_t d = g[c * ld + l];
      if (d <= 0) continue;
      // sorted insert-if-absent
      int lo = 0, hi = m;
      while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (cols[mid] < d) lo = mid + 1; else hi = mid;
      }
      if (lo < m && cols[lo] == d) continue;
      if (m == GM_MAXCAND) { atomicExch(err, 1); return; }
      for (int k = m; k > lo; --k) cols[k] = cols[k - 1];
      cols[lo] = d;
      ++m;
    }
  }
  if (PASS == 0) {
    len_or_ptr[r] = m;
    return;
  }
  const int64_t base = len_or_ptr[r];
  if (idx)
    for (int k = 0; k < m; ++k) idx[base + k] = cols[k] - 1;
  for (int64_t p = a0; p < a1; ++p) {
    const int32_t e = adj[p];
    const int64_t c = e / ld;
    const int M = e - (int32_t)(c * ld);
    const int fM = gm_field_of(M, nn);
    for (int l = 0; l < ld; ++l) {
      uint16_t rk = 0xFFFF;
      const int32_t d = g[c * ld + l];
      if (d > 0 && gm_touched(fM, gm_field_of(l, nn))) {
        int lo = 0, hi = m;
        while (lo < hi) {
          int mid = (lo + hi) >> 1;
          if (cols[mid] < d) lo = mid + 1; else hi = mid;
        }
        rk = (uint16_t)lo;
      }
      if (rank) rank[(c * ld + l) * ld + M] = rk;
      if (grank && gm_touched(fM, gm_field_of(l, nn))) {
        // compact major-major layout of the gather path (gm_gather_kernel.cuh): [U majors][T majors]; P majors apart
        const int su = ld, st = 3 * nn, offT = 2 * nn * su, nec = offT + nn * st;
        const int s = (fM == 2 && l >= 2 * nn) ? l - 3 : l;
        if (fM == 1) grankp[c * 6 * nn + (M - 2 * nn) * 2 * nn + s] = rk;
        else grank[c * nec + (fM == 0 ? M * su : offT + (M - 2 * nn - 3) * st) + s] = rk;
      }
    }
  }
}

__global__ void k_fill_u16(uint16_t *p, int64_t n, uint16_t v) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t < n) p[t] = v;
}

static inline unsigned gm_blocks(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

static int gm_build_pattern(gm_handle *h) {
  const int64_t n = h->n_total, nent = h->ncells * h->ld;
  const int BS = 256;
  int32_t *cnt = nullptr, *adj = nullptr;
  int64_t *adjptr = nullptr, *len = nullptr;
  uint16_t *grank = nullptr, *grankp = nullptr;
  int *err = nullptr;
  void *tmp = nullptr;
  size_t tmpb = 0;
  int rc = GM_OK;
  int64_t keep = h->device_bytes;
#define BAIL(x) do { rc = (x); if (rc != GM_OK) goto done; } while (0)
#define BAILC(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { gm_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); rc = e__ == cudaErrorMemoryAllocation ? GM_ERR_OOM : GM_ERR_CUDA; goto done; } } while (0)
  BAIL(gm_dalloc(h, &cnt, n + 1));
  BAIL(gm_dalloc(h, &adjptr, n + 1));
  BAIL(gm_dalloc(h, &adj, nent));
  BAIL(gm_dalloc(h, &len, n + 1));
  BAIL(gm_dalloc(h, &err, 1));
  BAILC(cudaMemsetAsync(cnt, 0, sizeof(int32_t) * (n + 1), h->stream));
  BAILC(cudaMemsetAsync(err, 0, sizeof(int), h->stream));
  k_count_adj<<<gm_blocks(nent, BS), BS, 0, h->stream>>>(nent, h->gdofs, cnt);
  cub::DeviceScan::ExclusiveSum(nullptr, tmpb, cnt, adjptr, n + 1, h->stream);
  {
    size_t t2 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, t2, len, len, n + 1, h->stream);
    if (t2 > tmpb) tmpb = t2;
  }
  BAILC(cudaMalloc(&tmp, tmpb ? tmpb : 16));
  BAILC(cub::DeviceScan::ExclusiveSum(tmp, tmpb, cnt, adjptr, n + 1, h->stream));
  BAILC(cudaMemsetAsync(cnt, 0, sizeof(int32_t) * (n + 1), h->stream));
  k_fill_adj<<<gm_blocks(nent, BS), BS, 0, h->stream>>>(nent, h->gdofs, adjptr, cnt, adj);
  k_sort_adj<<<gm_blocks(n, BS), BS, 0, h->stream>>>(n, adjptr, adj);
  BAILC(cudaMemsetAsync(len, 0, sizeof(int64_t) * (n + 1), h->stream));
  k_rows<0><<<gm_blocks(n, 128), 128, 0, h->stream>>>(n, h->nn, h->gdofs, adjptr, adj, len, nullptr, nullptr, nullptr, nullptr, err);
  {
    // max row length (for diagnostics / uint16 safety) and prefix sum
    BAIL(gm_dalloc(h, &h->ptr, n + 1));
    keep += (int64_t)sizeof(int64_t) * (n + 1);
    BAILC(cub::DeviceScan::ExclusiveSum(tmp, tmpb, len, h->ptr, n + 1, h->stream));
    int herr = 0;
    BAILC(cudaMemcpyAsync(&herr, err, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    BAILC(cudaMemcpyAsync(&h->nnz, h->ptr + n, sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    BAILC(cudaStreamSynchronize(h->stream));
    if (herr) {
      gm_set_error("gm_pattern: a row has more than %d distinct entries", GM_MAXCAND);
      rc = GM_ERR_ARG;
      goto done;
    }
  }
  BAIL(gm_dalloc(h, &h->idx, h->nnz));
  keep += (int64_t)sizeof(int32_t) * (h->nnz > 0 ? h->nnz : 1);
  if (h->active_path == GM_PATH_GATHER) {  // compact map of the gather path; the adjacency is kept (gm_gather.cuh)
    const int64_t nec = (int64_t)2 * h->nn * h->ld + 3 * h->nn * h->nn, nep = 6 * h->nn;
    BAIL(gm_dalloc(h, &grank, h->ncells * nec));
    BAIL(gm_dalloc(h, &grankp, h->ncells * nep));
    keep += (int64_t)sizeof(uint16_t) * h->ncells * (nec + nep) + (int64_t)sizeof(int32_t) * nent + (int64_t)sizeof(int64_t) * (n + 1);
    k_fill_u16<<<gm_blocks(h->ncells * nec, BS), BS, 0, h->stream>>>(grank, h->ncells * nec, 0xFFFF);
    k_fill_u16<<<gm_blocks(h->ncells * nep, BS), BS, 0, h->stream>>>(grankp, h->ncells * nep, 0xFFFF);
  } else {
    BAIL(gm_dalloc(h, &h->rank, h->ncells * h->ld * h->ld));
    keep += (int64_t)sizeof(uint16_t) * h->ncells * h->ld * h->ld;
    k_fill_u16<<<gm_blocks(h->ncells * h->ld * h->ld, BS), BS, 0, h->stream>>>(h->rank, h->ncells * h->ld * h->ld, 0xFFFF);
  }
  k_rows<1><<<gm_blocks(n, 128), 128, 0, h->stream>>>(n, h->nn, h->gdofs, adjptr, adj, h->ptr, h->idx, h->rank, grank, grankp, err);
  BAILC(cudaGetLastError());
  BAILC(cudaStreamSynchronize(h->stream));
  h->has_pattern = true;
  if (h->active_path == GM_PATH_GATHER) {
    h->g_adjptr = adjptr; h->g_adj = adj; h->g_rank = grank; h->g_rankp = grankp;
    adjptr = nullptr; adj = nullptr; grank = nullptr; grankp = nullptr;
  }
done:
  cudaFree(tmp);
  cudaFree(cnt);
  cudaFree(adjptr);
  cudaFree(adj);
  cudaFree(grank);
  cudaFree(grankp);
  cudaFree(len);
  cudaFree(err);
  h->device_bytes = keep;
#undef BAIL
#undef BAILC
  return rc;
}
