// gm_multi.cuh -- multi-GPU row blocks: interface-row exchange.
//
// One process per GPU.  Rank r assembles the cells of its own cell rows only; the value lists and
// residual entries of the dofs on the mesh line it shares with rank r-1 (its LOWER interface) are
// partial sums that belong to rank r-1 (owner = lowest row block touching the dof).  They are packed
// by a gather kernel, sent with ncclSend/ncclRecv over NVLink, and added into the owner's lists by
// a fused unpack-add kernel.  Both ranks enumerate the interface dofs in ascending global id and
// hold the complete sparsity pattern of those rows (each rank's tables carry one ghost cell row),
// so the packed layouts agree without any index exchange.  No broadcast back: the owner keeps the row.
#pragma once
#include <nccl.h>
#include <cub/cub.cuh>

#include <algorithm>
#include <vector>

#include "gm_internal.h"

struct gm_iface {
  int64_t ndof = 0, nlist = 0;   // dofs on the line, total list entries
  int32_t *dof = nullptr;        // [ndof] global dof ids (1-based), ascending
  int64_t *base = nullptr;       // [ndof] list base in nz
  int64_t *off = nullptr;        // [ndof+1] packed offsets
  double *buf = nullptr;         // [nlist + ndof] packed lists, then residual entries
};

struct gm_multi {
  gm_iface lo, hi;
  ncclComm_t comm = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;  // exchange (pack + NCCL + unpack-add) start / end, for gm_times.exchange_ms
  cudaStream_t xstream = nullptr;            // high-priority stream: interface strip + pack + send / receive
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  bool timed = false;                        // ev0 / ev1 were recorded by the last assembly
};

__global__ void k_iface_ptr(int64_t n, const int32_t *__restrict__ dof, const int64_t *__restrict__ ptr,
                            int64_t *__restrict__ base, int64_t *__restrict__ len) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t g = dof[i];
  base[i] = ptr[g - 1];
  len[i] = ptr[g] - ptr[g - 1];
}

// one warp per interface dof
template <bool UNPACK>
__global__ void k_iface_move(int64_t ndof, int64_t nlist, const int32_t *__restrict__ dof, const int64_t *__restrict__ base,
                             const int64_t *__restrict__ off, double *nz, double *b, double *buf, unsigned flags) {
  const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= ndof) return;
  if (flags & GM_JACOBIAN) {
    const int64_t o = off[i], n = off[i + 1] - o, bs = base[i];
    for (int64_t k = lane; k < n; k += 32) {
      if (UNPACK) nz[bs + k] += buf[o + k];
      else buf[o + k] = nz[bs + k];
    }
  }
  if ((flags & GM_RESIDUAL) && lane == 0) {
    if (UNPACK) b[dof[i] - 1] += buf[nlist + i];
    else buf[nlist + i] = b[dof[i] - 1];
  }
}

static void gm_iface_free(gm_iface &f) {
  cudaFree(f.dof); cudaFree(f.base); cudaFree(f.off); cudaFree(f.buf);
  f = gm_iface();
}

static void gm_multi_destroy(gm_handle *h) {
  gm_multi *m = (gm_multi *)h->multi;
  if (!m) return;
  gm_iface_free(m->lo);
  gm_iface_free(m->hi);
  if (m->comm) ncclCommDestroy(m->comm);
  if (m->ev0) cudaEventDestroy(m->ev0);
  if (m->ev1) cudaEventDestroy(m->ev1);
  if (m->ev_fork) cudaEventDestroy(m->ev_fork);
  if (m->ev_join) cudaEventDestroy(m->ev_join);
  if (m->xstream) cudaStreamDestroy(m->xstream);
  delete m;
  h->multi = nullptr;
}

// dofs of the nodes on one horizontal line of a cell row: bottom nodes {0,1,4} or top nodes {2,3,5}
static int gm_iface_build(gm_handle *h, gm_iface &f, int row, bool top) {
  const int nx = h->d.nx;
  std::vector<int32_t> g((size_t)nx * 30);
  GM_CUDA(cudaMemcpy(g.data(), h->gdofs + (int64_t)row * nx * 30, sizeof(int32_t) * g.size(), cudaMemcpyDeviceToHost));
  static const int bot[3] = {0, 1, 4}, tp[3] = {2, 3, 5};
  const int *nd = top ? tp : bot;
  std::vector<int32_t> ids;
  for (int c = 0; c < nx; ++c)
    for (int k = 0; k < 3; ++k)
      for (int fo : {0, 9, 21}) {
        const int32_t d = g[(size_t)c * 30 + fo + nd[k]];
        if (d > 0) ids.push_back(d);
      }
  std::sort(ids.begin(), ids.end());
  ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
  f.ndof = (int64_t)ids.size();
  if (f.ndof == 0) return GM_OK;
  int64_t *len = nullptr;
  GM_CUDA(cudaMalloc(&f.dof, sizeof(int32_t) * f.ndof));
  GM_CUDA(cudaMalloc(&f.base, sizeof(int64_t) * f.ndof));
  GM_CUDA(cudaMalloc(&f.off, sizeof(int64_t) * (f.ndof + 1)));
  GM_CUDA(cudaMalloc(&len, sizeof(int64_t) * f.ndof));
  GM_CUDA(cudaMemcpy(f.dof, ids.data(), sizeof(int32_t) * f.ndof, cudaMemcpyHostToDevice));
  k_iface_ptr<<<gm_blocks(f.ndof, 256), 256, 0, h->stream>>>(f.ndof, f.dof, h->ptr, f.base, len);
  std::vector<int64_t> hl((size_t)f.ndof), ho((size_t)f.ndof + 1, 0);
  GM_CUDA(cudaMemcpyAsync(hl.data(), len, sizeof(int64_t) * f.ndof, cudaMemcpyDeviceToHost, h->stream));
  GM_CUDA(cudaStreamSynchronize(h->stream));
  for (int64_t i = 0; i < f.ndof; ++i) ho[i + 1] = ho[i] + hl[i];
  f.nlist = ho[f.ndof];
  GM_CUDA(cudaMemcpy(f.off, ho.data(), sizeof(int64_t) * (f.ndof + 1), cudaMemcpyHostToDevice));
  GM_CUDA(cudaMalloc(&f.buf, sizeof(double) * (f.nlist + f.ndof)));
  GM_CUDA(cudaMemset(f.buf, 0, sizeof(double) * (f.nlist + f.ndof)));
  cudaFree(len);
  h->device_bytes += 8 * (f.nlist + f.ndof) + 20 * f.ndof;
  return GM_OK;
}

// owned dofs of a row-block rank = dofs of its owned cell rows minus the lower interface line (owner = lowest block)
__global__ void k_mark_owned(int64_t first, int64_t n, const int32_t *__restrict__ g, uint8_t *__restrict__ flag) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t < n) {
    const int32_t d = g[first + t];
    if (d > 0) flag[d - 1] = 1;
  }
}
__global__ void k_unmark(int64_t n, const int32_t *__restrict__ dof, uint8_t *__restrict__ flag) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t < n) flag[dof[t] - 1] = 0;
}

// the handle's multi-GPU state (created on first use by gm_pattern or gm_comm_init)
static int gm_multi_get(gm_handle *h, gm_multi **out) {
  gm_multi *m = (gm_multi *)h->multi;
  if (!m) {
    m = new (std::nothrow) gm_multi();
    if (!m) return GM_ERR_OOM;
    h->multi = m;
    GM_CUDA(cudaEventCreate(&m->ev0));
    GM_CUDA(cudaEventCreate(&m->ev1));
    GM_CUDA(cudaEventCreateWithFlags(&m->ev_fork, cudaEventDisableTiming));
    GM_CUDA(cudaEventCreateWithFlags(&m->ev_join, cudaEventDisableTiming));
    int lo = 0, hi = 0;
    GM_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));  // (hi = numerically lowest = highest priority)
    GM_CUDA(cudaStreamCreateWithPriority(&m->xstream, cudaStreamNonBlocking, hi));
  }
  *out = m;
  return GM_OK;
}

static int gm_multi_after_pattern(gm_handle *h) {
  if (h->d.ghost_lo == 0 && h->d.ghost_hi == 0) return GM_OK;
  gm_multi *m = nullptr;
  {
    const int rc0 = gm_multi_get(h, &m);
    if (rc0) return rc0;
  }
  int rc = GM_OK;
  if (h->d.ghost_lo) rc = gm_iface_build(h, m->lo, h->d.ghost_lo, false);               // bottom line of the first owned row
  if (rc == GM_OK && h->d.ghost_hi) rc = gm_iface_build(h, m->hi, h->d.ny - h->d.ghost_hi - 1, true);  // top line of the last owned row
  if (rc == GM_OK) {  // gm_info: number of owned dofs
    uint8_t *flag = nullptr;
    int64_t *cnt = nullptr;
    void *tmp = nullptr;
    size_t tb = 0;
    const int64_t first = (int64_t)h->d.ghost_lo * h->d.nx * 30, n = (int64_t)(h->d.ny - h->d.ghost_lo - h->d.ghost_hi) * h->d.nx * 30;
    if (cudaMalloc(&flag, (size_t)h->n_total) == cudaSuccess && cudaMalloc(&cnt, 8) == cudaSuccess) {
      cudaMemsetAsync(flag, 0, (size_t)h->n_total, h->stream);
      k_mark_owned<<<gm_blocks(n, 256), 256, 0, h->stream>>>(first, n, h->gdofs, flag);
      if (m->lo.ndof) k_unmark<<<gm_blocks(m->lo.ndof, 256), 256, 0, h->stream>>>(m->lo.ndof, m->lo.dof, flag);
      cub::DeviceReduce::Sum(nullptr, tb, flag, cnt, h->n_total, h->stream);
      if (cudaMalloc(&tmp, tb ? tb : 16) == cudaSuccess) {
        cub::DeviceReduce::Sum(tmp, tb, flag, cnt, h->n_total, h->stream);
        cudaMemcpyAsync(&h->n_owned, cnt, 8, cudaMemcpyDeviceToHost, h->stream);
        cudaStreamSynchronize(h->stream);
      }
    }
    cudaFree(flag); cudaFree(cnt); cudaFree(tmp);
  }
  return rc;
}

static int gm_iface_run(gm_handle *h, gm_iface &f, bool unpack, double *nz, double *b, double *buf, unsigned flags, cudaStream_t stream = nullptr) {
  if (f.ndof == 0) return GM_OK;
  if (!stream) stream = h->stream;
  const unsigned nb = gm_blocks(f.ndof * 32, 256);
  if (unpack) k_iface_move<true><<<nb, 256, 0, stream>>>(f.ndof, f.nlist, f.dof, f.base, f.off, nz, b, buf, flags);
  else k_iface_move<false><<<nb, 256, 0, stream>>>(f.ndof, f.nlist, f.dof, f.base, f.off, nz, b, buf, flags);
  GM_CUDA(cudaGetLastError());
  return GM_OK;
}

#define GM_NCCL(call)                                                                      \
  do {                                                                                     \
    ncclResult_t r__ = (call);                                                             \
    if (r__ != ncclSuccess) {                                                              \
      gm_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, ncclGetErrorString(r__)); \
      return GM_ERR_NCCL;                                                                  \
    }                                                                                      \
  } while (0)

// pack lower interface -> send to rank-1 ; receive from rank+1 -> add into the upper interface rows
static int gm_multi_exchange(gm_handle *h, double *d_nz, double *d_b, unsigned flags, int64_t *launches) {
  gm_multi *m = (gm_multi *)h->multi;
  if (!m || !m->comm) return GM_OK;
  GM_CUDA(cudaEventRecord(m->ev0, h->stream));
  m->timed = true;
  int rc = GM_OK;
  if (h->d.ghost_lo) { rc = gm_iface_run(h, m->lo, false, d_nz, d_b, m->lo.buf, flags); ++*launches; }
  if (rc) return rc;
  GM_NCCL(ncclGroupStart());
  if (h->d.ghost_lo) GM_NCCL(ncclSend(m->lo.buf, (size_t)(m->lo.nlist + m->lo.ndof), ncclDouble, h->d.rank - 1, m->comm, h->stream));
  if (h->d.ghost_hi) GM_NCCL(ncclRecv(m->hi.buf, (size_t)(m->hi.nlist + m->hi.ndof), ncclDouble, h->d.rank + 1, m->comm, h->stream));
  GM_NCCL(ncclGroupEnd());
  if (h->d.ghost_hi) { rc = gm_iface_run(h, m->hi, true, d_nz, d_b, m->hi.buf, flags); ++*launches; }
  GM_CUDA(cudaEventRecord(m->ev1, h->stream));
  return rc;
}

// Overlapped variant (k_mma): the strip next to the lower interface was launched on m->xstream, the others on h->stream.
// xstream: pack the lower interface -> send to rank-1 / receive from rank+1 (grouped).  h->stream: after its own strips AND
// the receive, add the received lists into the upper interface rows.  Only the tiny unpack-add is serialised.
static int gm_multi_exchange_overlapped(gm_handle *h, double *d_nz, double *d_b, unsigned flags, int64_t *launches) {
  gm_multi *m = (gm_multi *)h->multi;
  GM_CUDA(cudaEventRecord(m->ev0, m->xstream));
  m->timed = true;
  int rc = GM_OK;
  if (h->d.ghost_lo) { rc = gm_iface_run(h, m->lo, false, d_nz, d_b, m->lo.buf, flags, m->xstream); ++*launches; }
  if (rc) return rc;
  GM_NCCL(ncclGroupStart());
  if (h->d.ghost_lo) GM_NCCL(ncclSend(m->lo.buf, (size_t)(m->lo.nlist + m->lo.ndof), ncclDouble, h->d.rank - 1, m->comm, m->xstream));
  if (h->d.ghost_hi) GM_NCCL(ncclRecv(m->hi.buf, (size_t)(m->hi.nlist + m->hi.ndof), ncclDouble, h->d.rank + 1, m->comm, m->xstream));
  GM_NCCL(ncclGroupEnd());
  GM_CUDA(cudaEventRecord(m->ev1, m->xstream));  // exchange_ms = pack + send / receive on the exchange stream (incl. waiting for the peers)
  GM_CUDA(cudaEventRecord(m->ev_join, m->xstream));
  GM_CUDA(cudaStreamWaitEvent(h->stream, m->ev_join, 0));
  if (h->d.ghost_hi) { rc = gm_iface_run(h, m->hi, true, d_nz, d_b, m->hi.buf, flags); ++*launches; }
  return rc;
}
