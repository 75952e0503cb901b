// gm_internal.h -- handle layout and helpers shared by the kernels of libgemotion_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/gemotion_b200.h"
#include "gm_const.h"

struct gm_params {
  double Pr, Ra, n, reg, gx, gy, jac_scaling;
};

struct gm_handle {
  gm_desc d;  // scalars only; every pointer member is nulled after gm_create
  int dev = 0;
  int num_sms = 148;  // multiprocessors of the device (grid sizing)
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int ld = 0, nn = 0, nq = 0, nv = 0;
  int64_t ncells = 0, n_total = 0, n_diri_total = 0;
  int64_t off[3] = {0, 0, 0};
  GmConst hc;         // host copy of the constant block
  gm_params prm;
  // device tables
  int32_t *gdofs = nullptr;  // [ncells*ld] signed; >0 global free id (1-based); <0 -> dvals[-id-1]
  double *dvals = nullptr;   // [n_diri_total] (U | P | T)
  double *coords = nullptr;  // [nverts*2]
  int32_t *cellv = nullptr;  // [ncells*nv] 0-based
  // pattern
  bool has_pattern = false;
  int64_t nnz = 0;
  int64_t *ptr = nullptr;    // [n_total+1]
  int32_t *idx = nullptr;    // [nnz] 0-based minor indices (may be released for huge problems)
  uint16_t *rank = nullptr;  // [ncells*ld*ld] position of (major M, minor m) inside M's list: [c][m][M]
  int max_row_len = 0;
  // scatter colouring
  int ncolors = 0;
  int32_t *color_cells = nullptr;      // generic meshes: cells grouped by colour
  std::vector<int64_t> color_ptr;      // [ncolors+1]
  // structured path (gm_cart.cuh)
  int active_path = GM_PATH_SCATTER;
  void *cart = nullptr;
  void *multi = nullptr;  // gm_multi (row-block interface exchange)
  void *post = nullptr;   // gm_post (post-processing assemblies, gm_post.cuh)
  // generic owner-computes path (gm_gather.cuh): adjacency dof -> (cell*ld + local) in ascending order, compact cell->nnz map
  int64_t *g_adjptr = nullptr;
  int32_t *g_adj = nullptr;
  uint16_t *g_rank = nullptr, *g_rankp = nullptr;
  void *gather = nullptr;  // gm_gather: workspaces + device copy of the constants
  int64_t n_own_cells = 0;        // unstructured partition: own cells first, then ghosts (== ncells on one GPU)
  int64_t *ghost_ids = nullptr;   // [ncells - n_own_cells] global ids of the ghost cells
  // staging for the host API
  double *d_x = nullptr, *d_b = nullptr, *d_nz = nullptr;
  int64_t d_nz_cap = 0;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  gm_times times;
  bool times_pending = false;  // kernel_ms / exchange_ms of the last device assembly not read back yet
  int64_t n_owned = 0;         // dofs owned by this rank (row blocks); 0: all
  // host API on a rank of a partition: only the dof ranges its cells touch are copied (iterate in, residual out)
  std::vector<std::pair<int64_t, int64_t>> xruns;  // [begin, end) runs, ascending; empty: everything
  int64_t device_bytes = 0;
};

void gm_set_error(const char *fmt, ...);

#define GM_CUDA(call)                                                                       \
  do {                                                                                      \
    cudaError_t e__ = (call);                                                               \
    if (e__ != cudaSuccess) {                                                               \
      gm_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));  \
      return (e__ == cudaErrorMemoryAllocation) ? GM_ERR_OOM : GM_ERR_CUDA;                 \
    }                                                                                       \
  } while (0)

template <typename T>
static inline int gm_dalloc(gm_handle *h, T **p, int64_t count) {
  if (count <= 0) count = 1;
  cudaError_t e = cudaMalloc((void **)p, sizeof(T) * (size_t)count);
  if (e != cudaSuccess) {
    gm_set_error("cudaMalloc(%lld bytes) -> %s", (long long)(sizeof(T) * count), cudaGetErrorString(e));
    *p = nullptr;
    return GM_ERR_OOM;
  }
  h->device_bytes += (int64_t)sizeof(T) * count;
  return GM_OK;
}
