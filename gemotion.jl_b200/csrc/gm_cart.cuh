// gm_cart.cuh -- structured numeric path for Cartesian quad meshes: owner-computes row gather.
//
// Replaces the same reference interface as gm_scatter.cuh (jacobian!/residual!/
// residual_and_jacobian! of the operator built at /root/reference/src/Solver.jl:157, integrands
// Solver.jl:111-152, arithmetic SURVEY.md A.4) but never read-modify-writes nzval in HBM:
//
//  * A CTA owns a strip of TH=4 cell rows and sweeps it column by column in x.  The value lists
//    (CSC columns / CSR rows) of the P2 nodes of the current column are accumulated in SHARED
//    memory; a list is complete when the sweep has passed its node, and is then streamed to HBM
//    exactly once, contiguously, by a TMA bulk copy.  Cells of the strip's upper neighbour row (and
//    of the column left of an x-chunk) are evaluated as halo, only for the lists this CTA owns, so
//    no entry of any cell matrix is ever computed twice and no inter-CTA communication or atomics
//    are needed.
//  * Inside a column, a warp evaluates one third of a cell matrix: lane = (major node, minor
//    node) pair.  The pair's reference-element products (phi_i phi_j, d phi_i d phi_j, ... times
//    w|J|) are loop-invariant on a Cartesian mesh and live in REGISTERS; the per-cell state
//    (mu, grad u, grad T, S, u.grad phi) is broadcast from shared memory.  (Measured on B200:
//    DFMA with one operand streamed from the constant bank or shared memory tops out at 2.8-5.6
//    TDFMA/s, with register operands at 15.5 - profiles/r01_ubench_dfma_operand_source.txt.)
//  * The cell->nnz map is table driven (rank classes): cells whose 30x30 rank table is identical
//    share one 900-byte class table, so whatever numbering Gridap produced is honoured while the
//    map costs 4 bytes per cell instead of 837 x 8.
//  The kernel itself (pipeline, warp roles, buffers) is documented in gm_cart_kernel.cuh.
//
// Summation order differs from Gridap's (cells of a node are added left column first, even rows
// before odd rows); values agree with the oracle to ~1e-15 relative (tests/test_gpu_parity.py).
#pragma once
#include <cub/cub.cuh>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "gm_internal.h"
#include "gm_scatter.cuh"
#include "gm_cart_tables.h"

#define CT_XCH 64        // columns per x-chunk
#define CT_MAXCLASS 4096

// shared-memory list slots (doubles).  Node blocks: [list0][list1][list2][3 residual slots].
// Every slot is two elements longer than its longest list: a list is stored at slot + (global base & 1)
// so that shared and global addresses are 16-byte aligned together for the TMA bulk copy, and element
// cap-2(+parity) is a trash element that absorbs the entries of Dirichlet minors without a branch.
#define CT_VTX_L1 90
#define CT_VTX_L2 180
#define CT_VTX_RES 258
#define CT_VTX_BLK 262
#define CT_EDG_L1 54
#define CT_EDG_L2 108
#define CT_EDG_RES 156
#define CT_EDG_BLK 160
#define CT_CEN_L1 32
#define CT_CEN_L2 64
#define CT_CEN_RES 94
#define CT_CEN_BLK 98
#define CT_P_L 20
#define CT_P_RES 60
#define CT_P_BLK 64
#define CT_EVEN_STRIDE (CT_VTX_BLK + CT_EDG_BLK)                  // 422 per cell row
#define CT_ODD_STRIDE (CT_EDG_BLK + CT_CEN_BLK)                   // 258
#define CT_UNI 16                                                 // per (cell,q) uniform state
#define CT_NOD 6                                                  // per (cell,q,node) state

struct gm_cart {
  int nclass = 0;
  int32_t *cls = nullptr;     // [ncells] rank class of each cell
  uint8_t *ctab = nullptr;    // [nclass][900]   rank[minor][major]
  uint8_t *ptab = nullptr;    // [nclass][81][16] pair-major copy for k_gath (gm_cart_pair_ranks)
  double *visc = nullptr;     // [ncells][4][2]  (mu, kappa)
  CartPairTab *d_tab = nullptr;
  CartCell *d_cell = nullptr;
  gm_params last_prm;
  bool tab_valid = false;
  int th = 8;  // cell rows per strip (template parameter of k_cart)
  // kernel generation (gm_desc.kernel): GM_KERNEL_MMA (default) k_mma, GM_KERNEL_GATH / GM_KERNEL_GATH1 k_gath with
  // two / one trio of matrix warps, GM_KERNEL_SWEEP first-generation k_cart.  The older kernels stay selectable for A/B
  // measurements and are parity-tested.
  int kernel = GM_KERNEL_MMA;
  bool use_sweep = false;
  int gath_ng = 2;
};


// ------------------------------------------------------------------------------------------
// set-up kernels: rank classes + verification of the structural assumptions
// ------------------------------------------------------------------------------------------
__global__ void k_ct_hash(int64_t ncells, const uint16_t *__restrict__ rank, uint64_t *__restrict__ hash,
                          int32_t *__restrict__ cellid) {
  int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= ncells) return;
  const uint16_t *r = rank + c * 900;
  uint64_t h = 0x9E3779B97F4A7C15ull;
  for (int k = 0; k < 900; ++k) {
    h ^= (uint64_t)r[k] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h *= 0xBF58476D1CE4E5B9ull;
    h ^= h >> 29;
  }
  hash[c] = h;
  cellid[c] = (int32_t)c;
}

__global__ void k_ct_heads(int64_t n, const uint64_t *__restrict__ hs, int32_t *__restrict__ head) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) head[i] = (i == 0 || hs[i] != hs[i - 1]) ? 1 : 0;
}

__global__ void k_ct_assign(int64_t n, const int32_t *__restrict__ head, const int32_t *__restrict__ cid_incl,
                            const int32_t *__restrict__ cell_sorted, int32_t *__restrict__ cls, int32_t *__restrict__ rep,
                            int maxclass) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int id = cid_incl[i] - 1;
  if (id >= maxclass) return;
  cls[cell_sorted[i]] = id;
  if (head[i]) rep[id] = cell_sorted[i];
}

__global__ void k_ct_table(int nclass, const int32_t *__restrict__ rep, const uint16_t *__restrict__ rank,
                           uint8_t *__restrict__ ctab, int *__restrict__ err) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nclass * 900) return;
  const int k = t / 900, e = t - k * 900;
  const uint16_t r = rank[(int64_t)rep[k] * 900 + e];
  if (r != 0xFFFF && r >= 255) atomicExch(err, 2);
  ctab[t] = r == 0xFFFF ? 0xFF : (uint8_t)r;
}

// slot capacity of the list of local dof l (node type decides)
__device__ __forceinline__ int ct_cap(int l) {
  if (l >= 18 && l < 21) return 18;
  const bool isT = l >= 21;
  const int node = isT ? l - 21 : l % 9;
  if (node < 4) return isT ? 75 : 87;
  if (node < 8) return isT ? 45 : 51;
  return isT ? 27 : 30;
}

__global__ void k_ct_verify(int nx, int ny, const int32_t *__restrict__ g, const int64_t *__restrict__ ptr,
                            const uint16_t *__restrict__ rank, const int32_t *__restrict__ cls,
                            const uint8_t *__restrict__ ctab, int *__restrict__ err) {
  int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (c >= (int64_t)nx * ny) return;
  const int x = (int)(c % nx), y = (int)(c / nx);
  const int32_t *gc = g + c * 30;
  // (1) rank table equals the class table
  const uint16_t *r = rank + c * 900;
  const uint8_t *t = ctab + (int64_t)cls[c] * 900;
  for (int k = 0; k < 900; ++k) {
    const uint16_t a = r[k];
    if ((a == 0xFFFF ? 0xFF : (uint8_t)a) != t[k]) { atomicExch(err, 3); return; }
  }
  // (2) list lengths fit the slots
  for (int l = 0; l < 30; ++l) {
    const int32_t d = gc[l];
    if (d > 0 && ptr[d] - ptr[d - 1] > ct_cap(l)) { atomicExch(err, 4); return; }
  }
  // (3) conforming tensor-order sharing with the right and upper neighbours, U (both comps) and T
  const int offs[3] = {0, 9, 21};
  if (x + 1 < nx) {
    const int32_t *gr = gc + 30;
    for (int f = 0; f < 3; ++f)
      if (gc[offs[f] + 1] != gr[offs[f] + 0] || gc[offs[f] + 3] != gr[offs[f] + 2] || gc[offs[f] + 7] != gr[offs[f] + 6]) {
        atomicExch(err, 5); return;
      }
  }
  if (y + 1 < ny) {
    const int32_t *gu = gc + (int64_t)nx * 30;
    for (int f = 0; f < 3; ++f)
      if (gc[offs[f] + 2] != gu[offs[f] + 0] || gc[offs[f] + 3] != gu[offs[f] + 1] || gc[offs[f] + 5] != gu[offs[f] + 4]) {
        atomicExch(err, 5); return;
      }
  }
}

// ------------------------------------------------------------------------------------------
// viscosity pre-pass (n != 1): mu = gamma^((n-1)/2), kappa = ((n-1)/2) gamma^((n-3)/2) 4
// (Solver.jl:113-120).  One thread per (cell, quadrature point); 64 B/cell each way.
// ------------------------------------------------------------------------------------------
__global__ void k_ct_visc(int64_t ncells, const int32_t *__restrict__ g, const double *__restrict__ dvals,
                          const double *__restrict__ x, const CartCell *__restrict__ cc, double reg, double n,
                          double *__restrict__ visc) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= ncells * 4) return;
  const int64_t c = t >> 2;
  const int q = (int)(t & 3);
  double G00 = 0, G01 = 0, G10 = 0, G11 = 0;
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    const int32_t d0 = g[c * 30 + i], d1 = g[c * 30 + 9 + i];
    const double ux = d0 > 0 ? x[d0 - 1] : dvals[-d0 - 1];
    const double uy = d1 > 0 ? x[d1 - 1] : dvals[-d1 - 1];
    const double dx = cc->sdx[q][i], dy = cc->sdy[q][i];
    G00 += dx * ux; G01 += dx * uy; G10 += dy * ux; G11 += dy * uy;
  }
  const double D01 = 0.5 * (G01 + G10);
  const double gam = reg + 2.0 * (G00 * G00 + 2.0 * D01 * D01 + G11 * G11);
  const double mu = pow(gam, 0.5 * (n - 1.0));
  visc[t * 2] = mu;
  visc[t * 2 + 1] = 0.5 * (n - 1.0) * pow(gam, 0.5 * (n - 3.0)) * 4.0;
}

#include "gm_cart_kernel.cuh"
#include "gm_gath_kernel.cuh"
#include "gm_mma_kernel.cuh"

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static void gm_cart_destroy(gm_handle *h) {
  gm_cart *c = (gm_cart *)h->cart;
  if (!c) return;
  cudaFree(c->cls); cudaFree(c->ctab); cudaFree(c->ptab); cudaFree(c->visc); cudaFree(c->d_tab); cudaFree(c->d_cell);
  delete c;
  h->cart = nullptr;
}

static int gm_cart_create(gm_handle *h) {
  if (!h->d.cartesian || h->d.cell_type != GM_QUAD) { gm_set_error("structured path needs Cartesian quads"); return GM_ERR_ARG; }
  gm_cart *c = new (std::nothrow) gm_cart();
  if (!c) return GM_ERR_OOM;
  h->cart = c;
  c->kernel = h->d.kernel == GM_KERNEL_AUTO ? GM_KERNEL_MMA : h->d.kernel;
  c->use_sweep = c->kernel == GM_KERNEL_SWEEP;
  c->gath_ng = c->kernel == GM_KERNEL_GATH1 ? 1 : 2;
  h->active_path = GM_PATH_CART;  // confirmed (or revoked) by gm_cart_after_pattern
  return GM_OK;
}

// Builds rank classes from the per-cell rank table and verifies the structural assumptions.
// On any mismatch the handle silently stays on the scatter path (GM_PATH_AUTO) or fails (GM_PATH_CART).
static int gm_cart_after_pattern(gm_handle *h) {
  gm_cart *c = (gm_cart *)h->cart;
  const int64_t nc = h->ncells;
  uint64_t *hash = nullptr, *hash2 = nullptr;
  int32_t *cell = nullptr, *cell2 = nullptr, *head = nullptr, *cid = nullptr, *rep = nullptr;
  int *err = nullptr;
  void *tmp = nullptr;
  size_t tmpb = 0, t2 = 0;
  int rc = GM_OK, herr = 0, nclass = 0;
  const char *why = "";
#define CB(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { gm_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); rc = e__ == cudaErrorMemoryAllocation ? GM_ERR_OOM : GM_ERR_CUDA; goto done; } } while (0)
  CB(cudaMalloc(&hash, 8 * nc)); CB(cudaMalloc(&hash2, 8 * nc));
  CB(cudaMalloc(&cell, 4 * nc)); CB(cudaMalloc(&cell2, 4 * nc));
  CB(cudaMalloc(&head, 4 * nc)); CB(cudaMalloc(&cid, 4 * nc));
  CB(cudaMalloc(&rep, 4 * CT_MAXCLASS)); CB(cudaMalloc(&err, 4));
  CB(cudaMemsetAsync(err, 0, 4, h->stream));
  k_ct_hash<<<gm_blocks(nc, 128), 128, 0, h->stream>>>(nc, h->rank, hash, cell);
  cub::DeviceRadixSort::SortPairs(nullptr, tmpb, hash, hash2, cell, cell2, nc, 0, 64, h->stream);
  cub::DeviceScan::InclusiveSum(nullptr, t2, head, cid, nc, h->stream);
  if (t2 > tmpb) tmpb = t2;
  CB(cudaMalloc(&tmp, tmpb));
  CB(cub::DeviceRadixSort::SortPairs(tmp, tmpb, hash, hash2, cell, cell2, nc, 0, 64, h->stream));
  k_ct_heads<<<gm_blocks(nc, 256), 256, 0, h->stream>>>(nc, hash2, head);
  CB(cub::DeviceScan::InclusiveSum(tmp, tmpb, head, cid, nc, h->stream));
  CB(cudaMemcpyAsync(&nclass, cid + nc - 1, 4, cudaMemcpyDeviceToHost, h->stream));
  CB(cudaStreamSynchronize(h->stream));
  if (nclass > CT_MAXCLASS) { why = "too many distinct cell->nnz rank tables (irregular numbering)"; herr = 1; goto decided; }
  CB(cudaMalloc(&c->cls, 4 * nc));
  CB(cudaMalloc(&c->ctab, (size_t)nclass * 900));
  k_ct_assign<<<gm_blocks(nc, 256), 256, 0, h->stream>>>(nc, head, cid, cell2, c->cls, rep, CT_MAXCLASS);
  k_ct_table<<<gm_blocks((int64_t)nclass * 900, 256), 256, 0, h->stream>>>(nclass, rep, h->rank, c->ctab, err);
  k_ct_verify<<<gm_blocks(nc, 128), 128, 0, h->stream>>>(h->d.nx, h->d.ny, h->gdofs, h->ptr, h->rank, c->cls, c->ctab, err);
  CB(cudaMemcpyAsync(&herr, err, 4, cudaMemcpyDeviceToHost, h->stream));
  CB(cudaStreamSynchronize(h->stream));
  if (herr) why = herr == 2 ? "a list is longer than 254 entries" : herr == 3 ? "rank-class hash collision"
                : herr == 4 ? "a list exceeds its shared-memory slot" : "dof tables are not a conforming tensor-order Cartesian numbering";
decided:
  if (herr) {
    if (h->d.path == GM_PATH_CART) {
      gm_set_error("structured path unavailable: %s", why);
      rc = GM_ERR_ARG;
    } else {
      gm_cart_destroy(h);
      h->active_path = GM_PATH_SCATTER;
    }
    goto done;
  }
  c->nclass = nclass;
  h->device_bytes += 4 * nc + (int64_t)nclass * 900;
  {  // pair-major rank records for k_gath
    std::vector<uint8_t> ct((size_t)nclass * 900), pt((size_t)nclass * 81 * 16);
    CB(cudaMemcpy(ct.data(), c->ctab, ct.size(), cudaMemcpyDeviceToHost));
    gm_cart_pair_ranks(nclass, ct.data(), pt.data());
    CB(cudaMalloc(&c->ptab, pt.size()));
    CB(cudaMemcpy(c->ptab, pt.data(), pt.size(), cudaMemcpyHostToDevice));
    h->device_bytes += (int64_t)pt.size();
  }
  CB(cudaMalloc(&c->visc, sizeof(double) * 8 * nc));
  CB(cudaMalloc(&c->d_tab, sizeof(CartPairTab) * 81));
  CB(cudaMalloc(&c->d_cell, sizeof(CartCell)));
  h->device_bytes += 64 * nc;
  // the per-cell rank table is no longer needed (the class table replaces it)
  cudaFree(h->rank);
  h->rank = nullptr;
  h->device_bytes -= (int64_t)sizeof(uint16_t) * nc * 900;
  {
    c->th = 4;  // TH=4 (2 CTAs/SM) measured faster than TH=8 (1 CTA/SM: 30.8 vs 24.2 ms at 2048^2, v4); later versions are TH=4 only
  }
done:
  cudaFree(hash); cudaFree(hash2); cudaFree(cell); cudaFree(cell2); cudaFree(head); cudaFree(cid); cudaFree(rep);
  cudaFree(err); cudaFree(tmp);
#undef CB
  return rc;
}

// pair / cell constant tables from the reference tables and the current parameters
static int gm_cart_tables(gm_handle *h) {
  gm_cart *c = (gm_cart *)h->cart;
  if (c->tab_valid && memcmp(&c->last_prm, &h->prm, sizeof(gm_params)) == 0) return GM_OK;
  const GmConst &k = h->hc;
  const gm_params &p = h->prm;
  static thread_local CartPairTab tab[81];
  static thread_local CartCell cell;
  gm_cart_fill_tables(k.w, k.phi, k.dphi, k.psi, k.hx, k.hy, p.Pr, p.Ra, p.gx, p.gy, p.jac_scaling, tab, &cell);
  GM_CUDA(cudaMemcpyAsync(c->d_tab, tab, sizeof tab, cudaMemcpyHostToDevice, h->stream));
  GM_CUDA(cudaMemcpyAsync(c->d_cell, &cell, sizeof cell, cudaMemcpyHostToDevice, h->stream));
  GM_CUDA(cudaStreamSynchronize(h->stream));  // the staging arrays are reused by the next call
  c->last_prm = h->prm;
  c->tab_valid = true;
  return GM_OK;
}

template <int TH, bool CSR, bool NEWT>
static int gm_cart_launch1(gm_handle *h, const CartArgs &a) {
  const size_t sm = sizeof(CtSmem<TH>);
  // per-DEVICE attribute: set on every launch (a process may hold handles on several devices)
  GM_CUDA(cudaFuncSetAttribute(k_cart<TH, CSR, NEWT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  dim3 grid((a.nx + CT_XCH - 1) / CT_XCH, (a.row_end - a.row_begin + TH - 1) / TH);
  k_cart<TH, CSR, NEWT><<<grid, (3 * TH / 2 + 2) * 32, sm, h->stream>>>(a);
  return GM_OK;
}
template <int TH>
static int gm_cart_launch(gm_handle *h, const CartArgs &a, bool csr, bool newt) {
  if (csr) return newt ? gm_cart_launch1<TH, true, true>(h, a) : gm_cart_launch1<TH, true, false>(h, a);
  return newt ? gm_cart_launch1<TH, false, true>(h, a) : gm_cart_launch1<TH, false, false>(h, a);
}

// ---- k_gath (gm_gath_kernel.cuh): NG trios of matrix warps + GA_NAUX aux warps, one CTA per SM
template <int NG, bool CSR, bool NEWT, int XCH>
static int gm_gath_launch2(gm_handle *h, const CartArgs &a, int nstrips) {
  const size_t sm = sizeof(GaSmem<NG>);
  GM_CUDA(cudaFuncSetAttribute(k_gath<NG, CSR, NEWT, XCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  dim3 grid((a.nx + 1 + XCH - 1) / XCH, nstrips);
  k_gath<NG, CSR, NEWT, XCH><<<grid, (3 * NG + GA_NAUX) * 32, sm, h->stream>>>(a);
  return GM_OK;
}
template <int NG, bool CSR, bool NEWT>
static int gm_gath_launch1(gm_handle *h, const CartArgs &a) {
  const int nstrips = (a.row_end - a.row_begin + 1 + 4 * NG - 1) / (4 * NG);
  const int xch = h->d.xchunk ? h->d.xchunk : gm_gath_chunk(a.nx, nstrips, h->num_sms);
  return xch <= 128 ? gm_gath_launch2<NG, CSR, NEWT, 128>(h, a, nstrips) : gm_gath_launch2<NG, CSR, NEWT, 256>(h, a, nstrips);
}
template <int NG>
static int gm_gath_launch(gm_handle *h, const CartArgs &a, bool csr, bool newt) {
  if (csr) return newt ? gm_gath_launch1<NG, true, true>(h, a) : gm_gath_launch1<NG, true, false>(h, a);
  return newt ? gm_gath_launch1<NG, false, true>(h, a) : gm_gath_launch1<NG, false, false>(h, a);
}

// ---- k_mma (gm_mma_kernel.cuh): 4 main + 4 aux warps, 8 quad rows, two CTAs per SM.  Strips [strip0, strip0 + nstrips) on `stream`.
template <bool CSR, bool NEWT>
static int gm_mma_launch2(gm_handle *h, const CartArgs &a, int xch, int strip0, int nstrips, cudaStream_t stream) {
  const size_t sm = sizeof(MmSmem);
  GM_CUDA(cudaFuncSetAttribute(k_mma<CSR, NEWT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));  // (per device: set on every launch)
  GM_CUDA(cudaFuncSetAttribute(k_mma<CSR, NEWT>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));  // two CTAs per SM
  MmArgs ma;
  ma.a = a;
  memcpy(ma.out, MM_OUT, sizeof ma.out);
  ma.xch = xch;
  ma.strip0 = strip0;
  dim3 grid((a.nx + 1 + xch - 1) / xch, nstrips);
  k_mma<CSR, NEWT><<<grid, MM_NT, sm, stream>>>(ma);
#ifdef MM_PROF
  {  // development build: cycles per warp and phase of the middle CTA (profiles/r02_k_mma_phase_timing.txt)
    long long hp[8][8];
    cudaStreamSynchronize(stream);
    cudaMemcpyFromSymbol(hp, mm_prof, sizeof hp);
    for (int w = 0; w < 8; ++w) {
      fprintf(stderr, "mm_prof warp %d:", w);
      for (int k = 0; k < 8; ++k) fprintf(stderr, " %9lld", hp[w][k]);
      fprintf(stderr, "\n");
    }
  }
#endif
  return GM_OK;
}
static int gm_mma_launch(gm_handle *h, const CartArgs &a, bool csr, bool newt, int xch, int strip0, int nstrips, cudaStream_t stream) {
  if (csr) return newt ? gm_mma_launch2<true, true>(h, a, xch, strip0, nstrips, stream) : gm_mma_launch2<true, false>(h, a, xch, strip0, nstrips, stream);
  return newt ? gm_mma_launch2<false, true>(h, a, xch, strip0, nstrips, stream) : gm_mma_launch2<false, false>(h, a, xch, strip0, nstrips, stream);
}

static int gm_cart_run(gm_handle *h, const double *d_x, double *d_nz, double *d_b, unsigned flags, int64_t *launches) {
  gm_cart *c = (gm_cart *)h->cart;
  int rc = gm_cart_tables(h);
  if (rc) return rc;
  const gm_params &p = h->prm;
  CartArgs a;
  a.nx = h->d.nx; a.ny = h->d.ny;
  a.row_begin = h->d.ghost_lo; a.row_end = h->d.ny - h->d.ghost_hi;
  a.gdofs = h->gdofs; a.dvals = h->dvals; a.ptr = h->ptr; a.cls = c->cls; a.ctab = c->ctab; a.ptab = c->ptab; a.visc = c->visc;
  a.tab = c->d_tab; a.cc = c->d_cell; a.x = d_x; a.nz = d_nz; a.b = d_b; a.flags = flags;
  a.Pr = p.Pr; a.RaPr = p.Ra * p.Pr; a.gx = p.gx; a.gy = p.gy; a.reg = p.reg; a.n = p.n; a.js = p.jac_scaling;
  const bool newt = p.n == 1.0;
  if (!newt) {
    k_ct_visc<<<gm_blocks(h->ncells * 4, 256), 256, 0, h->stream>>>(h->ncells, h->gdofs, h->dvals, d_x, c->d_cell, p.reg, p.n, c->visc);
    ++*launches;
  }
  const bool csr = h->d.layout == GM_LAYOUT_CSR;
  if (c->kernel == GM_KERNEL_MMA) {
    const int nstrips = (a.row_end - a.row_begin + 1 + MM_QH - 1) / MM_QH;
    const int xch = h->d.xchunk > 0 ? h->d.xchunk : gm_mma_chunk(a.nx, nstrips, 2 * h->num_sms, MM_WARM);
    gm_multi *m = (gm_multi *)h->multi;
    const bool overlap = m && m->comm && !(flags & GM_NO_EXCHANGE) && (h->d.ghost_lo || h->d.ghost_hi);
    if (overlap) {
      // the strip next to the lower interface first, on the high-priority exchange stream; its lists are packed and
      // sent to rank-1 (and the lists of rank+1 received) while the remaining strips run on the main stream
      GM_CUDA(cudaEventRecord(m->ev_fork, h->stream));
      GM_CUDA(cudaStreamWaitEvent(m->xstream, m->ev_fork, 0));
      const int first = (h->d.ghost_lo && nstrips > 1) ? 1 : 0;  // strips [0, first) go to the exchange stream
      if (first) { rc = gm_mma_launch(h, a, csr, newt, xch, 0, first, m->xstream); ++*launches; }
      if (rc) return rc;
      rc = gm_mma_launch(h, a, csr, newt, xch, first, nstrips - first, h->stream);
      ++*launches;
      if (rc) return rc;
      GM_CUDA(cudaGetLastError());
      if (h->d.ghost_lo && !first) {  // a single strip: the pack must wait for it
        GM_CUDA(cudaEventRecord(m->ev_join, h->stream));
        GM_CUDA(cudaStreamWaitEvent(m->xstream, m->ev_join, 0));
      }
      return gm_multi_exchange_overlapped(h, d_nz, d_b, flags, launches);
    }
    rc = gm_mma_launch(h, a, csr, newt, xch, 0, nstrips, h->stream);
  } else if (c->use_sweep) rc = gm_cart_launch<4>(h, a, csr, newt);
  else rc = c->gath_ng == 1 ? gm_gath_launch<1>(h, a, csr, newt) : gm_gath_launch<2>(h, a, csr, newt);
  if (rc) return rc;
  ++*launches;
  GM_CUDA(cudaGetLastError());
  if (!(flags & GM_NO_EXCHANGE)) return gm_multi_exchange(h, d_nz, d_b, flags, launches);
  return GM_OK;
}
