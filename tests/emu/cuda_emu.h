// cuda_emu.h -- TEST INFRASTRUCTURE ONLY: runs a CUDA kernel's source as host C++, one pthread per CUDA
// thread, so that the index logic of gm_gath_kernel.cuh can be checked against the oracle in a container
// without a GPU (tests/test_gath_emu.py).  Nothing under gemotion.jl_b200/ links or loads this.
//
// Asynchronous copies are modelled pessimistically: cp.async copies are performed when the issuing thread
// waits for them, bulk shared->global copies read shared memory only at wait_group.read -- so data used
// before its wait, or a staging buffer recycled too early, shows up as wrong numbers.  Shared memory is
// poisoned at block start.
#pragma once
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <vector>

#define GM_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __constant__
#define __maxnreg__(...)
#define __restrict__
#define __align__(x)

struct emu_dim3 { unsigned x = 1, y = 1, z = 1; };
static thread_local emu_dim3 threadIdx, blockIdx;
static emu_dim3 blockDim, gridDim;
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
struct uint4 { uint32_t x, y, z, w; };
using std::min;
using std::max;
using std::fma;

#define GM_JACOBIAN 1u
#define GM_RESIDUAL 2u

static unsigned char *g_emu_smem = nullptr;
static inline unsigned char *emu_smem() { return g_emu_smem; }
static pthread_barrier_t g_emu_bar_all, g_emu_bar_warp[64];
static double g_emu_xbuf[64][32];
static inline void __syncthreads() { pthread_barrier_wait(&g_emu_bar_all); }
template <int NT>
static inline void ga_bar_step() { pthread_barrier_wait(&g_emu_bar_all); }
// full-warp butterfly shuffle: all 32 threads of the warp must call it
static inline double ga_shfl_xor(double v, int mask) {
  const unsigned w = threadIdx.x >> 5, l = threadIdx.x & 31;
  g_emu_xbuf[w][l] = v;
  pthread_barrier_wait(&g_emu_bar_warp[w]);
  const double r = g_emu_xbuf[w][l ^ mask];
  pthread_barrier_wait(&g_emu_bar_warp[w]);
  return r;
}

// ---- fp64 tensor instruction mma.sync.m8n8k4 (k_mma): all 32 threads of the warp must call it.
// Fragments as in PTX: A[row = lane / 4][k = lane % 4], B[k = lane % 4][col = lane / 4], D[row = lane / 4][col = 2 (lane % 4) + {0, 1}].
static double g_emu_ma[64][32], g_emu_mb[64][32];
static inline void mm_dmma(double (&d)[2], double a, double b) {
  const unsigned w = threadIdx.x >> 5, l = threadIdx.x & 31;
  g_emu_ma[w][l] = a; g_emu_mb[w][l] = b;
  pthread_barrier_wait(&g_emu_bar_warp[w]);
  const unsigned g = l >> 2, j = l & 3;
  for (int c = 0; c < 2; ++c) {
    double acc = d[c];
    for (int k = 0; k < 4; ++k) acc = fma(g_emu_ma[w][4 * g + k], g_emu_mb[w][4 * (2 * j + c) + k], acc);
    d[c] = acc;
  }
  pthread_barrier_wait(&g_emu_bar_warp[w]);
}
// ---- mbarrier: low half = pending arrivals, high half = parity of the current (incomplete) phase
static int g_emu_mbar_count = 0;
static inline void mm_mbar_init(unsigned long long *bar, int count) { g_emu_mbar_count = count; __atomic_store_n(bar, (unsigned long long)count, __ATOMIC_RELEASE); }
static inline void mm_mbar_arrive(unsigned long long *bar) {
  unsigned long long old = __atomic_load_n(bar, __ATOMIC_ACQUIRE), nw;
  do {
    unsigned pend = (unsigned)old - 1, ph = (unsigned)(old >> 32);
    if (pend == 0) { pend = (unsigned)g_emu_mbar_count; ph ^= 1; }
    nw = ((unsigned long long)ph << 32) | pend;
  } while (!__atomic_compare_exchange_n(bar, &old, nw, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE));
}
static inline void mm_mbar_wait(unsigned long long *bar, unsigned parity) {
  while ((unsigned)(__atomic_load_n(bar, __ATOMIC_ACQUIRE) >> 32) == parity) sched_yield();
}
static inline void mm_syncwarp() { pthread_barrier_wait(&g_emu_bar_warp[threadIdx.x >> 5]); }
static inline double mm_mask(double v, uint32_t mw, uint32_t sign) {
  uint64_t b;
  memcpy(&b, &v, 8);
  b = (((b >> 32) & mw) ^ sign) << 32 | ((uint32_t)b & mw);
  memcpy(&v, &b, 8);
  return v;
}

struct EmuCopy { void *dst; const void *src; int bytes; };
static thread_local std::vector<EmuCopy> t_emu_cp, t_emu_bulk;
static inline void ga_cp4(void *d, const void *s) { t_emu_cp.push_back({d, s, 4}); }
static inline void ga_cp8(void *d, const void *s) { t_emu_cp.push_back({d, s, 8}); }
static inline void ga_cp16(void *d, const void *s) { t_emu_cp.push_back({d, s, 16}); }
static inline void ga_cp_wait_all() {
  for (auto &c : t_emu_cp) memcpy(c.dst, c.src, c.bytes);
  t_emu_cp.clear();
}
// groups (k_mma): copies are performed as LATE as the program allows -- when the thread waits for their group
static thread_local std::vector<size_t> t_emu_cp_groups;  // end index (in t_emu_cp) of every committed group
static inline void ga_cp_commit() { t_emu_cp_groups.push_back(t_emu_cp.size()); }
static inline void ga_cp_wait_prev() {  // wait_group 1: everything but the newest committed group
  if (t_emu_cp_groups.size() < 2) return;
  const size_t n = t_emu_cp_groups[t_emu_cp_groups.size() - 2];
  for (size_t k = 0; k < n; ++k) memcpy(t_emu_cp[k].dst, t_emu_cp[k].src, t_emu_cp[k].bytes);
  t_emu_cp.erase(t_emu_cp.begin(), t_emu_cp.begin() + n);
  const size_t last = t_emu_cp_groups.back() - n;
  t_emu_cp_groups.clear();
  t_emu_cp_groups.push_back(last);
}
static int64_t g_emu_bulk_misaligned = 0;
static inline void ga_bulk_store(double *g, const double *s, int bytes) {
  if (((uintptr_t)g & 15) || ((uintptr_t)s & 15) || (bytes & 15) || bytes <= 0) __atomic_add_fetch(&g_emu_bulk_misaligned, 1, __ATOMIC_RELAXED);
  t_emu_bulk.push_back({g, s, bytes});
}
static inline void ga_bulk_commit() {}
static inline void ga_bulk_wait_read() {
  for (auto &c : t_emu_bulk) memcpy(c.dst, c.src, c.bytes);
  t_emu_bulk.clear();
}
static inline void ga_fence_async() {}

// ---- launch: blocks one after another, threads of a block concurrently
template <typename Args>
struct EmuLaunch {
  void (*kernel)(Args);
  Args args;
  emu_dim3 bidx;
  unsigned t;
};
template <typename Args>
static void *emu_thread_main(void *p) {
  EmuLaunch<Args> *l = (EmuLaunch<Args> *)p;
  threadIdx.x = l->t; threadIdx.y = 0; threadIdx.z = 0;
  blockIdx = l->bidx;
  t_emu_cp.clear(); t_emu_bulk.clear(); t_emu_cp_groups.clear();
  l->kernel(l->args);
  return nullptr;
}
template <typename Args>
static int emu_launch(void (*kernel)(Args), Args args, unsigned gx, unsigned gy, unsigned nthreads, unsigned naux, size_t smem_bytes) {
  gridDim.x = gx; gridDim.y = gy; blockDim.x = nthreads;
  std::vector<unsigned char> smem(smem_bytes + 64);
  g_emu_smem = (unsigned char *)(((uintptr_t)smem.data() + 63) & ~(uintptr_t)63);
  std::vector<pthread_t> th(nthreads);
  std::vector<EmuLaunch<Args>> la(nthreads);
  pthread_attr_t attr;
  pthread_attr_init(&attr);
  pthread_attr_setstacksize(&attr, 1 << 20);
  for (unsigned by = 0; by < gy; ++by)
    for (unsigned bx = 0; bx < gx; ++bx) {
      memset(g_emu_smem, 0x7F, smem_bytes);  // poison: huge doubles, huge positive ints (a stale dof id faults)
      pthread_barrier_init(&g_emu_bar_all, nullptr, nthreads);
      for (unsigned w = 0; w < nthreads / 32; ++w) pthread_barrier_init(&g_emu_bar_warp[w], nullptr, 32);
      for (unsigned t = 0; t < nthreads; ++t) {
        la[t].kernel = kernel; la[t].args = args; la[t].bidx.x = bx; la[t].bidx.y = by; la[t].t = t;
        if (pthread_create(&th[t], &attr, emu_thread_main<Args>, &la[t])) return -1;
      }
      for (unsigned t = 0; t < nthreads; ++t) pthread_join(th[t], nullptr);
      pthread_barrier_destroy(&g_emu_bar_all);
      for (unsigned w = 0; w < nthreads / 32; ++w) pthread_barrier_destroy(&g_emu_bar_warp[w]);
    }
  pthread_attr_destroy(&attr);
  return 0;
}
