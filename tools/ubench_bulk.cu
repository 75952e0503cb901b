// Microbenchmark (design evidence): issue rate of small shared->global TMA bulk copies
// (cp.async.bulk.global.shared::cta, SASS UBLKCP) per SM as a function of the copy size, next to plain 16-byte
// vector stores of the same bytes.  k_gath / k_mma write every completed value list (~240-720 B) with one bulk copy.
// Build on the GPU box:  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -cudart shared -o /tmp/ubench_bulk tools/ubench_bulk.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

// MODE 0: every thread of NTH threads issues its own bulk copies (one list per thread, as the kernels do)
// MODE 1: plain STG.128: a warp writes one list with consecutive lanes
template <int MODE>
__global__ void k(double *out, int bytes, int ncopies, int nth, long long *cyc) {
  extern __shared__ __align__(128) unsigned char sm[];
  for (int i = threadIdx.x; i < 48 * 1024 / 8; i += blockDim.x) ((double *)sm)[i] = i;
  __syncthreads();
  const long long t0 = clock64();
  double *base = out + (size_t)blockIdx.x * (64u << 20) / 8;  // 64 MB window per CTA, walked linearly
  if (MODE == 0) {
    if (threadIdx.x < nth) {
      for (int c = threadIdx.x; c < ncopies; c += nth) {
        const uint32_t sa = (uint32_t)__cvta_generic_to_shared(sm + ((c * bytes) % (40 * 1024) & ~15));
        double *g = base + (size_t)c * bytes / 8;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(sa), "r"(bytes) : "memory");
        if ((c / nth) % 4 == 3) { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory"); }
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
  } else {
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int c = w; c < ncopies; c += nw) {
      const double2 *s = (const double2 *)(sm + ((c * bytes) % (40 * 1024) & ~15));
      double2 *g = (double2 *)(base + (size_t)c * bytes / 8);
      for (int k2 = l; k2 < bytes / 16; k2 += 32) g[k2] = s[k2];
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  const int nsm = 148;
  double *out; long long *cyc, hc;
  cudaMalloc(&out, (size_t)nsm * 2 * (64u << 20));
  cudaMalloc(&cyc, 8);
  cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
  cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
  printf("# shared->global copies, one CTA per SM (148) or two (296); cycles from clock64 of CTA 0; 1965 MHz\n");
  for (int ctas = 1; ctas <= 2; ++ctas)
    for (int mode = 0; mode < 2; ++mode)
      for (int bytes : {240, 432, 720, 1440, 2880, 5760}) {
        for (int nth : {32, 128}) {
          if (mode == 1 && nth == 32) continue;
          const int ncopies = (int)((48u << 20) / bytes);  // 48 MB per CTA
          cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
          float ms = 0;
          for (int it = 0; it < 2; ++it) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<nsm * ctas, 128, 48 * 1024>>>(out, bytes, ncopies, nth, cyc);
            else k<1><<<nsm * ctas, 128, 48 * 1024>>>(out, bytes, ncopies, nth, cyc);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
          }
          cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
          const double tot = (double)ncopies * bytes * nsm * ctas;
          printf("%s ctas/SM=%d bytes=%5d issuing threads=%3d : %8.3f ms  %7.1f GB/s  %7.1f cycles per copy per SM  (%s)\n", mode == 0 ? "bulk" : "stg ",
                 ctas, bytes, nth, ms, tot / ms * 1e-6, (double)hc / ncopies / ctas, cudaGetErrorString(cudaGetLastError()));
        }
      }
  return 0;
}
