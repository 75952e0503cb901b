// Microbenchmark (design evidence for the structured path): rate and latency of the fp64
// tensor instruction (mma.sync .f64: SASS DMMA) on B200 next to plain DFMA, and whether the two share an issue
// port.  Build on the GPU box:  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -cudart shared -o /tmp/ubench_dmma tools/ubench_dmma.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1684(double (&d)[4], double a0, double a1, double b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a0), "d"(a1), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&d)[4], const double (&a)[4], double b0, double b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b0), "d"(b1));
}
__device__ __forceinline__ void dmma16816(double (&d)[4], const double (&a)[8], const double (&b)[4]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// MODE 0: m8n8k4, NACC independent accumulator pairs   1: m16n8k4   2: m16n8k8   3: m16n8k16
// MODE 4: DFMA only, NACC*2 independent chains          5: m8n8k4 + NF DFMA per DMMA (mixed issue)
template <int MODE, int NACC, int NF>
__global__ void k(double *out, int reps, long long *cyc) {
  const int lane = threadIdx.x & 31;
  double a[8], b[4];
  for (int i = 0; i < 8; ++i) a[i] = 1e-3 * (lane + i);
  for (int i = 0; i < 4; ++i) b[i] = 1e-3 * (lane - i);
  double acc[NACC][4];
  double f[NF > 0 ? NF : 1];
  for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) acc[i][j] = i + j;
  for (int i = 0; i < (NF > 0 ? NF : 1); ++i) f[i] = i;
  const long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      if (MODE == 0 || MODE == 5) dmma884(acc[i][0], acc[i][1], a[i & 7], b[i & 3]);
      if (MODE == 1) dmma1684(acc[i], a[i & 7], a[(i + 1) & 7], b[i & 3]);
      if (MODE == 2) { double aa[4] = {a[0], a[1], a[2], a[3]}; dmma1688(acc[i], aa, b[0], b[1]); }
      if (MODE == 3) dmma16816(acc[i], a, b);
      if (MODE == 4) { acc[i][0] = fma(a[i & 7], b[i & 3], acc[i][0]); acc[i][1] = fma(a[(i + 1) & 7], b[i & 3], acc[i][1]); }
      if (MODE == 5) {
#pragma unroll
        for (int j = 0; j < NF; ++j) f[j] = fma(a[j & 7], b[j & 3], f[j]);
      }
    }
  }
  const long long t1 = clock64();
  double v = 0;
  for (int i = 0; i < NACC; ++i) for (int j = 0; j < 4; ++j) v += acc[i][j];
  for (int i = 0; i < (NF > 0 ? NF : 1); ++i) v += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = v;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE, int NACC, int NF>
void run(const char *name, int nw) {
  const int nsm = 148, reps = 4000;
  double *out; long long *cyc, hc;
  cudaMalloc(&out, sizeof(double) * nsm * nw * 32);
  cudaMalloc(&cyc, 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0;
  for (int it = 0; it < 2; ++it) {
    cudaEventRecord(e0);
    k<MODE, NACC, NF><<<nsm, nw * 32>>>(out, reps, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
  }
  cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost);
  const double instr = (double)reps * NACC;  // tensor (or DFMA-pair) instructions per warp
  const double fma_per = MODE == 0 || MODE == 5 ? 256 : MODE == 1 ? 512 : MODE == 2 ? 1024 : MODE == 3 ? 2048 : 64;
  const double tf = instr * fma_per * nw * nsm / (ms * 1e-3) * 1e-12;
  const double extra = MODE == 5 ? instr * NF * 32.0 * nw * nsm / (ms * 1e-3) * 1e-12 : 0;
  printf("%-14s NACC=%2d NF=%2d warps/SM=%2d : %8.3f ms  %6.2f TFMA/s (+%5.2f TDFMA/s plain)  %7.2f cycles per instr per warp, %6.2f cycles per instr per SM  (%s)\n",
         name, NACC, NF, nw, ms, tf, extra, (double)hc / instr, (double)hc / (instr * nw), cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  printf("# fp64 tensor instruction (DMMA) vs DFMA, 148 CTAs (one per SM), 4000 iterations; cycles from clock64 of CTA 0\n");
  run<0, 1, 0>("m8n8k4 latency", 1);
  run<0, 1, 0>("m8n8k4", 4);
  run<0, 2, 0>("m8n8k4", 4);
  run<0, 4, 0>("m8n8k4", 4);
  run<0, 8, 0>("m8n8k4", 4);
  run<0, 8, 0>("m8n8k4", 8);
  run<0, 8, 0>("m8n8k4", 12);
  run<0, 8, 0>("m8n8k4", 16);
  run<0, 4, 0>("m8n8k4", 32);
  run<1, 1, 0>("m16n8k4 lat", 1);
  run<1, 4, 0>("m16n8k4", 4);
  run<1, 4, 0>("m16n8k4", 8);
  run<1, 4, 0>("m16n8k4", 16);
  run<2, 1, 0>("m16n8k8 lat", 1);
  run<2, 4, 0>("m16n8k8", 4);
  run<2, 4, 0>("m16n8k8", 8);
  run<2, 4, 0>("m16n8k8", 16);
  run<3, 1, 0>("m16n8k16 lat", 1);
  run<3, 4, 0>("m16n8k16", 4);
  run<3, 4, 0>("m16n8k16", 8);
  run<3, 4, 0>("m16n8k16", 16);
  run<4, 1, 0>("dfma latency", 1);
  run<4, 8, 0>("dfma", 4);
  run<4, 8, 0>("dfma", 8);
  run<4, 8, 0>("dfma", 16);
  run<5, 4, 2>("m8n8k4+2dfma", 8);
  run<5, 4, 4>("m8n8k4+4dfma", 8);
  run<5, 4, 8>("m8n8k4+8dfma", 8);
  run<5, 4, 8>("m8n8k4+8dfma", 16);
  run<5, 4, 16>("m8n8k4+16dfma", 8);
  return 0;
}
