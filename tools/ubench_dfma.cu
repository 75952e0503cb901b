// Microbenchmark (design evidence for the structured path): DFMA throughput when one
// operand streams from (A) the constant bank as an instruction operand, (B) shared memory via
// broadcast LDS.128, (C) registers; as a function of the table size and of whether the warps of
// a CTA walk the same table slice or different slices.
#include <cstdio>
#include <cuda_runtime.h>
#define TSMAX 4096
__constant__ double c_tab[TSMAX];

template <int TS, int NW, bool SLICE, int MODE>
__global__ void __launch_bounds__(NW * 32) k(double *out, const double *gtab, int reps) {
  extern __shared__ double s_tab[];
  if (MODE == 1) {
    for (int i = threadIdx.x; i < TS; i += blockDim.x) s_tab[i] = gtab[i];
    __syncthreads();
  }
  const int w = threadIdx.x >> 5;
  double s[4], acc[16];
  for (int i = 0; i < 4; ++i) s[i] = 1.0 + 1e-9 * (threadIdx.x + i);
  for (int i = 0; i < 16; ++i) acc[i] = i;
  constexpr int LEN = SLICE ? TS / NW : TS;
  for (int r = 0; r < reps; ++r) {
    if (SLICE) {
      // each warp owns a different, compile-time-known slice -> different straight-line code
#define BODY(W)                                                                            \
  case W: {                                                                                \
    _Pragma("unroll") for (int i = 0; i < LEN; i += 16) {                                  \
      _Pragma("unroll") for (int k2 = 0; k2 < 16; ++k2) {                                  \
        const int idx = (W % NW) * LEN + i + k2;                                           \
        double t;                                                                          \
        if (MODE == 0) t = c_tab[idx];                                                     \
        else if (MODE == 1) t = s_tab[idx];                                                \
        else t = s[(k2 + 1) & 3];                                                          \
        acc[k2] = fma(s[k2 & 3], t, acc[k2]);                                              \
      }                                                                                    \
    }                                                                                      \
  } break;
      switch (w) { BODY(0) BODY(1) BODY(2) BODY(3) BODY(4) BODY(5) BODY(6) BODY(7) }
    } else {
#pragma unroll
      for (int i = 0; i < LEN; i += 16) {
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2) {
          const int idx = i + k2;
          double t;
          if (MODE == 0) t = c_tab[idx];
          else if (MODE == 1) t = s_tab[idx];
          else t = s[(k2 + 1) & 3];
          acc[k2] = fma(s[k2 & 3], t, acc[k2]);
        }
      }
    }
  }
  double v = 0;
  for (int i = 0; i < 16; ++i) v += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = v;
}

template <int TS, int NW, bool SLICE, int MODE>
void run(const char *name, int ctas_per_sm) {
  double *out, *gtab;
  const int nsm = 148;
  cudaMalloc(&out, sizeof(double) * nsm * ctas_per_sm * NW * 32);
  cudaMalloc(&gtab, sizeof(double) * TSMAX);
  cudaMemset(gtab, 0, sizeof(double) * TSMAX);
  const int reps = SLICE ? 4000 : 500;
  cudaFuncSetAttribute(k<TS, NW, SLICE, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, TS * 8);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  for (int it = 0; it < 2; ++it) {
    cudaEventRecord(a);
    k<TS, NW, SLICE, MODE><<<nsm * ctas_per_sm, NW * 32, MODE == 1 ? TS * 8 : 0>>>(out, gtab, reps);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
  }
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double fma_per_thread = (double)reps * (SLICE ? TS / NW : TS);
  const double total = fma_per_thread * NW * 32 * nsm * ctas_per_sm;
  printf("%-10s TS=%5d NW=%d slice=%d ctas/sm=%d : %8.3f ms  %7.2f TDFMA/s  (%s)\n", name, TS, NW, (int)SLICE, ctas_per_sm, ms,
         total / ms * 1e-9, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(gtab);
}

int main() {
  double h[TSMAX];
  for (int i = 0; i < TSMAX; ++i) h[i] = 1e-3 * i;
  cudaMemcpyToSymbol(c_tab, h, sizeof h);
  run<256, 8, false, 2>("regs", 1);
  run<256, 8, false, 2>("regs", 2);
  run<256, 8, false, 0>("const", 1);
  run<1024, 8, false, 0>("const", 1);
  run<2304, 8, false, 0>("const", 1);
  run<4096, 8, false, 0>("const", 1);
  run<2304, 8, false, 0>("const", 2);
  run<1024, 8, true, 0>("const", 1);
  run<2304, 8, true, 0>("const", 1);
  run<4096, 8, true, 0>("const", 1);
  run<2304, 8, true, 0>("const", 2);
  run<4096, 8, true, 0>("const", 2);
  run<2304, 8, false, 1>("smem", 1);
  run<2304, 8, true, 1>("smem", 1);
  run<2304, 8, true, 1>("smem", 2);
  return 0;
}
