// Do DMMA.8x8x4 and DFMA share the FP64 units on this GPU?  Warps 0..1 of every sub-partition pair run DMMA chains,
// the others DFMA chains; compare the mixed run with each half alone.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o variants/overlap tools/dmma_dfma_overlap.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// mode bit 0: DMMA warps active, bit 1: DFMA warps active.  256 threads = 8 warps: warps 0-3 DMMA, 4-7 DFMA
// (one of each kind per sub-partition).
__global__ void __launch_bounds__(256) k(double* out, const double* in, int itersM, int itersF, int mode) {
  const int warp = threadIdx.x >> 5;
  double d[16], s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) d[i] = 0;
  const double a = in[threadIdx.x], b = in[threadIdx.x + 1];
  if (warp < 4) {
    if (mode & 1)
      for (int it = 0; it < itersM; it++) {
#pragma unroll
        for (int j = 0; j < 8; j++) mma884(d[2 * j], d[2 * j + 1], a, b);   // 8 x 256 FMA
      }
  } else {
    if (mode & 2)
      for (int it = 0; it < itersF; it++) {
#pragma unroll
        for (int j = 0; j < 16; j++) d[j] = fma(a, d[j], b);                // 16 x 32 FMA, two-operand-reuse form
      }
  }
#pragma unroll
  for (int i = 0; i < 16; i++) s += d[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  double *out, *in; cudaMalloc(&out, 8 * 148 * 256 * 4); cudaMalloc(&in, 8 * 1024); cudaMemset(in, 0, 8 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int itersM = 40000;          // per DMMA warp: 40000 x 8 x 256 FMA  = 8.19e7 FMA -> x16 cycles/DMMA = 5.12e6 cycles alone
  const int itersF = 160000;         // per DFMA warp: 160000 x 16 x 32 FMA = 8.19e7 FMA -> x2 cycles/DFMA  = 5.12e6 cycles alone
  for (int mode = 1; mode <= 3; mode++) {
    float ms = 0;
    for (int r = 0; r < 2; r++) { cudaEventRecord(e0); k<<<148, 256>>>(out, in, itersM, itersF, mode); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); }
    const double fma = 148.0 * 4 * (((mode & 1) ? (double)itersM * 8 * 256 : 0) + ((mode & 2) ? (double)itersF * 16 * 32 : 0));
    printf("mode %d (%s): %8.3f ms  %6.2f FMA/clk/SM\n", mode, mode == 1 ? "DMMA only" : mode == 2 ? "DFMA only" : "both", ms, fma / (ms * 1e-3) / 148 / 1.965e9);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
