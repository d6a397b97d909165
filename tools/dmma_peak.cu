// FP64 tensor-core (DMMA) issue rate on this GPU: mma.sync m8n8k4 / m16n8k4 / m16n8k8 / m16n8k16 f64, independent accumulator chains.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dmma tools/dmma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma1688(double* d, const double* a, const double* b) {   // m16n8k8: A 4 regs, B 2, C 4
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void mma1684(double* d, const double* a, double b) {   // m16n8k4: A 2 regs, B 1, C 4
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void mma16816(double* d, const double* a, const double* b) {   // m16n8k16: A 8 regs, B 4, C 4
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
               : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
               : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int SHAPE>
__global__ void __launch_bounds__(128) k(double* out, const double* in, int iters) {
  double d[32];
#pragma unroll
  for (int i = 0; i < 32; i++) d[i] = 0;
  double a[8], b[4];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = in[threadIdx.x + i];
#pragma unroll
  for (int i = 0; i < 4; i++) b[i] = in[threadIdx.x + 8 + i];
  for (int it = 0; it < iters; it++) {
    if (SHAPE == 0) {
#pragma unroll
      for (int j = 0; j < 16; j++) mma884(d[2 * j], d[2 * j + 1], a[j & 7], b[j & 3]);       // 16 x 256 FMA
    } else if (SHAPE == 1) {
#pragma unroll
      for (int j = 0; j < 8; j++) mma1684(d + 4 * j, a + 2 * (j & 3), b[j & 3]);              // 8 x 512
    } else if (SHAPE == 2) {
#pragma unroll
      for (int j = 0; j < 8; j++) mma1688(d + 4 * j, a + 4 * (j & 1), b + 2 * (j & 1));       // 8 x 1024
    } else {
#pragma unroll
      for (int j = 0; j < 8; j++) mma16816(d + 4 * j, a, b);                                  // 8 x 2048
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 32; i++) s += d[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int SHAPE>
void run(const char* name, double fmaPerIter) {
  double *out, *in; cudaMalloc(&out, 8 * 148 * 16 * 128); cudaMalloc(&in, 8 * 1024); cudaMemset(in, 0, 8 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000;
  for (int bps = 1; bps <= 4; bps *= 2) {
    float ms = 0;
    for (int r = 0; r < 2; r++) { cudaEventRecord(e0); k<SHAPE><<<148 * bps, 128>>>(out, in, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); }
    const double fma = 148.0 * bps * 4 * fmaPerIter * iters;
    printf("%s warps/SM %2d: %8.3f ms  %7.2f fma/clk/SM @1.965GHz  = %.2f TFLOP/s\n", name, bps * 4, ms, fma / (ms * 1e-3) / 148 / 1.965e9, 2 * fma / (ms * 1e-3) / 1e12);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
int main() {
  run<0>("m8n8k4  ", 16 * 256.0);
  run<1>("m16n8k4 ", 8 * 512.0);
  run<2>("m16n8k8 ", 8 * 1024.0);
  run<3>("m16n8k16", 8 * 2048.0);
}
