// Micro-benchmarks pinning the issue-limited pipe peaks the roofline uses
// (SURVEY.md §8d): FP64 DFMA, FP32 FFMA, MUFU (sin.approx) per SM per clock on B200.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipe_peaks tools/pipe_peaks.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CH>
__global__ void k_dfma(double* out, double a, double b, int iters) {
  double v[CH];
#pragma unroll
  for (int i = 0; i < CH; i++) v[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CH; i++) v[i] = fma(v[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; i++) s += v[i];
  if (s == 123.456) out[0] = s;
}
template <int CH>
__global__ void k_ffma(float* out, float a, float b, int iters) {
  float v[CH];
#pragma unroll
  for (int i = 0; i < CH; i++) v[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CH; i++) v[i] = fmaf(v[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < CH; i++) s += v[i];
  if (s == 123.456f) out[0] = s;
}
template <int CH>
__global__ void k_mufu(float* out, float a, int iters) {
  float v[CH];
#pragma unroll
  for (int i = 0; i < CH; i++) v[i] = threadIdx.x * 1e-3f + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CH; i++) v[i] = __sinf(v[i]) ;
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < CH; i++) s += v[i];
  if (s == 123.456f) out[0] = s + a;
}
// DFMA with a concurrent stream of integer/ALU work: shows whether non-FP64 issue is hidden
template <int CH>
__global__ void k_dfma_mix(double* out, double a, double b, int iters) {
  double v[CH]; unsigned u[CH];
#pragma unroll
  for (int i = 0; i < CH; i++) { v[i] = threadIdx.x * 1e-3 + i; u[i] = threadIdx.x + i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CH; i++) { v[i] = fma(v[i], a, b); u[i] = (u[i] ^ (u[i] >> 3)) + it; }
  }
  double s = 0; unsigned t = 0;
#pragma unroll
  for (int i = 0; i < CH; i++) { s += v[i]; t += u[i]; }
  if (s == 123.456 || t == 77) out[0] = s;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  printf("device %s sms %d clock_khz %d\n", p.name, sms, p.clockRate);
  double* dout; cudaMalloc(&dout, 64);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int CH = 8; const int iters = 20000;
  for (int wps = 4; wps <= 32; wps *= 2) {   // warps per SM
    int threads = 32 * wps > 1024 ? 1024 : 32 * wps; int blocks = sms * ((32 * wps) / threads);
    float ms;
    for (int rep = 0; rep < 2; rep++) { cudaEventRecord(e0); k_dfma<CH><<<blocks, threads>>>(dout, 1.0000001, 1e-9, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); }
    cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)blocks * threads * CH * iters;
    printf("DFMA  warps/SM %2d: %.3f ms  %.2f Gfma/s  = %.2f fma/clk/SM @1.965GHz\n", wps, ms, ops / ms * 1e-6, ops / (ms * 1e-3) / sms / 1.965e9);
    for (int rep = 0; rep < 2; rep++) { cudaEventRecord(e0); k_dfma_mix<CH><<<blocks, threads>>>(dout, 1.0000001, 1e-9, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); }
    cudaEventElapsedTime(&ms, e0, e1);
    printf("DFMA+ALU warps/SM %2d: %.3f ms  %.2f Gfma/s  = %.2f fma/clk/SM\n", wps, ms, ops / ms * 1e-6, ops / (ms * 1e-3) / sms / 1.965e9);
    for (int rep = 0; rep < 2; rep++) { cudaEventRecord(e0); k_ffma<CH><<<blocks, threads>>>((float*)dout, 1.0000001f, 1e-9f, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); }
    cudaEventElapsedTime(&ms, e0, e1);
    printf("FFMA  warps/SM %2d: %.3f ms  %.2f Gfma/s  = %.2f fma/clk/SM\n", wps, ms, ops / ms * 1e-6, ops / (ms * 1e-3) / sms / 1.965e9);
    for (int rep = 0; rep < 2; rep++) { cudaEventRecord(e0); k_mufu<CH><<<blocks, threads>>>((float*)dout, 1.0f, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); }
    cudaEventElapsedTime(&ms, e0, e1);
    printf("MUFU  warps/SM %2d: %.3f ms  %.2f Gop/s  = %.2f op/clk/SM\n", wps, ms, ops / ms * 1e-6, ops / (ms * 1e-3) / sms / 1.965e9);
  }
  // sustained DFMA for ~3 s to see the clock under power cap
  {
    int threads = 512, blocks = sms * 2; float ms;
    cudaEventRecord(e0);
    for (int r = 0; r < 40; r++) k_dfma<CH><<<blocks, threads>>>(dout, 1.0000001, 1e-9, iters * 4);
    cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    double ops = 40.0 * blocks * threads * CH * iters * 4;
    printf("DFMA sustained: %.1f ms  %.2f Gfma/s = %.2f fma/clk/SM @1.965GHz\n", ms, ops / ms * 1e-6, ops / (ms * 1e-3) / sms / 1.965e9);
  }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
