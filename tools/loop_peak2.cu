// Variants of the hot loop's accumulate ordering (register-operand reuse experiments).
#include <cstdio>
#include <cuda_runtime.h>
template <int VAR>
__global__ void __launch_bounds__(128, 4) k_loop(double* out, const double* in, int iters) {
  double acc[32];
#pragma unroll
  for (int i = 0; i < 32; i++) acc[i] = 0.0;
  double A0 = in[threadIdx.x], A1 = in[threadIdx.x + 1], cf = in[2], x = in[3 + threadIdx.x], xo = in[4], cd = in[5], sd = in[6];
#pragma unroll 2
  for (int it = 0; it < iters; it++) {
    double v[16];
    v[0] = x; v[1] = fma(xo, sd, x * cd);
    if (VAR == 0) {          // baseline: chain and accumulate interleaved per k
      acc[0] = fma(A0, v[0], acc[0]); acc[1] = fma(A1, v[0], acc[1]);
      acc[2] = fma(A0, v[1], acc[2]); acc[3] = fma(A1, v[1], acc[3]);
#pragma unroll
      for (int k = 2; k < 16; k++) { v[k] = fma(cf, v[k - 1], -v[k - 2]); acc[2 * k] = fma(A0, v[k], acc[2 * k]); acc[2 * k + 1] = fma(A1, v[k], acc[2 * k + 1]); }
    } else if (VAR == 1) {   // whole chain first, then component-major accumulation (A0 reused 16x, then A1)
#pragma unroll
      for (int k = 2; k < 16; k++) v[k] = fma(cf, v[k - 1], -v[k - 2]);
#pragma unroll
      for (int k = 0; k < 16; k++) acc[2 * k] = fma(A0, v[k], acc[2 * k]);
#pragma unroll
      for (int k = 0; k < 16; k++) acc[2 * k + 1] = fma(A1, v[k], acc[2 * k + 1]);
    } else {                 // snake
#pragma unroll
      for (int k = 2; k < 16; k++) v[k] = fma(cf, v[k - 1], -v[k - 2]);
#pragma unroll
      for (int k = 0; k < 16; k += 2) {
        acc[2 * k] = fma(A0, v[k], acc[2 * k]); acc[2 * k + 1] = fma(A1, v[k], acc[2 * k + 1]);
        acc[2 * k + 3] = fma(A1, v[k + 1], acc[2 * k + 3]); acc[2 * k + 2] = fma(A0, v[k + 1], acc[2 * k + 2]);
      }
    }
    x = v[15] * 0.999; xo = v[14];
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 32; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int VAR> void run(double* out, double* in, const char* name) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 100000, sms = 148; float ms;
  for (int r = 0; r < 2; r++) { cudaEventRecord(e0); k_loop<VAR><<<sms * 4, 128>>>(out, in, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); }
  cudaEventElapsedTime(&ms, e0, e1);
  printf("%s: %.2f ms  %.2f fp64 op/clk/SM\n", name, ms, (double)sms * 4 * 128 * 49.0 * iters / (ms * 1e-3) / sms / 1.965e9);
}
int main() {
  double *out, *in; cudaMalloc(&out, 8 * 148 * 8 * 128 * 4); cudaMalloc(&in, 8 * 1024); cudaMemset(in, 0, 8 * 1024);
  run<0>(out, in, "interleaved"); run<1>(out, in, "component-major"); run<2>(out, in, "snake");
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
