// Attainable FP64-pipe rate of the hot loop's instruction mix in isolation (no smem, no prep):
// per iteration 14 chain DFMAs + 32 accumulate DFMAs + 2 seed ops, NC=2, TW=16.
#include <cstdio>
#include <cuda_runtime.h>
template <int UNROLL>
__global__ void __launch_bounds__(128, 4) k_loop(double* out, const double* in, int iters) {
  double acc[32];
#pragma unroll
  for (int i = 0; i < 32; i++) acc[i] = 0.0;
  double A0 = in[threadIdx.x], A1 = in[threadIdx.x + 1], cf = in[2], x = in[3 + threadIdx.x], xo = in[4], cd = in[5], sd = in[6];
#pragma unroll UNROLL
  for (int it = 0; it < iters; it++) {
    double vm = x;
    double v = fma(xo, sd, x * cd);
    acc[0] = fma(A0, vm, acc[0]); acc[1] = fma(A1, vm, acc[1]);
    acc[2] = fma(A0, v, acc[2]); acc[3] = fma(A1, v, acc[3]);
#pragma unroll
    for (int k = 2; k < 16; k++) {
      const double vn = fma(cf, v, -vm); vm = v; v = vn;
      acc[2 * k] = fma(A0, v, acc[2 * k]); acc[2 * k + 1] = fma(A1, v, acc[2 * k + 1]);
    }
    x = v * 0.999; xo = vm;   // next iteration's operands depend on data (keeps the chain honest)
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 32; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  int sms = 148; double *out, *in;
  cudaMalloc(&out, 8 * 148 * 8 * 128 * 4); cudaMalloc(&in, 8 * 1024); cudaMemset(in, 0, 8 * 1024);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 100000;
  for (int bps = 1; bps <= 4; bps++) {
    float ms;
    for (int r = 0; r < 2; r++) { cudaEventRecord(e0); k_loop<2><<<sms * bps, 128>>>(out, in, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); }
    cudaEventElapsedTime(&ms, e0, e1);
    double dfma = (double)sms * bps * 128 * 49.0 * iters;
    printf("unroll2 warps/SM %2d: %.2f ms  %.2f fp64 op/clk/SM @1.965GHz\n", bps * 4, ms, dfma / (ms * 1e-3) / sms / 1.965e9);
    for (int r = 0; r < 2; r++) { cudaEventRecord(e0); k_loop<1><<<sms * bps, 128>>>(out, in, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); }
    cudaEventElapsedTime(&ms, e0, e1);
    printf("unroll1 warps/SM %2d: %.2f ms  %.2f fp64 op/clk/SM\n", bps * 4, ms, dfma / (ms * 1e-3) / sms / 1.965e9);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
