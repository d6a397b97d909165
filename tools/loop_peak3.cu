// Experiment: per-step broadcast operands (A0, A1, cf) moved to UNIFORM registers with redux.sync, so
// that every DFMA of the hot loop reads at most two vector-register pairs.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double uni(double x) {   // x is warp-uniform: OR-reduce -> uniform datapath
  unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)__double2loint(x));
  unsigned hi = __reduce_or_sync(0xffffffffu, (unsigned)__double2hiint(x));
  return __hiloint2double((int)hi, (int)lo);
}
template <int VAR>
__global__ void __launch_bounds__(128, 4) k_loop(double* out, const double* in, int iters) {
  __shared__ double rec[32][4];
  if (threadIdx.x < 32) { rec[threadIdx.x][0] = in[threadIdx.x]; rec[threadIdx.x][1] = in[threadIdx.x + 1]; rec[threadIdx.x][2] = in[2]; rec[threadIdx.x][3] = 0; }
  __syncthreads();
  double acc[32];
#pragma unroll
  for (int i = 0; i < 32; i++) acc[i] = 0.0;
  double x = in[3 + threadIdx.x], xo = in[4], cd = in[5], sd = in[6];
#pragma unroll 2
  for (int it = 0; it < iters; it++) {
    double A0 = rec[it & 31][0], A1 = rec[it & 31][1], cf = rec[it & 31][2];
    if (VAR == 1) { A0 = uni(A0); A1 = uni(A1); cf = uni(cf); }
    if (VAR == 2) { cf = uni(cf); }
    double vm = x, v = fma(xo, sd, x * cd);
    acc[0] = fma(A0, vm, acc[0]); acc[1] = fma(A1, vm, acc[1]);
    acc[2] = fma(A0, v, acc[2]); acc[3] = fma(A1, v, acc[3]);
#pragma unroll
    for (int k = 2; k < 16; k++) {
      const double vn = fma(cf, v, -vm); vm = v; v = vn;
      acc[2 * k] = fma(A0, v, acc[2 * k]); acc[2 * k + 1] = fma(A1, v, acc[2 * k + 1]);
    }
    x = v * 0.999; xo = vm;
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
  run<0>(out, in, "smem broadcast (vector regs)"); run<1>(out, in, "A0,A1,cf via redux -> uniform"); run<2>(out, in, "cf only via redux");
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
