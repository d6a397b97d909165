// How the FP64 pipe of an SM sub-partition is shared between ONE warp issuing DMMA.8x8x4 and other warps issuing
// dependent DFMA chains (the consumer / producer mix of srb_ws.cuh).  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ws_pipe_mix tools/ws_pipe_mix.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// warps [0, nDmma) per block: 16 independent DMMA accumulators; the others: ILP independent DFMA chains
template <int ILP>
__global__ void __launch_bounds__(512) k(double* out, const double* in, int nDmmaWarps, int itersDmma, int itersDfma) {
  const int warp = threadIdx.x >> 5;
  double s = 0;
  if (warp < nDmmaWarps) {
    double d[32];
#pragma unroll
    for (int i = 0; i < 32; i++) d[i] = 0;
    double a[8], b[2];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = in[threadIdx.x + i];
    b[0] = in[threadIdx.x + 8]; b[1] = in[threadIdx.x + 9];
    for (int it = 0; it < itersDmma; it++) {
#pragma unroll
      for (int j = 0; j < 16; j++) mma884(d[2 * j], d[2 * j + 1], a[j >> 1], b[j & 1]);
    }
#pragma unroll
    for (int i = 0; i < 32; i++) s += d[i];
  } else {
    double v[ILP];
    const double c = in[threadIdx.x], e = in[threadIdx.x + 1];
#pragma unroll
    for (int i = 0; i < ILP; i++) v[i] = in[threadIdx.x + 2 + i];
    for (int it = 0; it < itersDfma; it++) {
#pragma unroll
      for (int r = 0; r < 16; r++)
#pragma unroll
        for (int i = 0; i < ILP; i++) v[i] = fma(v[i], c, e);
    }
#pragma unroll
    for (int i = 0; i < ILP; i++) s += v[i];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
float run(int warps, int nDmma, int itD, int itF, double* out, double* in) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms = 0;
  for (int r = 0; r < 2; r++) { cudaEventRecord(e0); k<ILP><<<148, warps * 32>>>(out, in, nDmma, itD, itF); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); }
  return ms;
}

int main() {
  double *out, *in; cudaMalloc(&out, 8 * 148 * 512); cudaMalloc(&in, 8 * 1024); cudaMemset(in, 0, 8 * 1024);
  const double clk = 1.965e9;
  const int itD = 20000;
  // DMMA alone: 1, 2, 3, 4 warps per sub-partition
  for (int w = 4; w <= 16; w += 4) {
    const float ms = run<1>(w, w, itD, 0, out, in);
    printf("DMMA only, %2d warps/SM: %8.3f ms  %6.2f fma/clk/SM  (cycles per DMMA per SMSP-warp: %.1f)\n", w, ms,
           (double)w * itD * 16 * 256 / (ms * 1e-3) / clk, ms * 1e-3 * clk / (itD * 16.0));
  }
  // DFMA chains alone: 8 warps/SM (2 per sub-partition), ILP 1 / 2 / 4
  const int itF = 20000;
  { float ms = run<1>(8, 0, 0, itF, out, in); printf("DFMA only, 8 warps ILP1: %8.3f ms  cycles per dependent DFMA: %.1f\n", ms, ms * 1e-3 * clk / (itF * 16.0)); }
  { float ms = run<2>(8, 0, 0, itF, out, in); printf("DFMA only, 8 warps ILP2: %8.3f ms  cycles per DFMA per warp: %.1f\n", ms, ms * 1e-3 * clk / (itF * 32.0)); }
  { float ms = run<4>(8, 0, 0, itF, out, in); printf("DFMA only, 8 warps ILP4: %8.3f ms  cycles per DFMA per warp: %.1f\n", ms, ms * 1e-3 * clk / (itF * 64.0)); }
  // mix: 4 DMMA warps (1 per sub-partition) + 8 DFMA warps; DFMA work sized to ~ the producer share (245 ops per 128 DMMA)
  for (int ilp = 1; ilp <= 4; ilp *= 2) {
    for (int share = 1; share <= 2; share++) {
      // per DMMA warp 20000*16 DMMA; per DFMA warp: share * (245/128)/2 ops per DMMA
      const int nF = (int)(itD * 16.0 * share * 245.0 / 128.0 / 2.0 / (16.0 * ilp));
      float ms = ilp == 1 ? run<1>(12, 4, itD, nF, out, in) : ilp == 2 ? run<2>(12, 4, itD, nF, out, in) : run<4>(12, 4, itD, nF, out, in);
      const double pipe = (itD * 16.0 * 16 + 2.0 * nF * 16.0 * ilp * 2) / (ms * 1e-3 * clk);
      printf("mix 4 DMMA + 8 DFMA warps, ILP%d, DFMA ops per 128 DMMA = %3d: %8.3f ms  (DMMA alone would take %.3f ms)  FP64 pipe busy %.2f\n",
             ilp, share * 245, ms, itD * 16.0 * 16 / clk * 1e3, pipe);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
