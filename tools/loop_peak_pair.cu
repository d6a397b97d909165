// Attainable rate of the REAL pair-kernel main phase (srb::main_pair, all-pass path) in isolation:
// the per-warp staging area is filled once, then every warp replays the 32-step sub-batch `reps` times.
// Answers: how much of the FP64 pipe can the main phase use at 4/8/12/16 warps per SM when no warp is
// ever in the prep phase?  (DFMA form of the fp64 loop, before the DMMA rewrite: 44.6 op/clk/SM = 70 % of the
// 63.6 DFMA peak at any occupancy -- register-operand bound.)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I synchrad_b200/csrc -o /tmp/lpp tools/loop_peak_pair.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "srb_literal.cuh"

using namespace srb;

template <class C, int NW, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB) k_main(double* out, int reps) {
  extern __shared__ __align__(16) unsigned char raw[];
  WarpSmem<C>* smAll = reinterpret_cast<WarpSmem<C>*>(raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WarpSmem<C>& sm = smAll[warp];
  for (int k = 0; k < C::NREC; k++) sm.rec[lane][k] = (typename C::TM)(1e-3 * (k + 1) + 1e-5 * lane);
  for (int k = 0; k < C::NSEED; k++) sm.seeds[k][lane] = (typename C::TM)(0.5 + 1e-3 * k + 1e-5 * lane);
  sm.rng[lane] = 0;
  __syncwarp();
  ThreadState<C> st;
#pragma unroll
  for (int i = 0; i < C::NACC; i++) st.acc[i] = 0;
  Params P{}; Geom g{};
  for (int r = 0; r < reps; r++) {
    if constexpr (C::MMA) main_pair_mma<C>(P, g, sm, 32, 0xffffffffu, 0xffffffffu, lane, st);
    else main_pair<C>(P, g, sm, 32, 0xffffffffu, 0xffffffffu, lane, st);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < C::NACC; i++) s += (double)st.acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class C, int MINB>
void run(const char* name, double opsPerLaneStep) {
  int sms = 148; double* out;
  cudaMalloc(&out, 8 * 148 * 8 * 128 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int reps = 4000;
  const size_t smem = 4 * sizeof(WarpSmem<C>);
  cudaFuncSetAttribute(k_main<C, 4, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int bps = 1; bps <= MINB; bps++) {
    float ms = 0;
    for (int r = 0; r < 2; r++) {
      cudaEventRecord(e0); k_main<C, 4, MINB><<<sms * bps, 128, smem>>>(out, reps); cudaEventRecord(e1); cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
    }
    const double ops = (double)sms * bps * 128 * opsPerLaneStep * 32.0 * reps;
    printf("%s warps/SM %2d: %8.2f ms  %6.2f op/clk/SM @1.965GHz  (%.3e node updates/s)\n", name, bps * 4, ms,
           ops / (ms * 1e-3) / sms / 1.965e9, (double)sms * bps * 128 * C::TW * 32.0 * reps / (ms * 1e-3));
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  // fp64: tensor-core main phase (DMMA); "op" = FMA-equivalents per lane and step (32 in DMMA + 4 for X)
  run<Cfg<double, double, MODE_FAR, KIND_PAIR, 8, false, 2>, 3>("fp64 pair TW8 DMMA (<=168 regs)", 4 + 32);
  run<Cfg<double, double, MODE_FAR, KIND_PAIR, 8, false, 2>, 4>("fp64 pair TW8 DMMA (<=128 regs)", 4 + 32);
  run<Cfg<double, double, MODE_FAR, KIND_PAIR, 8, false, 3>, 3>("fp64 pair TW8 NC3 DMMA", 4 + 48);
  run<Cfg<double, float, MODE_FAR, KIND_PAIR, 8, false, 2>, 4>("fp32 pair TW8", 4 + 32);
  run<Cfg<double, float, MODE_FAR, KIND_PAIR, 16, false, 2>, 4>("fp32 pair TW16", 4 + 64);
  return 0;
}
