// synchrad_b200 — __global__ entry points and the C ABI of include/synchrad_b200.h.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/synchrad_b200.h"
#include "srb_literal.cuh"
#include "srb_ws.cuh"

namespace {

thread_local std::string g_err;
thread_local srb_launch_info g_info;

int fail(const std::string& m) { g_err = m; return -1; }
#define SRB_CUDA(call)                                                                  \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess)                                                              \
      return fail(std::string(#call) + ": " + cudaGetErrorString(e_));                  \
  } while (0)

// Block shape (measured on B200, profiles/r01_tuning.md): 4 warps (= 4 virtual directions) per
// block; the recurrence kernels run 4 blocks/SM (16 warps, <=128 registers), the direct kernels
// (48 accumulators + sincos temporaries) 2 blocks/SM.
#ifndef SRB_NW
#define SRB_NW 4
#endif
#ifndef SRB_MINB
#define SRB_MINB 4
#endif
#ifndef SRB_MINB_MMA
#define SRB_MINB_MMA 3
#endif
#ifndef SRB_MINB_DIRECT
#define SRB_MINB_DIRECT 2
#endif
constexpr int NW = SRB_NW;   // warps (= virtual directions) per block

// resident blocks per SM the register budget is tuned for: accumulators must stay in registers
template <class C> constexpr int min_blocks() {
  constexpr int accRegs = C::NACC * (int)sizeof(typename C::TM) / 4;
  if (C::KIND == srb::KIND_DREC) return C::NACC <= 24 ? 3 : 2;       // 17 KB of seeds per warp; 48-96 accumulator registers
  if (C::KIND == srb::KIND_DIRECT || C::KIND == srb::KIND_LITERAL) return accRegs <= 48 ? 4 : SRB_MINB_DIRECT;   // fp32 direct: 48 accumulator registers
  if (C::MMA) return C::NACC > 48 ? 2 : SRB_MINB_MMA;   // 16-node tiles: 64 fp64 accumulators per lane
  if (C::PAIR && sizeof(typename C::TM) == 8) return SRB_MINB > 3 ? 3 : SRB_MINB;   // measured: 168 regs beat 128
  return accRegs <= 64 ? SRB_MINB : (SRB_MINB > 3 ? 3 : SRB_MINB);
}

template <class C>
__global__ void __launch_bounds__(NW * 32, min_blocks<C>()) k_integrate(const srb::Params P) {
  extern __shared__ __align__(128) unsigned char smraw[];
  const uint32_t warp = threadIdx.x >> 5;
  srb::WarpSmem<C>* sm = reinterpret_cast<srb::WarpSmem<C>*>(smraw) + warp;
  const uint32_t nVDtiles = (P.nVD + NW - 1) / NW;
  const uint32_t vd = (blockIdx.x % nVDtiles) * NW + warp;   // neighbouring blocks share the tracks
  const uint32_t pc = blockIdx.x / nVDtiles;
  if (vd >= P.nVD || srb::deselected(P)) return;
  srb::ThreadState<C> st;
  srb::warp_task<C>(P, vd, pc, *sm, &st);
}

// ------------------------------------------------------------------------------------------- warp-specialised pair kernel
// (srb_ws.cuh: 1 DMMA consumer warp + NP producer warps per virtual direction, NU directions per block, mbarrier ring,
//  TMA-staged inputs)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\nbra WAIT_LOOP;\nWAIT_DONE:\n}" ::"r"(smem_u32(b)), "r"(parity) : "memory");
}
// TMA bulk copy global -> shared, completion counted in bytes on an mbarrier (both addresses and the size 16-byte aligned)
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <class C, int NU, int NP, int NS>
struct alignas(128) WsSmem {
  srb::WsStage<C> stage[NU][NS];
  double in[NU][NP][2][srb::WS_IN][srb::WS_NIN];     // TMA-staged input records of a producer's next (pair of) sub-batches
  uint64_t full[NU][NS], empty[NU][NS], inbar[NU][NP][2], xfer[NU][NS];
};

// NCW consumer warps per unit.  With two, the sub-batches alternate between them (a split of the GEMM's K dimension:
// each holds a full accumulator set) so that one warp's LDS / rotation / bookkeeping instructions issue under the
// other's 16-cycle DMMA issue stalls; at a flush the second warp hands its sums to the first through the stage.
template <class C, int NU, int NP, int NS, int NCW>
__global__ void __launch_bounds__((NCW + NP) * NU * 32, 1) k_integrate_ws(const srb::Params P) {
  using namespace srb;
  extern __shared__ __align__(128) unsigned char smraw[];
  WsSmem<C, NU, NP, NS>& S = *reinterpret_cast<WsSmem<C, NU, NP, NS>*>(smraw);
  static_assert(sizeof(WsStage<C>) >= 32 * C::NACC * sizeof(double), "stage too small for the flush transposes");
  static_assert(sizeof(WsStage<C>) % 16 == 0, "stages keep the input rows 16-byte aligned");
  // a producer must observe every phase of the empty barriers it waits on (parity waits): with stage = k % NS and
  // producer = k % NP that holds iff each stage is always served by the same producer
  static_assert(NS % NP == 0, "ring stages must map to a fixed producer");
  const int warp = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31u);
  if (deselected(P)) return;           // (uniform over the block: before any barrier)
  // masked steps of the tensor-core stream multiply stale stage data by 0: keep it finite from the start
  for (uint32_t k = threadIdx.x; k < sizeof(S.stage) / 4u; k += blockDim.x) reinterpret_cast<uint32_t*>(&S.stage)[k] = 0u;
  if (threadIdx.x == 0) {
    for (int u = 0; u < NU; u++) {
      for (int i = 0; i < NS; i++) { mbar_init(&S.full[u][i], 32); mbar_init(&S.empty[u][i], 32); mbar_init(&S.xfer[u][i], 32); }
      for (int q = 0; q < NP; q++) { mbar_init(&S.inbar[u][q][0], 1); mbar_init(&S.inbar[u][q][1], 1); }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const bool consumer = warp < NCW * NU;
  const int u = warp % NU;                                       // (warp w sits on sub-partition w % 4: one unit each)
  const int cidx = consumer ? warp / NU : 0;
  const int pidx = consumer ? 0 : (warp - NCW * NU) / NU;
  const uint32_t nVDtiles = (P.nVD + NU - 1) / NU;
  const uint32_t vd = (blockIdx.x % nVDtiles) * NU + (uint32_t)u;   // neighbouring blocks share the tracks
  const uint32_t pc = blockIdx.x / nVDtiles;
  if (vd >= P.nVD) return;
  Geom g;
  make_geom<C>(P, vd, g);
  uint32_t t0, t1;
  chunk_tracks(P, pc, t0, t1);
  Walk<C> w;
  w.init(t0, t1, pc);
  WalkItem it;
  uint32_t k = 0;

  if (consumer) {
    // Ring slot k lives in stage k % NS and belongs to consumer warp k % NCW.  With two consumer warps (NS even) the
    // even and the odd slots are two independent pipelines -- each warp waits only on the barriers of its own stages,
    // so it observes every one of their phases (a parity wait means nothing to a waiter that skips phases: an earlier
    // version let both warps wait on the flush slot and a warp that had drifted ahead took an older phase of the same
    // parity for it).  The pipelines meet at a flush, which therefore takes TWO consecutive slots, one in each:
    //   slot kA -> the "giver" (warp kA % 2): dumps its accumulators into stage A, restarts from zero;
    //   slot kB = kA + 1 -> the "taker": adds the dump (xfer barrier of stage A, phase tracked by counting flushes per
    //   stage), releases stage A, flushes through stage B and carries the cumulative amplitude from there on.
    static_assert(NCW == 1 || (NCW == 2 && NS % 2 == 0), "one or two consumer warps");
    ThreadState<C> st;
    uint32_t xph = 0u;                 // per-stage phase bits of the hand-over barriers
    auto flush_through = [&](double* buf, uint32_t iSnap) {
      // fragment layout -> one tile per lane (through the handed-over stage), common flush code, and back (snapshots
      // are cumulative)
      ws_store_frag<C>(buf, lane, st);
      __syncwarp();
      ws_load_tile<C>(buf, lane, st);
      flush_lane<C>(P, g, w.tv, pc, iSnap, lane, &st);
      ws_store_tile<C>(buf, lane, st);
      __syncwarp();
      ws_load_frag<C>(buf, lane, st);
      __syncwarp();
    };
    while (w.next(P, it)) {
      if (it.newTrack) {
#pragma unroll
        for (int q = 0; q < C::NACC; q++) st.acc[q] = 0.0;
      }
      const int sgi = (int)(k % (uint32_t)NS);
      const uint32_t par = (k / (uint32_t)NS) & 1u;
      if (it.kind == 0) {
        if ((int)(k % (uint32_t)NCW) == cidx) {
          mbar_wait(&S.full[u][sgi], par);
          ws_main<C>(P, g, S.stage[u][sgi], lane, st);
          mbar_arrive(&S.empty[u][sgi]);
        }
        k++;
      } else if (NCW == 1) {
        mbar_wait(&S.full[u][sgi], par);
        flush_through(reinterpret_cast<double*>(&S.stage[u][sgi]), it.iSnap);
        mbar_arrive(&S.empty[u][sgi]);
        k++;
      } else {
        const int sgB = (int)((k + 1u) % (uint32_t)NS);
        double* bufA = reinterpret_cast<double*>(&S.stage[u][sgi]);
        if ((int)(k % 2u) == cidx) {               // giver
          mbar_wait(&S.full[u][sgi], par);
#pragma unroll
          for (int q = 0; q < C::NACC; q++) { bufA[q * 32 + lane] = st.acc[q]; st.acc[q] = 0.0; }
          mbar_arrive(&S.xfer[u][sgi]);
        } else {                                   // taker
          mbar_wait(&S.full[u][sgB], ((k + 1u) / (uint32_t)NS) & 1u);
          mbar_wait(&S.xfer[u][sgi], (xph >> sgi) & 1u);
#pragma unroll
          for (int q = 0; q < C::NACC; q++) st.acc[q] += bufA[q * 32 + lane];
          mbar_arrive(&S.empty[u][sgi]);
          flush_through(reinterpret_cast<double*>(&S.stage[u][sgB]), it.iSnap);
          mbar_arrive(&S.empty[u][sgB]);
        }
        xph ^= 1u << sgi;
        k += 2u;
      }
    }
    return;
  }

  // ---- producer pidx of unit u.
  // PAIRS (NS == 2 * NP): a producer owns the ring slots 2m, 2m + 1 with m % NP == pidx -- always the same two stages --
  // and prepares the two sub-batches TOGETHER: the scalar chains of a step (two sincos, powers, the tile recurrence)
  // are long and share the FP64 pipe with the DMMA stream (~26 cycles per dependent op under that load, against 9
  // alone: tools/ws_pipe_mix.cu); two independent sub-batches in one instruction stream double the instruction-level
  // parallelism of the warp.  Otherwise slot k belongs to producer k % NP.
#ifndef SRB_WS_PAIRS
#define SRB_WS_PAIRS 0     // measured: 2.69e12 with pairs against 2.93e12 without (the ring loses depth), profiles/r02_ws_tuning.md
#endif
  constexpr bool PAIRS = SRB_WS_PAIRS && NS == 2 * NP;
  struct Own { uint32_t kind, base, k, itStart; int cnt; uint64_t idx; bool valid; };   // idx: element of step `base` in the arrays
  bool flush2 = false;               // two consumer warps: a flush takes two consecutive ring slots (see above)
  auto mine = [&](uint32_t kk) -> bool { return (PAIRS ? (kk >> 1) : kk) % (uint32_t)NP == (uint32_t)pidx; };
  auto next_own = [&](Own& o) {
    for (;;) {
      if (flush2) {
        flush2 = false;
        const uint32_t kk = k++;
        if (mine(kk)) { o.kind = 1u; o.base = 0u; o.cnt = 0; o.k = kk; o.idx = 0; o.valid = true; return; }
      }
      if (!w.next(P, it)) { o.valid = false; return; }
      const uint32_t kk = k++;
      if (it.kind == 1u && NCW > 1) flush2 = true;
      if (mine(kk)) {
        o.kind = it.kind; o.base = it.base; o.cnt = it.cnt; o.k = kk; o.itStart = w.tv.itStart;
        o.idx = (uint64_t)((const double*)w.tv.x - (const double*)P.x) + it.base;
        o.valid = true;
        return;
      }
    }
  };
  auto next_pair = [&](Own& a, Own& b) {          // b: the odd slot of a's pair, if there is one
    next_own(a);
    b.valid = false;
    if (PAIRS && a.valid && (a.k & 1u) == 0u) next_own(b);
  };
  const uint64_t totalSteps = P.offsets[P.nTracks];
  // Staged inputs of the sub-batch whose first step is element idx: the 34 packed records (k_prepass: x, y, z, a, b per
  // step, 72 bytes) from the even index at or below idx - 1 (previous step + 16-byte alignment), ONE bulk copy.
  // Sub-batches at the edges of the arrays use plain loads.
  auto can_tma = [&](const Own& o) -> bool { return P.tmaOK && o.idx >= 2 && o.idx + 33 <= totalSteps; };
  auto issue = [&](const Own& o, int slot) {
    if (lane == 0 && o.valid && o.kind == 0u && can_tma(o)) {
      uint64_t* bar = &S.inbar[u][pidx][slot];
      const uint64_t e0 = (o.idx - 1) & ~(uint64_t)1;
      mbar_arrive_expect_tx(bar, (uint32_t)(WS_NIN * WS_IN * 8));
      tma_load_1d(&S.in[u][pidx][slot][0][0], P.pre + e0 * WS_NIN, (uint32_t)(WS_NIN * WS_IN * 8), bar);
    }
  };
  WsConst kc;
  ws_const<C>(P, g, kc);
  unsigned long long nPass = 0, nAll = 0;
  uint32_t inphase = 0u;              // bit `slot`: phase of that input buffer's barrier
  // inputs of sub-batch o (staged in `slot` or, at the edges of the arrays, plain loads) -> guard + amplitude
  auto guard = [&](const Own& o, int slot, WsStep& ws) {
    const uint32_t itl = o.base + (uint32_t)lane;
    const bool active = lane < o.cnt;
    double x, y, z, a[3], b[3], xp = 0, yp = 0, zp = 0;
    if (can_tma(o)) {
      mbar_wait(&S.inbar[u][pidx][slot], (inphase >> slot) & 1u);
      inphase ^= 1u << slot;
      const double (*in)[WS_NIN] = S.in[u][pidx][slot];
      const int sl = (int)(o.idx - ((o.idx - 1) & ~(uint64_t)1)) + lane;      // 1 or 2, + lane
      x = in[sl][0]; y = in[sl][1]; z = in[sl][2];
#pragma unroll
      for (int c = 0; c < 3; c++) { a[c] = in[sl][3 + c]; b[c] = in[sl][6 + c]; }
      if (lane == 0) { xp = in[sl - 1][0]; yp = in[sl - 1][1]; zp = in[sl - 1][2]; }
    } else {                          // edge of the arrays: plain loads
      x = y = z = 0.0;
#pragma unroll
      for (int c = 0; c < 3; c++) a[c] = b[c] = 0.0;
      const uint64_t e = o.idx + (uint64_t)lane;
      if (active) {
        const double* q = P.pre + e * WS_NIN;
        x = q[0]; y = q[1]; z = q[2];
#pragma unroll
        for (int c = 0; c < 3; c++) { a[c] = q[3 + c]; b[c] = q[6 + c]; }
      }
      if (lane == 0 && itl > 0) { const double* q = P.pre + (e - 1) * WS_NIN; xp = q[0]; yp = q[1]; zp = q[2]; }
    }
    // tau = t - n.r in the reference's operation order (kernel_farfield.cl:65-67); the previous step's tau from the
    // neighbouring lane, lane 0 recomputes it (0 for it == 0: phasePrev starts at 0, Q1)
    const double tau = ssub(smul((double)(o.itStart + itl), P.dt), sdot3(x, y, z, g.nx, g.ny, g.nz));
    double tauPrev = __shfl_up_sync(0xffffffffu, tau, 1);
    if (lane == 0)
      tauPrev = itl == 0 ? 0.0 : ssub(smul((double)(o.itStart + itl - 1), P.dt), sdot3(xp, yp, zp, g.nx, g.ny, g.nz));
    ws_prep_guard<C>(P, g, kc, active, tau, tauPrev, a, b, ws, nPass, nAll);
  };
  auto finish = [&](const Own& o, WsStage<C>& sg, uint32_t fl) {
    const uint32_t fullMask = __ballot_sync(0xffffffffu, fl == 1u);
    const uint32_t anyMask = __ballot_sync(0xffffffffu, fl != 0u);
    if (lane == 0) { sg.cnt = (uint32_t)o.cnt; sg.fullMask = fullMask; sg.anyMask = anyMask; }
  };
  Own A, B, nA, nB;
  next_pair(A, B);
  issue(A, 0); issue(B, 1);
  while (A.valid) {
    next_pair(nA, nB);
    const int sgA = (int)(A.k % (uint32_t)NS), sgB = (int)((A.k + 1u) % (uint32_t)NS);
    mbar_wait(&S.empty[u][sgA], ((A.k / (uint32_t)NS) & 1u) ^ 1u);
    if (B.valid) mbar_wait(&S.empty[u][sgB], ((B.k / (uint32_t)NS) & 1u) ^ 1u);
#if defined(SRB_WS_FAKE_PRODUCER)      // tuning aid: stages are filled once, then only handed over (consumer-only timing)
    if (A.k >= (uint32_t)(2 * NS * NP)) { mbar_arrive(&S.full[u][sgA]); if (B.valid) mbar_arrive(&S.full[u][sgB]); A = nA; B = nB; continue; }
#endif
    const bool subA = A.kind == 0u, subB = B.valid && B.kind == 0u;
    WsStep wa, wb;
    wa.flag = wb.flag = 0u;
    if (subA) guard(A, 0, wa);
    if (subB) guard(B, 1, wb);
    // Every staged input has been CONSUMED by now (tau, tauPrev, the amplitude), not merely requested: a shared-memory
    // load still queued in the LSU could be overtaken by the TMA engine's write (seen as a ~1e-8 run-to-run wobble
    // of the spectrum when the copy was issued right after the loads).  The input buffers are free for this
    // producer's next sub-batches, whose copies then have the phasor part of this item (~2/3 of it) to land.
    __syncwarp();
    issue(nA, 0); issue(nB, 1);
    WsStage<C>& sa = S.stage[u][sgA];
    WsStage<C>& sb = S.stage[u][sgB];
    if (subA && subB && __all_sync(0xffffffffu, wa.flag == 1u && wb.flag == 1u)) {
      // both sub-batches all-pass at every lane (the common case): ONE straight-line block for the two phasor chains
      ws_seeds<C>(P, kc, wa.tau, wa.A, sa, lane);
      ws_seeds<C>(P, kc, wb.tau, wb.A, sb, lane);
      sa.rec[lane][0] = Dbl2{wa.A[0], wa.A[1]}; sa.rec[lane][1] = Dbl2{wa.A[2], wa.tau};
      sb.rec[lane][0] = Dbl2{wb.A[0], wb.A[1]}; sb.rec[lane][1] = Dbl2{wb.A[2], wb.tau};
      sa.rng[lane] = wa.lo | (wa.hi << 10) | (1u << 30);
      sb.rng[lane] = wb.lo | (wb.hi << 10) | (1u << 30);
      if (lane == 0) { sa.cnt = 32u; sa.fullMask = 0xffffffffu; sa.anyMask = 0xffffffffu; sb.cnt = 32u; sb.fullMask = 0xffffffffu; sb.anyMask = 0xffffffffu; }
    } else {
      if (subA) finish(A, sa, ws_prep_store<C>(P, kc, wa, sa, lane));
      if (subB) finish(B, sb, ws_prep_store<C>(P, kc, wb, sb, lane));
    }
    mbar_arrive(&S.full[u][sgA]);
    if (B.valid) mbar_arrive(&S.full[u][sgB]);
    A = nA; B = nB;
  }
  if (P.counters) {      // one atomic pair per warp
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { nPass += __shfl_xor_sync(0xffffffffu, nPass, o); nAll += __shfl_xor_sync(0xffffffffu, nAll, o); }
    if (lane == 0 && nAll) { atomicAdd(P.counters, nPass); atomicAdd(P.counters + 1, nAll); }
  }
}

// Direction-independent per-step kinematics, once per call (instead of once per direction):
// far: a = (beta_{it+1}-beta_it)/dt and b = (beta_{it+1}+beta_it)/2 ; near: beta_it.
// Same strict operation order as the in-kernel path, so results are bit-identical.
__global__ void k_prepass(const srb::Params P, double* __restrict__ pre, uint64_t total, uint64_t stride) {
  if (srb::deselected(P)) return;
  const double dtInv = srb::sdiv(1.0, P.dt);
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t a = 0, b = P.nTracks;          // last track with offsets[t] <= i
    while (b - a > 1) { const uint32_t mid = (a + b) >> 1; if (P.offsets[mid] <= i) a = mid; else b = mid; }
    const uint64_t o = P.offsets[a];
    const uint64_t n = P.offsets[a + 1] - o;
    const uint64_t it = i - o;
    if (P.mode == srb::MODE_FAR) {
      double av[3] = {0, 0, 0}, bv[3] = {0, 0, 0};
      if (it + 1 < n) srb::far_step_kinematics<double>(P.ux, P.uy, P.uz, i, dtInv, av, bv);
      if (P.prePacked) {      // one record per step for the warp-specialised kernel: one TMA copy stages 34 steps
        double* q = pre + i * 9;
        q[0] = ((const double*)P.x)[i]; q[1] = ((const double*)P.y)[i]; q[2] = ((const double*)P.z)[i];
#pragma unroll
        for (int c = 0; c < 3; c++) { q[3 + c] = av[c]; q[6 + c] = bv[c]; }
        continue;
      }
#pragma unroll
      for (int c = 0; c < 3; c++) { pre[c * stride + i] = av[c]; pre[(3 + c) * stride + i] = bv[c]; }
    } else {
      double bv[3];
      srb::near_step_kinematics<double>(P.ux, P.uy, P.uz, i, bv);
#pragma unroll
      for (int c = 0; c < 3; c++) pre[c * stride + i] = bv[c];
    }
  }
}

// Guard statistics of a random sample of (track, step, direction) triples (far field): how many steps
// pass the Nyquist guard at every node of the grid ("full"), at some nodes only ("partial"), or nowhere.
// The pair kernel is faster on full steps (5.0 vs 6.0 FP64 ops per update) but ~3x slower on partial
// ones (its accumulators are sums over node PAIRS), so the planner picks by this ratio.
// The corrected-recurrence kernel rotates by (1 + i c), c ~ one ulp of the phase: its second-order term c^2 / 2 stays
// below 1e-10 up to phases of ~6e10 rad (ulp 1.4e-5); beyond this limit the planner takes the direct kernel instead
// (found by the round-2 fuzz run: near field, omega * L = 7.5e11 rad, 1.2e-9 off).
#define SRB_DREC_MAX_PHASE 3.0e10
__global__ void k_probe(const srb::Params P, uint64_t total, double wFirst, double wLast, unsigned int* out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t h = (uint64_t)(i + 1) * 0x9E3779B97F4A7C15ull; h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
  const uint64_t gs = h % total;
  uint32_t a = 0, b = P.nTracks;
  while (b - a > 1) { const uint32_t mid = (a + b) >> 1; if (P.offsets[mid] <= gs) a = mid; else b = mid; }
  const uint64_t o = P.offsets[a], n = P.offsets[a + 1] - o, it = gs - o;
  if (it == 0 || it + 1 >= n) return;
  const uint32_t d = (uint32_t)((h >> 33) % ((uint64_t)P.nA2 * P.nPhi));
  const uint32_t iA2 = d % P.nA2, iPhi = d / P.nA2;
  const double sT = ((const double*)P.axA)[iA2], cT = ((const double*)P.axB)[iA2];
  const double sP = ((const double*)P.sinPhi)[iPhi], cP = ((const double*)P.cosPhi)[iPhi];
  const double* x = (const double*)P.x; const double* y = (const double*)P.y; const double* z = (const double*)P.z;
  const double dtau = fabs(P.dt - ((x[gs] - x[gs - 1]) * sT * cP + (y[gs] - y[gs - 1]) * sT * sP + (z[gs] - z[gs - 1]) * cT));
  atomicAdd(out, 1u);
  bool some = true;
  if (wLast * dtau < 3.14159265358979323846) atomicAdd(out + 1, 1u);
  else if (wFirst * dtau < 3.14159265358979323846) atomicAdd(out + 2, 1u);
  else some = false;
  // passing steps whose phase at the last node is beyond the 2^18 limit of the phase-tracking kernels (SI-unit tracks)
  const double tau = (double)(P.itStart[a] + (uint32_t)it) * P.dt - (x[gs] * sT * cP + y[gs] * sT * sP + z[gs] * cT);
  if (some && fabs(wLast * tau) > 262144.0) atomicAdd(out + 3, 1u);
  if (some && fabs(wLast * tau) > SRB_DREC_MAX_PHASE) atomicAdd(out + 4, 1u);     // beyond the corrected recurrence's first-order correction
}

// phasor = AUTO: the decision from the probe counts, on the device (no host round trip): partial steps cost the pair
// kernel ~3x, so it is taken only when they are < 10 % of the all-pass ones.  sel[0] = chosen kind; the same code is
// left in counters[2] for the caller.
__global__ void k_decide(const unsigned int* probe, int32_t* sel, unsigned long long* counters, int haveDrec) {
  const bool preferRecur = probe[0] > 0 && (double)probe[2] > 0.1 * (double)probe[1];
  // mostly all-pass steps with huge phases (SI-unit tracks): neither phase-tracking kernel applies (every step would be
  // evaluated node by node), the corrected-recurrence kernel does (srb_drec.cuh)
  const bool big = haveDrec && probe[4] == 0u && 2.0 * (double)probe[3] > (double)probe[1] + (double)probe[2];
  sel[0] = preferRecur ? srb::KIND_RECUR : (big ? srb::KIND_DREC : srb::KIND_PAIR);
  if (counters) counters[2] = (unsigned long long)sel[0];
}

// out[c][i] += sum over particle chunks of the private partial spectra (fixed order: deterministic)
__global__ void k_reduce_slabs(srb::Params P, int nOut, size_t perOut) {
  if (srb::deselected(P)) return;
  const size_t n = (size_t)nOut * perOut;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    double acc = 0.0;
    for (uint32_t s = 0; s + 1 < P.nPC; s++) acc += P.slabs[(size_t)s * P.slabStride + i];
    const int c = (int)(i / perOut);
    P.out[c][i - (size_t)c * perOut] += acc;
  }
}

// time-axis split, second pass (srb_core.cuh: combine_node): thread = (snapshot, node)
template <int MODE, int NCF>
__global__ void k_combine_segments(const srb::Params P, size_t n) {
  if (srb::deselected(P)) return;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    srb::combine_node<MODE, NCF>(P, i);
}

__global__ void k_swap_axes(const double* __restrict__ src, double* __restrict__ dst, uint32_t nSnaps,
                            uint32_t nO, uint32_t nA, uint32_t nP) {
  // src (nSnaps, nP, nA, nO) -> dst (nSnaps, nO, nA, nP)
  const size_t n = (size_t)nSnaps * nO * nA * nP;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t io = (uint32_t)(i % nO); size_t r = i / nO;
    const uint32_t ia = (uint32_t)(r % nA); r /= nA;
    const uint32_t ip = (uint32_t)(r % nP); const size_t is = r / nP;
    dst[((is * nO + io) * nA + ia) * nP + ip] = src[i];
  }
}

// Angle integrals of utils.py:75-95 on the device: thread = (omega node, phi slice); the phi slices of a node are
// summed through shared memory in a fixed order.
struct SpecPtrs { const double* p[6]; };
__global__ void k_energy_spectrum(SpecPtrs S, int nK, int coherent, int far, int layout, uint32_t nO, uint32_t nA, uint32_t nP,
                                  uint32_t iSnap, const double* __restrict__ ax, double dphi, double* __restrict__ out) {
  __shared__ double part[8][33];
  const uint32_t j = blockIdx.x * 32u + threadIdx.x;
  double acc = 0.0;
  if (j < nO) {
    for (uint32_t ip = threadIdx.y; ip < nP; ip += 8u) {
      // element (iSnap, j, a, ip) in either layout
      const size_t base = layout == 0 ? ((size_t)iSnap * nP + ip) * nA * (size_t)nO + j
                                      : ((size_t)iSnap * nO + j) * nA * (size_t)nP + ip;
      const size_t sA = layout == 0 ? (size_t)nO : (size_t)nP;
      auto val = [&](uint32_t a) {
        double v = 0.0;
        for (int k = 0; k < nK; k++) { const double s = S.p[k][base + (size_t)a * sA]; v += coherent ? s * s : s; }
        return v;
      };
      double sum = 0.0;
      if (far) {
        if (nA >= 3) {
          double v0 = val(0), v1 = val(1);
          double tm0 = 0.5 * (ax[1] + ax[0]);
          double y0 = 0.5 * (v1 + v0) * sin(tm0);
          for (uint32_t a = 1; a + 1 < nA; a++) {
            const double v2 = val(a + 1);
            const double tm1 = 0.5 * (ax[a + 1] + ax[a]);
            const double y1 = 0.5 * (v2 + v1) * sin(tm1);
            sum += 0.5 * (y1 + y0) * (tm1 - tm0);
            v1 = v2; tm0 = tm1; y0 = y1;
          }
        }
      } else {
        if (nA >= 2) {
          double y0 = val(0) * ax[0];
          for (uint32_t a = 1; a < nA; a++) {
            const double y1 = val(a) * ax[a];
            sum += 0.5 * (y1 + y0) * (ax[a] - ax[a - 1]);
            y0 = y1;
          }
        }
      }
      acc += sum;
    }
  }
  part[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && j < nO) {
    double t = 0.0;
    for (int k = 0; k < 8; k++) t += part[k][threadIdx.x];
    out[j] = dphi * t;
  }
}

// Issue-rate micro-kernels: the denominators of the compute roofline, measured on
// the device the benchmark runs on.  8 independent FMA chains per thread, 16 warps per SM.
template <typename T, bool MUFU>
__global__ void k_pipe_peak(T* out, T a, T b, int iters) {
  T v[8];
#pragma unroll
  for (int i = 0; i < 8; i++) v[i] = (T)threadIdx.x * (T)1e-3 + (T)i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MUFU) v[i] = (T)__sinf((float)v[i]);
      else v[i] = fma(v[i], a, b);
    }
  }
  T s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += v[i];
  if (s == (T)123.456) out[0] = s;
}

struct Launcher {
  void (*kernel)(const srb::Params);
  size_t smem;
  int chunk;
  int units;      // virtual directions per block
  int threads;    // block size
  int minBlocks;  // resident blocks per SM the register budget was tuned for (shared-memory carve-out request)
};

template <class C> Launcher make_launcher() {
  return Launcher{&k_integrate<C>, sizeof(srb::WarpSmem<C>) * NW, C::CHUNK, NW, NW * 32, min_blocks<C>()};
}
template <class C, int NU, int NP, int NS, int NCW> Launcher make_launcher_ws() {
  static_assert(sizeof(WsSmem<C, NU, NP, NS>) <= 227 * 1024, "ring does not fit the shared memory of an SM");
  return Launcher{&k_integrate_ws<C, NU, NP, NS, NCW>, sizeof(WsSmem<C, NU, NP, NS>), C::CHUNK, NU, (NCW + NP) * NU * 32, 1};
}
#ifndef SRB_WS_NP
#define SRB_WS_NP 2     // producer warps per unit
#endif
#ifndef SRB_WS_NS
#define SRB_WS_NS 4     // stages of a unit's ring
#endif
#ifndef SRB_WS_NCW
#define SRB_WS_NCW 2    // DMMA consumer warps per unit
#endif
// warp-specialised form of the fp64 pair kernel on the tensor cores (tile width x components % 8 == 0); units per block
// chosen so that the two-stage rings + input buffers fit the 227 KB of an SM
bool pick_ws(int tw, int nc, Launcher* L) {
  using srb::Cfg; using srb::KIND_PAIR; using srb::MODE_FAR;
  if (tw == 8 && nc == 2) { *L = make_launcher_ws<Cfg<double, double, MODE_FAR, KIND_PAIR, 8, false, 2>, 4, SRB_WS_NP, SRB_WS_NS, SRB_WS_NCW>(); return true; }
  if (tw == 4 && nc == 2) { *L = make_launcher_ws<Cfg<double, double, MODE_FAR, KIND_PAIR, 4, false, 2>, 4, SRB_WS_NP, SRB_WS_NS, SRB_WS_NCW>(); return true; }
  // 16-node tiles (grids with more than 256 omega nodes) and the three-component spheric kernels keep the
  // warp-autonomous form of srb_pair.cuh: their stages (16 KB) do not leave room for a four-deep ring
  return false;
}

using srb::KIND_DREC; using srb::Cfg; using srb::KIND_DIRECT; using srb::KIND_RECUR; using srb::KIND_LITERAL; using srb::KIND_PAIR; using srb::KIND_PAIR_FMA; using srb::MODE_FAR; using srb::MODE_NEAR;

// kind, mode, dtype, native, tile width, far components -> kernel
bool pick(int kind, int mode, int dtype, bool native, int tw, int nc, Launcher* L) {
#define SRB_CASE(K, M, D, NAT, TWV, NCV, TM)                                                        \
  if (kind == K && mode == M && dtype == D && native == NAT && tw == TWV && nc == NCV) {            \
    *L = make_launcher<Cfg<double, TM, M, K, TWV, NAT, NCV>>(); return true; }
#define SRB_BOTH(K, M, NAT, TWV, NCV) SRB_CASE(K, M, 0, NAT, TWV, NCV, double) SRB_CASE(K, M, 1, NAT, TWV, NCV, float)
  // recurrence, far: transverse (NC=2) for total/cartesian(_complex), NC=3 for the spheric kernels
  SRB_BOTH(KIND_RECUR, MODE_FAR, false, 16, 2) SRB_BOTH(KIND_RECUR, MODE_FAR, false, 8, 2) SRB_BOTH(KIND_RECUR, MODE_FAR, false, 4, 2)
  SRB_BOTH(KIND_RECUR, MODE_FAR, false, 16, 3) SRB_BOTH(KIND_RECUR, MODE_FAR, false, 8, 3) SRB_BOTH(KIND_RECUR, MODE_FAR, false, 4, 3)
  // recurrence, near
  SRB_BOTH(KIND_RECUR, MODE_NEAR, false, 8, 3) SRB_BOTH(KIND_RECUR, MODE_NEAR, false, 4, 3) SRB_BOTH(KIND_RECUR, MODE_NEAR, false, 2, 3)
  // direct
  SRB_BOTH(KIND_DIRECT, MODE_FAR, false, 8, 3) SRB_BOTH(KIND_DIRECT, MODE_FAR, false, 4, 3) SRB_BOTH(KIND_DIRECT, MODE_FAR, false, 2, 3)
  SRB_BOTH(KIND_DIRECT, MODE_NEAR, false, 8, 3) SRB_BOTH(KIND_DIRECT, MODE_NEAR, false, 4, 3) SRB_BOTH(KIND_DIRECT, MODE_NEAR, false, 2, 3)
  // direct layout, corrected recurrence: fp64, uniform grids, any phase magnitude (srb_drec.cuh)
  SRB_CASE(KIND_DREC, MODE_FAR, 0, false, 8, 3, double) SRB_CASE(KIND_DREC, MODE_FAR, 0, false, 4, 3, double) SRB_CASE(KIND_DREC, MODE_FAR, 0, false, 2, 3, double)
  SRB_CASE(KIND_DREC, MODE_NEAR, 0, false, 8, 3, double) SRB_CASE(KIND_DREC, MODE_NEAR, 0, false, 4, 3, double) SRB_CASE(KIND_DREC, MODE_NEAR, 0, false, 2, 3, double)
  // direct, fp32 native sincos
  SRB_CASE(KIND_DIRECT, MODE_FAR, 1, true, 8, 3, float) SRB_CASE(KIND_DIRECT, MODE_FAR, 1, true, 4, 3, float)
  SRB_CASE(KIND_DIRECT, MODE_FAR, 1, true, 2, 3, float)
  SRB_CASE(KIND_DIRECT, MODE_NEAR, 1, true, 8, 3, float) SRB_CASE(KIND_DIRECT, MODE_NEAR, 1, true, 4, 3, float)
  SRB_CASE(KIND_DIRECT, MODE_NEAR, 1, true, 2, 3, float)
  // symmetric-pair kernel: far field, transverse basis
  SRB_BOTH(KIND_PAIR, MODE_FAR, false, 8, 2) SRB_BOTH(KIND_PAIR, MODE_FAR, false, 4, 2) SRB_BOTH(KIND_PAIR, MODE_FAR, false, 2, 2)
  SRB_BOTH(KIND_PAIR, MODE_FAR, false, 16, 2)      // grids with > 256 omega nodes (fp64: tensor-core layout, 2 blocks/SM)
  SRB_BOTH(KIND_PAIR, MODE_FAR, false, 8, 3) SRB_BOTH(KIND_PAIR, MODE_FAR, false, 4, 3) SRB_BOTH(KIND_PAIR, MODE_FAR, false, 2, 3)   // spheric kernels
  // the same kernel without tensor cores, where KIND_PAIR uses them (fp64, TW*NC % 8 == 0); phasor = SRB_PHASOR_PAIR_FMA
  SRB_CASE(KIND_PAIR_FMA, MODE_FAR, 0, false, 8, 2, double) SRB_CASE(KIND_PAIR_FMA, MODE_FAR, 0, false, 4, 2, double)
  SRB_CASE(KIND_PAIR_FMA, MODE_FAR, 0, false, 8, 3, double) SRB_CASE(KIND_PAIR_FMA, MODE_FAR, 0, false, 16, 2, double)
  // literal fp32 (dtype 2)
  SRB_CASE(KIND_LITERAL, MODE_FAR, 2, false, 8, 3, float) SRB_CASE(KIND_LITERAL, MODE_FAR, 2, false, 4, 3, float)
  SRB_CASE(KIND_LITERAL, MODE_FAR, 2, false, 2, 3, float)
  SRB_CASE(KIND_LITERAL, MODE_NEAR, 2, false, 8, 3, float) SRB_CASE(KIND_LITERAL, MODE_NEAR, 2, false, 4, 3, float)
  SRB_CASE(KIND_LITERAL, MODE_NEAR, 2, false, 2, 3, float)
#undef SRB_BOTH
#undef SRB_CASE
  return false;
}

struct Plan {
  int kind, tw, nc;
  size_t preDoubles;      // doubles of scratch used by the per-step pre-pass (0 = computed in-kernel)
  bool native, ws;
  Launcher L;
  uint32_t chunkNodes, nChunks, nVD, nVDtiles, nPC;
  uint32_t nTS;           // time segments per track (1: off); then nPC = nTracks * nTS and `ampDoubles` replace the slabs
  int blocksPerSM, numSM;
  size_t perOut, slabDoubles, ampDoubles;
  int nOut;
  size_t work_doubles() const { return nTS > 1 ? ampDoubles : (size_t)(nPC - 1) * slabDoubles; }   // after the pre-pass area
};

int validate(const srb_grid* g, const srb_tracks* t) {
  if (!g || !t) return fail("null grid/tracks");
  if (g->mode != SRB_MODE_FAR && g->mode != SRB_MODE_NEAR) return fail("bad mode");
  if (srb_num_spectra(g->mode, g->comp) < 0)
    return fail("no such kernel: comp/mode combination does not exist in the reference");
  if (g->dtype != SRB_DTYPE_F64 && g->dtype != SRB_DTYPE_F32 && g->dtype != SRB_DTYPE_F32_LITERAL) return fail("bad dtype");
  if (g->nOmega == 0 || g->nAxis2 == 0 || g->nPhi == 0) return fail("empty grid");
  if (g->nSnaps == 0) return fail("nSnaps must be >= 1");
  if ((uint64_t)g->nOmega * g->nAxis2 * g->nPhi >= (1ull << 32)) return fail("grid too large (uint32 node index, as in the reference)");
  if (!g->omega || !g->sinPhi || !g->cosPhi) return fail("null grid table");
  if (g->mode == SRB_MODE_FAR && (!g->sinTheta || !g->cosTheta)) return fail("far mode needs sinTheta/cosTheta");
  if (g->mode == SRB_MODE_NEAR && !g->radius) return fail("near mode needs radius");
  if (t->nTracks && (!t->x || !t->y || !t->z || !t->ux || !t->uy || !t->uz || !t->offsets || !t->w ||
                     !t->itStart || !t->itEnd || !t->itSnaps))
    return fail("null track array");
  if (t->itSnapsStride != 0 && t->itSnapsStride != g->nSnaps) return fail("itSnapsStride must be 0 or nSnaps");
  return 0;
}

int make_plan(const srb_grid* g, const srb_tracks* t, size_t scratch_bytes, bool unlimited, Plan* p,
              bool preferRecur = false, bool forceDrec = false) {
  int dev = 0;
  SRB_CUDA(cudaGetDevice(&dev));
  SRB_CUDA(cudaDeviceGetAttribute(&p->numSM, cudaDevAttrMultiProcessorCount, dev));
  const bool uniform = g->omega_uniform && g->nOmega >= 2 && g->omega_last_host > g->omega_first_host;
  if (g->phasor == SRB_PHASOR_RECUR && !uniform) return fail("phasor recurrence needs an ascending uniform omega grid");
  const bool spheric = g->comp == SRB_COMP_SPHERIC || g->comp == SRB_COMP_SPHERIC_COMPLEX;
  const bool pairOk = uniform && g->mode == SRB_MODE_FAR;
  if ((g->phasor == SRB_PHASOR_PAIR || g->phasor == SRB_PHASOR_PAIR_FMA) && !pairOk) return fail("the pair kernel needs the far field and an ascending uniform omega grid");
  if (g->phasor == SRB_PHASOR_DIRECT || !uniform) p->kind = KIND_DIRECT;
  else if (g->phasor == SRB_PHASOR_RECUR || !pairOk || (preferRecur && g->phasor == SRB_PHASOR_AUTO)) p->kind = KIND_RECUR;
  else p->kind = KIND_PAIR;
  // near field: phase = omega*(t+R) ~ omega*L.  Beyond 2^18 rad the recurrence cannot track the
  // reference's rounded phase to 1e-9 (srb_core.cuh, flag 3), every step would fall back: the corrected-recurrence
  // kernel (srb_drec.cuh: direct layout, per-update correction onto the rounded phase) takes over.
  if (g->phasor == SRB_PHASOR_AUTO && g->mode == SRB_MODE_NEAR && g->dtype == SRB_DTYPE_F64 && uniform &&
      std::fabs(g->omega_last_host * g->L_screen) > 262144.0)
    p->kind = std::fabs(g->omega_last_host * g->L_screen) <= SRB_DREC_MAX_PHASE ? KIND_DREC : KIND_DIRECT;
  if (g->phasor == SRB_PHASOR_DREC) {
    if (!uniform || g->dtype != SRB_DTYPE_F64) return fail("the corrected-recurrence kernel needs fp64 and an ascending uniform omega grid");
    if (g->mode == SRB_MODE_NEAR && std::fabs(g->omega_last_host * g->L_screen) > SRB_DREC_MAX_PHASE)
      return fail("the corrected-recurrence kernel is first-order in one ulp of the phase: omega * L beyond 3e10 rad needs phasor = direct (or auto)");
    p->kind = KIND_DREC;
  }
  if (forceDrec && uniform && g->dtype == SRB_DTYPE_F64) p->kind = KIND_DREC;     // third candidate of phasor = AUTO (far field)
  if (g->dtype == SRB_DTYPE_F32_LITERAL) p->kind = KIND_LITERAL;
  p->native = (p->kind == KIND_DIRECT && g->dtype == SRB_DTYPE_F32 && g->native != 0);   // Q9
  const int tiles = p->kind == KIND_RECUR ? 16 : 32;
  if (p->kind == KIND_LITERAL && g->phasor == SRB_PHASOR_RECUR) return fail("the literal fp32 kernels have no recurrence variant");
  int twMax, twMin;
  if (p->kind == KIND_RECUR) { twMax = g->mode == SRB_MODE_FAR ? 16 : 8; twMin = twMax / 4; }
  else if (p->kind == KIND_PAIR) { twMax = !spheric ? 16 : 8; twMin = 2; }
  else if (p->kind == KIND_DREC) { twMax = 8; twMin = 2; }
  else { twMax = g->dtype == SRB_DTYPE_F64 ? 4 : 8; twMin = 2; }   // fp64 direct: 4 nodes/lane measured fastest
  p->tw = twMax;
  for (int tw = twMin; tw <= twMax; tw *= 2) if ((uint32_t)(tiles * tw) >= g->nOmega) { p->tw = tw; break; }
  if (const char* f = std::getenv("SRB_FORCE_TW")) {        // tuning aid
    const int tw = std::atoi(f);
    if (tw >= twMin && tw <= twMax && (tw & (tw - 1)) == 0) p->tw = tw;
  }
  p->nc = ((p->kind == KIND_PAIR || p->kind == KIND_RECUR) && g->mode == SRB_MODE_FAR && !spheric) ? 2 : 3;
  // SRB_PHASOR_PAIR_FMA: scalar-pipe accumulation; only differs where the pair kernel would use DMMA
  if (g->phasor == SRB_PHASOR_PAIR_FMA && p->kind == KIND_PAIR && g->dtype == SRB_DTYPE_F64 && (p->tw * p->nc) % 8 == 0)
    p->kind = KIND_PAIR_FMA;
  if (!pick(p->kind, g->mode, g->dtype, p->native, p->tw, p->nc, &p->L)) return fail("internal: no kernel for this configuration");
  // scratch = [pre-pass planes][private partial spectra]; the pre-pass is used when it fits
  // (plane stride rounded up to even: every plane starts 16-byte aligned, as the TMA staging of srb_ws.cuh wants)
  p->preDoubles = p->kind == KIND_LITERAL ? 0 : (size_t)(g->mode == SRB_MODE_FAR ? 6 : 3) * ((t->totalSteps_host + 1) & ~(uint64_t)1);
  if (!unlimited) {
    if (scratch_bytes >= p->preDoubles * 8) scratch_bytes -= p->preDoubles * 8;
    else p->preDoubles = 0;
  }
  // fp64 pair kernel on the tensor cores: warp-specialised form (srb_ws.cuh) whenever the pre-pass planes exist
  // (SRB_WS=0 keeps the round-1 warp-autonomous form: A/B measurements)
  p->ws = false;
  {
    const char* e = std::getenv("SRB_WS");
    if (p->kind == KIND_PAIR && g->dtype == SRB_DTYPE_F64 && (p->tw * p->nc) % 8 == 0 && p->preDoubles > 0 &&
        !(e && std::atoi(e) == 0))
      p->ws = pick_ws(p->tw, p->nc, &p->L);
  }
  if (p->ws) {             // packed 9-double records (x, y, z, a, b) instead of 6 planes
    const size_t packed = 9 * (size_t)t->totalSteps_host + 2;
    if (unlimited || scratch_bytes + p->preDoubles * 8 >= packed * 8) {
      if (!unlimited) scratch_bytes = scratch_bytes + p->preDoubles * 8 - packed * 8;
      p->preDoubles = packed;
    } else {               // scratch too small for the records: the warp-autonomous form with the planes
      p->ws = false;
      pick(p->kind, g->mode, g->dtype, p->native, p->tw, p->nc, &p->L);
    }
  }
  p->chunkNodes = (uint32_t)p->L.chunk;
  p->nChunks = (g->nOmega + p->chunkNodes - 1) / p->chunkNodes;
  p->nVD = g->nPhi * g->nAxis2 * p->nChunks;
  p->nVDtiles = (p->nVD + p->L.units - 1) / p->L.units;
  p->nOut = srb_num_spectra(g->mode, g->comp);
  p->perOut = (size_t)g->nSnaps * g->nOmega * g->nAxis2 * g->nPhi;
  p->slabDoubles = p->perOut * p->nOut;
  SRB_CUDA(cudaFuncSetAttribute(p->L.kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p->L.smem));
  SRB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&p->blocksPerSM, p->L.kernel, p->L.threads, p->L.smem));
  if (p->blocksPerSM < p->L.minBlocks) {
    // the default L1 / shared-memory split leaves fewer resident blocks than the registers were tuned for (seen on the
    // corrected-recurrence kernel: 2 instead of 3): ask for the carve-out that fits them.  Only then -- the kernels
    // that stream tracks through L1 lose ~8 % with a smaller L1 (C4 recipe, profiles/r02_guard_dominated.md).
    const size_t want = (p->L.smem + 1024) * (size_t)p->L.minBlocks;
    const int pct = (int)std::min<size_t>(100, want * 100 / (228 * 1024) + 1);
    const char* ce = std::getenv("SRB_CARVEOUT");             // (0: leave the driver's default; A/B measurements)
    if (!(ce && std::atoi(ce) == 0)) {
      SRB_CUDA(cudaFuncSetAttribute(p->L.kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct));
      SRB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&p->blocksPerSM, p->L.kernel, p->L.threads, p->L.smem));
    }
  }
  if (p->blocksPerSM < 1) return fail("kernel does not fit on an SM");
  // particle chunks: fill whole waves of the machine; bounded by tracks, scratch and 64 waves
  const uint64_t slots = (uint64_t)p->numSM * p->blocksPerSM;
  uint64_t maxPC = t->nTracks ? t->nTracks : 1;
  const uint64_t slabCap = unlimited ? (uint64_t)(1ull << 30) / (p->slabDoubles * 8) : scratch_bytes / (p->slabDoubles * 8);
  maxPC = std::min<uint64_t>(maxPC, 1 + slabCap);
  // ... and by the L2: every direction block of a particle chunk streams the chunk's tracks, concurrently running blocks
  // drift apart, so a chunk that does not stay L2-resident is re-read from DRAM by most of them (12 500 x 10^4 steps in
  // 37 chunks of 243 MB: 262 GB of DRAM traffic per launch against 6 GB of tracks, profiles/r02_ncu_headline.json of
  // that build).  Chunks of <= 32 MB of streamed input (a quarter of the 126 MB L2) when tracks and scratch allow.
  const uint64_t streamBytes = (p->preDoubles ? p->preDoubles * 8 : 0) + (p->ws ? 0 : 24 * t->totalSteps_host);
  const uint64_t l2PC = (streamBytes + (32ull << 20) - 1) / (32ull << 20);
  maxPC = std::min<uint64_t>(maxPC, std::max<uint64_t>(std::max<uint64_t>(1, (64 * slots) / p->nVDtiles), l2PC));
  maxPC = std::min<uint64_t>(maxPC, 4096);
  maxPC = std::min<uint64_t>(maxPC, std::max<uint64_t>(1, 0x7fffffffull / p->nVDtiles));   // gridDim.x limit
  double best = -1.0; uint32_t bestN = 1;
  for (uint64_t n = 1; n <= maxPC; n++) {
    const uint64_t blocks = (uint64_t)p->nVDtiles * n;
    const uint64_t waves = (blocks + slots - 1) / slots;
    const double eff = (double)blocks / (double)(waves * slots);
    if (eff >= best - 1e-12) { best = eff; bestN = (uint32_t)n; }
  }
  p->nPC = bestN;
  // Time-axis split (the reference's serial loop kernel_farfield.cl:59): with few particles even one track
  // per block leaves the machine part-empty (single electron on a 128x16x16 grid: 64 blocks of 4 directions for 148
  // SMs), so every track is cut into nTS segments of whole sub-batches and a second pass squares the summed partial
  // amplitudes.  A block has a fixed cost (pipeline start-up, flush: ~13 us for the warp-specialised kernel, measured),
  // so the split is taken only when the particle chunks fill less than 3/4 of the block slots of their waves, with
  // segments of at least 8 sub-batches.  SRB_TIME_SPLIT=0 disables, =N forces N (tests).
  p->nTS = 1; p->ampDoubles = 0;
  if (p->kind != KIND_LITERAL && t->nTracks > 0) {
    const int ncf = g->mode == SRB_MODE_FAR ? p->nc : 3;
    const size_t ampPer = (size_t)t->nTracks * g->nSnaps * 2 * ncf * ((size_t)g->nOmega * g->nAxis2 * g->nPhi);   // doubles per segment count
    const uint64_t capBytes = unlimited ? (1ull << 30) : scratch_bytes;
    const uint64_t avgSteps = t->totalSteps_host / t->nTracks;
    const uint64_t fit = std::min<uint64_t>(capBytes / (ampPer * 8),
                                            std::max<uint64_t>(1, 0x7fffffffull / ((uint64_t)p->nVDtiles * t->nTracks)));
    uint32_t want = 0;
    const char* e = std::getenv("SRB_TIME_SPLIT");
    if (e) want = (uint32_t)std::atoi(e);
    if (e && want >= 2) p->nTS = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(want, fit));
    else if (!e && best < 0.75) {
      const uint64_t maxTS = std::min<uint64_t>(std::min<uint64_t>(64, avgSteps / 256), fit);
      auto effOf = [&](uint64_t blocks) { const uint64_t waves = (blocks + slots - 1) / slots; return (double)blocks / (double)(waves * slots); };
      double bestTS = 0.0;
      for (uint64_t n = 2; n <= maxTS; n++) bestTS = std::max(bestTS, effOf((uint64_t)p->nVDtiles * t->nTracks * n));
      for (uint64_t n = 2; n <= maxTS; n++)
        if (effOf((uint64_t)p->nVDtiles * t->nTracks * n) >= bestTS - 0.02) { if (bestTS > best + 0.1) p->nTS = (uint32_t)n; break; }
    }
    if (p->nTS > 1) { p->nPC = t->nTracks * p->nTS; p->ampDoubles = ampPer * p->nTS; }
  }
  return 0;
}

}  // namespace

extern "C" {

int srb_version(void) { return SRB_ABI_VERSION; }
const char* srb_last_error(void) { return g_err.c_str(); }

int srb_num_spectra(int mode, int comp) {
  if (mode != SRB_MODE_FAR && mode != SRB_MODE_NEAR) return -1;
  switch (comp) {
    case SRB_COMP_TOTAL: return 1;
    case SRB_COMP_CARTESIAN: return 3;
    case SRB_COMP_CARTESIAN_COMPLEX: return 6;
    case SRB_COMP_SPHERIC: return mode == SRB_MODE_FAR ? 3 : -1;
    case SRB_COMP_SPHERIC_COMPLEX: return mode == SRB_MODE_FAR ? 6 : -1;
    default: return -1;
  }
}

size_t srb_scratch_bytes(const srb_grid* grid, const srb_tracks* tracks) {
  if (validate(grid, tracks) != 0) return 0;
  Plan p, q;
  if (make_plan(grid, tracks, 0, true, &p, false) != 0) return 0;
  if (make_plan(grid, tracks, 0, true, &q, true) != 0) return 0;
  const size_t a = (p.work_doubles() + p.preDoubles) * sizeof(double);
  const size_t b = (q.work_doubles() + q.preDoubles) * sizeof(double);
  size_t c = 0;
  if (grid->dtype == SRB_DTYPE_F64 && grid->mode == SRB_MODE_FAR && grid->phasor == SRB_PHASOR_AUTO) {
    Plan r;
    if (make_plan(grid, tracks, 0, true, &r, false, true) == 0) c = (r.work_doubles() + r.preDoubles) * sizeof(double);
  }
  return 64 + std::max(std::max(a, b), c);        // 64-byte header: guard probe counts + on-device kernel choice (phasor = AUTO)
}

// one plan's launches: [pre-pass] [zero the private partial spectra] kernel [reduce]; `sel`/`want`: see Params::sel
static int launch_plan(const srb_grid* g, const srb_tracks* t, const Plan& p, double* const* spectra, void* scratch,
                       uint64_t* counters, cudaStream_t stream, const int32_t* sel, int want, uint32_t* launched) {
  srb::Params P;
  std::memset(&P, 0, sizeof P);
  P.mode = g->mode; P.comp = g->comp;
  P.nOmega = g->nOmega; P.nA2 = g->nAxis2; P.nPhi = g->nPhi; P.nSnaps = g->nSnaps;
  P.omega = g->omega;
  P.axA = g->mode == SRB_MODE_FAR ? g->sinTheta : g->radius;
  P.axB = g->mode == SRB_MODE_FAR ? g->cosTheta : nullptr;
  P.sinPhi = g->sinPhi; P.cosPhi = g->cosPhi; P.formFactor = g->formFactor;
  P.L = g->L_screen; P.dt = g->dt;
  P.descending = g->omega_last_host < g->omega_first_host ? 1 : 0;
  P.domega = g->nOmega > 1 ? (g->omega_last_host - g->omega_first_host) / (double)(g->nOmega - 1) : 0.0;
  P.chunkNodes = p.chunkNodes; P.nChunks = p.nChunks; P.nVD = p.nVD;
  P.nTracks = t->nTracks;
  P.x = t->x; P.y = t->y; P.z = t->z; P.ux = t->ux; P.uy = t->uy; P.uz = t->uz;
  P.offsets = t->offsets; P.w = t->w; P.itStart = t->itStart; P.itEnd = t->itEnd; P.itSnaps = t->itSnaps;
  P.snapStride = t->itSnapsStride;
  for (int c = 0; c < p.nOut; c++) P.out[c] = spectra[c];
  double* preBuf = p.preDoubles ? (double*)scratch : nullptr;
  P.pre = preBuf; P.preStride = (t->totalSteps_host + 1) & ~(uint64_t)1;
  // TMA bulk copies need 16-byte aligned sources
  P.tmaOK = (((uintptr_t)preBuf & 15u) == 0 && !std::getenv("SRB_WS_NOTMA")) ? 1 : 0;      // (env: debugging aid)
  P.prePacked = p.ws ? 1 : 0;
  P.slabs = (double*)scratch + p.preDoubles; P.slabStride = p.slabDoubles; P.nPC = p.nPC;
  P.nTS = p.nTS; P.amp = p.nTS > 1 ? (double*)scratch + p.preDoubles : nullptr;
  P.counters = (unsigned long long*)counters;
  P.sel = sel; P.selWant = want;
  if (preBuf) {
    const uint64_t total = t->totalSteps_host;
    const int pb = (int)std::min<uint64_t>((total + 255) / 256, (uint64_t)p.numSM * 16);
    k_prepass<<<pb, 256, 0, stream>>>(P, preBuf, total, P.preStride);
    SRB_CUDA(cudaGetLastError());
    (*launched)++;
  }
  if (p.work_doubles()) SRB_CUDA(cudaMemsetAsync(P.slabs, 0, p.work_doubles() * sizeof(double), stream));
  const uint32_t blocks = p.nVDtiles * p.nPC;
  p.L.kernel<<<blocks, p.L.threads, p.L.smem, stream>>>(P);
  SRB_CUDA(cudaGetLastError());
  (*launched)++;
  if (p.nTS > 1) {
    const size_t n = p.perOut;       // (snapshot, node) elements
    const int rb = (int)std::min<size_t>((n + 127) / 128, (size_t)p.numSM * 16);
    const int ncf = g->mode == SRB_MODE_FAR ? p.nc : 3;
    if (g->mode == SRB_MODE_NEAR) k_combine_segments<srb::MODE_NEAR, 3><<<rb, 128, 0, stream>>>(P, n);
    else if (ncf == 2) k_combine_segments<srb::MODE_FAR, 2><<<rb, 128, 0, stream>>>(P, n);
    else k_combine_segments<srb::MODE_FAR, 3><<<rb, 128, 0, stream>>>(P, n);
    SRB_CUDA(cudaGetLastError());
    (*launched)++;
  } else if (p.nPC > 1) {
    const size_t n = p.slabDoubles;
    const int rb = (int)std::min<size_t>((n + 255) / 256, (size_t)p.numSM * 8);
    k_reduce_slabs<<<rb, 256, 0, stream>>>(P, p.nOut, p.perOut);
    SRB_CUDA(cudaGetLastError());
    (*launched)++;
  }
  return 0;
}

static void fill_info(const Plan& p, uint32_t launched, int kind) {
  g_info.kind = kind; g_info.tile_width = p.tw; g_info.chunk_nodes = p.chunkNodes; g_info.n_chunks = p.nChunks;
  g_info.n_virtual_dirs = p.nVD; g_info.n_particle_chunks = p.nPC; g_info.grid_blocks = p.nVDtiles * p.nPC;
  g_info.block_threads = (uint32_t)p.L.threads; g_info.n_components = (uint32_t)p.nc; g_info.smem_bytes = (uint32_t)p.L.smem;
  g_info.kernels_launched = launched; g_info.n_time_segments = p.nTS;
}

int srb_integrate(const srb_grid* g, const srb_tracks* t, double* const* spectra, int n_spectra,
                  void* scratch, size_t scratch_bytes, uint64_t* counters, void* stream_) {
  if (validate(g, t) != 0) return -1;
  cudaStream_t stream = (cudaStream_t)stream_;
  const int nOut = srb_num_spectra(g->mode, g->comp);
  if (n_spectra != nOut || !spectra) return fail("n_spectra does not match comp");
  for (int c = 0; c < nOut; c++) if (!spectra[c]) return fail("null spectrum buffer");
  std::memset(&g_info, 0, sizeof g_info);
  if (counters) SRB_CUDA(cudaMemsetAsync(counters, 0, 4 * sizeof(uint64_t), stream));
  if (t->nTracks == 0) return 0;
  uint32_t launched = 0;
  // phasor = AUTO with both uniform-grid kernels eligible (far field): the pair kernel is faster on all-pass steps but
  // ~3x slower on partially passing ones, so a probe kernel samples the Nyquist-guard statistics of 8192 random
  // (track, step, direction) triples and the choice is made ON THE DEVICE: both candidates are enqueued and the one
  // that was not chosen returns at once.  No host synchronisation; the choice is left in counters[2].
  const bool uniform = g->omega_uniform && g->nOmega >= 2 && g->omega_last_host > g->omega_first_host;
  if (g->phasor == SRB_PHASOR_AUTO && uniform && g->mode == SRB_MODE_FAR && g->dtype != SRB_DTYPE_F32_LITERAL &&
      scratch && scratch_bytes >= 256 && t->totalSteps_host > 2) {
    unsigned int* probeBuf = (unsigned int*)scratch;              // 64-byte header of the scratch: probe counts, decision
    int32_t* sel = (int32_t*)scratch + 8;
    void* body = (char*)scratch + 64;
    Plan pPair, pRec;
    if (make_plan(g, t, scratch_bytes - 64, false, &pPair, false) != 0) return -1;
    if (make_plan(g, t, scratch_bytes - 64, false, &pRec, true) != 0) return -1;
    if (pPair.kind != pRec.kind) {
      srb::Params Q;
      std::memset(&Q, 0, sizeof Q);
      Q.nA2 = g->nAxis2; Q.nPhi = g->nPhi; Q.axA = g->sinTheta; Q.axB = g->cosTheta; Q.sinPhi = g->sinPhi; Q.cosPhi = g->cosPhi;
      Q.dt = g->dt; Q.nTracks = t->nTracks; Q.x = t->x; Q.y = t->y; Q.z = t->z; Q.offsets = t->offsets; Q.itStart = t->itStart;
      // third candidate (fp64): the corrected-recurrence kernel, for mostly all-pass steps with phases beyond 2^18
      Plan pDrec;
      const bool haveDrec = g->dtype == SRB_DTYPE_F64 && make_plan(g, t, scratch_bytes - 64, false, &pDrec, false, true) == 0 &&
                            pDrec.kind == KIND_DREC;
      SRB_CUDA(cudaMemsetAsync(probeBuf, 0, 64, stream));
      k_probe<<<64, 128, 0, stream>>>(Q, t->totalSteps_host, g->omega_first_host, g->omega_last_host, probeBuf);
      k_decide<<<1, 1, 0, stream>>>(probeBuf, sel, (unsigned long long*)counters, haveDrec ? 1 : 0);
      SRB_CUDA(cudaGetLastError());
      launched += 2;
      // the candidates share the scratch body (only one of them runs): each one's launch sequence clears its own work
      // area, then the (conditional) pre-pass, kernel and reduction
      if (launch_plan(g, t, pPair, spectra, body, counters, stream, sel, pPair.kind, &launched) != 0) return -1;
      if (launch_plan(g, t, pRec, spectra, body, counters, stream, sel, pRec.kind, &launched) != 0) return -1;
      if (haveDrec && launch_plan(g, t, pDrec, spectra, body, counters, stream, sel, pDrec.kind, &launched) != 0) return -1;
      fill_info(pPair, launched, SRB_KIND_ON_DEVICE);
      return 0;
    }
  }
  Plan p;
  if (make_plan(g, t, scratch ? scratch_bytes : 0, false, &p, false) != 0) return -1;
  if (launch_plan(g, t, p, spectra, scratch, counters, stream, nullptr, 0, &launched) != 0) return -1;
  fill_info(p, launched, p.kind);
  return 0;
}

int srb_pipe_peak(int which, double* ops_per_second) {
  if (!ops_per_second) return fail("null result pointer");
  int dev = 0, numSM = 0;
  SRB_CUDA(cudaGetDevice(&dev));
  SRB_CUDA(cudaDeviceGetAttribute(&numSM, cudaDevAttrMultiProcessorCount, dev));
  void* buf = nullptr;
  SRB_CUDA(cudaMalloc(&buf, 64));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 40000, blocks = numSM * 2, threads = 256;
  float best = 1e30f;
  for (int rep = 0; rep < 4; rep++) {
    cudaEventRecord(e0);
    if (which == 0) k_pipe_peak<double, false><<<blocks, threads>>>((double*)buf, 1.0000001, 1e-9, iters);
    else if (which == 1) k_pipe_peak<float, false><<<blocks, threads>>>((float*)buf, 1.0000001f, 1e-9f, iters);
    else k_pipe_peak<float, true><<<blocks, threads>>>((float*)buf, 1.0f, 0.0f, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(buf);
  SRB_CUDA(cudaGetLastError());
  *ops_per_second = (double)blocks * threads * 8.0 * iters / (best * 1e-3);
  return 0;
}

int srb_last_launch(srb_launch_info* info) {
  if (!info) return fail("null info");
  *info = g_info;
  return 0;
}

int srb_swap_axes(const double* src, double* dst, uint32_t nSnaps, uint32_t nO, uint32_t nA, uint32_t nP, void* stream) {
  if (!src || !dst) return fail("null buffer");
  const size_t n = (size_t)nSnaps * nO * nA * nP;
  if (n == 0) return 0;
  int dev = 0, numSM = 0;
  SRB_CUDA(cudaGetDevice(&dev));
  SRB_CUDA(cudaDeviceGetAttribute(&numSM, cudaDevAttrMultiProcessorCount, dev));
  const int blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)numSM * 8);
  k_swap_axes<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, dst, nSnaps, nO, nA, nP);
  SRB_CUDA(cudaGetLastError());
  return 0;
}

int srb_energy_spectrum(int mode, int layout, const double* const* spectra, int n_spectra, int coherent, uint32_t nO, uint32_t nA,
                        uint32_t nP, uint32_t nSnaps, uint32_t iSnap, const double* axis2, double dphi, double* out,
                        void* stream) {
  if (mode != SRB_MODE_FAR && mode != SRB_MODE_NEAR) return fail("srb_energy_spectrum: bad mode");
  if (layout != 0 && layout != 1) return fail("srb_energy_spectrum: layout must be 0 (device) or 1 (swapped)");
  if (!spectra || n_spectra < 1 || n_spectra > 6) return fail("srb_energy_spectrum: 1..6 spectra expected");
  if (!axis2 || !out || nO == 0 || nA == 0 || nP == 0) return fail("srb_energy_spectrum: empty grid or null buffer");
  if (iSnap >= nSnaps) return fail("srb_energy_spectrum: snapshot index out of range");
  SpecPtrs S{};
  for (int k = 0; k < n_spectra; k++) { if (!spectra[k]) return fail("srb_energy_spectrum: null spectrum"); S.p[k] = spectra[k]; }
  k_energy_spectrum<<<(nO + 31) / 32, dim3(32, 8), 0, (cudaStream_t)stream>>>(S, n_spectra, coherent, mode == SRB_MODE_FAR,
                                                                               layout, nO, nA, nP, iSnap, axis2, dphi, out);
  SRB_CUDA(cudaGetLastError());
  return 0;
}

int srb_integrate_host(const srb_grid* g, const srb_tracks* t, double* const* spectra, int n_spectra,
                       uint64_t* counters_host, int device) {
  if (validate(g, t) != 0) return -1;
  SRB_CUDA(cudaSetDevice(device));
  const int nOut = srb_num_spectra(g->mode, g->comp);
  if (n_spectra != nOut || !spectra) return fail("n_spectra does not match comp");
  const size_t es = 8;   // tables and tracks are float64 for both dtypes
  std::vector<void*> owned;
  auto cleanup = [&]() { for (void* q : owned) cudaFree(q); };
  cudaError_t err = cudaSuccess;
  auto up = [&](const void* h, size_t bytes) -> void* {
    if (!h || bytes == 0 || err != cudaSuccess) return nullptr;
    void* d = nullptr;
    err = cudaMalloc(&d, bytes);
    if (err != cudaSuccess) return nullptr;
    owned.push_back(d);
    err = cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice);
    return d;
  };
  srb_grid gd = *g; srb_tracks td = *t;
  gd.omega = (const double*)up(g->omega, g->nOmega * es);
  gd.sinTheta = (const double*)up(g->sinTheta, g->nAxis2 * es); gd.cosTheta = (const double*)up(g->cosTheta, g->nAxis2 * es);
  gd.radius = (const double*)up(g->radius, g->nAxis2 * es);
  gd.sinPhi = (const double*)up(g->sinPhi, g->nPhi * es); gd.cosPhi = (const double*)up(g->cosPhi, g->nPhi * es);
  gd.formFactor = (const double*)up(g->formFactor, g->nOmega * es);
  const uint64_t total = t->totalSteps_host;
  td.x = (const double*)up(t->x, total * es); td.y = (const double*)up(t->y, total * es); td.z = (const double*)up(t->z, total * es);
  td.ux = (const double*)up(t->ux, total * es); td.uy = (const double*)up(t->uy, total * es); td.uz = (const double*)up(t->uz, total * es);
  td.offsets = (const uint64_t*)up(t->offsets, (size_t)(t->nTracks + 1) * 8);
  td.w = (const double*)up(t->w, t->nTracks * es);
  td.itStart = (const uint32_t*)up(t->itStart, (size_t)t->nTracks * 4);
  td.itEnd = (const uint32_t*)up(t->itEnd, (size_t)t->nTracks * 4);
  td.itSnaps = (const uint32_t*)up(t->itSnaps, (size_t)(t->itSnapsStride ? (size_t)t->nTracks * g->nSnaps : g->nSnaps) * 4);
  if (err != cudaSuccess) { cleanup(); return fail(std::string("upload: ") + cudaGetErrorString(err)); }
  const size_t per = (size_t)g->nSnaps * g->nOmega * g->nAxis2 * g->nPhi;
  std::vector<double*> dsp(nOut);
  for (int c = 0; c < nOut; c++) {
    void* d = nullptr; err = cudaMalloc(&d, per * 8);
    if (err == cudaSuccess) { owned.push_back(d); err = cudaMemset(d, 0, per * 8); }
    if (err != cudaSuccess) { cleanup(); return fail(std::string("spectrum alloc: ") + cudaGetErrorString(err)); }
    dsp[c] = (double*)d;
  }
  size_t sb = t->nTracks ? srb_scratch_bytes(&gd, &td) : 0;
  void* scratch = nullptr;
  if (sb) { if (cudaMalloc(&scratch, sb) == cudaSuccess) owned.push_back(scratch); else { scratch = nullptr; sb = 0; cudaGetLastError(); } }
  uint64_t* dcnt = nullptr;
  if (counters_host) { if (cudaMalloc((void**)&dcnt, 32) == cudaSuccess) owned.push_back(dcnt); else dcnt = nullptr; }
  int rc = srb_integrate(&gd, &td, dsp.data(), nOut, scratch, sb, dcnt, nullptr);
  if (rc == 0) {
    std::vector<double> h(per);
    for (int c = 0; c < nOut && err == cudaSuccess; c++) {
      err = cudaMemcpy(h.data(), dsp[c], per * 8, cudaMemcpyDeviceToHost);
      if (err == cudaSuccess) for (size_t i = 0; i < per; i++) spectra[c][i] += h[i];
    }
    if (err == cudaSuccess && dcnt) err = cudaMemcpy(counters_host, dcnt, 16, cudaMemcpyDeviceToHost);
    if (err != cudaSuccess) rc = fail(std::string("download: ") + cudaGetErrorString(err));
  }
  cleanup();
  return rc;
}

}  // extern "C"
