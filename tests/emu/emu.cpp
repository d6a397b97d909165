// TEST INFRASTRUCTURE ONLY — single-warp CPU emulation of the device code in
// synchrad_b200/csrc/srb_core.cuh (lanes run as a loop, shared memory is a plain struct).
// It exists so the kernel LOGIC (flush chains, pass ranges, seeds, recurrences, tile indexing)
// can be debugged against the oracle in a container without a GPU.  It is not a fallback: the
// product package never loads it and fails loudly without the CUDA library.
// Build: g++ -O2 -mfma -ffp-contract=off -fopenmp -shared -fPIC -o tests/emu/libsrb_emu.so tests/emu/emu.cpp
#include <cstring>
#include <vector>

#include "../../include/synchrad_b200.h"
#include "../../synchrad_b200/csrc/srb_literal.cuh"
#include "../../synchrad_b200/csrc/srb_ws.cuh"

using namespace srb;

template <class C>
static void run_all(const Params& P0, unsigned long long* counters) {
  unsigned long long tot[2] = {0, 0};
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : tot[:2])
  for (long long task = 0; task < (long long)P0.nVD * P0.nPC; task++) {
    Params P = P0;
    unsigned long long cnt[2] = {0, 0};
    P.counters = cnt;
    const uint32_t vd = (uint32_t)(task % P0.nVD), pc = (uint32_t)(task / P0.nVD);
    WarpSmem<C> sm;
    ThreadState<C> st[32];
    std::memset(&sm, 0, sizeof sm);
    warp_task<C>(P, vd, pc, sm, st);
    tot[0] += cnt[0]; tot[1] += cnt[1];
  }
  if (counters) { counters[0] = tot[0]; counters[1] = tot[1]; }
}

// warp-specialised pair kernel (srb_ws.cuh): producer and consumer of a unit run in sequence per item
template <class C>
static void run_all_ws(const Params& P0, unsigned long long* counters) {
  unsigned long long tot[2] = {0, 0};
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : tot[:2])
  for (long long task = 0; task < (long long)P0.nVD * P0.nPC; task++) {
    Params P = P0;
    unsigned long long cnt[2] = {0, 0};
    P.counters = cnt;
    ws_emulate_task<C>(P, (uint32_t)(task % P0.nVD), (uint32_t)(task / P0.nVD));
    tot[0] += cnt[0]; tot[1] += cnt[1];
  }
  if (counters) { counters[0] = tot[0]; counters[1] = tot[1]; }
}

// exposes the device sincos for accuracy tests
extern "C" void srb_emu_sincos(const double* x, double* s, double* c, long n) {
  for (long i = 0; i < n; i++) sincos_big(x[i], s + i, c + i);
}

extern "C" int srb_emu_integrate(const srb_grid* g, const srb_tracks* t, double* const* spectra, int nOut,
                                 int kind, int tw, uint32_t nPC, unsigned long long* counters, int prepass, uint32_t nTS) {
  Params P;
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
  if (g->dtype == SRB_DTYPE_F32_LITERAL) kind = KIND_LITERAL;
  const bool ws = kind == 6;          // 'pair_ws': the warp-specialised form of the fp64 pair kernel (needs the pre-pass)
  const int tiles = kind == KIND_RECUR ? 16 : 32;
  P.chunkNodes = (uint32_t)(tiles * tw);
  P.nChunks = (g->nOmega + P.chunkNodes - 1) / P.chunkNodes;
  P.nVD = g->nPhi * g->nAxis2 * P.nChunks;
  P.nTracks = t->nTracks;
  P.x = t->x; P.y = t->y; P.z = t->z; P.ux = t->ux; P.uy = t->uy; P.uz = t->uz;
  P.offsets = t->offsets; P.w = t->w; P.itStart = t->itStart; P.itEnd = t->itEnd; P.itSnaps = t->itSnaps;
  P.snapStride = t->itSnapsStride;
  for (int c = 0; c < nOut; c++) P.out[c] = spectra[c];
  const size_t perOut = (size_t)g->nSnaps * g->nOmega * g->nAxis2 * g->nPhi;
  std::vector<double> slabs((size_t)(nPC > 1 ? nPC - 1 : 0) * perOut * nOut, 0.0);
  P.slabs = slabs.data(); P.slabStride = perOut * nOut; P.nPC = nPC;
  // time-axis split (nTS > 1): one (track, segment) per chunk, partial amplitudes combined afterwards
  std::vector<double> amp;
  P.nTS = nTS > 1 ? nTS : 1;
  if (nTS > 1) {
    if (g->dtype == SRB_DTYPE_F32_LITERAL) return -2;
    P.nPC = nPC = t->nTracks * nTS;
    amp.assign((size_t)nPC * 6 * perOut, 0.0);
    P.amp = amp.data();
  }
  // optional pre-pass planes (same code path the GPU pre-pass kernel uses)
  std::vector<double> pre;
  if (prepass && t->nTracks) {
    const uint64_t total = t->totalSteps_host;
    pre.assign((size_t)(g->mode == SRB_MODE_FAR ? 6 : 3) * total, 0.0);
    const double dtInv = sdiv(1.0, g->dt);
    for (uint32_t tr = 0; tr < t->nTracks; tr++) {
      const uint64_t o = t->offsets[tr], n = t->offsets[tr + 1] - o;
      for (uint64_t it = 0; it < n; it++) {
        const uint64_t i = o + it;
        if (g->mode == SRB_MODE_FAR) {
          double av[3] = {0, 0, 0}, bv[3] = {0, 0, 0};
          if (it + 1 < n) far_step_kinematics<double>(t->ux, t->uy, t->uz, i, dtInv, av, bv);
          for (int c = 0; c < 3; c++) { pre[c * total + i] = av[c]; pre[(3 + c) * total + i] = bv[c]; }
        } else {
          double bv[3];
          near_step_kinematics<double>(t->ux, t->uy, t->uz, i, bv);
          for (int c = 0; c < 3; c++) pre[c * total + i] = bv[c];
        }
      }
    }
    P.pre = pre.data(); P.preStride = total;
  }
  const bool f32 = g->dtype == SRB_DTYPE_F32;
  if (kind == KIND_LITERAL) { P.pre = nullptr; }
  bool ok = false;
  const bool spheric = g->comp == SRB_COMP_SPHERIC || g->comp == SRB_COMP_SPHERIC_COMPLEX;
  const int nc = (kind == KIND_RECUR && g->mode == SRB_MODE_FAR && !spheric) ? 2 : 3;
#define EMU_CASE1(K, M, TWV, NCV)                                                               \
  if (kind == K && g->mode == M && tw == TWV && nc == NCV) {                                    \
    if (f32) run_all<Cfg<double, float, M, K, TWV, false, NCV>>(P, counters);                   \
    else run_all<Cfg<double, double, M, K, TWV, false, NCV>>(P, counters);                      \
    ok = true; }
#define EMU_CASE(K, M, TWV) EMU_CASE1(K, M, TWV, 2) EMU_CASE1(K, M, TWV, 3)
  EMU_CASE(KIND_RECUR, MODE_FAR, 16) EMU_CASE(KIND_RECUR, MODE_FAR, 8) EMU_CASE(KIND_RECUR, MODE_FAR, 4)
  EMU_CASE(KIND_RECUR, MODE_NEAR, 8) EMU_CASE(KIND_RECUR, MODE_NEAR, 4) EMU_CASE(KIND_RECUR, MODE_NEAR, 2)
  EMU_CASE(KIND_DIRECT, MODE_FAR, 8) EMU_CASE(KIND_DIRECT, MODE_FAR, 4) EMU_CASE(KIND_DIRECT, MODE_FAR, 2)
  EMU_CASE(KIND_DIRECT, MODE_NEAR, 8) EMU_CASE(KIND_DIRECT, MODE_NEAR, 4) EMU_CASE(KIND_DIRECT, MODE_NEAR, 2)
#define EMU_DREC(M, TWV) if (kind == KIND_DREC && !f32 && g->mode == M && tw == TWV) { run_all<Cfg<double, double, M, KIND_DREC, TWV, false, 3>>(P, counters); ok = true; }
  EMU_DREC(MODE_FAR, 8) EMU_DREC(MODE_FAR, 4) EMU_DREC(MODE_FAR, 2) EMU_DREC(MODE_NEAR, 8) EMU_DREC(MODE_NEAR, 4) EMU_DREC(MODE_NEAR, 2)
#undef EMU_DREC
#undef EMU_CASE
#undef EMU_CASE1
#define EMU_PAIR3(TWV) if (kind == KIND_PAIR && g->mode == MODE_FAR && tw == TWV && spheric) { if (f32) run_all<Cfg<double, float, MODE_FAR, KIND_PAIR, TWV, false, 3>>(P, counters); else run_all<Cfg<double, double, MODE_FAR, KIND_PAIR, TWV, false, 3>>(P, counters); ok = true; }
  EMU_PAIR3(8) EMU_PAIR3(4) EMU_PAIR3(2)
#undef EMU_PAIR3
#define EMU_PAIR(TWV) if (kind == KIND_PAIR && g->mode == MODE_FAR && tw == TWV && !spheric) { if (f32) run_all<Cfg<double, float, MODE_FAR, KIND_PAIR, TWV, false, 2>>(P, counters); else run_all<Cfg<double, double, MODE_FAR, KIND_PAIR, TWV, false, 2>>(P, counters); ok = true; }
  EMU_PAIR(16) EMU_PAIR(8) EMU_PAIR(4) EMU_PAIR(2)
#undef EMU_PAIR
  // scalar-pipe form of the fp64 pair kernel where KIND_PAIR takes the tensor-core layout
  if (kind == KIND_PAIR_FMA && g->mode == MODE_FAR && !f32) {
    if (tw == 8 && !spheric) { run_all<Cfg<double, double, MODE_FAR, KIND_PAIR_FMA, 8, false, 2>>(P, counters); ok = true; }
    if (tw == 4 && !spheric) { run_all<Cfg<double, double, MODE_FAR, KIND_PAIR_FMA, 4, false, 2>>(P, counters); ok = true; }
    if (tw == 8 && spheric) { run_all<Cfg<double, double, MODE_FAR, KIND_PAIR_FMA, 8, false, 3>>(P, counters); ok = true; }
  }
#define EMU_LIT(M, TWV) if (kind == KIND_LITERAL && g->mode == M && tw == TWV) { run_all<Cfg<double, float, M, KIND_LITERAL, TWV, false, 3>>(P, counters); ok = true; }
  EMU_LIT(MODE_FAR, 8) EMU_LIT(MODE_FAR, 4) EMU_LIT(MODE_FAR, 2) EMU_LIT(MODE_NEAR, 8) EMU_LIT(MODE_NEAR, 4) EMU_LIT(MODE_NEAR, 2)
#undef EMU_LIT
  if (ws) {
    ok = false;
    if (g->mode == MODE_FAR && !f32 && P.pre) {
      if (tw == 8 && !spheric) { run_all_ws<Cfg<double, double, MODE_FAR, KIND_PAIR, 8, false, 2>>(P, counters); ok = true; }

      if (tw == 4 && !spheric) { run_all_ws<Cfg<double, double, MODE_FAR, KIND_PAIR, 4, false, 2>>(P, counters); ok = true; }

    }
  }
  if (!ok) return -1;
  if (nTS > 1) {
    const bool spheric_ = g->comp == SRB_COMP_SPHERIC || g->comp == SRB_COMP_SPHERIC_COMPLEX;
    const int ncf = g->mode == SRB_MODE_NEAR ? 3 : (((kind == KIND_RECUR || kind == KIND_PAIR || kind == KIND_PAIR_FMA || ws) && !spheric_) ? 2 : 3);
    P.counters = nullptr;
    for (size_t i = 0; i < perOut; i++) {
      if (g->mode == SRB_MODE_NEAR) combine_node<MODE_NEAR, 3>(P, i);
      else if (ncf == 2) combine_node<MODE_FAR, 2>(P, i);
      else combine_node<MODE_FAR, 3>(P, i);
    }
    return 0;
  }
  for (uint32_t s = 0; s + 1 < nPC; s++)
    for (int c = 0; c < nOut; c++)
      for (size_t i = 0; i < perOut; i++) P.out[c][i] += slabs[(size_t)s * P.slabStride + (size_t)c * perOut + i];
  return 0;
}

static_assert(sizeof(WarpSmem<Cfg<double, double, MODE_FAR, KIND_PAIR, 8, false, 2>>) == 32 * 20 * 8 + 128 + 24 * 33 * 8, "other kinds keep their footprint");
