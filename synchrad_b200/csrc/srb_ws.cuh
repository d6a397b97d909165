// synchrad_b200 — warp-specialised form of the fp64 symmetric-pair kernel (the Cfg::MMA configurations of
// srb_pair.cuh): the headline far-field kernel on uniform omega grids.
//
// Round 1 ran prep phase -> main phase strictly in sequence inside each warp, so the latency-bound scalar chains of
// the prep phase (lane = time step) sat on the critical path of the FP64-tensor-core stream (ncu: half of the warp
// time, FP64 units 78 % busy).  Here a block is NU "units" = NU virtual directions, and a unit is
//     1 consumer warp  — nothing but the DMMA.8x8x4 stream  U[64 x 8NT] += X[64 x 32] * Q'[32 x 8NT]  per 32-step
//                        sub-batch (plus the rare lane-by-lane path for steps that pass the Nyquist guard partially and
//                        the flush once per track and snapshot), accumulators in registers;
//     NP producer warps — the per-(direction, step) work, lane = time step: tau = t - n.r in the reference's exact
//                        operation order, the Nyquist pass range, the Lienard-Wiechert amplitude, the tile phasors
//                        X_m = exp(i(phi_c + m*delta)) and the pair phasors Q'_pc = A_c exp(i(16+32p)delta), written
//                        in MMA-fragment-friendly layout into a ring of NS shared-memory stages;
//   hand-off by mbarriers (full[stage]: 32 producer lanes arrive; empty[stage]: 32 consumer lanes arrive), no block-wide
//   barrier after start-up.  Producers feed themselves with TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx)
//   of their next sub-batch's input records (x, y, z, acceleration, mean beta of 34 steps: one 2448-byte copy from
//   the packed records the pre-pass kernel writes), so the global-memory latency is off the chain as well (BASELINE
//   north_star: "track chunks staged into shared memory with TMA").
// A flush is an item of the same ring: the producer hands over an empty stage, which the consumer uses as scratch
// for the fragment-layout <-> tile-layout transposes around the common flush code.
//
// Versus round 1 the tile phasors X_m moved from the consumer (Y_a * Z_b, 128 FP64 ops per lane and sub-batch) into
// the producer (lane = step: 16 ops for tiles 0..7); the consumer forms tiles 8..31 by three rotations with the staged
// R^8, R^16 (96 ops per lane and sub-batch), so that a stage is 11 KB and the ring can be FOUR deep: with one stage per
// producer the cycle per stage is produce -> consume -> produce and the period (Tp + Tc) / 2 (measured: 4100 cycles
// per sub-batch against 2300 for the DMMA stream alone, profiles/r02_ws_tuning.md); a ring of NS stages fed by NP
// producers gives max(Tc, Tp / NP, (Tp + Tc) / NS), and Tp (a latency-bound chain sharing the FP64 pipe with the DMMA
// stream: ~35 cycles per dependent op) is 6-8 thousand cycles.
//
// The per-sub-batch functions below are shared with the CPU emulation (tests/emu: lanes as a loop, producer then
// consumer in sequence); mbarriers, TMA and the role split are GPU-only (srb_api.cu: k_integrate_ws).
#pragma once
#include "srb_pair.cuh"

namespace srb {

constexpr int WS_XS = 36;   // (Re, Im) pairs per row of the stage arrays: lane = step 128-bit stores and the fragment
                            // loads of the consumer (128-bit X / W, 64-bit halves of Q') are all bank-conflict free
struct alignas(16) Dbl2 { double x, y; };
constexpr int WS_IN = 34;   // staged input records per sub-batch (32 steps + the previous step + 16-byte alignment slack)
constexpr int WS_NIN = 9;   // doubles per input record: x, y, z, a0..a2, b0..b2 (written by the pre-pass kernel)

// per-unit constants of the omega chunk, kept in registers by the producers
struct WsConst { double wLo, wHi, wCen; };
template <class C>
SRB_HD void ws_const(const Params& P, const Geom& g, WsConst& k) {
  using TI = typename C::TI;
  constexpr uint32_t CEN = 16u * (uint32_t)(C::TW - 1);
  k.wLo = (double)((const TI*)P.omega)[g.cLo];
  k.wHi = (double)((const TI*)P.omega)[g.cHi - 1];
  // frequency of tile 0's centre: node cLo + 16(TW-1) of the table when the chunk has that many nodes (the exact
  // rounded phase the reference forms there), else the same frequency extrapolated on the uniform grid
  k.wCen = g.cLo + CEN < P.nOmega ? (double)((const TI*)P.omega)[g.cLo + CEN] : k.wLo + (double)CEN * P.domega;
}

template <class C>
struct alignas(16) WsStage {
  Dbl2 X[8][WS_XS];                 // [tile][step] = (Re, Im) X, tiles 0..7; tiles 8.. = X_b * R^8 / R^16 / R^24 (consumer)
  Dbl2 W[2][WS_XS];                 // R^8 and R^16: [power][step]
  Dbl2 Q[4 * C::NT][WS_XS];         // [pair*NC + comp][step] = A_comp * (cos, sin) of the pair offset
  Dbl2 rec[SUB][2];                 // (A0, A1), (A2, tau) of the step, for the lane-by-lane path
  uint32_t rng[SUB];                // lo | hi<<10 | flag<<30 (chunk-relative pass range)
  uint32_t cnt, fullMask, anyMask, pad;
  // a flush item hands the stage to the consumer as scratch for 32 * NACC doubles (16-node tiles need a little more)
  static constexpr int BASE_DOUBLES = 2 * (8 + 2 + 4 * C::NT) * WS_XS + SUB * 4 + SUB / 2 + 2;
  static constexpr int PAD_DOUBLES = 32 * C::NACC > BASE_DOUBLES ? 32 * C::NACC - BASE_DOUBLES : 2;
  double flushPad[PAD_DOUBLES];
};

struct WalkItem {
  uint32_t kind;      // 0: sub-batch of up to 32 steps, 1: flush of snapshot iSnap
  uint32_t base;      // first loop index `it` of the sub-batch
  uint32_t iSnap;
  int cnt;
  bool newTrack;      // first item of its track: accumulators restart from zero
};

// The loop nest of warp_task (tracks -> snapshot intervals -> sub-batches -> flush; kernel_farfield.cl:54-63,100-106
// with the closed-form flush chain) as a warp-uniform iterator, so that producers and consumer walk the
// same item sequence independently.
template <class C>
struct Walk {
  uint32_t t, t1, pc;
  TrackView tv;       // track of the item returned last
  uint32_t loopEnd, nComp, iSnap, cur, stop, base, afterFlush, segLo, segHi;
  int state;          // 0: next track, 1: inside a snapshot interval, 3: open the next interval
  bool fresh;
  SRB_HD void init(uint32_t t0_, uint32_t t1_, uint32_t pc_) { t = t0_; t1 = t1_; pc = pc_; state = 0; fresh = false; }
  SRB_HD bool next(const Params& P, WalkItem& it) {
    for (;;) {
      if (state == 0) {
        if (t >= t1) return false;
        load_track<C>(P, t, tv);
        loopEnd = tv.itEnd > 0 ? tv.itEnd - 1 : 0;
        nComp = tv.n > 0 ? (tv.n - 1 < loopEnd ? tv.n - 1 : loopEnd) : 0;
        iSnap = 0;
        while (iSnap < P.nSnaps && !(tv.itStart < tv.snaps[iSnap])) iSnap++;   // :54-57
        seg_range(P, pc, nComp, segLo, segHi);     // time-axis split: this chunk's share of the steps
        cur = 0; fresh = true; state = 3;
      }
      if (state == 3) {
        if (iSnap >= P.nSnaps) { t++; state = 0; continue; }
        // the flush test `it_glob + 2 == itSnaps[iSnap]` (:100) fires at it = itf, if reachable
        const long long itf = (long long)tv.snaps[iSnap] - 2 - (long long)tv.itStart;
        if (itf < (long long)cur || itf >= (long long)loopEnd) { t++; state = 0; continue; }   // never fires again (Q3/Q4)
        stop = (uint32_t)(itf + 1) < nComp ? (uint32_t)(itf + 1) : nComp;
        if (stop > segHi) stop = segHi;
        base = cur > segLo ? cur : segLo; afterFlush = (uint32_t)(itf + 1);
        state = 1;
      }
      it.iSnap = iSnap; it.newTrack = fresh; fresh = false;
      if (base < stop) {
        it.kind = 0; it.base = base;
        it.cnt = (int)(stop - base < (uint32_t)SUB ? stop - base : (uint32_t)SUB);
        base += SUB;
        return true;
      }
      it.kind = 1; it.base = 0; it.cnt = 0;
      cur = afterFlush; iSnap++; state = 3;
      return true;
    }
  }
};

// ------------------------------------------------------------------------------------------------ producer side
// Tile and pair phasors of one step (lane = step s) into the stage.
//   X_m = exp(i(phi_c + m*d)), m = 0..7: phi_c = the reference's own rounded phase at the centre node of tile 0,
//   d = domega*tau, by the three-term recurrence x[k+1] = 2cos(d) x[k] - x[k-1] (6 steps: error <= 8 ulp * |cot d|,
//   as the recurrence kernel's tiles).  Tiles 8..31 are X_b R^8, X_b R^16, (X_b R^8) R^16, formed by the consumer
//   from the staged R^8, R^16 (the full 64 x 32 operand does not leave room for a ring deep enough to cover the
//   producers' latency).  Q'_pc = A_c R^(16+32p).
template <class C>
SRB_HD void ws_seeds(const Params& P, const WsConst& kc, double tau, const double* A, WsStage<C>& sg, int s) {
  constexpr int TW = C::TW, NC = C::NC;
  double er, ei, sd, cd;
  sincos_big(smul(kc.wCen, tau), &ei, &er);
  sincos_big(P.domega * tau, &sd, &cd);
  const double cf = 2.0 * cd;
  double r2r = cd * cd - sd * sd, r2i = 2.0 * cd * sd;               // R^2
  double r4r = r2r * r2r - r2i * r2i, r4i = 2.0 * r2r * r2i;         // R^4
  const double r8r = r4r * r4r - r4i * r4i, r8i = 2.0 * r4r * r4i;   // R^8
  {
    double x0r = er, x0i = ei;
    double x1r = er * cd - ei * sd, x1i = er * sd + ei * cd;
    sg.X[0][s] = Dbl2{x0r, x0i}; sg.X[1][s] = Dbl2{x1r, x1i};
#pragma unroll
    for (int k = 2; k < 8; k++) {
      const double x2r = fma(cf, x1r, -x0r), x2i = fma(cf, x1i, -x0i);
      sg.X[k][s] = Dbl2{x2r, x2i};
      x0r = x1r; x0i = x1i; x1r = x2r; x1i = x2i;
    }
  }
  double qr = r8r * r8r - r8i * r8i, qi = 2.0 * r8r * r8i;           // R^16
  sg.W[0][s] = Dbl2{r8r, r8i}; sg.W[1][s] = Dbl2{qr, qi};
  const double wr = qr * qr - qi * qi, wi = 2.0 * qr * qi;           // R^32
#pragma unroll
  for (int p = 0; p < TW / 2; p++) {
#pragma unroll
    for (int c = 0; c < NC; c++) {
      sg.Q[p * NC + c][s] = Dbl2{A[c] * qr, A[c] * qi};
    }
    if (p + 1 < TW / 2) { const double t = qr * wr - qi * wi; qi = qr * wi + qi * wr; qr = t; }
  }
}

// Amplitude in the transverse basis with fewer FP64 ops than far_amplitude (47 -> 33): the denominator 1 - b.n, where
// the cancellation is (relative error ~ gamma^2 ulp, it has to match the reference's rounding), stays in the
// reference's operation order; its reciprocal by MUFU + two Newton steps (<= 1 ulp from the correctly rounded
// quotient, a RELATIVE error of 1e-16 on the amplitude) and the two projections with fused multiply-adds.
SRB_HD void far_amplitude_tr(const Geom& g, const double a[3], const double b[3], double A[3]) {
  double c1 = sdot3(a[0], a[1], a[2], g.nx, g.ny, g.nz);
  const double d = ssub(1.0, sdot3(b[0], b[1], b[2], g.nx, g.ny, g.nz));
#if defined(__CUDA_ARCH__)
  double c2;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(c2) : "d"(d));
  c2 = fma(fma(-d, c2, 1.0), c2, c2);
  c2 = fma(fma(-d, c2, 1.0), c2, c2);
#else
  const double c2 = 1.0 / d;
#endif
  c1 = smul(smul(c1, c2), c2);
  const double A0 = ssub(smul(c1, ssub(g.nx, b[0])), smul(c2, a[0]));
  const double A1 = ssub(smul(c1, ssub(g.ny, b[1])), smul(c2, a[1]));
  const double A2 = ssub(smul(c1, ssub(g.nz, b[2])), smul(c2, a[2]));
  A[0] = fma(g.tx, A0, fma(g.ty, A1, g.tz * A2));
  A[1] = fma(g.px, A0, g.py * A1);          // e_phi = (-sin phi, cos phi, 0)
  A[2] = 0.0;
}

// One step of the producer's sub-batch (lane = step), in two parts.  tau / tauPrev come from the caller (it owns the
// inputs: staged by TMA on the GPU); a, b: the step's acceleration and mean beta from the pre-pass.
//   part 1 (ws_prep_guard): Nyquist pass range and the amplitude vector -- consumes every staged input, so the caller
//           may recycle its input buffer afterwards;  part 2 (ws_prep_store): phasors and the stores into the stage.
struct WsStep { uint32_t lo, hi, flag; double tau, A[3]; };

template <class C>
SRB_HD void ws_prep_guard(const Params& P, const Geom& g, const WsConst& kc, bool active, double tau, double tauPrev,
                          const double* a, const double* b, WsStep& w, unsigned long long& nPass, unsigned long long& nAll) {
  w.lo = w.hi = w.flag = 0u; w.tau = tau; w.A[0] = w.A[1] = w.A[2] = 0.0;
  if (active) {
    const uint32_t n = g.cHi - g.cLo;
    // Nyquist guard on the reference's rounded predicate (kernel_farfield.cl:68-72), monotone in omega: the chunk's
    // largest and smallest node decide the common cases, pass_range settles a cut-off inside the chunk
    if (fabs(ssub(smul(kc.wHi, tau), smul(kc.wHi, tauPrev))) < 3.14159265358979323846) { w.hi = n; w.flag = 1u; }
    else if (fabs(ssub(smul(kc.wLo, tau), smul(kc.wLo, tauPrev))) < 3.14159265358979323846) {
      pass_range<C>(P, g, tau, tauPrev, w.lo, w.hi);
      w.flag = (w.hi <= w.lo) ? 0u : ((w.lo == 0 && w.hi == n) ? 1u : 2u);
    }
    nAll += n;
    if (w.flag) {
      if (C::NC == 2) far_amplitude_tr(g, a, b, w.A); else far_amplitude<C>(P, g, a, b, w.A);
      // beyond |phase| = 2^18 the seed arithmetic cannot track the reference's rounded phase to 1e-9: node by node
      if (fabs(kc.wHi * tau) > 262144.0) w.flag = 3u;
      nPass += w.hi - w.lo;
    }
  }
}

template <class C>
SRB_HD uint32_t ws_prep_store(const Params& P, const WsConst& kc, const WsStep& w, WsStage<C>& sg, int lane) {
  if (w.flag) {
    if (w.flag != 3u) ws_seeds<C>(P, kc, w.tau, w.A, sg, lane);
    sg.rec[lane][0] = Dbl2{w.A[0], w.A[1]}; sg.rec[lane][1] = Dbl2{w.A[2], w.tau};
  }
  sg.rng[lane] = w.lo | (w.hi << 10) | (w.flag << 30);
  return w.flag;
}

// ------------------------------------------------------------------------------------------------ consumer side
// steps that pass the guard partially or carry a huge phase (flag 3): one lane = the (tile, q) slots it owns in the
// MMA accumulator layout; same half-weight arithmetic as pair_update (exact)
template <class C>
SRB_HD void ws_partial(const Params& P, const Geom& g, const WsStage<C>& sg, uint32_t rest, int lane, ThreadState<C>& st) {
  using TI = typename C::TI;
  constexpr int NC = C::NC, NP = C::TW / 2, NT = C::NT;
  const int ks = lane & 3, b = lane >> 2;
  for (int s = 0; s < 32; s++) {
    if (!((rest >> s) & 1u)) continue;
    const uint32_t r = sg.rng[s];
    const uint32_t flag = r >> 30;
    const int hiN = (int)((r >> 10) & 0x3ffu);          // passing chunk-relative nodes: [0, hiN)
    const double tau = sg.rec[s][1].y;                  // flag 3 only
#pragma unroll
    for (int a = 0; a < 4; a++) {
      const int m = 8 * a + b;
      double xr = 0, xi = 0;
      if (flag != 3) {
        xr = sg.X[b][s].x; xi = sg.X[b][s].y;
        if (a & 1) { const double wr = sg.W[0][s].x, wi = sg.W[0][s].y; const double t = fma(xr, wr, -(xi * wi)); xi = fma(xr, wi, xi * wr); xr = t; }
        if (a & 2) { const double wr = sg.W[1][s].x, wi = sg.W[1][s].y; const double t = fma(xr, wr, -(xi * wi)); xi = fma(xr, wi, xi * wr); xr = t; }
      }
#pragma unroll
      for (int t = 0; t < NT; t++) {
        const int q = 4 * t + ks, p = q / NC, c = q - p * NC;
        const int km = NP - 1 - p, kp = NP + p;         // tile-local indices of the (-) and (+) node
        if (m + 32 * km >= hiN) continue;               // (-) fails, so does (+)
        const bool pp = m + 32 * kp < hiN;
        double cp = 0, sp = 0, cm, sm_;
        if (flag == 3) {
          const double A = c == 0 ? sg.rec[s][0].x : (c == 1 ? sg.rec[s][0].y : sg.rec[s][1].x);
          const uint32_t jb = g.cLo + (uint32_t)m;
          sincos_big(smul((double)((const TI*)P.omega)[jb + 32 * km], tau), &sm_, &cm);
          if (pp) sincos_big(smul((double)((const TI*)P.omega)[jb + 32 * kp], tau), &sp, &cp);
          cp *= A; sp *= A; cm *= A; sm_ *= A;
        } else {
          const double qc = sg.Q[q][s].x, qs = sg.Q[q][s].y;
          if (pp) { cp = fma(xr, qc, -(xi * qs)); sp = fma(xr, qs, xi * qc); }   // X * A Q
          cm = fma(xr, qc, xi * qs); sm_ = fma(xi, qc, -(xr * qs));              // X * A conj(Q)
        }
        st.acc[mma_acc_index<C>(a, t, 0)] += 0.5 * (cp + cm);
        st.acc[mma_acc_index<C>(a, t, 1)] += 0.5 * (cm - cp);
        st.acc[mma_acc_index<C>(a, t, 2)] += 0.5 * (sp - sm_);
        st.acc[mma_acc_index<C>(a, t, 3)] += 0.5 * (sp + sm_);
      }
    }
  }
}

#if defined(__CUDA_ARCH__)
// fragments of one k-group (4 steps): the lane's B operands (masked to 0 for steps that are not all-pass), its X of
// tiles b and 8 + b, and W of its step
template <class C>
struct WsFrag { double bq[C::NT], xr, xi, ur, ui, wr, wi; };
template <class C>
SRB_HD void ws_load_frag_k(const WsStage<C>& sg, int j, int ks, int b, uint32_t fullMask, WsFrag<C>& f) {
  const int s = 4 * j + ks;
  const bool on = (fullMask >> s) & 1u;
#pragma unroll
  for (int t = 0; t < C::NT; t++) {       // column 8t + b of Q' = (cos|sin)[b & 1] of q = 4t + b/2
    const Dbl2& qq = sg.Q[4 * t + (b >> 1)][s];
    const double v = (b & 1) ? qq.y : qq.x;
    f.bq[t] = on ? v : 0.0;
  }
  const Dbl2 xx = sg.X[b][s], uu = sg.W[0][s], ww = sg.W[1][s];
  f.xr = xx.x; f.xi = xx.y; f.ur = uu.x; f.ui = uu.y; f.wr = ww.x; f.wi = ww.y;
}
template <class C>
SRB_HD void ws_mma_k(const WsFrag<C>& f, ThreadState<C>& st) {
  constexpr int NT = C::NT;
  double xr[4], xi[4];
  xr[0] = f.xr; xi[0] = f.xi;                                                            // tile b
  xr[1] = fma(f.xr, f.ur, -(f.xi * f.ui)); xi[1] = fma(f.xr, f.ui, f.xi * f.ur);         // tile 8 + b
  xr[2] = fma(f.xr, f.wr, -(f.xi * f.wi)); xi[2] = fma(f.xr, f.wi, f.xi * f.wr);         // tile 16 + b
  xr[3] = fma(xr[1], f.wr, -(xi[1] * f.wi)); xi[3] = fma(xr[1], f.wi, xi[1] * f.wr);     // tile 24 + b
#pragma unroll
  for (int a = 0; a < 4; a++) {
#pragma unroll
    for (int t = 0; t < NT; t++) {
      dmma884(st.acc[((a * 2 + 0) * NT + t) * 2], st.acc[((a * 2 + 0) * NT + t) * 2 + 1], xr[a], f.bq[t]);
      dmma884(st.acc[((a * 2 + 1) * NT + t) * 2], st.acc[((a * 2 + 1) * NT + t) * 2 + 1], xi[a], f.bq[t]);
    }
  }
}
// one sub-batch on the tensor cores: 8 k-groups x 4 a x 2 (Re|Im X) x NT DMMA.8x8x4, operands straight from the stage;
// the fragments of k-group j + 1 are fetched before the DMMA burst of k-group j is issued
template <class C>
SRB_HD void ws_main(const Params& P, const Geom& g, const WsStage<C>& sg, int lane, ThreadState<C>& st) {
  const int ks = lane & 3, b = lane >> 2;
  const uint32_t fullMask = sg.fullMask, anyMask = sg.anyMask;
  if (fullMask == 0xffffffffu) {
    WsFrag<C> f0, f1;
    ws_load_frag_k<C>(sg, 0, ks, b, fullMask, f0);
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      ws_load_frag_k<C>(sg, j + 1, ks, b, fullMask, f1);
      ws_mma_k<C>(f0, st);
      if (j + 2 < 8) ws_load_frag_k<C>(sg, j + 2, ks, b, fullMask, f0);
      ws_mma_k<C>(f1, st);
    }
  } else if (fullMask) {
    for (int j = 0; j < 8; j++) {
      if (!((fullMask >> (4 * j)) & 0xfu)) continue;     // warp-uniform
      WsFrag<C> f;
      ws_load_frag_k<C>(sg, j, ks, b, fullMask, f);
      ws_mma_k<C>(f, st);
    }
  }
  const uint32_t rest = anyMask & ~fullMask;
  if (rest) ws_partial<C>(P, g, sg, rest, lane, st);
}
#else
// CPU emulation (tests/emu): the same fragments, the MMA spelled out over the 32 lanes of the warp
template <class C>
inline void ws_main(const Params& P, const Geom& g, const WsStage<C>& sg, ThreadState<C>* st) {
  constexpr int NT = C::NT;
  const uint32_t fullMask = sg.fullMask, anyMask = sg.anyMask;
  for (int j = 0; j < 8 && fullMask; j++) {
    if (!((fullMask >> (4 * j)) & 0xfu)) continue;
    double bq[NT][32], xr[4][32], xi[4][32];
    for (int lane = 0; lane < 32; lane++) {
      const int ks = lane & 3, b = lane >> 2, s = 4 * j + ks;
      const bool on = (fullMask >> s) & 1u;
      for (int t = 0; t < NT; t++) bq[t][lane] = on ? ((b & 1) ? sg.Q[4 * t + (b >> 1)][s].y : sg.Q[4 * t + (b >> 1)][s].x) : 0.0;
      xr[0][lane] = sg.X[b][s].x; xi[0][lane] = sg.X[b][s].y;
      const double ur = sg.W[0][s].x, ui = sg.W[0][s].y, wr = sg.W[1][s].x, wi = sg.W[1][s].y;
      xr[1][lane] = fma(xr[0][lane], ur, -(xi[0][lane] * ui)); xi[1][lane] = fma(xr[0][lane], ui, xi[0][lane] * ur);
      for (int a = 2; a < 4; a++) {
        xr[a][lane] = fma(xr[a - 2][lane], wr, -(xi[a - 2][lane] * wi));
        xi[a][lane] = fma(xr[a - 2][lane], wi, xi[a - 2][lane] * wr);
      }
    }
    for (int a = 0; a < 4; a++)
      for (int k1 = 0; k1 < 2; k1++)
        for (int t = 0; t < NT; t++)
          for (int lane = 0; lane < 32; lane++) {        // D[row][col] += sum_k A[row][k] B[k][col]
            const int row = lane >> 2;
            for (int e = 0; e < 2; e++) {
              const int col = 2 * (lane & 3) + e;
              double d = st[lane].acc[((a * 2 + k1) * NT + t) * 2 + e];
              for (int k = 0; k < 4; k++) d = fma((k1 ? xi : xr)[a][4 * row + k], bq[t][4 * col + k], d);
              st[lane].acc[((a * 2 + k1) * NT + t) * 2 + e] = d;
            }
          }
  }
  const uint32_t rest = anyMask & ~fullMask;
  if (rest) for (int lane = 0; lane < 32; lane++) ws_partial<C>(P, g, sg, rest, lane, st[lane]);
}
#endif

// fragment layout <-> tile layout (lane = tile m, acc[(p*NC + c)*4 + u]) through a scratch area of 32*NACC doubles
// (the stage handed over with the flush item)
template <class C>
SRB_HD void ws_store_frag(double* buf, int lane, const ThreadState<C>& st) {
  const int ks = lane & 3, b = lane >> 2;
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int t = 0; t < C::NT; t++)
#pragma unroll
      for (int u = 0; u < 4; u++) buf[(8 * a + b) * C::NACC + (4 * t + ks) * 4 + u] = st.acc[mma_acc_index<C>(a, t, u)];
}
template <class C>
SRB_HD void ws_load_frag(const double* buf, int lane, ThreadState<C>& st) {
  const int ks = lane & 3, b = lane >> 2;
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int t = 0; t < C::NT; t++)
#pragma unroll
      for (int u = 0; u < 4; u++) st.acc[mma_acc_index<C>(a, t, u)] = buf[(8 * a + b) * C::NACC + (4 * t + ks) * 4 + u];
}
template <class C>
SRB_HD void ws_load_tile(const double* buf, int lane, ThreadState<C>& st) {
#pragma unroll
  for (int k = 0; k < C::NACC; k++) st.acc[k] = buf[lane * C::NACC + k];
}
template <class C>
SRB_HD void ws_store_tile(double* buf, int lane, const ThreadState<C>& st) {
#pragma unroll
  for (int k = 0; k < C::NACC; k++) buf[lane * C::NACC + k] = st.acc[k];
}

#if !defined(__CUDA_ARCH__)
// ------------------------------------------------------------------------------------------------ CPU emulation
// One unit (virtual direction vd, particle chunk pc) with producer and consumer run in sequence per item.
template <class C>
inline void ws_emulate_task(const Params& P, uint32_t vd, uint32_t pc) {
  using TI = typename C::TI;
  static_assert(sizeof(WsStage<C>) >= 32 * C::NACC * sizeof(double), "stage too small for the flush transposes");
  Geom g;
  make_geom<C>(P, vd, g);
  uint32_t t0, t1;
  chunk_tracks(P, pc, t0, t1);
  WsStage<C>* sg = new WsStage<C>();
  std::memset(sg, 0, sizeof *sg);
  ThreadState<C> st[32];
  unsigned long long nPass = 0, nAll = 0;
  WsConst kc;
  ws_const<C>(P, g, kc);
  Walk<C> w;
  w.init(t0, t1, pc);
  WalkItem it;
  while (w.next(P, it)) {
    const TrackView& tv = w.tv;
    if (it.newTrack)
      for (int lane = 0; lane < 32; lane++)
        for (int k = 0; k < C::NACC; k++) st[lane].acc[k] = 0.0;
    if (it.kind == 0) {
      uint32_t fullMask = 0, anyMask = 0;
      for (int lane = 0; lane < 32; lane++) {
        const bool active = lane < it.cnt;
        double tau = 0, tauPrev = 0, a[3] = {0, 0, 0}, b[3] = {0, 0, 0};
        if (active) {
          const uint32_t i = it.base + (uint32_t)lane;
          tau = far_tau<C>(P, g, tv, i);
          tauPrev = i == 0 ? 0.0 : far_tau<C>(P, g, tv, i - 1);
          for (int c = 0; c < 3; c++) { a[c] = tv.pre[c * P.preStride + i]; b[c] = tv.pre[(3 + c) * P.preStride + i]; }
        }
        WsStep ws;
        ws_prep_guard<C>(P, g, kc, active, tau, tauPrev, a, b, ws, nPass, nAll);
        const uint32_t fl = ws_prep_store<C>(P, kc, ws, *sg, lane);
        if (fl == 1u) fullMask |= 1u << lane;
        if (fl != 0u) anyMask |= 1u << lane;
      }
      sg->cnt = (uint32_t)it.cnt; sg->fullMask = fullMask; sg->anyMask = anyMask;
      ws_main<C>(P, g, *sg, st);
    } else {
      double* buf = reinterpret_cast<double*>(sg);
      for (int lane = 0; lane < 32; lane++) ws_store_frag<C>(buf, lane, st[lane]);
      for (int lane = 0; lane < 32; lane++) ws_load_tile<C>(buf, lane, st[lane]);
      for (int lane = 0; lane < 32; lane++) flush_lane<C>(P, g, tv, pc, it.iSnap, lane, st);
      for (int lane = 0; lane < 32; lane++) ws_store_tile<C>(buf, lane, st[lane]);
      for (int lane = 0; lane < 32; lane++) ws_load_frag<C>(buf, lane, st[lane]);
      std::memset(sg, 0, sizeof *sg);      // (the GPU stage holds finite numbers after a flush; keep the emulation clean)
    }
  }
  if (P.counters) { P.counters[0] += nPass; P.counters[1] += nAll; }
  delete sg;
}
#endif

}  // namespace srb
