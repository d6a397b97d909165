// synchrad_b200 — "gridding" main phase (KIND_SPREAD, phasor = SRB_PHASOR_SPREAD): the sum over time steps as a
// type-1 non-uniform FFT.  EXPERIMENTAL (explicit opt-in; far field, fp64, total / cartesian(_complex), uniform
// omega grids with <= 256 nodes).
//
// For a step that passes the Nyquist guard at every node, the contribution to node j = jc + k is
//       A_s exp(i w_j tau_s) = [A_s exp(i w_jc tau_s)] * exp(i k x_s),      x_s = domega * tau_s  (mod 2 pi),
// i.e. F_k = sum_s c_s exp(i k x_s) with complex strengths c_s at the non-uniform points x_s: instead of
// 4 FMAs per (node, step) — the floor of any direct summation — each step is SPREAD onto SP_W = 13 cells of a
// 2x oversampled periodic grid (SP_N = 512 cells) with the "exponential of semicircle" kernel
// psi(u) = exp(beta (sqrt(1 - (2u/W)^2) - 1)), beta = 2.30 W (Barnett, Magland, af Klinteberg 2019), one 512-point
// FFT per (track, snapshot interval, component) turns the grid into the sums, and dividing by the kernel's
// Fourier transform removes the smoothing.  Error ~1e-12 * sum|c_s| (measured against the direct sum:
// <= 4e-12 of max|F| on rough amplitudes), far inside the 1e-9 parity budget.
//
// Mapping: tau grows monotonically with the step (d tau/dt = 1 - n.beta > 0), so the 13 cells a step touches
// drift slowly upwards: lane = cell of a 32-cell WINDOW whose 4 sums (Re, Im of the two transverse components)
// live in registers.  The kernel value of lane's cell is the polynomial piece p = lane - m of degree SP_DEG in the
// warp-uniform fractional offset, with the piece's coefficients held in registers and reloaded from shared
// memory only when the integer shift m changes; the window is written to the shared-memory grid only when the
// points leave it.  Per step: SP_DEG + 4 DFMAs per lane, against 36 + prep in the pair kernel.
// Steps that pass the guard partially (or with |phase| > 2^18) are evaluated node by node (main_direct) into the
// same per-node accumulators, so every quirk of the reference's guard is kept.
#pragma once
#include "srb_core.cuh"

namespace srb {

constexpr int SP_TAB_COEF = 0;                              // [SP_DEG+1][16]
constexpr int SP_TAB_DECONV = (SP_DEG + 1) * 16;            // [256]: 1 / Psi(2 pi k / SP_N), k = -128..127
constexpr int SP_TAB_TWID = SP_TAB_DECONV + 256;            // [256][2]: exp(+2 pi i q / SP_N)
constexpr int SP_TAB_SIZE = SP_TAB_TWID + 512;
constexpr double SP_BETA = 2.30 * SP_W;
#if defined(SRB_SPREAD_V2) && defined(__CUDACC__)
// SRB_SPREAD_V2 (NOT the shipped configuration; see the end of this file): the polynomial pieces as constant-bank
// operands of the prep phase's DFMAs (uploaded next to g_spread_tab by srb_integrate)
__constant__ double c_spread_coef[(SP_DEG + 1) * 16];
#endif
constexpr int SP_Y = 4, SP_C = 8;   // rec row of an all-pass step: [V0, V1, tau, -, y, y^2, y^4, y^8, c0re, c0im, c1re, c1im]

// Host: the three tables (kernel pieces as centred monomials in y = 2t - 1, deconvolution factors, twiddles).
inline void spread_build_tables(double* tab) {
  typedef long double L;
  const L PI = 3.141592653589793238462643383279502884L;
  auto psi = [](L u) -> L {
    const L z = 2 * u / SP_W;
    return (z * z < 1) ? std::exp((L)SP_BETA * (std::sqrt(1 - z * z) - 1)) : (L)0;
  };
  constexpr int D = SP_DEG;
  for (int k = 0; k <= D; k++) for (int p = 0; p < 16; p++) tab[SP_TAB_COEF + k * 16 + p] = 0.0;
  for (int p = 0; p < SP_W; p++) {
    // Chebyshev interpolation of piece p on t in [0,1] (u = p + t - W/2), then Chebyshev -> monomial in y
    L a[D + 1];
    for (int k = 0; k <= D; k++) {
      L acc = 0;
      for (int i = 0; i <= D; i++) {
        const L th = PI * (i + 0.5L) / (D + 1);
        acc += psi(p + (std::cos(th) + 1) / 2 - (L)SP_W / 2) * std::cos(k * th);
      }
      a[k] = acc * 2 / (D + 1);
    }
    a[0] /= 2;
    L mono[D + 1] = {0}, Tkm1[D + 1] = {0}, Tk[D + 1] = {0};
    Tkm1[0] = 1;                    // T_0
    if (D >= 1) Tk[1] = 1;          // T_1
    mono[0] += a[0];
    for (int i = 0; i <= D && D >= 1; i++) mono[i] += a[1] * Tk[i];
    for (int k = 2; k <= D; k++) {
      L Tn[D + 1] = {0};
      for (int i = 0; i <= D; i++) { Tn[i] -= Tkm1[i]; if (i + 1 <= D) Tn[i + 1] += 2 * Tk[i]; }
      for (int i = 0; i <= D; i++) { mono[i] += a[k] * Tn[i]; Tkm1[i] = Tk[i]; Tk[i] = Tn[i]; }
    }
    for (int k = 0; k <= D; k++) tab[SP_TAB_COEF + k * 16 + p] = (double)mono[k];
  }
  // Psi(xi) = int psi(u) cos(xi u) du over [-W/2, W/2]: psi vanishes to all orders that matter at the ends, so the
  // trapezoid rule converges spectrally
  const int M = 4096;
  for (int i = 0; i < 256; i++) {
    const L xi = 2 * PI * (i - 128) / SP_N;
    L acc = 0;
    for (int q = 1; q < M; q++) { const L u = -(L)SP_W / 2 + (L)SP_W * q / M; acc += psi(u) * std::cos(xi * u); }
    tab[SP_TAB_DECONV + i] = (double)(1 / (acc * SP_W / M));
  }
  for (int q = 0; q < 256; q++) {
    tab[SP_TAB_TWID + 2 * q] = (double)std::cos(2 * PI * q / SP_N);
    tab[SP_TAB_TWID + 2 * q + 1] = (double)std::sin(2 * PI * q / SP_N);
  }
}

// prep phase (lane = step): strengths c = A exp(i w_jc tau) and grid position of the step
template <class C>
SRB_HD int spread_stage(const Params& P, const Geom& g, double tau, const double* V, WarpSmem<C>& sm, int s) {
  using TI = typename C::TI;
  static_assert(C::NC == 2 && C::MODE == MODE_FAR && sizeof(typename C::TM) == 8, "gridding kind: far field, transverse basis, fp64");
  const uint32_t jc = (g.cHi - g.cLo) / 2u;
  const double wc = (double)((const TI*)P.omega)[g.cLo + jc];
  double sn, cs;
  sincos_big(smul(wc, tau), &sn, &cs);                 // the reference's own rounded phase at the centre node
  double q = (P.domega * tau) * 0.15915494309189533577;   // cycles
  q -= floor(q);
  const double a = q * (double)SP_N - 0.5 * (double)SP_W;
  const double L0 = ceil(a);
  const double t = L0 - a;                             // in [0, 1)
  const double y = 2.0 * t - 1.0, y2 = y * y, y4 = y2 * y2;
  // 16-byte aligned groups so that the main phase reads a step with four 128-bit broadcast loads
  sm.rec[s][SP_Y] = y; sm.rec[s][SP_Y + 1] = y2; sm.rec[s][SP_Y + 2] = y4; sm.rec[s][SP_Y + 3] = y4 * y4;
#if defined(SRB_SPREAD_V2)
  {
    // lane = step: all 13 kernel values of this step, 13 independent Estrin chains with warp-uniform coefficients
    const double y8 = y4 * y4;
#if defined(__CUDA_ARCH__)
    const double* cf = c_spread_coef;
#else
    const double* cf = P.spreadTab + SP_TAB_COEF;
#endif
#pragma unroll
    for (int p = 0; p < SP_W; p++) {
      const double a0 = fma(cf[1 * 16 + p], y, cf[0 * 16 + p]), a1 = fma(cf[3 * 16 + p], y, cf[2 * 16 + p]);
      const double a2 = fma(cf[5 * 16 + p], y, cf[4 * 16 + p]), a3 = fma(cf[7 * 16 + p], y, cf[6 * 16 + p]);
      const double a4 = fma(cf[9 * 16 + p], y, cf[8 * 16 + p]), a5 = fma(cf[11 * 16 + p], y, cf[10 * 16 + p]);
      const double b0 = fma(a1, y2, a0), b1 = fma(a3, y2, a2), b2 = fma(a5, y2, a4);
      sm.kv[s][p] = fma(b2, y8, fma(b1, y4, b0));
    }
    sm.kv[s][SP_W] = 0.0;
  }
#endif
#pragma unroll
  for (int c = 0; c < 2; c++) { sm.rec[s][SP_C + 2 * c] = V[c] * cs; sm.rec[s][SP_C + 2 * c + 1] = V[c] * sn; }   // (Re, Im) per component
  return (int)L0;
}

template <class C>
SRB_HD void spread_load_piece(const WarpSmem<C>& sm, int lane, int m, ThreadState<C>& st) {
  const int p = lane - m;
  const bool on = p >= 0 && p < SP_W;
#pragma unroll
  for (int k = 0; k <= SP_DEG; k++) st.cf[k] = on ? sm.coef[k][p] : 0.0;
  st.curM = m;
}

// once per warp task: tables into shared memory, empty grid, closed window
template <class C>
SRB_HD void spread_init(const Params& P, WarpSmem<C>& sm, int lane, ThreadState<C>& st) {
  for (int i = lane; i < (SP_DEG + 1) * 16; i += 32) (&sm.coef[0][0])[i] = P.spreadTab[SP_TAB_COEF + i];
  for (int i = lane; i < SP_N * 4; i += 32) (&sm.grid[0][0])[i] = 0.0;
  st.W0 = 0; st.curM = -1; st.have = 0; st.dirty = 0;
#pragma unroll
  for (int c = 0; c < 4; c++) st.sacc[c] = 0.0;
}

// the window's sums go to their grid cells (one distinct cell per lane)
template <class C>
SRB_HD void spread_close_window(WarpSmem<C>& sm, int lane, ThreadState<C>& st) {
  if (!st.have) return;
  double* cell = sm.grid[(st.W0 + lane) & (SP_N - 1)];
#pragma unroll
  for (int c = 0; c < 4; c++) { cell[c] += st.sacc[c]; st.sacc[c] = 0.0; }
  st.have = 0;
  st.dirty = 1;
#if defined(__CUDA_ARCH__)
  __syncwarp();
#endif
}

struct SpreadStep { double y, y2, y4, y8, c[4]; };
#if defined(SRB_SPREAD_V2)
template <class C> SRB_HD void main_spread_v2(const Params&, WarpSmem<C>&, int, uint32_t, int, ThreadState<C>&);
template <class C> SRB_HD void spread_close_window(WarpSmem<C>&, int, ThreadState<C>&);
#endif

template <class C>
SRB_HD SpreadStep spread_read_step(const WarpSmem<C>& sm, int s) {
  SpreadStep r;
#if defined(__CUDA_ARCH__)
  const double2* q = reinterpret_cast<const double2*>(&sm.rec[s][SP_Y]);
  const double2 a = q[0], b = q[1], c0 = q[2], c1 = q[3];
  r.y = a.x; r.y2 = a.y; r.y4 = b.x; r.y8 = b.y; r.c[0] = c0.x; r.c[1] = c0.y; r.c[2] = c1.x; r.c[3] = c1.y;
#else
  r.y = sm.rec[s][SP_Y]; r.y2 = sm.rec[s][SP_Y + 1]; r.y4 = sm.rec[s][SP_Y + 2]; r.y8 = sm.rec[s][SP_Y + 3];
  for (int c = 0; c < 4; c++) r.c[c] = sm.rec[s][SP_C + c];
#endif
  return r;
}

// kernel value of this lane's cell: Estrin evaluation of the degree-11 piece (11 FMAs, depth 4; the powers of the
// warp-uniform offset come from the prep phase)
template <class C>
SRB_HD double spread_kernel_value(const ThreadState<C>& st, const SpreadStep& r) {
  static_assert(SP_DEG == 11, "Estrin scheme below is written for degree 11");
  const double a0 = fma(st.cf[1], r.y, st.cf[0]), a1 = fma(st.cf[3], r.y, st.cf[2]), a2 = fma(st.cf[5], r.y, st.cf[4]);
  const double a3 = fma(st.cf[7], r.y, st.cf[6]), a4 = fma(st.cf[9], r.y, st.cf[8]), a5 = fma(st.cf[11], r.y, st.cf[10]);
  const double b0 = fma(a1, r.y2, a0), b1 = fma(a3, r.y2, a2), b2 = fma(a5, r.y2, a4);
  return fma(b2, r.y8, fma(b1, r.y4, b0));
}

// main phase over the all-pass steps of a sub-batch (lane = window cell; control flow is warp-uniform).
// sm.rng[s] holds the first cell L0 of an all-pass step.  `same` marks steps whose predecessor is an all-pass
// step with the same L0: they need no window / piece bookkeeping, and four of them in a row are evaluated
// together (independent dependency chains).
template <class C>
SRB_HD void main_spread(const Params& P, WarpSmem<C>& sm, int cnt, uint32_t fullMask, int lane, ThreadState<C>& st) {
#if defined(SRB_SPREAD_V2)
  main_spread_v2<C>(P, sm, cnt, fullMask, lane, st);
  return;
#endif
  uint32_t same = 0u;
#if defined(__CUDA_ARCH__)
  {
    const uint32_t mine = sm.rng[lane];
    const uint32_t prev = __shfl_up_sync(0xffffffffu, mine, 1);
    same = __ballot_sync(0xffffffffu, lane > 0 && lane < cnt && (mine >> 30) == 1u && mine == prev);
  }
#else
  for (int i = 1; i < cnt; i++) if ((sm.rng[i] >> 30) == 1u && sm.rng[i] == sm.rng[i - 1]) same |= 1u << i;
#endif
  uint32_t todo = fullMask;
  while (todo) {
#if defined(__CUDA_ARCH__)
    const int s = __ffs((int)todo) - 1;
#else
    const int s = __builtin_ctz(todo);
#endif
    if (!((same >> s) & 1u)) {
      const int L0 = (int)(sm.rng[s] & 0x3ffu) - 16;
      int m = st.have ? ((L0 - st.W0) & (SP_N - 1)) : 0;
      if (st.have && m > 32 - SP_W) { spread_close_window<C>(sm, lane, st); m = 0; }
      if (!st.have) { st.W0 = L0; st.have = 1; st.curM = -1; }
      if (m != st.curM) spread_load_piece<C>(sm, lane, m, st);
    }
    if (s + 3 < 32 && ((same >> (s + 1)) & 7u) == 7u) {       // s+1..s+3 continue the run
      SpreadStep r[4];
      double k[4];
#pragma unroll
      for (int i = 0; i < 4; i++) r[i] = spread_read_step<C>(sm, s + i);
#pragma unroll
      for (int i = 0; i < 4; i++) k[i] = spread_kernel_value<C>(st, r[i]);
#pragma unroll
      for (int i = 0; i < 4; i++) {
#pragma unroll
        for (int c = 0; c < 4; c++) st.sacc[c] = fma(k[i], r[i].c[c], st.sacc[c]);
      }
      todo &= ~(0xfu << s);
    } else {
      const SpreadStep r = spread_read_step<C>(sm, s);
      const double k = spread_kernel_value<C>(st, r);
#pragma unroll
      for (int c = 0; c < 4; c++) st.sacc[c] = fma(k, r.c[c], st.sacc[c]);
      todo &= todo - 1u;
    }
  }
}

#if defined(SRB_SPREAD_V2)
// SRB_SPREAD_V2 main phase (lane = window cell): the kernel values were computed by the prep phase with lane = step
// (143 DFMAs per lane and sub-batch instead of 11 per lane and STEP), so what is left per step is one predicated
// load of this cell's value, the broadcast strengths and 4 FMAs -- a dense, branch-free loop over the 16 steps of
// a half sub-batch once a pre-check has established that they all fit the 32-cell window.
// Validated on the CPU emulation only (tests/test_emulated_kernels.py::test_gridding_kernel_v2_logic); to be
// measured and GPU-validated before it replaces the loop above.
template <class C>
SRB_HD void main_spread_v2(const Params& P, WarpSmem<C>& sm, int cnt, uint32_t fullMask, int lane, ThreadState<C>& st) {
  for (int half = 0; half < 2; half++) {
    const uint32_t mask = fullMask & (half ? 0xffff0000u : 0x0000ffffu);
    if (!mask) continue;
#if defined(__CUDA_ARCH__)
    const int first = __ffs((int)mask) - 1, last = 31 - __clz((int)mask);
#else
    const int first = __builtin_ctz(mask), last = 31 - __builtin_clz(mask);
#endif
    const int Lf = (int)(sm.rng[first] & 0x3ffu) - 16, Ll = (int)(sm.rng[last] & 0x3ffu) - 16;
    if (!st.have) { st.W0 = Lf; st.have = 1; }
    if (((Ll - st.W0) & (SP_N - 1)) > 32 - SP_W) {          // the half would leave the window: write it out, re-anchor
      spread_close_window<C>(sm, lane, st);
      st.W0 = Lf; st.have = 1;
    }
    const bool fits = ((Ll - st.W0) & (SP_N - 1)) <= 32 - SP_W;   // tau is monotone, so every step of the half fits
    const int base = 16 * half;
    if (fits) {
#pragma unroll
      for (int i = 0; i < 16; i++) {
        const int s = base + i;
        if (!((mask >> s) & 1u)) continue;
        const int d = ((int)(sm.rng[s] & 0x3ffu) - 16 - st.W0) & (SP_N - 1);
        const int p = lane - d;
        const double k = (unsigned)p < (unsigned)SP_W ? sm.kv[s][p] : 0.0;
#pragma unroll
        for (int c = 0; c < 4; c++) st.sacc[c] = fma(k, sm.rec[s][SP_C + c], st.sacc[c]);
      }
    } else {                                                  // > 19 cells of drift within 16 steps: step by step
      for (int i = 0; i < 16; i++) {
        const int s = base + i;
        if (!((mask >> s) & 1u)) continue;
        const int L0 = (int)(sm.rng[s] & 0x3ffu) - 16;
        if (((L0 - st.W0) & (SP_N - 1)) > 32 - SP_W) { spread_close_window<C>(sm, lane, st); st.W0 = L0; st.have = 1; }
        const int p = lane - ((L0 - st.W0) & (SP_N - 1));
        const double k = (unsigned)p < (unsigned)SP_W ? sm.kv[s][p] : 0.0;
#pragma unroll
        for (int c = 0; c < 4; c++) st.sacc[c] = fma(k, sm.rec[s][SP_C + c], st.sacc[c]);
      }
    }
  }
}
#endif

SRB_HD int sp_bitrev9(int v) {
  int r = 0;
#pragma unroll
  for (int b = 0; b < 9; b++) r |= ((v >> b) & 1) << (8 - b);
  return r;
}

// one radix-2 decimation-in-frequency stage of X_k = sum_l x_l exp(+2 pi i k l / 512), both components at once;
// after the 9 stages X_k sits at cell bitrev9(k)
template <class C>
SRB_HD void spread_fft_stage(WarpSmem<C>& sm, const Params& P, int stage, int lane) {
  const int h = (SP_N / 2) >> stage;               // 256, 128, ..., 1
  const double* tw = P.spreadTab + SP_TAB_TWID;
  for (int r = 0; r < SP_N / 64; r++) {
    const int q = lane + 32 * r;
    const int j = q & (h - 1);
    const int i0 = ((q - j) << 1) + j, i1 = i0 + h;
    const int ti = j << stage;                       // j * (256 / h)
    const double wr = tw[2 * ti], wi = tw[2 * ti + 1];
    double* a = sm.grid[i0];
    double* b = sm.grid[i1];
#pragma unroll
    for (int c = 0; c < 2; c++) {
      const double ar = a[2 * c], ai = a[2 * c + 1], br = b[2 * c], bi = b[2 * c + 1];
      a[2 * c] = ar + br; a[2 * c + 1] = ai + bi;
      const double dr = ar - br, di = ai - bi;
      b[2 * c] = dr * wr - di * wi; b[2 * c + 1] = dr * wi + di * wr;
    }
  }
}

// deconvolved sums of this lane's nodes are added to the per-node accumulators (direct layout)
template <class C>
SRB_HD void spread_extract(const Params& P, const Geom& g, WarpSmem<C>& sm, int lane, ThreadState<C>& st) {
  const int n = (int)(g.cHi - g.cLo), jc = n / 2;
  const double* dec = P.spreadTab + SP_TAB_DECONV;
#pragma unroll
  for (int k = 0; k < C::TW; k++) {
    const int jj = lane + 32 * k;
    if (jj >= n) continue;
    const int kk = jj - jc;                          // in [-128, 128)
    const double* cell = sm.grid[sp_bitrev9(kk & (SP_N - 1))];
    const double d = dec[kk + 128];
    // grid cell layout: (Re c0, Im c0, Re c1, Im c1); accumulators: Re c0, Re c1, Im c0, Im c1
    st.acc[k * C::NPN + 0] += cell[0] * d; st.acc[k * C::NPN + 1] += cell[2] * d;
    st.acc[k * C::NPN + 2] += cell[1] * d; st.acc[k * C::NPN + 3] += cell[3] * d;
  }
}

template <class C>
SRB_HD void spread_clear(WarpSmem<C>& sm, int lane, ThreadState<C>& st) {
  for (int i = lane; i < SP_N * 4; i += 32) (&sm.grid[0][0])[i] = 0.0;
  st.dirty = 0;
}

}  // namespace srb
