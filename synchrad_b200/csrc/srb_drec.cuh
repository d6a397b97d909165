// synchrad_b200 — KIND_DREC: "direct layout, corrected recurrence".  Uniform ascending omega grids, fp64, phases of ANY
// magnitude: the near field at large screen distances (omega*(t + R) ~ 1e10 rad for the reference's own near-field
// test, kernel_nearfield.cl:64-92) and SI-unit far fields, where the phase-tracking kernels (recurrence, pair) stop
// at |phase| = 2^18 and the direct kernel pays a 25-op sincos per update.
//
// The reference evaluates sin/cos of the ROUNDED product p_j = fl(w_j * tau).  At 1e10 rad one ulp of the phase is
// 2e-6 rad, so that rounding is visible at the 1e-6 level and has to be reproduced to match the reference to 1e-9.
// Exactly:   w_j * tau = p_j + e_j,   e_j = fma(w_j, tau, -p_j)                    (the product's rounding error)
//            w_j       = w_c + (j - c) * dw + eps_j                                (table node vs ideal grid node;
//                                                                                   eps_j in double-double, once per lane)
//   =>  p_j = theta_j + c_j,   theta_j = (w_c + (j - c) dw) * tau  exactly,   c_j = eps_j * tau - e_j   (|c_j| ~ 1 ulp(p))
//   =>  exp(i p_j) = exp(i theta_j) * (1 + i c_j)                                  (second order: c^2/2 < 1e-11)
// exp(i theta_j) comes from unit-modulus phasor arithmetic whose error does not grow with |phase|:
//   prep phase (lane = step): E = exp(i w_c tau) and R = exp(i dw tau), each as sincos of the rounded product times
//   (1 + i * product's rounding error); the 32 lanes' first-node phasors S_m = (E R^b) R^(8a), m = 8a + b, as two-level
//   seeds, and W = R^32;   main phase (lane = tile {lane + 32k}): v_k = 2 Re(W)... three-term recurrence with 2cos(32 d),
//   per node: 3 ops for c_j (one DMUL + two DFMA), 2 for the correction, the accumulation of the direct kernel.
// Per update: far 2 + 5 + 6 = 13 FP64 slots (direct kernel: 32), near 2 + 5 + 14 = 21 (direct: 40).
#pragma once
#include "srb_core.cuh"

namespace srb {

// eps_k of the lane's nodes j = cLo + lane + 32k relative to the ideal grid anchored at the chunk's first node
// (w_c = omega[cLo], c = 0): double-double evaluation of omega[j] - (omega[cLo] + (j - cLo) * domega)
template <class C>
SRB_HD void drec_init_lane(const Params& P, const Geom& g, int lane, ThreadState<C>& st) {
  using TI = typename C::TI;
  const double w0 = (double)((const TI*)P.omega)[g.cLo];
#pragma unroll
  for (int k = 0; k < C::TW; k++) {
    const uint32_t j = g.cLo + (uint32_t)(lane + 32 * k);
    double e = 0.0;
    if (j < g.cHi) {
      const double m = (double)(lane + 32 * k);
      const double hi = smul(m, P.domega), lo = fma(m, P.domega, -hi);       // m * dw = hi + lo exactly
      const double s = sadd(w0, hi);                                          // two-sum: w0 + hi = s + t exactly
      const double bb = ssub(s, w0);
      const double t = sadd(ssub(w0, ssub(s, bb)), ssub(hi, bb));
      const double wj = (double)((const TI*)P.omega)[j];
      e = ssub(ssub(ssub(wj, s), t), lo);                                     // wj - s is exact (neighbouring doubles)
    }
    st.eps[k] = e;
  }
}

// prep phase, lane = step s: S_m = exp(i (w0 + m dw) tau) for the 32 lanes' first nodes m = 0..31 (exact products, see
// header), W = exp(i 32 dw tau).  out: W.re, W.im, 2 W.re, -.
template <class C>
SRB_HD void make_seeds_drec(const Params& P, const Geom& g, double tau, WarpSmem<C>& sm, int s, double out[4]) {
  using TI = typename C::TI;
  const double w0 = (double)((const TI*)P.omega)[g.cLo];
  const double p0 = smul(w0, tau), e0 = fma(w0, tau, -p0);
  const double d = smul(P.domega, tau), ed = fma(P.domega, tau, -d);
  double s0, c0, sd, cd;
  sincos_big(p0, &s0, &c0);
  sincos_big(d, &sd, &cd);
  // first-order rotation by the rounding errors of the two products (|e| <= ulp/2 of a phase <= 1e13: < 1e-3; second
  // order e^2/2 kept for e0 so that the error stays below 1e-12 up to phases of ~1e11)
  const double k0 = 1.0 - 0.5 * e0 * e0;
  double er = fma(-e0, s0, c0 * k0), ei = fma(e0, c0, s0 * k0);
  const double kd = 1.0 - 0.5 * ed * ed;
  const double rr = fma(-ed, sd, cd * kd), ri = fma(ed, cd, sd * kd);
  // Z_b = E R^b, b = 0..7: three-term recurrence with 2 Re(R) (R is unit-modulus up to 1e-16)
  double x0r = er, x0i = ei;
  double x1r = er * rr - ei * ri, x1i = er * ri + ei * rr;
  const double cf = 2.0 * rr;
  sm.seed[s][0] = Cpx{x0r, x0i}; sm.seed[s][1] = Cpx{x1r, x1i};
#pragma unroll
  for (int m = 2; m < 8; m++) {
    const double x2r = fma(cf, x1r, -x0r), x2i = fma(cf, x1i, -x0i);
    sm.seed[s][m] = Cpx{x2r, x2i};
    x0r = x1r; x0i = x1i; x1r = x2r; x1i = x2i;
  }
  // Y_a = R^(8a), a = 0..3, and W = R^32
  double wr = rr, wi = ri;
#pragma unroll
  for (int i = 0; i < 3; i++) { const double t = wr * wr - wi * wi; wi = 2.0 * wr * wi; wr = t; }   // R^8
  const double r8r = wr, r8i = wi;
  { const double t = wr * wr - wi * wi; wi = 2.0 * wr * wi; wr = t; }                                // R^16
  sm.seed[s][8] = Cpx{1.0, 0.0}; sm.seed[s][9] = Cpx{r8r, r8i}; sm.seed[s][10] = Cpx{wr, wi};
  sm.seed[s][11] = Cpx{wr * r8r - wi * r8i, wr * r8i + wi * r8r};                                    // R^24
  { const double t = wr * wr - wi * wi; wi = 2.0 * wr * wi; wr = t; }                                // R^32
  out[0] = wr; out[1] = wi; out[2] = 2.0 * wr; out[3] = 0.0;
}

// one node: corrected phasor of the reference's rounded phase, then the direct kernel's accumulation
template <class C>
SRB_HD void drec_update(const double* V, double w, double eps, double tau, double vr, double vi, int k, ThreadState<C>& st) {
  const double p = smul(w, tau);
  const double c = fma(eps, tau, -fma(w, tau, -p));       // c_j = eps_j tau - e_j
  const double cs = fma(-c, vi, vr), sn = fma(c, vr, vi);
  if (C::MODE == MODE_FAR) {
#pragma unroll
    for (int q = 0; q < C::NC; q++) {
      st.acc[k * C::NPN + q] = fma(V[q], cs, st.acc[k * C::NPN + q]);
      st.acc[k * C::NPN + C::NC + q] = fma(V[q], sn, st.acc[k * C::NPN + C::NC + q]);
    }
  } else {
    const double t1 = w * sn, t2 = w * cs;
#pragma unroll
    for (int q = 0; q < 3; q++) {   // Re += -c1*sin + c2*cos ; Im += c1*cos + c2*sin
      st.acc[k * 6 + q] = fma(V[3 + q], cs, fma(-V[q], t1, st.acc[k * 6 + q]));
      st.acc[k * 6 + 3 + q] = fma(V[3 + q], sn, fma(V[q], t2, st.acc[k * 6 + 3 + q]));
    }
  }
}

template <class C>
SRB_HD void main_drec(const Params& P, const Geom& g, const WarpSmem<C>& sm, int cnt, uint32_t fullMask,
                      uint32_t anyMask, int lane, ThreadState<C>& st) {
  constexpr int TW = C::TW, NV = C::NV;
  const uint32_t allMask = cnt >= 32 ? 0xffffffffu : ((1u << cnt) - 1u);
  const bool allFull = fullMask == allMask;     // warp-uniform: every node of the chunk passes at every step
  for (int s = 0; s < cnt; s++) {
    if (!((anyMask >> s) & 1u)) continue;
    double V[NV];
#pragma unroll
    for (int q = 0; q < NV; q++) V[q] = sm.rec[s][q];
    const double tau = sm.rec[s][NV], wr = sm.rec[s][NV + 1], wi = sm.rec[s][NV + 2], cf = sm.rec[s][NV + 3];
    const Cpx zb = sm.seed[s][lane & 7], ya = sm.seed[s][8 + (lane >> 3)];
    const Cpx sd{fma(zb.re, ya.re, -(zb.im * ya.im)), fma(zb.re, ya.im, zb.im * ya.re)};      // S_lane = Z_b Y_a
    // v_0 = S_lane, v_{-1} = S conj(W); v_{k+1} = 2 Re(W) v_k - v_{k-1}: the TW phasors of the lane first, then ONE
    // straight-line block of TW independent correction + accumulation chains
    double vr[TW], vi[TW];
    vr[0] = sd.re; vi[0] = sd.im;
    double pr = fma(sd.re, wr, sd.im * wi), pi = fma(sd.im, wr, -(sd.re * wi));
#pragma unroll
    for (int k = 1; k < TW; k++) {
      vr[k] = fma(cf, vr[k - 1], -pr); vi[k] = fma(cf, vi[k - 1], -pi);
      pr = vr[k - 1]; pi = vi[k - 1];
    }
    if (allFull) {
#pragma unroll
      for (int k = 0; k < TW; k++) drec_update<C>(V, st.wl[k], st.eps[k], tau, vr[k], vi[k], k, st);
    } else {
      const uint32_t r = sm.rng[s];
      const int lo = tile_lo((int)(r & 0x3ffu), lane, 32), hi = tile_lo((int)((r >> 10) & 0x3ffu), lane, 32);
#pragma unroll
      for (int k = 0; k < TW; k++) {
        // a failing node adds exactly 0: zero phasor (the amplitudes are finite), no branch
        const bool on = k >= lo && k < hi;
        drec_update<C>(V, st.wl[k], st.eps[k], tau, on ? vr[k] : 0.0, on ? vi[k] : 0.0, k, st);
      }
    }
  }
}

}  // namespace srb
