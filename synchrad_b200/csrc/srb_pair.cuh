// synchrad_b200 — "symmetric pair" main phase (KIND_PAIR): the fastest far-field path on uniform omega grids.
//
// A lane owns the interleaved tile {m, m+32, ..., m+32(TW-1)} of its chunk.  The tile is symmetric about
// its centre phase phi_c = phi0 + (m + 16(TW-1)) d  (d = domega*tau), and node pairs sit at
// phi_c -/+ (16+32p) d, p = 0..TW/2-1.  With X = exp(i phi_c) and Q_p = exp(i (16+32p) d):
//       exp(i phi_+-) = X * (cos_p +- i sin_p)
// The pair offsets are the SAME for every lane and tile, so (cos_p, sin_p) are computed once per
// (direction, step) in the prep phase and broadcast; the lane only forms B = A*X (per step) and
// accumulates four real sums per pair and component
//       U1 += Br cos_p   U2 += Bi sin_p   U3 += Br sin_p   U4 += Bi cos_p
//       F(+) = (U1 - U2) + i (U3 + U4)      F(-) = (U1 + U2) + i (U4 - U3)
// => 4 FMAs per (node, step) for both transverse components together with NO per-lane recurrence: the
// three-operand recurrence ops that cap the KIND_RECUR loop at 85 % of the DFMA rate are gone, and the
// op count drops from 6.0 to 5.0 per update (incl. X = Y_a * Z_b, the two-level seed product).
// Seeds: Y_a = exp(i(phi_c0 + 8a d)), a = 0..3 and Z_b = exp(i b d), b = 0..7 (m = 8a + b): 24 doubles per
// step in shared memory instead of 64.
//
// Steps that pass the guard only partially, or whose phase is too large for the seed arithmetic to
// track the reference's rounded phase (flag 3), go through `pair_update`, which adds one node's
// contribution in the U basis with half weights; it is exact (multiplication by 1/2).
#pragma once
#include "srb_core.cuh"

namespace srb {

template <class C>
SRB_HD void make_seeds_pair(const Params& P, const Geom& g, double tau, WarpSmem<C>& sm, int s) {
  using TI = typename C::TI; using TM = typename C::TM;
  constexpr int TW = C::TW;
  const double w0 = (double)((const TI*)P.omega)[g.cLo];
  double s0, c0, sd, cd;
  sincos_big(smul(w0, tau), &s0, &c0);
  sincos_big(P.domega * tau, &sd, &cd);
  double pr[8], pi[8];                     // R^(2^i)
  pr[0] = cd; pi[0] = sd;
#pragma unroll
  for (int i = 1; i < 8; i++) { pr[i] = pr[i - 1] * pr[i - 1] - pi[i - 1] * pi[i - 1]; pi[i] = 2.0 * pr[i - 1] * pi[i - 1]; }
  // Z_b = R^b
  double zr0 = 1.0, zi0 = 0.0, zr1 = cd, zi1 = sd;
  const double cf = 2.0 * cd;
  sm.seeds[8][s] = (TM)zr0; sm.seeds[9][s] = (TM)zi0; sm.seeds[10][s] = (TM)zr1; sm.seeds[11][s] = (TM)zi1;
#pragma unroll
  for (int b = 2; b < 8; b++) {
    const double zr2 = cf * zr1 - zr0, zi2 = cf * zi1 - zi0;
    sm.seeds[8 + 2 * b][s] = (TM)zr2; sm.seeds[9 + 2 * b][s] = (TM)zi2;
    zr0 = zr1; zi0 = zi1; zr1 = zr2; zi1 = zi2;
  }
  // E_c = E0 * R^(16(TW-1))
  double er = c0, ei = s0;
#pragma unroll
  for (int i = 4; i < 8; i++) {
    if ((16 * (TW - 1)) & (1 << i)) { const double t = er * pr[i] - ei * pi[i]; ei = er * pi[i] + ei * pr[i]; er = t; }
  }
  // Y_a = E_c * R^(8a)
#pragma unroll
  for (int a = 0; a < 4; a++) {
    sm.seeds[2 * a][s] = (TM)er; sm.seeds[2 * a + 1][s] = (TM)ei;
    const double t = er * pr[3] - ei * pi[3]; ei = er * pi[3] + ei * pr[3]; er = t;
  }
  // Q_p = R^(16+32p), all TW/2 of them staged (measured: letting the lanes advance p with a three-term
  // recurrence from Q_0 trades 8 broadcast loads for 6 three-operand FP64 ops and is 6 % slower)
  double qr = pr[4], qi = pi[4];
#pragma unroll
  for (int p = 0; p < TW / 2; p++) {
    sm.rec[s][C::NV + 2 * p] = (TM)qr; sm.rec[s][C::NV + 2 * p + 1] = (TM)qi;
    const double t = qr * pr[5] - qi * pi[5]; qi = qr * pi[5] + qi * pr[5]; qr = t;
  }
}

// adds (re_p, im_p) to node (+) and (re_m, im_m) to node (-) of pair p, component c, in the U basis
template <class C>
SRB_HD void pair_update(ThreadState<C>& st, int p, int c, typename C::TM re_p, typename C::TM im_p,
                        typename C::TM re_m, typename C::TM im_m) {
  using TM = typename C::TM;
  typename C::TM* U = &st.acc[(p * C::NC + c) * 4];
  const TM h = (TM)0.5;
  U[0] += h * (re_p + re_m); U[1] += h * (re_m - re_p);
  U[2] += h * (im_p - im_m); U[3] += h * (im_p + im_m);
}

template <class C>
SRB_HD void main_pair(const Params& P, const Geom& g, const WarpSmem<C>& sm, int cnt, uint32_t fullMask,
                      uint32_t anyMask, int lane, ThreadState<C>& st) {
  using TM = typename C::TM; using TI = typename C::TI;
  constexpr int TW = C::TW, NC = C::NC, NP = TW / 2;
  const int ia = 2 * (lane >> 3), ib = 8 + 2 * (lane & 7);
  const uint32_t allMask = cnt >= 32 ? 0xffffffffu : ((1u << cnt) - 1u);
  if (fullMask == allMask) {
    TM yr = sm.seeds[ia][0], yi = sm.seeds[ia + 1][0], zr = sm.seeds[ib][0], zi = sm.seeds[ib + 1][0];
#pragma unroll 2
    for (int s = 0; s < cnt; s++) {
      const int sn = s + 1 < cnt ? s + 1 : s;
      const TM nyr = sm.seeds[ia][sn], nyi = sm.seeds[ia + 1][sn], nzr = sm.seeds[ib][sn], nzi = sm.seeds[ib + 1][sn];
      const TM xr = fma(yr, zr, -(yi * zi)), xi = fma(yr, zi, yi * zr);
      TM br[NC], bi[NC];
#pragma unroll
      for (int c = 0; c < NC; c++) { const TM a = sm.rec[s][c]; br[c] = a * xr; bi[c] = a * xi; }
#pragma unroll
      for (int p = 0; p < NP; p++) {
        const TM qc = sm.rec[s][NC + 2 * p], qs = sm.rec[s][NC + 2 * p + 1];
#pragma unroll
        for (int c = 0; c < NC; c++) {
          TM* U = &st.acc[(p * NC + c) * 4];
          U[0] = fma(br[c], qc, U[0]); U[1] = fma(bi[c], qs, U[1]);
          U[2] = fma(br[c], qs, U[2]); U[3] = fma(bi[c], qc, U[3]);
        }
      }
      yr = nyr; yi = nyi; zr = nzr; zi = nzi;
    }
    return;
  }
  for (int s = 0; s < cnt; s++) {
    if (!((anyMask >> s) & 1u)) continue;
    const uint32_t r = sm.rng[s];
    const uint32_t flag = r >> 30;
    const int hiN = (int)((r >> 10) & 0x3ffu);          // passing chunk-relative nodes: [0, hiN)
    TM A[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) A[c] = sm.rec[s][c];
    TM xr = 0, xi = 0;
    if (flag != 3) {
      const TM yr = sm.seeds[ia][s], yi = sm.seeds[ia + 1][s], zr = sm.seeds[ib][s], zi = sm.seeds[ib + 1][s];
      xr = fma(yr, zr, -(yi * zi)); xi = fma(yr, zi, yi * zr);
    }
    const TM tau = sm.rec[s][NC];                       // flag 3 only
    const uint32_t jb = g.cLo + (uint32_t)lane;
#pragma unroll
    for (int p = 0; p < NP; p++) {
      const int km = NP - 1 - p, kp = NP + p;           // tile-local indices of the (-) and (+) node
      const TM qcp = sm.rec[s][NC + 2 * p], qsp = sm.rec[s][NC + 2 * p + 1];
      if (lane + 32 * km >= hiN) continue;              // (-) fails, so does (+); wider pairs have a LOWER (-) node
      const bool pp = lane + 32 * kp < hiN;
      TM cp, sp, cm, sm_;
      if (flag == 3) {
        sincos_t(tmul((TM)((const TI*)P.omega)[jb + 32 * km], tau), &sm_, &cm);
        cp = 0; sp = 0;
        if (pp) sincos_t(tmul((TM)((const TI*)P.omega)[jb + 32 * kp], tau), &sp, &cp);
      } else {
        cp = fma(xr, qcp, -(xi * qsp)); sp = fma(xr, qsp, xi * qcp);       // X * Q
        cm = fma(xr, qcp, xi * qsp);    sm_ = fma(xi, qcp, -(xr * qsp));   // X * conj(Q)
      }
#pragma unroll
      for (int c = 0; c < NC; c++)
        pair_update<C>(st, p, c, pp ? A[c] * cp : (TM)0, pp ? A[c] * sp : (TM)0, A[c] * cm, A[c] * sm_);
    }
  }
}

// complex amplitude of tile-local node k from the U sums
template <class C>
SRB_HD void pair_node_amp(const ThreadState<C>& me, int k, double re[3], double im[3]) {
  constexpr int NP = C::TW / 2;
  const bool plus = k >= NP;
  const int p = plus ? k - NP : NP - 1 - k;
#pragma unroll
  for (int c = 0; c < C::NC; c++) {
    const typename C::TM* U = &me.acc[(p * C::NC + c) * 4];
    re[c] = plus ? (double)U[0] - (double)U[1] : (double)U[0] + (double)U[1];
    im[c] = plus ? (double)U[2] + (double)U[3] : (double)U[3] - (double)U[2];
  }
}

}  // namespace srb
