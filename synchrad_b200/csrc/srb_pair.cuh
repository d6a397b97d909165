// synchrad_b200 — "symmetric pair" main phase (KIND_PAIR): the fastest far-field path on uniform omega grids.
//
// A lane owns the interleaved tile {m, m+32, ..., m+32(TW-1)} of its chunk.  The tile is symmetric about
// its centre phase phi_c = phi0 + (m + 16(TW-1)) d  (d = domega*tau), and node pairs sit at
// phi_c -/+ (16+32p) d, p = 0..TW/2-1.  With X = exp(i phi_c) and Q_p = exp(i (16+32p) d):
//       exp(i phi_+-) = X * (cos_p +- i sin_p)
// The pair offsets are the SAME for every lane and tile, so (cos_p, sin_p) are computed once per
// (direction, step) in the prep phase, multiplied there by the step's amplitude (Q'_pc = A_c Q_p) and
// broadcast; the lane accumulates four real sums per pair and component
//       U1 += Xr A cos_p   U2 += Xi A sin_p   U3 += Xr A sin_p   U4 += Xi A cos_p
//       F(+) = (U1 - U2) + i (U3 + U4)      F(-) = (U1 + U2) + i (U4 - U3)
// => 4 FMAs per (node, step) for both transverse components together with NO per-lane recurrence: the
// three-operand recurrence ops that cap the KIND_RECUR loop at 85 % of the DFMA rate are gone, and the
// op count drops from 6.0 to 4.5 per update (incl. X = Y_a * Z_b, the two-level seed product; the
// amplitude costs nothing in the main phase because it rides on the broadcast operand).
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
SRB_HD void make_seeds_pair(const Params& P, const Geom& g, double tau, const double* V, WarpSmem<C>& sm, int s) {
  using TI = typename C::TI; using TM = typename C::TM;
  constexpr int TW = C::TW;
  // Phase of tile 0's centre: node cLo + 16(TW-1) of the table when the chunk has that many nodes (the exact
  // rounded phase the reference forms there), else the same frequency extrapolated on the uniform grid.
  constexpr uint32_t CEN = 16u * (uint32_t)(TW - 1);
  const double wc = g.cLo + CEN < P.nOmega ? (double)((const TI*)P.omega)[g.cLo + CEN]
                                           : (double)((const TI*)P.omega)[g.cLo] + (double)CEN * P.domega;
  double er, ei, sd, cd;
  sincos_big(smul(wc, tau), &ei, &er);
  sincos_big(P.domega * tau, &sd, &cd);
  double pr[6], pi[6];                     // R^(2^i)
  pr[0] = cd; pi[0] = sd;
#pragma unroll
  for (int i = 1; i < 6; i++) { pr[i] = pr[i - 1] * pr[i - 1] - pi[i - 1] * pi[i - 1]; pi[i] = 2.0 * pr[i - 1] * pi[i - 1]; }
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
  // Y_a = E_c * R^(8a)
#pragma unroll
  for (int a = 0; a < 4; a++) {
    sm.seeds[2 * a][s] = (TM)er; sm.seeds[2 * a + 1][s] = (TM)ei;
    if (a < 3) { const double t = er * pr[3] - ei * pi[3]; ei = er * pi[3] + ei * pr[3]; er = t; }
  }
  // Q'_pc = A_c R^(16+32p), all of them staged (measured: letting the lanes advance p with a three-term
  // recurrence from Q_0 trades broadcast loads for three-operand FP64 ops and is 6 % slower)
  double qr = pr[4], qi = pi[4];
#pragma unroll
  for (int p = 0; p < TW / 2; p++) {
#pragma unroll
    for (int c = 0; c < C::NC; c++) {
      sm.rec[s][C::QOFF + 2 * (p * C::NC + c)] = (TM)(V[c] * qr);
      sm.rec[s][C::QOFF + 2 * (p * C::NC + c) + 1] = (TM)(V[c] * qi);
    }
    if (p + 1 < TW / 2) { const double t = qr * pr[5] - qi * pi[5]; qi = qr * pi[5] + qi * pr[5]; qr = t; }
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
  static_assert(!C::MMA, "the tensor-core layout has its own main phase (main_pair_mma)");
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
#pragma unroll
      for (int p = 0; p < NP; p++) {
#pragma unroll
        for (int c = 0; c < NC; c++) {
          const TM qc = sm.rec[s][C::QOFF + 2 * (p * NC + c)], qs = sm.rec[s][C::QOFF + 2 * (p * NC + c) + 1];
          TM* U = &st.acc[(p * NC + c) * 4];
          U[0] = fma(xr, qc, U[0]); U[1] = fma(xi, qs, U[1]);
          U[2] = fma(xr, qs, U[2]); U[3] = fma(xi, qc, U[3]);
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
      if (lane + 32 * km >= hiN) continue;              // (-) fails, so does (+); wider pairs have a LOWER (-) node
      const bool pp = lane + 32 * kp < hiN;
      if (flag == 3) {
        TM cp = 0, sp = 0, cm, sm_;
        sincos_t(tmul((TM)((const TI*)P.omega)[jb + 32 * km], tau), &sm_, &cm);
        if (pp) sincos_t(tmul((TM)((const TI*)P.omega)[jb + 32 * kp], tau), &sp, &cp);
#pragma unroll
        for (int c = 0; c < NC; c++) pair_update<C>(st, p, c, A[c] * cp, A[c] * sp, A[c] * cm, A[c] * sm_);
      } else {
#pragma unroll
        for (int c = 0; c < NC; c++) {
          const TM qc = sm.rec[s][C::QOFF + 2 * (p * NC + c)], qs = sm.rec[s][C::QOFF + 2 * (p * NC + c) + 1];
          const TM cp = fma(xr, qc, -(xi * qs)), sp = fma(xr, qs, xi * qc);      // X * A Q
          const TM cm = fma(xr, qc, xi * qs), sm_ = fma(xi, qc, -(xr * qs));     // X * A conj(Q)
          pair_update<C>(st, p, c, pp ? cp : (TM)0, pp ? sp : (TM)0, cm, sm_);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------- tensor-core main phase (Cfg::MMA)
// The all-pass accumulation of one 32-step sub-batch is the real GEMM
//     U[64 x 8NT] += Xm[64 x 32] * Q'[32 x 8NT]
//   rows    (a, k1, b): tile m = 8a + b of the chunk, k1 = Re|Im of X_m(step) = Y_a Z_b
//   columns n = 2q + k2 : q = p*NC + c (pair p, component c), k2 = cos|sin       (the Q' row of `rec`)
//   K       the steps of the sub-batch
// issued as DMMA.8x8x4 (mma.sync m8n8k4 f64): 8 k-groups x 4 a x 2 k1 x NT instructions.  Measured on B200:
// DMMA runs at the DFMA peak (63.8 FMA/clk/SM, tools/dmma_peak.cu) but reads 4 doubles per lane for 8 FMAs
// instead of 3 per FMA, so the register-operand ceiling that holds the DFMA loop at 70 % of the pipe
// (tools/loop_peak_pair.cu) does not apply, and the shared-memory traffic drops from 12 loads per step and
// lane to ~2.
// Fragment layout (PTX ISA, m8n8k4 .f64): A: lane l holds A[l>>2][l&3]; B: B[l&3][l>>2]; C/D: C[l>>2][2(l&3)+e].
// => lane l = 4b + ks owns, for every a and k1, the sums of tile m = 8a + b and columns q = 4t + ks:
//    acc[((a*2 + k1)*NT + t)*2 + e]  with  (k1,e): (0,0)=U1 (0,1)=U3 (1,0)=U4 (1,1)=U2  of the header comment.
// Steps that are not all-pass are masked out of the B fragments and handled lane by lane below
// (same half-weight arithmetic as pair_update).  The staging area is zeroed at kernel start, so a masked
// step's stale seeds are finite and contribute exactly 0.
#if defined(__CUDA_ARCH__)
SRB_HD void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
#endif

// index of U_u (u = 0..3 in pair_update's order: U[0]..U[3]) inside the MMA accumulator of (a, t)
template <class C>
SRB_HD constexpr int mma_acc_index(int a, int t, int u) {
  // u=0: Xr*cos (k1=0,e=0)   u=1: Xi*sin (1,1)   u=2: Xr*sin (0,1)   u=3: Xi*cos (1,0)
  return ((a * 2 + ((u == 1 || u == 3) ? 1 : 0)) * C::NT + t) * 2 + ((u == 1 || u == 2) ? 1 : 0);
}

// the steps of `rest` (partial pass or flag 3), one lane = the (tile, q) slots it owns in the MMA layout
template <class C>
SRB_HD void pair_mma_partial(const Params& P, const Geom& g, const WarpSmem<C>& sm, uint32_t rest, int lane,
                             ThreadState<C>& st) {
  using TM = typename C::TM; using TI = typename C::TI;
  constexpr int NC = C::NC, NP = C::TW / 2, NT = C::NT;
  const int ks = lane & 3, b = lane >> 2;
  for (int s = 0; s < 32; s++) {
    if (!((rest >> s) & 1u)) continue;
    const uint32_t r = sm.rng[s];
    const uint32_t flag = r >> 30;
    const int hiN = (int)((r >> 10) & 0x3ffu);          // passing chunk-relative nodes: [0, hiN)
    const TM tau = sm.rec[s][NC];                       // flag 3 only
    TM zr = 0, zi = 0;
    if (flag != 3) { zr = sm.seeds[8 + 2 * b][s]; zi = sm.seeds[9 + 2 * b][s]; }
#pragma unroll
    for (int a = 0; a < 4; a++) {
      const int m = 8 * a + b;
      TM xr = 0, xi = 0;
      if (flag != 3) {
        const TM yr = sm.seeds[2 * a][s], yi = sm.seeds[2 * a + 1][s];
        xr = fma(yr, zr, -(yi * zi)); xi = fma(yr, zi, yi * zr);
      }
#pragma unroll
      for (int t = 0; t < NT; t++) {
        const int q = 4 * t + ks, p = q / NC, c = q - p * NC;
        const int km = NP - 1 - p, kp = NP + p;         // tile-local indices of the (-) and (+) node
        if (m + 32 * km >= hiN) continue;               // (-) fails, so does (+)
        const bool pp = m + 32 * kp < hiN;
        TM cp = 0, sp = 0, cm, sm_;
        if (flag == 3) {
          const TM A = sm.rec[s][c];
          const uint32_t jb = g.cLo + (uint32_t)m;
          sincos_t(tmul((TM)((const TI*)P.omega)[jb + 32 * km], tau), &sm_, &cm);
          if (pp) sincos_t(tmul((TM)((const TI*)P.omega)[jb + 32 * kp], tau), &sp, &cp);
          cp *= A; sp *= A; cm *= A; sm_ *= A;
        } else {
          const TM qc = sm.rec[s][C::QOFF + 2 * q], qs = sm.rec[s][C::QOFF + 2 * q + 1];
          if (pp) { cp = fma(xr, qc, -(xi * qs)); sp = fma(xr, qs, xi * qc); }   // X * A Q
          cm = fma(xr, qc, xi * qs); sm_ = fma(xi, qc, -(xr * qs));              // X * A conj(Q)
        }
        const TM h = (TM)0.5;
        st.acc[mma_acc_index<C>(a, t, 0)] += h * (cp + cm);
        st.acc[mma_acc_index<C>(a, t, 1)] += h * (cm - cp);
        st.acc[mma_acc_index<C>(a, t, 2)] += h * (sp - sm_);
        st.acc[mma_acc_index<C>(a, t, 3)] += h * (sp + sm_);
      }
    }
  }
}

#if defined(__CUDA_ARCH__)
template <class C>
SRB_HD void main_pair_mma(const Params& P, const Geom& g, const WarpSmem<C>& sm, int cnt, uint32_t fullMask,
                          uint32_t anyMask, int lane, ThreadState<C>& st) {
  constexpr int NT = C::NT;
  const int ks = lane & 3, b = lane >> 2;
  if (fullMask) {
#pragma unroll 2
    for (int j = 0; j < 8; j++) {
      if (!((fullMask >> (4 * j)) & 0xfu)) continue;     // warp-uniform
      const int s = 4 * j + ks;
      const bool on = (fullMask >> s) & 1u;
      double bq[NT];
#pragma unroll
      for (int t = 0; t < NT; t++) { const double v = sm.rec[s][C::QOFF + 8 * t + b]; bq[t] = on ? v : 0.0; }
      const double zr = sm.seeds[8 + 2 * b][s], zi = sm.seeds[9 + 2 * b][s];
#pragma unroll
      for (int a = 0; a < 4; a++) {
        const double yr = sm.seeds[2 * a][s], yi = sm.seeds[2 * a + 1][s];
        const double xr = fma(yr, zr, -(yi * zi)), xi = fma(yr, zi, yi * zr);
#pragma unroll
        for (int t = 0; t < NT; t++) {
          dmma884(st.acc[((a * 2 + 0) * NT + t) * 2], st.acc[((a * 2 + 0) * NT + t) * 2 + 1], xr, bq[t]);
          dmma884(st.acc[((a * 2 + 1) * NT + t) * 2], st.acc[((a * 2 + 1) * NT + t) * 2 + 1], xi, bq[t]);
        }
      }
    }
  }
  const uint32_t rest = anyMask & ~fullMask;
  if (rest) pair_mma_partial<C>(P, g, sm, rest, lane, st);
}
#else
// CPU emulation (tests/emu): the same fragments, the MMA spelled out over the 32 lanes of the warp
template <class C>
inline void main_pair_mma(const Params& P, const Geom& g, const WarpSmem<C>& sm, int cnt, uint32_t fullMask,
                          uint32_t anyMask, ThreadState<C>* st) {
  constexpr int NT = C::NT;
  for (int j = 0; j < 8 && fullMask; j++) {
    if (!((fullMask >> (4 * j)) & 0xfu)) continue;
    double bq[NT][32], xr[4][32], xi[4][32];
    for (int lane = 0; lane < 32; lane++) {
      const int ks = lane & 3, b = lane >> 2, s = 4 * j + ks;
      const bool on = (fullMask >> s) & 1u;
      for (int t = 0; t < NT; t++) bq[t][lane] = on ? sm.rec[s][C::QOFF + 8 * t + b] : 0.0;
      const double zr = sm.seeds[8 + 2 * b][s], zi = sm.seeds[9 + 2 * b][s];
      for (int a = 0; a < 4; a++) {
        const double yr = sm.seeds[2 * a][s], yi = sm.seeds[2 * a + 1][s];
        xr[a][lane] = fma(yr, zr, -(yi * zi)); xi[a][lane] = fma(yr, zi, yi * zr);
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
  if (rest) for (int lane = 0; lane < 32; lane++) pair_mma_partial<C>(P, g, sm, rest, lane, st[lane]);
}
#endif

// MMA layout <-> tile layout (lane = tile m, acc[(p*NC + c)*4 + u]) through the warp's staging area, which is
// dead between sub-batches; used around the flush (once per track and snapshot).  The staging area only
// ever holds finite numbers afterwards too (finite sums).
template <class C>
SRB_HD void pair_mma_store_frag(WarpSmem<C>& sm, int lane, const ThreadState<C>& st) {
  static_assert(sizeof(WarpSmem<C>) >= 32 * C::NACC * sizeof(typename C::TM), "staging area too small for the transpose");
  typename C::TM* buf = reinterpret_cast<typename C::TM*>(&sm);
  const int ks = lane & 3, b = lane >> 2;
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int t = 0; t < C::NT; t++)
#pragma unroll
      for (int u = 0; u < 4; u++) buf[(8 * a + b) * C::NACC + (4 * t + ks) * 4 + u] = st.acc[mma_acc_index<C>(a, t, u)];
}
template <class C>
SRB_HD void pair_mma_load_frag(const WarpSmem<C>& sm, int lane, ThreadState<C>& st) {
  const typename C::TM* buf = reinterpret_cast<const typename C::TM*>(&sm);
  const int ks = lane & 3, b = lane >> 2;
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int t = 0; t < C::NT; t++)
#pragma unroll
      for (int u = 0; u < 4; u++) st.acc[mma_acc_index<C>(a, t, u)] = buf[(8 * a + b) * C::NACC + (4 * t + ks) * 4 + u];
}
template <class C>
SRB_HD void pair_mma_load_tile(const WarpSmem<C>& sm, int lane, ThreadState<C>& st) {
  const typename C::TM* buf = reinterpret_cast<const typename C::TM*>(&sm);
#pragma unroll
  for (int k = 0; k < C::NACC; k++) st.acc[k] = buf[lane * C::NACC + k];
}
template <class C>
SRB_HD void pair_mma_store_tile(WarpSmem<C>& sm, int lane, const ThreadState<C>& st) {
  typename C::TM* buf = reinterpret_cast<typename C::TM*>(&sm);
#pragma unroll
  for (int k = 0; k < C::NACC; k++) buf[lane * C::NACC + k] = st.acc[k];
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
