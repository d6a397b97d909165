// synchrad_b200 — LITERAL single-precision kernels (dtype = SRB_DTYPE_F32_LITERAL).
//
// The default 'float' mode is mixed precision (srb_core.cuh).  This file provides the other reading of
// "follow the reference's dtype option": every operation of kernel_farfield.cl:30-108 /
// kernel_nearfield.cl:29-103 carried out in fp32, in the reference's order, without FMA contraction,
// including the PER-NODE Nyquist guard on fp32-rounded phases (which is not monotone in omega at fp32
// resolution, so it cannot be hoisted) and the per-step FormFactor multiply of cartesian_complex.
// Inputs arrive as float64 arrays and are rounded to fp32 on load — the reference's
// `astype(float32)` (calc.py:585-597); the tables must already hold fp32-representable values computed
// the way `_init_data` computes them (NumPy float32 sin/cos, calc.py:494-512).
// It agrees with the strict fp32 oracle to ~1e-6 (only sinf/cosf differ, by an ulp) and is as far from
// the fp64 answer as the reference's own single-precision path is (5.7 % of max on its undulator test).
// Work is still hoisted where that is bit-neutral: tau, beta, the amplitude vector are computed once
// per (direction, step) in the prep phase, identically to what every node of the reference computes.
#pragma once
#include "srb_core.cuh"
#include "srb_pair.cuh"
#include "srb_drec.cuh"

namespace srb {

SRB_HD float fm(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  return a * b;
#endif
}
SRB_HD float fa(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fadd_rn(a, b);
#else
  return a + b;
#endif
}
SRB_HD float fs(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fsub_rn(a, b);
#else
  return a - b;
#endif
}
SRB_HD float fd(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fdiv_rn(a, b);
#else
  return a / b;
#endif
}
SRB_HD float fsq(float a) {
#if defined(__CUDA_ARCH__)
  return __fsqrt_rn(a);
#else
  return sqrtf(a);
#endif
}
SRB_HD float fdot3(float ax, float ay, float az, float bx, float by, float bz) {
  return fa(fa(fm(ax, bx), fm(ay, by)), fm(az, bz));
}

// prep phase, lane = step: rec = far {A0,A1,A2,tau} ; near {D0,D1,D2,N0,N1,N2,tau,rInv}
template <class C>
SRB_HD void lit_prep_phase(const Params& P, const Geom& g, const TrackView& tv, uint32_t itBase, int cnt,
                           int lane, WarpSmem<C>& sm) {
  if (lane >= cnt) return;
  const uint32_t it = itBase + (uint32_t)lane;
  const float dt = (float)P.dt;
  const float time = fm((float)(tv.itStart + it), dt);
  const float x = (float)ldv<double>(tv.x, it), y = (float)ldv<double>(tv.y, it), z = (float)ldv<double>(tv.z, it);
  float u0 = (float)ldv<double>(tv.ux, it), u1 = (float)ldv<double>(tv.uy, it), u2 = (float)ldv<double>(tv.uz, it);
  if (C::MODE == MODE_FAR) {
    const float nx = (float)g.nx, ny = (float)g.ny, nz = (float)g.nz;
    const float tau = fs(time, fdot3(x, y, z, nx, ny, nz));
    float v0 = (float)ldv<double>(tv.ux, it + 1), v1 = (float)ldv<double>(tv.uy, it + 1), v2 = (float)ldv<double>(tv.uz, it + 1);
    float gi = fd(1.0f, fsq(fa(1.0f, fdot3(u0, u1, u2, u0, u1, u2))));
    u0 = fm(u0, gi); u1 = fm(u1, gi); u2 = fm(u2, gi);
    gi = fd(1.0f, fsq(fa(1.0f, fdot3(v0, v1, v2, v0, v1, v2))));
    v0 = fm(v0, gi); v1 = fm(v1, gi); v2 = fm(v2, gi);
    const float dtInv = fd(1.0f, dt);
    const float a0 = fm(fs(v0, u0), dtInv), a1 = fm(fs(v1, u1), dtInv), a2 = fm(fs(v2, u2), dtInv);
    const float b0 = fm(0.5f, fa(v0, u0)), b1 = fm(0.5f, fa(v1, u1)), b2 = fm(0.5f, fa(v2, u2));
    float c1 = fdot3(a0, a1, a2, nx, ny, nz);
    float c2 = fs(1.0f, fdot3(b0, b1, b2, nx, ny, nz));
    c2 = fd(1.0f, c2);
    c1 = fm(fm(c1, c2), c2);
    float A0 = fs(fm(c1, fs(nx, b0)), fm(c2, a0));
    float A1 = fs(fm(c1, fs(ny, b1)), fm(c2, a1));
    float A2 = fs(fm(c1, fs(nz, b2)), fm(c2, a2));
    if (P.comp == COMP_SPH || P.comp == COMP_SPH_CPLX) {
      const float s0 = fdot3(nx, ny, nz, A0, A1, A2);
      const float s1 = fdot3((float)g.tx, (float)g.ty, (float)g.tz, A0, A1, A2);
      const float s2 = fdot3((float)g.px, (float)g.py, (float)g.pz, A0, A1, A2);
      A0 = s0; A1 = s1; A2 = s2;
    }
    sm.rec[lane][0] = A0; sm.rec[lane][1] = A1; sm.rec[lane][2] = A2; sm.rec[lane][3] = tau;
  } else {
    const float r0 = fs((float)g.nx, x), r1 = fs((float)g.ny, y), r2 = fs((float)g.nz, z);
    const float rL = fsq(fdot3(r0, r1, r2, r0, r1, r2));
    const float tau = fa(time, rL);
    const float rInv = fd(1.0f, rL);
    const float n0 = fm(rInv, r0), n1 = fm(rInv, r1), n2 = fm(rInv, r2);
    const float gi = fd(1.0f, fsq(fa(1.0f, fdot3(u0, u1, u2, u0, u1, u2))));
    u0 = fm(u0, gi); u1 = fm(u1, gi); u2 = fm(u2, gi);
    const float ri2 = fm(rInv, rInv);
    sm.rec[lane][0] = fs(u0, n0); sm.rec[lane][1] = fs(u1, n1); sm.rec[lane][2] = fs(u2, n2);
    sm.rec[lane][3] = fm(ri2, n0); sm.rec[lane][4] = fm(ri2, n1); sm.rec[lane][5] = fm(ri2, n2);
    sm.rec[lane][6] = tau; sm.rec[lane][7] = rInv;
  }
}

// sin/cos of an fp32 phase (the value the reference feeds to sin()/cos()): range reduction to [-pi, pi], then the SFU
// approximations (abs error 2^-21.4 = 4e-7 on that range).  |x| < 48000: three-term Cody-Waite reduction in fp32
// (2*pi = 6.28125 + 1.9350052e-3 + 3.0199e-7, q*C exact for the first two; error 1.2e-7); larger arguments (the near
// field's 1e10 rad phases) are reduced exactly on the FP64 pipe.  2 MUFU + 7 FP32 ops against ~25 for sincosf(), no
// conversions on the common path.  The error is far below the fp32 phase noise the mode reproduces (|phase| * 6e-8)
// and the 1e-4 tolerance; measured against the reference's fp32 vectors: see tests.
template <bool BIG>
SRB_HD void lit_sincos(float x, float* sn, float* cs) {
#if defined(__CUDA_ARCH__)
  float r;
  if (!BIG) {
    const float q = __fadd_rn(fmaf(x, 0.15915494309189535f, 12582912.0f), -12582912.0f);
    r = fmaf(q, -6.28125f, x);
    r = fmaf(q, -1.9350051879882812e-3f, r);
    r = fmaf(q, -3.019916050561733e-7f, r);
  } else {
    const double xd = (double)x;
    const double t = fma(xd, 0.15915494309189535, 6755399441055744.0);
    const double qd = t - 6755399441055744.0;
    r = (float)fma(qd, -6.283185307179586, xd);
  }
  *sn = __sinf(r);
  *cs = __cosf(r);
#else
  *sn = sinf(x); *cs = cosf(x);
#endif
}

// the accumulation of one step for the lane's TW nodes: one straight-line block (eight independent sincos/accumulate
// chains per lane); BIG: some phase of the warp is beyond the fp32 Cody-Waite range
template <class C, bool BIG>
SRB_HD void lit_accumulate(const float* R, const float* ph, uint32_t pass, bool useFF, ThreadState<C>& st) {
  constexpr int TW = C::TW;
#pragma unroll
  for (int k = 0; k < TW; k++) {
    float sn, cs;
    lit_sincos<BIG>(ph[k], &sn, &cs);
    const bool on = (pass >> k) & 1u;
    sn = on ? sn : 0.0f; cs = on ? cs : 0.0f;              // a failed node adds exactly 0 (amplitudes are finite)
    if (C::MODE == MODE_FAR) {
      if (useFF) { sn = fm(sn, st.ff[k]); cs = fm(cs, st.ff[k]); }
#pragma unroll
      for (int c = 0; c < 3; c++) {
        st.acc[k * 6 + c] = fmaf(R[c], cs, st.acc[k * 6 + c]);
        st.acc[k * 6 + 3 + c] = fmaf(R[c], sn, st.acc[k * 6 + 3 + c]);
      }
    } else {
      const float wr = fm(st.wl[k], R[7]);
      const float ws = fm(wr, sn), wc = fm(wr, cs);
#pragma unroll
      for (int c = 0; c < 3; c++) {
        st.acc[k * 6 + c] = fmaf(R[3 + c], cs, fmaf(-R[c], ws, st.acc[k * 6 + c]));
        st.acc[k * 6 + 3 + c] = fmaf(R[3 + c], sn, fmaf(R[c], wc, st.acc[k * 6 + 3 + c]));
      }
    }
  }
}

// main phase, lane = interleaved tile {lane + 32k}: the reference's per-node loop body.  What is kept literally is
// everything that shapes the result at the 1e-4 level: tracks, tables, tau, the amplitude and the PHASE in fp32
// (rounded product, kernel_farfield.cl:65-67) and the per-node Nyquist guard on those phases (:68-72).  The
// accumulation uses fused multiply-adds (OpenCL leaves the contraction of `Re += A*cos` to the driver, Q8) and
// sin/cos of the fp32 phase come from the SFU (lit_sincos) instead of libm.
template <class C>
SRB_HD void lit_main_phase(const Params& P, const Geom& g, const WarpSmem<C>& sm, int cnt, int lane,
                           ThreadState<C>& st) {
  constexpr int TW = C::TW;
  const bool useFF = (C::MODE == MODE_FAR && P.comp == COMP_CART_CPLX && P.formFactor != nullptr);
  const float PI_F = (float)3.14159265358979323846;
  // the lane's nodes are j = cLo + lane + 32k: its first nValid = ceil((n - lane) / 32) of them exist
  const int nNodes = (int)(g.cHi - g.cLo);
  int nValid = lane < nNodes ? (nNodes - 1 - lane) / 32 + 1 : 0;
  nValid = nValid < TW ? nValid : TW;
#if defined(__CUDA_ARCH__)
  asm volatile("" : "+r"(nValid));      // (keeps ptxas from re-deriving it from eight predicates inside the loop)
#endif
  uint32_t nPass = 0;
  for (int s = 0; s < cnt; s++) {
    float R[8];
#pragma unroll
    for (int k = 0; k < C::NREC; k++) R[k] = sm.rec[s][k];
    // the guard of the lane's nodes first (branch-free), then one straight-line block for all of them: eight
    // independent sincos/accumulate chains per lane instead of eight basic blocks executed one after the other
    float ph[TW];
    uint32_t pass = 0u;
    bool big = false;
#pragma unroll
    for (int k = 0; k < TW; k++) {
      ph[k] = fm(st.wl[k], C::MODE == MODE_FAR ? R[3] : R[6]);
      const float dPhase = fabsf(fs(ph[k], st.pprev[k]));
      st.pprev[k] = ph[k];
      pass |= (k < nValid && dPhase < PI_F) ? (1u << k) : 0u;
      big = big || !(fabsf(ph[k]) < 48000.0f);
    }
    nPass += (uint32_t)__builtin_popcount(pass);
#if defined(__CUDA_ARCH__)
    if (!__any_sync(0xffffffffu, pass != 0u)) continue;      // (guard-dominated inputs: whole steps fail)
    big = __any_sync(0xffffffffu, big);
#else
    if (!pass) continue;
#endif
    if (big) lit_accumulate<C, true>(R, ph, pass, useFF, st);
    else lit_accumulate<C, false>(R, ph, pass, useFF, st);
  }
  st.nPass += nPass;
  st.nAll += (unsigned long long)(nValid * cnt);
}

template <class C>
SRB_HD void lit_flush_lane(const Params& P, const Geom& g, const TrackView& tv, uint32_t pc, uint32_t iSnap,
                           int lane, const ThreadState<C>& me) {
  constexpr int TW = C::TW;
  const size_t nTotal = (size_t)P.nOmega * P.nA2 * P.nPhi;
  const bool cplx = (P.comp == COMP_CART_CPLX || P.comp == COMP_SPH_CPLX);
  const float wp = (float)tv.w, dt = (float)P.dt;
  const float wpdt2 = fm(fm(wp, dt), dt);
  const float wpdt = fm(fsq(wp), dt);
#pragma unroll
  for (int k = 0; k < TW; k++) {
    const uint32_t j = g.cLo + (uint32_t)(lane + 32 * k);
    if (j >= g.cHi) continue;
    const size_t idx = (size_t)j + (size_t)P.nOmega * (g.iA2 + (size_t)P.nA2 * g.iPhi) + nTotal * iSnap;
    const float* re = &me.acc[k * 6];
    const float* im = &me.acc[k * 6 + 3];
    if (!cplx) {
      if (P.comp == COMP_TOTAL) {
        dest<C>(P, pc, 0)[idx] += (double)fm(wpdt2, fa(fdot3(re[0], re[1], re[2], re[0], re[1], re[2]),
                                                       fdot3(im[0], im[1], im[2], im[0], im[1], im[2])));
      } else {
#pragma unroll
        for (int c = 0; c < 3; c++)
          dest<C>(P, pc, c)[idx] += (double)fm(wpdt2, fa(fm(re[c], re[c]), fm(im[c], im[c])));
      }
    } else {
#pragma unroll
      for (int c = 0; c < 3; c++) {
        dest<C>(P, pc, 2 * c)[idx] += (double)fm(wpdt, re[c]);
        dest<C>(P, pc, 2 * c + 1)[idx] += (double)fm(wpdt, im[c]);
      }
    }
  }
}

}  // namespace srb
