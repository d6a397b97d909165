// synchrad_b200 — warp-level core of the spectral-integration kernels (sm_100a).
//
// Replaces the per-node OpenCL loops of the reference (kernel_farfield.cl:30-108 and the four
// epilogue variants; kernel_nearfield.cl:29-103 and two variants) with a batched, warp-autonomous
// formulation.  This is NOT a translation of those kernels:
//
//   reference: one work-item per (omega,theta,phi) node, serial over steps, one launch per
//              particle, every node recomputes beta/gamma/acceleration and calls sin+cos.
//   here:      one WARP owns a "virtual direction" (theta|R, phi, omega-chunk) for a whole slice of
//              particles.  Steps are processed in sub-batches of 32:
//                 prep phase  (lane = time step): everything that does not depend on omega —
//                     tau = t - n.r, the Lienard-Wiechert vector A, the Nyquist cut-off range and,
//                     for uniform omega grids, phasor seeds for 16 omega-tiles — is computed ONCE per
//                     (direction, step), in the reference's exact operation order and without FMA
//                     contraction (SURVEY §7 "hard parts"), and staged in shared memory;
//                 main phase  (lane = omega tile; a tile is the INTERLEAVED node set {m, m+T, m+2T, ...}
//                     of its chunk, T = tiles per chunk, so that a Nyquist cut-off [0, j*) spreads
//                     evenly over the lanes): each thread sweeps its tile with a three-term
//                     phasor recurrence v[k+1] = 2cos(d)*v[k] - v[k-1] (cos and sin parts live in
//                     different lanes) and accumulates A*v into registers; for total/cartesian only the two
//                     transverse components of A are carried (A is orthogonal to n): 6 FP64 issue slots
//                     per (particle,step,node) update instead of ~30 with a per-node sincos.
//              Non-uniform grids ('wavelengthGrid', 'logGrid', calc.py:390-399) use the DIRECT
//              main phase (per-node sincos of the same rounded phase the reference forms).
//              |A|^2*w is added to the spectrum once per (track, snapshot) by the owning thread —
//              no per-step global traffic, no atomics (deterministic).
//
// The fastest far-field main phase (symmetric node pairs, no per-lane recurrence) lives in srb_pair.cuh;
// an all-fp32 reproduction of the reference's single-precision kernels in srb_literal.cuh.
//
// The file is written so that the same code can be compiled by g++ for a single-warp CPU
// emulation (tests/emu/, test infrastructure only: it lets the kernel logic be debugged in a
// container without a GPU).  The product never runs it on the CPU.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <type_traits>

#if defined(__CUDACC__)
#define SRB_HD __host__ __device__ __forceinline__
#else
#define SRB_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define SRB_LANES_BEGIN { const int lane = (int)(threadIdx.x & 31u);
#define SRB_LANES_END } __syncwarp();
#define SRB_ST (st[0])
#else
#define SRB_LANES_BEGIN for (int lane = 0; lane < 32; ++lane) {
#define SRB_LANES_END }
#define SRB_ST (st[lane])
#endif

namespace srb {

enum { MODE_FAR = 0, MODE_NEAR = 1 };
enum { COMP_TOTAL = 0, COMP_CART = 1, COMP_CART_CPLX = 2, COMP_SPH = 3, COMP_SPH_CPLX = 4 };
enum { KIND_DIRECT = 0, KIND_RECUR = 1, KIND_LITERAL = 2, KIND_PAIR = 3, KIND_PAIR_FMA = 4, KIND_DREC = 5 };   // LITERAL: srb_literal.cuh, PAIR: srb_pair.cuh
// KIND_DREC (srb_drec.cuh): the direct kernel's layout with the per-node sincos replaced by a per-lane recurrence along
// omega plus a per-update first-order correction onto the reference's ROUNDED phase -- phases of any magnitude (near
// field at large L: 1e10 rad) on uniform grids, fp64.
// (kind 5 was the experimental gridding / type-1 NUFFT kernel of round 1: correct but 0.75-0.81x the pair kernel after two
//  iterations, profiles/r02_spread_v2_decision.txt; removed from the library, kept on the git branch `spread-kernel`)
// KIND_PAIR_FMA: the pair kernel with its accumulation on the scalar FP64 pipe (DFMA) where KIND_PAIR uses DMMA
constexpr int SUB = 32;  // steps per sub-batch (= lanes of the prep phase)

// ---- strict (uncontracted, round-to-nearest) double arithmetic: the oracle's operation order
SRB_HD double smul(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}
SRB_HD double sadd(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}
SRB_HD double ssub(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __dsub_rn(a, b);
#else
  return a - b;
#endif
}
SRB_HD double sdiv(double a, double b) {
#if defined(__CUDA_ARCH__)
  return __ddiv_rn(a, b);
#else
  return a / b;
#endif
}
SRB_HD double ssqrt(double a) {
#if defined(__CUDA_ARCH__)
  return __dsqrt_rn(a);
#else
  return sqrt(a);
#endif
}
SRB_HD double sdot3(double ax, double ay, double az, double bx, double by, double bz) {
  return sadd(sadd(smul(ax, bx), smul(ay, by)), smul(az, bz));  // OpenCL dot(), Q8 association
}
SRB_HD void sincos_d(double x, double* s, double* c) {
#if defined(__CUDA_ARCH__)
  sincos(x, s, c);
#else
  *s = sin(x); *c = cos(x);
#endif
}
// v with its sign bit XOR-ed by mask (0 or 0x80000000): a lane-dependent sign without an FP64-pipe op
SRB_HD double flipsign(double v, uint32_t mask) {
#if defined(__CUDA_ARCH__)
  return __hiloint2double(__double2hiint(v) ^ (int)mask, __double2loint(v));
#else
  return mask ? -v : v;
#endif
}
SRB_HD float flipsign(float v, uint32_t mask) {
#if defined(__CUDA_ARCH__)
  return __int_as_float(__float_as_int(v) ^ (int)mask);
#else
  return mask ? -v : v;
#endif
}
// sin/cos of a double of ANY magnitude (|x| < 2^50) at fast-path cost, FP64 pipe only (no FRND /
// F2I conversions, which run at 1/4 rate).  Near-field phases omega*(t+R) reach 1e10 rad and
// SI-unit far-field phases 1e6, far beyond the fast path of CUDA's sincos().
//   x/pi = h + l as an exact double-double product; q = rint(2(h+l)) by the magic-number trick
//   (the quadrant is read from the mantissa); r = (h - q/2) + l in [-1/4, 1/4]; y = pi*r;
//   fdlibm's __kernel_sin/__kernel_cos minimax polynomials on [-pi/4, pi/4].
// Max abs error 2.0e-16 for |x| up to 1e13 (checked against sinl/cosl, tests/test_emulated_kernels.py).
SRB_HD void sincos_big(double x, double* sn, double* cs) {
  const double h = smul(x, 0.3183098861837907);
  double l = fma(x, 0.3183098861837907, -h);
  l = fma(x, -1.9678676675182486e-17, l);
  const double t = fma(h, 2.0, 6755399441055744.0);
  const double qd = ssub(t, 6755399441055744.0);
#if defined(__CUDA_ARCH__)
  const uint32_t q = (uint32_t)__double2loint(t);
#else
  uint64_t tb; memcpy(&tb, &t, 8);
  const uint32_t q = (uint32_t)tb;
#endif
  const double r = sadd(fma(qd, -0.5, h), l);
  const double y = fma(r, 3.141592653589793, smul(r, 1.2246467991473532e-16));
  const double z = smul(y, y);
  double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
  ps = fma(z, ps, 2.75573137070700676789e-06);
  ps = fma(z, ps, -1.98412698298579493134e-04);
  ps = fma(z, ps, 8.33333333332248946124e-03);
  ps = fma(z, ps, -1.66666666666666324348e-01);
  const double s = fma(smul(y, z), ps, y);
  double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
  pc = fma(z, pc, -2.75573143513906633035e-07);
  pc = fma(z, pc, 2.48015872894767294178e-05);
  pc = fma(z, pc, -1.38888888888741095749e-03);
  pc = fma(z, pc, 4.16666666666666019037e-02);
  const double c = fma(smul(z, z), pc, fma(z, -0.5, 1.0));
  const double a = (q & 1u) ? c : s, b = (q & 1u) ? s : c;
  *sn = flipsign(a, (q & 2u) << 30);
  *cs = flipsign(b, ((q + 1u) & 2u) << 30);
}
SRB_HD void sincos_t(double x, double* s, double* c) { sincos_big(x, s, c); }
SRB_HD double tmul(double a, double b) { return smul(a, b); }
SRB_HD float tmul(float a, float b) {
#if defined(__CUDA_ARCH__)
  return __fmul_rn(a, b);
#else
  return a * b;
#endif
}
SRB_HD void sincos_t(float x, float* s, float* c) {
#if defined(__CUDA_ARCH__)
  sincosf(x, s, c);
#else
  *s = sinf(x); *c = cosf(x);
#endif
}
// (fp32 'native' path, calc.py:612-615 -> native_sin/native_cos: see sincos_mixed<true>)
// fp32 main phase of the DIRECT kind (mixed precision): the phase omega*tau is formed in fp64 (a phase rounded to
// fp32 is already 2^-24*|phase| = 1e-4 rad off at 2000 rad) and reduced there -- 1 DMUL + 4 DFMA on the FP64 pipe,
// which this kernel leaves idle -- to r in [-pi/4, pi/4] + quadrant; sin/cos of r are fp32 minimax polynomials
// (Cephes sinf/cosf kernels, ~1e-7).  NATIVE: reduction to [-pi, pi], then the MUFU approximations.
template <bool NATIVE>
SRB_HD void sincos_mixed(double w, double tau, float* sn, float* cs) {
  const double ph = smul(w, tau);
#if defined(__CUDA_ARCH__)
  if (NATIVE) {
    const double t = fma(ph, 0.15915494309189535, 6755399441055744.0);
    const double qd = ssub(t, 6755399441055744.0);
    double r = fma(qd, -6.283185307179586, ph);
    r = fma(qd, -2.4492935982947064e-16, r);
    __sincosf((float)r, sn, cs);
    return;
  }
#endif
  const double t = fma(ph, 0.6366197723675814, 6755399441055744.0);
  const double qd = ssub(t, 6755399441055744.0);
#if defined(__CUDA_ARCH__)
  const uint32_t q = (uint32_t)__double2loint(t);
#else
  uint64_t tb; memcpy(&tb, &t, 8);
  const uint32_t q = (uint32_t)tb;
#endif
  double rd = fma(qd, -1.5707963267948966, ph);
  rd = fma(qd, -6.123233995736766e-17, rd);
  const float r = (float)rd, z = r * r;
  float ps = fmaf(z, -1.9515295891e-4f, 8.3321608736e-3f);
  ps = fmaf(z, ps, -1.6666654611e-1f);
  const float s = fmaf(r * z, ps, r);
  float pc = fmaf(z, 2.443315711809948e-5f, -1.388731625493765e-3f);
  pc = fmaf(z, pc, 4.166664568298827e-2f);
  const float c = fmaf(z * z, pc, fmaf(z, -0.5f, 1.0f));
  const float a = (q & 1u) ? c : s, b = (q & 1u) ? s : c;
  *sn = flipsign(a, (q & 2u) << 30);
  *cs = flipsign(b, ((q + 1u) & 2u) << 30);
}

// ---- kernel parameter block (plain data, passed by value)
struct Params {
  // grid (calc.py:486-512 tables, in the compute dtype TI)
  int32_t mode, comp;
  uint32_t nOmega, nA2, nPhi, nSnaps;
  const void *omega, *axA, *axB, *sinPhi, *cosPhi, *formFactor;
  double L, dt;
  int32_t descending;   // omega table is descending ('wavelengthGrid')
  double domega;        // uniform spacing of the 2*pi*omega table (recurrence kernels only)
  uint32_t chunkNodes;  // omega nodes per virtual direction
  uint32_t nChunks;
  uint32_t nVD;         // nPhi * nA2 * nChunks
  // tracks (SoA, concatenated)
  uint32_t nTracks;
  const void *x, *y, *z, *ux, *uy, *uz;
  const uint64_t* offsets;  // [nTracks+1]
  const void* w;            // [nTracks] TI
  const uint32_t *itStart, *itEnd, *itSnaps;
  uint32_t snapStride;      // 0: one itSnaps[nSnaps] for all tracks, else per-track stride
  // output: fp64 spectra in the reference's device layout (nSnaps, nPhi, nA2, nOmega)
  double* out[6];
  double* slabs;            // (nPC-1) private partial spectra, reduced afterwards
  size_t slabStride;        // doubles per slab = nOut*nSnaps*nTotal
  uint32_t nPC;             // particle chunks
  unsigned long long* counters;  // [0] passed updates, [1] all updates (may be null)
  // optional per-step pre-pass (direction independent, computed once per call by k_prepass):
  // SoA planes of `preStride` doubles; far: a0,a1,a2,b0,b1,b2 (acceleration, mean beta),
  // near: beta0,beta1,beta2.  nullptr -> the prep phase computes them itself.
  const double* pre;
  uint64_t preStride;
  int32_t tmaOK;        // pre is 16-byte aligned (TMA staging of srb_ws.cuh)
  int32_t prePacked;    // pre holds one 9-double record per step: x, y, z, a0..a2, b0..b2 (warp-specialised kernel)
  // phasor = AUTO with two eligible kernels: both are launched and the one the device-side guard probe did not
  // select returns at once (sel == nullptr: unconditional)
  const int32_t* sel;
  int32_t selWant;
  // time-axis split (few-particle configurations): every track is cut into nTS segments of whole 32-step sub-batches,
  // particle chunk pc = track * nTS + segment; a flush stores the segment's partial complex amplitudes in `amp`
  // [(pc * nSnaps + iSnap) * 2 NCF + (Re c | NCF + Im c)][node] and combine_node() squares their sum.  nTS <= 1: off.
  uint32_t nTS;
  double* amp;
};
SRB_HD bool deselected(const Params& P) { return P.sel != nullptr && *P.sel != P.selWant; }

// NC: amplitude components carried per node in far-field mode.
//   3 — Cartesian (or, for the spheric kernels, the (n, e_theta, e_phi) projections);
//   2 — transverse basis (e_theta, e_phi): the Lienard-Wiechert vector A = c1 (n - beta) - c2 a is
//       orthogonal to n identically (A.n = c1/c2 - c2 (a.n) = 0; in floating point the residue is
//       ~1 ulp of |A|), so two components carry everything and the Cartesian ones are rebuilt at
//       flush time as F = F_theta e_theta + F_phi e_phi.  Used for total / cartesian(_complex).
template <class TI_, class TM_, int MODE_, int KIND_, int TW_, bool NATIVE_, int NC_ = 3>
struct Cfg {
  using TI = TI_;   // dtype of tables / tracks
  using TM = TM_;   // arithmetic type of the main phase
  static constexpr int MODE = MODE_, KIND = KIND_, TW = TW_, NC = NC_;
  static constexpr bool PAIR = KIND_ == KIND_PAIR || KIND_ == KIND_PAIR_FMA;
  static constexpr bool NATIVE = NATIVE_;
  static constexpr int TILES = (KIND_ == KIND_RECUR) ? 16 : 32;   // omega tiles per chunk
  static constexpr int CHUNK = TILES * TW_;
  static constexpr int NV = (MODE_ == MODE_FAR) ? NC_ : 6;        // per-step vector entries in `rec`
  // accumulators per node: split layout (recurrence) holds one part (cos or sin) of NV sums,
  // the direct layout holds Re and Im of the 3 (far: NC) amplitude components
  static constexpr int NPN = (KIND_ == KIND_RECUR) ? NV
      : ((MODE_ == MODE_FAR && (KIND_ == KIND_DIRECT || KIND_ == KIND_DREC || PAIR)) ? 2 * NC_ : 6);
  static constexpr int NACC = NPN * TW_;
  // rec row: V[NV], then (recurrence) 2cos(d), cos(d), sin(d) | (direct) tau (fp32: tau_hi, tau_lo) ; padded to even
  //          (pair) V[NC], tau (flag 3 only), pad to QOFF, then A_c*(cos,sin) of the TW/2 pair offsets for
  //          every component c: [QOFF + 2(p*NC + c)]; padded to a multiple of 4 (16-byte rows in fp32)
  // fp64 pair kernel on the FP64 tensor cores (DMMA.8x8x4, srb_pair.cuh): the accumulation is a GEMM
  //   U[(tile, Re|Im X), (pair, component, cos|sin)] += X[(tile), step] * Q'[step, (pair, component)]
  // with K = steps; it needs the column count TW*NC to be a multiple of 8 (n-tiles of the MMA).
  static constexpr bool MMA = KIND_ == KIND_PAIR && sizeof(TM_) == 8 && (TW_ * NC_) % 8 == 0;
  static constexpr int NT = MMA ? TW_ * NC_ / 8 : 0;            // 8-column tiles of Q'
  static constexpr int QOFF = 4;
  static constexpr int NSEED = (KIND_ == KIND_RECUR) ? 32 : (PAIR ? 24 : 1);
  // the tensor-core layout transposes its accumulators through the staging area at a flush: 32*NACC values
  // must fit in rec + rng + seeds (only the 16-node tiles need the extra row padding)
  static constexpr int NREC_PAIR = (QOFF + NC_ * TW_ + 3) & ~3;
  static constexpr int NREC_FLUSH = MMA ? (((32 * NACC - 16 - NSEED * 33 + 31) / 32 + 3) & ~3) : 0;
  static constexpr int NREC = KIND_ == KIND_LITERAL ? (MODE_ == MODE_FAR ? 4 : 8)
      : PAIR ? (NREC_PAIR > NREC_FLUSH ? NREC_PAIR : NREC_FLUSH)
      : KIND_ == KIND_DREC ? ((NV + 5 + 1) & ~1)   // V[NV], tau, R^32 (Re, Im), conj-rotated... see srb_drec.cuh: tau, wr, wi, 2cos, pad
      : KIND_ == KIND_RECUR ? (sizeof(TM_) == 4 ? ((NV + 3 + 3) & ~3) : ((NV + 3 + 1) & ~1))   // 16-byte rows: rec_row() reads them with 128-bit loads
      : (((NV + (sizeof(TM_) == 4 ? 2 : 1)) + 1) & ~1);   // direct, fp32: tau as hi + lo
  static constexpr bool DIRECTLIKE = KIND_ == KIND_DIRECT || KIND_ == KIND_DREC;         // lane = tile {lane + 32k}, Re and Im per node
  using TW_T = typename std::conditional<DIRECTLIKE, double, TM_>::type;                 // type of the lane's omega nodes
};

// KIND_DREC: per step the two-level phasor seeds of the 32 lanes' first nodes m = 8a + b: Z_b = E R^b (b = 0..7) and
// Y_a = R^(8a) (a = 0..3), S_m = Y_a Z_b; rows padded to 13 pairs (lane = step 128-bit stores conflict free)
struct alignas(16) Cpx { double re, im; };
template <bool ON> struct DrecSmem {};
template <> struct alignas(16) DrecSmem<true> { Cpx seed[SUB][13]; };

template <class C>
struct alignas(16) WarpSmem : DrecSmem<C::KIND == KIND_DREC> {
  typename C::TM rec[SUB][C::NREC];       // per-step record (see Cfg::NREC)
  uint32_t rng[SUB];                      // lo | hi<<10 | flag<<30 (chunk-relative pass range)
  typename C::TM seeds[C::NSEED][SUB + 1];  // [part*16+tile][step]: cos|sin of the tile's first node
};

template <class C>
struct ThreadState {
  typename C::TM acc[C::NACC];
  typename C::TW_T wl[(C::DIRECTLIKE || C::KIND == KIND_LITERAL) ? C::TW : 1];   // this lane's omega nodes
  double eps[C::KIND == KIND_DREC ? C::TW : 1];   // KIND_DREC: table node minus ideal uniform-grid node (double-double residue)
  typename C::TM pprev[C::KIND == KIND_LITERAL ? C::TW : 1];   // literal kind: per-node phasePrev
  typename C::TM ff[C::KIND == KIND_LITERAL ? C::TW : 1];      // literal kind: per-node FormFactor
  unsigned long long nPass, nAll;
};

struct Geom {            // one virtual direction
  uint32_t iPhi, iA2, cLo, cHi;   // omega nodes [cLo, cHi)
  double nx, ny, nz;     // far: unit vector n ; near: point on the screen
  double tx, ty, tz, px, py, pz;  // far spheric: e_theta, e_phi
};

struct TrackView {
  const void *x, *y, *z, *ux, *uy, *uz;   // already offset to the track start
  const double* pre;                      // pre-pass planes offset to the track start, or null
  uint32_t n, itStart, itEnd;
  const uint32_t* snaps;
  double w;
};

// direction-independent part of a far-field step: beta_it, beta_it+1 -> acceleration a and
// mean beta b (kernel_farfield.cl:74-83), strict operation order
template <class TI>
SRB_HD void far_step_kinematics(const void* ux, const void* uy, const void* uz, size_t it, double dtInv,
                                double a[3], double b[3]) {
  double u0 = (double)((const TI*)ux)[it], u1 = (double)((const TI*)uy)[it], u2 = (double)((const TI*)uz)[it];
  double v0 = (double)((const TI*)ux)[it + 1], v1 = (double)((const TI*)uy)[it + 1], v2 = (double)((const TI*)uz)[it + 1];
  double gi = sdiv(1.0, ssqrt(sadd(1.0, sdot3(u0, u1, u2, u0, u1, u2))));
  u0 = smul(u0, gi); u1 = smul(u1, gi); u2 = smul(u2, gi);
  gi = sdiv(1.0, ssqrt(sadd(1.0, sdot3(v0, v1, v2, v0, v1, v2))));
  v0 = smul(v0, gi); v1 = smul(v1, gi); v2 = smul(v2, gi);
  a[0] = smul(ssub(v0, u0), dtInv); a[1] = smul(ssub(v1, u1), dtInv); a[2] = smul(ssub(v2, u2), dtInv);
  b[0] = smul(0.5, sadd(v0, u0)); b[1] = smul(0.5, sadd(v1, u1)); b[2] = smul(0.5, sadd(v2, u2));
}
// near field: beta of the single sample (kernel_nearfield.cl:78-81)
template <class TI>
SRB_HD void near_step_kinematics(const void* ux, const void* uy, const void* uz, size_t it, double b[3]) {
  const double u0 = (double)((const TI*)ux)[it], u1 = (double)((const TI*)uy)[it], u2 = (double)((const TI*)uz)[it];
  const double gi = sdiv(1.0, ssqrt(sadd(1.0, sdot3(u0, u1, u2, u0, u1, u2))));
  b[0] = smul(u0, gi); b[1] = smul(u1, gi); b[2] = smul(u2, gi);
}

// pair kind (srb_pair.cuh)
template <class C> SRB_HD void make_seeds_pair(const Params&, const Geom&, double, const double*, WarpSmem<C>&, int);
template <class C> SRB_HD void main_pair(const Params&, const Geom&, const WarpSmem<C>&, int, uint32_t, uint32_t, int, ThreadState<C>&);
#if defined(__CUDA_ARCH__)
template <class C> SRB_HD void main_pair_mma(const Params&, const Geom&, const WarpSmem<C>&, int, uint32_t, uint32_t, int, ThreadState<C>&);
#else
template <class C> inline void main_pair_mma(const Params&, const Geom&, const WarpSmem<C>&, int, uint32_t, uint32_t, ThreadState<C>*);
#endif
template <class C> SRB_HD void pair_mma_store_frag(WarpSmem<C>&, int, const ThreadState<C>&);
template <class C> SRB_HD void pair_mma_load_frag(const WarpSmem<C>&, int, ThreadState<C>&);
template <class C> SRB_HD void pair_mma_load_tile(const WarpSmem<C>&, int, ThreadState<C>&);
template <class C> SRB_HD void pair_mma_store_tile(WarpSmem<C>&, int, const ThreadState<C>&);
template <class C> SRB_HD void pair_node_amp(const ThreadState<C>&, int, double*, double*);
// corrected-recurrence kind (srb_drec.cuh)
template <class C> SRB_HD void drec_init_lane(const Params&, const Geom&, int, ThreadState<C>&);
template <class C> SRB_HD void make_seeds_drec(const Params&, const Geom&, double, WarpSmem<C>&, int, double*);
template <class C> SRB_HD void main_drec(const Params&, const Geom&, const WarpSmem<C>&, int, uint32_t, uint32_t, int, ThreadState<C>&);
// literal fp32 kind (srb_literal.cuh), used by warp_task below
template <class C> SRB_HD void lit_prep_phase(const Params&, const Geom&, const TrackView&, uint32_t, int, int, WarpSmem<C>&);
template <class C> SRB_HD void lit_main_phase(const Params&, const Geom&, const WarpSmem<C>&, int, int, ThreadState<C>&);
template <class C> SRB_HD void lit_flush_lane(const Params&, const Geom&, const TrackView&, uint32_t, uint32_t, int, const ThreadState<C>&);

template <class TI> SRB_HD double ldv(const void* p, size_t i) { return (double)((const TI*)p)[i]; }

// -------------------------------------------------------------------------------- prep phase
// Everything below is per (direction, step) and follows the oracle's operation order exactly
// (no contraction), so that for TI=double tau and the amplitude vector are bit-identical to the
// strict restatement of kernel_farfield.cl:65-94 / kernel_nearfield.cl:64-85.
// Lienard-Wiechert amplitude vector of one far-field step from the step's acceleration a and mean beta b
// (kernel_farfield.cl:84-94), in the basis the configuration carries (see Cfg::NC)
template <class C>
SRB_HD void far_amplitude(const Params& P, const Geom& g, const double a[3], const double b[3], double A[3]) {
  double c1 = sdot3(a[0], a[1], a[2], g.nx, g.ny, g.nz);
  double c2 = ssub(1.0, sdot3(b[0], b[1], b[2], g.nx, g.ny, g.nz));
  c2 = sdiv(1.0, c2);
  c1 = smul(smul(c1, c2), c2);
  const double A0 = ssub(smul(c1, ssub(g.nx, b[0])), smul(c2, a[0]));
  const double A1 = ssub(smul(c1, ssub(g.ny, b[1])), smul(c2, a[1]));
  const double A2 = ssub(smul(c1, ssub(g.nz, b[2])), smul(c2, a[2]));
  if (C::NC == 2) {              // transverse basis (see Cfg)
    A[0] = sdot3(g.tx, g.ty, g.tz, A0, A1, A2);
    A[1] = sdot3(g.px, g.py, g.pz, A0, A1, A2);
    A[2] = 0.0;
  } else if (P.comp == COMP_SPH || P.comp == COMP_SPH_CPLX) {   // kernel_farfield.cl:442-445
    A[0] = sdot3(g.nx, g.ny, g.nz, A0, A1, A2);
    A[1] = sdot3(g.tx, g.ty, g.tz, A0, A1, A2);
    A[2] = sdot3(g.px, g.py, g.pz, A0, A1, A2);
  } else { A[0] = A0; A[1] = A1; A[2] = A2; }
}

template <class C>
SRB_HD void prep_far(const Params& P, const Geom& g, const TrackView& tv, uint32_t it,
                     double dtInv, double A[3]) {
  using TI = typename C::TI;
  double a[3], b[3];
  if (tv.pre) {
#pragma unroll
    for (int c = 0; c < 3; c++) { a[c] = tv.pre[c * P.preStride + it]; b[c] = tv.pre[(3 + c) * P.preStride + it]; }
  } else far_step_kinematics<TI>(tv.ux, tv.uy, tv.uz, it, dtInv, a, b);
  far_amplitude<C>(P, g, a, b, A);
}

// tau of step it-1 (the reference's phasePrev/omega), or 0 for it == 0 (Q1)
template <class C>
SRB_HD double far_tau(const Params& P, const Geom& g, const TrackView& tv, uint32_t it) {
  using TI = typename C::TI;
  const double tp = smul((double)(tv.itStart + it), P.dt);
  return ssub(tp, sdot3(ldv<TI>(tv.x, it), ldv<TI>(tv.y, it), ldv<TI>(tv.z, it), g.nx, g.ny, g.nz));
}

template <class C>
SRB_HD void near_tau(const Params& P, const Geom& g, const TrackView& tv, uint32_t it,
                     double& tau, double& r0, double& r1, double& r2, double& rL) {
  using TI = typename C::TI;
  const double time = smul((double)(tv.itStart + it), P.dt);
  r0 = ssub(g.nx, ldv<TI>(tv.x, it)); r1 = ssub(g.ny, ldv<TI>(tv.y, it)); r2 = ssub(g.nz, ldv<TI>(tv.z, it));
  rL = ssqrt(sdot3(r0, r1, r2, r0, r1, r2));
  tau = sadd(time, rL);
}

// near: B = rInv*(beta - n), Cv = rInv^2 * n ; the reference's c1 = omega*B, c2 = Cv
template <class C>
SRB_HD void prep_near(const Params& P, const Geom& g, const TrackView& tv, uint32_t it,
                      double r0, double r1, double r2, double rL, double B[3], double Cv[3]) {
  using TI = typename C::TI;
  const double rInv = sdiv(1.0, rL);
  const double n0 = smul(rInv, r0), n1 = smul(rInv, r1), n2 = smul(rInv, r2);
  double u[3];
  if (tv.pre) {
#pragma unroll
    for (int c = 0; c < 3; c++) u[c] = tv.pre[c * P.preStride + it];
  } else near_step_kinematics<TI>(tv.ux, tv.uy, tv.uz, it, u);
  B[0] = smul(rInv, ssub(u[0], n0)); B[1] = smul(rInv, ssub(u[1], n1)); B[2] = smul(rInv, ssub(u[2], n2));
  const double ri2 = smul(rInv, rInv);
  Cv[0] = smul(ri2, n0); Cv[1] = smul(ri2, n1); Cv[2] = smul(ri2, n2);
}

// Nyquist guard (kernel_farfield.cl:68-72), hoisted: for fixed (direction, step) the test
// |fl(w_j*tau) - fl(w_j*tauPrev)| < pi is monotone in omega, so the passing nodes of a chunk form
// one interval.  Its end is found by bisection on the reference's EXACT rounded predicate, so the
// per-node decisions are reproduced.  Returns chunk-relative [lo, hi).
template <class C>
SRB_HD void pass_range(const Params& P, const Geom& g, double tau, double tauPrev,
                       uint32_t& lo, uint32_t& hi) {
  using TI = typename C::TI;
  const TI* om = (const TI*)P.omega;
  auto pass = [&](uint32_t j) -> bool {
    const double w = (double)om[j];
    return fabs(ssub(smul(w, tau), smul(w, tauPrev))) < 3.14159265358979323846;
  };
  const uint32_t n = g.cHi - g.cLo;
  const uint32_t jSmall = P.descending ? g.cHi - 1 : g.cLo;   // smallest omega of the chunk
  const uint32_t jLarge = P.descending ? g.cLo : g.cHi - 1;
  if (pass(jLarge)) { lo = 0; hi = n; return; }
  if (!pass(jSmall)) { lo = 0; hi = 0; return; }
  if (C::KIND == KIND_RECUR || C::PAIR) {
    // uniform ascending grid: estimate the boundary from pi/|dtau|, then settle it with the exact
    // predicate (a couple of evaluations instead of a full bisection)
    const double je = (3.14159265358979323846 / fabs(ssub(tau, tauPrev)) - (double)om[g.cLo]) / P.domega;
    uint32_t b = (uint32_t)fmin(fmax(ceil(je), 1.0), (double)(n - 1));
    while (b > 1 && !pass(g.cLo + b - 1)) b--;
    while (b < n - 1 && pass(g.cLo + b)) b++;
    lo = 0; hi = b; return;
  }
  // invariant: pass(a) true, pass(b) false, a and b chunk-relative positions ordered by omega
  uint32_t a = 0, b = n - 1;            // positions in ascending-omega order
  while (b - a > 1) {
    const uint32_t mid = (a + b) >> 1;
    const uint32_t j = P.descending ? g.cHi - 1 - mid : g.cLo + mid;
    if (pass(j)) a = mid; else b = mid;
  }
  if (P.descending) { lo = n - b; hi = n; } else { lo = 0; hi = b; }
}

// Phasor seeds for the 16 interleaved omega tiles of a chunk (uniform grid): tile m starts at chunk
// node m, X_m = exp(i(phi0 + m*d)), generated with the three-term recurrence in m; consecutive nodes
// of a tile are 16 grid steps apart, so the main phase advances with exp(i*16*d) (out[] = 2cos(16d),
// cos(16d), sin(16d)).  phi0 is the reference's own rounded phase at the chunk's first node.
template <class C>
SRB_HD void make_seeds(const Params& P, const Geom& g, double tau, WarpSmem<C>& sm, int s, double out[3]) {
  using TI = typename C::TI; using TM = typename C::TM;
  const double w0 = (double)((const TI*)P.omega)[g.cLo];
  double s0, c0, sd, cd;
  sincos_big(smul(w0, tau), &s0, &c0);
  sincos_big(P.domega * tau, &sd, &cd);
  double cw = cd, sw = sd;
#pragma unroll
  for (int i = 1; i < 16; i <<= 1) { const double t = cw * cw - sw * sw; sw = 2.0 * cw * sw; cw = t; }
  out[0] = 2.0 * cw; out[1] = cw; out[2] = sw;
  double xr0 = c0, xi0 = s0;
  double xr1 = c0 * cd - s0 * sd, xi1 = c0 * sd + s0 * cd;
  const double cf = 2.0 * cd;
  sm.seeds[0][s] = (TM)xr0;  sm.seeds[16][s] = (TM)xi0;
  sm.seeds[1][s] = (TM)xr1;  sm.seeds[17][s] = (TM)xi1;
#pragma unroll
  for (int m = 2; m < 16; m++) {
    const double xr2 = cf * xr1 - xr0, xi2 = cf * xi1 - xi0;
    sm.seeds[m][s] = (TM)xr2; sm.seeds[16 + m][s] = (TM)xi2;
    xr0 = xr1; xi0 = xi1; xr1 = xr2; xi1 = xi2;
  }
}

// The prep phase of one step (lane = step) in two parts.
//   prep_guard: tau, the Nyquist pass range and flag (0 nothing passes, 1 all nodes of the chunk, 2 some, 3 = 1|2 with a
//               phase too large for the seed arithmetic), the amplitude vector -- everything stays in registers;
//   prep_store: phasor seeds and the shared-memory record for the lane = tile main phases.
// Between the two the warp may decide to evaluate a guard-dominated sub-batch with lane = step (main_sparse below),
// which needs neither seeds nor records.
struct PrepStep { uint32_t flag, lo, hi; double tau, V[6]; };

template <class C>
SRB_HD void prep_guard(const Params& P, const Geom& g, const TrackView& tv, uint32_t itBase, int cnt,
                       double dtInv, int lane, ThreadState<C>& st, PrepStep& ps) {
  using TM = typename C::TM;
  ps.flag = ps.lo = ps.hi = 0u; ps.tau = 0.0;
#pragma unroll
  for (int k = 0; k < 6; k++) ps.V[k] = 0.0;
  if (lane >= cnt) return;
  const uint32_t it = itBase + (uint32_t)lane;
  double tau, tauPrev;
  double r0 = 0, r1 = 0, r2 = 0, rL = 1;
  if (C::MODE == MODE_FAR) tau = far_tau<C>(P, g, tv, it);
  else near_tau<C>(P, g, tv, it, tau, r0, r1, r2, rL);
  // tau of the previous step: the neighbouring lane has it (lanes [0,cnt) are all here)
#if defined(__CUDA_ARCH__)
  tauPrev = __shfl_up_sync(cnt >= 32 ? 0xffffffffu : ((1u << cnt) - 1u), tau, 1);
  if (lane == 0)
#endif
  {
    if (it == 0) tauPrev = 0.0;        // phasePrev starts at 0 (Q1)
    else if (C::MODE == MODE_FAR) tauPrev = far_tau<C>(P, g, tv, it - 1);
    else { double q0, q1, q2, qL; near_tau<C>(P, g, tv, it - 1, tauPrev, q0, q1, q2, qL); }
  }
  uint32_t lo, hi;
  pass_range<C>(P, g, tau, tauPrev, lo, hi);
  const uint32_t n = g.cHi - g.cLo;
  uint32_t flag = (hi <= lo) ? 0u : ((lo == 0 && hi == n) ? 1u : 2u);
  st.nAll += n;
  ps.tau = tau;
  if (flag == 0u) return;                              // nothing passes the guard: no amplitude, no seeds
  if (C::MODE == MODE_FAR) prep_far<C>(P, g, tv, it, dtInv, ps.V);
  else prep_near<C>(P, g, tv, it, r0, r1, r2, rL, ps.V, ps.V + 3);
  if (C::PAIR || C::KIND == KIND_RECUR) {
    // The seed arithmetic reproduces the reference's rounded phase fl(w_j*tau) only to ~4 ulp(phase);
    // beyond |phase| ~ 2^18 that exceeds the 1e-9 parity budget, so such steps are evaluated
    // node by node (flag 3).  fp32 main phase: the seeds are fp64, no such limit.
    const double wl = (double)((const typename C::TI*)P.omega)[P.descending ? g.cLo : g.cHi - 1];
    if (sizeof(TM) == 8 && fabs(wl * tau) > 262144.0) flag = 3u;
  }
  ps.flag = flag; ps.lo = lo; ps.hi = hi;
  st.nPass += hi - lo;
}

template <class C>
SRB_HD void prep_store(const Params& P, const Geom& g, const PrepStep& ps, int cnt, int lane, WarpSmem<C>& sm) {
  using TM = typename C::TM;
  if (lane >= cnt) return;
  const uint32_t flag = ps.flag;
  if (flag == 0u) { sm.rng[lane] = 0u; return; }
  const double tau = ps.tau;
  double last[3] = {tau, 0.0, 0.0};
  if constexpr (C::PAIR) {
    if (flag != 3u) make_seeds_pair<C>(P, g, tau, ps.V, sm, lane);
  }
  if (C::KIND == KIND_RECUR) {
    if (flag != 3u) make_seeds<C>(P, g, tau, sm, lane, last);
  }
  if constexpr (C::KIND == KIND_DREC) {
    double w4[4];
    make_seeds_drec<C>(P, g, tau, sm, lane, w4);
    sm.rec[lane][C::NV + 1] = w4[0]; sm.rec[lane][C::NV + 2] = w4[1]; sm.rec[lane][C::NV + 3] = w4[2]; sm.rec[lane][C::NV + 4] = w4[3];
  }
  sm.rng[lane] = ps.lo | (ps.hi << 10) | (flag << 30);
#pragma unroll
  for (int k = 0; k < C::NV; k++) sm.rec[lane][k] = (TM)ps.V[k];
  if (!C::PAIR || flag == 3u)
    sm.rec[lane][C::NV] = (TM)last[0];   // recurrence: 2cos(d) (flag 3: tau) ; direct: tau ; pair: flag 3 only
  if (C::KIND == KIND_DIRECT && sizeof(TM) == 4)
    sm.rec[lane][C::NV + 1] = (TM)ssub(tau, (double)(TM)tau);   // fp32 direct: tau = hi + lo (48 bits)
  if (C::KIND == KIND_RECUR) { sm.rec[lane][C::NV + 1] = (TM)last[1]; sm.rec[lane][C::NV + 2] = (TM)last[2]; }
}

// Guard-dominated sub-batches (wiggler / betatron regime: a few low-omega nodes pass at every step) on the recurrence
// kernels: the lane = tile main phase walks the 32 steps one after the other with a handful of lanes busy.  Here the
// lanes stay the STEPS: for every chunk node j below the largest pass bound of the sub-batch each lane evaluates its
// step's phasor at the reference's rounded phase (sincos_big of fl(w_j*tau): exact at any magnitude, so SI-unit
// phases need no special case) times its amplitude, the warp sums the 2 NV products, and the node's two owner
// lanes (cos part, sin part of tile j % 16) take them.  ~100 instructions per node instead of ~80 per step.
// Which steps of a sub-batch go through main_sparse: a cost model in instructions per warp, minimised over the split
// threshold H (steps with at most H passing nodes -> lane = step, the others -> lane = tile):
//   lane = step : SRB_SPARSE_ROW per node row (up to the largest pass bound among its steps) + a fixed part;
//   lane = tile : per step 50 + 40 per node row of 16 for steps with huge phases (per-node sincos), 30 + 6 for ordinary
//                 ones (recurrence), + the seeds / records of the prep phase once.
// Measured on the C3 / C4 recipes (profiles/r02_guard_dominated.md).
#ifndef SRB_SPARSE_ROW
#define SRB_SPARSE_ROW 85
#endif
SRB_HD uint32_t dense_step_cost(uint32_t flag, uint32_t hi) {
  if (!flag) return 0u;
  const uint32_t rows = (hi + 15u) / 16u;
  return flag == 3u ? 50u + 40u * rows : 30u + 6u * rows;
}
constexpr uint32_t SPARSE_FIXED = 64u, DENSE_FIXED = 150u;
constexpr int SPARSE_NH = 8;
SRB_HD uint32_t sparse_threshold(int i) { return i < 7 ? (2u << i) : 0x3ffu; }     // H = 2, 4, ..., 128, everything
#if defined(__CUDA_ARCH__)
template <class C>
SRB_HD void main_sparse(const Params& P, const Geom& g, const PrepStep& ps, uint32_t hi, uint32_t maxHi, int lane,
                        WarpSmem<C>& sm, ThreadState<C>& st) {     // hi: this lane's pass bound, 0 if its step is not taken here
  using TI = typename C::TI; using TM = typename C::TM;
  constexpr int NV = C::NV, TW = C::TW;
  // G nodes at a time: each lane (= step) stores its G * 2 NV products (V_c cos, V_c sin) as columns of a [value][lane]
  // array in the (idle) staging area; lane L then sums half L & 1 of row L >> 1, one shuffle joins the halves, and the
  // owner lanes of the G nodes fetch their NV sums: ~12 instructions per node for the reduction against 60 for a
  // shuffle butterfly of every value.
  constexpr int G = NV == 2 ? 4 : (NV == 3 ? 2 : 1), NVAL = G * 2 * NV, ROW = 33;
  static_assert(NVAL <= 16 && 16 % G == 0, "one row per lane pair");
  static_assert(sizeof(WarpSmem<C>) >= (size_t)NVAL * ROW * sizeof(double), "staging area too small for the sparse reduction");
  double* buf = reinterpret_cast<double*>(&sm);
  const uint32_t mine = (uint32_t)(lane & 15);
  const int part = lane >> 4, row = lane >> 1, half = lane & 1;
  const uint32_t jLast = g.cHi - g.cLo - 1u;
  for (uint32_t k = 0; 16u * k < maxHi; k++) {
    double tmp[NV];
#pragma unroll
    for (int c = 0; c < NV; c++) tmp[c] = 0.0;
    for (uint32_t m0 = 0; m0 < 16u && 16u * k + m0 < maxHi; m0 += (uint32_t)G) {
#pragma unroll
      for (int q = 0; q < G; q++) {
        const uint32_t j = 16u * k + m0 + (uint32_t)q;
        double sn, cs;
        sincos_big(smul((double)((const TI*)P.omega)[g.cLo + (j < jLast ? j : jLast)], ps.tau), &sn, &cs);
        if (!(j < hi)) { sn = 0.0; cs = 0.0; }
#pragma unroll
        for (int c = 0; c < NV; c++) {
          buf[(q * 2 * NV + c) * ROW + lane] = ps.V[c] * cs;
          buf[(q * 2 * NV + NV + c) * ROW + lane] = ps.V[c] * sn;
        }
      }
      __syncwarp();
      double tot = 0.0;
      if (row < NVAL) {
#pragma unroll
        for (int i = 0; i < 16; i++) tot += buf[row * ROW + half * 16 + i];
      }
      tot += __shfl_xor_sync(0xffffffffu, tot, 1);
      const int q = (int)mine - (int)m0;        // index of the node this lane owns in the group, if it is in it
      const bool owns = q >= 0 && q < G;
#pragma unroll
      for (int c = 0; c < NV; c++) {
        const double f = __shfl_sync(0xffffffffu, tot, 2 * ((owns ? q : 0) * 2 * NV + part * NV + c));
        if (owns) tmp[c] = f;
      }
      __syncwarp();
    }
#pragma unroll
    for (int kk = 0; kk < TW; kk++) {
      if ((uint32_t)kk == k) {        // warp-uniform
#pragma unroll
        for (int c = 0; c < NV; c++) st.acc[kk * NV + c] += (TM)tmp[c];
      }
    }
  }
}
#else
template <class C>
inline void main_sparse(const Params& P, const Geom& g, const PrepStep* ps, uint32_t maxHi, ThreadState<C>* st) {
  using TI = typename C::TI; using TM = typename C::TM;
  constexpr int NV = C::NV;
  for (uint32_t j = 0; j < maxHi; j++) {
    double a[NV], b[NV];
    for (int c = 0; c < NV; c++) a[c] = b[c] = 0.0;
    for (int lane = 0; lane < 32; lane++) {
      if (!ps[lane].flag || j >= ps[lane].hi) continue;
      double sn, cs;
      sincos_big(smul((double)((const TI*)P.omega)[g.cLo + j], ps[lane].tau), &sn, &cs);
      for (int c = 0; c < NV; c++) { a[c] += ps[lane].V[c] * cs; b[c] += ps[lane].V[c] * sn; }
    }
    for (int c = 0; c < NV; c++) {
      st[j & 15u].acc[(j >> 4) * NV + c] += (TM)a[c];
      st[(j & 15u) + 16u].acc[(j >> 4) * NV + c] += (TM)b[c];
    }
  }
}
#endif

// -------------------------------------------------------------------------------- main phase
// index of the lowest set bit (warp-uniform step masks are walked bit by bit instead of testing all 32 steps)
SRB_HD int low_bit(uint32_t m) {
#if defined(__CUDA_ARCH__)
  return __ffs((int)m) - 1;
#else
  return __builtin_ctz(m);
#endif
}
// smallest tile-local index k with m + T*k >= j (chunk-relative node bound j >= 0): ceil((j-m)/T)
SRB_HD int tile_lo(int j, int m, int T) { const int d = j - m; return d <= 0 ? 0 : (d + T - 1) / T; }

// the NV + 3 entries of a recurrence-kernel record (V, 2cos(d) | tau, cos(d), sin(d)) with 128-bit shared-memory loads
template <class C>
SRB_HD void rec_row(const WarpSmem<C>& sm, int s, typename C::TM* R) {
  using TM = typename C::TM;
  constexpr int N = C::NV + 3;
#if defined(__CUDA_ARCH__)
  if constexpr (sizeof(TM) == 4 && C::NREC % 4 == 0) {
    const float4* q = reinterpret_cast<const float4*>(&sm.rec[s][0]);
#pragma unroll
    for (int i = 0; i < (N + 3) / 4; i++) {
      const float4 t = q[i];
      if (4 * i + 0 < N) R[4 * i + 0] = t.x;
      if (4 * i + 1 < N) R[4 * i + 1] = t.y;
      if (4 * i + 2 < N) R[4 * i + 2] = t.z;
      if (4 * i + 3 < N) R[4 * i + 3] = t.w;
    }
  } else if constexpr (sizeof(TM) == 8 && C::NREC % 2 == 0) {
    const double2* q = reinterpret_cast<const double2*>(&sm.rec[s][0]);
#pragma unroll
    for (int i = 0; i < (N + 1) / 2; i++) {
      const double2 t = q[i];
      if (2 * i + 0 < N) R[2 * i + 0] = t.x;
      if (2 * i + 1 < N) R[2 * i + 1] = t.y;
    }
  } else
#endif
  {
#pragma unroll
    for (int k = 0; k < N; k++) R[k] = sm.rec[s][k];
  }
}

// one full step of one tile: v0, v1 = first two nodes of the tile, then the three-term recurrence
template <class C>
SRB_HD void tile_step_full(const typename C::TM* V, typename C::TM coef, typename C::TM vm, typename C::TM v,
                           ThreadState<C>& st) {
  using TM = typename C::TM;
  constexpr int TW = C::TW;
  constexpr int NV = C::NV;
#pragma unroll
  for (int c = 0; c < NV; c++) st.acc[c] = fma(V[c], vm, st.acc[c]);
#pragma unroll
  for (int c = 0; c < NV; c++) st.acc[NV + c] = fma(V[c], v, st.acc[NV + c]);
#pragma unroll
  for (int k = 2; k < TW; k++) {
    const TM vn = fma(coef, v, -vm); vm = v; v = vn;
#pragma unroll
    for (int c = 0; c < NV; c++) st.acc[k * NV + c] = fma(V[c], v, st.acc[k * NV + c]);
  }
}

template <class C>
SRB_HD void main_recur(const Params& P, const Geom& g, const WarpSmem<C>& sm, int cnt, uint32_t fullMask,
                       uint32_t anyMask, int lane, ThreadState<C>& st) {
  using TM = typename C::TM; using TI = typename C::TI;
  constexpr int TW = C::TW;
  constexpr int NV = C::NV;
  const int m = lane & 15;
  const uint32_t sgn = (lane >> 4) ? 0u : 0x80000000u;   // cos lanes subtract s0*sd
  const uint32_t allMask = cnt >= 32 ? 0xffffffffu : ((1u << cnt) - 1u);
  if (fullMask == allMask) {
    // hot path: every step of the sub-batch passes the guard at every node of the chunk.
    // Operands of step s+1 are fetched from shared memory while step s is being accumulated.
    TM R[NV + 3], nR[NV + 3], x, xo, nx, nxo;
    rec_row<C>(sm, 0, R);
    x = sm.seeds[lane][0]; xo = sm.seeds[lane ^ 16][0];
#pragma unroll 2
    for (int s = 0; s < cnt; s++) {
      const int sn = s + 1 < cnt ? s + 1 : s;
      rec_row<C>(sm, sn, nR);
      nx = sm.seeds[lane][sn]; nxo = sm.seeds[lane ^ 16][sn];
      // second node of the tile: cos lane c1 = c0*cd - s0*sd ; sin lane s1 = s0*cd + c0*sd
      const TM v1 = fma(xo, flipsign(R[NV + 2], sgn), x * R[NV + 1]);
      tile_step_full<C>(R, R[NV], x, v1, st);
#pragma unroll
      for (int k = 0; k < NV + 3; k++) R[k] = nR[k];
      x = nx; xo = nxo;
    }
    return;
  }
  for (uint32_t todo = anyMask; todo; todo &= todo - 1u) {   // steps where anything passes the guard (warp-uniform)
    const int s = low_bit(todo);
    const uint32_t r = sm.rng[s];
    const uint32_t flag = r >> 30;
    TM V[NV + 3];                         // V, 2cos(d) (flag 3: tau), cos(d), sin(d)
    rec_row<C>(sm, s, V);
    const TM coef = V[NV];
    // tile-local k passes iff m + 16k < hi (ascending uniform grid: the pass range starts at node 0) ;
    // kmax = largest count over the tiles (tile 0), uniform
    constexpr int lo = 0;
    const int hi = tile_lo((int)((r >> 10) & 0x3ffu), m, 16);
    const int kmax = tile_lo((int)((r >> 10) & 0x3ffu), 0, 16);
    if (flag == 3) {   // direct evaluation of this step (phase too large for the recurrence)
      const uint32_t j0 = g.cLo + (uint32_t)m;
      // (a rolled loop over k with a warp-uniform accumulator switch: this rare path is kept small -- the kernel's
      //  top stall was instruction fetch, profiles/r02_ncu_c3_recurrence_kernel.txt)
#pragma unroll 1
      for (int k = 0; k < kmax && k < TW; k++) {
        TM cur = (TM)0;
        if (k >= lo && k < hi) {
          TM sn, cs;
          sincos_t(tmul((TM)((const TI*)P.omega)[j0 + 16 * k], coef), &sn, &cs);   // coef slot holds tau
          cur = (lane >> 4) ? sn : cs;
        }
#pragma unroll
        for (int kk = 0; kk < TW; kk++) {
          if (kk == k) {
#pragma unroll
            for (int c = 0; c < NV; c++) st.acc[kk * NV + c] = fma(V[c], cur, st.acc[kk * NV + c]);
          }
        }
      }
      continue;
    }
    TM vm = sm.seeds[lane][s];
    TM v = fma(sm.seeds[lane ^ 16][s], flipsign(V[NV + 2], sgn), vm * V[NV + 1]);
    if (flag == 1) {
      tile_step_full<C>(V, coef, vm, v, st);
    } else {
#pragma unroll
      for (int k = 0; k < TW; k++) {
        if (k >= kmax) break;                      // warp-uniform: no tile has a passing node beyond
        TM cur;
        if (k == 0) cur = vm; else if (k == 1) cur = v;
        else { const TM vn = fma(coef, v, -vm); vm = v; v = vn; cur = vn; }
        if (k >= lo && k < hi) {
#pragma unroll
          for (int c = 0; c < NV; c++) st.acc[k * NV + c] = fma(V[c], cur, st.acc[k * NV + c]);
        }
      }
    }
  }
}

// Direct main phase: lane = tile of TW nodes, per-node sincos of the reference's rounded phase.
template <class C>
SRB_HD void direct_update(const typename C::TM* V, double w_, double tau, int k, ThreadState<C>& st) {
  using TM = typename C::TM;
  TM sn, cs;
  const TM w = (TM)w_;
  if constexpr (sizeof(TM) == 4) sincos_mixed<C::NATIVE>(w_, tau, (float*)&sn, (float*)&cs);
  else sincos_t(smul(w_, tau), &sn, &cs);  // single rounding == the reference's omega*(time - n.r)
  if (C::MODE == MODE_FAR) {
#pragma unroll
    for (int c = 0; c < C::NC; c++) {
      st.acc[k * C::NPN + c] = fma(V[c], cs, st.acc[k * C::NPN + c]);
      st.acc[k * C::NPN + C::NC + c] = fma(V[c], sn, st.acc[k * C::NPN + C::NC + c]);
    }
  } else {
    const TM t1 = w * sn, t2 = w * cs;
#pragma unroll
    for (int c = 0; c < 3; c++) {   // Re += -c1*sin + c2*cos ; Im += c1*cos + c2*sin
      st.acc[k * 6 + c] = fma(V[3 + c], cs, fma(-V[c], t1, st.acc[k * 6 + c]));
      st.acc[k * 6 + 3 + c] = fma(V[3 + c], sn, fma(V[c], t2, st.acc[k * 6 + 3 + c]));
    }
  }
}

template <class C>
SRB_HD void main_direct(const Params& P, const Geom& g, const WarpSmem<C>& sm, int cnt, uint32_t fullMask,
                        uint32_t anyMask, int lane, ThreadState<C>& st) {
  using TM = typename C::TM;
  constexpr int TW = C::TW;
  constexpr int NV = C::NV;
  const uint32_t allMask = cnt >= 32 ? 0xffffffffu : ((1u << cnt) - 1u);
  if (fullMask == allMask) {       // every node of the chunk passes at every step: no predicates
    for (int s = 0; s < cnt; s++) {
      TM V[NV];
#pragma unroll
      for (int k = 0; k < NV; k++) V[k] = sm.rec[s][k];
      const double tau = sizeof(TM) == 4 ? (double)sm.rec[s][NV] + (double)sm.rec[s][NV + 1] : (double)sm.rec[s][NV];
#pragma unroll
      for (int k = 0; k < TW; k++) direct_update<C>(V, st.wl[k], tau, k, st);
    }
    return;
  }
  for (uint32_t todo = anyMask; todo; todo &= todo - 1u) {
    const int s = low_bit(todo);
    const uint32_t r = sm.rng[s];
    const int lo = tile_lo((int)(r & 0x3ffu), lane, 32), hi = tile_lo((int)((r >> 10) & 0x3ffu), lane, 32);
    if (hi <= 0 || lo >= TW) continue;
    TM V[NV];
#pragma unroll
    for (int k = 0; k < NV; k++) V[k] = sm.rec[s][k];
    const double tau = sizeof(TM) == 4 ? (double)sm.rec[s][NV] + (double)sm.rec[s][NV + 1] : (double)sm.rec[s][NV];
    const int kmin = tile_lo((int)(r & 0x3ffu), 31, 32), kmax = tile_lo((int)((r >> 10) & 0x3ffu), 0, 32);   // warp-uniform
#pragma unroll
    for (int k = 0; k < TW; k++) {
      if (k >= kmax) break;
      if (k < kmin) continue;
      if (k >= lo && k < hi) direct_update<C>(V, st.wl[k], tau, k, st);
    }
  }
}

// -------------------------------------------------------------------------------- flush
// Adds this track's contribution to snapshot iSnap (kernel_farfield.cl:100-106 and variants).
// Accumulators are NOT reset (cumulative snapshots, Q3).  `xchg(state, idx)` returns the
// value `f(partner lane)`; on the GPU it is a shuffle, in the emulator a direct read.
template <class C>
SRB_HD double* dest(const Params& P, uint32_t pc, int c) {
  const size_t nTotal = (size_t)P.nOmega * P.nA2 * P.nPhi;
  return pc == 0 ? P.out[c] : P.slabs + (size_t)(pc - 1) * P.slabStride + (size_t)c * P.nSnaps * nTotal;
}

// One node's contribution to snapshot iSnap from its complex amplitude (re, im: NCF components in the basis the
// kernel carries): the epilogues of kernel_farfield.cl:100-106 and its variants.  doRe / doIm: which parts this lane
// writes (the split layout of the recurrence kernels holds each node in two lanes).
template <int MODE, int NCF>
SRB_HD void emit_node(const Params& P, const Geom& g, double w, uint32_t pc, size_t idx, uint32_t j, double* re, double* im,
                      bool doRe, bool doIm) {
  const size_t nTotal = (size_t)P.nOmega * P.nA2 * P.nPhi;
  auto dst = [&](int c) -> double* {
    return pc == 0 ? P.out[c] : P.slabs + (size_t)(pc - 1) * P.slabStride + (size_t)c * P.nSnaps * nTotal;
  };
  const bool cplx = (P.comp == COMP_CART_CPLX || P.comp == COMP_SPH_CPLX);
  const double wpdt2 = smul(smul(w, P.dt), P.dt);
  if (MODE == MODE_FAR && NCF == 2) {
    if (P.comp == COMP_TOTAL) {
      // |F|^2 in the orthonormal transverse basis
      if (doRe) dst(0)[idx] += wpdt2 * ((re[0] * re[0] + re[1] * re[1]) + (im[0] * im[0] + im[1] * im[1]));
      return;
    }
    // Cartesian components F = F_theta e_theta + F_phi e_phi
    const double rt = re[0], rp = re[1], it_ = im[0], ip = im[1];
    re[0] = rt * g.tx + rp * g.px; re[1] = rt * g.ty + rp * g.py; re[2] = rt * g.tz + rp * g.pz;
    im[0] = it_ * g.tx + ip * g.px; im[1] = it_ * g.ty + ip * g.py; im[2] = it_ * g.tz + ip * g.pz;
  }
  if (!cplx) {
    if (doRe) {
      if (P.comp == COMP_TOTAL) {
        dst(0)[idx] += wpdt2 * (((re[0] * re[0] + re[1] * re[1]) + re[2] * re[2]) +
                               ((im[0] * im[0] + im[1] * im[1]) + im[2] * im[2]));
      } else {
#pragma unroll
        for (int c = 0; c < 3; c++) dst(c)[idx] += wpdt2 * (re[c] * re[c] + im[c] * im[c]);
      }
    }
  } else {
    const double wpdt = smul(ssqrt(w), P.dt);
    const bool useFF = (MODE == MODE_FAR && P.comp == COMP_CART_CPLX && P.formFactor != nullptr);
    const double ff = useFF ? ((const double*)P.formFactor)[j] : 1.0;
#pragma unroll
    for (int c = 0; c < 3; c++) {
      if (doRe) dst(2 * c)[idx] += wpdt * (re[c] * ff);
      if (doIm) dst(2 * c + 1)[idx] += wpdt * (im[c] * ff);
    }
  }
}

// Split layout (recurrence kernels): lanes l and l^16 hold the cos and sin parts of the same 16
// tiles, so the partner's accumulators of node k are fetched (GPU: shuffles; emulator: direct
// read) and both lanes reconstruct the full complex amplitude; lane part 0 then writes.
// one node of the flush: ma = this lane's NPN accumulators of tile-local node k, pa = the partner lane's (split layout
// of the recurrence kernels only)
template <class C>
SRB_HD void flush_node(const Params& P, const Geom& g, const TrackView& tv, uint32_t pc, uint32_t iSnap, int lane, int k,
                       const double* ma, const double* pa) {
  using TI = typename C::TI;
  constexpr int NPN = C::NPN;
  constexpr int NCF = (C::MODE == MODE_FAR) ? C::NC : 3;   // complex amplitude components held
  const size_t nTotal = (size_t)P.nOmega * P.nA2 * P.nPhi;
  const int tile = (C::KIND == KIND_RECUR) ? (lane & 15) : lane;
  const int part = (C::KIND == KIND_RECUR) ? (lane >> 4) : 0;
  const uint32_t j = g.cLo + (uint32_t)(tile + C::TILES * k);
  if (!(j < g.cHi)) return;
  const size_t idx = (size_t)j + (size_t)P.nOmega * (g.iA2 + (size_t)P.nA2 * g.iPhi) + nTotal * iSnap;
  double re[3], im[3];
  if constexpr (C::DIRECTLIKE || C::PAIR) {
#pragma unroll
    for (int c = 0; c < NCF; c++) { re[c] = ma[c]; im[c] = ma[NCF + c]; }
  } else {
    const double* cosS = part ? pa : ma;
    const double* sinS = part ? ma : pa;
    if (C::MODE == MODE_FAR) {
#pragma unroll
      for (int c = 0; c < NCF; c++) { re[c] = cosS[c]; im[c] = sinS[c]; }
    } else {
      // near: acc = {P = sum B*v, Q = sum Cv*v}; Re = Qcos - w*Psin, Im = Qsin + w*Pcos
      const double wj = (double)((const TI*)P.omega)[j];
#pragma unroll
      for (int c = 0; c < 3; c++) {
        re[c] = cosS[(NPN > 3 ? 3 : 0) + c] - wj * sinS[c];
        im[c] = sinS[(NPN > 3 ? 3 : 0) + c] + wj * cosS[c];
      }
    }
  }
  if (P.nTS > 1u) {   // time-axis split: this segment's partial amplitude, squared after the segments are summed
    if (part == 0) {
      double* a = P.amp + (((size_t)pc * P.nSnaps + iSnap) * (size_t)(2 * NCF)) * nTotal + (idx - nTotal * iSnap);
#pragma unroll
      for (int c = 0; c < NCF; c++) { a[(size_t)c * nTotal] = re[c]; a[(size_t)(NCF + c) * nTotal] = im[c]; }
    }
    return;
  }
  emit_node<C::MODE, NCF>(P, g, tv.w, pc, idx, j, re, im, part == 0, C::KIND != KIND_RECUR || part == 1);
}

// `stage` (GPU, non-pair kinds): the warp's idle staging area.  The accumulators go through it so that the loop over the
// lane's nodes can stay ROLLED: unrolled over 16-node tiles the flush is ~90 KB of code, executed once per (track,
// direction) -- on short tracks it kept evicting the hot loops from the instruction cache (C3 recipe: `no_instruction`
// was the top stall, profiles/r02_guard_dominated.md).
template <class C>
SRB_HD void flush_lane(const Params& P, const Geom& g, const TrackView& tv, uint32_t pc, uint32_t iSnap,
                       int lane, const ThreadState<C>* st, typename C::TM* stage = nullptr) {
  using TM = typename C::TM;
  constexpr int TW = C::TW;
  constexpr int NPN = C::NPN;
#if defined(__CUDA_ARCH__)
  const ThreadState<C>& me = st[0];
  constexpr int CAP = (int)(sizeof(WarpSmem<C>) / (32 * sizeof(TM)));     // accumulators per lane that fit
  constexpr int KC = CAP / NPN >= TW ? TW : CAP / NPN;                     // nodes per pass (0: the direct kinds' small area)
  if constexpr (!C::PAIR && KC >= 1) {
    if (stage) {
#pragma unroll
      for (int k0 = 0; k0 < TW; k0 += KC) {
        __syncwarp();
#pragma unroll
        for (int q = 0; q < KC * NPN; q++)
          if (k0 * NPN + q < C::NACC) stage[q * 32 + lane] = me.acc[k0 * NPN + q];
        __syncwarp();
#pragma unroll 1
        for (int kk = 0; kk < KC; kk++) {
          if (k0 + kk >= TW) break;
          double ma[NPN], pa[NPN];
#pragma unroll
          for (int c = 0; c < NPN; c++) {
            ma[c] = (double)stage[(kk * NPN + c) * 32 + lane];
            pa[c] = C::KIND == KIND_RECUR ? (double)stage[(kk * NPN + c) * 32 + (lane ^ 16)] : 0.0;
          }
          flush_node<C>(P, g, tv, pc, iSnap, lane, k0 + kk, ma, pa);
        }
      }
      __syncwarp();
      return;
    }
  }
#else
  const ThreadState<C>& me = st[lane];
  (void)stage;
#endif
#pragma unroll
  for (int k = 0; k < TW; k++) {
    double ma[NPN > 6 ? NPN : 6], pa[NPN > 6 ? NPN : 6];
    if constexpr (C::PAIR) {
      pair_node_amp<C>(me, k, ma, ma + ((C::MODE == MODE_FAR) ? C::NC : 3));
    } else {
#pragma unroll
      for (int c = 0; c < NPN; c++) {
        ma[c] = (double)me.acc[k * NPN + c];
        if (C::KIND == KIND_RECUR) {
#if defined(__CUDA_ARCH__)
          pa[c] = __shfl_xor_sync(0xffffffffu, ma[c], 16);
#else
          pa[c] = (double)st[lane ^ 16].acc[k * NPN + c];
#endif
        }
      }
    }
    flush_node<C>(P, g, tv, pc, iSnap, lane, k, ma, pa);
  }
}

// geometry of virtual direction vd: node range of the omega chunk, unit vector n (far) / screen point (near), e_theta, e_phi
template <class C>
SRB_HD void make_geom(const Params& P, uint32_t vd, Geom& g) {
  using TI = typename C::TI;
  const uint32_t ch = vd % P.nChunks; const uint32_t d = vd / P.nChunks;
  g.iA2 = d % P.nA2; g.iPhi = d / P.nA2;
  g.cLo = ch * P.chunkNodes; g.cHi = g.cLo + P.chunkNodes < P.nOmega ? g.cLo + P.chunkNodes : P.nOmega;
  const double sP = ldv<TI>(P.sinPhi, g.iPhi), cP = ldv<TI>(P.cosPhi, g.iPhi);
  if (C::MODE == MODE_FAR) {
    const double sT = ldv<TI>(P.axA, g.iA2), cT = ldv<TI>(P.axB, g.iA2);
    if (C::KIND != KIND_LITERAL) { g.nx = smul(sT, cP); g.ny = smul(sT, sP); g.nz = cT; g.tx = smul(cT, cP); g.ty = smul(cT, sP); }
    else {  // the reference forms n in fp32 (kernel_farfield.cl:40-42)
      g.nx = (double)((float)sT * (float)cP); g.ny = (double)((float)sT * (float)sP); g.nz = cT;
      g.tx = (double)((float)cT * (float)cP); g.ty = (double)((float)cT * (float)sP);
    }
    g.tz = -sT; g.px = -sP; g.py = cP; g.pz = 0.0;
  } else {
    const double r = ldv<TI>(P.axA, g.iA2);
    if (C::KIND != KIND_LITERAL) { g.nx = smul(r, cP); g.ny = smul(r, sP); }
    else { g.nx = (double)((float)r * (float)cP); g.ny = (double)((float)r * (float)sP); }
    g.nz = P.L;
    g.tx = g.ty = g.tz = g.px = g.py = g.pz = 0.0;
  }
}

// particle chunk -> track range [t0, t1), balanced by cumulative steps (offsets is a prefix sum)
SRB_HD void chunk_tracks(const Params& P, uint32_t pc, uint32_t& t0, uint32_t& t1) {
  if (P.nTS > 1u) { t0 = pc / P.nTS; t1 = t0 + 1u; return; }   // time-axis split: one (track, segment) per chunk
  const uint64_t total = P.offsets[P.nTracks];
  auto bound = [&](uint32_t c) -> uint32_t {
    if (c == 0) return 0u;
    if (c >= P.nPC) return P.nTracks;
    const uint64_t target = (total / P.nPC) * c + ((total % P.nPC) * c) / P.nPC;
    uint32_t a = 0, b = P.nTracks;      // first track whose start offset >= target
    while (a < b) { const uint32_t mid = (a + b) >> 1; if (P.offsets[mid] < target) a = mid + 1; else b = mid; }
    return a;
  };
  t0 = bound(pc); t1 = bound(pc + 1);
}

// time-axis split: loop indices [segLo, segHi) of segment pc % nTS of a track with nComp computed steps -- whole 32-step
// sub-batches, the same number for every segment (the last ones may be short or empty)
SRB_HD void seg_range(const Params& P, uint32_t pc, uint32_t nComp, uint32_t& segLo, uint32_t& segHi) {
  if (P.nTS <= 1u) { segLo = 0u; segHi = 0xffffffffu; return; }
  const uint32_t seg = pc % P.nTS;
  const uint32_t nSub = (nComp + (uint32_t)SUB - 1u) / (uint32_t)SUB;
  const uint32_t per = (nSub + P.nTS - 1u) / P.nTS;
  const uint64_t lo = (uint64_t)seg * per * SUB, hi = lo + (uint64_t)per * SUB;
  segLo = lo < nComp ? (uint32_t)lo : nComp;
  segHi = hi < nComp ? (uint32_t)hi : nComp;
}

// time-axis split, second pass: element i = (iSnap, node) of the spectra.  Per track (in order: deterministic) the
// segments' partial amplitudes are summed and the common epilogue adds the track's contribution.  Snapshots that did
// not fire for a track hold zeros (the buffer is cleared before the integration kernel) and add exactly 0.
template <int MODE, int NCF>
SRB_HD void combine_node(const Params& P, size_t i) {
  const size_t nTotal = (size_t)P.nOmega * P.nA2 * P.nPhi;
  const size_t iSnap = i / nTotal, node = i - iSnap * nTotal;
  const uint32_t j = (uint32_t)(node % P.nOmega);
  const size_t d = node / P.nOmega;
  Geom g;
  g.iA2 = (uint32_t)(d % P.nA2); g.iPhi = (uint32_t)(d / P.nA2);
  g.tx = g.ty = g.tz = g.px = g.py = g.pz = 0.0;
  if (MODE == MODE_FAR && NCF == 2) {     // e_theta, e_phi as make_geom forms them
    const double sP = ((const double*)P.sinPhi)[g.iPhi], cP = ((const double*)P.cosPhi)[g.iPhi];
    const double sT = ((const double*)P.axA)[g.iA2], cT = ((const double*)P.axB)[g.iA2];
    g.tx = smul(cT, cP); g.ty = smul(cT, sP); g.tz = -sT; g.px = -sP; g.py = cP;
  }
  for (uint32_t t = 0; t < P.nTracks; t++) {
    double re[3] = {0, 0, 0}, im[3] = {0, 0, 0};
    for (uint32_t s = 0; s < P.nTS; s++) {
      const double* a = P.amp + ((((size_t)t * P.nTS + s) * P.nSnaps + iSnap) * (size_t)(2 * NCF)) * nTotal + node;
#pragma unroll
      for (int c = 0; c < NCF; c++) { re[c] += a[(size_t)c * nTotal]; im[c] += a[(size_t)(NCF + c) * nTotal]; }
    }
    emit_node<MODE, NCF>(P, g, ((const double*)P.w)[t], 0u, i, j, re, im, true, true);
  }
}

// track t of the batch
template <class C>
SRB_HD void load_track(const Params& P, uint32_t t, TrackView& tv) {
  using TI = typename C::TI;
  const uint64_t o = P.offsets[t];
  tv.n = (uint32_t)(P.offsets[t + 1] - o);
  tv.x = (const TI*)P.x + o; tv.y = (const TI*)P.y + o; tv.z = (const TI*)P.z + o;
  tv.ux = (const TI*)P.ux + o; tv.uy = (const TI*)P.uy + o; tv.uz = (const TI*)P.uz + o;
  tv.pre = P.pre ? P.pre + o : nullptr;
  tv.itStart = P.itStart[t]; tv.itEnd = P.itEnd[t];
  tv.snaps = P.itSnaps + (size_t)P.snapStride * t;
  tv.w = ldv<TI>(P.w, t);
}

// -------------------------------------------------------------------------------- warp task
// One warp integrates all tracks of particle chunk `pc` for virtual direction `vd`.
template <class C>
SRB_HD void warp_task(const Params& P, uint32_t vd, uint32_t pc, WarpSmem<C>& sm, ThreadState<C>* st) {
  using TI = typename C::TI; using TM = typename C::TM;
  Geom g;
  make_geom<C>(P, vd, g);
  uint32_t t0, t1;
  chunk_tracks(P, pc, t0, t1);
  const double dtInv = sdiv(1.0, P.dt);
  if constexpr (C::MMA) {
    // masked steps of the tensor-core main phase multiply stale staging data by 0: keep it finite from the start
    SRB_LANES_BEGIN
      uint32_t* raw = reinterpret_cast<uint32_t*>(&sm);
      for (uint32_t k = (uint32_t)lane; k < sizeof(WarpSmem<C>) / 4u; k += 32u) raw[k] = 0u;
    SRB_LANES_END
  }
  SRB_LANES_BEGIN
    SRB_ST.nPass = 0; SRB_ST.nAll = 0;
    if constexpr (C::KIND == KIND_DREC) drec_init_lane<C>(P, g, lane, SRB_ST);
    if (C::DIRECTLIKE || C::KIND == KIND_LITERAL) {
#pragma unroll
      for (int k = 0; k < C::TW; k++) {
        const uint32_t j = g.cLo + (uint32_t)(lane + 32 * k);
        SRB_ST.wl[k] = j < g.cHi ? (typename C::TW_T)((const TI*)P.omega)[j] : (typename C::TW_T)0;
        if (C::KIND == KIND_LITERAL)
          SRB_ST.ff[k] = (j < g.cHi && P.formFactor) ? (TM)((const TI*)P.formFactor)[j] : (TM)1;
      }
    }
  SRB_LANES_END

  for (uint32_t t = t0; t < t1; t++) {
    TrackView tv;
    load_track<C>(P, t, tv);
    SRB_LANES_BEGIN
#pragma unroll
      for (int k = 0; k < C::NACC; k++) SRB_ST.acc[k] = (TM)0;
      if (C::KIND == KIND_LITERAL) {
#pragma unroll
        for (int k = 0; k < C::TW; k++) SRB_ST.pprev[k] = (TM)0;      // phasePrev starts at 0 (Q1)
      }
    SRB_LANES_END
    // loop bounds of kernel_farfield.cl:59-63 (uint wrap-around for itEnd==0 / nSteps==0 not reproduced)
    const uint32_t loopEnd = tv.itEnd > 0 ? tv.itEnd - 1 : 0;
    const uint32_t nComp = tv.n > 0 ? (tv.n - 1 < loopEnd ? tv.n - 1 : loopEnd) : 0;
    uint32_t iSnap = 0;
    while (iSnap < P.nSnaps && !(tv.itStart < tv.snaps[iSnap])) iSnap++;   // :54-57
    uint32_t segLo, segHi;   // time-axis split: this chunk's share of the steps (everything otherwise)
    seg_range(P, pc, nComp, segLo, segHi);
    uint32_t cur = 0;    // next loop index `it` to process
    while (iSnap < P.nSnaps) {
      // the flush test `it_glob + 2 == itSnaps[iSnap]` (:100) fires at it = itf, if reachable
      const long long itf = (long long)tv.snaps[iSnap] - 2 - (long long)tv.itStart;
      if (itf < (long long)cur || itf >= (long long)loopEnd) break;   // never fires again (Q3/Q4)
      uint32_t stop = (uint32_t)(itf + 1) < nComp ? (uint32_t)(itf + 1) : nComp;
      if (stop > segHi) stop = segHi;
      for (uint32_t base = cur > segLo ? cur : segLo; base < stop; base += SUB) {
        const int cnt = (int)(stop - base < (uint32_t)SUB ? stop - base : (uint32_t)SUB);
        if constexpr (C::KIND == KIND_LITERAL) {
          SRB_LANES_BEGIN
            lit_prep_phase<C>(P, g, tv, base, cnt, lane, sm);
          SRB_LANES_END
          SRB_LANES_BEGIN
            lit_main_phase<C>(P, g, sm, cnt, lane, SRB_ST);
          SRB_LANES_END
        } else {
#if defined(__CUDA_ARCH__)
        // the prep phase of the NEXT sub-batch starts with dependent global loads: warm L1 during this main phase
        // (measured +2 % on the C5 shard)
        if (base + SUB < stop) {
          const uint32_t itn = base + SUB + (threadIdx.x & 31u);
          if (itn < stop) {
            asm volatile("prefetch.global.L1 [%0];" ::"l"((const TI*)tv.x + itn));
            asm volatile("prefetch.global.L1 [%0];" ::"l"((const TI*)tv.y + itn));
            asm volatile("prefetch.global.L1 [%0];" ::"l"((const TI*)tv.z + itn));
            if (tv.pre) {
#pragma unroll
              for (int c = 0; c < (C::MODE == MODE_FAR ? 6 : 3); c++)
                asm volatile("prefetch.global.L1 [%0];" ::"l"(tv.pre + c * P.preStride + itn));
            }
          }
        }
#endif
        uint32_t fullMask = 0u, anyMask = 0u;   // bit s: step s of the sub-batch is all-pass / has any pass
        // recurrence kernels: the steps where only a few low nodes pass go through the lane = step evaluation
        // (main_sparse), the others through seeds + shared-memory records + the lane = tile main phase; the split is
        // chosen per sub-batch from a cost model (sparse_threshold, dense_step_cost)
        uint32_t sparseMask = 0u, maxHi = 0u;
#if defined(__CUDA_ARCH__)
        PrepStep psv[1];
#else
        PrepStep psv[32];
#endif
        SRB_LANES_BEGIN
#if defined(__CUDA_ARCH__)
          PrepStep& ps = psv[0];
#else
          PrepStep& ps = psv[lane];
#endif
          prep_guard<C>(P, g, tv, base, cnt, dtInv, lane, SRB_ST, ps);
#if defined(__CUDA_ARCH__)
          fullMask = __ballot_sync(0xffffffffu, ps.flag == 1u);
          anyMask = __ballot_sync(0xffffffffu, ps.flag != 0u);
          if constexpr (C::KIND == KIND_RECUR) {
            // steps with huge phases (per-node sincos in either form) that pass at some nodes only: which form for which?
            // (ordinary phases: the recurrence form of the lane = tile loop is not beaten, measured on the C4 recipe)
            if (anyMask != fullMask && __ballot_sync(0xffffffffu, ps.flag == 3u)) {
              const uint32_t mine = dense_step_cost(ps.flag, ps.hi);
              uint32_t best = __reduce_add_sync(0xffffffffu, mine) + DENSE_FIXED;      // everything lane = tile
#pragma unroll
              for (int i = 0; i < SPARSE_NH; i++) {
                const bool small = ps.flag == 3u && ps.hi <= sparse_threshold(i);
                const uint32_t mh = __reduce_max_sync(0xffffffffu, small ? ps.hi : 0u);
                if (mh == 0u) continue;                                                 // warp-uniform
                const uint32_t rest = __reduce_add_sync(0xffffffffu, small ? 0u : mine);
                const uint32_t cost = mh * (uint32_t)SRB_SPARSE_ROW + SPARSE_FIXED + (rest ? rest + DENSE_FIXED : 0u);
                if (cost < best) { best = cost; sparseMask = __ballot_sync(0xffffffffu, small); maxHi = mh; }
              }
            }
          }
#else
          if (ps.flag == 1u) fullMask |= 1u << lane;
          if (ps.flag != 0u) anyMask |= 1u << lane;
#endif
        SRB_LANES_END
#if defined(SRB_SKIP_MAIN)      // tuning aid: the cost of the guard part of the prep phase alone
        continue;
#endif
#if !defined(__CUDA_ARCH__)
        if (C::KIND == KIND_RECUR && anyMask != fullMask) {      // the same decision, lanes as a loop
          uint32_t all = 0u;
          for (int l = 0; l < 32; l++) all += dense_step_cost(psv[l].flag, psv[l].hi);
          uint32_t best = all + DENSE_FIXED;
          for (int i = 0; i < SPARSE_NH; i++) {
            uint32_t mh = 0u, rest = 0u, mask = 0u;
            for (int l = 0; l < 32; l++) {
              const bool small = psv[l].flag == 3u && psv[l].hi <= sparse_threshold(i);
              if (small) { mask |= 1u << l; if (psv[l].hi > mh) mh = psv[l].hi; }
              else rest += dense_step_cost(psv[l].flag, psv[l].hi);
            }
            if (mh == 0u) continue;
            const uint32_t cost = mh * (uint32_t)SRB_SPARSE_ROW + SPARSE_FIXED + (rest ? rest + DENSE_FIXED : 0u);
            if (cost < best) { best = cost; sparseMask = mask; maxHi = mh; }
          }
        }
#endif
        if constexpr (C::KIND == KIND_RECUR) {
          if (sparseMask) {
#if defined(__CUDA_ARCH__)
            const bool taken = (sparseMask >> (threadIdx.x & 31u)) & 1u;
            main_sparse<C>(P, g, psv[0], taken ? psv[0].hi : 0u, maxHi, (int)(threadIdx.x & 31u), sm, st[0]);
            if (taken) psv[0].flag = 0u;
            __syncwarp();
#else
            PrepStep sel[32];
            for (int l = 0; l < 32; l++) { sel[l] = psv[l]; if (!((sparseMask >> l) & 1u)) sel[l].flag = 0u; else psv[l].flag = 0u; }
            main_sparse<C>(P, g, sel, maxHi, st);
#endif
            anyMask &= ~sparseMask;
            if (!anyMask) continue;
          }
        }
        SRB_LANES_BEGIN
#if defined(__CUDA_ARCH__)
          prep_store<C>(P, g, psv[0], cnt, lane, sm);
#else
          prep_store<C>(P, g, psv[lane], cnt, lane, sm);
#endif
        SRB_LANES_END
#if !defined(__CUDA_ARCH__)
        if constexpr (C::MMA) main_pair_mma<C>(P, g, sm, cnt, fullMask, anyMask, st);   // warp-collective: emulated over all lanes
        else
#endif
        {
        SRB_LANES_BEGIN
          if constexpr (C::KIND == KIND_RECUR) main_recur<C>(P, g, sm, cnt, fullMask, anyMask, lane, SRB_ST);
          else if constexpr (C::MMA) {
#if defined(__CUDA_ARCH__)
            main_pair_mma<C>(P, g, sm, cnt, fullMask, anyMask, lane, SRB_ST);
#endif
          }
          else if constexpr (C::PAIR) main_pair<C>(P, g, sm, cnt, fullMask, anyMask, lane, SRB_ST);
          else if constexpr (C::KIND == KIND_DREC) main_drec<C>(P, g, sm, cnt, fullMask, anyMask, lane, SRB_ST);
          else main_direct<C>(P, g, sm, cnt, fullMask, anyMask, lane, SRB_ST);
        SRB_LANES_END
        }
        }
      }
      if constexpr (C::MMA) {
        // tensor-core accumulator layout -> one tile per lane for the flush, and back (cumulative snapshots)
        SRB_LANES_BEGIN
          pair_mma_store_frag<C>(sm, lane, SRB_ST);
        SRB_LANES_END
        SRB_LANES_BEGIN
          pair_mma_load_tile<C>(sm, lane, SRB_ST);
        SRB_LANES_END
      }
      SRB_LANES_BEGIN
        if constexpr (C::KIND == KIND_LITERAL) lit_flush_lane<C>(P, g, tv, pc, iSnap, lane, SRB_ST);
        else flush_lane<C>(P, g, tv, pc, iSnap, lane, st, C::PAIR ? nullptr : reinterpret_cast<typename C::TM*>(&sm));
      SRB_LANES_END
      if constexpr (C::MMA) {
        SRB_LANES_BEGIN
          pair_mma_store_tile<C>(sm, lane, SRB_ST);
        SRB_LANES_END
        SRB_LANES_BEGIN
          pair_mma_load_frag<C>(sm, lane, SRB_ST);
        SRB_LANES_END
      }
      cur = (uint32_t)(itf + 1);
      iSnap++;
    }
  }
  if (P.counters) {
    SRB_LANES_BEGIN
#if defined(__CUDA_ARCH__)
      // one atomic pair per WARP (32 lanes x thousands of warps on the same two addresses serialise in L2: ~1 ns each,
      // visible as soon as blocks are short -- the time-axis split)
      unsigned long long np_ = SRB_ST.nPass, na_ = SRB_ST.nAll;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { np_ += __shfl_xor_sync(0xffffffffu, np_, o); na_ += __shfl_xor_sync(0xffffffffu, na_, o); }
      if (lane == 0 && na_) { atomicAdd(P.counters, np_); atomicAdd(P.counters + 1, na_); }
#else
      P.counters[0] += SRB_ST.nPass; P.counters[1] += SRB_ST.nAll;
#endif
    SRB_LANES_END
  }
}

}  // namespace srb
