/* synchrad_b200 — C ABI of the B200 spectral-integration path (libsynchrad_b200.so).
 *
 * The reference (hightower8083/synchrad) has no FFI: its device boundary is the positional
 * argument list that `SynchRad._process_track` marshals into the PyOpenCL kernels, ONE PARTICLE
 * PER LAUNCH (synchrad/calc.py:292-353; kernel prototypes kernel_farfield.cl:6-28,
 * kernel_nearfield.cl:5-27).  This header is that argument list turned into a batched C ABI:
 * the same tables, the same per-track scalars, all particles of a rank in one call.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`; the caller (torch,
 *     cudaMalloc, ...) owns every buffer, the library allocates nothing persistent;
 *   - `stream` is a cudaStream_t passed as void* (0 = default stream); calls are asynchronous
 *     with respect to the host unless stated;
 *   - spectra are ACCUMULATED INTO (`+=`), like the reference kernels' `spectrum[...] +=`
 *     (kernel_farfield.cl:102); zeroing is the caller's job (calc.py:482-484);
 *   - spectra are float64 in the reference's device layout (nSnaps, nPhi, nAxis2, nOmega)
 *     (calc.py:455), whatever the compute dtype (documented deviation from Q5: the
 *     cross-particle sum is carried in fp64, the host result is fp64 anyway, calc.py:576-577);
 *   - return value 0 = success, <0 = error; the message is in srb_last_error() (thread-local).
 *     No C++ exception crosses this boundary.
 */
#ifndef SYNCHRAD_B200_H
#define SYNCHRAD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SRB_ABI_VERSION 3

/* mode: which kernel file of the reference is replaced (calc.py:617-620) */
#define SRB_MODE_FAR 0  /* kernel_farfield.cl  */
#define SRB_MODE_NEAR 1 /* kernel_nearfield.cl */

/* comp: which __kernel of that file (dispatch of calc.py:324-353) */
#define SRB_COMP_TOTAL 0             /* total                    -> 1 spectrum  */
#define SRB_COMP_CARTESIAN 1         /* cartesian_comps          -> 3 spectra   */
#define SRB_COMP_CARTESIAN_COMPLEX 2 /* cartesian_comps_complex  -> 6 spectra   */
#define SRB_COMP_SPHERIC 3           /* spheric_comps (far only) -> 3 spectra   */
#define SRB_COMP_SPHERIC_COMPLEX 4   /* spheric_comps_complex    -> 6 spectra   */

/* dtype: the Mako variable `my_dtype` (calc.py:610-611).
 * F64: everything in double, bit-faithful to the strict reading of the reference kernels.
 * F32: MIXED precision — tables, tracks and all per-(direction, step) work stay float64; only
 *      the per-omega phasor and accumulation arithmetic is float32.  (The reference's literal
 *      float32 path loses ~0.4 rad of phase to the float32 rounding of tracks and direction
 *      cosines alone; it has no reproducible answer to be bit-compatible with.) */
#define SRB_DTYPE_F64 0
#define SRB_DTYPE_F32 1
/* F32_LITERAL: every operation of the reference kernels in fp32, in the reference's order, per-node
 *      Nyquist guard on fp32 phases (srb_literal.cuh).  Arrays are still float64 buffers; tracks are
 *      rounded to fp32 on load (the reference's astype(float32)); tables must hold fp32-representable
 *      values computed as `_init_data` computes them in float32.  Reproduces the reference's single-
 *      precision behaviour, including its ~0.4 rad phase noise. */
#define SRB_DTYPE_F32_LITERAL 2

/* phasor: how exp(i*omega*tau) is evaluated per node */
#define SRB_PHASOR_AUTO 0   /* uniform grid, far field: pair kernel, recurrence (many partially passing steps) or -- fp64,
                               mostly all-pass steps with phases beyond 2^18 -- corrected recurrence, chosen ON THE DEVICE
                               from a sampled Nyquist-guard / phase statistic (all candidates are enqueued, the others
                               return at once; no host synchronisation; choice reported in counters[2]); near field:
                               recurrence or corrected recurrence by omega*L; non-uniform grids: direct */
#define SRB_PHASOR_DIRECT 1 /* per-node sincos of the reference's rounded phase */
#define SRB_PHASOR_RECUR 2  /* three-term recurrence along omega (uniform grids only) */
#define SRB_PHASOR_PAIR 3   /* symmetric node pairs about the tile centre, broadcast pair phasors (uniform grids, far
                               field); fp64: the accumulation over steps runs as a GEMM on the FP64 tensor cores
                               (DMMA.8x8x4) when tile width x components is a multiple of 8 */
#define SRB_PHASOR_PAIR_FMA 4 /* the pair kernel with the accumulation kept on the scalar FP64 pipe (DFMA) */
#define SRB_PHASOR_DREC 5   /* direct layout, per-lane recurrence along omega + per-update first-order correction onto the
                               reference's rounded phase: phases up to 3e10 rad (near field at large L; refused beyond:
                               one ulp of the phase then exceeds 1e-5 rad), uniform grids, fp64 */

/* srb_launch_info.kind when the kernel was chosen on the device (read counters[2]) */
#define SRB_KIND_ON_DEVICE (-1)

/* Spectral grid + run constants: the `args_axes + args_res + args_aux` of calc.py:306-322.
 * Tables are float64 arrays with the content `_init_data` uploads (calc.py:486-512): omega is
 * already multiplied by 2*pi. */
typedef struct srb_grid {
  int32_t mode;    /* SRB_MODE_*  */
  int32_t comp;    /* SRB_COMP_*  */
  int32_t dtype;   /* SRB_DTYPE_* */
  int32_t native;  /* `f_native` (calc.py:612-615); honoured for F32 + direct phasor only (Q9) */
  int32_t phasor;  /* SRB_PHASOR_* */
  int32_t omega_uniform; /* host hint: table is an ascending uniform grid (no Features) */
  uint32_t nOmega, nAxis2, nPhi; /* `gridNodeNums`; nAxis2 = nTheta (far) | nRadius (near) */
  uint32_t nSnaps;
  const double* omega;      /* [nOmega]  2*pi*omega            */
  const double* sinTheta;   /* [nAxis2]  far                   */
  const double* cosTheta;   /* [nAxis2]  far                   */
  const double* radius;     /* [nAxis2]  near                  */
  const double* sinPhi;     /* [nPhi]                          */
  const double* cosPhi;     /* [nPhi]                          */
  const double* formFactor; /* [nOmega] or NULL; applied by far cartesian_complex only,
                             as in the reference (kernel_farfield.cl:271,324-325) */
  double L_screen;        /* near: `distanceToScreen`        */
  double dt;              /* `timeStep` (c*dt)               */
  double omega_first_host, omega_last_host; /* host copies of omega[0], omega[nOmega-1] */
} srb_grid;

/* All tracks of this rank, SoA and concatenated: the `args_track` of calc.py:306-307 for
 * every particle at once.  Track t occupies [offsets[t], offsets[t+1]) of each coordinate array. */
typedef struct srb_tracks {
  uint32_t nTracks;
  const double *x, *y, *z, *ux, *uy, *uz;
  const uint64_t* offsets;              /* [nTracks+1] */
  const double* w;                      /* [nTracks] `wp` */
  const uint32_t* itStart;              /* [nTracks] */
  const uint32_t* itEnd;                /* [nTracks] `np.uint32(it_range[-1])` (calc.py:307) */
  const uint32_t* itSnaps;              /* snapshot iterations (calc.py:626-630) */
  uint32_t itSnapsStride;               /* 0: one [nSnaps] table shared by all tracks;
                                           nSnaps: per-track rows (it_range=None, calc.py:297-301) */
  uint64_t totalSteps_host;             /* host copy of offsets[nTracks] (work partitioning) */
} srb_tracks;

/* Statistics of the last srb_integrate on this stream (device memory, 2 x uint64):
 * [0] (node,step) updates that passed the Nyquist guard, [1] updates visited. */

int srb_version(void);
const char* srb_last_error(void);

/* Number of spectra `comp` produces in `mode` (1, 3 or 6); <0 if the combination does not
 * exist (the reference has no near-field spheric kernels, calc.py:342 would raise). */
int srb_num_spectra(int mode, int comp);

/* Bytes of scratch srb_integrate wants for this problem (pre-pass planes, private partial spectra of the
 * particle chunks or, with few particles, partial amplitudes of the time segments).  Never fails for valid
 * inputs; 0 is possible. */
size_t srb_scratch_bytes(const srb_grid* grid, const srb_tracks* tracks);

/* The hot path: replaces the whole `for itr in calc_iterator: ... _process_track` loop of
 * calc.py:257-267 and the kernels it launches.
 *   spectra[n_spectra] : float64 device buffers, each nSnaps*nPhi*nAxis2*nOmega, `+=`
 *   scratch            : device buffer of >= srb_scratch_bytes() (may be NULL if that is 0);
 *                        a smaller buffer is accepted and only reduces parallelism
 *   counters           : device uint64[4] or NULL; zeroed and filled by the call: [0] updates that passed the Nyquist
 *                        guard, [1] updates visited, [2] with phasor = AUTO and two eligible kernels: the kind chosen
 *                        on the device (srb_launch_info.kind == SRB_KIND_ON_DEVICE then), [3] reserved
 * Asynchronous with respect to the host on `stream`: no call of this library synchronises the stream. */
int srb_integrate(const srb_grid* grid, const srb_tracks* tracks, double* const* spectra,
                  int n_spectra, void* scratch, size_t scratch_bytes, uint64_t* counters,
                  void* stream);

/* Same computation through HOST buffers (all pointers of grid/tracks/spectra are host
 * pointers here): uploads, integrates, downloads and adds into the host spectra.
 * Synchronous.  This is the entry a non-Python host would bind. */
int srb_integrate_host(const srb_grid* grid, const srb_tracks* tracks, double* const* spectra,
                       int n_spectra, uint64_t* counters_host, int device);

/* Replaces `_spectr_from_device` + `_gather_result_mpi` plumbing on the device side
 * (calc.py:560-577): dst[nSnaps][nOmega][nAxis2][nPhi] = swapaxes(src,-1,-3). */
int srb_swap_axes(const double* src, double* dst, uint32_t nSnaps, uint32_t nOmega,
                  uint32_t nAxis2, uint32_t nPhi, void* stream);

/* Angle-integrated energy spectrum of snapshot `iSnap`, on the device -- the integrals of utils.py:75-95
 * (`get_energy_spectrum`) without the unit prefactors.  layout = 0: spectra in the device layout
 * (nSnaps, nPhi, nAxis2, nOmega) as srb_integrate leaves them; layout = 1: (nSnaps, nOmega, nAxis2, nPhi) as
 * srb_swap_axes leaves them.  val = sum_k spectra[k] (coherent != 0: sum_k spectra[k]^2, the *_complex comps);
 *   far : out[j] = dphi * sum_phi trapz( 0.5 (val[a+1] + val[a]) * sin(th_mid[a]) ; th_mid ),  th_mid = mid-points of theta
 *   near: out[j] = dphi * sum_phi trapz( val[a] * r[a] ; r )
 * axis2 = theta (far) or radius (near), float64[nAxis2] on the device; out = float64[nOmega] on the device,
 * overwritten.  Deterministic (fixed summation order). */
int srb_energy_spectrum(int mode, int layout, const double* const* spectra, int n_spectra, int coherent,
                        uint32_t nOmega, uint32_t nAxis2, uint32_t nPhi, uint32_t nSnaps, uint32_t iSnap,
                        const double* axis2, double dphi, double* out, void* stream);

/* How the last srb_integrate was configured (for benchmarks/diagnostics). */
typedef struct srb_launch_info {
  int32_t kind;        /* 0 direct, 1 recurrence, 2 literal fp32, 3 pair, 4 pair on the scalar pipe, 5 gridding */
  int32_t tile_width;  /* omega nodes per thread */
  uint32_t chunk_nodes, n_chunks, n_virtual_dirs, n_particle_chunks;
  uint32_t grid_blocks, block_threads, smem_bytes;
  uint32_t kernels_launched;
  uint32_t n_components;  /* far-field amplitude components carried per node (2 transverse | 3) */
  uint32_t n_time_segments; /* > 1: time-axis split (few particles): every track cut into this many step segments, one
                               (track, segment) per particle chunk, partial amplitudes summed before squaring */
} srb_launch_info;
int srb_last_launch(srb_launch_info* info);

/* Measured issue-limited peak of one pipe on the current device, in lane-operations per second:
 * which = 0 FP64 FMA, 1 FP32 FMA, 2 MUFU (sin.approx).  Runs a ~50 ms micro-kernel; synchronous.
 * These are the denominators of the compute roofline (the path is issue-bound, not HBM-bound). */
int srb_pipe_peak(int which, double* ops_per_second);

#ifdef __cplusplus
}
#endif
#endif /* SYNCHRAD_B200_H */
