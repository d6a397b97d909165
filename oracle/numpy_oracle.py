"""ORACLE — TEST INFRASTRUCTURE ONLY.

Independent, vectorised NumPy restatement of the reference's `total` / `cartesian_comps`
kernels (kernel_farfield.cl:30-108,208-222 and kernel_nearfield.cl:29-103,197-211): arrays in
the device layout (nPhi, nAxis2, nOmega) (calc.py:455), one Python loop over time steps, a
boolean mask for the Nyquist guard, phasePrev updated unconditionally.  It exists to
cross-check oracle_kernels.cpp (two independent restatements agreeing bit-for-bit in double is
the strongest pin available without an OpenCL runtime) and for tiny cases.  Pure NumPy: every
binary operation rounds once, so there is no FMA contraction by construction.
"""
import numpy as np


def _dot(a, b):
    return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]


def particle(mode, comp, T, track, wp, itStart, itEnd, omega2pi, axisA, axisB, sinPhi, cosPhi,
             L_screen, dt, itSnaps):
    """One kernel launch.  Returns a list of 1 or 3 arrays (nSnaps, nPhi, nAxis2, nOmega)."""
    x, y, z, ux, uy, uz = [np.asarray(c).astype(T) for c in track]
    nSteps = x.size
    om = omega2pi.astype(T)[None, None, :]
    sP = sinPhi.astype(T)[:, None, None]
    cP = cosPhi.astype(T)[:, None, None]
    shape = (sinPhi.size, axisA.size, omega2pi.size)
    nSnaps = len(itSnaps)
    dt = T(dt)
    wp = T(wp)
    one = T(1)
    if mode == 'far':
        sT = axisA.astype(T)[None, :, None]
        cT = axisB.astype(T)[None, :, None]
        n = [np.broadcast_to(sT * cP, shape), np.broadcast_to(sT * sP, shape),
             np.broadcast_to(cT + 0 * cP, shape)]
    else:
        r = axisA.astype(T)[None, :, None]
        X = [np.broadcast_to(r * cP, shape), np.broadcast_to(r * sP, shape),
             np.full(shape, T(L_screen))]
    dtInv = one / dt
    wpdt2 = wp * dt * dt
    phasePrev = np.zeros(shape, T)
    Re = [np.zeros(shape, T) for _ in range(3)]
    Im = [np.zeros(shape, T) for _ in range(3)]
    ncomp = 1 if comp == 'total' else 3
    out = [np.zeros((nSnaps,) + shape, T) for _ in range(ncomp)]
    iSnap = 0
    while iSnap < nSnaps and not (itStart < itSnaps[iSnap]):
        iSnap += 1
    PI = T(np.pi)
    for it in range(0, max(int(itEnd) - 1, 0)):
        it_glob = itStart + it
        if it < nSteps - 1:
            time = T(it_glob) * dt
            xl = [x[it], y[it], z[it]]
            if mode == 'far':
                phase = om * (time - _dot(xl, n))
            else:
                rV = [X[k] - xl[k] for k in range(3)]
                rL = np.sqrt(_dot(rV, rV))
                phase = om * (time + rL)
            dPhase = np.abs(phase - phasePrev)
            phasePrev = phase
            m = dPhase < PI
            s = np.sin(phase)
            c = np.cos(phase)
            u = [ux[it], uy[it], uz[it]]
            gi = one / np.sqrt(one + _dot(u, u))
            u = [uk * gi for uk in u]
            if mode == 'far':
                un = [ux[it + 1], uy[it + 1], uz[it + 1]]
                gi = one / np.sqrt(one + _dot(un, un))
                un = [uk * gi for uk in un]
                a = [(un[k] - u[k]) * dtInv for k in range(3)]
                u = [T(0.5) * (un[k] + u[k]) for k in range(3)]
                c1 = _dot(a, n)
                c2 = one - _dot(u, n)
                c2 = one / c2
                c1 = c1 * c2 * c2
                for k in range(3):
                    amp = c1 * (n[k] - u[k]) - c2 * a[k]
                    Re[k] = np.where(m, Re[k] + amp * c, Re[k])
                    Im[k] = np.where(m, Im[k] + amp * s, Im[k])
            else:
                rInv = one / rL
                for k in range(3):
                    nk = rInv * rV[k]
                    c1 = (om * rInv) * (u[k] - nk)
                    c2 = (rInv * rInv) * nk
                    Re[k] = np.where(m, Re[k] + ((-c1) * s + c2 * c), Re[k])
                    Im[k] = np.where(m, Im[k] + (c1 * c + c2 * s), Im[k])
        if iSnap < nSnaps and it_glob + 2 == itSnaps[iSnap]:
            if comp == 'total':
                out[0][iSnap] += wpdt2 * (_dot(Re, Re) + _dot(Im, Im))
            else:
                for k in range(3):
                    out[k][iSnap] += wpdt2 * (Re[k] * Re[k] + Im[k] * Im[k])
            iSnap += 1
    return out
