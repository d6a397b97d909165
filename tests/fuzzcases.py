"""Random small problems for differential tests (emulated kernels on CPU, CUDA kernels on GPU vs the
oracle): ragged track lengths, it_start, global/per-track ranges, snapshot counts, every comp, near
and far, uniform / log / wavelength grids, omega ranges that make the Nyquist guard bite."""
import numpy as np

import cases


def rand_case(rs):
    nt = rs.randint(1, 5)
    base, dt = cases.c5_tracks_numpy(nt, 120, seed=rs.randint(1 << 30))
    tracks = []
    for t in base:
        n = rs.randint(1, 120)
        tr = [c[:n].copy() for c in t[:6]] + [float(rs.uniform(0.5, 2.0))]
        if rs.rand() < 0.7:
            tr.append(int(rs.randint(0, 40)))
        tracks.append(tr)
    mode = 'far' if rs.rand() < 0.7 else 'near'
    nw = int(rs.choice([1, 2, 3, 17, 40, 64, 100, 130, 256, 300]))
    grid = (nw, int(rs.randint(1, 4)), int(rs.randint(1, 4)))
    A = cases.c5_args(grid=grid)
    kw = {}
    if mode == 'near':
        A['mode'] = 'near'
        L = float(rs.choice([2.0, 50.0, 1e4]))
        A['grid'][1] = (0.0, 0.03 * L)
        kw['L_screen'] = L
        comp = rs.choice(['total', 'cartesian', 'cartesian_complex'])
    else:
        comp = rs.choice(['total', 'cartesian', 'cartesian_complex', 'spheric', 'spheric_complex'])
    if rs.rand() < 0.3 and nw > 2:
        A['Features'] = [str(rs.choice(['logGrid', 'wavelengthGrid']))]
    if rs.rand() < 0.5:
        A['grid'][0] = (A['grid'][0][0], A['grid'][0][1] * float(rs.choice([3, 30, 300])))
    kw['comp'] = str(comp)
    kw['nSnaps'] = int(rs.choice([1, 1, 2, 3, 7]))
    if rs.rand() < 0.5:
        a = int(rs.randint(0, 20))
        kw['it_range'] = (a, a + int(rs.randint(1, 200)))
    if rs.rand() < 0.3:
        kw['sigma_particle'] = 1e-5
    return A, tracks, dt, kw


def vector_errors(got, ref):
    """max |d| / (max over ALL components of |ref|): components that vanish by orthogonality (spheric r,
    Cartesian z on the axis) are rounding residue in the reference and exact zeros in the transverse-basis
    kernels, so they are judged on the scale of the whole vector."""
    big = max(np.abs(r).max() for r in ref.values())
    if big == 0:
        return max(np.abs(got[k]).max() for k in ref)
    return max(np.abs(got[k] - ref[k]).max() for k in ref) / big
