"""Deterministic input recipes for the BASELINE.json configs.

C1/C2 follow the reference's own test scripts (tests/test_undulator_analytic.py:8-60 and
tests/test_undulator_analytic_near.py:8-64) with the RNG removed (Np=1, gamma == g0 exactly) or
seeded; C3/C5-like generators are synthetic recipes of the named shapes.  Nothing here reads
/root/reference at run time.
"""
import numpy as np


def undulator_tracks(Np=1, near=False, seed=None, K0=0.1, Periods=50, g0=100.0):
    """Planar-undulator tracks of the reference's analytic tests.

    Returns (particleTracks, dt, info).  seed=None -> every particle has gamma == g0 exactly
    (the deterministic variant of BASELINE.md §2); otherwise g0 + 1e-4*g0*randn (seeded).
    """
    dg = 1e-4 * g0
    StepsPerPeriod = 64 if near else 32
    gg = g0 / (1.0 + K0 ** 2 / 2) ** 0.5
    k_res = 2 * gg ** 2
    dt = 1.0 / StepsPerPeriod
    Steps2Do = int((Periods + 2) / dt) + 1

    def ux_fun(z):
        val = K0 * np.sin(2 * np.pi * z)
        val *= (z > 0) * (z < 1.5) * z / 1.5 + (z > 1.5)
        val *= (z > Periods - 1.5) * (z < Periods) * (Periods - z) / 1.5 + (z < Periods - 1.5)
        return val

    t = np.linspace(-1, Periods + 1, Steps2Do)
    y = np.zeros_like(t)
    uy = np.zeros_like(t)
    if seed is None:
        gammas = np.full(Np, g0)
    else:
        gammas = g0 + dg * np.random.RandomState(seed).randn(Np)
    tracks = []
    for g0_p in gammas:
        ggp = g0_p / (1.0 + K0 ** 2 / 2) ** 0.5
        vb = (1.0 - ggp ** -2) ** 0.5
        z = vb * t
        ux = ux_fun(z - 0.5 * dt)
        uz = (g0 ** 2 - 1 - ux ** 2) ** 0.5
        x = ux[0] / g0_p * dt / 2 + np.cumsum(ux / g0_p) * dt
        tracks.append([x, y, z, ux, uy, uz, 1.0, 0])
    info = dict(K0=K0, Periods=Periods, g0=g0, k_res=k_res, Np=Np)
    return tracks, dt, info


def undulator_args(info, near=False, grid=None, L_scr=1e5, dtype='double'):
    k_res, g0 = info['k_res'], info['g0']
    if near:
        g = grid or (128, 256, 32)
        A = {"grid": [(0.02 * k_res, 1.1 * k_res), (0, L_scr * 1 / g0), (0.0, 2 * np.pi), g],
             "mode": "near"}
    else:
        g = grid or (128, 32, 32)
        A = {"grid": [(0.02 * k_res, 1.1 * k_res), (0, 2.0 / g0), (0.0, 2 * np.pi), g]}
    A['dtype'] = dtype
    A['ctx'] = [0, 0]
    return A


def undulator_energy_theory(info, J_in_um):
    """Analytic estimate the reference tests compare with (test_undulator_analytic.py:78-87)."""
    return (info['Np'] * info['k_res'] * J_in_um * (7 * np.pi / 24) / 137.0
            * info['K0'] ** 2 * (1 + info['K0'] ** 2 / 2) * info['Periods'])


def wiggler_tracks(Np=8, n=256, seed=0, K0=20.0, gamma0=1000.0, spread=0.1, si_scale=1.0):
    """Betatron-like (wiggler regime) ensemble: strongly guard-dominated.

    Lengths in units of the oscillation period times `si_scale` (set ~1e-3 to mimic SI-unit
    magnitudes: large omega, small coordinates, |phase| ~ 1e5-1e6).
    """
    rs = np.random.RandomState(seed)
    osc = 4
    dt = osc / (n - 1.0)
    t = np.arange(n) * dt
    tracks = []
    for _ in range(Np):
        g = gamma0 * (1 + spread * rs.randn())
        K = K0 * (1 + 0.1 * rs.randn())
        ph, ps = rs.uniform(0, 2 * np.pi, 2)
        ux = K * np.cos(2 * np.pi * t + ph)
        uy = 0.3 * K * np.sin(2 * np.pi * t + ps)
        uz = np.sqrt(g ** 2 - 1 - ux ** 2 - uy ** 2)
        x = np.cumsum(ux / g) * dt
        y = np.cumsum(uy / g) * dt
        z = np.cumsum(uz / g) * dt
        w = 1.0 + rs.rand()
        tracks.append([x * si_scale, y * si_scale, z * si_scale, ux, uy, uz, w, 0])
    return tracks, dt * si_scale, dict(K0=K0, gamma0=gamma0)


def wiggler_args(info, grid=(64, 8, 8), dtype='double', si_scale=1.0, features=()):
    g, K = info['gamma0'], info['K0']
    w_c = 1.5 * K * g ** 2 / si_scale
    A = {"grid": [(1e-3 * w_c, 1.0 * w_c), (0, 2 * K / g), (0.0, 2 * np.pi), tuple(grid)],
         "dtype": dtype, "ctx": [0, 0]}
    if features:
        A['Features'] = list(features)
    return A


def c5_tracks_numpy(Np, n=1000, seed=1234, dt=0.01):
    """Host (NumPy) version of the C5 synthetic recipe for small parity cases."""
    rs = np.random.RandomState(seed)
    t = (np.arange(n) * dt)
    tracks = []
    for _ in range(Np):
        g = 200 * (1 + 0.05 * rs.randn())
        K = 2 * (1 + 0.1 * rs.randn())
        ph, ps = rs.uniform(0, 2 * np.pi, 2)
        ux = K * np.cos(2 * np.pi * t + ph)
        uy = 0.5 * K * np.sin(2 * np.pi * t + ps)
        uz = np.sqrt(g ** 2 - 1 - ux ** 2 - uy ** 2)
        x = (np.cumsum(ux / g) - 0.5 * ux / g) * dt
        y = (np.cumsum(uy / g) - 0.5 * uy / g) * dt
        z = (np.cumsum(uz / g) - 0.5 * uz / g) * dt
        tracks.append([x, y, z, ux, uy, uz, 1.0, 0])
    return tracks, dt


def c5_args(grid=(256, 32, 32), dtype='double'):
    g, K = 200.0, 2.0
    w1 = 2 * g ** 2 / (1 + K ** 2 / 2)
    return {"grid": [(0.02 * w1, 1.5 * w1), (0, 3 * K / g), (0.0, 2 * np.pi), tuple(grid)],
            "dtype": dtype, "ctx": [0, 0]}


def betatron_tracks(Np=1000, seed=0, Num_osc=4, K0=20.0, energy_MeV=500.0, n_p=8e24, energy_spread=0.1,
                    samples_per_osc=64, substeps=16):
    """C3: matched Gaussian beam performing betatron oscillations in an ion channel, SI
    units, parameters of tutorials/Betatron_Example.ipynb (cells 3-4); tracks from a deterministic
    vectorised leap-frog (momenta at t_k, coordinates staggered by dt/2, as the notebook does).
    Returns (tracks, c*dt, info)."""
    from scipy.constants import c, e, m_e, physical_constants
    r_e = physical_constants['classical electron radius'][0]
    rs = np.random.RandomState(seed)
    g0 = energy_MeV * 1e6 * e / (m_e * c ** 2)
    pz0 = (g0 ** 2 - 1) ** 0.5
    g0m = (1 + pz0 ** 2 + K0 ** 2) ** 0.5
    w_p = c * (4 * np.pi * r_e * n_p) ** 0.5
    w_ch = w_p / (2 * g0m) ** 0.5
    lam_ch = 2 * np.pi * c / w_ch
    w_crit = 1.5 * K0 * g0 ** 2 * w_ch
    T_fin = Num_osc * lam_ch / c
    Nt = int(Num_osc * samples_per_osc)
    dt = T_fin / (Nt - 1)
    R_match = K0 * c / w_p * (2 / g0) ** 0.5
    x = R_match / 2 ** .5 * rs.randn(Np); y = R_match / 2 ** .5 * rs.randn(Np); z = 1e-6 * rs.randn(Np)
    ux = K0 / 2 ** .5 * rs.randn(Np); uy = K0 / 2 ** .5 * rs.randn(Np)
    uz = pz0 * (1 + energy_spread * rs.randn(Np))
    X = np.empty((6, Np, Nt))
    h = dt / substeps
    k = 0.5 * w_p ** 2 / c
    # momenta live at integer times, coordinates at half-integer times (kick-drift leap-frog)
    gam = np.sqrt(1 + ux ** 2 + uy ** 2 + uz ** 2)
    xs, ys, zs = x + 0.5 * h * c * ux / gam, y + 0.5 * h * c * uy / gam, z + 0.5 * h * c * uz / gam
    for it in range(Nt):
        X[3, :, it], X[4, :, it], X[5, :, it] = ux, uy, uz
        for s in range(substeps):
            if s == substeps // 2:
                X[0, :, it], X[1, :, it], X[2, :, it] = (xs - 0.5 * h * c * ux / gam, ys - 0.5 * h * c * uy / gam,
                                                       zs - 0.5 * h * c * uz / gam)
            ux = ux - h * k * xs; uy = uy - h * k * ys
            gam = np.sqrt(1 + ux ** 2 + uy ** 2 + uz ** 2)
            xs = xs + h * c * ux / gam; ys = ys + h * c * uy / gam; zs = zs + h * c * uz / gam
    tracks = [[X[0, i].copy(), X[1, i].copy(), X[2, i].copy(), X[3, i].copy(), X[4, i].copy(), X[5, i].copy(), 1.0, 0]
              for i in range(Np)]
    info = dict(K0=K0, gamma0=g0, omega_crit_1m=w_crit / (2 * np.pi * c))
    return tracks, c * dt, info


def betatron_args(info, grid=(256, 32, 32), dtype='double'):
    wc = info['omega_crit_1m']
    return {"grid": [(1e-3 * wc, wc), (0, 2 * info['K0'] / info['gamma0']), (0.0, 2 * np.pi), tuple(grid)],
            "dtype": dtype, "ctx": [0, 0]}


def spiral_tracks(Np=10000, seed=0, N_winds=1, L_b=0.4e-6, energy_MeV=100.0, n_p=5e24, K0=4.0, Num_osc=3,
                  samples_per_osc=64, energy_spread=0.001, pr_spread=0.0, substeps=16):
    """C4 (BASELINE configs[3]): the spiral beam of tutorials/Spiral_Beam_Part1.ipynb:56-75 (cells 2-3:
    100 MeV, n_p = 5e24 m^-3, K0 = 4, one spiral winding over 0.4 um, 3 channel oscillations x 64 samples = 192
    samples) in SI units, seeded.  The notebook integrates each particle with scipy's Radau solver (cell 4); the
    tracks here come from a deterministic vectorised kick-drift leap-frog of the same equations of motion
    (momenta at t_k, coordinates staggered by dt/2 as the notebook stores them) -- parity needs identical inputs
    on both sides, not the notebook's integrator.  Returns (tracks, c*dt, info)."""
    from scipy.constants import c, e, m_e, physical_constants
    r_e = physical_constants['classical electron radius'][0]
    rs = np.random.RandomState(seed)
    g0 = energy_MeV * 1e6 * e / (m_e * c ** 2)
    pz0 = (g0 ** 2 - K0 ** 2 - 1) ** .5
    w_p = c * (4 * np.pi * r_e * n_p) ** 0.5
    w_ch = w_p / (2 * g0) ** 0.5
    R_match = K0 * c / w_p * (2 / g0) ** 0.5
    lam_ch = 2 * np.pi * c / w_ch
    w_crit = 1.5 * K0 * g0 ** 2 * w_ch
    k_sp = 2 * np.pi / (L_b / N_winds)
    beta_phs = pz0 / g0 + w_ch / (k_sp * c)
    theta_vc = float(np.arccos(1 / beta_phs))
    theta_beta = K0 / g0
    T_fin = Num_osc * lam_ch / c
    Nt = int(Num_osc * samples_per_osc)
    dt = T_fin / (Nt - 1)
    z = np.linspace(0, L_b, Np)
    if Np > 1:
        z = z + 0.5 * (z[1] - z[0]) * (rs.rand(Np) - 0.5)
    x = R_match * np.sin(k_sp * z); y = R_match * np.cos(k_sp * z)
    ux = -K0 * np.cos(k_sp * z) * (1 + pr_spread * rs.randn(Np))
    uy = K0 * np.sin(k_sp * z) * (1 + pr_spread * rs.randn(Np))
    uz = pz0 * (1 + energy_spread * rs.randn(Np))
    X = np.empty((6, Np, Nt))
    h = dt / substeps
    k = 0.5 * w_p ** 2 / c
    gam = np.sqrt(1 + ux ** 2 + uy ** 2 + uz ** 2)
    xs, ys, zs = x + 0.5 * h * c * ux / gam, y + 0.5 * h * c * uy / gam, z + 0.5 * h * c * uz / gam
    for it in range(Nt):
        X[3, :, it], X[4, :, it], X[5, :, it] = ux, uy, uz
        for s in range(substeps):
            if s == substeps // 2:
                X[0, :, it], X[1, :, it], X[2, :, it] = (xs - 0.5 * h * c * ux / gam, ys - 0.5 * h * c * uy / gam,
                                                       zs - 0.5 * h * c * uz / gam)
            ux = ux - h * k * xs; uy = uy - h * k * ys
            gam = np.sqrt(1 + ux ** 2 + uy ** 2 + uz ** 2)
            xs = xs + h * c * ux / gam; ys = ys + h * c * uy / gam; zs = zs + h * c * uz / gam
    tracks = [[X[0, i].copy(), X[1, i].copy(), X[2, i].copy(), X[3, i].copy(), X[4, i].copy(), X[5, i].copy(), 1.0, 0]
              for i in range(Np)]
    info = dict(K0=K0, gamma0=g0, omega_crit_1m=w_crit / (2 * np.pi * c), theta_vc=theta_vc, theta_beta=theta_beta,
                L_b=L_b)
    return tracks, c * dt, info


def spiral_args(info, grid=(512, 64, 64), dtype='float'):
    """Grid of tutorials/Spiral_Beam_Part1.ipynb:255-264 (cell 6) at the BASELINE C4 node counts."""
    th_max = 1.2 * max(info['theta_beta'], info['theta_vc'])
    return {"grid": [(1., 3 * info['omega_crit_1m']), (0, th_max), (0.0, 2 * np.pi), tuple(grid)],
            "dtype": dtype, "ctx": [0, 0]}
