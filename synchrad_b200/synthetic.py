"""Synthetic track generators of the named benchmark shapes, written with
torch so the same code fills a GPU-resident batch or a host batch.

C5 recipe (normalised units, length unit = oscillation period): for particle p
    gamma_p = 200 (1 + 0.05 xi1),  K_p = 2 (1 + 0.1 xi2),  phi_p, psi_p ~ U[0, 2 pi)
    ux = K_p cos(2 pi t + phi_p),  uy = 0.5 K_p sin(2 pi t + psi_p),
    uz = sqrt(gamma_p^2 - 1 - ux^2 - uy^2),  (x, y, z) = cumsum(u / gamma_p) dt (staggered half a step)
    dt = 0.01, n = 10^4 samples, w = 1.
Grid (256, 32, 32): omega in [0.02, 1.5] * 2 gamma^2 / (1 + K^2/2) at gamma = 200, K = 2;
theta in [0, 3K/gamma]; phi in [0, 2 pi).  Guard-pass fraction ~ 1, max |phase| ~ 1e4.
"""
import numpy as np
import torch

C5_DT = 0.01
C5_STEPS = 10_000


def c5_args(grid=(256, 32, 32), dtype='double'):
    g, K = 200.0, 2.0
    w1 = 2 * g ** 2 / (1 + K ** 2 / 2)
    return {"grid": [(0.02 * w1, 1.5 * w1), (0, 3 * K / g), (0.0, 2 * np.pi), tuple(grid)],
            "dtype": dtype}


def c5_batch(n_particles, n_steps=C5_STEPS, seed=1234, device='cpu', dt=C5_DT, chunk=2048):
    """Returns a dict in the C-ABI SoA layout with torch tensors on `device`:
    x,y,z,ux,uy,uz [n_particles*n_steps] float64; offsets int64 [n+1]; w float64 [n];
    itStart,itEnd int32 [n]; itSnaps int32 [n,1] (per-track range, nSnaps=1); n,total,snapStride."""
    dev = torch.device(device)
    gen = torch.Generator(device='cpu')
    gen.manual_seed(int(seed))
    xi = torch.randn(n_particles, 2, generator=gen, dtype=torch.float64)
    ph = torch.rand(n_particles, 2, generator=gen, dtype=torch.float64) * (2 * np.pi)
    total = n_particles * n_steps
    out = {k: torch.empty(total, dtype=torch.float64, device=dev) for k in ('x', 'y', 'z', 'ux', 'uy', 'uz')}
    t = (torch.arange(n_steps, dtype=torch.float64, device=dev) * dt)[None, :]
    for a in range(0, n_particles, chunk):
        b = min(a + chunk, n_particles)
        g = (200.0 * (1 + 0.05 * xi[a:b, 0])).to(dev)[:, None]
        K = (2.0 * (1 + 0.1 * xi[a:b, 1])).to(dev)[:, None]
        p0 = ph[a:b, 0].to(dev)[:, None]
        p1 = ph[a:b, 1].to(dev)[:, None]
        ux = K * torch.cos(2 * np.pi * t + p0)
        uy = 0.5 * K * torch.sin(2 * np.pi * t + p1)
        uz = torch.sqrt(g * g - 1 - ux * ux - uy * uy)
        sl = slice(a * n_steps, b * n_steps)
        for name, u in (('x', ux), ('y', uy), ('z', uz)):
            v = u / g
            out[name][sl] = ((torch.cumsum(v, dim=1) - 0.5 * v) * dt).reshape(-1)
            out['u' + name][sl] = u.reshape(-1)
    out['offsets'] = torch.arange(n_particles + 1, dtype=torch.int64, device=dev) * n_steps
    out['w'] = torch.ones(n_particles, dtype=torch.float64, device=dev)
    out['itStart'] = torch.zeros(n_particles, dtype=torch.int32, device=dev)
    out['itEnd'] = torch.full((n_particles,), n_steps, dtype=torch.int32, device=dev)
    out['itSnaps'] = torch.full((n_particles, 1), n_steps, dtype=torch.int32, device=dev)
    out['n'], out['total'], out['snapStride'] = n_particles, total, 1
    return out


def batch_to_track_list(batch, first=0, count=None):
    """View a (host) batch as the reference's list-of-tracks input (NumPy arrays, no copies)."""
    n = batch['n'] if count is None else min(batch['n'], first + count)
    off = batch['offsets'].cpu().numpy()
    arrs = {k: batch[k].cpu().numpy() for k in ('x', 'y', 'z', 'ux', 'uy', 'uz')}
    w = batch['w'].cpu().numpy()
    return [[arrs[k][off[i]:off[i + 1]] for k in ('x', 'y', 'z', 'ux', 'uy', 'uz')] + [float(w[i]), 0]
            for i in range(first, n)]
