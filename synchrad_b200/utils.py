"""Host-side glue the reference's analytic tests call on the finished spectrum
(`calc.get_energy(lambda0_um=1)`, `J_in_um`; tests/test_undulator_analytic.py:4,77).

The unit scalings, the theta/R/phi/omega integrals and the spot maps of /root/reference/synchrad/utils.py
:16-19, :23-160 are provided (NumPy on a <= 50 MB array) -- what the reference's tests and
tutorial notebooks call on a finished calculation.  `on_device=True` evaluates the angle integrals on the GPU
from the device-resident result (`srb_energy_spectrum`).  Every method is checked against the
reference's own utils.py on the reference's stored spectra (tests/test_reference_pin.py).
VTK export (needs tvtk) and the track converters are out of scope.
"""
import numpy as np
from scipy.constants import m_e, c, e, epsilon_0, hbar
from scipy.constants import alpha as alpha_fs

J_in_um = 2e6 * np.pi * hbar * c
r_e = e ** 2 / (4 * np.pi * epsilon_0 * m_e * c ** 2)
omega_1m = 2 * np.pi * c
energy_1m_eV = omega_1m * hbar / e


_trapz = getattr(np, 'trapezoid', None) or np.trapz      # NumPy >= 2.0 / 1.x (the reference calls np.trapz)


class Utilities:
    """Mixin base of SynchRad (reference: `class SynchRad(Utilities)`, calc.py:21)."""

    def get_full_spectrum(self, spect_filter=None, phot_num=False, lambda0_um=None,
                          normalize_to_weights=False, comp='total', iteration=-1):
        rad = self.Data['radiation']
        coherent = self.Args['comp'].split('_')[-1] == 'complex'
        if coherent and comp != 'total':
            val = rad[comp + 're'][iteration].astype(np.complex128) \
                + 1.0j * rad[comp + 'im'][iteration].astype(np.complex128)
        elif comp == 'total':
            val = 0.0
            for key in rad:
                part = rad[key][iteration].astype(np.double)
                val = val + (part ** 2 if coherent else part)
        else:
            val = 0.0 + rad[comp][iteration].astype(np.double)

        if self.Args['mode'] == 'far':
            val = alpha_fs / (4 * np.pi ** 2) * val
        elif self.Args['mode'] == 'near':
            val = alpha_fs * np.pi / 4 * val
            val = val / (2 * np.pi) ** 2
        if spect_filter is not None:
            val = val * spect_filter
        if normalize_to_weights:
            val = val / self.total_weight
        if phot_num:
            val = val / self.Args['omega'][:, None, None]
        elif lambda0_um is not None:
            val = val * (J_in_um / lambda0_um)
        return val

    def _energy_spectrum_on_device(self, phot_num=False, lambda0_um=None, normalize_to_weights=False,
                                   comp='total', iteration=-1):
        """`get_energy_spectrum` evaluated on the GPU from the spectra the last `calculate_spectrum` left there
        (SURVEY §8f-4): only the [n_omega] result crosses PCIe.  Same prefactors as `get_full_spectrum`."""
        dev = getattr(self, '_dev_radiation', None)
        if dev is None:
            raise RuntimeError('no device-resident spectrum: run calculate_spectrum first (rank 0 holds the result)')
        if comp != 'total':
            raise ValueError("on_device=True integrates comp='total' (the sum over the stored components)")
        from . import engine
        coherent = self.Args['comp'].split('_')[-1] == 'complex'
        far = self.Args['mode'] == 'far'
        n_w, n_2, n_p = (int(v) for v in self.Args['gridNodeNums'])
        spectra = list(dev.values())
        nSnaps = int(spectra[0].shape[0])
        val = engine.energy_spectrum(self.Args['mode'], spectra, coherent, nSnaps, n_w, n_2, n_p, iteration,
                                     self.Args['theta'] if far else self.Args['radius'], float(self.Args['dph']))
        val = val.cpu().numpy()
        val = alpha_fs / (4 * np.pi ** 2) * val if far else alpha_fs * np.pi / 4 * val / (2 * np.pi) ** 2
        if normalize_to_weights:
            val = val / self.total_weight
        if phot_num:
            val = val / np.asarray(self.Args['omega'], dtype=np.double)
        elif lambda0_um is not None:
            val = val * (J_in_um / lambda0_um)
        return val

    def get_energy_spectrum(self, spect_filter=None, phot_num=False, lambda0_um=None, on_device=False, **kw):
        if on_device:
            if spect_filter is not None:
                raise ValueError('on_device=True does not take a spect_filter; use the host path')
            return self._energy_spectrum_on_device(phot_num=phot_num, lambda0_um=lambda0_um, **kw)
        val = self.get_full_spectrum(spect_filter=spect_filter, phot_num=phot_num,
                                     lambda0_um=lambda0_um, **kw)
        if self.Args['mode'] == 'far':
            th = self.Args['theta']
            th_mid = 0.5 * (th[1:] + th[:-1])
            v_mid = 0.5 * (val[:, 1:, :] + val[:, :-1, :])
            inner = _trapz(v_mid * np.sin(th_mid)[None, :, None], th_mid, axis=1)
        else:
            r = self.Args['radius']
            inner = _trapz(val * r[None, :, None], r, axis=1)
        return self.Args['dph'] * inner.sum(-1)

    def get_energy(self, spect_filter=None, phot_num=False, lambda0_um=None, on_device=False, **kw):
        val = self.get_energy_spectrum(spect_filter=spect_filter, phot_num=phot_num,
                                       lambda0_um=lambda0_um, on_device=on_device, **kw)
        return _trapz(val, self.Args['omega'])

    def get_spot(self, k0=None, spect_filter=None, phot_num=False, lambda0_um=None, **kw):
        """Angular map (theta|R, phi): the spectrum integrated over omega, or its slice at the node closest to
        `k0` (utils.py:104-127).  As in the reference, a single-node omega axis is weighted with `dw`, and a `k0`
        whose nearest-from-below node is the last one raises IndexError."""
        val = self.get_full_spectrum(spect_filter=spect_filter, phot_num=phot_num,
                                     lambda0_um=lambda0_um, **kw)
        omega = self.Args['omega']
        if k0 is None:
            if val.shape[0] > 1:
                return _trapz(val, omega, axis=0)
            return val[0] * self.Args['dw']
        below = int((omega < k0).sum())
        if np.abs(omega[below + 1] - k0) < np.abs(omega[below] - k0):
            below += 1
        return val[below]

    def get_spot_cartesian(self, k0=None, th_part=1.0, bins=(200, 200), spect_filter=None, phot_num=False,
                           lambda0_um=None, **kw):
        """`get_spot` resampled from the polar (theta|R, phi) nodes onto a Cartesian bins[0] x bins[1] raster of
        half-width th_part * max(theta|R) by linear (Delaunay) interpolation, zero outside the hull
        (utils.py:129-158).  Returns (map, [-m, m, -m, m])."""
        from scipy.interpolate import griddata
        spot = self.get_spot(spect_filter=spect_filter, k0=k0, phot_num=phot_num, lambda0_um=lambda0_um, **kw)
        if self.Args['mode'] == 'far':
            rho = self.Args['theta']
        elif self.Args['mode'] == 'near':
            rho = self.Args['radius']
        else:
            raise ValueError("get_spot_cartesian is for 'far' and 'near' modes only")
        phi_nodes, rho_nodes = np.meshgrid(self.Args['phi'], rho)          # both (n_rho, n_phi), like the spot
        points = ((rho_nodes * np.cos(phi_nodes)).ravel(), (rho_nodes * np.sin(phi_nodes)).ravel())
        half = th_part * rho_nodes.max()
        raster = np.mgrid[-half:half:bins[0] * 1j, -half:half:bins[1] * 1j]
        out = griddata(points, spot.ravel(), (raster[0].ravel(), raster[1].ravel()), fill_value=0.,
                       method='linear').reshape(raster[0].shape)
        return out, np.array([-half, half, -half, half])

    def exportToVTK(self, *a, **kw):
        """utils.py:174-228 needs tvtk (mayavi); like the reference without it, report and return."""
        print('TVTK API is not found')

    def get_spectral_axis(self):
        return self.Args['omega']
