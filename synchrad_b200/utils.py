"""Host-side glue the reference's analytic tests call on the finished spectrum
(`calc.get_energy(lambda0_um=1)`, `J_in_um`; tests/test_undulator_analytic.py:4,77).

Only the unit scalings and the theta/R/phi/omega integrals of /root/reference/synchrad/utils.py
:16-19, :23-102 are provided (NumPy on a <= 50 MB array).  `on_device=True` evaluates the
angle integrals on the GPU from the device-resident result (`srb_energy_spectrum`).
Spot maps, VTK export and the track converters are out of scope.
"""
import numpy as np
from scipy.constants import m_e, c, e, epsilon_0, hbar
from scipy.constants import alpha as alpha_fs

J_in_um = 2e6 * np.pi * hbar * c
r_e = e ** 2 / (4 * np.pi * epsilon_0 * m_e * c ** 2)
omega_1m = 2 * np.pi * c
energy_1m_eV = omega_1m * hbar / e


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
            inner = np.trapezoid(v_mid * np.sin(th_mid)[None, :, None], th_mid, axis=1)
        else:
            r = self.Args['radius']
            inner = np.trapezoid(val * r[None, :, None], r, axis=1)
        return self.Args['dph'] * inner.sum(-1)

    def get_energy(self, spect_filter=None, phot_num=False, lambda0_um=None, on_device=False, **kw):
        val = self.get_energy_spectrum(spect_filter=spect_filter, phot_num=phot_num,
                                       lambda0_um=lambda0_um, on_device=on_device, **kw)
        return np.trapezoid(val, self.Args['omega'])

    def get_spectral_axis(self):
        return self.Args['omega']
