"""Host-side glue the reference's analytic tests call on the finished spectrum
(`calc.get_energy(lambda0_um=1)`, `J_in_um`; tests/test_undulator_analytic.py:4,77).

Only the unit scalings and the theta/R/phi/omega integrals of /root/reference/synchrad/utils.py
:16-19, :23-102 are provided (NumPy on a <= 50 MB array; not an acceleration target, SURVEY §2
row 6).  Spot maps, VTK export and the track converters are out of scope.
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

    def get_energy_spectrum(self, spect_filter=None, phot_num=False, lambda0_um=None, **kw):
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

    def get_energy(self, spect_filter=None, phot_num=False, lambda0_um=None, **kw):
        val = self.get_energy_spectrum(spect_filter=spect_filter, phot_num=phot_num,
                                       lambda0_um=lambda0_um, **kw)
        return np.trapezoid(val, self.Args['omega'])

    def get_spectral_axis(self):
        return self.Args['omega']
