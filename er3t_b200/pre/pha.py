"""
Phase-function objects with the contract of er3t.pre.pha (er3t/pre/pha/pha_hg.py:31-66, pha_mie.py:72-223):
`.data['id'|'ang'|'pha'|'asy'|'ssa'|'ref']['data']`, pha of shape (Nang, Npf).

`pha_hg` restates the reference's analytic table.  `pha_mie_wc` replaces the look-up in libRadtran's wc.sol.mie.cdf
(absent) by an own Lorenz-Mie computation (Bohren & Huffman recurrences) integrated over a gamma size
distribution, on the reference's default 498-angle grid (pha_mie.py:106-113).
"""

import numpy as np

__all__ = ['cal_hg_pha_func', 'pha_hg', 'pha_mie_wc', 'mie_angles_default']


def cal_hg_pha_func(asy, ang):
    """Henyey-Greenstein phase function normalised to int P dmu = 1 (er3t/pre/pha/pha_hg.py:10-27)."""
    mu = np.cos(np.deg2rad(ang))
    return 0.5 * (1.0 - asy ** 2.0) / ((1.0 - 2.0 * asy * mu + asy ** 2.0) ** 1.5)


class pha_hg:

    ID = 'Henyey-Greenstein'

    def __init__(self, asy_params=(-0.85, 0.85), angles=None):
        if angles is None:
            angles = np.linspace(0.0, 180.0, 1801)
        asy_params = np.array(asy_params, dtype=np.float64)
        angles = np.array(angles, dtype=np.float64)
        pha = np.stack([cal_hg_pha_func(g, angles) for g in asy_params], axis=1)
        self.data = {
            'id': {'data': 'HG', 'name': 'Henyey-Greenstein', 'unit': 'N/A'},
            'ang': {'data': angles, 'name': 'Angle', 'unit': 'degree'},
            'asy': {'data': asy_params, 'name': 'Asymmetry parameter', 'unit': 'N/A'},
            'pha': {'data': pha, 'name': 'Phase function', 'unit': 'N/A'},
        }


def mie_angles_default():
    """The reference's default Mie angle grid, 498 angles (er3t/pre/pha/pha_mie.py:106-113)."""
    return np.concatenate((np.arange(0.0, 2.0, 0.01), np.arange(2.0, 5.0, 0.05), np.arange(5.0, 10.0, 0.1),
                           np.arange(10.0, 15.0, 0.5), np.arange(15.0, 176.0, 1.0), np.arange(176.0, 180.1, 0.25)))


def _refractive_index_water(wvl_nm):
    """Hale & Querry (1973) water refractive index, the same table er3t/pre/sfc/util.py:58-88 uses (visible-NIR part)."""
    w = np.array([400., 450., 500., 550., 600., 650., 700., 750., 800., 850., 900., 950., 1000., 1200., 1400., 1600., 1800., 2000., 2200.])
    nr = np.array([1.339, 1.337, 1.335, 1.333, 1.332, 1.331, 1.331, 1.330, 1.329, 1.329, 1.328, 1.327, 1.327, 1.324, 1.321, 1.317, 1.312, 1.306, 1.296])
    ni = np.array([1.86e-9, 1.02e-9, 1.00e-9, 1.96e-9, 1.09e-8, 1.64e-8, 3.35e-8, 1.56e-7, 1.25e-7, 2.93e-7, 4.86e-7, 2.93e-6, 2.89e-6,
                   9.89e-6, 1.38e-4, 8.55e-5, 1.15e-4, 1.10e-3, 2.89e-4])
    return np.interp(wvl_nm, w, nr), np.interp(wvl_nm, w, ni)


def _mie_s11(x, m, mu):
    """Unpolarised scattered intensity (|S1|^2 + |S2|^2)/2 at cos(angle) = mu, plus Qext, Qsca for size parameters x
    (1-D array) and complex index m.  Vectorised over sizes; upward recurrence for psi/xi, downward for D_n."""
    x = np.asarray(x, dtype=np.float64)
    nx = x.size
    nmax = int(np.max(x) + 4.05 * np.max(x) ** (1.0 / 3.0) + 2)
    nstop = (x + 4.05 * x ** (1.0 / 3.0) + 2).astype(int)
    mx = m * x
    nmx = int(max(nmax, np.max(np.abs(mx))) + 16)
    # logarithmic derivative D_n(mx), downward
    D = np.zeros((nmx + 1, nx), dtype=np.complex128)
    for n in range(nmx, 0, -1):
        D[n - 1] = n / mx - 1.0 / (D[n] + n / mx)
    psi0, psi1 = np.cos(x), np.sin(x)
    chi0, chi1 = -np.sin(x), np.cos(x)
    xi1 = psi1 - 1j * chi1
    s1 = np.zeros((nx, mu.size), dtype=np.complex128)
    s2 = np.zeros((nx, mu.size), dtype=np.complex128)
    pi0 = np.zeros_like(mu)
    pi1 = np.ones_like(mu)
    qext = np.zeros(nx)
    qsca = np.zeros(nx)
    for n in range(1, nmax + 1):
        psi = (2.0 * n - 1.0) / x * psi1 - psi0
        chi = (2.0 * n - 1.0) / x * chi1 - chi0
        xi = psi - 1j * chi
        dn = D[n]
        an = ((dn / m + n / x) * psi - psi1) / ((dn / m + n / x) * xi - xi1)
        bn = ((m * dn + n / x) * psi - psi1) / ((m * dn + n / x) * xi - xi1)
        act = (n <= nstop)
        an = np.where(act, an, 0.0)
        bn = np.where(act, bn, 0.0)
        tau = n * mu * pi1 - (n + 1.0) * pi0
        fn = (2.0 * n + 1.0) / (n * (n + 1.0))
        s1 += fn * (an[:, None] * pi1[None, :] + bn[:, None] * tau[None, :])
        s2 += fn * (an[:, None] * tau[None, :] + bn[:, None] * pi1[None, :])
        qext += (2.0 * n + 1.0) * (an + bn).real
        qsca += (2.0 * n + 1.0) * (np.abs(an) ** 2 + np.abs(bn) ** 2)
        pi_next = ((2.0 * n + 1.0) * mu * pi1 - (n + 1.0) * pi0) / n
        pi0, pi1 = pi1, pi_next
        psi0, psi1 = psi1, psi
        chi0, chi1 = chi1, chi
        xi1 = psi1 - 1j * chi1
    qext *= 2.0 / x ** 2
    qsca *= 2.0 / x ** 2
    return 0.5 * (np.abs(s1) ** 2 + np.abs(s2) ** 2), qext, qsca


class pha_mie_wc:

    """
    Water-cloud Mie phase functions for effective radii `reff` (um) at `wavelength` (nm).

    Gamma size distribution n(r) ~ r^((1-3 veff)/veff) exp(-r / (reff veff)) with effective variance `veff`
    (default 0.1, the value libRadtran's wc.sol.mie.cdf tables were made with), integrated on `nr` radii.
    Output keys match er3t/pre/pha/pha_mie.py:206-216: 'id', 'wvl0', 'wvl', 'ang', 'pha' (Nang, Nreff),
    'ssa' (Nreff), 'asy' (Nreff), 'ref' (Nreff).  pha is normalised like the reference's HG table: int P dmu = 2.
    """

    ID = 'Mie (Water Clouds)'

    def __init__(self, wavelength=555.0, angles=None, reff=None, veff=0.1, nr=128, verbose=False, **kwargs):
        if angles is None:
            angles = mie_angles_default()
        if reff is None:
            reff = np.arange(1.0, 26.0, 1.0)
        self.wvl0 = float(wavelength)
        ang = np.asarray(angles, dtype=np.float64)
        reff = np.asarray(reff, dtype=np.float64)
        mu = np.cos(np.deg2rad(ang))
        nr_, ni_ = _refractive_index_water(self.wvl0)
        m = complex(nr_, ni_)
        lam_um = self.wvl0 * 1.0e-3
        pha = np.zeros((ang.size, reff.size))
        ssa = np.zeros(reff.size)
        asy = np.zeros(reff.size)
        for i, re in enumerate(reff):
            b = re * veff
            alpha = (1.0 - 3.0 * veff) / veff
            r = np.exp(np.linspace(np.log(max(0.02, re * 0.04)), np.log(re * 3.6), nr))
            wts = r ** alpha * np.exp(-r / b) * r                     # * r: d(ln r) integration measure
            x = 2.0 * np.pi * r / lam_um
            # sizes are processed in blocks of similar size parameter so that each block stops its series early
            s11 = np.zeros((nr, mu.size)); qe = np.zeros(nr); qs = np.zeros(nr)
            with np.errstate(all='ignore'):
                for j0 in range(0, nr, 16):
                    sl = slice(j0, min(nr, j0 + 16))
                    s11[sl], qe[sl], qs[sl] = _mie_s11(x[sl], m, mu)
            area = wts * r ** 2
            cext = np.sum(area * qe)
            csca = np.sum(area * qs)
            p = (s11 * (wts / (x / r) ** 2)[:, None]).sum(axis=0)     # sum sigma_diff ~ S11 / k^2
            # normalise to int P dmu = 2  (i.e. (1/4pi) int P dOmega = 1)
            order = np.argsort(mu)
            norm = np.trapezoid(p[order], mu[order])
            p = p * 2.0 / norm
            pha[:, i] = p
            ssa[i] = min(1.0, csca / cext)
            asy[i] = 0.5 * np.trapezoid((p * mu)[order], mu[order])
        self.data = {
            'id': {'data': 'Mie', 'name': 'Mie', 'unit': 'N/A'},
            'wvl0': {'data': self.wvl0, 'name': 'Given wavelength', 'unit': 'nm'},
            'wvl': {'data': self.wvl0, 'name': 'Actual wavelength', 'unit': 'nm'},
            'ang': {'data': ang, 'name': 'Angle', 'unit': 'degree'},
            'pha': {'data': pha, 'name': 'Phase function', 'unit': 'N/A'},
            'ssa': {'data': ssa, 'name': 'Single scattering albedo', 'unit': 'N/A'},
            'asy': {'data': asy, 'name': 'Asymmetry parameter', 'unit': 'N/A'},
            'ref': {'data': reff, 'name': 'Effective radius', 'unit': 'mm'},
        }
