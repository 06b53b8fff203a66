"""
The Lorenz-Mie generator behind `pha_mie_wc` (er3t_b200/pre/pha.py) -- it stands in for libRadtran's wc.sol.mie.cdf
(er3t/pre/pha/pha_mie.py:98, absent offline) and feeds every config, so it is pinned against
  * the worked example of Bohren & Huffman (1983), appendix A (BHMIE): m = 1.55, x = 5.213 -> Qext = Qsca = 3.10543,
    Qback = 2.92534,
  * an independent evaluation of the Mie coefficients with scipy's spherical Bessel functions (no recurrences),
  * the large-sphere limit Qext -> 2, the optical theorem, and the normalisation / asymmetry of the water-cloud table.
"""

import numpy as np
from scipy import special

from er3t_b200.pre import pha as P


def mie_scipy(x, m, nmax=None):
    """Qext, Qsca, g from a_n, b_n written with spherical Bessel functions (Bohren & Huffman eq. 4.53, 4.61, 4.62)."""
    nmax = int(x + 4.05 * x ** (1.0 / 3.0) + 2) if nmax is None else nmax
    n = np.arange(1, nmax + 1)
    jx, jmx = special.spherical_jn(n, x), special.spherical_jn(n, m * x)
    yx = special.spherical_yn(n, x)
    djx, djmx = special.spherical_jn(n, x, derivative=True), special.spherical_jn(n, m * x, derivative=True)
    dyx = special.spherical_yn(n, x, derivative=True)
    hx, dhx = jx + 1j * yx, djx + 1j * dyx
    # Riccati-Bessel derivatives: [z j(z)]' = j + z j'
    dpsi_x, dpsi_mx, dxi_x = jx + x * djx, jmx + m * x * djmx, hx + x * dhx
    a = (m * m * jmx * dpsi_x - jx * dpsi_mx) / (m * m * jmx * dxi_x - hx * dpsi_mx)
    b = (jmx * dpsi_x - jx * dpsi_mx) / (jmx * dxi_x - hx * dpsi_mx)
    qext = 2.0 / x ** 2 * np.sum((2 * n + 1) * (a + b).real)
    qsca = 2.0 / x ** 2 * np.sum((2 * n + 1) * (np.abs(a) ** 2 + np.abs(b) ** 2))
    g = 4.0 / (x ** 2 * qsca) * (np.sum(n[:-1] * (n[:-1] + 2.0) / (n[:-1] + 1.0) * (a[:-1] * np.conj(a[1:]) + b[:-1] * np.conj(b[1:])).real)
                                 + np.sum((2 * n + 1.0) / (n * (n + 1.0)) * (a * np.conj(b)).real))
    return qext, qsca, g


def test_bohren_huffman_worked_example():
    x = 2.0 * np.pi * 0.525 / 0.6328
    s11, qe, qs = P._mie_s11(np.array([x]), complex(1.55, 0.0), np.array([-1.0, 1.0]))
    assert abs(x - 5.213) < 5e-4
    assert abs(qe[0] - 3.10543) < 2e-5 and abs(qs[0] - 3.10543) < 2e-5
    assert abs(4.0 * s11[0, 0] / x ** 2 - 2.92534) < 2e-5            # Qback = 4 |S(180 deg)|^2 / x^2


def test_against_independent_bessel_evaluation_and_g():
    mu = np.cos(np.deg2rad(np.linspace(0.0, 180.0, 7201)))
    for x, m in ((10.0, complex(1.33, 1e-6)), (40.0, complex(1.331, 1.64e-8)), (3.0, complex(1.5, 0.01))):
        s11, qe, qs = P._mie_s11(np.array([x]), m, mu)
        qe2, qs2, g2 = mie_scipy(x, m)
        assert abs(qe[0] / qe2 - 1.0) < 1e-9 and abs(qs[0] / qs2 - 1.0) < 1e-9
        # asymmetry parameter from the angular distribution itself
        p = s11[0]
        g = np.trapezoid(p * mu, mu) / np.trapezoid(p, mu)
        assert abs(g - g2) < 2e-4, (x, g, g2)
        # optical theorem / normalisation: int S11 dOmega = pi x^2 Qsca
        assert abs(2.0 * np.pi * abs(np.trapezoid(p, mu)) / (np.pi * x ** 2 * qs[0]) - 1.0) < 2e-4


def test_large_sphere_limit():
    x = np.array([500.0, 2000.0])
    with np.errstate(all='ignore'):                                   # terms beyond a size's own series length are masked
        _, qe, qs = P._mie_s11(x, complex(1.331, 1e-8), np.array([1.0]))
    assert np.all(qe > 2.0) and np.all(qe < 2.0 + 4.0 * x ** (-2.0 / 3.0))      # Qext -> 2 + 1.99 x^(-2/3) + ripple (van de Hulst)
    assert np.all(qs <= qe + 1e-12)


def test_water_cloud_table_10um_650nm():
    pha = P.pha_mie_wc(wavelength=650.0, reff=[10.0], nr=96)
    ang, p = pha.data['ang']['data'], pha.data['pha']['data'][:, 0]
    mu = np.cos(np.deg2rad(ang))
    assert ang.size == 498                                            # the reference's default grid (pha_mie.py:106-113)
    assert abs(0.5 * abs(np.trapezoid(p, mu)) - 1.0) < 1e-12          # (1/2) int P dmu = 1
    assert abs(pha.data['asy']['data'][0] - 0.862) < 0.008            # water droplets, r_eff = 10 um, 650 nm: g ~ 0.86
    assert 0.9999 < pha.data['ssa']['data'][0] <= 1.0
    assert p[0] > 1000.0 and p.min() > 0.0                            # diffraction peak; strictly positive
    i140 = np.argmin(np.abs(ang - 140.0)); i120 = np.argmin(np.abs(ang - 120.0))
    assert p[i140] > 1.5 * p[i120]                                    # primary rainbow near 140 deg
