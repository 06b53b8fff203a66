"""
Gas-absorption objects with the contract of er3t.pre.abs.abs_16g (er3t/pre/abs/abs_crk.py:27-706):
`.Ng`, `.wvl`, `.wvl_info`, `.coef['abso_coef'|'slit_func'|'solar'|'weight']['data']` where abso_coef is the layer
absorption OPTICAL DEPTH (Nz, Ng) (abs_crk.py:622-628; consumed by mca_atm_1d at er3t/rtm/mca/mca_atm.py:90).

The reference derives the coefficients from er3t/data/abs/abs_16g.h5 (absent) -- spectroscopy is out
of scope.  `abs_16g` here keeps the 16 hard-coded quadrature weights of the reference
(abs_crk.py:693-702) and fills abso_coef with a documented synthetic ladder; `abs_gen` wraps arbitrary arrays
(e.g. an O2 A-band g sweep for config 4).
"""

import numpy as np

__all__ = ['abs_16g', 'abs_gen']

# er3t/pre/abs/abs_crk.py:693-702
WEIGHT_16G = np.array([0.1527534276, 0.1491729617, 0.1420961469, 0.1316886544,
                       0.1181945205, 0.1019300893, 0.0832767040, 0.0626720116,
                       0.0424925000, 0.0046269894, 0.0038279891, 0.0030260086,
                       0.0022199750, 0.0014140010, 0.0005330000, 0.000075])


class abs_gen:

    """Generic container: abso_coef (Nz, Ng) layer absorption optical depth, weight (Ng), solar (Ng), slit (Nz, Ng)."""

    ID = 'Gas absorption (generic)'

    def __init__(self, wavelength, abso_coef, weight, solar=None, slit_func=None):
        abso_coef = np.asarray(abso_coef, dtype=np.float64)
        self.Nz, self.Ng = abso_coef.shape
        self.wvl = float(wavelength)
        self.nwl = 1
        self.wvl_info = '%.2f nm (applied SSFR slit)' % self.wvl
        weight = np.asarray(weight, dtype=np.float64)
        solar = np.full(self.Ng, 1.0) if solar is None else np.asarray(solar, dtype=np.float64)
        slit_func = np.ones((self.Nz, self.Ng)) if slit_func is None else np.asarray(slit_func, dtype=np.float64)
        self.coef = {
            'wvl': {'name': 'Wavelength', 'data': self.wvl, 'units': 'nm'},
            'abso_coef': {'name': 'Absorption Coefficient (Nz, Ng)', 'data': abso_coef},
            'slit_func': {'name': 'Slit Function (Nz, Ng)', 'data': slit_func},
            'solar': {'name': 'Solar Factor (Ng)', 'data': solar},
            'weight': {'name': 'Weight (Ng)', 'data': weight},
        }


class abs_16g(abs_gen):

    """
    16-g stand-in.  abso_coef[:, g] = tau_col(g) * (air column fraction of the layer) with a geometric ladder
    tau_col(g) = tau_min * (tau_max / tau_min) ** (g / 15): the strongest weights see the weakest absorption, like a
    sorted k-distribution.  solar = `solar_flux` (W m^-2 nm^-1) for every g.
    """

    ID = 'Gas absorption (16 g, synthetic ladder)'

    def __init__(self, wavelength=650.0, atm_obj=None, fname=None, overwrite=False, verbose=False,
                 tau_min=1.0e-3, tau_max=2.0, solar_flux=1.6):
        if atm_obj is None:
            raise OSError('Error [abs_16g]: please provide an \'atm\' object for <atm_obj>.')
        p = atm_obj.lev['pressure']['data']
        frac = (p[:-1] - p[1:]) / (p[0] - p[-1])
        g = np.arange(16)
        tau_col = tau_min * (tau_max / tau_min) ** (g / 15.0)
        abso = frac[:, None] * tau_col[None, :]
        super().__init__(wavelength, abso, WEIGHT_16G.copy(), solar=np.full(16, solar_flux))
