"""
Atmosphere object with the contract of er3t.pre.atm.atm_atmmod (er3t/pre/atm/atm_atmmod.py:17-240):
`.lev` and `.lay` dictionaries of {'name', 'units', 'data'} holding altitude (km), pressure (hPa), temperature (K),
thickness (km) and gas number densities (cm^-3).

The reference interpolates the AFGL US-standard profile from er3t/data/atmmod/afglus.dat (absent).
This stand-in evaluates the 1976 US Standard Atmosphere analytically (temperature lapse-rate segments, hydrostatic
pressure) and fixed mixing ratios; the layer means follow the reference's convention (layer = mid-level values,
atm_atmmod.py:115-146).
"""

import numpy as np

__all__ = ['atm_atmmod']

_KB = 1.380649e-23


def _us76(z_km):
    """US Standard Atmosphere 1976 up to 86 km: temperature (K) and pressure (hPa) at geometric height z (km)."""
    zb = np.array([0.0, 11.0, 20.0, 32.0, 47.0, 51.0, 71.0, 86.0])
    lr = np.array([-6.5, 0.0, 1.0, 2.8, 0.0, -2.8, -2.0])
    g0, R = 9.80665, 287.053
    tb = [288.15]
    pb = [1013.25]
    for i in range(len(lr)):
        dzb = zb[i + 1] - zb[i]
        t1 = tb[-1] + lr[i] * dzb
        if lr[i] == 0.0:
            p1 = pb[-1] * np.exp(-g0 * dzb * 1000.0 / (R * tb[-1]))
        else:
            p1 = pb[-1] * (t1 / tb[-1]) ** (-g0 / (R * lr[i] / 1000.0))
        tb.append(t1)
        pb.append(p1)
    z = np.atleast_1d(np.asarray(z_km, dtype=np.float64))
    # geopotential height
    h = z * 6356.766 / (6356.766 + z)
    i = np.clip(np.searchsorted(zb, h, side='right') - 1, 0, len(lr) - 1)
    tbv, pbv, lrv, zbv = np.array(tb)[i], np.array(pb)[i], lr[i], zb[i]
    t = tbv + lrv * (h - zbv)
    with np.errstate(divide='ignore', invalid='ignore'):
        p_grad = pbv * (t / tbv) ** (-g0 / (R * np.where(lrv == 0, 1.0, lrv) / 1000.0))
    p_iso = pbv * np.exp(-g0 * (h - zbv) * 1000.0 / (R * tbv))
    p = np.where(lrv == 0.0, p_iso, p_grad)
    return t, p


class atm_atmmod:

    ID = 'Atmosphere 1D (US Standard 1976, analytic)'

    def __init__(self, levels=None, fname=None, fname_atmmod=None, overwrite=False, verbose=False, lat=30.0):
        if levels is None:
            raise OSError('Error [atm_atmmod]: Please provide <levels> (km) to proceed.')
        self.levels = np.asarray(levels, dtype=np.float64)
        self.layers = 0.5 * (self.levels[1:] + self.levels[:-1])
        self.verbose = verbose
        self.lat = lat
        self.lev = self._make(self.levels)
        self.lay = self._make(self.layers)
        self.lay['thickness'] = {'name': 'Thickness', 'units': 'km', 'data': self.levels[1:] - self.levels[:-1]}

    @staticmethod
    def _make(z):
        t, p = _us76(z)
        air = p * 100.0 / (_KB * t) * 1.0e-6                   # cm^-3
        h2o = air * 7.75e-3 * np.exp(-z / 2.0)                 # ~ US-standard column, 2 km scale height
        o3 = air * (3.0e-8 + 7.5e-6 * np.exp(-0.5 * ((z - 32.0) / 8.0) ** 2))
        d = {
            'altitude': {'name': 'Altitude', 'units': 'km', 'data': np.asarray(z, dtype=np.float64)},
            'pressure': {'name': 'Pressure', 'units': 'mb', 'data': p},
            'temperature': {'name': 'Temperature', 'units': 'K', 'data': t},
            'air': {'name': 'Air number density', 'units': 'cm-3', 'data': air},
            'o3': {'name': 'o3 number density', 'units': 'cm-3', 'data': o3},
            'o2': {'name': 'o2 number density', 'units': 'cm-3', 'data': air * 0.20948},
            'h2o': {'name': 'h2o number density', 'units': 'cm-3', 'data': h2o},
            'co2': {'name': 'co2 number density', 'units': 'cm-3', 'data': air * 4.0e-4},
            'no2': {'name': 'no2 number density', 'units': 'cm-3', 'data': air * 2.3e-11},
            'ch4': {'name': 'ch4 number density', 'units': 'cm-3', 'data': air * 1.8e-6},
        }
        return d
