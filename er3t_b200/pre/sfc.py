"""
2-D surface container and Cox-Munk parameterisation with the contract of er3t.pre.sfc
(er3t/pre/sfc/sfc_gen.py:18-159, er3t/pre/sfc/util.py:14-150).

`sfc_2d_gen(sfc_2d=...)` accepts a 2-D albedo array (Lambertian), a dict with fiso/fgeo/fvol (LSRT; stored order
f_iso, f_geo, f_vol -- sfc_gen.py:123-125) or a dict with diffuse_alb/diffuse_frac/refrac_r/refrac_i/slope
(DSM / Cox-Munk; sfc_gen.py:137-141).
"""

import numpy as np

__all__ = ['sfc_2d_gen', 'cal_ocean_brdf']


class sfc_2d_gen:

    ID = 'Surface 2D (generic)'

    def __init__(self, sfc_2d=None, fname=None, overwrite=False, verbose=False, **kwargs):
        if sfc_2d is None:
            sfc_2d = kwargs.get('alb_2d', None)
        if sfc_2d is None:
            raise OSError('\nError [sfc_2d_gen]: Please provide <sfc_2d> to proceed.')
        self.sfc = sfc_2d
        self.verbose = verbose
        self.pre_sfc()

    def pre_sfc(self):
        self.data = {}
        if isinstance(self.sfc, np.ndarray):
            Nx, Ny = self.sfc.shape
            sfc = np.zeros((Nx, Ny, 1), dtype=np.float32)
            sfc[:, :, 0] = self.sfc
            name = 'Surface albedo (Lambertian)'
        elif isinstance(self.sfc, dict):
            keys = {key.lower().replace('_', ''): key for key in self.sfc.keys()}
            if all(k in keys for k in ('fiso', 'fvol', 'fgeo')):
                order = ('fiso', 'fgeo', 'fvol')
                name = 'Surface BRDF-LSRT (Isotropic, LiSparseR, RossThick)'
            elif all(k in keys for k in ('diffusealb', 'diffusefrac', 'refracr', 'refraci', 'slope')):
                order = ('diffusealb', 'diffusefrac', 'refracr', 'refraci', 'slope')
                name = 'Surface BRDF-DSM (Diffuse-Specular Mixture)'
            else:
                raise OSError('\nError [sfc_2d_gen]: Currently we only support 2D surface albedo or BRDF.')
            Nx, Ny = np.asarray(self.sfc[keys[order[0]]]).shape
            sfc = np.zeros((Nx, Ny, len(order)), dtype=np.float32)
            for i, k in enumerate(order):
                sfc[:, :, i] = self.sfc[keys[k]]
        else:
            raise OSError('\nError [sfc_2d_gen]: Currently we only support 2D numpy array or Python dictionary.')
        self.data['nx'] = {'data': Nx, 'name': 'Nx', 'units': 'N/A'}
        self.data['ny'] = {'data': Ny, 'name': 'Ny', 'units': 'N/A'}
        self.data['sfc'] = {'data': sfc, 'name': name, 'units': 'N/A'}
        self.Nx, self.Ny = Nx, Ny


# Hale & Querry (1973) refractive index of water; Koepke (1984) whitecap reflectance -- the published tables the
# reference embeds at er3t/pre/sfc/util.py:58-88,120-130
_HQ_WVL = np.array([0.250, 0.275, 0.300, 0.325, 0.345, 0.375, 0.400, 0.425, 0.445, 0.475, 0.500, 0.525, 0.550, 0.575,
                    0.600, 0.625, 0.650, 0.675, 0.700, 0.725, 0.750, 0.775, 0.800, 0.825, 0.850, 0.875, 0.900, 0.925,
                    0.950, 0.975, 1.000, 1.200, 1.400, 1.600, 1.800, 2.000, 2.200, 2.400, 2.600, 2.650, 2.700, 2.750,
                    2.800, 2.850, 2.900, 2.950, 3.000, 3.050, 3.100, 3.150, 3.200, 3.250, 3.300, 3.350, 3.400, 3.450,
                    3.500, 3.600, 3.700, 3.800, 3.900, 4.000]) * 1000.0
_HQ_REAL = np.array([1.362, 1.354, 1.349, 1.346, 1.343, 1.341, 1.339, 1.338, 1.337, 1.336, 1.335, 1.334, 1.333, 1.333,
                     1.332, 1.332, 1.331, 1.331, 1.331, 1.330, 1.330, 1.330, 1.329, 1.329, 1.329, 1.328, 1.328, 1.328,
                     1.327, 1.327, 1.327, 1.324, 1.321, 1.317, 1.312, 1.306, 1.296, 1.279, 1.242, 1.219, 1.188, 1.157,
                     1.142, 1.149, 1.201, 1.292, 1.371, 1.426, 1.467, 1.483, 1.478, 1.467, 1.450, 1.432, 1.420, 1.410,
                     1.400, 1.385, 1.374, 1.364, 1.357, 1.351])
_HQ_IMAG = np.array([3.35E-08, 2.35E-08, 1.60E-08, 1.08E-08, 6.50E-09, 3.50E-09, 1.86E-09, 1.30E-09, 1.02E-09, 9.35E-10,
                     1.00E-09, 1.32E-09, 1.96E-09, 3.60E-09, 1.09E-08, 1.39E-08, 1.64E-08, 2.23E-08, 3.35E-08, 9.15E-08,
                     1.56E-07, 1.48E-07, 1.25E-07, 1.82E-07, 2.93E-07, 3.91E-07, 4.86E-07, 1.06E-06, 2.93E-06, 3.48E-06,
                     2.89E-06, 9.89E-06, 1.38E-04, 8.55E-05, 1.15E-04, 1.10E-03, 2.89E-04, 9.56E-04, 3.17E-03, 6.70E-03,
                     1.90E-02, 5.90E-02, 1.15E-01, 1.85E-01, 2.68E-01, 2.98E-01, 2.72E-01, 2.40E-01, 1.92E-01, 1.35E-01,
                     9.24E-02, 6.10E-02, 3.68E-02, 2.61E-02, 1.95E-02, 1.32E-02, 9.40E-03, 5.15E-03, 3.60E-03, 3.40E-03,
                     3.80E-03, 4.60E-03])
_WC_WVL = np.arange(200.0, 4001.0, 100.0)
_WC_REF = np.array([0.220, 0.220, 0.220, 0.220, 0.220, 0.220, 0.215, 0.210, 0.200, 0.190, 0.175, 0.155, 0.130, 0.080,
                    0.100, 0.105, 0.100, 0.080, 0.045, 0.055, 0.065, 0.060, 0.055, 0.040, 0.000, 0.000, 0.000, 0.000,
                    0.000, 0.000, 0.000, 0.000, 0.000, 0.000, 0.000, 0.000, 0.000, 0.000, 0.000])


def cal_ocean_brdf(wvl=650.0, u10=1.0, sal=34.3, pcl=0.01, whitecaps=True):
    """
    Cox-Munk / whitecap parameters of the DSM surface (er3t/pre/sfc/util.py:14-150):
    refractive index of sea water (Hale & Querry 1973 + Friedman 1969 salinity correction 0.006 * sal / 34.3),
    slope variance 0.00512 * u10 + 0.003 (Cox & Munk 1954), whitecap fraction 2.95e-6 * u10 ** 3.52 with Koepke (1984)
    effective reflectance.  `u10` may be a scalar or a 2-D array (NumPy-2 clean, unlike the reference's np.float_).
    """
    u10 = np.asarray(u10, dtype=np.float64)
    wvl_ = np.zeros_like(u10) + wvl
    refrac_r = np.interp(wvl_, _HQ_WVL, _HQ_REAL) + 0.006 * (sal / 34.3)
    refrac_i = np.interp(wvl_, _HQ_WVL, _HQ_IMAG)
    slope = 0.00512 * u10 + 0.003
    if whitecaps:
        diffuse_frac = 2.95e-06 * (u10 ** 3.52)
        diffuse_alb = np.interp(wvl_, _WC_WVL, _WC_REF)
    else:
        diffuse_frac = 0.0 * u10
        diffuse_alb = 0.0 * u10
    out = {'diffuse_alb': diffuse_alb, 'diffuse_frac': diffuse_frac, 'refrac_r': refrac_r, 'refrac_i': refrac_i, 'slope': slope}
    if u10.ndim == 0:
        out = {k: float(v) for k, v in out.items()}
    return out
