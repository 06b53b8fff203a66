"""
Helper functions the photon-transport path needs from er3t.util: only the ones on
the hot path are restated here, with the reference line they follow.
"""

import datetime

import numpy as np

__all__ = ['cal_sol_fac', 'cal_mol_ext', 'cal_mol_ext_0', 'g0_calc', 'g_alt_calc', 'get_lay_index', 'nice_array_str',
           'cal_r_twostream', 'cal_t_twostream', 'cal_ext', 'add_reference', 'print_reference', 'host_zeros', 'pin_array']

_references = []


def add_reference(reference):
    """Collect literature references like er3t.util.add_reference (used by mcarats_ng, er3t/rtm/mca/mcarats.py:101)."""
    if reference not in _references:
        _references.append(reference)


def print_reference():
    for r in _references:
        print(r)


def cal_sol_fac(dtime):
    """Sun-Earth distance factor 1/r^2 for the day of year of `dtime` (er3t/util/util.py:934-950)."""
    doy = dtime.timetuple().tm_yday
    rsun = 1.0 - 0.0167086 * np.cos(0.017202124161707175 * (doy - 4.0))
    return 1.0 / (rsun * rsun)


def g0_calc(lat):
    """Sea-level gravity (m/s^2), Bodhaine et al. 1999 eq. 11 (er3t/util/util.py:1005-1012)."""
    c2 = np.cos(2.0 * np.deg2rad(lat))
    return 9.806160 * (1.0 - 0.0026373 * c2 + 0.0000059 * c2 * c2)


def g_alt_calc(g0, lat, z):
    """Gravity at height z (m), Bodhaine et al. 1999 eq. 10 (er3t/util/util.py:1014-1028); g0 in m/s^2."""
    c2 = np.cos(2.0 * np.deg2rad(lat))
    g = g0 * 100.0 - (3.085462e-4 + 2.27e-7 * c2) * z + (7.254e-11 + 1.0e-13 * c2) * z ** 2 - (1.517e-17 + 6.0e-20 * c2) * z ** 3
    return g / 100.0


def _bodhaine_ratio(wv0):
    num = 1.0455996 - 341.29061 * wv0 ** (-2.0) - 0.90230850 * wv0 ** 2.0
    den = 1.0 + 0.0027059889 * wv0 ** (-2.0) - 85.968563 * wv0 ** 2.0
    return num / den


def cal_mol_ext_0(wv0, pz1, pz2):
    """Rayleigh optical depth between pressures pz1 > pz2 (hPa) at wavelength wv0 (micron), fixed surface constant
    0.00210966 (er3t/util/util.py:1080-1100)."""
    return 0.00210966 * _bodhaine_ratio(wv0) * (pz1 - pz2) / 1013.25


def cal_mol_ext(wv0, pz1, pz2, atm0=None):
    """
    Rayleigh optical depth per layer as the reference computes it today (er3t/util/util.py:1030-1077): the Bodhaine
    wavelength ratio times a surface constant p_sfc * N_A / (g0 * m_a) * 1e-28 that depends on the atmosphere object
    (surface pressure, CO2 mixing ratio, latitude 30 deg unless `atm0.lat` exists).  Without an atmosphere object the
    constant of `cal_mol_ext_0` is used.
    """
    if atm0 is None:
        return cal_mol_ext_0(wv0, pz1, pz2)
    lat = getattr(atm0, 'lat', 30.0)
    g0 = g0_calc(lat) * 100.0                                   # cm/s^2
    ma = 28.9595 + 15.0556 * atm0.lay['co2']['data'][0] / atm0.lay['air']['data'][0]
    p_sfc = atm0.lev['pressure']['data'][0] * 1000.0            # dyne/cm^2
    const_sfc = p_sfc * 6.02214179e23 / (g0 * ma) * 1e-28
    return const_sfc * _bodhaine_ratio(wv0) * (pz1 - pz2) / 1013.25


def get_lay_index(lay, lay_ref):
    """Index of every altitude in `lay` inside `lay_ref` (nearest), error if farther than half the largest spacing
    (er3t/util/util.py:804-831)."""
    lay = np.atleast_1d(np.asarray(lay, dtype=np.float64))
    lay_ref = np.asarray(lay_ref, dtype=np.float64)
    threshold = (lay_ref[1:] - lay_ref[:-1]).max() / 2.0
    index = np.argmin(np.abs(lay[:, None] - lay_ref[None, :]), axis=1)
    dd = np.abs(lay - lay_ref[index])
    if np.any(dd > threshold):
        raise ValueError('Error [get_layer_index]: Mismatch between layer and reference layer: ' + str(dd.max()))
    return index.astype(np.int32)


def nice_array_str(array1d, numPerLine=6):
    """1-D array -> lines of `numPerLine` values formatted '  %12g' (er3t/util/util.py:191-221)."""
    array1d = np.asarray(array1d)
    if array1d.ndim > 1:
        raise ValueError('Error [nice_array_str]: Only support 1-D array.')
    lines = []
    for i in range(0, array1d.size, numPerLine):
        lines.append(''.join('  %12g' % v for v in array1d[i:i + numPerLine]) + '\n')
    return ''.join(lines)


def cal_r_twostream(tau, a=0.0, g=0.85, mu=1.0):
    """Two-stream reflectance without absorption (er3t/util/util.py:1135-1151)."""
    x = 2.0 * mu / (1.0 - g) / (1.0 - a)
    return (tau + a * x) / (tau + x)


def cal_t_twostream(tau, a=0.0, g=0.85, mu=1.0):
    """Two-stream transmittance without absorption (er3t/util/util.py:1155-1171)."""
    x = 2.0 * mu / (1.0 - g) / (1.0 - a)
    return x / (tau + x)


def cal_ext(cot, cer, dz=1.0, Qe=2.0):
    """Extinction (1/m) from optical thickness and effective radius (um) over thickness dz (km)
    (er3t/util/util.py:1103-1131)."""
    lwp = 2.0 / 3000.0 * cot * cer
    lwc = lwp / dz
    return 0.75 * Qe * lwc / cer * 1.0e3


def host_zeros(shape, dtype=np.float32):
    """Zero-filled host array; page-locked (pinned) when a CUDA device is present so that the host -> device copy of the
    3-D fields in b200rt_upload_scene runs at full PCIe speed.  Falls back to plain numpy memory otherwise."""
    try:
        import torch
        if torch.cuda.is_available():
            tdt = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64, np.dtype(np.int32): torch.int32}[np.dtype(dtype)]
            return torch.zeros(tuple(shape), dtype=tdt, pin_memory=True).numpy()
    except Exception:
        pass
    return np.zeros(shape, dtype=dtype)


def pin_array(a):
    """Copy of `a` in page-locked host memory (same dtype, C order) when a CUDA device is present, else `a` made
    C-contiguous.  For inputs that are uploaded repeatedly (cloud fields of a scene that is re-run): cudaMemcpy from
    page-locked memory runs at PCIe speed, from pageable memory at roughly a third of it."""
    a = np.ascontiguousarray(a)
    try:
        import torch
        if torch.cuda.is_available():
            t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0].copy()).dtype, pin_memory=True)
            out = t.numpy()
            out[...] = a
            return out
    except Exception:
        pass
    return a


def default_date():
    return datetime.datetime.now()
