"""
1-D and 3-D "atmosphere" adapters with the interface of er3t/rtm/mca/mca_atm.py.

Same namelist keys, units and array shapes as the reference (`.nml` is what `mcarats_ng` consumes), with two
differences that do not change the payload:
  * the 3-D binary is written only when `fname` is given, with one vectorised `ndarray.tofile` per array instead of
    `struct.pack(*every_voxel)` (er3t/rtm/mca/mca_atm.py:383-388, the dominant Python cost at config 2);
  * the debug prints of the reference (mca_atm.py:236-237,295-297) are dropped.
"""

import copy
import os
import warnings

import numpy as np

from er3t_b200.util import cal_mol_ext, get_lay_index, host_zeros

__all__ = ['mca_atm_1d', 'mca_atm_3d']


class _LazyArray:
    """A numpy array that is computed the first time somebody looks at it (np.asarray, indexing, .shape ...).
    `mca_atm_3d(device_props=True)` uses it for the fields the GPU derives itself, so that the namelist payload of the
    reference (er3t/rtm/mca/mca_atm.py:324-337) stays available without being paid for on every run."""

    def __init__(self, fn, shape, dtype=np.float32):
        self._fn, self._val, self.shape, self.dtype = fn, None, tuple(shape), np.dtype(dtype)
        self.ndim = len(self.shape)

    def get(self):
        if self._val is None:
            self._val = self._fn()
        return self._val

    def __array__(self, dtype=None, copy=None):
        a = self.get()
        return a if dtype is None else a.astype(dtype, copy=False)

    def __getitem__(self, key):
        return self.get()[key]

    def __setitem__(self, key, val):
        self.get()[key] = val

    def __getattr__(self, name):           # anything else behaves like the array
        return getattr(self.get(), name)


class mca_atm_1d:

    """
    atm_obj: atmosphere object (e.g. er3t_b200.pre.atm_atmmod or er3t.pre.atm.atm_atmmod)
    abs_obj: absorption object (e.g. er3t_b200.pre.abs_16g or er3t.pre.abs.abs_16g)

    self.nml[ig]: 'Atm_zgrd0' (m), 'Atm_wkd0', 'Atm_mtprof', 'Atm_tmp1d', 'Atm_nkd', 'Atm_np1d', 'Atm_nz',
                  'Atm_ext1d(1:, k)', 'Atm_omg1d(1:, k)', 'Atm_apf1d(1:, k)', 'Atm_abs1d(1:, 1)'
    """

    ID = 'MCARaTS 1D Atmosphere'

    def __init__(self, atm_obj=None, abs_obj=None):
        if atm_obj is None:
            raise OSError('Error [mca_atm_1d]: please provide an \'atm\' object for <atm_obj>.')
        if abs_obj is None:
            raise OSError('Error [mca_atm_1d]: please provide an \'abs\' object for <abs_obj>.')
        self.atm = atm_obj
        self.abs = abs_obj
        self.Ng = self.abs.Ng
        self.wvl_info = self.abs.wvl_info
        self.pre_mca_1d_atm()

    def pre_mca_1d_atm(self):
        # er3t/rtm/mca/mca_atm.py:68-102
        lev, lay = self.atm.lev, self.atm.lay
        dz_m = lay['thickness']['data'] * 1000.0
        nz = lay['altitude']['data'].size
        # Bodhaine Rayleigh optical depth per layer / thickness -> scattering coefficient (1/m)
        atm_sca = cal_mol_ext(self.abs.wvl * 0.001, lev['pressure']['data'][:-1], lev['pressure']['data'][1:], self.atm) / dz_m
        self.nml = {}
        for ig in range(self.Ng):
            d = {}
            d['Atm_zgrd0'] = {'data': lev['altitude']['data'] * 1000.0, 'units': 'm', 'name': 'Layer boundaries'}
            d['Atm_wkd0'] = {'data': 1.0, 'units': 'N/A', 'name': 'Weight coefficients'}
            d['Atm_mtprof'] = {'data': 0, 'units': 'N/A', 'name': 'Temperature profile flag'}
            d['Atm_tmp1d'] = {'data': lay['temperature']['data'], 'units': 'K', 'name': 'Temperature profile'}
            d['Atm_nkd'] = {'data': 1, 'units': 'N/A', 'name': 'Number of K-distribution'}
            d['Atm_np1d'] = {'data': 1, 'units': 'N/A', 'name': 'Number of 1D atmospheric constituents'}
            d['Atm_nz'] = {'data': nz, 'units': 'N/A', 'name': 'Number of z grid points'}
            d['Atm_abs1d(1:, 1)'] = {'data': self.abs.coef['abso_coef']['data'][:, ig] / dz_m, 'units': '/m', 'name': 'Absorption coefficients'}
            d['Atm_ext1d(1:, 1)'] = {'data': atm_sca, 'units': '/m', 'name': 'Extinction coefficients'}
            d['Atm_omg1d(1:, 1)'] = {'data': np.repeat(1.0, nz), 'units': 'N/A', 'name': 'Single scattering albedo'}
            d['Atm_apf1d(1:, 1)'] = {'data': np.repeat(-1, nz), 'units': 'N/A', 'name': 'Phase function'}
            self.nml[ig] = d

    def add_mca_1d_atm(self, ext1d=None, omg1d=None, apf1d=None, z_bottom=None, z_top=None):
        # er3t/rtm/mca/mca_atm.py:105-139 (z_bottom / z_top in km, compared with layer-centre altitudes)
        if (ext1d is None) or (omg1d is None) or (apf1d is None):
            raise OSError('Error [mca_atm_1d]: Please provide values of <ext1d>, <omg1d>, and <apf1d>.')
        alt = self.atm.lay['altitude']['data']
        nz = alt.size
        keep = np.ones(nz, dtype=bool)
        if z_bottom is not None:
            keep &= ~(alt < z_bottom)
        if z_top is not None:
            keep &= ~(alt > z_top)
        for ig in range(self.Ng):
            k = self.nml[ig]['Atm_np1d']['data'] + 1
            for key, val, units, name in (('Atm_ext1d', ext1d, '/m', 'Extinction coefficients'),
                                          ('Atm_omg1d', omg1d, 'N/A', 'Single scattering albedo'),
                                          ('Atm_apf1d', apf1d, 'N/A', 'Phase function')):
                prof = np.zeros(nz)
                prof[:] = val
                prof[~keep] = 0.0
                self.nml[ig]['%s(1:, %d)' % (key, k)] = {'data': prof, 'units': units, 'name': name}
            self.nml[ig]['Atm_np1d']['data'] += 1


class mca_atm_3d:

    """
    atm_obj: atmosphere object; cld_obj: cloud object with `.lay` (nx, ny, dx, dy, altitude, thickness, extinction,
    cer, temperature); pha_obj: phase-function object or None.

    self.nml: 'Atm_nx', 'Atm_ny', 'Atm_dx' (m), 'Atm_dy' (m), 'Atm_nz3', 'Atm_iz3l', 'Atm_np3d',
              'Atm_tmpa3d' (nx, ny, nz3), 'Atm_abst3d' / 'Atm_extp3d' / 'Atm_omgp3d' / 'Atm_apfp3d' (nx, ny, nz3, np3d)
    """

    ID = 'MCARaTS 3D Atmosphere'

    def __init__(self, atm_obj=None, cld_obj=None, pha_obj=None, fname=None, overwrite=True, force=False,
                 verbose=False, quiet=False, *, device_props=False):
        """device_props=True (new, keyword only): with a Mie `pha_obj`, single-scattering albedo and asymmetry parameter
        of the cloudy voxels are NOT interpolated on the host (mca_atm.py:291-303, 0.1-10 s at config 2); the effective
        radius field and the (ref, ssa, asy) tables are handed to the solver, whose scene-packing kernel derives the
        same float32 values on the GPU (include/b200rt.h, scene.cer3d).  `nml['Atm_omgp3d']['data']` etc. remain
        available as lazily evaluated arrays."""
        self.device_props = bool(device_props)
        self.overwrite = overwrite
        self.verbose = verbose
        self.quiet = quiet
        if atm_obj is None:
            raise OSError('Error [mca_atm_3d]: Please provide an \'atm\' object for <atm_obj>.')
        if cld_obj is None:
            raise OSError('Error [mca_atm_3d]: Please provide an \'cld\' object for <cld_obj>.')
        self.atm = atm_obj
        self.cld = cld_obj
        if pha_obj is None and self.verbose:
            warnings.warn('Warning [mca_atm_3d]: No phase function set specified - ignore thermodynamic phase/effective radius with g=0.85 (Henyey-Greenstein).')
        self.pha = pha_obj
        if self.cld.lay['altitude']['data'].size != self.cld.lay['thickness']['data'].size:
            msg = 'Error [mca_atm_3d]: Incorrect number of cloud layers (%d) vs layer thicknesses (%d).' % \
                  (self.cld.lay['altitude']['data'].size, self.cld.lay['thickness']['data'].size)
            raise ValueError(msg)
        self.pre_mca_3d_atm()
        # the reference always writes 'mca_atm_3d.bin' into the CWD when fname is None; here the tensors are handed to
        # the solver in memory and the file is an optional artefact
        if fname is not None:
            if self.overwrite or ((not os.path.exists(fname)) and (not force)):
                self.gen_mca_3d_atm_file(fname)
            else:
                self.nml['Atm_inpfile'] = {'data': fname}
        else:
            self.nml['Atm_inpfile'] = {'data': 'mca_atm_3d.bin'}

    def pre_mca_3d_atm(self):
        # er3t/rtm/mca/mca_atm.py:231-337
        cld, atm = self.cld.lay, self.atm.lay
        lay_index = get_lay_index(cld['altitude']['data'], atm['altitude']['data'])
        nx, ny = int(cld['nx']['data']), int(cld['ny']['data'])
        nz3 = int(lay_index.size)
        iz3l = int(lay_index[0] + 1)
        if (iz3l + nz3) > atm['altitude']['data'].size:
            raise ValueError('Error [mca_atm_3d]: Non-homogeneous layer top exceeds atmosphere top.')

        ext_in = cld['extinction']['data']
        ext_in = ext_in.data if isinstance(ext_in, np.ma.MaskedArray) else np.asarray(ext_in)
        pid = None if self.pha is None else self.pha.data['id']['data'].lower()
        self.cer3d = None
        self.cer_tables = None
        if self.device_props and pid == 'mie':
            self._pre_device(cld, atm, lay_index, ext_in, nx, ny, nz3, iz3l)
            return
        atm_tmp = np.asarray(cld['temperature']['data'], dtype=np.float32) - atm['temperature']['data'][lay_index].astype(np.float32)[None, None, :]
        atm_abs = np.zeros((nx, ny, nz3, 1), dtype=np.float32)
        # the three fields the solver reads live in page-locked memory when a GPU is present (fast H2D)
        atm_ext = host_zeros((nx, ny, nz3, 1), np.float32)
        atm_ext[..., 0] = ext_in[:, :, :nz3]
        atm_omg = host_zeros((nx, ny, nz3, 1), np.float32)
        atm_omg[...] = 1.0
        atm_apf = host_zeros((nx, ny, nz3, 1), np.float32)

        if self.pha is None:
            atm_apf[...] = 0.85
        else:
            # Rayleigh everywhere, then the cloudy voxels
            atm_apf[...] = -1.0
            logic_cld = ext_in[:, :, :nz3] > 0.0
            pid = self.pha.data['id']['data'].lower()
            if pid == 'hg':
                atm_apf[logic_cld, 0] = np.argmin(np.abs(self.pha.data['asy']['data'] - 0.85)) + 1.0
            elif pid == 'mie':
                cer = cld['cer']['data']
                cer = cer.data if isinstance(cer, np.ma.MaskedArray) else np.asarray(cer)
                ref = np.asarray(self.pha.data['ref']['data'], dtype=np.float64)
                c = cer[:, :, :nz3][logic_cld].astype(np.float64)
                # HG with the Mie-derived asymmetry parameter: the reference interpolates ssa and asy (not the table
                # index) linearly in effective radius with linear extrapolation (mca_atm.py:291-303)
                atm_omg[logic_cld, 0] = _interp_extrap(c, ref, self.pha.data['ssa']['data'])
                atm_apf[logic_cld, 0] = _interp_extrap(c, ref, self.pha.data['asy']['data'])

        self._fill_nml(cld, nz3, iz3l, atm_tmp, atm_abs, atm_ext, atm_omg, atm_apf)

    def _pre_device(self, cld, atm, lay_index, ext_in, nx, ny, nz3, iz3l):
        """device_props: only extinction and effective radius are touched on the host (views of the cloud object's own
        arrays when those are float32 and C-ordered, else one float32 copy each into page-locked memory); omega / apf /
        temperature deviation are lazy."""
        cer = cld['cer']['data']
        cer = cer.data if isinstance(cer, np.ma.MaskedArray) else np.asarray(cer)

        def as_field(a):
            # ZERO-COPY when the cloud object already holds float32 C-ordered (nx, ny, nz3) data: the solver then reads
            # the caller's buffer directly (at full PCIe speed when that buffer is page-locked, e.g.
            # er3t_b200.util.host_zeros / pin_array); anything else is converted once into page-locked memory
            v = a[:, :, :nz3]
            if v.dtype == np.float32 and v.flags['C_CONTIGUOUS'] and v.shape == (nx, ny, nz3):
                return v.reshape(nx, ny, nz3, 1)
            out = host_zeros((nx, ny, nz3, 1), np.float32)
            out[..., 0] = v
            return out
        atm_ext = as_field(ext_in)
        cer3 = as_field(cer)
        self.cer3d = cer3
        ref = np.ascontiguousarray(self.pha.data['ref']['data'], dtype=np.float64)
        ssa = np.ascontiguousarray(self.pha.data['ssa']['data'], dtype=np.float64)
        asy = np.ascontiguousarray(self.pha.data['asy']['data'], dtype=np.float64)
        self.cer_tables = (ref, ssa, asy)

        def props(which):
            def f():
                out = np.full((nx, ny, nz3, 1), 1.0 if which == 0 else -1.0, dtype=np.float32)
                m = atm_ext[..., 0] > 0.0
                # the float32 effective radius the GPU sees, interpolated in float64 like mca_atm.py:291-303
                out[m, 0] = _interp_extrap(cer3[..., 0][m].astype(np.float64), ref, ssa if which == 0 else asy)
                return out
            return f

        def tmpa():
            return np.asarray(cld['temperature']['data'], dtype=np.float32) - atm['temperature']['data'][lay_index].astype(np.float32)[None, None, :]

        shp = (nx, ny, nz3, 1)
        self._fill_nml(cld, nz3, iz3l, _LazyArray(tmpa, (nx, ny, nz3)), _LazyArray(lambda: np.zeros(shp, dtype=np.float32), shp), atm_ext,
                       _LazyArray(props(0), shp), _LazyArray(props(1), shp))

    def _fill_nml(self, cld, nz3, iz3l, atm_tmp, atm_abs, atm_ext, atm_omg, atm_apf):
        self.nml = {}
        self.nml['Atm_nx'] = copy.deepcopy(cld['nx'])
        self.nml['Atm_ny'] = copy.deepcopy(cld['ny'])
        self.nml['Atm_dx'] = {'data': cld['dx']['data'] * 1000.0, 'name': cld['dx'].get('name', 'dx'), 'units': 'm'}
        self.nml['Atm_dy'] = {'data': cld['dy']['data'] * 1000.0, 'name': cld['dy'].get('name', 'dy'), 'units': 'm'}
        self.nml['Atm_nz3'] = {'data': nz3, 'unit': 'N/A', 'name': 'number of 3D layer'}
        # NOTE the reference emits lay_index[0] + 2 here (iz3l = lay_index[0] + 1 at mca_atm.py:242, then iz3l + 1 at
        # :330), one layer above the documented meaning of Atm_iz3l.  Reproduced verbatim so
        # that results match "MCARaTS as driven by er3t"; mcarats_ng(..., iz3l_fix=True) undoes it.
        self.nml['Atm_iz3l'] = {'data': iz3l + 1, 'unit': 'N/A', 'name': 'layer index of first 3D layer'}
        self.nml['Atm_tmpa3d'] = {'data': atm_tmp, 'units': 'K', 'name': 'Temperature deviation'}
        self.nml['Atm_abst3d'] = {'data': atm_abs, 'units': '/m', 'name': 'Absorption coefficients deviation'}
        self.nml['Atm_extp3d'] = {'data': atm_ext, 'units': '/m', 'name': 'Extinction coefficients'}
        self.nml['Atm_omgp3d'] = {'data': atm_omg, 'units': 'N/A', 'name': 'Single scattering Albedo'}
        self.nml['Atm_apfp3d'] = {'data': atm_apf, 'units': 'N/A', 'name': 'Phase function'}
        self.nml['Atm_np3d'] = {'data': 1, 'units': 'N/A', 'name': 'Number of 3D atmospheric constituents'}

    def add_mca_3d_atm(self, ext3d=None, omg3d=None, apf3d=None):
        # er3t/rtm/mca/mca_atm.py:340-370
        if (ext3d is None) or (omg3d is None) or (apf3d is None):
            raise OSError('Error [mca_atm_3d]: Please provide an <ext3d>, <omg3d>, and <apf3d>.')
        for name, arr in (('ext3d', ext3d), ('omg3d', omg3d), ('apf3d', apf3d)):
            if isinstance(arr, np.ndarray) and arr.ndim != 3:
                raise ValueError('Error [mca_atm_3d]: <%s> should be in the dimension of (nx, ny, nz).' % name)
        for key, arr in (('Atm_extp3d', ext3d), ('Atm_omgp3d', omg3d), ('Atm_apfp3d', apf3d)):
            new = np.zeros(self.nml[key]['data'].shape[:3] + (1,), dtype=np.float32)
            new[..., 0] = arr
            self.nml[key]['data'] = np.concatenate((self.nml[key]['data'], new), axis=-1)
        self.nml['Atm_np3d']['data'] += 1

    def gen_mca_3d_atm_file(self, fname):
        # layout of er3t/rtm/mca/mca_atm.py:373-392: tmpa3d, abst3d, then per component ext, omg, apf;
        # little-endian float32, x fastest
        if not self.quiet:
            print('Message [mca_atm_3d]: Creating 3D atm file <%s> for MCARaTS ...' % fname)
        fname = os.path.abspath(fname)
        self.nml['Atm_inpfile'] = {'data': fname}
        with open(fname, 'wb') as f:
            _write_f(f, self.nml['Atm_tmpa3d']['data'])
            _write_f(f, self.nml['Atm_abst3d']['data'])
            for i in range(self.nml['Atm_np3d']['data']):
                _write_f(f, self.nml['Atm_extp3d']['data'][..., i])
                _write_f(f, self.nml['Atm_omgp3d']['data'][..., i])
                _write_f(f, self.nml['Atm_apfp3d']['data'][..., i])
        if not self.quiet:
            print('Message [mca_atm_3d]: File <%s> is created.' % fname)


def _write_f(f, arr):
    np.asarray(arr).astype('<f4').flatten(order='F').tofile(f)


def _interp_extrap(x, xp, fp):
    """linear interpolation with linear extrapolation (scipy interp1d(..., fill_value='extrapolate'))"""
    xp = np.asarray(xp, dtype=np.float64)
    fp = np.asarray(fp, dtype=np.float64)
    i = np.clip(np.searchsorted(xp, x, side='right') - 1, 0, xp.size - 2)
    t = (x - xp[i]) / (xp[i + 1] - xp[i])
    return fp[i] + t * (fp[i + 1] - fp[i])
