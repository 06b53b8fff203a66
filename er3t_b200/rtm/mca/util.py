"""
Callers of the hot path that er3t/rtm/mca/util.py provides: reflectance-vs-COT look-up tables.

The reference builds the table by looping over the COT values and running one complete `mcarats_ng` job set
(Nrun x Ng MCARaTS processes) per value (er3t/rtm/mca/util.py:105-195,311-398), i.e. 35 x 48 tiny subprocess jobs for a
typical table.  Here ALL COT values are traced in ONE launch: the plane-parallel clouds are laid
side by side as the columns of one scene and the solver runs in IPA mode, where a photon never leaves the column it
entered -- every column is an independent plane-parallel problem, exactly what the reference computes value by value.
The source illuminates the columns uniformly, so each COT receives `Nphoton` photons on average.

Same constructor arguments, attributes (`cot ref ref_std rad rad_std ref_2s toa0 mu0`) and methods
(`get_cot_from_ref`, `get_ref_from_cot`, `run_all`, `load_all`) as the reference; per-COT result files
`<fdir>/<output_tag>_cot-XXXXX.X_cer-XX.X.h5` with `mean/rad`, `mean/rad_std`, `mean/toa` are written like the
reference's (`.h5` through h5py when it is installed, else the same keys in a zip/npz container under the same name).
"""

import datetime
import os

import numpy as np
from scipy.interpolate import interp1d

import er3t_b200.pre as _pre
from er3t_b200.util import cal_r_twostream, default_date
from .mca_atm import mca_atm_1d, mca_atm_3d
from .mca_out import mca_out_ng
from .mca_sca import mca_sca
from .mcarats import mcarats_ng

__all__ = ['func_ref_vs_cot', 'func_ref_vs_cot_multi_pixel']


def _name_tag(cot0, cer0):
    return 'cot-%05.1f_cer-%04.1f' % (cot0, cer0)


def _dump_one(fname, rad, rad_std, toa):
    items = {'rad': np.asarray(rad), 'rad_std': np.asarray(rad_std), 'toa': np.asarray(toa)}
    try:
        import h5py
    except ImportError:
        h5py = None
    if h5py is not None:
        with h5py.File(fname, 'w') as f:
            g = f.create_group('mean')
            for k, v in items.items():
                g[k] = v
    else:
        with open(fname, 'wb') as f:
            np.savez_compressed(f, **{'mean/%s' % k: v for k, v in items.items()})


def _load_one(fname):
    with open(fname, 'rb') as f:
        magic = f.read(2)
    if magic == b'PK':
        z = np.load(fname, allow_pickle=False)
        return z['mean/rad'], z['mean/rad_std'], z['mean/toa']
    import h5py
    with h5py.File(fname, 'r') as f:
        return f['mean/rad'][...], f['mean/rad_std'][...], f['mean/toa'][...]


class _lut_base:

    """Shared machinery: one batched IPA launch over all COT values, then the reference's post-processing."""

    tabulated = False     # True: cloud scatters with the tabulated Mie function of the nearest r_eff (1-D variant)

    def _setup(self, cot, cer0, fdir, date, wavelength, surface_albedo, solar_zenith_angle, solar_azimuth_angle,
               sensor_zenith_angle, sensor_azimuth_angle, sensor_altitude, cloud_top_height, cloud_geometrical_thickness,
               solver, Nphoton, atm0, Ncpu, output_tag, overwrite, Nx, Ny, dx, dy, seed, device, solver_obj, Nrun, pha0, abs0):
        self.cot = np.asarray(cot, dtype=np.float64)
        self.cer0 = cer0
        self.wvl0 = wavelength
        self.sza0 = solar_zenith_angle
        self.saa0 = solar_azimuth_angle
        self.vza0 = sensor_zenith_angle
        self.vaa0 = sensor_azimuth_angle
        self.alt0 = sensor_altitude
        self.cth0 = cloud_top_height
        self.cbh0 = cloud_top_height - cloud_geometrical_thickness
        self.alb0 = surface_albedo
        self.fdir = fdir
        self.output_tag = output_tag
        self.photon0 = Nphoton
        self.solver0 = solver
        self.cpu0 = Ncpu
        self.date0 = date
        self.atm0 = atm0
        self.Nx, self.Ny, self.dx, self.dy = Nx, Ny, dx, dy
        self.seed, self.device, self._solver_obj, self.Nrun = seed, device, solver_obj, Nrun
        self._pha0, self._abs0 = pha0, abs0
        self.mca = None

        self.mu0 = np.cos(np.deg2rad(self.sza0))
        self.ref_2s = cal_r_twostream(self.cot, a=self.alb0, mu=self.mu0)

        if not overwrite:
            try:
                self.load_all()
            except Exception:
                self.run_all()
                self.load_all()
        else:
            self.run_all()
            self.load_all()

    # ------------------------------------------------------------------ results (er3t/rtm/mca/util.py:73-101)
    def load_all(self):
        rad, rad_std, toa0 = [], [], None
        for i in range(self.cot.size):
            if self.fdir is not None:
                r, s, toa0 = _load_one('%s/%s_%s.h5' % (self.fdir, self.output_tag, _name_tag(self.cot[i], self.cer0)))
            else:
                r, s, toa0 = self._mem[i]
            rad.append(np.mean(r))
            rad_std.append(np.mean(s))
        self.rad = np.array(rad)
        self.rad_std = np.array(rad_std)
        self.toa0 = toa0
        self.ref = np.pi * self.rad / (toa0 * self.mu0)
        self.ref_std = np.pi * self.rad_std / (toa0 * self.mu0)

    # ------------------------------------------------------------------ one launch for the whole table
    def run_all(self):
        if self.fdir is not None:
            os.makedirs(self.fdir, exist_ok=True)
        atm0 = self.atm0
        if atm0 is None:
            atm0 = _pre.atm_atmmod(levels=np.arange(0.0, 20.1, 0.1))
        abs0 = self._abs0 if self._abs0 is not None else _pre.abs_16g(wavelength=self.wvl0, atm_obj=atm0)
        pha0 = self._pha0 if self._pha0 is not None else _pre.pha_mie_wc(wavelength=self.wvl0)
        sca0 = mca_sca(pha_obj=pha0)

        alt = atm0.lay['altitude']['data']
        altitude0 = alt[(alt >= self.cbh0) & (alt <= self.cth0)]
        if altitude0.size == 0:
            raise OSError('Error [func_ref_vs_cot]: no atmospheric layer between cloud base and cloud top.')
        ncot = self.cot.size
        nxb = ncot * self.Nx
        cld0 = _pre.cld_gen_hom(cot0=1.0, cer0=self.cer0, altitude=altitude0, atm_obj=atm0, Nx=nxb, Ny=self.Ny, dx=self.dx, dy=self.dy)
        # column block i carries COT i, split evenly over the cloud layers like cld_gen_hom (pre/cld/cld_gen.py:659-696);
        # thickness from the atmosphere's own layers so that the column optical depth is exact on a non-uniform grid
        thick = atm0.lay['thickness']['data'][(alt >= self.cbh0) & (alt <= self.cth0)] * 1000.0
        ext = np.repeat(self.cot, self.Nx)[:, None, None] / altitude0.size / thick[None, None, :]
        cld0.lay['extinction']['data'] = np.broadcast_to(ext, (nxb, self.Ny, altitude0.size)).astype(np.float32).copy()
        cld0.lay['cot']['data'] = cld0.lay['extinction']['data'] * thick[None, None, :]

        atm1d0 = mca_atm_1d(atm_obj=atm0, abs_obj=abs0)
        atm3d0 = mca_atm_3d(cld_obj=cld0, atm_obj=atm0, pha_obj=pha0, quiet=True)
        if self.tabulated:
            # er3t/rtm/mca/util.py:144-149: table of the nearest effective radius, its single-scattering albedo
            iref = int(np.argmin(np.abs(pha0.data['ref']['data'] - self.cer0)))
            cloudy = atm3d0.nml['Atm_extp3d']['data'][..., 0] > 0.0
            atm3d0.nml['Atm_omgp3d']['data'][cloudy, 0] = pha0.data['ssa']['data'][iref]
            atm3d0.nml['Atm_apfp3d']['data'][cloudy, 0] = iref + 1

        fdir_run = 'tmp-data/lut' if self.fdir is None else '%s/%s_batched/rad' % (self.fdir, self.output_tag)
        self.mca = mcarats_ng(
            date=self.date0, atm_1ds=[atm1d0], atm_3ds=[atm3d0], sca=sca0, target='radiance',
            surface_albedo=self.alb0, solar_zenith_angle=self.sza0, solar_azimuth_angle=self.saa0,
            sensor_zenith_angle=self.vza0, sensor_azimuth_angle=self.vaa0, sensor_altitude=self.alt0,
            fdir=fdir_run, Nrun=self.Nrun, Ng=abs0.Ng, weights=abs0.coef['weight']['data'],
            photons=self.photon0 * ncot, solver='IPA', Ncpu=self.cpu0, mp_mode='py', overwrite=True, quiet=True,
            seed=self.seed, device=self.device, solver_obj=self._solver_obj, iz3l_fix=True)
        out0 = mca_out_ng(mca_obj=self.mca, abs_obj=abs0, mode='mean', squeeze=False, quiet=True)
        rad = np.asarray(out0.data['rad']['data']).reshape(nxb, self.Ny)
        rad_std = np.asarray(out0.data['rad_std']['data']).reshape(nxb, self.Ny)
        toa = out0.data['toa']['data']
        self._mem = []
        for i in range(ncot):
            r = rad[i * self.Nx:(i + 1) * self.Nx]
            s = rad_std[i * self.Nx:(i + 1) * self.Nx]
            self._mem.append((r, s, toa))
            if self.fdir is not None:
                _dump_one('%s/%s_%s.h5' % (self.fdir, self.output_tag, _name_tag(self.cot[i], self.cer0)), r, s, toa)

    # ------------------------------------------------------------------ table look-ups (er3t/rtm/mca/util.py:197-213)
    def get_cot_from_ref(self, ref, method='cubic', mode='rt'):
        if mode == '2s':
            f = interp1d(self.ref_2s, self.cot, kind=method, bounds_error=False, fill_value='extrapolate')
        elif mode == 'rt':
            f = interp1d(self.ref, self.cot, kind=method, bounds_error=False, fill_value='extrapolate')
        return f(ref)

    def get_ref_from_cot(self, cot, method='cubic', mode='rt'):
        if mode == '2s':
            f = interp1d(self.cot, self.ref_2s, kind=method, bounds_error=False)
        elif mode == 'rt':
            f = interp1d(self.cot, self.ref, kind=method, bounds_error=False)
        return f(cot)


class func_ref_vs_cot(_lut_base):

    """Plane-parallel cloud with the tabulated Mie phase function of the nearest r_eff (er3t/rtm/mca/util.py:19-213)."""

    tabulated = True

    def __init__(self, cot, cer0=10.0, fdir='tmp-data', date=None, wavelength=650.0, surface_albedo=0.03,
                 atmospheric_profile=None, solar_zenith_angle=30.0, solar_azimuth_angle=0.0, sensor_zenith_angle=0.0,
                 sensor_azimuth_angle=0.0, sensor_altitude=705000.0, cloud_top_height=2.0, cloud_geometrical_thickness=1.0,
                 solver='3d', Nphoton=1e6, atm0=None, Ncpu='auto', output_tag='er3t', overwrite=True,
                 *, seed=None, device=0, solver_obj=None, Nrun=3, pha0=None, abs0=None):
        self.fname_atm = atmospheric_profile
        self._setup(cot, cer0, fdir, date if date is not None else default_date(), wavelength, surface_albedo, solar_zenith_angle,
                    solar_azimuth_angle, sensor_zenith_angle, sensor_azimuth_angle, sensor_altitude, cloud_top_height,
                    cloud_geometrical_thickness, solver, Nphoton, atm0, Ncpu, output_tag, overwrite, 1, 1, 0.1, 0.1,
                    seed, device, solver_obj, Nrun, pha0, abs0)


class func_ref_vs_cot_multi_pixel(_lut_base):

    """Nx x Ny homogeneous 3-D cloud per COT, HG with the Mie-derived asymmetry parameter (what mca_atm_3d assigns),
    IPA solver (er3t/rtm/mca/util.py:218-416)."""

    tabulated = False

    def __init__(self, cot, cer0=10.0, fdir='tmp-data', date=None, wavelength=650.0, surface_albedo=0.03,
                 solar_zenith_angle=30.0, solar_azimuth_angle=0.0, sensor_zenith_angle=0.0, sensor_azimuth_angle=0.0,
                 sensor_altitude=705000.0, Nphoton=1e6, cloud_top_height=2.0, cloud_geometrical_thickness=1.0,
                 solver='ipa', Nx=2, Ny=2, dx=0.1, dy=0.1, Ncpu=12, atm0=None, output_tag='er3t', overwrite=True,
                 *, seed=None, device=0, solver_obj=None, Nrun=3, pha0=None, abs0=None):
        self._setup(cot, cer0, fdir, date if date is not None else datetime.datetime.now(), wavelength, surface_albedo,
                    solar_zenith_angle, solar_azimuth_angle, sensor_zenith_angle, sensor_azimuth_angle, sensor_altitude,
                    cloud_top_height, cloud_geometrical_thickness, solver, Nphoton, atm0, Ncpu, output_tag, overwrite, Nx, Ny, dx, dy,
                    seed, device, solver_obj, Nrun, pha0, abs0)
