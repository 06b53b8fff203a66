"""
File-free 3-D cloud generators with the `.lay` / `.lev` payload of er3t.pre.cld
(er3t/pre/cld/cld_gen.py:19-696, er3t/pre/cld/cld_les.py:211-230): altitude/thickness in km, extinction in 1/m,
cer in micron, arrays shaped (Nx, Ny, Nz).

  cld_gen_hom   homogeneous slab, same arithmetic as the reference (cld_gen.py:600-696)
  cld_gen_hem   hemispherical clouds dropped at seeded random positions (statistics of cld_gen.py:180-358;
                explicit `seed` instead of the global numpy RNG)
  cld_gen_les   seeded synthetic stand-in for the LES scene that cld_les reads from les.nc (absent):
                cumulus-like towers from a smoothed random field, adiabatic-like extinction growing with height
"""

import numpy as np

from .atm import atm_atmmod

__all__ = ['cld_gen_hom', 'cld_gen_hem', 'cld_gen_les']


def _frame(obj, altitude, Nx, Ny, dx, dy, atm_obj=None):
    altitude = np.asarray(altitude, dtype=np.float64)
    dz = float(altitude[1] - altitude[0]) if altitude.size > 1 else 1.0
    obj.altitude = altitude
    obj.Nx, obj.Ny, obj.Nz = int(Nx), int(Ny), altitude.size
    obj.dx, obj.dy, obj.dz = dx, dy, dz
    obj.x = np.arange(Nx) * dx
    obj.y = np.arange(Ny) * dy
    obj.z = altitude - altitude[0]
    alt_lev = np.append(altitude - dz / 2.0, altitude[-1] + dz / 2.0)
    obj.lev = {'altitude': {'data': alt_lev, 'name': 'Altitude', 'units': 'km'}}
    obj.lay = {
        'x': {'data': obj.x, 'name': 'X', 'units': 'km'}, 'y': {'data': obj.y, 'name': 'Y', 'units': 'km'},
        'z': {'data': obj.z, 'name': 'Z', 'units': 'km'},
        'nx': {'data': obj.Nx, 'name': 'Nx', 'units': 'N/A'}, 'ny': {'data': obj.Ny, 'name': 'Ny', 'units': 'N/A'},
        'nz': {'data': obj.Nz, 'name': 'Nz', 'units': 'N/A'},
        'dx': {'data': dx, 'name': 'dx', 'units': 'km'}, 'dy': {'data': dy, 'name': 'dy', 'units': 'km'},
        'dz': {'data': dz, 'name': 'dz', 'units': 'km'},
        'altitude': {'data': altitude, 'name': 'Altitude', 'units': 'km'},
        'thickness': {'data': alt_lev[1:] - alt_lev[:-1], 'name': 'Layer thickness', 'units': 'km'},
    }
    if atm_obj is None:
        t_1d = atm_atmmod(levels=alt_lev).lay['temperature']['data']
    else:
        t_1d = np.interp(altitude, atm_obj.lay['altitude']['data'], atm_obj.lay['temperature']['data'])
    t_3d = np.empty((obj.Nx, obj.Ny, obj.Nz), dtype=np.float32)
    t_3d[...] = t_1d[None, None, :]
    obj.lay['temperature'] = {'data': t_3d, 'name': 'Temperature', 'units': 'K'}


def _fill(obj, ext, cer):
    ext = np.asarray(ext, dtype=np.float32)
    obj.lay['extinction'] = {'data': ext, 'name': 'Extinction coefficients', 'units': 'm^-1'}
    obj.lay['cer'] = {'data': np.asarray(cer, dtype=np.float32), 'name': 'Cloud effective radius', 'units': 'micron'}
    cot = ext * (obj.lay['thickness']['data'][None, None, :] * 1000.0)
    obj.lay['cot'] = {'data': cot, 'name': 'Cloud optical thickness', 'units': 'N/A'}
    obj.lev['cot_2d'] = {'data': cot.sum(axis=-1), 'name': 'Cloud optical thickness', 'units': 'N/A'}


class cld_gen_hom:

    ID = 'Homogeneous Cloud 3D'

    def __init__(self, fname=None, altitude=np.arange(1.5, 2.5, 0.5), Nx=10, Ny=10, dx=0.1, dy=0.1, cot0=10.0, cer0=10.0,
                 atm_obj=None, overwrite=False, verbose=False):
        self.fname = fname
        self.verbose = verbose
        _frame(self, altitude, Nx, Ny, dx, dy, atm_obj=atm_obj)
        self.cal_cld_opt_prop(cot0=cot0, cer0=cer0)

    def cal_cld_opt_prop(self, cot0=10.0, cer0=10.0, cot_scale=1.0):
        # cld_gen.py:659-696: the column optical thickness is split evenly over the Nz layers
        cot_lay = cot0 * cot_scale / self.Nz
        ext0 = cot_lay / self.dz / 1000.0
        shape = (self.Nx, self.Ny, self.Nz)
        _fill(self, np.full(shape, ext0), np.full(shape, cer0))


class cld_gen_hem:

    ID = 'Hemispherical Cloud 3D'

    def __init__(self, fname=None, altitude=np.arange(1.5, 6.7, 0.1), Nx=400, Ny=400, dx=0.1, dy=0.1,
                 radii=(1.0, 2.0, 4.0), weights=None, w2h_ratio=1.0, min_dist=0.2, cloud_frac_tgt=0.2, seed=0,
                 ext0=0.03, cer0=12.0, overwrite=False, verbose=False):
        self.fname = fname
        self.verbose = verbose
        altitude = np.asarray(altitude, dtype=np.float64)
        dz = altitude[1] - altitude[0]
        top = min(altitude[-1], max(radii) / w2h_ratio + altitude[0])
        altitude = np.arange(altitude[0], top + dz, dz)
        _frame(self, altitude, Nx, Ny, dx, dy)
        rng = np.random.default_rng(seed)
        radii = np.asarray(radii, dtype=np.float64)
        space = np.zeros((Nx, Ny, self.Nz), dtype=np.float32)
        taken = np.zeros((Nx, Ny), dtype=bool)
        X, Y = np.meshgrid(self.x, self.y, indexing='ij')
        Lx, Ly = Nx * dx, Ny * dy
        self.clouds = []
        tries = 0
        while taken.mean() < cloud_frac_tgt and tries < 20000:
            tries += 1
            r = rng.choice(radii, p=weights)
            cx, cy = rng.random() * Lx, rng.random() * Ly
            ddx = np.minimum(np.abs(X - cx), Lx - np.abs(X - cx))     # cyclic domain
            ddy = np.minimum(np.abs(Y - cy), Ly - np.abs(Y - cy))
            d2 = ddx ** 2 + ddy ** 2
            if np.any(taken & (d2 < (r + min_dist) ** 2)):
                continue
            h2 = (r * r - d2) / (w2h_ratio ** 2)
            height = np.sqrt(np.clip(h2, 0.0, None))                  # dome height above cloud base (km)
            space += (self.z[None, None, :] < height[:, :, None]) & (d2 < r * r)[:, :, None]
            taken |= d2 < r * r
            self.clouds.append({'x': cx, 'y': cy, 'radius': r})
        space = np.clip(space, 0.0, 1.0)
        self.space_3d = space
        self.cloud_frac = float(taken.mean())
        _fill(self, ext0 * space, cer0 * (space > 0))
        self.lev['cth_2d'] = {'data': (space * self.dz).sum(axis=-1) + self.lev['altitude']['data'][0],
                              'name': 'Cloud top height', 'units': 'km'}


def _smooth_periodic(f, sigma_px):
    """Gaussian smoothing of a periodic 2-D field by FFT."""
    nx, ny = f.shape
    kx = np.fft.fftfreq(nx)[:, None]
    ky = np.fft.fftfreq(ny)[None, :]
    filt = np.exp(-2.0 * (np.pi * sigma_px) ** 2 * (kx * kx + ky * ky))
    return np.real(np.fft.ifft2(np.fft.fft2(f) * filt))


class cld_gen_les:

    """
    Synthetic shallow-cumulus field with the shape of the LES scene used by examples/00_er3t_mca.py:973 and
    projects/05_cnn-les_rad-sim.py (480 x 480 columns at 100 m).

    A smoothed (correlation length `corr_km`) Gaussian random field is thresholded to the target cloud fraction; its
    excess over the threshold sets the cloud depth of each column; inside a cloud the extinction grows linearly with
    height above cloud base (adiabatic-like) and is scaled so that the median column optical thickness of cloudy
    columns is `cot_median`.  cer grows from `cer_base` to `cer_top` with height.  rng = default_rng(seed).
    """

    ID = 'LES-like Cloud 3D (synthetic)'

    def __init__(self, Nx=480, Ny=480, dx=0.1, dy=0.1, altitude=None, cloud_frac=0.25, corr_km=0.6, cot_median=8.0,
                 cer_base=6.0, cer_top=14.0, seed=2, atm_obj=None, verbose=False):
        if altitude is None:
            altitude = 0.52 + 0.04 * np.arange(100)               # 100 layers of 40 m from 0.5 km
        self.verbose = verbose
        _frame(self, altitude, Nx, Ny, dx, dy, atm_obj=atm_obj)
        rng = np.random.default_rng(seed)
        f = _smooth_periodic(rng.standard_normal((Nx, Ny)), corr_km / dx)
        f += 0.5 * _smooth_periodic(rng.standard_normal((Nx, Ny)), 0.25 * corr_km / dx) * f.std() / 0.5
        f = (f - f.mean()) / f.std()
        thr = np.quantile(f, 1.0 - cloud_frac)
        depth_frac = np.clip((f - thr) / (f.max() - thr), 0.0, 1.0) ** 0.7            # 0..1 of the block depth
        nz = self.Nz
        base = int(0.15 * nz)
        ztop = base + depth_frac * (nz - base)
        k = np.arange(nz)[None, None, :]
        incloud = (k >= base) & (k < ztop[:, :, None]) & (depth_frac[:, :, None] > 0)
        height = (k - base + 0.5) / max(1, nz - base)
        ext = np.where(incloud, 0.2 + height, 0.0)
        thick_m = self.lay['thickness']['data'] * 1000.0
        cot = (ext * thick_m[None, None, :]).sum(axis=-1)
        scale = cot_median / np.median(cot[cot > 0])
        ext = ext * scale
        cer = np.where(incloud, cer_base + (cer_top - cer_base) * height, 0.0)
        _fill(self, ext, cer)
        self.cloud_frac = float((cot > 0).mean())
