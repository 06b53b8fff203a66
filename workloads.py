"""
The five BASELINE.json configs as synthetic inputs ("C1" ... "C5"), built with the file-free input
builders of er3t_b200.pre and the er3t.rtm.mca-compatible adapters.  Every builder takes a `scale` in (0, 1] that
shrinks the horizontal grid and the photon count so that the same scene family can be compared with the CPU oracle in
seconds (tests/test_gpu_configs.py); scale = 1 is the named shape.

    kw, abs0 = workloads.build('C3', scale=0.0625)
    mca = er3t_b200.rtm.mca.mcarats_ng(**kw)          # target, solver, surface, sensors are part of kw

C1  plane-parallel water cloud tau = 10, r_eff = 10 um over a Lambertian surface, 650 nm flux, 1e6 photons
    (examples/00_er3t_mca.py style; 1-D cloud component as in er3t/rtm/mca/util.py:150-159)
C2  3-D LES-like field 480 x 480 x 100 voxels at 100 m, nadir radiance at 650 nm, 1e8 photons (projects/05 shape)
C3  MODIS-like scene 448 x 512 at 250 m, LSRT land BRDF per pixel, multi-angle views (projects/02 shape)
C4  OCO-2 O2 A-band: 768 x 960 at 250 m, 2-D Lambertian albedo map, wavelength x g sweep (projects/01 shape)
C5  SPN-S spectral irradiance: flux with gas absorption over a Cox-Munk ocean, 64 x 64 at 2 km (projects/03 shape)
"""

import datetime

import numpy as np

SEED = 20260101
DATE = datetime.datetime(2017, 8, 13)


_PHA_CACHE = {}


def _pha(wavelength, **kw):
    """pha_mie_wc memoised per process (the Mie table costs seconds and several configs share a wavelength)."""
    from er3t_b200.pre import pha_mie_wc
    key = (float(wavelength), tuple(sorted((k, tuple(v) if isinstance(v, (list, tuple)) else v) for k, v in kw.items())))
    if key not in _PHA_CACHE:
        _PHA_CACHE[key] = pha_mie_wc(wavelength=wavelength, **kw)
    return _PHA_CACHE[key]


def _round_even(v, lo=4):
    return max(lo, int(round(v / 2.0)) * 2)


def c1(scale=1.0, photons=1e6, hom3d=False):
    from er3t_b200.pre import atm_atmmod, abs_16g, pha_mie_wc, cld_gen_hom
    from er3t_b200.rtm.mca import mca_atm_1d, mca_atm_3d, mca_sca
    atm0 = atm_atmmod(levels=np.arange(0.0, 20.1, 1.0))
    abs0 = abs_16g(wavelength=650.0, atm_obj=atm0)
    pha0 = _pha(650.0, reff=[5.0, 10.0, 15.0, 20.0], nr=64)
    atm1d0 = mca_atm_1d(atm_obj=atm0, abs_obj=abs0)
    atm_3ds = []
    if hom3d:
        # the 2 x 2-column homogeneous-3D variant (cld_gen_hom, er3t/rtm/mca/util.py:340-364)
        cld0 = cld_gen_hom(cot0=10.0, cer0=10.0, altitude=np.array([1.5]), atm_obj=atm0, Nx=2, Ny=2, dx=0.1, dy=0.1)
        atm_3ds = [mca_atm_3d(cld_obj=cld0, atm_obj=atm0, pha_obj=pha0, quiet=True)]
    else:
        iref = int(np.argmin(np.abs(pha0.data['ref']['data'] - 10.0)))
        atm1d0.add_mca_1d_atm(ext1d=10.0 / 1000.0, omg1d=pha0.data['ssa']['data'][iref], apf1d=iref + 1, z_bottom=1.0, z_top=2.0)
    kw = dict(date=DATE, atm_1ds=[atm1d0], atm_3ds=atm_3ds, Ng=abs0.Ng, target='flux', surface_albedo=0.03, sca=mca_sca(pha_obj=pha0),
              solar_zenith_angle=30.0, solar_azimuth_angle=0.0, fdir='tmp-data/c1', Nrun=3, photons=max(1e4, photons * scale),
              weights=abs0.coef['weight']['data'], solver='3D', quiet=True, seed=SEED, iz3l_fix=True)
    return kw, abs0


def c2(scale=1.0, photons=1e8, nx=480, ny=480, nz3=100, nrun=3, parts=False):
    from er3t_b200.pre import atm_atmmod, abs_16g, pha_mie_wc, cld_gen_les
    from er3t_b200.rtm.mca import mca_atm_1d, mca_atm_3d, mca_sca
    s = np.sqrt(scale)
    nx, ny = _round_even(nx * s), _round_even(ny * s)
    nz3 = max(4, int(round(nz3 * min(1.0, 4.0 * s)))) if scale < 1.0 else nz3
    dz = 4.0 / nz3                                               # 3-D block spans 0.5 .. 4.5 km
    levels = np.concatenate(([0.0], 0.5 + dz * np.arange(nz3 + 1), np.arange(5.0, 20.1, 1.0)))
    atm0 = atm_atmmod(levels=levels)
    abs0 = abs_16g(wavelength=650.0, atm_obj=atm0)
    cld0 = cld_gen_les(Nx=nx, Ny=ny, dx=0.1, dy=0.1, altitude=0.5 + dz * (np.arange(nz3) + 0.5), seed=2, atm_obj=atm0)
    pha0 = _pha(650.0)
    kw = dict(date=DATE, atm_1ds=[mca_atm_1d(atm_obj=atm0, abs_obj=abs0)],
              atm_3ds=[mca_atm_3d(cld_obj=cld0, atm_obj=atm0, pha_obj=pha0, quiet=True)], Ng=abs0.Ng, target='radiance',
              surface_albedo=0.03, sca=mca_sca(pha_obj=pha0), solar_zenith_angle=28.9, solar_azimuth_angle=296.83,
              sensor_zenith_angle=0.0, sensor_azimuth_angle=0.0, sensor_altitude=705000.0, fdir='tmp-data/c2',
              Nrun=nrun, photons=max(1e4, photons * scale), weights=abs0.coef['weight']['data'], solver='3D', quiet=True, seed=SEED,
              iz3l_fix=True)
    if parts:
        return kw, abs0, dict(cld=cld0, atm=atm0, pha=pha0)
    return kw, abs0


def c3(scale=1.0, photons=1e8, views=((0.0, 0.0), (30.0, 90.0), (60.0, 200.0))):
    """LSRT land BRDF per pixel (f_iso, f_geo, f_vol ~ U(0.05-0.3, 0-0.05, 0-0.15), seeded); the reference traces one
    (VZA, VAA) per call (Rad_nrad = 1, er3t/rtm/mca/mcarats.py:301); further views ride along as extra sensors."""
    from er3t_b200.pre import atm_atmmod, abs_16g, pha_mie_wc, cld_gen_les, sfc_2d_gen
    from er3t_b200.rtm.mca import mca_atm_1d, mca_atm_3d, mca_sca, mca_sfc_2d
    s = np.sqrt(scale)
    nx, ny = _round_even(448 * s), _round_even(512 * s)
    atm0 = atm_atmmod(levels=np.arange(0.0, 20.1, 0.5))
    abs0 = abs_16g(wavelength=650.0, atm_obj=atm0)
    cld0 = cld_gen_les(Nx=nx, Ny=ny, dx=0.25, dy=0.25, altitude=np.arange(0.75, 4.0, 0.5), cloud_frac=0.4, corr_km=2.5,
                       cot_median=10.0, seed=3, atm_obj=atm0)
    pha0 = _pha(650.0)
    rng = np.random.default_rng(SEED)
    sfc0 = sfc_2d_gen(sfc_2d={'fiso': rng.uniform(0.05, 0.3, (nx, ny)), 'fgeo': rng.uniform(0.0, 0.05, (nx, ny)),
                              'fvol': rng.uniform(0.0, 0.15, (nx, ny))})
    extra = [dict(sensor_zenith_angle=v[0], sensor_azimuth_angle=v[1]) for v in views[1:]]
    kw = dict(date=DATE, atm_1ds=[mca_atm_1d(atm_obj=atm0, abs_obj=abs0)],
              atm_3ds=[mca_atm_3d(cld_obj=cld0, atm_obj=atm0, pha_obj=pha0, quiet=True)], Ng=abs0.Ng, target='radiance',
              surface_albedo=mca_sfc_2d(atm_obj=atm0, sfc_obj=sfc0, quiet=True), sca=mca_sca(pha_obj=pha0),
              solar_zenith_angle=35.0, solar_azimuth_angle=150.0, sensor_zenith_angle=views[0][0], sensor_azimuth_angle=views[0][1],
              sensor_altitude=705000.0, fdir='tmp-data/c3', Nrun=3, photons=max(1e4, photons * scale),
              weights=abs0.coef['weight']['data'], solver='3D', quiet=True, seed=SEED, iz3l_fix=True, extra_sensors=extra)
    return kw, abs0


def c4(scale=1.0, photons=1e9, nwvl=8):
    """O2 A-band sweep: `nwvl` wavelengths, each a correlated-k set of 8 ... 16 g with column absorption optical depths
    spanning 1e-3 ... 10 and a normalised slit as weights (er3t/pre/abs/abs_crk.py:1642,1658 contract).  Returns a LIST
    of (kw, abs) pairs -- one mcarats_ng call per wavelength like projects/01_oco2_rad-sim.py -- sharing one scene."""
    from er3t_b200.pre import atm_atmmod, abs_gen, pha_mie_wc, cld_gen_les, sfc_2d_gen
    from er3t_b200.rtm.mca import mca_atm_1d, mca_atm_3d, mca_sca, mca_sfc_2d
    s = np.sqrt(scale)
    nx, ny = _round_even(768 * s), _round_even(960 * s)
    atm0 = atm_atmmod(levels=np.arange(0.0, 20.1, 0.5))
    cld0 = cld_gen_les(Nx=nx, Ny=ny, dx=0.25, dy=0.25, altitude=np.arange(0.75, 3.0, 0.5), cloud_frac=0.3, corr_km=3.0,
                       cot_median=6.0, seed=4, atm_obj=atm0)
    pha0 = _pha(770.0)
    rng = np.random.default_rng(SEED + 4)
    sfc0 = sfc_2d_gen(sfc_2d=rng.uniform(0.1, 0.4, (nx, ny)).astype(np.float32))
    sfc = mca_sfc_2d(atm_obj=atm0, sfc_obj=sfc0, quiet=True)
    atm3d = mca_atm_3d(cld_obj=cld0, atm_obj=atm0, pha_obj=pha0, quiet=True)
    sca = mca_sca(pha_obj=pha0)
    p = atm0.lev['pressure']['data']
    frac = (p[:-1] - p[1:]) / (p[0] - p[-1])
    out = []
    for iw in range(nwvl):
        ng = 8 + (iw * 8) // max(1, nwvl - 1) if nwvl > 1 else 8                      # 8 ... 16 g
        tau_max = 10.0 ** (-1.0 + 2.0 * iw / max(1, nwvl - 1))                        # line wing ... line centre
        tau_col = 1.0e-3 * (tau_max / 1.0e-3) ** (np.arange(ng) / max(1, ng - 1))
        slit = np.exp(-0.5 * ((np.arange(ng) - 0.5 * (ng - 1)) / (0.35 * ng)) ** 2)
        wgt = slit / slit.sum()
        abs0 = abs_gen(759.0 + 1.0 * iw, frac[:, None] * tau_col[None, :], wgt, solar=np.full(ng, 1.25))
        kw = dict(date=DATE, atm_1ds=[mca_atm_1d(atm_obj=atm0, abs_obj=abs0)], atm_3ds=[atm3d], Ng=ng, target='radiance',
                  surface_albedo=sfc, sca=sca, solar_zenith_angle=40.0, solar_azimuth_angle=120.0, sensor_zenith_angle=0.3,
                  sensor_azimuth_angle=10.0, sensor_altitude=705000.0, fdir='tmp-data/c4/w%02d' % iw, Nrun=3,
                  photons=max(1e4, photons * scale / nwvl), weights=wgt, solver='3D', quiet=True, seed=SEED + 100 * iw, iz3l_fix=True)
        out.append((kw, abs0))
    return out


def c5(scale=1.0, photons=1e7, segment=0, pha0=None):
    """Flux with gas absorption over a Cox-Munk ocean (cal_ocean_brdf(745 nm, u10 = 5 m/s), er3t/pre/sfc/util.py:14-150)
    under broken 3-D clouds; one flight segment of projects/03_spns_flux-sim.py (2 km pixels, 1 km levels)."""
    from er3t_b200.pre import atm_atmmod, abs_16g, pha_mie_wc, cld_gen_les, sfc_2d_gen, cal_ocean_brdf
    from er3t_b200.rtm.mca import mca_atm_1d, mca_atm_3d, mca_sca, mca_sfc_2d
    s = np.sqrt(scale)
    nx = ny = _round_even(64 * s)
    atm0 = atm_atmmod(levels=np.arange(0.0, 20.1, 1.0))
    abs0 = abs_16g(wavelength=745.0, atm_obj=atm0, tau_max=1.0)
    cld0 = cld_gen_les(Nx=nx, Ny=ny, dx=2.0, dy=2.0, altitude=np.array([1.5, 2.5, 3.5]), cloud_frac=0.5, corr_km=12.0,
                       cot_median=12.0, seed=5 + 31 * segment, atm_obj=atm0)
    pha0 = _pha(745.0) if pha0 is None else pha0
    oc = cal_ocean_brdf(wvl=745.0, u10=np.full((nx, ny), 5.0))
    sfc0 = sfc_2d_gen(sfc_2d=oc)
    kw = dict(date=DATE, atm_1ds=[mca_atm_1d(atm_obj=atm0, abs_obj=abs0)],
              atm_3ds=[mca_atm_3d(cld_obj=cld0, atm_obj=atm0, pha_obj=pha0, quiet=True)], Ng=abs0.Ng, target='flux',
              surface_albedo=mca_sfc_2d(atm_obj=atm0, sfc_obj=sfc0, quiet=True), sca=mca_sca(pha_obj=pha0),
              solar_zenith_angle=45.0, solar_azimuth_angle=200.0, fdir='tmp-data/c5', Nrun=3, photons=max(1e4, photons * scale),
              weights=abs0.coef['weight']['data'], solver='3D', quiet=True, seed=SEED, iz3l_fix=True)
    return kw, abs0


def c5_segments(scale=1.0, photons=1e7, nseg=30):
    """The ~30 flight-track segments of projects/03_spns_flux-sim.py:48-53: one independent scene (own cloud field) and
    one mcarats_ng call per segment, 1e7 photons each.  Returns a LIST of (kw, abs) pairs like c4."""
    from er3t_b200.pre import pha_mie_wc
    pha0 = _pha(745.0)
    return [c5(scale, photons, segment=i, pha0=pha0) for i in range(nseg)]


def build(name, scale=1.0, **kw):
    """C1 ... C5, plus the variants SURVEY.md 8d names: C1H (2 x 2-column homogeneous-3-D variant of C1), C2R (C2 as the
    reference runs it: 10 layers of 400 m, L2-resident), C3V1 / C3V9 (one view -- what the reference traces per call --
    and nine views 0 ... 60 deg in one pass), C5S (30 flight segments, one scene each)."""
    name = name.upper()
    if name == 'C1H':
        return c1(scale, hom3d=True, **kw)
    if name == 'C2R':
        return c2(scale, nz3=10, **kw)
    if name == 'C3V1':
        return c3(scale, views=((0.0, 0.0),), **kw)
    if name == 'C3V9':
        return c3(scale, views=tuple((7.5 * i, 40.0 * i) for i in range(9)), **kw)
    if name == 'C5S':
        return c5_segments(scale, **kw)
    if name == 'C1':
        return c1(scale, **kw)
    if name == 'C2':
        return c2(scale, **kw)
    if name == 'C3':
        return c3(scale, **kw)
    if name == 'C4':
        return c4(scale, **kw)
    if name == 'C5':
        return c5(scale, **kw)
    raise ValueError('unknown config %s (C1 ... C5)' % name)
