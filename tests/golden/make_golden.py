"""
Generate golden vectors by running the REFERENCE's own host-side functions (hong-chen/er3t at /root/reference) on
small seeded inputs.  Run in the build container only (the reference tree does not exist on the GPU box):

    python tests/golden/make_golden.py

Writes tests/golden/reference_vectors.npz (+ a few byte-exact files).  The reference needs h5py / netCDF4 / pyhdf /
matplotlib / cartopy ... at import time; none is used by the functions exercised here, so they are stubbed.
The input objects (atmosphere, absorption, cloud, phase function, surface) come from er3t_b200.pre, because the
reference's own builders need data files that are not shipped; the reference classes under test are
duck-typed on those payloads.
"""

import contextlib
import datetime
import io
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, '..', '..'))
REF = '/root/reference'


class _Stub(types.ModuleType):
    def __init__(self, name):
        super().__init__(name)
        import importlib.machinery
        self.__spec__ = importlib.machinery.ModuleSpec(name, None)
        self.__path__ = []

    def __getattr__(self, name):
        if name.startswith('__'):
            raise AttributeError(name)
        m = _Stub(self.__name__ + '.' + name)
        sys.modules[m.__name__] = m
        return m

    def __call__(self, *a, **k):
        return _Stub('call')


def import_reference():
    for name in ['h5py', 'netCDF4', 'pyhdf', 'pyhdf.SD', 'matplotlib', 'matplotlib.pyplot', 'matplotlib.path', 'matplotlib.image',
                 'matplotlib.patches', 'matplotlib.gridspec', 'matplotlib.axes', 'matplotlib.colors', 'matplotlib.ticker',
                 'mpl_toolkits', 'mpl_toolkits.axes_grid1', 'cartopy', 'cartopy.crs', 'owslib', 'owslib.wmts', 'pysolar', 'pysolar.solar',
                 'bs4', 'geopy', 'geopy.distance', 'xarray', 'h5netcdf']:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = _Stub(name)
    if not hasattr(np, 'float_'):
        pass
    sys.path.insert(0, REF)
    sys.path.insert(0, ROOT)
    import er3t
    return er3t


def main():
    er3t = import_reference()
    import er3t.rtm.mca as rmca
    from er3t.rtm.mca.mcarats import distribute_photon, cal_mca_azimuth
    from er3t.rtm.mca.mca_run import rearrange_jobs
    from er3t.rtm.mca.mca_out import mca_out_raw, read_flux_mca_out, read_radiance_mca_out
    from er3t.util import cal_sol_fac, get_lay_index, cal_mol_ext, nice_array_str, cal_r_twostream
    from er3t.pre.pha import pha_hg
    from er3t.pre.sfc import cal_ocean_brdf, sfc_2d_gen
    import er3t_b200.pre as bpre
    from er3t_b200.rtm.mca import write_mca_out_raw

    out = {}
    tmp = os.path.join(HERE, '_tmp')
    os.makedirs(tmp, exist_ok=True)
    sink = io.StringIO()

    # ---- A1 distribute_photon (incl. the vector hard-coded at tests/00_test_util.py:249-252)
    w16 = np.array([0.1527534276, 0.1491729617, 0.1420961469, 0.1316886544, 0.1181945205, 0.1019300893, 0.0832767040, 0.0626720116,
                    0.0424925000, 0.0046269894, 0.0038279891, 0.0030260086, 0.0022199750, 0.0014140010, 0.0005330000, 0.000075])
    out['dp_w16_1e8'] = distribute_photon(1e8, w16)
    out['dp_w16_1e6_b02'] = distribute_photon(1e6, w16, base_ratio=0.2)
    out['dp_even_12345'] = distribute_photon(12345, np.repeat(1.0 / 7, 7))
    rng = np.random.default_rng(5)
    wr = rng.random(9); wr /= wr.sum()
    out['dp_rand_w'] = wr
    out['dp_rand_3e7'] = distribute_photon(3e7, wr, base_ratio=0.05)

    # ---- A2 cal_mca_azimuth
    az = np.array([-720.5, -90.0, 0.0, 45.0, 90.0, 180.0, 269.9, 270.0, 296.83, 360.0, 450.0, 725.0])
    out['az_in'] = az
    out['az_out'] = np.array([cal_mca_azimuth(a) for a in az])

    # ---- A11 rearrange_jobs
    for i, (ncpu, w) in enumerate([(5, np.tile(out['dp_w16_1e8'], 3)), (12, np.tile(out['dp_w16_1e8'], 3)), (3, rng.integers(1, 1000, 20)),
                                   (8, np.arange(1, 31)), (4, np.array([5, 5, 5, 5, 5, 5]))]):
        out['rj%d_ncpu' % i] = ncpu
        out['rj%d_w' % i] = np.asarray(w)
        out['rj%d_out' % i] = rearrange_jobs(ncpu, np.asarray(w))

    # ---- A15 cal_sol_fac, get_lay_index, nice_array_str, two-stream
    dates = [datetime.datetime(2017, 8, 13), datetime.datetime(2019, 1, 4), datetime.datetime(2020, 7, 4), datetime.datetime(2024, 12, 31)]
    out['solfac_doy'] = np.array([d.timetuple().tm_yday for d in dates])
    out['solfac'] = np.array([cal_sol_fac(d) for d in dates])
    lay_ref = 0.5 * (np.linspace(0, 20, 41)[1:] + np.linspace(0, 20, 41)[:-1])
    lay = np.array([1.3, 1.7, 2.25, 6.1])
    out['gli_lay'] = lay; out['gli_ref'] = lay_ref
    out['gli_out'] = get_lay_index(lay, lay_ref)
    arr = np.array([1.0, 2.5e-6, 3.25, 1e10, -4.0, 0.333333333, 7.0])
    out['nas_in'] = arr
    out['nas_out'] = np.array(nice_array_str(arr))
    out['r2s'] = np.array([cal_r_twostream(t, a=0.03, g=0.85, mu=0.866) for t in (0.5, 5.0, 50.0)])

    # ---- pha_hg
    ph = pha_hg(asy_params=[0.0, 0.5, 0.85], angles=np.linspace(0.0, 180.0, 181))
    out['hg_pha'] = ph.data['pha']['data']

    # ---- cal_ocean_brdf (scalar path; the 2-D path of the reference uses np.float_ and fails on NumPy 2)
    ob = cal_ocean_brdf(wvl=745.0, u10=5.0)
    out['ocean_745_5'] = np.array([ob['diffuse_alb'], ob['diffuse_frac'], ob['refrac_r'], ob['refrac_i'], ob['slope']], dtype=np.float64)
    ob = cal_ocean_brdf(wvl=650.0, u10=12.0, whitecaps=False)
    out['ocean_650_12_nowc'] = np.array([ob['diffuse_alb'], ob['diffuse_frac'], ob['refrac_r'], ob['refrac_i'], ob['slope']], dtype=np.float64)

    # ---- A4/A5 mca_atm_1d, A6/A7 mca_atm_3d, A8 mca_sca, A9 mca_sfc_2d on er3t_b200.pre objects
    atm0 = bpre.atm_atmmod(levels=np.linspace(0.0, 20.0, 21))
    abs0 = bpre.abs_16g(wavelength=650.0, atm_obj=atm0)
    with contextlib.redirect_stdout(sink):
        a1 = rmca.mca_atm_1d(atm_obj=atm0, abs_obj=abs0)
        a1.add_mca_1d_atm(ext1d=0.01, omg1d=0.999, apf1d=2.0, z_bottom=1.0, z_top=2.0)
        out['molext'] = cal_mol_ext(0.65, atm0.lev['pressure']['data'][:-1], atm0.lev['pressure']['data'][1:], atm0)
    for ig in (0, 7, 15):
        for key in ('Atm_zgrd0', 'Atm_ext1d(1:, 1)', 'Atm_abs1d(1:, 1)', 'Atm_omg1d(1:, 1)', 'Atm_apf1d(1:, 1)', 'Atm_ext1d(1:, 2)',
                    'Atm_omg1d(1:, 2)', 'Atm_apf1d(1:, 2)', 'Atm_tmp1d'):
            out['a1d_g%d_%s' % (ig, key)] = np.asarray(a1.nml[ig][key]['data'], dtype=np.float64)
    out['a1d_np1d'] = a1.nml[0]['Atm_np1d']['data']

    cld0 = bpre.cld_gen_hem(Nx=12, Ny=10, dx=0.1, dy=0.1, altitude=np.arange(1.25, 3.3, 0.5), radii=[0.3, 0.5], cloud_frac_tgt=0.3, seed=3)
    cld0.lay['cer']['data'] = np.where(cld0.lay['extinction']['data'] > 0, 6.0 + 10.0 * np.random.default_rng(1).random(cld0.lay['extinction']['data'].shape), 0.0).astype(np.float32)
    np.save(os.path.join(tmp, 'cer.npy'), cld0.lay['cer']['data'])
    pha_mie = types.SimpleNamespace(ID='Mie (Water Clouds)', data={
        'id': {'data': 'Mie'}, 'ang': {'data': np.linspace(0, 180, 19)}, 'pha': {'data': np.ones((19, 4))},
        'ssa': {'data': np.array([0.99999, 0.9999, 0.9995, 0.999])}, 'asy': {'data': np.array([0.80, 0.85, 0.87, 0.88])},
        'ref': {'data': np.array([4.0, 8.0, 12.0, 20.0])}})
    pha_hg3 = pha_hg(asy_params=[0.0, 0.8, 0.86], angles=np.linspace(0.0, 180.0, 37))
    fbin = os.path.join(tmp, 'atm3d.bin')
    for tag, pobj in (('none', None), ('hg', pha_hg3), ('mie', pha_mie)):
        with contextlib.redirect_stdout(sink):
            a3 = rmca.mca_atm_3d(atm_obj=atm0, cld_obj=cld0, pha_obj=pobj, fname=fbin, overwrite=True, quiet=True)
        for key in ('Atm_extp3d', 'Atm_omgp3d', 'Atm_apfp3d', 'Atm_tmpa3d', 'Atm_abst3d'):
            out['a3d_%s_%s' % (tag, key)] = a3.nml[key]['data']
        out['a3d_%s_meta' % tag] = np.array([a3.nml['Atm_nx']['data'], a3.nml['Atm_ny']['data'], a3.nml['Atm_nz3']['data'], a3.nml['Atm_iz3l']['data'],
                                            a3.nml['Atm_dx']['data'], a3.nml['Atm_dy']['data'], a3.nml['Atm_np3d']['data']], dtype=np.float64)
    with contextlib.redirect_stdout(sink):
        a3.add_mca_3d_atm(ext3d=np.full((12, 10, out['a3d_mie_Atm_extp3d'].shape[2]), 1e-4), omg3d=np.full((12, 10, out['a3d_mie_Atm_extp3d'].shape[2]), 0.9),
                          apf3d=np.full((12, 10, out['a3d_mie_Atm_extp3d'].shape[2]), 0.6))
        a3.gen_mca_3d_atm_file(fbin)
    out['a3d_file_bytes'] = np.fromfile(fbin, dtype=np.uint8)

    fsca = os.path.join(tmp, 'sca.bin')
    with contextlib.redirect_stdout(sink):
        sca = rmca.mca_sca(pha_obj=pha_hg3, fname=fsca, overwrite=True, quiet=True)
    out['sca_file_bytes'] = np.fromfile(fsca, dtype=np.uint8)
    out['sca_meta'] = np.array([sca.nml['Sca_npf']['data'], sca.nml['Sca_nangi']['data'], sca.nml['Sca_nskip']['data'], sca.nml['Sca_nanci']['data']])

    r2 = np.random.default_rng(11)
    sfc_in = {
        'lambert': (r2.random((6, 5)) * 1.4 - 0.2).astype(np.float64),
        'lsrt': {'fiso': r2.random((6, 5)) * 0.3, 'fvol': r2.random((6, 5)) * 0.1, 'fgeo': r2.random((6, 5)) * 0.05},
        'dsm': {'diffuse_alb': np.full((6, 5), 0.22), 'diffuse_frac': np.full((6, 5), 0.001), 'refrac_r': np.full((6, 5), 1.34),
                'refrac_i': np.full((6, 5), 1e-7), 'slope': r2.random((6, 5)) * 0.05 + 0.003},
    }
    np.save(os.path.join(tmp, 'sfc_lambert.npy'), sfc_in['lambert'])
    for tag, val in sfc_in.items():
        with contextlib.redirect_stdout(sink):
            s2 = sfc_2d_gen(sfc_2d=val if not isinstance(val, np.ndarray) else val.copy())
            fs = os.path.join(tmp, 'sfc_%s.bin' % tag)
            ms = rmca.mca_sfc_2d(atm_obj=atm0, sfc_obj=s2, fname=fs, overwrite=True, quiet=True)
        out['sfc_%s_jsfc' % tag] = ms.nml['Sfc_jsfc2d']['data']
        out['sfc_%s_psfc' % tag] = ms.nml['Sfc_psfc2d']['data']
        out['sfc_%s_file_bytes' % tag] = np.fromfile(fs, dtype=np.uint8)
        if isinstance(val, dict):
            for k, v in val.items():
                out['sfc_%s_in_%s' % (tag, k)] = v
        else:
            out['sfc_%s_in' % tag] = val

    # ---- A10 mca_inp_file text
    nml = {'Wld_mverb': 0, 'Wld_jseed': 12345, 'Wld_mbswap': 0, 'Wld_mtarget': 2, 'Wld_moptim': 0, 'Wld_njob': 1, 'Sca_inpfile': 'sca.bin',
           'Sca_npf': 3, 'Sca_nanci': 0, 'Sca_nangi': 37, 'Sca_nskip': 0, 'Atm_inpfile': 'atm3d.bin', 'Atm_np1d': 2, 'Atm_np3d': 1, 'Atm_nx': 12,
           'Atm_ny': 10, 'Atm_nz': 20, 'Atm_iz3l': 3, 'Atm_nz3': 5, 'Atm_nkd': 1, 'Atm_mtprof': 0, 'Atm_wkd0': 1.0, 'Atm_dx': 100.0, 'Atm_dy': 100.0,
           'Atm_zgrd0': a1.nml[0]['Atm_zgrd0']['data'], 'Atm_tmp1d': a1.nml[0]['Atm_tmp1d']['data'],
           'Atm_ext1d(1:, 1)': a1.nml[0]['Atm_ext1d(1:, 1)']['data'], 'Atm_omg1d(1:, 1)': a1.nml[0]['Atm_omg1d(1:, 1)']['data'],
           'Atm_apf1d(1:, 1)': a1.nml[0]['Atm_apf1d(1:, 1)']['data'], 'Atm_abs1d(1:, 1)': a1.nml[3]['Atm_abs1d(1:, 1)']['data'],
           'Atm_ext1d(1:, 2)': a1.nml[0]['Atm_ext1d(1:, 2)']['data'], 'Atm_omg1d(1:, 2)': a1.nml[0]['Atm_omg1d(1:, 2)']['data'],
           'Atm_apf1d(1:, 2)': a1.nml[0]['Atm_apf1d(1:, 2)']['data'],
           'Sfc_mbrdf': np.array([1, 0, 0, 0]), 'Sfc_mtype': 1, 'Sfc_param(1)': 0.03, 'Src_flx': 1.0, 'Src_qmax': 0.533133, 'Src_dwlen': 0.0,
           'Src_mtype': 1, 'Src_mphi': 0, 'Src_the': 150.0, 'Src_phi': 225.0, 'Rad_mrkind': 2, 'Rad_mplen': 0, 'Rad_mpmap': 1, 'Rad_nrad': 1,
           'Rad_difr0': 7.5, 'Rad_difr1': 0.0025, 'Rad_the': 180.0, 'Rad_phi': 270.0, 'Rad_zloc': 705000.0, 'Rad_nxr': 12, 'Rad_nyr': 10}
    ftxt = os.path.join(tmp, 'inp.txt')
    rmca.mca_inp_file(ftxt, nml, comment=False)
    out['inp_text'] = np.array(open(ftxt).read())
    np.savez(os.path.join(HERE, 'inp_nml.npz'), **{k.replace('(', '<').replace(')', '>').replace(':', ';').replace(', ', '|'): v for k, v in nml.items()})

    # ---- A12-A14 mca_out_raw + read_flux / read_radiance weighting on synthetic per-job files
    Nx, Ny, Nz, Ng, Nrun = 3, 2, 5, 4, 3
    r3 = np.random.default_rng(21)
    mca_flux = types.SimpleNamespace(Ng=Ng, Nrun=Nrun, date=datetime.datetime(2017, 8, 13), target='flux', photons=np.tile(np.array([4, 3, 2, 1]) * 1000, Nrun), fnames_out=[])
    mca_rad = types.SimpleNamespace(Ng=Ng, Nrun=Nrun, date=datetime.datetime(2017, 8, 13), target='radiance', photons=np.tile(np.array([4, 3, 2, 1]) * 1000, Nrun), fnames_out=[])
    raw_flux = r3.random((Nrun, Ng, 3, Nx, Ny, Nz)).astype(np.float32)
    raw_rad = r3.random((Nrun, Ng, Nx, Ny, 1)).astype(np.float32)
    for ir in range(Nrun):
        rowf, rowr = [], []
        for ig in range(Ng):
            ff = os.path.join(tmp, 'f_r%02d.g%03d.out.bin' % (ir, ig))
            write_mca_out_raw(ff, [('a1', 'Fdn0', raw_flux[ir, ig, 0]), ('a2', 'Fdn', raw_flux[ir, ig, 1]), ('a3', 'Fup', raw_flux[ir, ig, 2])])
            fr = os.path.join(tmp, 'r_r%02d.g%03d.out.bin' % (ir, ig))
            write_mca_out_raw(fr, [('b1', 'Radiance', raw_rad[ir, ig])])
            rowf.append(ff); rowr.append(fr)
        mca_flux.fnames_out.append(rowf); mca_rad.fnames_out.append(rowr)
    absx = types.SimpleNamespace(coef={'weight': {'data': np.array([0.4, 0.3, 0.2, 0.1])}, 'solar': {'data': np.array([1.5, 1.6, 1.7, 1.8])},
                                       'slit_func': {'data': 0.5 + r3.random((Nz - 1, Ng))}})
    out['outw_weight'] = absx.coef['weight']['data']; out['outw_solar'] = absx.coef['solar']['data']; out['outw_slit'] = absx.coef['slit_func']['data']
    out['outw_raw_flux'] = raw_flux; out['outw_raw_rad'] = raw_rad
    raw0 = mca_out_raw(mca_flux.fnames_out[1][2])
    out['raw_parse_dims'] = np.array(raw0.data[0]['dims'])
    out['raw_parse_v1'] = raw0.data[1]['data']
    for mode in ('mean', 'all'):
        df = read_flux_mca_out(mca_flux, absx, mode=mode, squeeze=True)
        for k in df:
            if isinstance(df[k]['data'], np.ndarray) or np.isscalar(df[k]['data']):
                out['outw_flux_%s_%s' % (mode, k)] = np.asarray(df[k]['data'])
        dr = read_radiance_mca_out(mca_rad, absx, mode=mode, squeeze=True)
        for k in dr:
            out['outw_rad_%s_%s' % (mode, k)] = np.asarray(dr[k]['data'])
    dfn = read_flux_mca_out(mca_flux, absx, mode='mean', squeeze=False)
    out['outw_flux_nosq_f_up'] = dfn['f_up']['data']

    np.savez_compressed(os.path.join(HERE, 'reference_vectors.npz'), **out)
    print('wrote %d arrays to %s' % (len(out), os.path.join(HERE, 'reference_vectors.npz')))
    import shutil
    shutil.copy(os.path.join(tmp, 'cer.npy'), os.path.join(HERE, 'cer.npy'))
    shutil.rmtree(tmp)


if __name__ == '__main__':
    main()
