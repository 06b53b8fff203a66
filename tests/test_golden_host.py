"""
Host-side adapter functions against golden vectors produced by the REFERENCE's own code
(tests/golden/make_golden.py imports hong-chen/er3t from /root/reference and runs its functions on seeded inputs).
Integer / byte results must match bit for bit; float arithmetic restated with the same operations matches to rounding.
"""

import datetime
import os
import types

import numpy as np
import pytest

import er3t_b200.pre as bpre
from er3t_b200.rtm import mca as bmca
from er3t_b200 import util as butil

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, 'golden', 'reference_vectors.npz'), allow_pickle=False)

W16 = np.array([0.1527534276, 0.1491729617, 0.1420961469, 0.1316886544, 0.1181945205, 0.1019300893, 0.0832767040, 0.0626720116,
                0.0424925000, 0.0046269894, 0.0038279891, 0.0030260086, 0.0022199750, 0.0014140010, 0.0005330000, 0.000075])


def test_distribute_photon_known_answer_of_reference_tests():
    # the only numeric vector in the reference's test tree: tests/00_test_util.py:249-252
    expected = np.array([14824075, 14483931, 13811633, 12822922, 11540979, 9995858, 8223786, 6266341, 4349287, 752063,
                         676158, 599970, 523397, 446830, 363135, 319635])
    got = bmca.distribute_photon(1e8, W16)
    assert np.array_equal(got, expected)
    assert got.sum() == 100000000
    assert np.array_equal(bpre.abs.WEIGHT_16G, W16)


def test_distribute_photon_golden():
    assert np.array_equal(bmca.distribute_photon(1e8, W16), G['dp_w16_1e8'])
    assert np.array_equal(bmca.distribute_photon(1e6, W16, base_ratio=0.2), G['dp_w16_1e6_b02'])
    assert np.array_equal(bmca.distribute_photon(12345, np.repeat(1.0 / 7, 7)), G['dp_even_12345'])
    assert np.array_equal(bmca.distribute_photon(3e7, G['dp_rand_w'], base_ratio=0.05), G['dp_rand_3e7'])


def test_cal_mca_azimuth_golden():
    got = np.array([bmca.cal_mca_azimuth(a) for a in G['az_in']])
    assert np.array_equal(got, G['az_out'])


@pytest.mark.parametrize('i', range(5))
def test_rearrange_jobs_golden(i):
    got = bmca.rearrange_jobs(int(G['rj%d_ncpu' % i]), G['rj%d_w' % i])
    assert np.array_equal(got, G['rj%d_out' % i])
    assert sorted(got.tolist()) == list(range(G['rj%d_w' % i].size))      # a permutation: every job runs exactly once


def test_small_utils_golden():
    dates = [datetime.datetime(2017, 8, 13), datetime.datetime(2019, 1, 4), datetime.datetime(2020, 7, 4), datetime.datetime(2024, 12, 31)]
    assert np.array_equal(np.array([butil.cal_sol_fac(d) for d in dates]), G['solfac'])
    assert np.array_equal(butil.get_lay_index(G['gli_lay'], G['gli_ref']), G['gli_out'])
    with pytest.raises(ValueError):
        butil.get_lay_index(np.array([25.0]), G['gli_ref'])
    assert butil.nice_array_str(G['nas_in']) == str(G['nas_out'])
    assert np.array_equal(np.array([butil.cal_r_twostream(t, a=0.03, g=0.85, mu=0.866) for t in (0.5, 5.0, 50.0)]), G['r2s'])
    ph = bpre.pha_hg(asy_params=[0.0, 0.5, 0.85], angles=np.linspace(0.0, 180.0, 181))
    assert np.array_equal(ph.data['pha']['data'], G['hg_pha'])


def test_cal_ocean_brdf_golden():
    ob = bpre.cal_ocean_brdf(wvl=745.0, u10=5.0)
    got = np.array([ob['diffuse_alb'], ob['diffuse_frac'], ob['refrac_r'], ob['refrac_i'], ob['slope']])
    assert np.allclose(got, G['ocean_745_5'], rtol=1e-14, atol=0)
    ob = bpre.cal_ocean_brdf(wvl=650.0, u10=12.0, whitecaps=False)
    got = np.array([ob['diffuse_alb'], ob['diffuse_frac'], ob['refrac_r'], ob['refrac_i'], ob['slope']])
    assert np.allclose(got, G['ocean_650_12_nowc'], rtol=1e-14, atol=0)
    # 2-D wind field (the reference's 2-D path is broken on NumPy 2; must equal the scalar path pixel by pixel)
    ob2 = bpre.cal_ocean_brdf(wvl=745.0, u10=np.full((3, 2), 5.0))
    assert ob2['slope'].shape == (3, 2) and np.allclose(ob2['slope'], G['ocean_745_5'][4])


def _atm_abs():
    atm0 = bpre.atm_atmmod(levels=np.linspace(0.0, 20.0, 21))
    abs0 = bpre.abs_16g(wavelength=650.0, atm_obj=atm0)
    return atm0, abs0


def test_mca_atm_1d_golden():
    atm0, abs0 = _atm_abs()
    a1 = bmca.mca_atm_1d(atm_obj=atm0, abs_obj=abs0)
    a1.add_mca_1d_atm(ext1d=0.01, omg1d=0.999, apf1d=2.0, z_bottom=1.0, z_top=2.0)
    assert a1.nml[0]['Atm_np1d']['data'] == int(G['a1d_np1d'])
    assert np.allclose(butil.cal_mol_ext(0.65, atm0.lev['pressure']['data'][:-1], atm0.lev['pressure']['data'][1:], atm0), G['molext'], rtol=1e-13)
    for ig in (0, 7, 15):
        for key in ('Atm_zgrd0', 'Atm_ext1d(1:, 1)', 'Atm_abs1d(1:, 1)', 'Atm_omg1d(1:, 1)', 'Atm_apf1d(1:, 1)', 'Atm_ext1d(1:, 2)',
                    'Atm_omg1d(1:, 2)', 'Atm_apf1d(1:, 2)', 'Atm_tmp1d'):
            ref = G['a1d_g%d_%s' % (ig, key)]
            got = np.asarray(a1.nml[ig][key]['data'], dtype=np.float64)
            assert np.allclose(got, ref, rtol=1e-13, atol=0), (ig, key)


def _cloud():
    cld0 = bpre.cld_gen_hem(Nx=12, Ny=10, dx=0.1, dy=0.1, altitude=np.arange(1.25, 3.3, 0.5), radii=[0.3, 0.5], cloud_frac_tgt=0.3, seed=3)
    cld0.lay['cer']['data'] = np.load(os.path.join(HERE, 'golden', 'cer.npy'))
    return cld0


@pytest.mark.parametrize('tag', ['none', 'hg', 'mie'])
def test_mca_atm_3d_golden(tag, tmp_path):
    atm0, abs0 = _atm_abs()
    cld0 = _cloud()
    pobj = None
    if tag == 'hg':
        pobj = bpre.pha_hg(asy_params=[0.0, 0.8, 0.86], angles=np.linspace(0.0, 180.0, 37))
    elif tag == 'mie':
        pobj = types.SimpleNamespace(ID='Mie (Water Clouds)', data={
            'id': {'data': 'Mie'}, 'ang': {'data': np.linspace(0, 180, 19)}, 'pha': {'data': np.ones((19, 4))},
            'ssa': {'data': np.array([0.99999, 0.9999, 0.9995, 0.999])}, 'asy': {'data': np.array([0.80, 0.85, 0.87, 0.88])},
            'ref': {'data': np.array([4.0, 8.0, 12.0, 20.0])}})
    a3 = bmca.mca_atm_3d(atm_obj=atm0, cld_obj=cld0, pha_obj=pobj, quiet=True)
    meta = G['a3d_%s_meta' % tag]
    got_meta = [a3.nml['Atm_nx']['data'], a3.nml['Atm_ny']['data'], a3.nml['Atm_nz3']['data'], a3.nml['Atm_iz3l']['data'],
                a3.nml['Atm_dx']['data'], a3.nml['Atm_dy']['data'], a3.nml['Atm_np3d']['data']]
    assert np.allclose(got_meta, meta, rtol=1e-14)
    for key in ('Atm_extp3d', 'Atm_omgp3d', 'Atm_apfp3d', 'Atm_tmpa3d', 'Atm_abst3d'):
        ref = G['a3d_%s_%s' % (tag, key)]
        got = a3.nml[key]['data']
        assert got.shape == ref.shape and got.dtype == ref.dtype, key
        if key == 'Atm_tmpa3d':
            assert np.allclose(got, ref, atol=1e-4)       # float32 subtraction order
        else:
            assert np.allclose(got, ref, rtol=2e-7, atol=0), key
    if tag == 'mie':
        nz3 = a3.nml['Atm_nz3']['data']
        a3.add_mca_3d_atm(ext3d=np.full((12, 10, nz3), 1e-4), omg3d=np.full((12, 10, nz3), 0.9), apf3d=np.full((12, 10, nz3), 0.6))
        f = str(tmp_path / 'atm3d.bin')
        a3.gen_mca_3d_atm_file(f)
        got = np.fromfile(f, dtype='<f4')
        ref = G['a3d_file_bytes'].view('<f4')
        assert got.size == ref.size
        nvox = 12 * 10 * nz3
        assert np.allclose(got[:nvox], ref[:nvox], atol=1e-4)            # tmpa3d block
        assert np.allclose(got[nvox:], ref[nvox:], rtol=2e-7, atol=0)    # abst3d + (ext, omg, apf) per component, x fastest


def test_mca_sca_golden(tmp_path):
    pobj = bpre.pha_hg(asy_params=[0.0, 0.8, 0.86], angles=np.linspace(0.0, 180.0, 37))
    f = str(tmp_path / 'sca.bin')
    sca = bmca.mca_sca(pha_obj=pobj, fname=f, quiet=True)
    assert np.array_equal(np.fromfile(f, dtype=np.uint8), G['sca_file_bytes'])
    assert [sca.nml['Sca_npf']['data'], sca.nml['Sca_nangi']['data'], sca.nml['Sca_nskip']['data'], sca.nml['Sca_nanci']['data']] == G['sca_meta'].tolist()


@pytest.mark.parametrize('tag', ['lambert', 'lsrt', 'dsm'])
def test_mca_sfc_2d_golden(tag, tmp_path):
    atm0, abs0 = _atm_abs()
    if tag == 'lambert':
        val = G['sfc_lambert_in'].copy()
    elif tag == 'lsrt':
        val = {k: G['sfc_lsrt_in_' + k] for k in ('fiso', 'fvol', 'fgeo')}
    else:
        val = {k: G['sfc_dsm_in_' + k] for k in ('diffuse_alb', 'diffuse_frac', 'refrac_r', 'refrac_i', 'slope')}
    s2 = bpre.sfc_2d_gen(sfc_2d=val)
    f = str(tmp_path / 'sfc.bin')
    ms = bmca.mca_sfc_2d(atm_obj=atm0, sfc_obj=s2, fname=f, quiet=True)
    assert np.array_equal(ms.nml['Sfc_jsfc2d']['data'], G['sfc_%s_jsfc' % tag])
    assert np.array_equal(np.asarray(ms.nml['Sfc_psfc2d']['data'], dtype=np.float32), np.asarray(G['sfc_%s_psfc' % tag], dtype=np.float32))
    assert np.array_equal(np.fromfile(f, dtype=np.uint8), G['sfc_%s_file_bytes' % tag])


def test_mca_inp_file_golden(tmp_path):
    z = np.load(os.path.join(HERE, 'golden', 'inp_nml.npz'), allow_pickle=False)
    nml = {}
    for k in z.files:
        key = k.replace('<', '(').replace('>', ')').replace(';', ':').replace('|', ', ')
        v = z[k]
        if v.ndim == 0:
            v = v.item()
        nml[key] = v
    f = str(tmp_path / 'inp.txt')
    bmca.mca_inp_file(f, nml, comment=False)
    assert open(f).read() == str(G['inp_text'])
    with pytest.raises(OSError):
        bmca.mca_inp_file(f, {'Wld_typo': 1})


def _fake_mca(kind, tmp_path, Nx=3, Ny=2, Nz=5, Ng=4, Nrun=3):
    m = types.SimpleNamespace(Ng=Ng, Nrun=Nrun, date=datetime.datetime(2017, 8, 13), target=kind,
                              photons=np.tile(np.array([4, 3, 2, 1]) * 1000, Nrun), fnames_out=[])
    for ir in range(Nrun):
        row = []
        for ig in range(Ng):
            f = str(tmp_path / ('%s_r%02d.g%03d.out.bin' % (kind, ir, ig)))
            if kind == 'flux':
                raw = G['outw_raw_flux'][ir, ig]
                bmca.write_mca_out_raw(f, [('a1', 'Fdn0', raw[0]), ('a2', 'Fdn', raw[1]), ('a3', 'Fup', raw[2])])
            else:
                bmca.write_mca_out_raw(f, [('b1', 'Radiance', G['outw_raw_rad'][ir, ig])])
            row.append(f)
        m.fnames_out.append(row)
    return m


def test_output_weighting_golden(tmp_path):
    """mca_out_raw parsing and the g-weighting / run statistics of read_flux_mca_out / read_radiance_mca_out."""
    absx = types.SimpleNamespace(coef={'weight': {'data': G['outw_weight']}, 'solar': {'data': G['outw_solar']}, 'slit_func': {'data': G['outw_slit']}})
    mf = _fake_mca('flux', tmp_path)
    raw0 = bmca.mca_out_raw(mf.fnames_out[1][2])
    assert raw0.data[0]['dims'] == G['raw_parse_dims'].tolist()
    assert np.array_equal(raw0.data[1]['data'], G['raw_parse_v1'])
    for mode in ('mean', 'all'):
        d = bmca.read_flux_mca_out(mf, absx, mode=mode, squeeze=True)
        for k in d:
            ref = G['outw_flux_%s_%s' % (mode, k)]
            got = np.asarray(d[k]['data'])
            assert got.shape == ref.shape, (mode, k)
            assert np.allclose(got, ref, rtol=3e-6, atol=1e-7), (mode, k)
        mr = _fake_mca('radiance', tmp_path)
        d = bmca.read_radiance_mca_out(mr, absx, mode=mode, squeeze=True)
        for k in d:
            ref = G['outw_rad_%s_%s' % (mode, k)]
            assert np.allclose(np.asarray(d[k]['data']), ref, rtol=3e-6, atol=1e-7), (mode, k)
    dn = bmca.read_flux_mca_out(mf, absx, mode='mean', squeeze=False)
    assert dn['f_up']['data'].shape == G['outw_flux_nosq_f_up'].shape
    assert np.allclose(dn['f_up']['data'], G['outw_flux_nosq_f_up'], rtol=3e-6)
    with pytest.raises(OSError):
        bmca.read_flux_mca_out(mf, absx, mode='median')
