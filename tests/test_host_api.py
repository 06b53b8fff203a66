"""Host logic of the er3t.rtm.mca mirror, exercised end to end on the CPU through a test double of the solver that is
backed by the oracle (tests/oracle_solver.py): namelist construction, scene hand-off, job list, fused vs raw weighting,
compatibility files, persistence, error behaviour."""

import datetime
import os

import numpy as np
import pytest

import er3t_b200.pre as bpre
from er3t_b200.rtm import mca as bmca
from oracle_solver import OracleSolver


@pytest.fixture(scope='module')
def inputs():
    atm0 = bpre.atm_atmmod(levels=np.linspace(0, 20, 21))
    abs0 = bpre.abs_16g(wavelength=650.0, atm_obj=atm0)
    cld0 = bpre.cld_gen_hom(altitude=np.array([1.5]), Nx=4, Ny=3, dx=0.1, dy=0.1, cot0=10.0, cer0=10.0, atm_obj=atm0)
    pha0 = bpre.pha_mie_wc(wavelength=650.0, reff=[5.0, 10.0, 15.0], nr=48)
    sca = bmca.mca_sca(pha_obj=pha0)
    atm3d0 = bmca.mca_atm_3d(cld_obj=cld0, atm_obj=atm0, pha_obj=pha0, quiet=True)
    atm1d0 = bmca.mca_atm_1d(atm_obj=atm0, abs_obj=abs0)
    return dict(atm0=atm0, abs0=abs0, sca=sca, atm3d0=atm3d0, atm1d0=atm1d0)


def run(inputs, tmp_path, **kw):
    args = dict(date=datetime.datetime(2017, 8, 13), atm_1ds=[inputs['atm1d0']], atm_3ds=[inputs['atm3d0']], Ng=16, target='radiance',
                surface_albedo=0.03, sca=inputs['sca'], solar_zenith_angle=30.0, solar_azimuth_angle=45.0, fdir=str(tmp_path / 'sim'),
                Nrun=3, photons=6e4, weights=inputs['abs0'].coef['weight']['data'], solver='3D', quiet=True, seed=1, solver_obj=OracleSolver())
    args.update(kw)
    return bmca.mcarats_ng(**args)


def test_namelist_matches_reference_contract(inputs, tmp_path):
    m = run(inputs, tmp_path)
    n = m.nml[0]
    # values er3t sets (er3t/rtm/mca/mcarats.py:250-307,374-399) -- SURVEY.md Appendix A
    assert n['Wld_mtarget'] == 2 and n['Wld_moptim'] == 0 and n['Rad_mrkind'] == 2 and n['Rad_nrad'] == 1
    assert n['Rad_the'] == 180.0 and n['Rad_phi'] == 270.0 and n['Rad_zloc'] == 705000.0
    assert n['Src_the'] == 150.0 and n['Src_phi'] == 225.0 and n['Src_qmax'] == 0.533133 and n['Src_flx'] == 1.0
    assert n['Sfc_mtype'] == 1 and n['Sfc_param(1)'] == 0.03 and list(n['Sfc_mbrdf']) == [1, 0, 0, 0]
    assert n['Atm_iz3l'] == 3 and n['Atm_nz3'] == 1 and n['Rad_nxr'] == 4 and n['Rad_nyr'] == 3       # iz3l quirk kept verbatim
    assert m.photons.shape == (48,) and m.photons_per_set == 60000 and m.Nx == 4 and m.Ny == 3
    assert m.fnames_out[2][15].endswith('r02.g015.out.bin')
    assert m.solver == '3D' and m.target == 'radiance' and not m.sfc_2d


def test_fused_weighting_equals_reference_style_weighting_of_raw_fields(inputs, tmp_path):
    fused = run(inputs, tmp_path)
    raw = run(inputs, tmp_path, raw=True, write_files=True)
    a = bmca.mca_out_ng(mca_obj=fused, abs_obj=inputs['abs0'], mode='mean', squeeze=True).data
    b = bmca.mca_out_ng(mca_obj=raw, abs_obj=inputs['abs0'], mode='mean', squeeze=True).data
    assert a['rad']['data'].shape == (4, 3) and a['rad']['data'].dtype == np.float32
    assert np.allclose(a['rad']['data'], b['rad']['data'], rtol=2e-6)
    assert np.allclose(a['rad_std']['data'], b['rad_std']['data'], rtol=1e-3, atol=1e-7)
    assert a['rad']['dims_info'] == ['Nx', 'Ny'] and a['rad']['units'] == 'W/m^2/nm/sr'
    # the compatibility files are what mca_out_raw parses; drop the in-memory copy and read them back
    raw.raw = None
    c = bmca.mca_out_ng(mca_obj=raw, abs_obj=inputs['abs0'], mode='mean', squeeze=True).data
    assert np.allclose(c['rad']['data'], b['rad']['data'], rtol=1e-6)
    assert os.path.isfile(raw.fnames_inp[0][0]) and 'Wld_jseed' in open(raw.fnames_inp[0][0]).read()
    # reflectance as downstream code computes it (er3t/rtm/mca/util.py:101) is physical
    refl = np.pi * a['rad']['data'].mean() / (a['toa']['data'] * np.cos(np.deg2rad(30.0)))
    assert 0.2 < refl < 0.7


def test_flux_and_heating_targets(inputs, tmp_path):
    m = run(inputs, tmp_path, target='heating rate', photons=4e4)
    d = bmca.mca_out_ng(mca_obj=m, abs_obj=inputs['abs0'], mode='mean', squeeze=True).data
    assert d['f_down']['data'].shape == (4, 3, 21) and d['absorbed']['data'].shape == (4, 3, 20)
    toa_down = d['f_down']['data'][:, :, -1].mean()
    assert abs(toa_down / (d['toa']['data'] * np.cos(np.deg2rad(30.0))) - 1.0) < 0.02
    assert np.allclose(d['f_down_diffuse']['data'], d['f_down']['data'] - d['f_down_direct']['data'], atol=1e-6)
    # column energy budget: net flux at TOA - net flux at the surface = absorbed in the atmosphere
    net_toa = (d['f_down']['data'][:, :, -1] - d['f_up']['data'][:, :, -1]).mean()
    net_sfc = (d['f_down']['data'][:, :, 0] - d['f_up']['data'][:, :, 0]).mean()
    assert abs((net_toa - net_sfc) - d['absorbed']['data'].sum(axis=-1).mean()) < 0.02 * net_toa
    a = bmca.mca_out_ng(mca_obj=m, abs_obj=inputs['abs0'], mode='all', squeeze=False).data
    assert a['f_up']['data'].shape == (4, 3, 21, 1, 3) and a['f_up']['dims_info'] == ['Nx', 'Ny', 'Nz', 'Nt', 'Nr']


def test_dump_and_load_roundtrip(inputs, tmp_path):
    m = run(inputs, tmp_path, photons=3e4)
    f = str(tmp_path / 'out.h5')
    a = bmca.mca_out_ng(fname=f, mca_obj=m, abs_obj=inputs['abs0'], mode='mean', squeeze=True, quiet=True)
    assert os.path.isfile(f)
    b = bmca.mca_out_ng(fname=f, mode='mean', quiet=True)           # reading mode: no objects needed (mca_out.py:160-162)
    assert np.array_equal(a.data['rad']['data'], b.data['rad']['data'])
    assert b.data['rad']['units'] == 'W/m^2/nm/sr'


def test_error_behaviour(inputs, tmp_path):
    with pytest.raises(OSError, match='Cannot understand <solver='):
        run(inputs, tmp_path, solver='4d')
    with pytest.raises(OSError, match='Cannot understand <target='):
        run(inputs, tmp_path, target='albedo')
    with pytest.raises(OSError, match='need <atm_1ds>'):
        run(inputs, tmp_path, atm_1ds=[])
    with pytest.raises(ValueError, match='Cannot ingest <surface_albedo>'):
        run(inputs, tmp_path, surface_albedo='bright')
    with pytest.raises(OSError, match='Please provide both'):
        bmca.mca_out_ng()
    with pytest.raises(OSError):
        bmca.mca_atm_1d(atm_obj=None, abs_obj=None)
    m = run(inputs, tmp_path, surface_albedo=1, Ncpu=1, photons=2e4)         # int albedo and Ncpu=1 are accepted (Appendix C)
    assert m.fused is not None


def test_2d_surface_object(inputs, tmp_path):
    alb = np.full((4, 3), 0.2)
    sfc = bmca.mca_sfc_2d(atm_obj=inputs['atm0'], sfc_obj=bpre.sfc_2d_gen(sfc_2d=alb), quiet=True)
    m = run(inputs, tmp_path, surface_albedo=sfc, photons=3e4)
    assert m.sfc_2d and m.scene.struct.sfc_nx == 4 and m.scene.struct.sfc_ny == 3
    m2 = run(inputs, tmp_path, surface_albedo=0.2, photons=3e4)
    a = bmca.mca_out_ng(mca_obj=m, abs_obj=inputs['abs0']).data['rad']['data'].mean()
    b = bmca.mca_out_ng(mca_obj=m2, abs_obj=inputs['abs0']).data['rad']['data'].mean()
    assert abs(a / b - 1.0) < 0.05


def test_all_sky_camera_through_the_public_api(inputs, tmp_path):
    # sensor_type='all-sky camera': Rad_mrkind = 1, qmax = 178, apsize = 0.05, 500 x 500 pixels (er3t/rtm/mca/mcarats.py:291-296,369-371)
    m = run(inputs, tmp_path, sensor_type='all-sky camera', sensor_zenith_angle=180.0, sensor_altitude=0.0, sensor_xpos=0.25, sensor_ypos=0.75,
            camera_pixels=(10, 8), photons=3e4)
    n = m.nml[0]
    assert n['Rad_mrkind'] == 1 and n['Rad_qmax'] == 178.0 and n['Rad_apsize'] == 0.05 and n['Rad_xpos'] == 0.25 and n['Rad_ypos'] == 0.75
    assert n['Rad_the'] == 0.0 and n['Rad_zloc'] == 0.0 and n['Rad_nxr'] == 10 and n['Rad_nyr'] == 8
    se = m.scene.sensors[0]
    assert se.kind == 1 and (se.nxr, se.nyr) == (10, 8) and se.qmax == 178.0 and se.xpos == 0.25
    out = bmca.mca_out_ng(mca_obj=m, abs_obj=inputs['abs0'], mode='mean', squeeze=True).data
    assert out['rad']['data'].shape == (10, 8) and np.all(np.isfinite(out['rad']['data'])) and out['rad']['data'].max() > 0.0
    # without the override the reference's fixed 500 x 500 grid is what the namelist carries
    m2 = run(inputs, tmp_path, sensor_type='all-sky camera', sensor_zenith_angle=180.0, sensor_altitude=0.0, dry_run=True)
    assert m2.nml[0]['Rad_nxr'] == 500 and m2.nml[0]['Rad_nyr'] == 500


def test_hdf5_dump_matches_the_reference_call_for_call(monkeypatch, tmp_path):
    """h5py is absent from the image, so the HDF5 branch of mca_out_ng.dump / load is exercised against a recording
    stand-in (tests/fake_h5py.py).  The expected log was recorded from the REFERENCE's own dump (mca_out.py:209-233) by
    tests/golden/make_h5_golden.py: same group, same dataset options (gzip 9, chunked), same attributes and types."""
    import json
    import os
    import sys
    import fake_h5py
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), 'golden'))
    from make_h5_golden import sample_data
    from er3t_b200.rtm.mca.mca_out import mca_out_ng
    monkeypatch.setitem(sys.modules, 'h5py', fake_h5py)

    class _M:
        target = 'radiance'

    o = object.__new__(mca_out_ng)
    o.data, o.fname, o.mode, o.quiet, o.verbose, o.mca = sample_data(), str(tmp_path / 'out.h5'), 'mean', True, False, _M()
    fake_h5py.reset()
    o.dump()
    want = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'h5_dump_calls.json')))
    got = json.loads(json.dumps(fake_h5py.LOG))
    assert got == want
    # and the reader gets the same dictionary back (HDF5 branch of load: the file does not start with the zip magic)
    with open(o.fname, 'wb') as f:
        f.write(b'\x89HDF\r\n\x1a\n')
    fake_h5py.FILES[o.fname] = fake_h5py.FILES.pop(o.fname) if o.fname in fake_h5py.FILES else None
    r = object.__new__(mca_out_ng)
    r.fname, r.mode, r.quiet, r.verbose = o.fname, 'mean', True, False
    r.load()
    for key, item in o.data.items():
        assert np.array_equal(np.asarray(r.data[key]['data']), np.asarray(item['data']))
        assert r.data[key]['name'] == item['name'] and r.data[key]['units'] == item['units']


def test_device_props_hands_over_views_and_lazy_fields_equal_the_host_interpolation():
    """mca_atm_3d(device_props=True): extinction and effective radius are handed on as VIEWS of the cloud object's own
    float32 C-ordered arrays (no host copy; other dtypes / layouts are converted once), omega / apf stay available as lazily
    evaluated arrays that equal what the default path (the restated er3t/rtm/mca/mca_atm.py:291-303) computes, and the scene
    carries cer3d + the (ref, ssa, asy) tables instead of omg3d / apf3d."""
    from er3t_b200.util import pin_array
    atm0 = bpre.atm_atmmod(levels=np.linspace(0, 20, 21))
    abs0 = bpre.abs_16g(wavelength=650.0, atm_obj=atm0)
    pha0 = bpre.pha_mie_wc(wavelength=650.0, reff=[5.0, 10.0, 15.0], nr=48)
    cld0 = bpre.cld_gen_les(Nx=12, Ny=10, dx=0.1, dy=0.1, altitude=np.array([1.5, 2.5]), seed=2, atm_obj=atm0)
    ref = bmca.mca_atm_3d(cld_obj=cld0, atm_obj=atm0, pha_obj=pha0, quiet=True)
    for key in ('extinction', 'cer'):
        cld0.lay[key]['data'] = pin_array(np.asarray(cld0.lay[key]['data'], dtype=np.float32))     # plain memory without a GPU
    dev = bmca.mca_atm_3d(cld_obj=cld0, atm_obj=atm0, pha_obj=pha0, quiet=True, device_props=True)
    assert np.shares_memory(dev.nml['Atm_extp3d']['data'], cld0.lay['extinction']['data'])
    assert np.shares_memory(dev.cer3d, cld0.lay['cer']['data'])
    for key in ('Atm_extp3d', 'Atm_omgp3d', 'Atm_apfp3d', 'Atm_abst3d', 'Atm_tmpa3d'):
        assert np.array_equal(np.asarray(dev.nml[key]['data']), np.asarray(ref.nml[key]['data'])), key
    # a float64 / Fortran-ordered field is converted once instead of viewed
    cld0.lay['extinction']['data'] = np.asfortranarray(cld0.lay['extinction']['data'].astype(np.float64))
    dev2 = bmca.mca_atm_3d(cld_obj=cld0, atm_obj=atm0, pha_obj=pha0, quiet=True, device_props=True)
    assert not np.shares_memory(dev2.nml['Atm_extp3d']['data'], cld0.lay['extinction']['data'])
    assert dev2.nml['Atm_extp3d']['data'].dtype == np.float32 and np.array_equal(dev2.nml['Atm_extp3d']['data'], ref.nml['Atm_extp3d']['data'])
    # scene hand-off
    m = bmca.mcarats_ng(date=datetime.datetime(2017, 8, 13), atm_1ds=[bmca.mca_atm_1d(atm_obj=atm0, abs_obj=abs0)], atm_3ds=[dev], Ng=16,
                        target='radiance', surface_albedo=0.03, sca=bmca.mca_sca(pha_obj=pha0), solar_zenith_angle=30.0,
                        solar_azimuth_angle=45.0, fdir='tmp-data/devprops', Nrun=1, photons=1e4, weights=abs0.coef['weight']['data'],
                        solver='3D', quiet=True, seed=1, dry_run=True)
    st = m.scene.struct
    assert st.cer3d and st.nref == 3 and not st.omg3d and not st.apf3d and st.layout3d == 1
