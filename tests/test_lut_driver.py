"""Batched IPA look-up-table driver (er3t_b200.rtm.mca.func_ref_vs_cot*) on the CPU through the
oracle-backed test double: one launch for all COT values must reproduce what the reference computes value by value
(er3t/rtm/mca/util.py:105-195) -- a plane-parallel cloud per COT."""

import datetime
import os

import numpy as np
import pytest

import er3t_b200.pre as bpre
from er3t_b200.rtm import mca as bmca
from oracle_solver import OracleSolver


@pytest.fixture(scope='module')
def common():
    atm0 = bpre.atm_atmmod(levels=np.linspace(0, 20, 21))
    abs0 = bpre.abs_16g(wavelength=650.0, atm_obj=atm0)
    pha0 = bpre.pha_mie_wc(wavelength=650.0, reff=[5.0, 10.0, 15.0], nr=48)
    return dict(atm0=atm0, abs0=abs0, pha0=pha0)


def test_batched_lut_equals_per_cot_plane_parallel_runs(common, tmp_path):
    cot = np.array([0.0, 2.0, 10.0, 40.0])
    f = bmca.func_ref_vs_cot(cot, cer0=10.0, fdir=str(tmp_path / 'lut'), date=datetime.datetime(2017, 8, 13), wavelength=650.0,
                             surface_albedo=0.03, solar_zenith_angle=30.0, Nphoton=4e4, atm0=common['atm0'], overwrite=True,
                             seed=11, solver_obj=OracleSolver(), pha0=common['pha0'], abs0=common['abs0'])
    assert f.ref.shape == (4,) and f.ref_std.shape == (4,) and f.ref_2s.shape == (4,)
    assert np.all(np.diff(f.ref) > 0)                         # reflectance grows with optical thickness
    assert f.mca.solver == 'IPA' and f.mca.Nx == 4           # one launch, one column per COT
    # per-COT files with the reference's names and keys
    for c in cot:
        assert os.path.exists('%s/er3t_cot-%05.1f_cer-10.0.h5' % (str(tmp_path / 'lut'), c))
    # the reference's way: one 1-D run per COT with the tabulated phase function (er3t/rtm/mca/util.py:144-181)
    sca0 = bmca.mca_sca(pha_obj=common['pha0'])
    iref = int(np.argmin(np.abs(common['pha0'].data['ref']['data'] - 10.0)))
    for i in (1, 2):
        atm1d0 = bmca.mca_atm_1d(atm_obj=common['atm0'], abs_obj=common['abs0'])
        atm1d0.add_mca_1d_atm(ext1d=cot[i] / 1000.0, omg1d=common['pha0'].data['ssa']['data'][iref], apf1d=iref + 1, z_bottom=1.0, z_top=2.0)
        m = bmca.mcarats_ng(date=datetime.datetime(2017, 8, 13), atm_1ds=[atm1d0], atm_3ds=[], sca=sca0, target='radiance',
                            surface_albedo=0.03, solar_zenith_angle=30.0, solar_azimuth_angle=0.0, fdir=str(tmp_path / 'one'), Nrun=3, Ng=16,
                            weights=common['abs0'].coef['weight']['data'], photons=4e4, solver='3d', quiet=True, seed=5,
                            solver_obj=OracleSolver())
        o = bmca.mca_out_ng(mca_obj=m, abs_obj=common['abs0'], mode='mean', squeeze=True, quiet=True)
        ref1 = np.pi * float(np.mean(o.data['rad']['data'])) / (o.data['toa']['data'] * f.mu0)
        sd = np.hypot(f.ref_std[i], np.pi * float(np.mean(o.data['rad_std']['data'])) / (o.data['toa']['data'] * f.mu0)) / np.sqrt(3.0)
        assert abs(f.ref[i] - ref1) < 4.0 * sd + 0.003, (cot[i], f.ref[i], ref1, sd)
    # reload without running (overwrite=False reads the files) and use the table
    g = bmca.func_ref_vs_cot(cot, cer0=10.0, fdir=str(tmp_path / 'lut'), solar_zenith_angle=30.0, overwrite=False, solver_obj=None)
    assert np.allclose(g.ref, f.ref)
    assert abs(float(g.get_cot_from_ref(g.ref[2], method='linear')) - 10.0) < 1e-6
    assert abs(float(g.get_ref_from_cot(10.0, method='linear')) - g.ref[2]) < 1e-9
    assert abs(float(g.get_cot_from_ref(g.ref_2s[1], method='linear', mode='2s')) - 2.0) < 1e-6


def test_multi_pixel_variant(common):
    cot = np.array([1.0, 8.0])
    f = bmca.func_ref_vs_cot_multi_pixel(cot, cer0=10.0, fdir=None, date=datetime.datetime(2017, 8, 13), solar_zenith_angle=30.0,
                                         Nphoton=2e4, Nx=2, Ny=2, atm0=common['atm0'], seed=3, solver_obj=OracleSolver(),
                                         pha0=common['pha0'], abs0=common['abs0'])
    assert f.mca.Nx == 4 and f.mca.Ny == 2 and f.ref.shape == (2,) and f.ref[1] > f.ref[0] > 0.0
    # two-stream sanity cross-check, the reference's own loose check (er3t/rtm/mca/util.py:66)
    assert abs(f.ref[1] - f.ref_2s[1]) < 0.15


def test_oracle_on_the_reference_benchmark_curve():
    """Pins the ORACLE on the reference's benchmark geometry (examples/00_er3t_bmk.py:24-46,470-579): reflectance against COT
    from func_ref_vs_cot with the oracle double against the deterministic adding-doubling curve
    (tests/golden/ad_cot_sweep.npz); the GPU takes the same test at all 35 COT values and 2e6 photons each
    (tests/test_gpu_deterministic.py)."""
    import sys
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
    sys.path.insert(0, gdir)
    import make_cot_sweep as mk
    fx = np.load(os.path.join(gdir, 'ad_cot_sweep.npz'))
    atm0, abs0, _ = mk.inputs()
    pha0 = bpre.pha_mie_wc(wavelength=650.0, reff=[5.0, 10.0, 15.0], nr=96)
    pick = [0, 4, 9, 19, 28]                                  # COT 0, 2, 10, 30, 100
    nrun = 4
    f = bmca.func_ref_vs_cot(fx['cot'][pick], cer0=10.0, fdir=None, date=mk.DATE, wavelength=650.0, surface_albedo=float(fx['albedo']),
                             solar_zenith_angle=float(fx['sza']), Nphoton=6e4, atm0=atm0, seed=21, solver_obj=OracleSolver(), pha0=pha0,
                             abs0=abs0, Nrun=nrun)
    sem = f.ref_std / np.sqrt(nrun)
    dev = np.abs(f.ref - fx['ref'][pick])
    assert np.all(dev < 4.0 * sem + 3.0 * fx['ref_conv'][pick] + 2e-3 * fx['ref'][pick]), (f.ref, fx['ref'][pick], sem)
