"""
Parity tests proper: the CUDA path (through the C-ABI) against oracle/ on the same seeded scenes.

Tolerances are the ones BASELINE.json's north_star states:
  * energy conservation R + T + A = 1 to fp64-accumulation precision,
  * domain-mean fluxes within 0.5 % of the oracle,
  * per-pixel radiance within 3 combined Monte Carlo standard deviations.
Monte Carlo on two different random streams cannot agree bit for bit; the integer parts of the path (Philox,
photon -> job mapping) are checked bit-exactly in test_gpu_bits.py.
"""

import numpy as np
import pytest

import oracle
import scenes
from er3t_b200 import abi

pytestmark = pytest.mark.gpu

FLUX_RTOL = 0.005      # north_star: domain-mean fluxes within 0.5 %
NSIGMA = 3.0           # north_star: per-pixel radiance within 3 combined sigma


def run_both(solver, sc, opt, jobs, jobs_gpu=None):
    solver.upload_scene(sc, opt)
    solver.run(jobs if jobs_gpu is None else jobs_gpu)
    g = solver.results()
    c = oracle.run(sc, opt, jobs)
    return g, c


def close(gm, cm, z, rtol=FLUX_RTOL, nsig=3.5):
    """within the stated relative tolerance, or -- when the Monte Carlo noise of the comparison itself is larger
    than that -- within nsig combined standard errors."""
    return abs(gm / cm - 1.0) < rtol or abs(z) < nsig


def check_energy(st, tol=1e-9):
    n = st['photons']
    bal = (st['w_toa_up'] + st['w_sfc_abs'] + st['w_atm_abs'] - st['w_roulette']) / n - 1.0
    assert abs(bal) < tol, bal


def mean_close(g, c, nslab):
    """domain-mean of a field: within 0.5 %, or within 3.5 combined standard errors of the comparison itself"""
    gm, gs = scenes.mean_sem(np.asarray(g).reshape(nslab, -1).mean(axis=1))
    cm, cs = scenes.mean_sem(np.asarray(c).reshape(nslab, -1).mean(axis=1))
    return close(gm, cm, (gm - cm) / np.hypot(gs, cs))


def zscores(g, c, nslab, shape):
    gm, gs = scenes.mean_sem(g.reshape((nslab,) + shape))
    cm, cs = scenes.mean_sem(c.reshape((nslab,) + shape))
    # fp32-vs-fp64 rounding floor: deterministic tallies (e.g. TOA down-flux) have zero Monte Carlo spread
    den = np.sqrt(gs ** 2 + cs ** 2) + 2e-6 * np.abs(cm)
    ok = den > 0
    z = np.zeros_like(gm)
    z[ok] = (gm[ok] - cm[ok]) / den[ok]
    return z, gm, cm


def assert_pixels(z, nslab):
    """|z| <= 3 sigma per pixel, allowing the tail mass a Student-t with nslab-1 dof gives on many pixels."""
    frac = np.mean(np.abs(z) > NSIGMA)
    assert frac <= 0.03, frac
    assert np.max(np.abs(z)) < 6.5, np.max(np.abs(z))
    rms = np.sqrt(np.mean(z ** 2))
    assert 0.5 < rms < 1.6, rms


@pytest.mark.parametrize('absorb', [False, True])
def test_plane_parallel_flux_and_radiance(solver, absorb):
    sc, absg = scenes.plane_parallel(absorb=absorb)
    nslab = 8
    opt = abi.make_options(target=abi.TARGET_FLUX | abi.TARGET_RADIANCE, nslab=nslab, wmin=0.0)
    jobs, keep = scenes.multi_seed_jobs(250000, nslab, abs1d=absg)
    jobs_gpu, keep2 = scenes.multi_seed_jobs(2000000, nslab, abs1d=absg)
    g, c = run_both(solver, sc, opt, jobs, jobs_gpu)
    check_energy(g['stats'])
    check_energy(c['stats'])
    nlev = sc.struct.nz + 1
    z, gm, cm = zscores(g['flux'], c['flux'], nslab, (3, nlev, 1, 1))
    # domain-mean flux at every level, all three components
    big = cm > 1e-3
    assert np.all(np.abs(gm[big] / cm[big] - 1.0) < FLUX_RTOL), np.max(np.abs(gm[big] / cm[big] - 1.0))
    assert np.max(np.abs(z)) < 5.0
    zr, grm, crm = zscores(g['rad'], c['rad'], nslab, (1,))
    assert abs(zr[0]) < 4.0
    assert close(grm[0], crm[0], zr[0]), (grm[0], crm[0], zr[0])


def test_plane_parallel_oblique_view_and_reciprocity(solver):
    """Oblique parallel-projection sensor over a plane-parallel atmosphere (no 3-D block: the local-estimate optical depth is
    the closed-form difference of cumulative profiles): against the oracle, and Helmholtz reciprocity
    I(sun a -> view b) / mu_a = I(sun b -> view a) / mu_b on the GPU alone."""
    za, zb, dphi = 30.0, 55.0, 40.0
    nslab = 8
    out = []
    for sza, vza in ((za, zb), (zb, za)):
        sc0, absg = scenes.plane_parallel(sza=sza, cot=2.0, g=0.7, omega=0.98, albedo=0.2, absorb=True, with_sensor=False, qmax=0.0)
        sc = abi.HostScene(sc0.zgrd, sc0.ext1d, sc0.omg1d, sc0.apf1d, sfc_type=1, sfc_param=(0.2, 0, 0, 0, 0), src_the=180.0 - sza,
                           src_phi=270.0, src_qmax=0.0, sensors=[dict(the=180.0 - vza, phi=270.0 + dphi, nxr=1, nyr=1)])
        opt = abi.make_options(target=abi.TARGET_RADIANCE, nslab=nslab, wmin=0.2)
        jobs, keep = scenes.multi_seed_jobs(200000, nslab, abs1d=absg)
        g, c = run_both(solver, sc, opt, jobs)
        check_energy(g['stats'])
        assert mean_close(g['rad'], c['rad'], nslab)
        m, se = scenes.mean_sem(g['rad'].reshape(nslab))
        out.append((m / np.cos(np.deg2rad(sza)), se / np.cos(np.deg2rad(sza))))
    (a, sa), (b, sb) = out
    assert abs(a - b) < 4.0 * np.hypot(sa, sb) + 0.005 * a, (a, b, sa, sb)


def test_plane_parallel_roulette_unbiased(solver):
    sc, absg = scenes.plane_parallel(omega=0.9, albedo=0.5)
    nslab = 8
    o0 = abi.make_options(target=abi.TARGET_FLUX, nslab=nslab, wmin=0.0)
    o1 = abi.make_options(target=abi.TARGET_FLUX, nslab=nslab, wmin=0.2, wfac=1.0)
    jobs, keep = scenes.multi_seed_jobs(100000, nslab)
    solver.upload_scene(sc, o0); solver.run(jobs); a = solver.results()
    solver.upload_scene(sc, o1); solver.run(jobs); b = solver.results()
    check_energy(a['stats']); check_energy(b['stats'])
    assert b['stats']['n_roulette_kill'] > 0
    z, am, bm = zscores(a['flux'], b['flux'], nslab, (3, sc.struct.nz + 1, 1, 1))
    assert np.max(np.abs(z)) < 5.0
    assert abs(bm[2, -1, 0, 0] / am[2, -1, 0, 0] - 1.0) < FLUX_RTOL


@pytest.mark.parametrize('sfc', ['lambert', 'lambert2d', 'lsrt', 'dsm'])
def test_3d_radiance_nadir(solver, sfc):
    sc = scenes.scene_3d(sfc=sfc)
    nslab = 8
    opt = abi.make_options(target=abi.TARGET_RADIANCE | abi.TARGET_FLUX, nslab=nslab, wmin=0.2)
    jobs, keep = scenes.multi_seed_jobs(200000, nslab)
    g, c = run_both(solver, sc, opt, jobs)
    check_energy(g['stats'])
    nx, ny = sc.struct.nx, sc.struct.ny
    z, gm, cm = zscores(g['rad'], c['rad'], nslab, (ny, nx))
    assert_pixels(z, nslab)
    assert mean_close(g['rad'], c['rad'], nslab)
    # domain-mean fluxes at TOA and surface
    gf = g['flux'].reshape(nslab, 3, -1, ny, nx).mean(axis=(0, 3, 4))
    cf = c['flux'].reshape(nslab, 3, -1, ny, nx).mean(axis=(0, 3, 4))
    for var, lev in ((2, -1), (1, 0), (0, 0), (2, 0)):
        assert abs(gf[var, lev] / cf[var, lev] - 1.0) < FLUX_RTOL, (var, lev, gf[var, lev], cf[var, lev])


@pytest.mark.parametrize('sv,cm,K', [((1, 1, 1), (1, 1, 1), 1), ((4, 4, 2), (1, 1, 1), 4), ((16, 12, 4), (2, 2, 2), 8),
                                     ((1, 1, 1), (4, 4, 4), 16), ((2, 2, 1), (2, 4, 2), 3), ((3, 5, 2), (4, 2, 1), 8)])
def test_3d_supervoxel_sizes_agree_with_exact_traversal(solver, sv, cm, K):
    """Null-collision tracking on any two-level majorant grid (fine cells + empty-space coarse cells) and any flight
    length must reproduce the oracle's exact traversal."""
    sc = scenes.scene_3d(sensors=[dict(the=180.0, phi=270.0, nxr=16, nyr=12)])
    nslab = 8
    opt = abi.make_options(target=abi.TARGET_RADIANCE, nslab=nslab, wmin=0.2, sv=sv, cm=cm, flight_steps=K)
    jobs, keep = scenes.multi_seed_jobs(200000, nslab)
    g, c = run_both(solver, sc, opt, jobs)
    z, gm, cm = zscores(g['rad'], c['rad'], nslab, (12, 16))
    assert_pixels(z, nslab)
    assert mean_close(g['rad'], c['rad'], nslab)


@pytest.mark.parametrize('sv,cm,runs,uniform', [((1, 1, 1), (1, 1, 1), 0, True), ((2, 2, 1), (2, 2, 2), 0, True), ((2, 2, 3), (4, 4, 1), 0, True),
                                                ((1, 1, 2), (2, 2, 2), -1, True), ((2, 2, 1), (2, 2, 2), 0, False), ((1, 1, 1), (1, 1, 3), 0, False)])
def test_3d_empty_runs_tall_block(solver, sv, cm, runs, uniform):
    """A 10-layer 3-D block with ragged cloud tops: boxes that merge vertical runs of empty coarse cells (3-D layers of
    equal thickness, `empty_runs` = 0) against the one-group boxes (`empty_runs` = -1, or layers of unequal thickness,
    which fall back to the layer search) -- all must reproduce the oracle's exact traversal."""
    dz3 = 100.0 if uniform else np.array([60.0, 140.0, 100.0, 80.0, 120.0, 100.0, 90.0, 110.0, 100.0, 100.0])
    sc = scenes.scene_3d(nz3=10, dz3=dz3, nlay=18, sza=50.0, sensors=[dict(the=180.0, phi=270.0, nxr=16, nyr=12)])
    nslab = 8
    opt = abi.make_options(target=abi.TARGET_RADIANCE, nslab=nslab, wmin=0.2, sv=sv, cm=cm, empty_runs=runs)
    jobs, keep = scenes.multi_seed_jobs(200000, nslab)
    g, c = run_both(solver, sc, opt, jobs)
    check_energy(g['stats'])
    z, gm, cm_ = zscores(g['rad'], c['rad'], nslab, (12, 16))
    assert_pixels(z, nslab)
    assert mean_close(g['rad'], c['rad'], nslab)


@pytest.mark.parametrize('nx,ny,nz3', [(1, 1, 3), (3, 5, 4), (7, 2, 1)])
def test_3d_ragged_and_degenerate_grids(solver, nx, ny, nz3):
    """Grids the automatic cell sizes do not divide (odd column counts, one column, one 3-D layer): flux and nadir
    radiance against the oracle."""
    sc = scenes.scene_3d(nx=nx, ny=ny, nz3=nz3, sensors=[dict(the=180.0, phi=270.0, nxr=nx, nyr=ny)], seed=5)
    nslab = 8
    opt = abi.make_options(target=abi.TARGET_RADIANCE | abi.TARGET_FLUX, nslab=nslab, wmin=0.2)
    jobs, keep = scenes.multi_seed_jobs(100000, nslab)
    g, c = run_both(solver, sc, opt, jobs)
    check_energy(g['stats'])
    z, gm, cm = zscores(g['rad'], c['rad'], nslab, (ny, nx))
    assert np.max(np.abs(z)) < 5.0
    assert mean_close(g['rad'], c['rad'], nslab)
    gf = g['flux'].reshape(nslab, 3, -1, ny, nx).mean(axis=(0, 3, 4))
    cf = c['flux'].reshape(nslab, 3, -1, ny, nx).mean(axis=(0, 3, 4))
    for var, lev in ((2, -1), (1, 0)):
        assert abs(gf[var, lev] / cf[var, lev] - 1.0) < 2 * FLUX_RTOL, (var, lev, gf[var, lev], cf[var, lev])


def test_empty_atmosphere_and_zero_photons(solver):
    """No extinction at all over a black surface: every photon reaches the ground unscattered (direct = total down-flux =
    mu0 at every level, no up-flux, T = 1); and a job list without photons runs and leaves the tallies at zero."""
    z = scenes.std_z()
    nz = z.size - 1
    sc = abi.HostScene(z, np.zeros((1, nz)), np.ones((1, nz)), -np.ones((1, nz)), sfc_type=1, sfc_param=(0.0, 0, 0, 0, 0),
                       src_the=180.0 - 60.0, src_phi=270.0, src_qmax=0.0, sensors=[dict(the=180.0, phi=270.0, nxr=1, nyr=1)])
    opt = abi.make_options(target=abi.TARGET_RADIANCE | abi.TARGET_FLUX, nslab=1, wmin=0.2)
    jobs, keep = abi.make_jobs([50000], [3], [0])
    solver.upload_scene(sc, opt); solver.run(jobs)
    r = solver.results()
    st = r['stats']
    assert st['n_coll'] == 0 and st['n_sfc'] == 50000
    assert abs(st['w_sfc_abs'] / st['photons'] - 1.0) < 1e-12 and st['w_toa_up'] == 0.0
    f = r['flux'].reshape(3, nz + 1)
    assert np.allclose(f[0], 0.5, rtol=1e-6) and np.allclose(f[1], 0.5, rtol=1e-6) and np.all(f[2] == 0.0)
    assert np.all(r['rad'] == 0.0)
    jobs0, keep0 = abi.make_jobs([0, 0], [1, 2], [0, 0])
    solver.run(jobs0)
    r0 = solver.results()
    assert r0['stats']['photons'] == 0 and np.all(r0['flux'] == 0.0) and np.all(r0['rad'] == 0.0)


def test_3d_oblique_multi_sensor(solver):
    sens = [dict(the=180.0, phi=270.0, nxr=16, nyr=12),
            dict(the=180.0 - 35.0, phi=30.0, nxr=16, nyr=12),
            dict(the=180.0 - 60.0, phi=250.0, nxr=8, nyr=6),
            dict(the=180.0 - 20.0, phi=100.0, nxr=16, nyr=12, zloc=900.0)]    # inside the cloud layer block
    sc = scenes.scene_3d(sensors=sens, sfc='lsrt')
    nslab = 8
    opt = abi.make_options(target=abi.TARGET_RADIANCE, nslab=nslab, wmin=0.2)
    jobs, keep = scenes.multi_seed_jobs(150000, nslab)
    g, c = run_both(solver, sc, opt, jobs)
    per = [16 * 12, 16 * 12, 8 * 6, 16 * 12]
    gr = g['rad'].reshape(nslab, -1); cr = c['rad'].reshape(nslab, -1)
    off = 0
    for n in per:
        z, gm, cm = zscores(gr[:, off:off + n].copy(), cr[:, off:off + n].copy(), nslab, (n,))
        assert_pixels(z, nslab)
        assert mean_close(gr[:, off:off + n], cr[:, off:off + n], nslab)
        off += n


@pytest.mark.parametrize('solver_mode', [abi.SOLVER_IPA, abi.SOLVER_PARTIAL_3D])
def test_3d_ipa_and_partial(solver, solver_mode):
    sc = scenes.scene_3d()
    nslab = 8
    opt = abi.make_options(solver=solver_mode, target=abi.TARGET_RADIANCE | abi.TARGET_FLUX, nslab=nslab, wmin=0.2)
    jobs, keep = scenes.multi_seed_jobs(150000, nslab)
    g, c = run_both(solver, sc, opt, jobs)
    z, gm, cm = zscores(g['rad'], c['rad'], nslab, (12, 16))
    assert_pixels(z, nslab)
    assert mean_close(g['rad'], c['rad'], nslab)


def test_3d_two_components_table_phase_heating(solver):
    sc = scenes.scene_3d(two_comp=True, apf_mode='table')
    nslab = 8
    absg = np.full(sc.struct.nz, 2.0e-5)
    opt = abi.make_options(target=abi.TARGET_RADIANCE | abi.TARGET_FLUX | abi.TARGET_HEATING, nslab=nslab, wmin=0.2)
    jobs, keep = scenes.multi_seed_jobs(150000, nslab, abs1d=absg)
    g, c = run_both(solver, sc, opt, jobs)
    check_energy(g['stats'])
    z, gm, cm = zscores(g['rad'], c['rad'], nslab, (12, 16))
    assert_pixels(z, nslab)
    assert mean_close(g['rad'], c['rad'], nslab)
    gh = g['heat'].reshape(nslab, sc.struct.nz, -1).mean(axis=(0, 2))
    ch = c['heat'].reshape(nslab, sc.struct.nz, -1).mean(axis=(0, 2))
    assert np.all(np.abs(gh / ch - 1.0) < 0.01), np.max(np.abs(gh / ch - 1.0))
    # absorbed energy in the heating tally equals the atmospheric absorption counter
    n = g['stats']['photons'] / nslab
    mu0 = np.cos(np.deg2rad(40.0))
    assert abs(g['heat'].reshape(nslab, -1).sum(axis=1).mean() / (12 * 16) / mu0 - g['stats']['w_atm_abs'] / nslab / n) < 1e-6


def test_cyclic_shift_invariance(solver):
    """3-D result is invariant under a cyclic shift of the field (SURVEY.md 8c identity)."""
    sc = scenes.scene_3d(sfc='lambert')
    nslab = 8
    opt = abi.make_options(target=abi.TARGET_RADIANCE, nslab=nslab, wmin=0.2)
    jobs, keep = scenes.multi_seed_jobs(200000, nslab)
    solver.upload_scene(sc, opt); solver.run(jobs); a = solver.read_rad().reshape(nslab, 12, 16)
    ext, omg, apf = sc.ext3d, sc.omg3d, sc.apf3d          # (nx, ny, nz3, np3d)
    sh = (5, 3)
    sc2 = abi.HostScene(sc.zgrd, sc.ext1d, sc.omg1d, sc.apf1d, nx=16, ny=12, dx=100.0, dy=100.0, iz3l=sc.struct.iz3l,
                        ext3d=np.roll(ext, sh, axis=(0, 1)), omg3d=np.roll(omg, sh, axis=(0, 1)), apf3d=np.roll(apf, sh, axis=(0, 1)),
                        sfc_type=1, sfc_param=(0.1, 0, 0, 0, 0), src_the=sc.struct.src_the, src_phi=sc.struct.src_phi,
                        sensors=[dict(the=180.0, phi=270.0, nxr=16, nyr=12)])
    jobs2, keep2 = scenes.multi_seed_jobs(200000, nslab, seed0=555)
    solver.upload_scene(sc2, opt); solver.run(jobs2); b = solver.read_rad().reshape(nslab, 12, 16)
    b = np.roll(b, (-sh[1], -sh[0]), axis=(1, 2))
    z, am, bm = zscores(a, b, nslab, (12, 16))
    assert_pixels(z, nslab)


def test_3d_all_sky_camera(solver):
    """Rad_mrkind = 1 (all-sky camera at the ground, looking up) next to a nadir satellite view in one launch: the point
    detector's pixels are heavy-tailed, so the camera is compared through its mean over sky regions, the satellite
    view pixel by pixel."""
    npx = 12
    sens = [dict(kind=1, the=0.0, phi=0.0, psi=30.0, zloc=0.0, nxr=npx, nyr=npx, xpos=0.4, ypos=0.6, qmax=150.0, umax=160.0, vmax=160.0, apsize=0.05),
            dict(kind=2, the=180.0, phi=270.0, nxr=16, nyr=12)]
    sc = scenes.scene_3d(sensors=sens, clear_below=True)
    nslab = 8
    opt = abi.make_options(target=abi.TARGET_RADIANCE, nslab=nslab, wmin=0.2)
    jobs, keep = scenes.multi_seed_jobs(400000, nslab)
    g, c = run_both(solver, sc, opt, jobs)
    check_energy(g['stats'])
    gr, cr = g['rad'].reshape(nslab, -1), c['rad'].reshape(nslab, -1)
    ncam = npx * npx
    z, gm, cm = zscores(gr[:, ncam:], cr[:, ncam:], nslab, (12, 16))
    assert_pixels(z, nslab)
    gcam, ccam = gr[:, :ncam].reshape(nslab, npx, npx), cr[:, :ncam].reshape(nslab, npx, npx)
    assert np.all(np.isfinite(gcam)) and gcam.min() >= 0.0
    # pixels outside the field-of-view cone stay empty on both sides
    assert np.array_equal(gcam.sum(axis=0) == 0.0, ccam.sum(axis=0) == 0.0)
    for sel in (np.s_[:, :], np.s_[: npx // 2, :], np.s_[npx // 2:, :], np.s_[:, : npx // 2], np.s_[:, npx // 2:]):
        a, b = gcam[(slice(None),) + sel].mean(axis=(1, 2)), ccam[(slice(None),) + sel].mean(axis=(1, 2))
        am, asem = scenes.mean_sem(a); bm, bsem = scenes.mean_sem(b)
        assert abs(am / bm - 1.0) < 0.03 or abs(am - bm) < 3.5 * np.hypot(asem, bsem), (am, bm, asem, bsem)


def test_3d_absorption_field_abst3d(solver):
    """Atm_abst3d != 0 (er3t/rtm/mca/mca_inp.py:232-235, mca_atm.py:249): the CUDA path treats the 3-D absorption field as one
    more component with omega = 0 (absorption at collisions), the oracle integrates it along the path -- same expectation."""
    sc0 = scenes.scene_3d()
    rng = np.random.default_rng(11)
    a3 = (3.0e-4 * rng.random((16, 12, 4))).astype(np.float32)
    a3[rng.random((16, 12, 4)) < 0.4] = 0.0
    sc = abi.HostScene(sc0.zgrd, sc0.ext1d, sc0.omg1d, sc0.apf1d, nx=16, ny=12, dx=100.0, dy=100.0, iz3l=sc0.struct.iz3l,
                       ext3d=sc0.ext3d, omg3d=sc0.omg3d, apf3d=sc0.apf3d, abs3d=a3, sfc_type=1, sfc_param=(0.1, 0, 0, 0, 0),
                       src_the=sc0.struct.src_the, src_phi=sc0.struct.src_phi, sensors=[dict(the=180.0, phi=270.0, nxr=16, nyr=12)])
    nslab = 8
    opt = abi.make_options(target=abi.TARGET_RADIANCE | abi.TARGET_FLUX | abi.TARGET_HEATING, nslab=nslab, wmin=0.2)
    jobs, keep = scenes.multi_seed_jobs(200000, nslab)
    g, c = run_both(solver, sc, opt, jobs)
    check_energy(g['stats'])
    n = g['stats']['photons']
    ag, ac = g['stats']['w_atm_abs'] / n, c['stats']['w_atm_abs'] / c['stats']['photons']
    assert ag > 0.01                                                     # the field does absorb
    assert abs(ag / ac - 1.0) < 0.01, (ag, ac)
    z, gm, cm = zscores(g['rad'], c['rad'], nslab, (12, 16))
    assert_pixels(z, nslab)
    assert mean_close(g['rad'], c['rad'], nslab)
    for var in range(3):
        assert mean_close(g['flux'][:, var], c['flux'][:, var], nslab)
    gh = g['heat'].reshape(nslab, sc.struct.nz, -1).mean(axis=(0, 2))
    ch = c['heat'].reshape(nslab, sc.struct.nz, -1).mean(axis=(0, 2))
    lay = ch > 1e-4 * ch.max()
    assert np.all(np.abs(gh[lay] / ch[lay] - 1.0) < 0.02), np.max(np.abs(gh[lay] / ch[lay] - 1.0))
    with pytest.raises(OSError, match='Atm_abst3d must be finite'):
        bad = a3.copy(); bad[0, 0, 0] = -1e-4
        solver.upload_scene(abi.HostScene(sc0.zgrd, sc0.ext1d, sc0.omg1d, sc0.apf1d, nx=16, ny=12, dx=100.0, dy=100.0, iz3l=sc0.struct.iz3l,
                                          ext3d=sc0.ext3d, omg3d=sc0.omg3d, apf3d=sc0.apf3d, abs3d=bad), opt)
