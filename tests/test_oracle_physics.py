"""
Pins for the CPU oracle (oracle/oracle_mc.cpp).  MCARaTS itself is not available offline and the reference ships no
golden vectors for transport, so the oracle is pinned against
  * a deterministic adding-doubling solver (oracle/adding_doubling.py): flux profile and nadir radiance,
  * analytic identities: direct beam, energy conservation, Lambertian surface without atmosphere, single scattering,
    IPA == 1-D on a horizontally uniform field, BRDF quadrature albedo.
Tolerances are a few Monte Carlo standard errors at the photon counts used (written next to each assertion).
"""

import ctypes as C

import numpy as np
import pytest

import oracle
import scenes
from oracle import adding_doubling as ad
from er3t_b200 import abi


def energy_balance(st):
    return (st['w_toa_up'] + st['w_sfc_abs'] + st['w_atm_abs'] - st['w_roulette']) / st['photons'] - 1.0


def test_against_adding_doubling_flux_profile_and_nadir_radiance():
    ns = 48
    mu, w = ad.radau_nodes(ns)
    k0 = int(np.argmin(abs(mu - 0.85)))
    mu0 = mu[k0]
    sza = np.rad2deg(np.arccos(mu0))
    z = scenes.std_z()
    nz = z.size - 1
    ext = np.zeros((2, nz)); omg = np.ones((2, nz)); apf = np.zeros((2, nz))
    ext[0] = scenes.rayleigh_ext(z); apf[0] = -1.0
    ext[1, 1] = 8.0 / 1000.0; apf[1, 1] = 0.8; omg[1, 1] = 0.99
    absg = np.zeros(nz); absg[:5] = 2e-5
    pr, ph = ad.rayleigh(), ad.hg(0.8)
    layers = [dict(dz=1000.0, comps=[(ext[0, iz], 1.0, pr), (ext[1, iz], omg[1, iz], ph)], absorb=absg[iz]) for iz in range(nz - 1, -1, -1)]
    res = ad.solve(layers, 0.2, k0, nstream=ns)
    sc = abi.HostScene(z, ext, omg, apf, sfc_type=1, sfc_param=(0.2, 0, 0, 0, 0), src_the=180 - sza, src_phi=270.0, src_qmax=0.0,
                       sensors=[dict(the=180.0, phi=270.0, nxr=1, nyr=1)])
    opt = abi.make_options(target=abi.TARGET_FLUX | abi.TARGET_RADIANCE, nslab=1, wmin=0.0)
    jobs, keep = abi.make_jobs([400000], [7], [0], abs1d=[absg])
    r = oracle.run(sc, opt, jobs)
    f = r['flux'][0, :, :, 0, 0]
    # 4e5 photons: standard error of a flux ~ 8e-4 * mu0; adding-doubling discretisation ~ 2e-4
    assert np.max(np.abs(f[2] - res['f_up'])) / mu0 < 3.5e-3
    assert np.max(np.abs(f[1] - res['f_down'])) / mu0 < 3.5e-3
    assert np.max(np.abs(f[0] - res['f_down_direct'])) / mu0 < 2e-3
    assert abs(r['rad'][0] / res['rad_nadir_toa'] - 1.0) < 0.01
    assert abs(energy_balance(r['stats'])) < 1e-10


def test_direct_beam_and_energy_conservation_with_roulette():
    sc, absg = scenes.plane_parallel(cot=3.0, omega=0.9, albedo=0.3, absorb=True, qmax=0.0, with_sensor=False)
    opt = abi.make_options(target=abi.TARGET_FLUX, nslab=1, wmin=0.2)
    jobs, keep = abi.make_jobs([300000], [3], [0], abs1d=[absg])
    r = oracle.run(sc, opt, jobs)
    st = r['stats']
    assert abs(energy_balance(st)) < 1e-10                      # fp64 accumulation
    assert st['n_roulette_kill'] > 0
    assert abs(st['w_roulette'] / st['photons']) < 5e-3          # roulette is unbiased: created == destroyed weight on average
    mu0 = np.cos(np.deg2rad(30.0))
    dz = np.diff(sc.zgrd)
    tau = np.cumsum(((sc.ext1d.sum(axis=0) + absg) * dz)[::-1])[::-1]      # optical depth from TOA down to each level
    expect = mu0 * np.exp(-np.concatenate([tau, [0.0]]) / mu0)
    got = r['flux'][0, 0, :, 0, 0]
    assert np.max(np.abs(got - expect)) < 3e-3                   # binomial error of 3e5 photons


def test_lambertian_surface_without_atmosphere():
    z = np.array([0.0, 1000.0])
    sc = abi.HostScene(z, [[1e-12]], [[1.0]], [[-1.0]], sfc_type=1, sfc_param=(0.35, 0, 0, 0, 0), src_the=180.0 - 40.0, src_phi=10.0,
                       sensors=[dict(the=180.0, phi=270.0, nxr=1, nyr=1), dict(the=180.0 - 50.0, phi=123.0, nxr=1, nyr=1)])
    opt = abi.make_options(target=abi.TARGET_FLUX | abi.TARGET_RADIANCE, nslab=1, wmin=0.0)
    jobs, keep = abi.make_jobs([20000], [1], [0])
    r = oracle.run(sc, opt, jobs)
    mu0 = np.cos(np.deg2rad(40.0))
    assert np.allclose(r['rad'], mu0 * 0.35 / np.pi, rtol=1e-4)      # deterministic up to the solar cone
    assert np.allclose(r['flux'][0, 2, :, 0, 0], mu0 * 0.35, rtol=1e-4)


def test_single_scattering_radiance_analytic():
    # thin homogeneous HG layer over a black surface: I = mu0 * omega * P(Theta) / (4 pi) * (1 - exp(-tau (1/mu0 + 1/muv))) / (mu0 + muv) ... (mu0 muv form)
    tau, g, omega = 0.5, 0.6, 0.9
    z = np.array([0.0, 1000.0])
    sza, vza = 35.0, 20.0
    sc = abi.HostScene(z, [[tau / 1000.0]], [[omega]], [[g]], sfc_type=1, sfc_param=(0.0, 0, 0, 0, 0), src_the=180.0 - sza, src_phi=0.0, src_qmax=0.0,
                       sensors=[dict(the=180.0 - vza, phi=180.0, nxr=1, nyr=1)])
    opt = abi.make_options(target=abi.TARGET_RADIANCE, nslab=1, wmin=0.0, iso_max=1)
    jobs, keep = abi.make_jobs([400000], [11], [0])
    r = oracle.run(sc, opt, jobs)
    mu0, muv = np.cos(np.deg2rad(sza)), np.cos(np.deg2rad(vza))
    # photon direction (sin sza, 0, -cos sza); toward sensor: -(view) = (sin vza cos(180)*-1 ...) -> computed explicitly
    d = np.array([np.sin(np.deg2rad(180 - sza)) * 1.0, 0.0, np.cos(np.deg2rad(180 - sza))])
    v = np.array([np.sin(np.deg2rad(180 - vza)) * np.cos(np.pi), 0.0, np.cos(np.deg2rad(180 - vza))])
    cosang = float(np.dot(d, -v))
    P = (1 - g * g) / (1 + g * g - 2 * g * cosang) ** 1.5
    expect = mu0 * omega * P / (4 * np.pi) * (1 - np.exp(-tau * (1 / mu0 + 1 / muv))) / (muv * (1 / mu0 + 1 / muv)) / mu0
    assert abs(r['rad'][0] / expect - 1.0) < 0.01


def test_reciprocity_of_multiple_scattering_at_oblique_angles():
    """Helmholtz reciprocity of a plane-parallel atmosphere over a Lambertian surface: with the results normalised per
    unit flux NORMAL to the beam, I(sun a -> view b) / mu_a = I(sun b -> view a) / mu_b at the same relative azimuth.
    Pins the oblique local estimate under multiple scattering (Rayleigh + HG cloud of optical depth 2 + surface)."""
    za, zb, dphi = 30.0, 55.0, 40.0
    out = []
    for sza, vza in ((za, zb), (zb, za)):
        sc, _ = scenes.plane_parallel(sza=sza, cot=2.0, g=0.7, omega=0.98, albedo=0.2, with_sensor=False, qmax=0.0)
        # sun travels toward azimuth 270 (src_phi); the sensor direction of travel (toward the sensor) is rotated by dphi
        sc2 = abi.HostScene(sc.zgrd, sc.ext1d, sc.omg1d, sc.apf1d, sfc_type=1, sfc_param=(0.2, 0, 0, 0, 0), src_the=180.0 - sza, src_phi=270.0,
                            src_qmax=0.0, sensors=[dict(the=180.0 - vza, phi=270.0 + dphi, nxr=1, nyr=1)])
        opt = abi.make_options(target=abi.TARGET_RADIANCE, nslab=4, wmin=0.0)
        jobs, keep = scenes.multi_seed_jobs(100000, 4, seed0=77)
        r = oracle.run(sc2, opt, jobs)
        m, se = scenes.mean_sem(r['rad'].reshape(4))
        out.append((m / np.cos(np.deg2rad(sza)), se / np.cos(np.deg2rad(sza))))
    (a, sa), (b, sb) = out
    assert abs(a - b) < 4.0 * np.hypot(sa, sb) + 0.005 * a, (a, b, sa, sb)


def test_ipa_equals_1d_on_uniform_field_and_3d_agrees():
    z = scenes.std_z(10, 10000.0)
    nz = 10
    ext1 = scenes.rayleigh_ext(z)[None, :]
    e3 = np.full((4, 3, 2), 0.004, dtype=np.float32)
    kw = dict(nx=4, ny=3, dx=200.0, dy=200.0, iz3l=2, ext3d=e3, omg3d=np.ones_like(e3), apf3d=np.full_like(e3, 0.7),
              sfc_type=1, sfc_param=(0.1, 0, 0, 0, 0), src_the=140.0, src_phi=30.0, sensors=[dict(the=180.0, phi=270.0, nxr=4, nyr=3)])
    sc3 = abi.HostScene(z, ext1, np.ones((1, nz)), -np.ones((1, nz)), **kw)
    # the same medium as two 1-D components
    ext2 = np.zeros((2, nz)); ext2[0] = ext1[0]; ext2[1, 1:3] = 0.004
    omg2 = np.ones((2, nz)); apf2 = np.zeros((2, nz)); apf2[0] = -1; apf2[1] = 0.7
    sc1 = abi.HostScene(z, ext2, omg2, apf2, sfc_type=1, sfc_param=(0.1, 0, 0, 0, 0), src_the=140.0, src_phi=30.0, sensors=[dict(the=180.0, phi=270.0, nxr=1, nyr=1)])
    res = {}
    for name, sc, solver in (('3d', sc3, abi.SOLVER_3D), ('ipa', sc3, abi.SOLVER_IPA), ('p3d', sc3, abi.SOLVER_PARTIAL_3D), ('1d', sc1, abi.SOLVER_3D)):
        opt = abi.make_options(solver=solver, target=abi.TARGET_FLUX | abi.TARGET_RADIANCE, nslab=1, wmin=0.0)
        jobs, keep = abi.make_jobs([150000], [5], [0])
        r = oracle.run(sc, opt, jobs)
        res[name] = (r['rad'].mean(), r['flux'][0, 2, -1].mean(), r['flux'][0, 1, 0].mean())
    for name in ('3d', 'ipa', 'p3d'):
        for a, b in zip(res[name], res['1d']):
            assert abs(a / b - 1.0) < 0.012, (name, a, b)        # ~3 sigma at 1.5e5 photons


@pytest.mark.parametrize('kind', ['lsrt', 'dsm'])
def test_brdf_sampling_matches_quadrature_albedo(kind):
    """surface_sample weights (oracle) integrate to the albedo obtained by quadrature of the BRDF the local estimate uses."""
    lib = oracle.load()
    if kind == 'lsrt':
        t, prm = abi.SFC_LSRT, np.array([0.2, 0.03, 0.1, 0, 0], dtype=np.float32)
    else:
        t, prm = abi.SFC_DSM, np.array([0.22, 0.05, 1.34, 1e-7, 0.04], dtype=np.float32)
    sza = 50.0
    mu0 = np.cos(np.deg2rad(sza))
    # quadrature over the upper hemisphere
    nmu, nphi = 400, 720
    mu = (np.arange(nmu) + 0.5) / nmu
    phi = (np.arange(nphi) + 0.5) * 2 * np.pi / nphi
    MU, PH = np.meshgrid(mu, phi, indexing='ij')
    st = np.sqrt(1 - MU ** 2)
    dout = np.stack([st * np.cos(PH), st * np.sin(PH), MU], axis=-1).reshape(-1, 3)
    din = np.tile(np.array([np.sin(np.deg2rad(sza)), 0.0, -mu0]), (dout.shape[0], 1))
    f = np.zeros(dout.shape[0])
    lib.oracle_brdf_eval(t, prm.ctypes.data, np.ascontiguousarray(din).ctypes.data, np.ascontiguousarray(dout).ctypes.data, f.ctypes.data, f.size)
    albedo = np.sum(f.reshape(nmu, nphi) * MU) * (1.0 / nmu) * (2 * np.pi / nphi)
    z = np.array([0.0, 100.0])
    sc = abi.HostScene(z, [[1e-12]], [[1.0]], [[-1.0]], sfc_type=np.full((1, 1), t, dtype=np.int32), sfc_param=prm.reshape(1, 1, 5),
                       src_the=180.0 - sza, src_phi=0.0, src_qmax=0.0)
    opt = abi.make_options(target=abi.TARGET_FLUX, nslab=1, wmin=0.0)
    jobs, keep = abi.make_jobs([400000], [2], [0])
    r = oracle.run(sc, opt, jobs)
    got = r['flux'][0, 2, -1, 0, 0] / mu0
    assert abs(got / albedo - 1.0) < 0.01, (got, albedo)


def test_tabulated_phase_function_sampling_is_consistent_with_evaluation():
    lib = oracle.load()
    ang, pha = scenes.synthetic_mie_table()
    sc = abi.HostScene(np.array([0.0, 1.0]), [[0.0]], [[1.0]], [[0.0]], ang=ang, pha=pha)
    xi = (np.arange(200000) + 0.5) / 200000
    mu_s = np.zeros_like(xi)
    lib.oracle_phase_sample(C.addressof(sc.struct), 2.0, xi.ctypes.data, mu_s.ctypes.data, xi.size)
    edges = np.linspace(-1, 1, 41)
    hist, _ = np.histogram(mu_s, bins=edges)
    mid = np.linspace(-1, 1, 4001)
    p = np.zeros_like(mid)
    lib.oracle_phase_eval(C.addressof(sc.struct), 2.0, mid.ctypes.data, p.ctypes.data, mid.size)
    assert abs(np.trapezoid(p, mid) - 2.0) < 2e-3                     # normalisation (1/2) int P dmu = 1
    cdf = np.concatenate([[0], np.cumsum(0.5 * (p[1:] + p[:-1]) * np.diff(mid))]) / 2.0
    expect = np.diff(np.interp(edges, mid, cdf)) * xi.size
    big = expect > 200
    assert np.max(np.abs(hist[big] / expect[big] - 1.0)) < 0.02
    g_sample = mu_s.mean()
    g_eval = 0.5 * np.trapezoid(p * mid, mid)
    assert abs(g_sample - g_eval) < 2e-3


def test_all_sky_camera_matches_parallel_projection_sensors():
    """Rad_mrkind = 1: in a plane-parallel atmosphere a ground camera looking up must record the same downward radiance
    field that 2nd-kind sensors (parallel projection, domain average) placed at the ground report for the same viewing
    vectors -- two independent estimators (1/R^2 point detector vs area average) in the SAME run.  Compared as
    azimuthal means over two rings of zenith angle (the point detector is noisy pixel by pixel).  The layer next to
    the camera is empty so that its estimator has bounded variance."""
    z = np.arange(0.0, 12001.0, 1000.0)
    nz = z.size - 1
    ext = np.zeros((2, nz)); omg = np.ones((2, nz)); apf = np.zeros((2, nz))
    ext[0] = scenes.rayleigh_ext(z, 0.2); apf[0] = -1.0
    ext[0, 0] = 0.0
    ext[1, 2] = 4.0 / 1000.0; apf[1, 2] = 0.8
    npx = 36
    rings = [(15.0, 10.0, 20.0, 6), (60.0, 50.0, 70.0, 12)]       # centre, inner, outer zenith angle (deg), azimuths
    dirs = [(t, 360.0 * k / n + 7.0) for t, lo, hi, n in rings for k in range(n)]
    sens = [dict(kind=1, the=0.0, phi=0.0, psi=0.0, zloc=0.0, nxr=npx, nyr=npx, xpos=0.5, ypos=0.5, qmax=178.0, umax=180.0, vmax=180.0, apsize=0.05)]
    sens += [dict(kind=2, the=t, phi=f, zloc=0.0, nxr=1, nyr=1) for t, f in dirs]
    sc = abi.HostScene(z, ext, omg, apf, dx=6.0e4, dy=6.0e4, sfc_type=1, sfc_param=(0.0, 0, 0, 0, 0), src_the=180.0 - 35.0, src_phi=270.0,
                       sensors=sens)
    nslab = 8
    opt = abi.make_options(target=abi.TARGET_RADIANCE, nslab=nslab, wmin=0.2)
    jobs, keep = scenes.multi_seed_jobs(500000, nslab)
    r = oracle.run(sc, opt, jobs)
    rad = r['rad'].reshape(nslab, -1)
    cam = rad[:, :npx * npx].reshape(nslab, npx, npx)            # [iy (V), ix (U)]
    du = np.pi / npx
    u = (np.arange(npx) + 0.5) * du - np.pi / 2
    U, V = np.meshgrid(u, u, indexing='xy')
    th = np.rad2deg(np.hypot(U, V))
    k0 = 0
    for t, lo, hi, n in rings:
        sel = (th > lo) & (th < hi)
        m_c, s_c = scenes.mean_sem(cam[:, sel].mean(axis=1))
        m_p, s_p = scenes.mean_sem(rad[:, npx * npx + k0:npx * npx + k0 + n].mean(axis=1))
        k0 += n
        assert m_p > 0.05
        assert abs(m_c / m_p - 1.0) < 0.08 or abs(m_c - m_p) < 3.5 * np.hypot(s_c, s_p), (t, m_c, m_p, s_c, s_p)
    # the sun is seen at zenith angle 35 deg, azimuth 90 deg: that half of the sky is the brighter one
    assert cam.mean(axis=0)[V > 0].mean() > 1.3 * cam.mean(axis=0)[V < 0].mean()


# ---------------------------------------------------------------------------------------------------------------------
# tabulated (Mie) phase functions and oblique views against the deterministic solver
# ---------------------------------------------------------------------------------------------------------------------
def test_beam_source_solver_matches_node_beam_solver():
    """Two independent formulations of the deterministic solver (solar beam as a quadrature node vs as a source term,
    Radau vs Gauss grid) agree on fluxes and oblique radiances."""
    ns = 40
    mu, w = ad.radau_nodes(ns)
    k0 = int(np.argmin(abs(mu - 0.85)))
    z = scenes.std_z()
    nz = z.size - 1
    e0 = scenes.rayleigh_ext(z)
    pr, ph = ad.rayleigh(), ad.hg(0.8)
    layers = [dict(dz=1000.0, comps=[(e0[iz], 1.0, pr), (8e-3 if iz == 1 else 0.0, 0.99, ph)], absorb=2e-5 if iz < 5 else 0.0) for iz in range(nz - 1, -1, -1)]
    views = [(0.0, 0.0), (30.0, 40.0), (60.0, 150.0)]
    a = ad.solve_views(layers, 0.2, k0, nstream=ns, views=views, nmode=12)
    b = ad.solve_beam(layers, 0.2, np.rad2deg(np.arccos(mu[k0])), nstream=ns, views=views, nmode=12)
    assert np.max(np.abs(a['f_up'] - b['f_up'])) < 2e-6
    assert np.max(np.abs(a['f_down'] - b['f_down'])) < 2e-6
    assert np.allclose(a['rad_views'], b['rad_views'], rtol=3e-6)
    assert abs(a['rad_views'][0] / ad.solve(layers, 0.2, k0, nstream=ns)['rad_nadir_toa'] - 1.0) < 1e-6   # (different azimuth quadratures)


def test_fixture_hg_reproduced_at_low_resolution():
    """The committed fixture is what oracle/adding_doubling.py produces (HG converges with few streams)."""
    fx = scenes.ad_fixture('hg')
    z = fx['z']; nz = z.size - 1
    pr, ph = ad.rayleigh(), ad.hg(0.85)
    layers = [dict(dz=float(z[iz + 1] - z[iz]), comps=[(fx['ext'][0, iz], 1.0, pr), (fx['ext'][1, iz], fx['omg'][1, iz], ph)], absorb=fx['absg'][iz])
              for iz in range(nz - 1, -1, -1)]
    r = ad.solve_beam(layers, float(fx['albedo']), float(fx['sza']), nstream=48, views=[tuple(v) for v in fx['views']], nmode=16)
    assert np.max(np.abs(r['f_up'] - fx['f_up'])) < 2e-5
    assert np.max(np.abs(r['f_down'] - fx['f_down'])) < 2e-5
    assert np.allclose(r['rad_views'], fx['rad_views'], rtol=2e-4)


@pytest.mark.parametrize('name', ['mie', 'hg'])
def test_oracle_against_deterministic_oblique_views_and_tables(name):
    """oracle_mc.cpp against the converged adding-doubling result: flux at all 21 levels, TOA radiance at nadir and four
    oblique directions, for the 498-angle Mie table (piecewise linear in mu) and for Henyey-Greenstein."""
    fx = scenes.ad_fixture(name)
    sc = scenes.ad_scene(fx)
    nslab = 8
    opt = abi.make_options(target=abi.TARGET_FLUX | abi.TARGET_RADIANCE, nslab=nslab, wmin=0.2)
    jobs, keep = scenes.multi_seed_jobs(250000, nslab, abs1d=fx['absg'])
    r = oracle.run(sc, opt, jobs)
    mu0 = float(fx['mu0'])
    f = r['flux'][:, :, :, 0, 0]                                   # (nslab, 3, nlev)
    for var, key in ((2, 'f_up'), (1, 'f_down'), (0, 'f_down_direct')):
        m, s = scenes.mean_sem(f[:, var])
        # 2e6 photons: standard error of a flux ~ 4e-4 * mu0
        assert np.max(np.abs(m - fx[key]) / (4.0 * s + 2e-4 * mu0)) < 1.0, (key, np.max(np.abs(m - fx[key])) / mu0)
    rad = r['rad'].reshape(nslab, -1)
    m, s = scenes.mean_sem(rad)
    z = (m - fx['rad_views']) / np.sqrt(s ** 2 + (3e-4 * fx['rad_views']) ** 2)
    assert np.max(np.abs(z)) < 4.0, (z, m / fx['rad_views'])
    assert np.max(np.abs(m / fx['rad_views'] - 1.0)) < 0.03        # and in absolute terms: a few per cent at 2e6 photons
    assert abs(energy_balance(r['stats'])) < 1e-10
