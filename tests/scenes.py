"""Seeded synthetic scenes shared by the oracle tests (CPU) and the parity tests (GPU)."""

import numpy as np

from er3t_b200 import abi


def std_z(nz=20, top=20000.0):
    return np.linspace(0.0, top, nz + 1)


def rayleigh_ext(z, tau_total=0.045):
    """Exponential (8 km scale height) Rayleigh extinction profile with the given total optical depth."""
    zc = 0.5 * (z[1:] + z[:-1])
    dz = z[1:] - z[:-1]
    e = np.exp(-zc / 8000.0)
    return e * tau_total / np.sum(e * dz)


def synthetic_mie_table(nang=361, g1=0.88, g2=-0.5, f=0.97):
    """Forward-peaked double Henyey-Greenstein tabulated on a regular angle grid (a stand-in for a Mie table)."""
    ang = np.linspace(0.0, 180.0, nang)
    mu = np.cos(np.deg2rad(ang))

    def hg(g):
        return (1 - g * g) / (1 + g * g - 2 * g * mu) ** 1.5
    tabs = []
    for gg in (g1, g1 - 0.03, g1 - 0.06):
        tabs.append(f * hg(gg) + (1 - f) * hg(g2))
    return ang, np.stack(tabs, axis=1)        # (nang, npf)


def plane_parallel(sza=30.0, cot=10.0, g=0.85, omega=1.0, albedo=0.03, absorb=False, with_sensor=True, qmax=0.533133,
                   apf_cloud=None, table=False):
    """Config-1 style scene: 20 x 1 km layers, Rayleigh background, cloud as 2nd 1-D component in 1-2 km
    (er3t/rtm/mca/util.py:150-159)."""
    z = std_z()
    nz = z.size - 1
    ext = np.zeros((2, nz)); omg = np.ones((2, nz)); apf = np.zeros((2, nz))
    ext[0] = rayleigh_ext(z); apf[0] = -1.0
    ext[1, 1] = cot / 1000.0; omg[1, 1] = omega
    apf[1, 1] = g if apf_cloud is None else apf_cloud
    kw = {}
    if table:
        ang, pha = synthetic_mie_table()
        kw.update(ang=ang, pha=pha)
    sensors = [dict(the=180.0, phi=270.0, nxr=1, nyr=1)] if with_sensor else []
    sc = abi.HostScene(z, ext, omg, apf, sfc_type=1, sfc_param=(albedo, 0, 0, 0, 0), src_the=180.0 - sza, src_phi=270.0,
                       src_qmax=qmax, sensors=sensors, **kw)
    absg = None
    if absorb:
        absg = np.zeros(nz)
        absg[:6] = 3.0e-5 * np.exp(-np.arange(6) / 2.0)
    return sc, absg


def cloud_field(nx=16, ny=12, nz3=4, seed=2, dx=100.0, dz=200.0, zbase=600.0, cf=0.35, ext0=0.03, two_comp=False,
                apf_mode='hg'):
    """Blocky random cumulus-like field: (nx, ny, nz3) extinction in 1/m, g (or table index) and omega per voxel."""
    rng = np.random.default_rng(seed)
    base = rng.random((nx, ny)) < cf
    # smooth a little so that clouds span a few columns
    base = base | np.roll(base, 1, axis=0) & (rng.random((nx, ny)) < 0.7)
    top = rng.integers(1, nz3 + 1, size=(nx, ny))
    ext = np.zeros((nx, ny, nz3), dtype=np.float32)
    for k in range(nz3):
        ext[:, :, k] = np.where(base & (k < top), ext0 * (0.5 + rng.random((nx, ny))), 0.0)
    omg = np.ones_like(ext)
    apf = np.full_like(ext, -1.0)
    cld = ext > 0
    if apf_mode == 'hg':
        apf[cld] = (0.80 + 0.08 * rng.random(ext.shape))[cld]
    elif apf_mode == 'table':
        apf[cld] = (1.0 + 2.0 * rng.random(ext.shape))[cld]
    omg[cld] = (0.98 + 0.02 * rng.random(ext.shape))[cld]
    return ext, omg.astype(np.float32), apf.astype(np.float32)


def scene_3d(nx=16, ny=12, nz3=4, sza=40.0, saa_phi=200.0, sensors=None, sfc='lambert', seed=2, two_comp=False,
             apf_mode='hg', dz3=200.0, zbase_layer=3, nlay=12, table=False, qmax=0.533133, clear_below=False):
    """
    Small 3-D cloud scene.  Atmosphere: `nlay` layers, non-uniform: fine (dz3) layers around the 3-D block.
    The 3-D block starts at 1-based layer `zbase_layer`.
    """
    # level grid: two coarse layers, nz3 fine, rest coarse
    z = [0.0]
    for i in range(zbase_layer - 1):
        z.append(z[-1] + 300.0)
    dz3s = np.broadcast_to(np.asarray(dz3, dtype=np.float64), (nz3,))      # a scalar, or one thickness per 3-D layer
    for i in range(nz3):
        z.append(z[-1] + float(dz3s[i]))
    while len(z) < nlay + 1:
        z.append(z[-1] + 1500.0)
    z = np.array(z)
    nz = z.size - 1
    ext1 = np.zeros((1, nz)); omg1 = np.ones((1, nz)); apf1 = -np.ones((1, nz))
    ext1[0] = rayleigh_ext(z, 0.05)
    if clear_below:
        ext1[0, :zbase_layer - 1] = 0.0      # nothing scatters next to a ground-based camera (bounded 1/R^2 variance)
    ext, omg, apf = cloud_field(nx, ny, nz3, seed=seed, apf_mode=apf_mode, dz=float(np.mean(dz3)))
    kw = {}
    if two_comp:
        rng = np.random.default_rng(seed + 100)
        ext2 = (2.0e-4 * rng.random(ext.shape)).astype(np.float32)        # thin absorbing aerosol everywhere
        omg2 = np.full_like(ext, 0.9)
        apf2 = np.full_like(ext, 0.6)
        ext = np.stack([ext, ext2], axis=-1); omg = np.stack([omg, omg2], axis=-1); apf = np.stack([apf, apf2], axis=-1)
    if table or apf_mode == 'table':
        ang, pha = synthetic_mie_table()
        kw.update(ang=ang, pha=pha)
    if sfc == 'lambert':
        st, sp = 1, (0.1, 0, 0, 0, 0)
    elif sfc == 'lambert2d':
        rng = np.random.default_rng(seed + 7)
        st = np.ones((nx, ny), dtype=np.int32)
        sp = np.zeros((nx, ny, 5), dtype=np.float32)
        sp[..., 0] = 0.05 + 0.3 * rng.random((nx, ny))
    elif sfc == 'lsrt':
        rng = np.random.default_rng(seed + 8)
        st = np.full((nx, ny), 4, dtype=np.int32)
        sp = np.zeros((nx, ny, 5), dtype=np.float32)
        sp[..., 0] = 0.05 + 0.25 * rng.random((nx, ny))
        sp[..., 1] = 0.05 * rng.random((nx, ny))
        sp[..., 2] = 0.15 * rng.random((nx, ny))
    elif sfc == 'dsm':
        st = np.full((nx, ny), 2, dtype=np.int32)
        sp = np.zeros((nx, ny, 5), dtype=np.float32)
        sp[..., 0] = 0.22; sp[..., 1] = 0.01; sp[..., 2] = 1.34; sp[..., 3] = 1.0e-7; sp[..., 4] = 0.0286
    else:
        raise ValueError(sfc)
    if sensors is None:
        sensors = [dict(the=180.0, phi=270.0, nxr=nx, nyr=ny)]
    sc = abi.HostScene(z, ext1, omg1, apf1, nx=nx, ny=ny, dx=100.0, dy=100.0, iz3l=zbase_layer, ext3d=ext, omg3d=omg,
                       apf3d=apf, sfc_type=st, sfc_param=sp, src_the=180.0 - sza, src_phi=saa_phi, src_qmax=qmax,
                       sensors=sensors, **kw)
    return sc


def multi_seed_jobs(nphot, nslab, seed0=1000, abs1d=None):
    """`nslab` independent repetitions (one slab each) -> mean and standard error per pixel."""
    return abi.make_jobs([nphot] * nslab, [seed0 + 17 * i for i in range(nslab)], list(range(nslab)),
                         abs1d=None if abs1d is None else [abs1d] * nslab)


def mean_sem(a):
    """mean and standard error over axis 0 (independent repetitions)."""
    a = np.asarray(a)
    return a.mean(axis=0), a.std(axis=0, ddof=1) / np.sqrt(a.shape[0])


# ---------------------------------------------------------------------------------------------------------------------
# deterministic plane-parallel benchmark cases (tests/golden/ad_fixtures.npz, made by tests/golden/make_ad_fixtures.py)
# ---------------------------------------------------------------------------------------------------------------------
def ad_fixture(name):
    """dict of the arrays of one adding-doubling benchmark case ('mie' or 'hg')."""
    import os
    f = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ad_fixtures.npz'))
    pre = name + '_'
    return {k[len(pre):]: f[k] for k in f.files if k.startswith(pre)}


def ad_scene(fx, hom3d=False, src_phi=200.0):
    """The benchmark case as a transport scene.  View k of the fixture is (vza, dphi) with dphi = azimuth of the photon's
    direction of travel toward the sensor minus the azimuth of the solar direction of travel; the sensor's VIEWING vector
    (include/b200rt.h: `the`, `phi`) points the other way, hence phi_view = src_phi + dphi - 180.
    hom3d: the cloud layer as a horizontally uniform 2 x 2-column 3-D block instead of a 1-D component (the
    cld_gen_hom variant of config 1, er3t/rtm/mca/util.py:340-364)."""
    z = fx['z']
    sensors = [dict(the=180.0 - float(v), phi=(src_phi + float(d) - 180.0) % 360.0, nxr=1, nyr=1) for v, d in fx['views']]
    kw = {}
    if 'ang' in fx:
        kw.update(ang=fx['ang'], pha=fx['pha'])
    ext, omg, apf = fx['ext'].copy(), fx['omg'].copy(), fx['apf'].copy()
    if hom3d:
        icl = int(np.argmax(ext[1]))
        e3 = np.full((2, 2, 1), ext[1, icl], dtype=np.float32)
        o3 = np.full((2, 2, 1), omg[1, icl], dtype=np.float32)
        a3 = np.full((2, 2, 1), apf[1, icl], dtype=np.float32)
        for s in sensors:
            s.update(nxr=2, nyr=2)
        kw.update(nx=2, ny=2, dx=100.0, dy=100.0, iz3l=icl + 1, ext3d=e3, omg3d=o3, apf3d=a3)
        ext, omg, apf = ext[:1], omg[:1], apf[:1]
    return abi.HostScene(z, ext, omg, apf, sfc_type=1, sfc_param=(float(fx['albedo']), 0, 0, 0, 0), src_the=180.0 - float(fx['sza']),
                         src_phi=src_phi, src_qmax=0.0, sensors=sensors, **kw)
