#!/usr/bin/env python
"""
Generate tests/golden/ad_fixtures.npz: plane-parallel benchmark cases solved by the DETERMINISTIC adding-doubling
code (oracle/adding_doubling.py::solve_beam) at a resolution where it has converged to ~1e-4 or better, so that
the CUDA path can be compared with something that is neither a Monte Carlo code nor written around the oracle:
flux profile at all 21 levels and TOA radiance at nadir and oblique views, for the config-1 atmosphere
(BASELINE.json configs[0]: Rayleigh + gas absorption + a tau = 10 water cloud in 1-2 km over a Lambertian
surface, SZA 30 deg) with (a) the 498-angle Mie table of r_eff = 10 um at 650 nm and (b) Henyey-Greenstein g = 0.85.

The role the reference gives to such a check: examples/00_er3t_bmk.py:470-579 (MCARaTS against libRadtran/DISORT).

    python tests/golden/make_ad_fixtures.py [--nstream 320] [--nmode 96]      # ~20 min on 8 cores

Each case is solved twice (nstream and 0.75 * nstream; nmode and 0.75 * nmode) and the change is stored as the
convergence estimate `*_conv` next to every result.
"""
import argparse
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, '..', '..'))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from oracle import adding_doubling as ad      # noqa: E402
import scenes                                  # noqa: E402
from er3t_b200.pre import pha_mie_wc           # noqa: E402

SZA = 30.0
# (view zenith angle, azimuth of the photon's direction of travel toward the sensor minus that of the solar beam)
VIEWS = [(0.0, 0.0), (30.0, 40.0), (60.0, 100.0), (45.0, 0.0), (60.0, 150.0)]


def build_case(kind, albedo):
    z = scenes.std_z()
    nz = z.size - 1
    ext = np.zeros((2, nz)); omg = np.ones((2, nz)); apf = np.zeros((2, nz))
    ext[0] = scenes.rayleigh_ext(z); apf[0] = -1.0
    ext[1, 1] = 10.0 / 1000.0
    absg = np.zeros(nz)
    absg[:8] = 1.0e-5 * np.exp(-np.arange(8) / 3.0)
    case = dict(z=z, ext=ext, omg=omg, apf=apf, absg=absg, albedo=albedo, sza=SZA, views=np.array(VIEWS))
    if kind == 'mie':
        pha0 = pha_mie_wc(wavelength=650.0, reff=[10.0], nr=96)
        case['ang'] = pha0.data['ang']['data']
        case['pha'] = pha0.data['pha']['data']                       # (nang, 1)
        omg[1, 1] = float(pha0.data['ssa']['data'][0])
        apf[1, 1] = 1.0                                              # table index 1
        pcloud = ad.table(case['ang'], case['pha'][:, 0])
    else:
        apf[1, 1] = 0.85
        pcloud = ad.hg(0.85)
    pr = ad.rayleigh()
    layers = [dict(dz=float(z[iz + 1] - z[iz]), comps=[(ext[0, iz], 1.0, pr), (ext[1, iz], omg[1, iz], pcloud)], absorb=absg[iz])
              for iz in range(nz - 1, -1, -1)]
    return case, layers


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--nstream', type=int, default=320)
    ap.add_argument('--nmode', type=int, default=96)
    ap.add_argument('--out', default=os.path.join(HERE, 'ad_fixtures.npz'))
    a = ap.parse_args()
    out = {'nstream': a.nstream, 'nmode': a.nmode}
    for name, kind, albedo in (('mie', 'mie', 0.03), ('hg', 'hg', 0.3)):
        case, layers = build_case(kind, albedo)
        t0 = time.time()
        hi = ad.solve_beam(layers, albedo, SZA, nstream=a.nstream, views=VIEWS, nmode=a.nmode)
        lo = ad.solve_beam(layers, albedo, SZA, nstream=int(0.75 * a.nstream), views=VIEWS, nmode=int(0.75 * a.nmode))
        print('%s: %.0f s' % (name, time.time() - t0), flush=True)
        for k, v in case.items():
            out['%s_%s' % (name, k)] = np.asarray(v)
        for k in ('f_up', 'f_down', 'f_down_direct', 'rad_views', 'ss_tail'):
            out['%s_%s' % (name, k)] = hi[k]
            out['%s_%s_conv' % (name, k)] = np.abs(hi[k] - lo[k])
        out['%s_mu0' % name] = hi['mu0']
        print(name, 'f_up(TOA)/mu0 %.6f  f_down(sfc)/mu0 %.6f' % (hi['f_up'][-1] / hi['mu0'], hi['f_down'][0] / hi['mu0']))
        print(name, 'rad', hi['rad_views'], 'rel. change vs 0.75 resolution', np.abs(hi['rad_views'] - lo['rad_views']) / hi['rad_views'])
        print(name, 'max rel. flux change', np.max(np.abs(hi['f_up'] - lo['f_up']) / hi['mu0']), np.max(np.abs(hi['f_down'] - lo['f_down']) / hi['mu0']), flush=True)
    np.savez_compressed(a.out, **out)
    print('wrote', a.out)


if __name__ == '__main__':
    main()
