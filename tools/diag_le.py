"""High-statistics GPU-vs-oracle comparison of the local-estimate radiance (diagnostic; run on a GPU box: python tools/diag_le.py)."""
import sys, os, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import scenes, oracle
from er3t_b200 import abi
from er3t_b200.solver import Solver
s = Solver(0)
def both(sc, opt, ng=20000000, nc=2000000, absg=None):
    nslab = opt.nslab
    jg, k1 = scenes.multi_seed_jobs(ng // nslab, nslab, abs1d=absg)
    jc, k2 = scenes.multi_seed_jobs(nc // nslab, nslab, seed0=77, abs1d=absg)
    s.upload_scene(sc, opt); s.run(jg); g = s.results()
    c = oracle.run(sc, opt, jc)
    gm, gs = scenes.mean_sem(g['rad'].reshape(nslab, -1).mean(axis=1))
    cm, cs = scenes.mean_sem(c['rad'].reshape(nslab, -1).mean(axis=1))
    sg, so = g['stats'], c['stats']
    print('  rad gpu %.6f +- %.6f  cpu %.6f +- %.6f  ratio %.5f  z %.2f' % (gm, gs, cm, cs, gm / cm, (gm - cm) / np.hypot(gs, cs)))
    print('  coll/phot gpu %.4f cpu %.4f ; le/phot gpu %.4f cpu %.4f ; sfc/phot gpu %.4f cpu %.4f ; R gpu %.5f cpu %.5f ; Mph/s %.1f' % (
        sg['n_coll'] / sg['photons'], so['n_coll'] / so['photons'], sg['n_le'] / sg['photons'], so['n_le'] / so['photons'],
        sg['n_sfc'] / sg['photons'], so['n_sfc'] / so['photons'], sg['w_toa_up'] / sg['photons'], so['w_toa_up'] / so['photons'],
        sg['photons'] / sg['elapsed_ms'] / 1e3))
for name, kw, okw in [
    ('base g=0.85', dict(), dict()),
    ('isotropic g=0', dict(g=0.0), dict()),
    ('g=0.85 single scatter', dict(), dict(iso_max=1)),
    ('g=0.85 two orders', dict(), dict(iso_max=2)),
    ('g=0.85 black surface', dict(albedo=0.0), dict()),
    ('g=0.85 sza=0 qmax=0', dict(sza=0.001, qmax=0.0), dict()),
    ('thin cot=0.5', dict(cot=0.5), dict()),
]:
    sc, absg = scenes.plane_parallel(**kw)
    opt = abi.make_options(target=abi.TARGET_RADIANCE, nslab=10, wmin=0.0, **okw)
    print(name)
    both(sc, opt)
