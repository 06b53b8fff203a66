"""Per-COT comparison of func_ref_vs_cot on the GPU with the adding-doubling curves (tests/golden/ad_cot_sweep.npz)."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests', 'golden'))
import numpy as np
import make_cot_sweep as mk
import er3t_b200.rtm.mca as bmca
from er3t_b200.solver import Solver
fx = np.load(os.path.join(ROOT, 'tests', 'golden', 'ad_cot_sweep.npz'))
atm0, abs0, pha0 = mk.inputs()
sol = Solver(0)
nrun = 8
nph = float(os.environ.get('NPH', '1e7'))
f = bmca.func_ref_vs_cot(fx['cot'], cer0=10.0, fdir=None, date=mk.DATE, wavelength=650.0, surface_albedo=0.03, solar_zenith_angle=float(fx['sza']),
                         solar_azimuth_angle=238.9053, sensor_zenith_angle=0.0, sensor_azimuth_angle=261.9049, Nphoton=nph, seed=7, solver_obj=sol,
                         atm0=atm0, pha0=pha0, abs0=abs0, Nrun=nrun)
sem = f.ref_std / np.sqrt(nrun)
print('photons/COT %.0e x %d runs; kernel %.1f ms, %.1f M photons/s' % (nph, nrun, f.mca.stats['elapsed_ms'], f.mca.stats['photons'] / f.mca.stats['elapsed_ms'] / 1e3))
print('   COT      gpu        sem     ad_240    ad_160   (gpu-ad240)/ad240   z')
for i, c in enumerate(fx['cot']):
    print('%6.1f  %.6f  %.1e  %.6f  %.6f  %+.2e  %+.1f' % (c, f.ref[i], sem[i], fx['ref_n240'][i], fx['ref_n160'][i], f.ref[i] / fx['ref_n240'][i] - 1.0,
                                                            (f.ref[i] - fx['ref_n240'][i]) / max(sem[i], 1e-12)))
