"""Where does the end-to-end step of bench.py go?  mca_atm_3d(device_props=True) / mcarats_ng scene build / upload (H2D +
packing kernels) / transport / D2H / mca_out_ng weighting, on the config-2 workload with the raw cloud fields in
page-locked host memory (what bench.py's e2e leg times as one number)."""
import sys, os, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from er3t_b200 import abi
from er3t_b200.util import pin_array
from er3t_b200.solver import Solver
from er3t_b200.rtm.mca import mcarats_ng, mca_out_ng, mca_atm_3d
kw, abs0, parts = bench.build_workload(480, 480, 100, float(os.environ.get('PHOT', '1e8')), parts=True)
pinned = os.environ.get('PIN', '1') == '1'
if pinned:
    for key in ('extinction', 'cer'):
        parts['cld'].lay[key]['data'] = pin_array(np.asarray(parts['cld'].lay[key]['data'], dtype=np.float32))
sol = Solver(0)
for it in range(4):
    t0 = time.time()
    a3 = mca_atm_3d(cld_obj=parts['cld'], atm_obj=parts['atm'], pha_obj=parts['pha'], quiet=True, device_props=True)
    t1 = time.time()
    m = mcarats_ng(**dict(kw, atm_3ds=[a3], dry_run=True))
    jobs, keep = abi.make_jobs(**m.jobs_args)
    t2 = time.time()
    sol.upload_scene(m.scene, m.options)
    t3 = time.time()
    sol.run(jobs)
    t4 = time.time()
    res = sol.results()
    t5 = time.time()
    m2 = mcarats_ng(**dict(kw, atm_3ds=[a3], solver_obj=sol))
    t6 = time.time()
    out = mca_out_ng(mca_obj=m2, abs_obj=abs0, mode='mean', squeeze=True)
    t7 = time.time()
    print('pinned %d | mca_atm_3d %.1f ms | nml+scene+jobs %.1f ms | upload (H2D %.0f MB + pack) %.1f ms | run+sync %.1f ms (kernel %.1f) | read %.1f ms '
          '|| whole mcarats_ng %.1f ms | mca_out_ng %.1f ms' % (pinned, 1e3 * (t1 - t0), 1e3 * (t2 - t1), m.scene.h2d_bytes() / 1e6 if hasattr(m.scene, 'h2d_bytes') else -1,
                                                              1e3 * (t3 - t2), 1e3 * (t4 - t3), sol.stats()['elapsed_ms'], 1e3 * (t5 - t4), 1e3 * (t6 - t5), 1e3 * (t7 - t6)), flush=True)
