"""Where does the end-to-end time of mcarats_ng + mca_out_ng go? (host packing / H2D / kernel / D2H / weighting)"""
import sys, os, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from er3t_b200 import abi
from er3t_b200.solver import Solver
from er3t_b200.rtm.mca import mcarats_ng, mca_out_ng
kw, abs0 = bench.build_workload(480, 480, 100, float(os.environ.get('PHOT', '1e8')))
sol = Solver(0)
for it in range(3):
    t0 = time.time()
    m = mcarats_ng(**dict(kw, dry_run=True))
    t1 = time.time()
    jobs, keep = abi.make_jobs(**m.jobs_args)
    sol.upload_scene(m.scene, m.options)
    t2 = time.time()
    sol.run(jobs)
    t3 = time.time()
    res = sol.results()
    t4 = time.time()
    m2 = mcarats_ng(**dict(kw, solver_obj=sol))
    t5 = time.time()
    out = mca_out_ng(mca_obj=m2, abs_obj=abs0)
    t6 = time.time()
    print('build nml+scene %.0f ms | upload %.0f ms | run+sync %.0f ms (kernel %.0f) | read %.0f ms || full mcarats_ng %.0f ms | mca_out_ng %.0f ms' % (
        1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2), sol.stats()['elapsed_ms'], 1e3 * (t4 - t3), 1e3 * (t5 - t4), 1e3 * (t6 - t5)), flush=True)
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
m = mcarats_ng(**dict(kw, dry_run=True))
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(12)
