"""Sweep super-voxel sizes / launch shapes on the C2 workload; prints photons/s and events per photon."""
import sys, os, time, json
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from er3t_b200 import abi
from er3t_b200.solver import Solver
from er3t_b200.rtm.mca import mcarats_ng
nz3 = int(os.environ.get('NZ3', '100'))
phot = float(os.environ.get('PHOT', '2e7'))
kw, abs0 = bench.build_workload(480, 480, nz3, phot, nrun=1)
sol = Solver(0)
configs = [tuple(int(v) for v in c.split(',')) for c in sys.argv[1:]]
for cfg in configs:
    sv = cfg[:3]
    cm = cfg[3:6] if len(cfg) >= 6 else (0, 0, 0)
    K = cfg[6] if len(cfg) > 6 else 0
    E = cfg[7] if len(cfg) > 7 else 0
    R = cfg[8] if len(cfg) > 8 else 0
    tpb = cfg[9] if len(cfg) > 9 else 0
    bps = cfg[10] if len(cfg) > 10 else 0
    m = mcarats_ng(**dict(kw, dry_run=True, supervoxel=sv))
    m.options.cmx, m.options.cmy, m.options.cmz = cm
    m.options.flight_steps = K
    m.options.event_min = E
    m.options.threads_per_block = tpb
    m.options.blocks_per_sm = bps
    jobs, keep = abi.make_jobs(**m.jobs_args)
    sol.upload_scene(m.scene, m.options)
    sol.run(jobs); sol.run(jobs)
    st = sol.stats()
    n = st['photons']
    print('sv=%s cm=%s K=%d E=%d R=%d tpb=%d bps=%d : %.1f Mph/s | per photon: cell %.1f tent %.1f coll %.1f sfc %.2f le %.1f visit %.1f | bytes/ph %.0f' % (
        sv, cm, K, E, R, tpb, bps, n / st['elapsed_ms'] / 1e3, st['n_cell'] / n, st['n_tent'] / n, st['n_coll'] / n, st['n_sfc'] / n, st['n_le'] / n,
        st['n_le_visit'] / n, st['bytes_alg'] / n), flush=True)
