"""Sweep the launch shape of the warp-pool transport kernel on the C2 workload.
usage: sweep_pool.py LIB:NP,TPB,BPS,K,E ...   (LIB = path of a libb200rt build or '-' for the in-tree one)"""
import sys, os
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT)
import bench
from er3t_b200 import abi
from er3t_b200.solver import Solver
from er3t_b200.rtm.mca import mcarats_ng
nz3 = int(os.environ.get('NZ3', '100'))
phot = float(os.environ.get('PHOT', '2e7'))
kw, abs0 = bench.build_workload(480, 480, nz3, phot, nrun=1)
m = mcarats_ng(**dict(kw, dry_run=True))
jobs, keep = abi.make_jobs(**m.jobs_args)
sols = {}
for arg in sys.argv[1:]:
    libp, cfg = arg.split(':')
    NP, tpb, bps, K, E = [int(v) for v in cfg.split(',')]
    if libp not in sols:
        sols[libp] = Solver(0, lib=abi.load_library(None if libp == '-' else libp))
    sol = sols[libp]
    m.options.pool_slots, m.options.threads_per_block, m.options.blocks_per_sm = NP, tpb, bps
    m.options.flight_steps, m.options.event_min = K, E
    try:
        sol.upload_scene(m.scene, m.options)
        sol.run(jobs); sol.run(jobs)
    except OSError as e:
        print('%s NP=%d tpb=%d bps=%d K=%d E=%d : FAILED %s' % (libp, NP, tpb, bps, K, E, e), flush=True)
        continue
    st = sol.stats()
    n = st['photons']
    print('%s NP=%d tpb=%d bps=%d K=%d E=%d : %.1f Mph/s | per photon: cell %.1f tent %.1f coll %.1f sfc %.2f le %.1f' % (
        libp, NP, tpb, bps, K, E, n / st['elapsed_ms'] / 1e3, st['n_cell'] / n, st['n_tent'] / n, st['n_coll'] / n, st['n_sfc'] / n, st['n_le'] / n), flush=True)
