#!/usr/bin/env python
"""
Throughput of the five BASELINE.json configs (workloads.py) at their NAMED shapes on one B200:
photons/s of the transport launch with the scene resident in HBM (b200rt_stats elapsed, CUDA events inside the
library), algorithmic bytes per photon, events per photon and the energy balance of every run.

    python tools/bench_configs.py [--scale 1.0] [--reps 3] [--out gpurun_out/configs.json]

C4 is a list of wavelengths sharing one scene: its line sums photons and kernel time over the sweep.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT)
import numpy as np

import workloads
from er3t_b200 import abi
from er3t_b200.solver import Solver
from er3t_b200.rtm.mca import mcarats_ng


def run_one(sol, kw, reps, pf=1.0, pool=0, smem_tally=0):
    prep = mcarats_ng(**dict(kw, dry_run=True, photons=kw['photons'] * pf))
    prep.options.pool_slots = pool
    prep.options.smem_tally = smem_tally
    jobs, keep = abi.make_jobs(**prep.jobs_args)
    t0 = time.time()
    sol.upload_scene(prep.scene, prep.options)
    t_up = time.time() - t0
    sol.run(jobs)                      # warm-up
    ms, st = [], None
    for _ in range(reps):
        sol.run(jobs)
        st = sol.stats()
        ms.append(st['elapsed_ms'])
    n = float(st['photons'])
    bal = (st['w_toa_up'] + st['w_sfc_abs'] + st['w_atm_abs'] - st['w_roulette']) / n - 1.0
    shape = 'plane-parallel' if prep.scene.struct.nz3 <= 0 else '%dx%dx%d' % (prep.scene.struct.nx, prep.scene.struct.ny, prep.scene.struct.nz3)
    return dict(photons=n, kernel_ms=float(np.median(ms)), upload_ms=1e3 * t_up, bytes_alg=float(st['bytes_alg']), balance=bal, shape=shape,
                njob=len(jobs), nrad=int(prep.scene.struct.nrad),
                per_photon={k: st[k] / n for k in ('n_cell', 'n_tent', 'n_coll', 'n_sfc', 'n_le', 'n_le_visit', 'n_tally')})


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--scale', type=float, default=1.0)
    ap.add_argument('--reps', type=int, default=3)
    ap.add_argument('--pool', type=int, default=0, help='photon slots per warp (0 = auto)')
    ap.add_argument('--smem-tally', type=int, default=0, help='-1 global atomics only, 0 auto, 2 one private copy per warp')
    ap.add_argument('--photon-factor', type=float, default=1.0, help='multiply the photon count only (ncu captures)')
    ap.add_argument('--configs', default='C1,C2,C3,C4,C5')
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'configs.json'))
    a = ap.parse_args()
    sol = Solver(device=0)
    rows = []
    for name in a.configs.split(','):
        built = workloads.build(name, scale=a.scale)
        pairs = built if isinstance(built, list) else [built]
        parts = [run_one(sol, kw, a.reps, a.photon_factor, a.pool, a.smem_tally) for kw, _ in pairs]
        kw0 = pairs[0][0]
        n = sum(p['photons'] for p in parts)
        ms = sum(p['kernel_ms'] for p in parts)
        b = sum(p['bytes_alg'] for p in parts)
        peak = 6461.5
        try:
            peak = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
        except Exception:
            pass
        row = dict(config=name, target=kw0['target'], shape=parts[0]['shape'], calls=len(parts), jobs_per_call=parts[0]['njob'], sensors=parts[0]['nrad'], photons=n, kernel_ms=ms,
                   mphotons_per_s=n / ms / 1e3, bytes_alg_per_photon=b / n, gbs_alg=b / ms / 1e6, roofline_frac_hbm=b / ms / 1e6 / peak,
                   mphotons_per_s_incl_upload=n / (ms + sum(p['upload_ms'] for p in parts)) / 1e3,
                   max_abs_balance=max(abs(p['balance']) for p in parts), upload_ms=sum(p['upload_ms'] for p in parts),
                   per_photon=parts[0]['per_photon'])
        rows.append(row)
        print(json.dumps(row), flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(rows, open(a.out, 'w'), indent=1)


if __name__ == '__main__':
    main()
