#!/usr/bin/env python
"""
The config-4 sweep (O2 A-band: 8 wavelengths x 8 ... 16 g x 3 runs over ONE 768 x 960 x 5 scene, workloads.c4,
projects/01_oco2_rad-sim.py:70-73 shape) sharded over the GPUs of one node, two ways:

  photons  every wavelength's photons are split over the ranks (index % world == rank, the library's shard options);
           one in-place NCCL all-reduce of the radiance tallies per wavelength (17.7 MB for 3 slabs);
  calls    whole wavelengths are handed to ranks, longest-processing-time first (er3t's `rearrange_jobs` idea,
           er3t/rtm/mca/mca_run.py:184-230, applied to calls instead of processes); no collective on the data path, the
           result of a wavelength stays with the rank that traced it (the reference writes one file per job).

    python -m torch.distributed.run --nproc-per-node N tools/c4_sweep.py [--scale 1.0] [--reps 2]

Timing: barrier + cuda synchronize on both sides, max over ranks; rank 0 prints one JSON line.  Also importable:
bench.py calls `sweep_both()` as its `c4_sweep` leg when N > 1.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT)
import numpy as np


def lpt_assign(costs, world):
    """Longest processing time first: indices of `costs` per rank."""
    order = sorted(range(len(costs)), key=lambda i: -costs[i])
    load = [0.0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: load[k])
        out[r].append(i)
        load[r] += costs[i]
    return out


def sweep_both(sol, rank, world, scale=1.0, reps=2, photons=1e9):
    import torch
    import torch.distributed as dist
    import workloads
    from er3t_b200 import abi, dist as edist
    from er3t_b200.rtm.mca import mcarats_ng

    pairs = workloads.c4(scale=scale, photons=photons)

    def prep(kw, shard):
        m = mcarats_ng(**dict(kw, dry_run=True, shard=shard))
        jobs, keep = abi.make_jobs(**m.jobs_args)
        return m, jobs, keep

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        fn()                                   # warm-up
        best = None
        for _ in range(reps):
            barrier()
            t0 = time.time()
            fn()
            barrier()
            dt = time.time() - t0
            if world > 1:
                t = torch.tensor([dt], dtype=torch.float64, device='cuda')
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            best = dt if best is None else min(best, dt)
        return best

    nphot_total = 0.0
    # ---- photons of every call split over the ranks
    preps = [prep(kw, (rank, world)) for kw, _ in pairs]
    nphot_total = float(sum(np.sum(m.jobs_args['nphot']) for m, _, _ in preps))
    sol.upload_scene(preps[0][0].scene, preps[0][0].options)
    kms = []

    def run_photons():
        del kms[:]
        for m, jobs, _ in preps:
            sol.run(jobs, sync=False)
            if world > 1:
                edist.allreduce_results(sol, to_host=False)
            else:
                sol.sync()
            kms.append(sol.stats()['elapsed_ms'])
        sol.results()                          # the last wavelength's tallies on the host (every rank holds the sum)
    t_ph = timed(run_photons)
    k_ph = float(sum(kms))

    # ---- whole calls per rank, longest first
    mine = lpt_assign([float(np.sum(m.jobs_args['nphot'])) for m, _, _ in preps], world)[rank]
    preps_c = [prep(pairs[i][0], (0, 1)) for i in mine]
    if preps_c:
        sol.upload_scene(preps_c[0][0].scene, preps_c[0][0].options)
    kms_c = []

    def run_calls():
        del kms_c[:]
        for m, jobs, _ in preps_c:
            sol.run(jobs)
            kms_c.append(sol.stats()['elapsed_ms'])
            sol.results()
    t_ca = timed(run_calls)
    k_ca = float(sum(kms_c))
    if world > 1:
        t = torch.tensor([k_ca, -k_ca], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        k_max, k_min = float(t[0].item()), -float(t[1].item())
    else:
        k_max = k_min = k_ca
    sc = preps[0][0].scene.struct
    return {'workload': 'C4 O2 A-band sweep: %d wavelengths x 8 ... 16 g x 3 runs, one %d x %d x %d scene, %.3g photons in total' % (
                len(pairs), sc.nx, sc.ny, sc.nz3, nphot_total),
            'photons': nphot_total, 'n_gpus': world,
            'photon_sharding': {'value': nphot_total / t_ph, 'unit': 'photons/s', 's_per_sweep': t_ph, 'kernel_ms_rank0': k_ph,
                                'collective': 'one in-place NCCL all-reduce of the radiance tallies per wavelength'},
            'whole_calls_lpt': {'value': nphot_total / t_ca, 'unit': 'photons/s', 's_per_sweep': t_ca, 'kernel_ms_max_rank': k_max,
                                'kernel_ms_min_rank': k_min, 'collective': 'none (results stay with the tracing rank)'}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--scale', type=float, default=1.0)
    ap.add_argument('--reps', type=int, default=2)
    ap.add_argument('--photons', type=float, default=1e9)
    ap.add_argument('--out', default='')
    a = ap.parse_args()
    import torch
    from er3t_b200 import dist as edist
    from er3t_b200.solver import Solver
    rank, world, local = edist.init_from_env()
    torch.cuda.set_device(local)
    sol = Solver(device=local)
    res = sweep_both(sol, rank, world, a.scale, a.reps, a.photons)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    if rank == 0:
        print(json.dumps(res))
        if a.out:
            os.makedirs(os.path.dirname(a.out), exist_ok=True)
            json.dump(res, open(a.out, 'w'), indent=1)


if __name__ == '__main__':
    main()
