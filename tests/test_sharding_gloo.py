"""Multi-process path on CPU: two gloo ranks, photons sharded by index, one all-reduce of the tallies.
The union of the shards is exactly the single-process photon set (counter-based streams), so the reduced result must
equal the one-process result up to floating-point summation order."""

import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import torch.distributed as dist
    import scenes
    from er3t_b200 import abi, dist as edist
    from oracle_solver import OracleSolver
    edist.init_from_env(backend='gloo')
    sc = scenes.scene_3d(nx=8, ny=6)
    opt = abi.make_options(target=abi.TARGET_RADIANCE | abi.TARGET_FLUX, nslab=2, wmin=0.2, shard_rank=rank, shard_world=world)
    jobs, keep = scenes.multi_seed_jobs(20001, 2)
    s = OracleSolver()
    s.upload_scene(sc, opt)
    s.run(jobs)
    res = edist.allreduce_results(s)
    if rank == 0:
        np.savez(out, rad=res['rad'], flux=res['flux'], photons=res['stats']['photons'])
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_equals_single_process(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    import oracle
    import scenes
    from er3t_b200 import abi
    out = str(tmp_path / 'r.npz')
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    sc = scenes.scene_3d(nx=8, ny=6)
    opt = abi.make_options(target=abi.TARGET_RADIANCE | abi.TARGET_FLUX, nslab=2, wmin=0.2)
    jobs, keep = scenes.multi_seed_jobs(20001, 2)
    ref = oracle.run(sc, opt, jobs)
    assert int(got['photons']) == 2 * 20001
    assert np.allclose(got['rad'], ref['rad'], rtol=1e-10, atol=1e-14)
    assert np.allclose(got['flux'], ref['flux'], rtol=1e-10, atol=1e-14)


def test_lpt_assignment_of_whole_calls():
    """tools/c4_sweep.py hands whole mcarats_ng calls (wavelengths) to ranks longest-first -- the idea of the reference's
    `rearrange_jobs` (er3t/rtm/mca/mca_run.py:184-230) applied to calls: every call lands on exactly one rank and the
    heaviest rank carries at most the lightest one's load plus one call."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tools'))
    from c4_sweep import lpt_assign
    rng = np.random.default_rng(4)
    for world in (1, 2, 3, 8):
        costs = list(rng.uniform(1.0, 5.0, 11))
        parts = lpt_assign(costs, world)
        assert sorted(i for p in parts for i in p) == list(range(len(costs)))
        loads = [sum(costs[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(costs) + 1e-12
    assert lpt_assign([1.0] * 8, 8) == [[i] for i in range(8)]
