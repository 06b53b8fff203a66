"""Multi-GPU path on real devices: two NCCL ranks, photons of every job sharded by index, ONE in-place
all-reduce of the library-owned tallies.  The union of the shards is exactly the one-GPU photon set (counter-based Philox
streams), so the reduced tallies must equal the one-GPU tallies up to fp64 summation order.  Skipped on boxes with one GPU."""

import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
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
    from er3t_b200.solver import Solver
    edist.init_from_env(backend='nccl')
    sc = scenes.scene_3d(nx=16, ny=12)
    opt = abi.make_options(target=abi.TARGET_RADIANCE | abi.TARGET_FLUX, nslab=2, wmin=0.2, shard_rank=rank, shard_world=world)
    jobs, keep = scenes.multi_seed_jobs(300001, 2)
    s = Solver(device=rank)
    s.upload_scene(sc, opt)
    s.run(jobs, sync=False)
    res = edist.allreduce_results(s)               # waits for the run, reduces in place on b200rt_tally_ptrs, reads back
    # the handle itself now holds the global tallies
    again = s.read_rad()
    assert np.array_equal(again, res['rad'])
    if rank == 0:
        np.savez(out, rad=res['rad'], flux=res['flux'], photons=res['stats']['photons'], n_coll=res['stats']['n_coll'])
    dist.barrier()
    dist.destroy_process_group()
    s.close()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_two_nccl_ranks_equal_one_gpu(solver, tmp_path):
    import scenes
    from er3t_b200 import abi
    out = str(tmp_path / 'r.npz')
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    sc = scenes.scene_3d(nx=16, ny=12)
    opt = abi.make_options(target=abi.TARGET_RADIANCE | abi.TARGET_FLUX, nslab=2, wmin=0.2)
    jobs, keep = scenes.multi_seed_jobs(300001, 2)
    solver.upload_scene(sc, opt)
    solver.run(jobs)
    ref = solver.results()
    assert int(got['photons']) == 2 * 300001 == ref['stats']['photons']
    assert int(got['n_coll']) == ref['stats']['n_coll']                     # the same histories, event for event
    assert np.allclose(got['rad'], ref['rad'], rtol=1e-9, atol=1e-15)
    assert np.allclose(got['flux'], ref['flux'].reshape(got['flux'].shape), rtol=1e-9, atol=1e-15)
