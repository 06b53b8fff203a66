"""
Multi-GPU plumbing: one process per GPU (torchrun), photons of every (run, g) job sharded by
`photon_index % world == rank` (disjoint Philox counters), and ONE NCCL all-reduce (sum, fp64) of the tallies over
NVLink at the end -- the only exchange step this path has.  The reference's counterpart is the
process pool of er3t/rtm/mca/mca_run.py:144-159 (no collective at all).
"""

import os

import numpy as np

__all__ = ['init_from_env', 'allreduce_results', 'shard_from_env']


def shard_from_env():
    return int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))


def init_from_env(backend=None):
    """Initialise torch.distributed from the torchrun environment (no-op for world size 1).  Returns (rank, world, local_rank)."""
    rank, world = shard_from_env()
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1:
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            if backend is None:
                backend = 'nccl' if torch.cuda.is_available() else 'gloo'
            if backend == 'nccl':
                torch.cuda.set_device(local)
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def allreduce_results(solver, to_host=True):
    """
    Sum the tallies of all ranks.  With NCCL the library's device buffers are copied device-to-device into torch
    tensors (b200rt_read_* accepts device destinations), reduced over NVLink, and only then brought to the host; with
    gloo (CPU tests) the host arrays are reduced.  Event counters are reduced with the same collective.
    """
    import torch
    import torch.distributed as dist
    res = {'flux': None, 'rad': None, 'heat': None}
    world = dist.get_world_size() if dist.is_initialized() else 1
    use_cuda = world > 1 and dist.get_backend() == 'nccl'
    ptrs = solver.tally_ptrs() if hasattr(solver, 'tally_ptrs') else None
    readers = {'flux': getattr(solver, 'read_flux', None), 'rad': getattr(solver, 'read_rad', None), 'heat': getattr(solver, 'read_heat', None)}
    host = solver.results() if not use_cuda else None
    st = solver.stats()
    for key in ('flux', 'rad', 'heat'):
        if use_cuda:
            n = ptrs[key][1]
            if n == 0:
                continue
            t = torch.empty(n, dtype=torch.float64, device='cuda')
            rc = solver.lib.b200rt_read_flux if key == 'flux' else (solver.lib.b200rt_read_rad if key == 'rad' else solver.lib.b200rt_read_heat)
            solver._check(rc(solver.handle, t.data_ptr(), n), 'b200rt_read_' + key)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            res[key] = t.cpu().numpy() if to_host else t
        else:
            a = host[key]
            if a is None:
                continue
            if world > 1:
                t = torch.from_numpy(np.ascontiguousarray(a))
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
                a = t.numpy()
            res[key] = a
    if world > 1:
        keys = [k for k in st.keys() if k not in ('elapsed_ms', 'launches')]
        v = torch.tensor([float(st[k]) for k in keys], dtype=torch.float64, device='cuda' if use_cuda else 'cpu')
        dist.all_reduce(v, op=dist.ReduceOp.SUM)
        for k, x in zip(keys, v.cpu().tolist()):
            st[k] = type(st[k])(x)
    if res['flux'] is not None and hasattr(solver, 'scene'):
        res['flux'] = res['flux'].reshape(solver.scene.flux_shape(solver.options.nslab)) if not hasattr(res['flux'], 'is_cuda') else res['flux']
    res['stats'] = st
    return res
