"""bench.py contract (CPU side): the reference arm prints exactly ONE JSON line on stdout with the keys the driver reads,
also when a launcher exports OMP_NUM_THREADS=1 (torchrun does); without a GPU the CUDA arm fails loudly."""

import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))


def _run(args, env_extra=None):
    env = dict(os.environ)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + args, capture_output=True, text=True, env=env, cwd=ROOT, timeout=600)


def test_reference_arm_prints_one_json_line():
    r = _run(['--impl', 'reference', '--gpus', '1', '--steps', '1', '--warmup', '1', '--photons', '5e4', '--nx', '32', '--nz3', '8'],
             {'OMP_NUM_THREADS': '1'})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'photons/s' and d['unit'] == 'photons/s' and d['higher_is_better'] is True
    assert d['value'] > 0 and d['steps'] == 1 and d['warmup'] == 1 and d['n_gpus'] == 1
    assert d['e2e'] == {'value': d['value'], 'unit': 'photons/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['value'] == d['value'] and cb['cores'] >= 1 and 'photons' in cb['sample']
    # every host thread is used although the launcher exported OMP_NUM_THREADS=1
    assert cb['cores'] == len(os.sched_getaffinity(0))
    assert 'workload' in d['config'] and 'model' not in d['config']


def test_reference_arm_other_ranks_print_nothing():
    r = _run(['--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '1', '--photons', '5e4', '--nx', '32', '--nz3', '8'],
             {'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'})
    assert r.returncode == 0 and r.stdout.strip() == ''


def test_cuda_arm_fails_loudly_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    r = _run(['--gpus', '1', '--steps', '1', '--warmup', '1', '--photons', '1e4', '--nx', '32', '--nz3', '8'])
    assert r.returncode != 0
    assert 'no CUDA device' in r.stderr or 'no CPU fallback' in r.stderr
