"""The C-ABI: the built library exports every symbol include/b200rt.h declares, the ctypes mirror has the same struct
layout as the C header (checked by compiling a probe with gcc), and missing libraries fail loudly.  No compute calls
(no GPU here)."""

import ctypes as C
import os
import re
import subprocess

import pytest

from er3t_b200 import abi

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
HEADER = os.path.join(ROOT, 'include', 'b200rt.h')


def test_header_declares_what_the_mirror_lists():
    txt = open(HEADER).read()
    declared = set(re.findall(r'\b(b200rt_[a-z_]+)\s*\(', txt))
    assert declared == set(abi.EXPORTS)


def test_library_exports_every_symbol():
    p = abi.library_path()
    if not os.path.isfile(p):
        import __graft_entry__ as g
        g.build()
    lib = C.CDLL(p)
    for name in abi.EXPORTS:
        assert hasattr(lib, name), name
    lib.b200rt_version.restype = C.c_int
    assert lib.b200rt_version() == 103


def test_struct_layout_matches_header(tmp_path):
    probe = tmp_path / 'probe.c'
    probe.write_text('''
#include <stdio.h>
#include <stddef.h>
#include "b200rt.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu\\n", sizeof(b200rt_sensor), sizeof(b200rt_scene), sizeof(b200rt_job), sizeof(b200rt_options), sizeof(b200rt_stats));
  printf("%zu %zu %zu %zu\\n", offsetof(b200rt_scene, zgrd), offsetof(b200rt_scene, sfc_param), offsetof(b200rt_scene, src_the), offsetof(b200rt_scene, sensors));
  printf("%zu %zu %zu\\n", offsetof(b200rt_job, abs1d), offsetof(b200rt_job, rad_scale), offsetof(b200rt_options, wmin));
  printf("%zu %zu\\n", offsetof(b200rt_stats, w_toa_up), offsetof(b200rt_stats, launches));
  return 0;
}''')
    exe = tmp_path / 'probe'
    subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), str(probe), '-o', str(exe)])
    out = subprocess.check_output([str(exe)]).decode().split()
    got = [int(v) for v in out]
    S = abi.SceneStruct
    exp = [C.sizeof(abi.Sensor), C.sizeof(S), C.sizeof(abi.Job), C.sizeof(abi.Options), C.sizeof(abi.Stats),
           S.zgrd.offset, S.sfc_param.offset, S.src_the.offset, S.sensors.offset,
           abi.Job.abs1d.offset, abi.Job.rad_scale.offset, abi.Options.wmin.offset,
           abi.Stats.w_toa_up.offset, abi.Stats.launches.offset]
    assert got == exp


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setenv('ER3T_B200_LIB', str(tmp_path / 'nope.so'))
    with pytest.raises(OSError, match='no CPU fallback'):
        abi.load_library(path=abi.library_path())


def test_solver_without_gpu_raises():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from er3t_b200.solver import Solver
    with pytest.raises(OSError, match='no CPU fallback'):
        Solver(device=0)
