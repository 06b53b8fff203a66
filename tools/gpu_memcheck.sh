#!/bin/bash
# compute-sanitizer (memcheck, then racecheck) over one small invocation of every kernel specialisation family:
# 3-D radiance + flux, oblique local estimates, column-frozen (IPA), all-sky camera, plane-parallel with block-private tallies.
# usage (under gpurun): bash tools/gpu_memcheck.sh TAG
TAG=${1:-x}
mkdir -p gpurun_out
cat > /tmp/memcheck_run.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), 'tests'))
import numpy as np
from er3t_b200 import abi
from er3t_b200.solver import Solver
import scenes
nphot = int(sys.argv[1])
s = Solver(device=0)
cam = [dict(kind=1, the=0.0, phi=0.0, zloc=0.0, nxr=8, nyr=8, xpos=0.5, ypos=0.5, qmax=120.0, umax=120.0, vmax=120.0, apsize=0.05)]
cases = [('3-D rad+flux', scenes.scene_3d(nz3=6, dz3=100.0, nlay=14), dict(target=abi.TARGET_RADIANCE | abi.TARGET_FLUX)),
         ('3-D oblique', scenes.scene_3d(nz3=6, dz3=100.0, nlay=14, sfc='dsm', sensors=[dict(the=140.0, phi=30.0, nxr=16, nyr=12)]), dict(target=abi.TARGET_RADIANCE)),
         ('IPA', scenes.scene_3d(), dict(target=abi.TARGET_RADIANCE, solver=abi.SOLVER_IPA)),
         ('camera', scenes.scene_3d(sensors=cam, clear_below=True), dict(target=abi.TARGET_RADIANCE)),
         ('plane-parallel', scenes.plane_parallel(absorb=True)[0], dict(target=abi.TARGET_FLUX | abi.TARGET_HEATING | abi.TARGET_RADIANCE))]
for name, sc, o in cases:
    opt = abi.make_options(nslab=2, wmin=0.2, **o)
    jobs, keep = scenes.multi_seed_jobs(nphot, 2)
    s.upload_scene(sc, opt); s.run(jobs)
    st = s.results()['stats']
    print(name, 'photons', st['photons'], 'coll', st['n_coll'], flush=True)
s.close()
PY
timeout 150 compute-sanitizer --tool memcheck --print-limit 20 python /tmp/memcheck_run.py 3000 > gpurun_out/memcheck_$TAG.log 2>&1
echo "exit $?" >> gpurun_out/memcheck_$TAG.log
timeout 150 compute-sanitizer --tool racecheck --print-limit 20 python /tmp/memcheck_run.py 600 > gpurun_out/racecheck_$TAG.log 2>&1
echo "exit $?" >> gpurun_out/racecheck_$TAG.log
tail -9 gpurun_out/memcheck_$TAG.log; tail -9 gpurun_out/racecheck_$TAG.log
