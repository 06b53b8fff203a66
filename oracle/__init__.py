"""
oracle/ -- TEST INFRASTRUCTURE ONLY (never imported by er3t_b200/).

* oracle_mc.cpp      fp64 CPU Monte Carlo restatement of the photon-transport path (PARITY UNPINNED
                     against MCARaTS, see the header of that file and DESIGN.md).
* adding_doubling.py deterministic plane-parallel solver (adding-doubling; tabulated phase functions, oblique views by
                     azimuthal Fourier modes) that pins oracle_mc AND, through tests/golden/ad_fixtures.npz, the CUDA path.
* philox_np.py       numpy Philox4x32-10 checked against the Random123 known-answer vectors.
(The reference's host-side arithmetic -- distribute_photon, cal_mca_azimuth, output weighting ... -- needs no restatement
here: tests/golden/make_golden.py runs the reference's own functions and commits their outputs.)

Only tests/, __graft_entry__.smoke() and the cpu_baseline / `--impl reference` legs of bench.py may import this.
"""

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, '_build', 'liboracle_mc.so')
_LIB = None


def build(force=False):
    """Compile oracle_mc.cpp with the committed Makefile (g++ only)."""
    deps = [os.path.join(_HERE, 'oracle_mc.cpp'), os.path.join(_HERE, '..', 'include', 'b200rt.h')]      # the ABI structs are shared
    if force or not os.path.isfile(_SO) or any(os.path.getmtime(_SO) < os.path.getmtime(d) for d in deps if os.path.isfile(d)):
        subprocess.check_call(['make', '-C', _HERE, '-s'] + (['-B'] if force else []))
    return _SO


def load():
    global _LIB
    if _LIB is None:
        if not os.path.isfile(_SO):
            build()
        lib = C.CDLL(_SO)
        vp = C.c_void_p
        lib.oracle_run.argtypes = [vp, vp, vp, C.c_int, vp, vp, vp, vp, C.c_int]
        lib.oracle_run.restype = C.c_int
        lib.oracle_philox_fill.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, vp, C.c_int64]
        lib.oracle_philox_fill.restype = None
        lib.oracle_phase_eval.argtypes = [vp, C.c_double, vp, vp, C.c_int64]
        lib.oracle_phase_sample.argtypes = [vp, C.c_double, vp, vp, C.c_int64]
        lib.oracle_brdf_eval.argtypes = [C.c_int32, vp, vp, vp, vp, C.c_int64]
        _LIB = lib
    return _LIB


def run(scene, options, jobs, nthreads=0):
    """
    Trace `jobs` on the CPU.

    scene   : er3t_b200.abi.HostScene (host arrays)
    options : er3t_b200.abi.Options
    jobs    : ctypes array of er3t_b200.abi.Job
    returns : dict(flux=..., rad=..., heat=..., stats=...)   (arrays are None when not in `target`)
    """
    from er3t_b200 import abi
    lib = load()
    nslab = options.nslab
    flux = rad = heat = None
    if options.target & abi.TARGET_FLUX:
        flux = np.zeros(scene.flux_shape(nslab), dtype=np.float64)
    if (options.target & abi.TARGET_RADIANCE) and scene.struct.nrad > 0:
        rad = np.zeros(scene.rad_size(nslab), dtype=np.float64)
    if options.target & abi.TARGET_HEATING:
        heat = np.zeros(scene.heat_shape(nslab), dtype=np.float64)
    st = abi.Stats()
    rc = lib.oracle_run(C.addressof(scene.struct), C.addressof(options), C.addressof(jobs), len(jobs),
                        None if flux is None else flux.ctypes.data,
                        None if rad is None else rad.ctypes.data,
                        None if heat is None else heat.ctypes.data,
                        C.addressof(st), int(nthreads))
    if rc != 0:
        raise OSError('Error [oracle]: oracle_run failed with code %d.' % rc)
    return {'flux': flux, 'rad': rad, 'heat': heat, 'stats': st.as_dict()}


def philox(seed, first, n, c2=0, c3=0):
    out = np.zeros(4 * n, dtype=np.uint32)
    load().oracle_philox_fill(int(seed), int(first), int(c2), int(c3), out.ctypes.data, int(n))
    return out.reshape(n, 4)
