"""
Thin object wrapper over the C-ABI (include/b200rt.h): one Solver == one handle == one GPU.

This is what `mcarats_ng` drives instead of `mca_run` (er3t/rtm/mca/mca_run.py:41-181).  Errors from the
library surface as OSError carrying the library's message, in the reference's message style
(`'Error [mcarats_ng]: ...'`, er3t/rtm/mca/mcarats.py:144-145,481-483).
"""

import ctypes as C

import numpy as np

from . import abi

__all__ = ['Solver']


class Solver:

    def __init__(self, device=0, lib=None):
        self.lib = lib if lib is not None else abi.load_library()
        self.handle = C.c_void_p()
        rc = self.lib.b200rt_create(C.byref(self.handle), int(device))
        if rc != 0:
            self.handle = None
            raise OSError('Error [er3t_b200]: cannot create a solver on CUDA device %d (code %d). A B200 GPU is required; there is no CPU fallback.' % (device, rc))
        self.device = device
        self.scene = None
        self.options = None
        self._keep = None

    # ------------------------------------------------------------------ helpers
    def _check(self, rc, where):
        if rc != 0:
            msg = self.lib.b200rt_last_error(self.handle)
            msg = msg.decode() if msg else ''
            raise OSError('Error [er3t_b200]: %s failed (code %d): %s' % (where, rc, msg))

    def close(self):
        if getattr(self, 'handle', None):
            self.lib.b200rt_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ API
    def upload_scene(self, scene, options):
        """scene: abi.HostScene (or any object with `.struct`); options: abi.Options"""
        self._check(self.lib.b200rt_upload_scene(self.handle, C.byref(scene.struct), C.byref(options)), 'b200rt_upload_scene')
        self.scene = scene
        self.options = options

    def run(self, jobs, accumulate=False, stream=None, sync=True):
        """jobs: ctypes array of abi.Job (see abi.make_jobs)."""
        self._keep = jobs
        self._check(self.lib.b200rt_run(self.handle, jobs, len(jobs), 1 if accumulate else 0, stream), 'b200rt_run')
        if sync:
            self.sync()

    def sync(self):
        self._check(self.lib.b200rt_sync(self.handle), 'b200rt_sync')

    def stats(self):
        st = abi.Stats()
        self._check(self.lib.b200rt_stats_get(self.handle, C.byref(st)), 'b200rt_stats_get')
        return st.as_dict()

    def tally_ptrs(self):
        vp = C.c_void_p
        f, r, h = vp(), vp(), vp()
        nf, nr, nh = C.c_int64(), C.c_int64(), C.c_int64()
        self._check(self.lib.b200rt_tally_ptrs(self.handle, C.byref(f), C.byref(nf), C.byref(r), C.byref(nr), C.byref(h), C.byref(nh)),
                    'b200rt_tally_ptrs')
        return {'flux': (f.value, nf.value), 'rad': (r.value, nr.value), 'heat': (h.value, nh.value)}

    def read_flux(self, out=None):
        shape = self.scene.flux_shape(self.options.nslab)
        if out is None:
            out = np.empty(shape, dtype=np.float64)
        self._check(self.lib.b200rt_read_flux(self.handle, out.ctypes.data, out.size), 'b200rt_read_flux')
        return out

    def read_rad(self, out=None):
        n = self.scene.rad_size(self.options.nslab)
        if out is None:
            out = np.empty(n, dtype=np.float64)
        self._check(self.lib.b200rt_read_rad(self.handle, out.ctypes.data, out.size), 'b200rt_read_rad')
        return out

    def read_heat(self, out=None):
        shape = self.scene.heat_shape(self.options.nslab)
        if out is None:
            out = np.empty(shape, dtype=np.float64)
        self._check(self.lib.b200rt_read_heat(self.handle, out.ctypes.data, out.size), 'b200rt_read_heat')
        return out

    def results(self):
        """Everything the target asked for, as numpy arrays (same keys as oracle.run)."""
        res = {'flux': None, 'rad': None, 'heat': None}
        t = self.options.target
        if t & abi.TARGET_FLUX:
            res['flux'] = self.read_flux()
        if (t & abi.TARGET_RADIANCE) and self.scene.struct.nrad > 0:
            res['rad'] = self.read_rad()
        if t & abi.TARGET_HEATING:
            res['heat'] = self.read_heat()
        res['stats'] = self.stats()
        return res

    # diagnostics
    def philox(self, seed, first, n, c2=0, c3=0):
        out = np.zeros(4 * n, dtype=np.uint32)
        self._check(self.lib.b200rt_philox_fill(self.handle, int(seed), int(first), int(c2), int(c3), out.ctypes.data, int(n)), 'b200rt_philox_fill')
        return out.reshape(n, 4)

    def phase_eval(self, apf, mu):
        mu = np.ascontiguousarray(mu, dtype=np.float64)
        out = np.empty_like(mu)
        self._check(self.lib.b200rt_phase_eval(self.handle, float(apf), mu.ctypes.data, out.ctypes.data, mu.size), 'b200rt_phase_eval')
        return out

    def phase_sample(self, apf, xi):
        xi = np.ascontiguousarray(xi, dtype=np.float64)
        out = np.empty_like(xi)
        self._check(self.lib.b200rt_phase_sample(self.handle, float(apf), xi.ctypes.data, out.ctypes.data, xi.size), 'b200rt_phase_sample')
        return out

    def brdf_eval(self, sfc_type, param5, dir_in, dir_out):
        p = np.ascontiguousarray(param5, dtype=np.float32)
        di = np.ascontiguousarray(dir_in, dtype=np.float64).reshape(-1, 3)
        do = np.ascontiguousarray(dir_out, dtype=np.float64).reshape(-1, 3)
        out = np.empty(di.shape[0], dtype=np.float64)
        self._check(self.lib.b200rt_brdf_eval(self.handle, int(sfc_type), p.ctypes.data, di.ctypes.data, do.ctypes.data, out.ctypes.data, out.size),
                    'b200rt_brdf_eval')
        return out
