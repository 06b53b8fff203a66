"""
Result assembly with the interface of er3t/rtm/mca/mca_out.py.

`mca_out_raw`  parses one MCARaTS `.bin` + `.ctl` pair (kept so that files written by the compatibility emitter
               `write_mca_out_raw`, or by a real MCARaTS, can be read back);
`mca_out_ng`   same constructor and the same `.data` dictionary (keys, names, units, dims_info, float32 arrays) as the
               reference, filled from the in-memory tallies of `mcarats_ng` instead of Nrun*Ng files.

The g-point weighting `factors[iz, ig] = sol_fac * solar[ig] * weight[ig] * slit[iz, ig] / sum_g(weight * slit[iz])`
(mca_out.py:319-327,444-452) is computed by `cal_factors`; `mcarats_ng` passes it to the GPU so that tallies arrive
already weighted and summed over g (one slab per run).
"""

import os

import numpy as np

from er3t_b200.util import cal_sol_fac

__all__ = ['mca_out_raw', 'mca_out_ng', 'write_mca_out_raw', 'cal_factors', 'read_flux_mca_out', 'read_radiance_mca_out']


def cal_factors(date, abs_obj, Nz, Ng):
    """factors (Nz, Ng) float32 and `toa`, as in read_flux_mca_out / read_radiance_mca_out (mca_out.py:313-327,374)."""
    zz = np.arange(Nz)
    if Nz > 1:
        zz[-1] = zz[-2]                 # the top level reuses the slit function of the top layer
    sol_fac = cal_sol_fac(date)
    weight = abs_obj.coef['weight']['data']
    slit = abs_obj.coef['slit_func']['data']
    solar = abs_obj.coef['solar']['data']
    # vectorised restatement of the reference's double loop (same operation order and float32 roundings: norm is rounded
    # to float32 first, the products are formed in float64 and rounded once)
    weight = np.asarray(weight); solar = np.asarray(solar); slit = np.asarray(slit)
    rows = np.ascontiguousarray(weight[np.newaxis, :Ng] * slit[zz, :Ng]) if weight.size == Ng and slit.shape[1] == Ng else None
    if rows is None:
        norm = np.array([sol_fac / (weight * slit[zz[iz], :]).sum() for iz in range(Nz)], dtype=np.float32)
    else:
        norm = (sol_fac / rows.sum(axis=1)).astype(np.float32)
    factors = (norm.astype(np.float64)[:, np.newaxis] * solar[np.newaxis, :Ng] * weight[np.newaxis, :Ng] * slit[zz, :Ng]).astype(np.float32)
    toa = np.sum(sol_fac * solar * weight)
    return factors, toa


class mca_out_raw:

    """Read one MCARaTS output binary using its GrADS control file (mca_out.py:29-103)."""

    def __init__(self, fname_bin):
        if not os.path.isfile(fname_bin):
            raise OSError('Error [mca_out_raw]: Cannot find <%s>.' % fname_bin)
        fname_ctl = fname_bin + '.ctl'
        if not os.path.isfile(fname_ctl):
            raise OSError('Error [mca_out_raw]: Cannot find <%s>.' % fname_ctl)
        self.fname_bin = fname_bin
        self.fname_ctl = fname_ctl
        self.data = []
        self.read_ctl()
        self.read_bin()

    def read_ctl(self):
        with open(self.fname_ctl, 'r') as f:
            lines = [line.strip() for line in f.readlines()]
        Nx = Ny = Nt = 1
        start = 0
        for i, line in enumerate(lines):
            if 'XDEF' in line:
                Nx = int(line.replace('XDEF', '').replace('LINEAR', '').split()[0])
            elif 'YDEF' in line:
                Ny = int(line.replace('YDEF', '').replace('LINEAR', '').split()[0])
            elif 'TDEF' in line:
                Nt = int(line.replace('TDEF', '').split()[0])
            elif 'VARS' in line and 'ENDVARS' not in line:
                self.Nvar = int(line.replace('VARS', '').strip())
                for j in range(i + 1, i + self.Nvar + 1):
                    words = lines[j].split()
                    Nz = int(words[1])
                    end = start + Nx * Ny * Nz * Nt
                    self.data.append({'name': ' '.join([words[0], '(%s)' % ' '.join(words[3:])]),
                                      'dims': [Nx, Ny, Nz, Nt], 'dims_info': ['Nx', 'Ny', 'Nz', 'Nt'],
                                      'Index_Start': start, 'Index_End': end})
                    start = end

    def read_bin(self, dtype='<f4'):
        raw = np.fromfile(self.fname_bin, dtype=dtype)
        for info in self.data:
            info['data'] = raw[info['Index_Start']:info['Index_End']].reshape(info['dims'], order='F')


def write_mca_out_raw(fname_bin, variables, dx=1.0, dy=1.0):
    """
    Compatibility emitter: write `variables` = [(name, description, array (Nx, Ny, Nz))] as a
    MCARaTS-style `.bin` (float32 LE, Fortran order, variables concatenated) plus `.bin.ctl`, readable by
    `mca_out_raw` here and in the reference.
    """
    Nx, Ny = variables[0][2].shape[:2]
    with open(fname_bin, 'wb') as f:
        for _, _, arr in variables:
            np.asarray(arr).astype('<f4').flatten(order='F').tofile(f)
    with open(fname_bin + '.ctl', 'w') as f:
        f.write('DSET ^%s\n' % os.path.basename(fname_bin))
        f.write('TITLE b200rt output (MCARaTS layout)\n')
        f.write('OPTIONS LITTLE_ENDIAN\n')
        f.write('UNDEF -9.99E33\n')
        f.write('XDEF %d LINEAR 1 1\n' % Nx)
        f.write('YDEF %d LINEAR 1 1\n' % Ny)
        f.write('ZDEF %d LINEAR 1 1\n' % max(v[2].shape[2] for v in variables))
        f.write('TDEF 1 LINEAR 00:00Z01JAN2000 1HR\n')
        f.write('VARS %d\n' % len(variables))
        for name, desc, arr in variables:
            f.write('%s %d 99 %s\n' % (name, arr.shape[2], desc))
        f.write('ENDVARS\n')


def _per_run_arrays(mca_obj, abs_obj, kind):
    """
    Per-run, g-weighted fields as float32 arrays shaped (Nx, Ny, Nz, Nt=1, Nrun).

    Two sources: (1) the fused tallies `mcarats_ng` keeps in memory (already weighted on the GPU with the factors of
    ITS absorption object); (2) raw per-job fields (`mca_obj.raw`, or `.bin` files on disk) weighted here exactly like
    the reference does (float32 accumulation, mca_out.py:344-366,467-481).
    """
    nvar = 3 if kind == 'flux' else 1
    fused = getattr(mca_obj, 'fused', None)
    if fused is not None and fused.get(kind) is not None:
        own_abs = getattr(mca_obj, 'abs', None)
        same = (abs_obj is own_abs) or (own_abs is not None and all(
            np.array_equal(np.asarray(abs_obj.coef[k]['data']), np.asarray(own_abs.coef[k]['data'])) for k in ('weight', 'solar', 'slit_func')))
        if not same:
            raise OSError('Error [mca_out_ng]: <abs_obj> differs from the absorption object the simulation was weighted with; rerun <mcarats_ng> with raw=True.')
        return [np.asarray(a, dtype=np.float32) for a in fused[kind]]
    # raw path
    def load(ir, ig):
        raw = getattr(mca_obj, 'raw', None)
        if raw is not None:
            return raw[ir][ig]
        return [v['data'] for v in mca_out_raw(mca_obj.fnames_out[ir][ig]).data]
    first = load(0, 0)
    Nx, Ny, Nz, Nt = first[0].shape
    factors, _ = cal_factors(mca_obj.date, abs_obj, Nz, mca_obj.Ng)
    out = [np.zeros((Nx, Ny, Nz, Nt, mca_obj.Nrun), dtype=np.float32) for _ in range(nvar)]
    for ir in range(mca_obj.Nrun):
        for ig in range(mca_obj.Ng):
            fields = load(ir, ig)
            for v in range(nvar):
                out[v][..., ir] += np.asarray(fields[v], dtype=np.float32) * factors[None, None, :, None, ig]
    return out


def _finish(arr, squeeze):
    """(Nx, Ny, Nz, Nt, Nrun) -> squeezed array + dims_info, like mca_out.py:333-338."""
    dims_info = ['Nx', 'Ny', 'Nz', 'Nt']
    dims = list(arr.shape[:4])
    if squeeze:
        dims_info = [dims_info[i] for i in range(4) if dims[i] > 1]
        dims = [d for d in dims if d > 1]
    return arr.reshape(dims + [arr.shape[-1]]), dims_info + ['Nr']


def read_flux_mca_out(mca_obj, abs_obj, mode='mean', squeeze=True):
    """Fluxes (W/m^2/nm): f_down, f_up, f_down_direct, f_down_diffuse (+ _std), toa, N_photon, N_run (mca_out.py:283-407)."""
    mode = mode.lower()
    f_down_direct, f_down, f_up = _per_run_arrays(mca_obj, abs_obj, 'flux')
    _, toa = cal_factors(mca_obj.date, abs_obj, 1, mca_obj.Ng)
    f_down_direct, dims_info = _finish(f_down_direct, squeeze)
    f_down, _ = _finish(f_down, squeeze)
    f_up, _ = _finish(f_up, squeeze)
    d = {'toa': {'data': toa, 'name': 'TOA without SZA', 'units': 'W/m^2/nm'}}
    fields = (('f_down', f_down, 'Global downwelling flux'), ('f_up', f_up, 'Global upwelling flux'),
              ('f_down_direct', f_down_direct, 'Direct downwelling flux'),
              ('f_down_diffuse', f_down - f_down_direct, 'Diffuse downwelling flux'))
    if mode == 'all':
        for key, arr, name in fields:
            d[key] = {'data': arr, 'name': name, 'units': 'W/m^2/nm', 'dims_info': dims_info}
    elif mode == 'mean':
        for key, arr, name in fields:
            d[key] = {'data': np.mean(arr, axis=-1), 'name': name + ' (mean)', 'units': 'W/m^2/nm', 'dims_info': dims_info[:-1]}
        for key, arr, name in fields:
            d[key + '_std'] = {'data': np.std(arr, axis=-1), 'name': name + ' (standard deviation)', 'units': 'W/m^2/nm', 'dims_info': dims_info[:-1]}
    else:
        raise OSError('Error [read_flux_mca_out]: Do not support <mode=%s>.' % mode)
    d['N_photon'] = {'data': mca_obj.photons, 'name': 'Number of photons', 'units': 'N/A'}
    d['N_run'] = {'data': mca_obj.Nrun, 'name': 'Number of runs', 'units': 'N/A'}
    return d


def read_radiance_mca_out(mca_obj, abs_obj, mode='mean', squeeze=True):
    """Radiance (W/m^2/nm/sr): rad (+ rad_std), toa, N_photon, N_run (mca_out.py:412-505)."""
    mode = mode.lower()
    rad, = _per_run_arrays(mca_obj, abs_obj, 'radiance')
    _, toa = cal_factors(mca_obj.date, abs_obj, 1, mca_obj.Ng)
    rad, dims_info = _finish(rad, squeeze)
    d = {'toa': {'data': toa, 'name': 'TOA without SZA', 'units': 'W/m^2/nm'}}
    if mode == 'all':
        d['rad'] = {'data': rad, 'name': 'Radiance', 'units': 'W/m^2/nm/sr', 'dims_info': dims_info}
    elif mode == 'mean':
        d['rad'] = {'data': np.mean(rad, axis=-1), 'name': 'Radiance (mean)', 'units': 'W/m^2/nm/sr', 'dims_info': dims_info[:-1]}
        d['rad_std'] = {'data': np.std(rad, axis=-1), 'name': 'Radiance (standard deviation)', 'units': 'W/m^2/nm/sr', 'dims_info': dims_info[:-1]}
    else:
        raise OSError('Error [read_radiance_mca_out]: Do not support <mode=%s>.' % mode)
    d['N_photon'] = {'data': mca_obj.photons, 'name': 'Number of photons', 'units': 'N/A'}
    d['N_run'] = {'data': mca_obj.Nrun, 'name': 'Number of runs', 'units': 'N/A'}
    return d


def read_heating_mca_out(mca_obj, abs_obj, mode='mean', squeeze=True):
    """Absorbed power per layer (W/m^2/nm) -- the reader the reference lacks for target='heating rate'
    (mcarats.py:279-283 vs mca_out.py:202-206)."""
    mode = mode.lower()
    heat, = _per_run_arrays(mca_obj, abs_obj, 'heating')
    heat, dims_info = _finish(heat, squeeze)
    d = {}
    if mode == 'all':
        d['absorbed'] = {'data': heat, 'name': 'Absorbed flux per layer', 'units': 'W/m^2/nm', 'dims_info': dims_info}
    elif mode == 'mean':
        d['absorbed'] = {'data': np.mean(heat, axis=-1), 'name': 'Absorbed flux per layer (mean)', 'units': 'W/m^2/nm', 'dims_info': dims_info[:-1]}
        d['absorbed_std'] = {'data': np.std(heat, axis=-1), 'name': 'Absorbed flux per layer (standard deviation)', 'units': 'W/m^2/nm', 'dims_info': dims_info[:-1]}
    else:
        raise OSError('Error [read_heating_mca_out]: Do not support <mode=%s>.' % mode)
    return d


class mca_out_ng:

    """
    fname=    : HDF5 (or .npz when h5py is unavailable) file to write/read, default None
    mca_obj=  : mcarats_ng object
    abs_obj=  : absorption object
    mode=     : 'mean' or 'all'
    overwrite=: overwrite the file `fname`
    squeeze=  : drop axes of length 1

    self.data['f_up'|'f_down'|'f_down_direct'|'f_down_diffuse'|...'_std'|'rad'|'rad_std'|'toa'|'N_photon'|'N_run']
    Same loading rules as the reference (mca_out.py:160-177).
    """

    def __init__(self, fname=None, mca_obj=None, abs_obj=None, mode='mean', overwrite=False, squeeze=True, quiet=False, verbose=False):
        self.mode = mode
        self.quiet = quiet
        self.verbose = verbose
        self.overwrite = overwrite
        self.squeeze = squeeze
        self.fname = fname
        self.mca = mca_obj
        self.abs = abs_obj
        have = (mca_obj is not None) and (abs_obj is not None)
        if (fname is not None) and os.path.exists(fname) and (not overwrite):
            self.load()
        elif have and (fname is not None):
            self.run()
            self.dump()
        elif have and (fname is None):
            self.run()
        else:
            raise OSError('Error [mca_out_ng]: Please provide both <mca_obj> and <abs_obj> to proceed.')

    def run(self):
        if self.verbose:
            print('Message [mca_out_ng]: Reading <%s> ...' % self.mca.target.lower())
        if self.mca.target in ['flux', 'flux0']:
            self.data = read_flux_mca_out(self.mca, self.abs, mode=self.mode, squeeze=self.squeeze)
        elif self.mca.target == 'radiance':
            self.data = read_radiance_mca_out(self.mca, self.abs, mode=self.mode, squeeze=self.squeeze)
        elif self.mca.target == 'heating rate':
            self.data = read_flux_mca_out(self.mca, self.abs, mode=self.mode, squeeze=self.squeeze)
            self.data.update(read_heating_mca_out(self.mca, self.abs, mode=self.mode, squeeze=self.squeeze))

    # ---- persistence: HDF5 group '<mode>/<key>' with attrs name/units/dims_info (mca_out.py:209-233)
    def dump(self):
        if not self.quiet:
            print('Message [mca_out_ng]: Saving <%s> into <%s> ...' % (self.mca.target.lower(), self.fname))
        mode = self.mode.lower()
        try:
            import h5py
        except ImportError:
            h5py = None
        if h5py is not None:
            # call for call what the reference asks of h5py (mca_out.py:216-231; pinned by tests/golden/h5_dump_calls.json):
            # every ndarray -- 0-d included -- becomes a gzip-9 chunked dataset, scalars plain members, `dims_info` an
            # array of byte strings (np.string_ of the list), the other attributes as they are
            with h5py.File(self.fname, 'w') as f:
                g = f.create_group(mode)
                for key, item in self.data.items():
                    if isinstance(item['data'], np.ndarray):
                        g.create_dataset(key, data=item['data'], compression='gzip', compression_opts=9, chunks=True)
                    else:
                        g[key] = item['data']
                    for k0, v0 in item.items():
                        if k0 != 'data':
                            g[key].attrs[k0] = np.bytes_(v0) if k0 == 'dims_info' else v0
        else:
            flat = {}
            for key, item in self.data.items():
                flat['%s/%s' % (mode, key)] = np.asarray(item['data'])
                for k0, v0 in item.items():
                    if k0 != 'data':
                        flat['%s/%s@%s' % (mode, key, k0)] = np.asarray(str(v0))
            with open(self.fname, 'wb') as f:
                np.savez_compressed(f, **flat)

    def load(self):
        self.data = {}
        mode = self.mode.lower()
        with open(self.fname, 'rb') as f:
            magic = f.read(4)
        if magic[:2] == b'PK':
            z = np.load(self.fname, allow_pickle=False)
            for k in z.files:
                if not k.startswith(mode + '/'):
                    continue
                key = k[len(mode) + 1:]
                if '@' in key:
                    key, attr = key.split('@')
                    self.data.setdefault(key, {})[attr] = str(z[k])
                else:
                    arr = z[k]
                    self.data.setdefault(key, {})['data'] = arr if arr.ndim > 0 else arr[()]
        else:
            import h5py
            with h5py.File(self.fname, 'r') as f:
                g = f[mode]
                for key in g.keys():
                    item = {'data': g[key][...]}
                    for k0, v0 in g[key].attrs.items():
                        item[k0] = v0.decode() if isinstance(v0, bytes) else v0
                    self.data[key] = item
