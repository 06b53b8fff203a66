"""2-D surface adapter with the interface of er3t/rtm/mca/mca_sfc.py."""

import copy
import os

import numpy as np

__all__ = ['mca_sfc_2d']


class mca_sfc_2d:

    """
    atm_obj: atmosphere object; sfc_obj: surface object with .data['nx'|'ny'|'sfc'] (e.g. er3t_b200.pre.sfc_2d_gen)

    self.nml: 'Sfc_nxb', 'Sfc_nyb', 'Sfc_tmps2d' (nx, ny), 'Sfc_jsfc2d' (nx, ny) int16, 'Sfc_psfc2d' (nx, ny, 5)
    Surface types (mca_sfc.py:89-133): 1 Lambertian (albedo clipped to [0, 1]), 4 LSRT (f_iso, f_geo, f_vol),
    2 DSM / Cox-Munk (diffuse_alb, diffuse_frac, refrac_r, refrac_i, slope variance).
    """

    ID = 'MCARaTS 2D Surface'

    def __init__(self, atm_obj=None, sfc_obj=None, fname=None, overwrite=True, force=False, verbose=False, quiet=False):
        self.overwrite = overwrite
        self.verbose = verbose
        self.quiet = quiet
        if atm_obj is None:
            raise OSError('\nError [mca_sfc_2d]: Please provide an <atm> object for <atm_obj>.')
        if sfc_obj is None:
            raise OSError('\nError [mca_sfc_2d]: Please provide an <sfc> object for <sfc_obj>.')
        self.atm = atm_obj
        self.sfc = sfc_obj
        self.pre_mca_2d_sfc()
        if fname is not None:
            if self.overwrite or ((not os.path.exists(fname)) and (not force)):
                self.gen_mca_2d_sfc_file(fname)
            else:
                self.nml['Sfc_inpfile'] = {'data': fname}
        else:
            self.nml['Sfc_inpfile'] = {'data': 'mca_sfc_2d.bin'}

    def pre_mca_2d_sfc(self):
        data = self.sfc.data['sfc']['data']
        name = self.sfc.data['sfc']['name'].lower()
        nx, ny = self.sfc.Nx, self.sfc.Ny
        self.nml = {}
        self.nml['Sfc_nxb'] = copy.deepcopy(self.sfc.data['nx'])
        self.nml['Sfc_nyb'] = copy.deepcopy(self.sfc.data['ny'])
        psfc = np.zeros((nx, ny, 5), dtype=np.float32)
        if ('lambertian' in name) and (np.squeeze(data).ndim == 2):
            jtype = 1
            psfc[:, :, 0] = np.clip(np.squeeze(data), 0.0, 1.0)
        elif ('brdf-lsrt' in name) or (data.shape[-1] == 3):
            jtype = 4
            psfc[:, :, :3] = data[:, :, :3]
        elif ('cox-munk' in name) or (data.shape[-1] == 5):
            jtype = 2
            psfc[:, :, :5] = data[:, :, :5]
        else:
            raise OSError('\nError [mca_sfc_2d]: Cannot determine surface type - currently only supports Lambertian surface and LSRT BRDF surface (e.g., MCD43A1).')
        self.nml['Sfc_tmps2d'] = dict(data=np.zeros((nx, ny), dtype=np.float32), name='Temperature anomalies', units='K')
        self.nml['Sfc_jsfc2d'] = dict(data=np.full((nx, ny), jtype, dtype=np.int16), name='Surface distribution type', units='N/A')
        self.nml['Sfc_psfc2d'] = dict(data=psfc, name='Surface distribution parameters', units='N/A')

    def gen_mca_2d_sfc_file(self, fname):
        # mca_sfc.py:136-146: tmps2d, jsfc2d (written as float32), psfc2d; Fortran order
        fname = os.path.abspath(fname)
        self.nml['Sfc_inpfile'] = {'data': fname}
        with open(fname, 'wb') as f:
            for key in ('Sfc_tmps2d', 'Sfc_jsfc2d', 'Sfc_psfc2d'):
                np.asarray(self.nml[key]['data']).astype('<f4').flatten(order='F').tofile(f)
        if not self.quiet:
            print('Message [mca_sfc_2d]: File <%s> is created.' % fname)
