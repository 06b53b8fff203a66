"""Tabulated phase-function set with the interface of er3t/rtm/mca/mca_sca.py."""

import os

import numpy as np

__all__ = ['mca_sca']


class mca_sca:

    """
    pha_obj: phase-function object with .data['ang'] (Nang) and .data['pha'] (Nang, Npf)

    self.nml: 'Sca_npf', 'Sca_nskip', 'Sca_nanci', 'Sca_nangi', 'Sca_inpfile'
    The binary (angles, then one phase function per column; float32 LE -- mca_sca.py:82-95) is written only when
    `fname` is given; the solver reads self.pha directly.
    """

    ID = 'MCARaTS Scattering'

    def __init__(self, pha_obj=None, fname=None, overwrite=True, force=False, verbose=False, quiet=False):
        self.overwrite = overwrite
        self.verbose = verbose
        self.quiet = quiet
        if pha_obj is None:
            raise OSError('Error [mca_sca]: Please provide an \'pha\' object for <pha_obj>.')
        self.pha = pha_obj
        self.pre_mca_sca()
        if fname is not None:
            if self.overwrite or ((not os.path.exists(fname)) and (not force)):
                self.gen_mca_sca_file(fname)
            else:
                self.nml['Sca_inpfile'] = {'data': fname}
        else:
            self.nml['Sca_inpfile'] = {'data': 'mca_sca.bin'}

    def pre_mca_sca(self, nskip=0, nanci=0):
        self.nml = {}
        self.nml['Sca_npf'] = dict(data=self.pha.data['pha']['data'].shape[1], name='Number of tabulated phase functions', units='N/A')
        self.nml['Sca_nskip'] = dict(data=nskip, name='Number of phase functions to be skipped', units='N/A')
        self.nml['Sca_nanci'] = dict(data=nanci, name='Number of ancillary data', units='N/A')
        self.nml['Sca_nangi'] = dict(data=self.pha.data['ang']['data'].size, name='Number of angles', units='N/A')

    def gen_mca_sca_file(self, fname):
        fname = os.path.abspath(fname)
        self.nml['Sca_inpfile'] = {'data': fname}
        with open(fname, 'wb') as f:
            np.asarray(self.pha.data['ang']['data']).astype('<f4').tofile(f)
            for i in range(self.nml['Sca_npf']['data']):
                np.asarray(self.pha.data['pha']['data'][:, i]).astype('<f4').tofile(f)
        if not self.quiet:
            print('Message [mca_sca]: File <%s> is created.' % fname)
