"""
Drop-in for er3t.rtm.mca (er3t/rtm/mca/__init__.py:1-8): same public names, same constructor arguments, same
`.nml` / `.data` payloads -- but `mcarats_ng` traces photons in-process on the B200 through include/b200rt.h instead of
writing namelists and spawning the MCARaTS binary.
"""

from .mcarats import mcarats_ng, cal_mca_azimuth, distribute_photon
from .mca_atm import mca_atm_1d, mca_atm_3d
from .mca_sca import mca_sca
from .mca_sfc import mca_sfc_2d
from .mca_out import mca_out_raw, mca_out_ng, write_mca_out_raw, cal_factors, read_flux_mca_out, read_radiance_mca_out
from .mca_inp import mca_inp_file, mca_inp_nml, load_mca_inp_nml
from .mca_run import mca_run, rearrange_jobs
from .util import func_ref_vs_cot, func_ref_vs_cot_multi_pixel

__all__ = ['mcarats_ng', 'cal_mca_azimuth', 'distribute_photon', 'mca_atm_1d', 'mca_atm_3d', 'mca_sca', 'mca_sfc_2d',
           'mca_out_raw', 'mca_out_ng', 'write_mca_out_raw', 'cal_factors', 'read_flux_mca_out', 'read_radiance_mca_out',
           'mca_inp_file', 'mca_inp_nml', 'load_mca_inp_nml', 'mca_run', 'rearrange_jobs',
           'func_ref_vs_cot', 'func_ref_vs_cot_multi_pixel']
