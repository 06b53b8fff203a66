"""
Namelist schema and text writer with the interface of er3t/rtm/mca/mca_inp.py.

The in-process solver does not read namelist files; the schema is kept because (a) it is the parameter contract
`mcarats_ng.nml[ig]` follows and (b) `mca_inp_file` stays available as a debug dump that
is byte-compatible with what the reference writes for MCARaTS (14 groups, `%-.16g` scalars, six `%12g` values per
line for arrays -- er3t/rtm/mca/mca_inp.py:636-697).
"""

import os
from collections import OrderedDict

import numpy as np

from er3t_b200.util import nice_array_str

__all__ = ['mca_inp_file', 'mca_inp_nml', 'load_mca_inp_nml']

# group -> ordered variable names (er3t/rtm/mca/mca_inp.py:36-382)
_SCHEMA = OrderedDict([
    ('mcarWld_nml_init', ['Wld_mverb', 'Wld_jseed', 'Wld_mbswap', 'Wld_mtarget', 'Wld_moptim', 'Wld_njob']),
    ('mcarSca_nml_init', ['Sca_inpfile', 'Sca_npf', 'Sca_nanci', 'Sca_nangi', 'Sca_nskip', 'Sca_ndfl', 'Sca_nchi', 'Sca_ntg', 'Sca_qtfmax']),
    ('mcarAtm_nml_init', ['Atm_inpfile', 'Atm_np1d', 'Atm_np3d', 'Atm_nx', 'Atm_ny', 'Atm_nz', 'Atm_iz3l', 'Atm_nz3', 'Atm_nkd',
                          'Atm_mtprof', 'Atm_nwl', 'Atm_nqlay', 'Atm_iipfd1d', 'Atm_iipfd3d']),
    ('mcarSfc_nml_init', ['Sfc_inpfile', 'Sfc_mbrdf', 'Sfc_nxb', 'Sfc_nyb', 'Sfc_nsco', 'Sfc_nsuz']),
    ('mcarSrc_nml_init', ['Src_nsrc']),
    ('mcarFlx_nml_init', ['Flx_mflx', 'Flx_mhrt', 'Flx_nxf', 'Flx_nyf', 'Flx_diff0', 'Flx_diff1', 'Flx_cf_dtau']),
    ('mcarRad_nml_init', ['Rad_mrkind', 'Rad_mpmap', 'Rad_mplen', 'Rad_nrad', 'Rad_nxr', 'Rad_nyr', 'Rad_nwf', 'Rad_ntp', 'Rad_tpmin', 'Rad_tpmax']),
    ('mcarVis_nml_init', ['Vis_mrend', 'Vis_epserr', 'Vis_fpsmth', 'Vis_fatten', 'Vis_nqhem']),
    ('mcarPho_nml_init', ['Pho_iso_SS', 'Pho_iso_tru', 'Pho_iso_max', 'Pho_wmin', 'Pho_wmax', 'Pho_wfac', 'Pho_pfpeak']),
    ('mcarWld_nml_job', ['Wld_nplcf']),
    ('mcarAtm_nml_job', ['Atm_idread', 'Atm_wkd0', 'Atm_dx', 'Atm_dy', 'Atm_zgrd0', 'Atm_tmp1d', 'Atm_ext1d', 'Atm_omg1d', 'Atm_apf1d',
                         'Atm_abs1d', 'Atm_fext1d', 'Atm_fext3d', 'Atm_fabs1d', 'Atm_fabs3d', 'Atm_mcs_rat', 'Atm_mcs_frc',
                         'Atm_mcs_dtauz', 'Atm_mcs_dtauxy']),
    ('mcarSfc_nml_job', ['Sfc_idread', 'Sfc_mtype', 'Sfc_param', 'Sfc_nudsm', 'Sfc_nurpv', 'Sfc_nulsrt', 'Sfc_nqpot', 'Sfc_rrmax', 'Sfc_rrexp']),
    ('mcarSrc_nml_job', ['Src_mtype', 'Src_dwlen', 'Src_mphi', 'Src_flx', 'Src_qmax', 'Src_the', 'Src_phi']),
    ('mcarRad_nml_job', ['Rad_mrproj', 'Rad_difr0', 'Rad_difr1', 'Rad_zetamin', 'Rad_npwrn', 'Rad_npwrf', 'Rad_cf_dmax', 'Rad_cf_taus',
                         'Rad_wfunc0', 'Rad_rmin0', 'Rad_rmid0', 'Rad_rmax0', 'Rad_phi', 'Rad_the', 'Rad_psi', 'Rad_umax', 'Rad_vmax',
                         'Rad_qmax', 'Rad_xpos', 'Rad_ypos', 'Rad_zloc', 'Rad_apsize', 'Rad_zref']),
])

# MCARaTS defaults of the knobs the transport physics depends on (documented at mca_inp.py:19-364); er3t never sets
# them, so they are part of the effective configuration
DEFAULTS = {'Pho_iso_SS': 1, 'Pho_iso_max': 1000000, 'Pho_wmin': 0.2, 'Pho_wmax': 3.0, 'Pho_wfac': 1.0,
            'Rad_zref': 0.0, 'Src_qmax': 0.0, 'Wld_moptim': 2, 'Atm_iz3l': 1, 'Atm_nz3': 0}


def load_mca_inp_nml():
    """Fresh schema: OrderedDict(group -> OrderedDict(variable -> None))."""
    return OrderedDict((g, OrderedDict((k, None) for k in keys)) for g, keys in _SCHEMA.items())


def mca_inp_nml(input_dict, verbose=True, comment=False):
    """
    Distribute `input_dict` over the namelist groups.  Array-slice keys such as 'Atm_ext1d(1:, 2)' are inserted after
    the last existing slice of the same base variable (or after the base variable); unknown keys raise OSError
    (er3t/rtm/mca/mca_inp.py:575-632).
    """
    nml = load_mca_inp_nml()
    order = []                       # [(key, group)] in output order
    for g, keys in _SCHEMA.items():
        order += [(k, g) for k in keys]
    names = [k for k, _ in order]
    for key, val in input_dict.items():
        if key in names:
            nml[order[names.index(key)][1]][key] = val
        elif '(' in key and ')' in key:
            base = key[:key.index('(')]
            if base not in names:
                raise OSError('Error [mca_inp_nml]: please check input variable <%s>.' % key)
            more = [k for k in names if base in k and '(' in k]
            idx = names.index(more[-1]) if more else names.index(base)
            grp = order[idx][1]
            order.insert(idx + 1, (key, grp))
            names.insert(idx + 1, key)
            nml[grp][key] = val
        else:
            raise OSError('Error [mca_inp_nml]: please check input variable <%s>.' % key)
    return nml, OrderedDict(order)


def mca_inp_file(input_fname, input_dict, verbose=True, comment=False):
    """Write the MCARaTS namelist text file for one job (debug dump; the CUDA solver does not read it)."""
    nml, where = mca_inp_nml(input_dict, verbose=verbose, comment=comment)
    input_fname = os.path.abspath(input_fname)
    os.makedirs(os.path.dirname(input_fname), exist_ok=True)
    with open(input_fname, 'w') as f:
        for grp in nml.keys():
            f.write('&%s\n' % grp)
            for key in [k for k, g in where.items() if g == grp]:
                var = nml[grp][key]
                if var is None:
                    continue
                if isinstance(var, (int, float, np.int32, np.int64, np.float32, np.float64)):
                    f.write(' %-15s = %-.16g\n' % (key, var))
                elif isinstance(var, str):
                    f.write((' %-15s = %s\n' if '*' in var else ' %-15s = \'%s\'\n') % (key, var))
                elif isinstance(var, np.ndarray):
                    if var.size > 1:
                        s = nice_array_str(var)
                        if len(s) <= 80:
                            f.write(' %-15s = %s\n' % (key, s))
                        else:
                            f.write(' %-15s =\n' % key)
                            f.write('%s\n' % s)
                    elif var.size == 1:
                        f.write(' %-15s = %-g\n' % (key, var.reshape(-1)[0]))
                else:
                    raise ValueError('Error [mca_inp_file]: only types of int, float, str, ndarray are supported (do not support <%s> as %s).' % (key, type(var)))
            f.write('/\n')
