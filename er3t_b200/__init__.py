"""
er3t_b200 -- B200-native, in-process replacement for the photon-transport path that EaR3T (hong-chen/er3t)
delegates to the external MCARaTS binary through er3t.rtm.mca.

    er3t_b200.abi      ctypes mirror of include/b200rt.h + loader of csrc/libb200rt.so (no CPU fallback)
    er3t_b200.solver   one handle per GPU
    er3t_b200.rtm.mca  mcarats_ng / mca_atm_1d / mca_atm_3d / mca_sca / mca_sfc_2d / mca_out_ng with the reference's API
    er3t_b200.pre      file-free input builders with the payload layout of er3t.pre.*
"""

__version__ = '0.1.0'
