"""
Drop-in for er3t.rtm.mca (er3t/rtm/mca/__init__.py:1-8): same public names, same constructor arguments, same
`.nml` / `.data` payloads -- but `mcarats_ng` traces photons in-process on the B200 through include/b200rt.h instead of
writing namelists and spawning the MCARaTS binary.
"""

from .mca_inp import *
from .mca_run import *
from .mca_atm import *
from .mca_sca import *
from .mca_sfc import *
from .mcarats import *
from .mca_out import *
from .util import *
