"""
File-free input builders with the payload layout of er3t.pre.* (SURVEY.md 8f rank 2).

The reference's builders read data files that are not shipped with the repository (er3t/data/...);
these stand-ins produce objects with the same attributes and dict-of-dict payloads (`.lev/.lay`, `.coef`, `.data`)
from analytic or seeded synthetic inputs, so that `er3t_b200.rtm.mca` -- and the unmodified reference classes, which
are duck-typed on those payloads -- can be driven without any download.
"""

from .atm import atm_atmmod
from .abs import abs_16g, abs_gen
from .pha import pha_hg, pha_mie_wc, cal_hg_pha_func
from .cld import cld_gen_hom, cld_gen_hem, cld_gen_les
from .sfc import sfc_2d_gen, cal_ocean_brdf
