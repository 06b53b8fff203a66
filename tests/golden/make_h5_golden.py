#!/usr/bin/env python
"""
Record what the REFERENCE's own `mca_out_ng.dump` (er3t/rtm/mca/mca_out.py:209-233) asks of h5py, with the recording
stand-in tests/fake_h5py.py in h5py's place: tests/golden/h5_dump_calls.json.  Build container only.

    python tests/golden/make_h5_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, '..'))
sys.path.insert(0, HERE)



def sample_data():
    """the dictionary layout read_radiance_mca_out / read_flux_mca_out produce (mca_out.py:399-407,500-505)"""
    rng = np.random.default_rng(7)
    return {
        'rad': {'data': rng.random((4, 3)).astype(np.float32), 'name': 'Radiance', 'units': 'W/m^2/nm/sr', 'dims_info': ['x', 'y']},
        'rad_std': {'data': rng.random((4, 3)).astype(np.float32), 'name': 'Radiance standard deviation', 'units': 'W/m^2/nm/sr', 'dims_info': ['x', 'y']},
        'toa': {'data': 1.6123, 'name': 'TOA without SZA', 'units': 'W/m^2/nm'},
        'N_photon': {'data': np.array([100, 200, 300]), 'name': 'Number of photons', 'units': 'N/A'},
        'N_run': {'data': 3, 'name': 'Number of runs', 'units': 'N/A'},
    }


def main():
    import fake_h5py
    sys.modules['h5py'] = fake_h5py
    if not hasattr(np, 'string_'):
        np.string_ = np.bytes_                    # NumPy 2 dropped the alias the reference uses (mca_out.py:229)
    import make_golden
    make_golden.import_reference()
    from er3t.rtm.mca.mca_out import mca_out_ng

    class _M:
        target = 'radiance'

    o = object.__new__(mca_out_ng)
    o.data, o.fname, o.mode, o.quiet, o.verbose, o.mca = sample_data(), 'golden.h5', 'mean', True, False, _M()
    fake_h5py.reset()
    o.dump()
    with open(os.path.join(HERE, 'h5_dump_calls.json'), 'w') as f:
        json.dump(fake_h5py.LOG, f, indent=1)
    print('recorded %d calls' % len(fake_h5py.LOG))


if __name__ == '__main__':
    main()
