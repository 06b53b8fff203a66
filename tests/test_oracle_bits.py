"""Integer parts of the path: Philox4x32-10 against the published Random123 known-answer vectors, and the oracle's
C++ implementation against the numpy one (bit-exact)."""

import numpy as np

import oracle
from oracle import philox_np


def test_philox_known_answer_vectors():
    # Random123 kat_vectors: philox4x32-10
    kat = [
        ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, exp in kat:
        got = philox_np.philox4x32_10(np.array([ctr], dtype=np.uint32), np.array(key, dtype=np.uint32))[0]
        assert tuple(int(v) for v in got) == exp


def test_oracle_philox_matches_numpy_bit_exact():
    for seed, first, c2, c3 in [(0, 0, 0, 0), (20260101, 5, 3, 0xB200), (0xDEADBEEFCAFEF00D, (1 << 32) - 7, 17, 0x0ACC1E)]:
        a = oracle.philox(seed, first, 64, c2=c2, c3=c3)
        b = philox_np.photon_block(seed, first, 64, c2=c2, c3=c3)
        assert np.array_equal(a, b)


def test_streams_are_disjoint_per_photon_and_job():
    a = philox_np.photon_block(1, 0, 1000)
    b = philox_np.photon_block(2, 0, 1000)
    c = philox_np.photon_block(1, 0, 1000, c2=1)
    assert len(set(map(tuple, np.concatenate([a, b, c]).tolist()))) == 3000
    u = (a.astype(np.float64) + 0.5) / 4294967296.0
    assert abs(u.mean() - 0.5) < 0.02 and abs(u.var() - 1.0 / 12.0) < 0.01
