"""TEST INFRASTRUCTURE.  numpy Philox4x32-10 (Salmon, Moraes, Dror & Shaw, SC'11), the counter-based generator the
CUDA path uses for every photon.  Checked in tests/test_oracle_bits.py against the Random123 known-answer vectors."""

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr: (n, 4) uint32, key: (2,) or (n, 2) uint32 -> (n, 4) uint32"""
    c = np.array(ctr, dtype=np.uint64).reshape(-1, 4).copy()
    k = np.array(key, dtype=np.uint64).reshape(-1, 2).copy()
    if k.shape[0] == 1:
        k = np.repeat(k, c.shape[0], axis=0)
    for _ in range(10):
        p0 = M0 * c[:, 0]
        p1 = M1 * c[:, 2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        n0 = hi1 ^ c[:, 1] ^ k[:, 0]
        n2 = hi0 ^ c[:, 3] ^ k[:, 1]
        c = np.stack([n0, lo1, n2, lo0], axis=1)
        k[:, 0] = (k[:, 0] + np.uint64(W0)) & MASK
        k[:, 1] = (k[:, 1] + np.uint64(W1)) & MASK
    return c.astype(np.uint32)


def photon_block(seed, first, n, c2=0, c3=0):
    """The layout b200rt_philox_fill / oracle_philox_fill use: key = seed (lo, hi), ctr = (index lo, index hi, c2, c3)."""
    idx = np.uint64(first) + np.arange(n, dtype=np.uint64)
    ctr = np.stack([idx & MASK, idx >> np.uint64(32), np.full(n, c2, dtype=np.uint64), np.full(n, c3, dtype=np.uint64)], axis=1)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint64)
    return philox4x32_10(ctr, key)
