"""Seeded synthetic inputs shared by the golden generator and the tests.

TEST INFRASTRUCTURE ONLY.  numpy PCG64 streams; each golden file stores a CRC32 of
the pool it was generated from, so a drift of the generator is detected.
"""

from __future__ import annotations

import numpy as np


def make_distinct(row: np.ndarray) -> int:
    """Nudge exact duplicates apart by one ulp (in place); returns how many were moved.

    The reference's ``torch.sort`` (code_base/arithmetic.py:127) orders equal logits in an
    implementation-defined way, so golden inputs are made tie-free; the tie-break rule of
    this framework (lower id first) is tested against the oracle separately.
    """
    order = np.argsort(row, kind="stable")
    vals = row[order]
    dup = np.nonzero(vals[1:] <= vals[:-1])[0]
    moved = 0
    while dup.size:
        for i in dup:
            if vals[i + 1] <= vals[i]:
                vals[i + 1] = np.nextafter(vals[i], np.float32(np.inf))
                moved += 1
        dup = np.nonzero(vals[1:] <= vals[:-1])[0]
    row[order] = vals
    return moved


def logits_pool(seed: int, T: int, V: int, scale: float) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64(seed))
    pool = rng.standard_normal((T, V), dtype=np.float32) * np.float32(scale)
    for t in range(T):
        make_distinct(pool[t])
    return pool


def message_bits(seed: int, n: int):
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.integers(0, 2, n).astype(np.uint8)


def rows_for(pool: np.ndarray, stream: int):
    T = pool.shape[0]
    return lambda t: pool[(t + 3 * stream) % T]
