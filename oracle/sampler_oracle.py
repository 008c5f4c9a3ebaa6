"""ORACLE (test infrastructure, not product code) -- the two batch samplers.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this module.

1. ``legacy_batch`` (parity mode) is NumPy itself: ``RandomState(epoch).shuffle(arange(N))[:B]``,
   which is what ``sklearn.utils.resample(replace=False, random_state=epoch)`` does at the
   reference's call sites ``src/fm.py:72-79`` and ``src/mf.py:88-95``. The product's C++
   replica (MT19937 + legacy Fisher-Yates with masked rejection) is checked against it.

2. ``feistel_batch`` (perf mode) is NOT in the reference -- parity unpinned by construction;
   it is this build's own device-side sampler (reproducing the legacy
   shuffle is O(N) sequential host work per epoch). It draws B distinct rows as the first B
   images of a keyed bijection of [0, N): a 6-round balanced Feistel network on
   2*ceil(log2(N)/2) bits with cycle walking. This NumPy statement is the specification the
   CUDA kernel is tested against, bit for bit.
"""
from __future__ import annotations

import numpy as np

from .fm_oracle import legacy_batch  # noqa: F401  (re-exported)

_M32 = np.uint64(0xFFFFFFFF)
FEISTEL_ROUNDS = 6


def _mix32(x):
    """murmur3 fmix32 on uint64 arrays holding 32-bit values."""
    x = x & _M32
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x85EBCA6B)) & _M32
    x ^= x >> np.uint64(13)
    x = (x * np.uint64(0xC2B2AE35)) & _M32
    x ^= x >> np.uint64(16)
    return x


def feistel_round_keys(seed: int, epoch: int):
    base = (np.uint64(seed & 0xFFFFFFFF) * np.uint64(0x9E3779B9)
            + np.uint64(epoch & 0xFFFFFFFF) * np.uint64(0x7F4A7C15)) & _M32
    r = np.arange(FEISTEL_ROUNDS, dtype=np.uint64)
    return _mix32((base + (r + np.uint64(1)) * np.uint64(0x632BE5AB)) & _M32)


def feistel_half_bits(n_rows: int) -> int:
    bits = max(2, int(n_rows - 1).bit_length())
    return (bits + 1) // 2


def feistel_permute(q, n_rows: int, seed: int, epoch: int):
    """Image of positions q (array of ints in [0, N)) under the epoch's bijection of [0, N)."""
    if n_rows > 2**32:
        raise ValueError("feistel sampler supports at most 2**32 rows")
    h = np.uint64(feistel_half_bits(n_rows))
    mask = (np.uint64(1) << h) - np.uint64(1)
    keys = feistel_round_keys(seed, epoch)
    x = np.asarray(q, dtype=np.uint64).copy()
    todo = np.ones(x.shape, dtype=bool)
    while todo.any():
        cur = x[todo]
        L, R = cur >> h, cur & mask
        for r in range(FEISTEL_ROUNDS):
            f = _mix32(R ^ keys[r]) & mask
            L, R = R, L ^ f
        cur = (L << h) | R
        x[todo] = cur
        todo[todo] = cur >= np.uint64(n_rows)
    return x.astype(np.int64)


def feistel_batch(n_rows: int, batch_size: int, epoch: int, seed: int = 0) -> np.ndarray:
    if batch_size > n_rows:
        raise ValueError(
            "Cannot sample %d out of arrays with dim %d when replace is False" % (batch_size, n_rows))
    return feistel_permute(np.arange(batch_size), n_rows, seed, epoch)
