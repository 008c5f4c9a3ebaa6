"""ORACLE (test infrastructure, not product code) -- NumPy specification of the device-side click generator
(``csrc/synth.cu``, ``rfm_factored_generate``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu legs may import this module.

Parity status: the simulation MODEL is the reference's (``utils/dataloader/kuairec/_click.py``):

* ``relevance``  -- ``_click.py:148-171``  gamma = clip(watch_ratio / relevance_clip, 0, 1)
* ``exposure``   -- ``_click.py:173-205``  theta_i = max(sigmoid(3 z_i - 1) ** exposure_bias, eps), z = standardised counts
* ``clicks``     -- ``_click.py:207-235``  O ~ Be(theta), R ~ Be(gamma), Y = O * R
* ``pscore``     -- ``kuairec/loader.py:167``  theta ** pow_used

``exposure`` is checked against the reference's own ``_sigmoid`` / formula on the same counts (tests). The random
STREAM cannot be the reference's: it draws from NumPy's legacy global generator, sequentially, from real KuaiRec
columns that are not shipped. This generator is counter-based (Philox4x32-10, Salmon et al. 2011: key = seed, counter =
(index, stream, lane)), which is what lets any GPU produce any shard; this file is its specification and the device
kernel is held to it: ids, labels bit for bit (up to values closer than 1e-12 to a decision boundary), reals to 1e-12.
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)
ST_PAIR, ST_CTX, ST_NOISE, ST_CLICK, ST_USER_FACTOR, ST_ITEM_FACTOR = 0, 1, 2, 3, 16, 17


def philox4x32(seed: int, index, stream: int, lane: int):
    """(n, 4) uint32 outputs of Philox4x32-10 for counters (index low, index high, stream, lane), key = seed."""
    index = np.asarray(index, dtype=np.uint64)
    c = [index & MASK, index >> np.uint64(32), np.full(index.shape, stream, dtype=np.uint64),
         np.full(index.shape, lane, dtype=np.uint64)]
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return np.stack(c, axis=-1).astype(np.uint32)


def u01(a, b):
    """53-bit uniform in (0, 1) from two 32-bit words (NumPy's double construction, shifted off zero)."""
    return ((a >> np.uint32(5)).astype(np.float64) * 67108864.0 + (b >> np.uint32(6)).astype(np.float64) + 0.5) \
        * (1.0 / 9007199254740992.0)


def normal_from(r):
    u1, u2 = u01(r[..., 0], r[..., 1]), u01(r[..., 2], r[..., 3])
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(np.pi * (2.0 * u2))


def cdf_search(cdf, u):
    """Smallest index whose inclusive cumulative probability exceeds u (the last index absorbs rounding)."""
    return np.minimum(np.searchsorted(cdf, u, side="right"), len(cdf) - 1)


def sigmoid_exposure(x, a=3.0, b=-1.0):
    return 1.0 / (1.0 + np.exp(-(a * x + b)))          # _click.py:238-240


def exposure(counts, exposure_bias: float, eps: float = 0.1):
    """theta per item from exposure counts, ``_click.py:193-202`` (pandas ``std`` is the sample standard deviation)."""
    counts = np.asarray(counts, dtype=np.float64)
    z = (counts - counts.mean()) / counts.std(ddof=1)
    return np.maximum(sigmoid_exposure(z) ** exposure_bias, eps)


def relevance(watch_ratio, relevance_clip: float = 2.0):
    return np.clip(watch_ratio / relevance_clip, 0.0, 1.0)     # _click.py:168-171


def generate(seed, row0, n_rows, user_cdf, item_cdf, theta, pow_used, n_ctx=1, n_hidden=8, hidden_scale=0.5,
             noise_scale=0.3, watch_shift=-0.35, relevance_clip=2.0):
    """Rows [row0, row0 + n_rows) of the log: dict of users, items, ctx (n_rows, n_ctx), labels Y, relevance R,
    pscores, targets Y / pscore."""
    g = np.arange(row0, row0 + n_rows, dtype=np.uint64)
    r = philox4x32(seed, g, ST_PAIR, 0)
    users = cdf_search(user_cdf, u01(r[:, 0], r[:, 1]))
    items = cdf_search(item_cdf, u01(r[:, 2], r[:, 3]))
    ctx = np.stack([normal_from(philox4x32(seed, g, ST_CTX, j)) for j in range(n_ctx)], axis=1) if n_ctx else \
        np.zeros((n_rows, 0))
    dot = np.zeros(n_rows)
    for f in range(n_hidden):
        dot = dot + normal_from(philox4x32(seed, users, ST_USER_FACTOR, f)) * \
            normal_from(philox4x32(seed, items, ST_ITEM_FACTOR, f))
    eps = normal_from(philox4x32(seed, g, ST_NOISE, 0))
    watch_ratio = np.exp(hidden_scale * dot + noise_scale * eps + watch_shift)
    gamma = relevance(watch_ratio, relevance_clip)
    th = np.asarray(theta)[items]
    r = philox4x32(seed, g, ST_CLICK, 0)
    uo, ur = u01(r[:, 0], r[:, 1]), u01(r[:, 2], r[:, 3])
    O = (uo < th).astype(np.int8)
    R = (ur < gamma).astype(np.int8)
    Y = O * R                                                   # _click.py:231
    ps = th ** pow_used                                         # kuairec/loader.py:167
    return dict(users=users.astype(np.int32), items=items.astype(np.int32), ctx=ctx, labels=Y, relevance=R,
                pscores=ps, targets=Y / ps, gamma=gamma, theta=th, u_exposure=uo, u_relevance=ur)
