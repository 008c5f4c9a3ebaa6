"""Device-side semi-synthetic interaction logs.

The reference simulates clicks from KuaiRec's watch ratios (``utils/dataloader/kuairec/_click.py:148-235``):
relevance ``gamma = clip(watch_ratio / 2, 0, 1)``, exposure ``theta_i = max(sigmoid(3 z_i - 1) ** bias, 0.1)``,
``R ~ Be(gamma)``, ``O ~ Be(theta)``, click ``Y = O * R``, propensity ``theta ** pow_used``. ``ClickModel`` holds the
same model over synthetic popularity / activity distributions and ``generate_rows`` draws any shard
``[row0, row0 + n_rows)`` of the log ON THE DEVICE (counter-based Philox: every GPU can produce every row), as
factored rows ready for ``FactorizationMachines.fit`` -- a 10^9-row log costs 24 bytes per row of HBM and no PCIe.
Specification: ``oracle/clicks_oracle.py``.
"""
from __future__ import annotations

import ctypes
from ctypes import byref, c_double, c_int32, c_int64, c_uint64, c_void_p
from dataclasses import dataclass

import numpy as np

from . import _capi
from ._capi import check, lib, ptr
from .factored import BLOCK_CTX, BLOCK_ID, BLOCK_TABLE, KEY_ITEM, KEY_USER, RowsBlock


class _ClickModelC(ctypes.Structure):
    """``rfm_click_model`` of include/rfm_b200.h."""
    _fields_ = [("seed", c_uint64), ("row0", c_int64), ("n_users", c_int64), ("n_items", c_int64),
                ("user_cdf", c_void_p), ("item_cdf", c_void_p), ("item_exposure", c_void_p), ("item_pscore", c_void_p), ("pow_used", c_double),
                ("n_hidden", c_int32), ("keep_labels", c_int32), ("hidden_scale", c_double), ("noise_scale", c_double),
                ("watch_shift", c_double), ("relevance_clip", c_double)]


def sigmoid_exposure(x, a: float = 3.0, b: float = -1.0):
    """``_sigmoid`` of ``kuairec/_click.py:238-240``."""
    return 1.0 / (1.0 + np.exp(-(a * x + b)))


@dataclass
class ClickModel:
    """Popularity / activity distributions and the reference's exposure model over them."""
    n_users: int
    n_items: int
    seed: int = 12345
    exposure_bias: float = 3.0      # conf/setting/kuairec.yaml:13
    pow_used: float = 0.5           # conf/setting/kuairec.yaml:47
    eps: float = 0.1                # kuairec/_click.py:176
    n_hidden: int = 8
    hidden_scale: float = 0.5
    noise_scale: float = 0.3
    watch_shift: float = -0.35
    relevance_clip: float = 2.0     # kuairec/_click.py:151

    def __post_init__(self):
        rng = np.random.default_rng(self.seed)
        rank = rng.permutation(self.n_items)
        pop = (rank + 10.0) ** -0.8
        self.item_prob = pop / pop.sum()
        act = rng.lognormal(sigma=0.6, size=self.n_users)
        self.user_prob = act / act.sum()
        self.user_cdf = np.cumsum(self.user_prob)
        self.user_cdf[-1] = 1.0
        self.item_cdf = np.cumsum(self.item_prob)
        self.item_cdf[-1] = 1.0
        # exposure from the items' (expected) exposure counts, kuairec/_click.py:193-202
        counts = self.item_prob * 1e6
        z = (counts - counts.mean()) / counts.std(ddof=1)
        self.item_exposure = np.maximum(sigmoid_exposure(z) ** self.exposure_bias, self.eps)
        self.item_pscore = self.item_exposure ** self.pow_used          # kuairec/loader.py:167

    def c_struct(self, row0: int, keep_labels: bool):
        return _ClickModelC(self.seed, row0, self.n_users, self.n_items, self.user_cdf.ctypes.data,
                            self.item_cdf.ctypes.data, self.item_exposure.ctypes.data, self.item_pscore.ctypes.data, self.pow_used, self.n_hidden,
                            int(keep_labels), self.hidden_scale, self.noise_scale, self.watch_shift, self.relevance_clip)


class GeneratedRows(_capi._Handle):
    """Factored rows generated on the device; stands where ``train["features"]`` does (labels / pscores: None)."""

    _destroy = "rfm_csr_destroy"
    factored = True

    def __init__(self, ctx, model: ClickModel, n_rows: int, blocks, row0: int = 0, dtype="float64", keep_labels=False):
        super().__init__()
        arr = (RowsBlock * len(blocks))()
        keep, n_cols = [], 0
        for slot, b in zip(arr, blocks):
            if b[0] == "id":
                slot.kind, slot.key = BLOCK_ID, KEY_USER if b[1] == "user" else KEY_ITEM
                slot.n_cols = slot.n_entities = int(b[2])
            elif b[0] == "table":
                t = b[2].tocsr()
                indptr = np.ascontiguousarray(t.indptr)
                is64 = indptr.dtype == np.int64
                if not is64:
                    indptr = _capi.as_array(indptr, np.int32)
                indices, data = _capi.as_array(t.indices, np.int32), _capi.as_array(t.data, np.float64)
                keep += [indptr, indices, data]
                slot.kind, slot.key = BLOCK_TABLE, KEY_USER if b[1] == "user" else KEY_ITEM
                slot.n_cols, slot.n_entities = t.shape[1], t.shape[0]
                slot.indptr, slot.indptr_is_int64 = indptr.ctypes.data, int(is64)
                slot.indices, slot.data = indices.ctypes.data, data.ctypes.data
            elif b[0] == "ctx":                      # ("ctx", n_columns): values are generated, ~ N(0, 1)
                slot.kind, slot.n_cols = BLOCK_CTX, int(b[1])
            else:
                raise ValueError("unknown block kind %r" % (b[0],))
            n_cols += int(slot.n_cols)
        cm = model.c_struct(row0, keep_labels)
        check(lib().rfm_factored_generate(ctx.handle, n_rows, byref(cm), arr, len(blocks), _capi.dtype_code(dtype),
                                          byref(self.handle)))
        self.ctx, self.dtype, self.n_rows, self.shape, self.h2d_bytes = ctx, dtype, n_rows, (n_rows, n_cols), 0
        self.n_ctx = sum(int(b[1]) for b in blocks if b[0] == "ctx")
        self.keep_labels = keep_labels

    def download(self, first: int = 0, n: int = None):
        n = self.n_rows - first if n is None else n
        users, items = np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int32)
        ctxv, targets = np.empty((n, self.n_ctx)), np.empty(n)
        labels = np.empty(n, dtype=np.int8) if self.keep_labels else None
        rel = np.empty(n, dtype=np.int8) if self.keep_labels else None
        check(lib().rfm_rows_download(self.handle, first, n, ptr(users), ptr(items), ptr(ctxv), ptr(targets), ptr(labels),
                                      ptr(rel)))
        return dict(users=users, items=items, ctx=ctxv, targets=targets, labels=labels, relevance=rel)
