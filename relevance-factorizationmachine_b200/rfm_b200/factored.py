"""``FactoredFeatures`` -- FM rows in the form the reference's data layer holds them BEFORE ``scipy.sparse.hstack``
(SURVEY.md section 8 row f3).

The reference builds its FM design matrix by stacking per-entity tables gathered per interaction:

* Coat (``utils/dataloader/coat/_preparer.py:154-170``)::

      sp.hstack([onehot_user_ids[user_ids], user_features[user_ids],
                 onehot_item_ids[item_ids], item_features[item_ids]])

* KuaiRec (``utils/dataloader/kuairec/_feature.py:169-209``)::

      hstack([I_user, I_item, csr_matrix(interaction columns), user table rows, video table rows])

A ``FactoredFeatures`` keeps those pieces apart -- the blocks, in column order, and one ``(user, item)`` pair per
interaction -- and can stand wherever the reference passes the stacked matrix: ``train["features"]``,
``val["features"]``, ``predict(X=...)``, ``evaluator.features[...]``. The device assembles each row on the fly in the
stacked matrix's column order, so results are bit-identical to the hstacked CSR while the upload shrinks from
``12 m + 16`` bytes per interaction (m stored non-zeros) to ``~8 + 8 n_ctx + 16``. ``tocsr()`` materialises the
stacked matrix (what the reference would have built) for parity checks.

Binding a maintainer would add (``coat/_preparer.py:168``)::

    fm_features[_df_name] = FactoredFeatures([("id", "user", n_users), ("table", "user", user_features),
                                              ("id", "item", n_items), ("table", "item", item_features)],
                                             users=user_ids, items=item_ids)
"""
from __future__ import annotations

import ctypes
from ctypes import POINTER, byref, c_double, c_int32, c_int64, c_void_p
from typing import List, Sequence, Tuple

import numpy as np

from . import _capi
from ._capi import check, lib, ptr

BLOCK_ID, BLOCK_TABLE, BLOCK_CTX = 0, 1, 2
KEY_USER, KEY_ITEM = 0, 1


class RowsBlock(ctypes.Structure):
    """``rfm_rows_block`` of include/rfm_b200.h."""
    _fields_ = [("kind", c_int32), ("key", c_int32), ("n_cols", c_int64), ("n_entities", c_int64),
                ("indptr", c_void_p), ("indptr_is_int64", c_int32), ("reserved", c_int32),
                ("indices", c_void_p), ("data", c_void_p), ("values", c_void_p)]


class FactoredFeatures:
    """blocks: sequence of ``("id", "user"|"item", n_ids)``, ``("table", "user"|"item", csr_table)`` and
    ``("ctx", dense)`` (``dense``: ``(n_rows,)`` or ``(n_rows, c)`` float array), in the column order of the stacked
    matrix. users / items: one id per interaction (int32 or int64)."""

    def __init__(self, blocks: Sequence[tuple], users, items):
        self.users = np.ascontiguousarray(users)
        self.items = np.ascontiguousarray(items)
        if self.users.ndim != 1 or self.users.shape != self.items.shape:
            raise ValueError("users and items must be 1-D arrays of equal length")
        for name, arr in (("users", self.users), ("items", self.items)):
            if arr.dtype not in (np.int32, np.int64):
                if arr.dtype.kind not in "iu":
                    raise ValueError("%s must be integer ids" % name)
                setattr(self, name, arr.astype(np.int64))
        n_rows = self.users.shape[0]
        self.blocks: List[tuple] = []
        n_cols = 0
        for blk in blocks:
            kind = blk[0]
            if kind == "id":
                _, key, n_ids = blk
                self._check_key(key)
                self.blocks.append(("id", key, int(n_ids)))
                n_cols += int(n_ids)
            elif kind == "table":
                _, key, table = blk
                self._check_key(key)
                table = table.tocsr()
                if not table.has_canonical_format:
                    table = table.copy()
                    table.sum_duplicates()
                self.blocks.append(("table", key, table))
                n_cols += table.shape[1]
            elif kind == "ctx":
                dense = np.asarray(blk[1], dtype=np.float64)
                if dense.ndim == 1:
                    dense = dense[:, None]
                if dense.ndim != 2 or dense.shape[0] != n_rows:
                    raise ValueError("a context block needs one row per interaction")
                self.blocks.append(("ctx", np.ascontiguousarray(dense)))
                n_cols += dense.shape[1]
            else:
                raise ValueError("unknown block kind %r" % (kind,))
        if not 1 <= len(self.blocks) <= 6:
            raise ValueError("between 1 and 6 blocks")
        if sum(b[1].shape[1] for b in self.blocks if b[0] == "ctx") > 32:
            raise ValueError("at most 32 context columns in total")
        self.shape = (n_rows, n_cols)

    @staticmethod
    def _check_key(key):
        if key not in ("user", "item"):
            raise ValueError("block key must be 'user' or 'item'")

    # ---- what callers of the stacked matrix use ------------------------------------------------------------
    def __len__(self):
        return self.shape[0]

    def __getitem__(self, rows):
        """Row selection like ``X[rows]`` on the stacked matrix (slices, index arrays, boolean masks)."""
        if isinstance(rows, tuple):
            raise IndexError("FactoredFeatures supports row selection only")
        blocks = [("ctx", b[1][rows]) if b[0] == "ctx" else b for b in self.blocks]
        return FactoredFeatures(blocks, self.users[rows], self.items[rows])

    def tocsr(self):
        """The matrix the reference would have built: ``scipy.sparse.hstack`` of the gathered blocks."""
        import scipy.sparse as sp
        parts = []
        n = self.shape[0]
        for b in self.blocks:
            if b[0] == "id":
                ids = self.users if b[1] == "user" else self.items
                parts.append(sp.csr_matrix((np.ones(n), (np.arange(n), ids)), shape=(n, b[2])))
            elif b[0] == "table":
                ids = self.users if b[1] == "user" else self.items
                parts.append(b[2][ids])
            else:
                parts.append(sp.csr_matrix(b[1]))
        X = sp.hstack(parts, format="csr")
        X.sort_indices()
        return X

    @property
    def nbytes(self):
        """Host bytes a device copy moves (ids + context values + tables)."""
        total = self.users.nbytes + self.items.nbytes
        for b in self.blocks:
            if b[0] == "table":
                total += b[2].indptr.nbytes + b[2].indices.nbytes + b[2].data.nbytes
            elif b[0] == "ctx":
                total += b[1].nbytes
        return total


class PerItem:
    """``pscores`` as a per-ITEM table: stands where ``train["pscores"]`` does when the features are factored. The
    reference gathers an item-level propensity per row (``coat/_preparer.py:56-62``, ``kuairec/loader.py:160-168``);
    handing the table over instead saves 8 bytes per interaction of upload. ``pscore[t] = table[item[t]]``."""

    def __init__(self, table):
        self.table = np.ascontiguousarray(table, dtype=np.float64)
        if self.table.ndim != 1:
            raise ValueError("a per-item pscore table is 1-D")

    def rows(self, items):
        return self.table[np.asarray(items)]


class FactoredRows(_capi._Handle):
    """Device copy of a ``FactoredFeatures`` (+ labels / pscores): an ``rfm_csr`` handle of the factored kind."""

    _destroy = "rfm_csr_destroy"
    factored = True

    def __init__(self, ctx, X: FactoredFeatures, labels=None, pscores=None, dtype="float64", row_range=None):
        """row_range=(begin, end): copy only those rows from the host (the object keeps the full shape; the caller fills
        the rest on the device and calls ``finalize()``, see rfm_b200.dist.sharded_factored_rows)."""
        super().__init__()
        n_rows = X.shape[0]
        self.ctx, self.shape, self.dtype, self.n_rows = ctx, X.shape, dtype, n_rows
        self.n_ctx = sum(b[1].shape[1] for b in X.blocks if b[0] == "ctx")
        self.has_targets = labels is not None
        arr = (RowsBlock * len(X.blocks))()
        keep = []
        for slot, b in zip(arr, X.blocks):
            if b[0] == "id":
                slot.kind, slot.key, slot.n_cols, slot.n_entities = BLOCK_ID, KEY_USER if b[1] == "user" else KEY_ITEM, b[2], b[2]
            elif b[0] == "table":
                t = b[2]
                indptr = np.ascontiguousarray(t.indptr)
                is64 = indptr.dtype == np.int64
                if not is64:
                    indptr = _capi.as_array(indptr, np.int32)
                indices = _capi.as_array(t.indices, np.int32)
                data = _capi.as_array(t.data, np.float64)
                keep += [indptr, indices, data]
                slot.kind, slot.key = BLOCK_TABLE, KEY_USER if b[1] == "user" else KEY_ITEM
                slot.n_cols, slot.n_entities = t.shape[1], t.shape[0]
                slot.indptr, slot.indptr_is_int64 = indptr.ctypes.data, int(is64)
                slot.indices, slot.data = indices.ctypes.data, data.ctypes.data
            else:
                slot.kind, slot.n_cols, slot.values = BLOCK_CTX, b[1].shape[1], b[1].ctypes.data
        by_item = isinstance(pscores, PerItem)
        ps = None if pscores is None else (pscores.table if by_item else _capi.as_array(pscores, np.float64))
        if labels is not None and (len(labels) != n_rows or ps is None or (not by_item and ps.shape[0] != n_rows)):
            raise ValueError("labels/pscores must have one entry per row")
        y, targets, label_bytes = None, None, 8
        if labels is not None:
            lab = np.asarray(labels)
            if lab.dtype in (np.int8, np.int32, np.int64):         # sent as held: 1, 4 or 8 bytes per row
                y, label_bytes = np.ascontiguousarray(lab), lab.dtype.itemsize
            else:
                y, targets = _capi.integer_labels(lab, pscores.rows(X.items) if by_item else ps)
        common = (ctx.handle, n_rows, ptr(X.users), int(X.users.dtype == np.int64), ptr(X.items),
                  int(X.items.dtype == np.int64), arr, len(X.blocks), ptr(y), label_bytes)
        if row_range is not None:
            if targets is not None:
                raise ValueError("a ranged upload takes integer labels")
            begin, end = row_range
            per_row, per_item = (None, ptr(ps)) if (by_item and y is not None) else (ptr(ps), None)
            check(lib().rfm_factored_create_range(*common, per_row, per_item, ps.shape[0] if per_item is not None else 0,
                                                  _capi.dtype_code(dtype), begin, end, byref(self.handle)))
            frac = (end - begin) / max(n_rows, 1)
            tables = sum(t.indptr.nbytes + t.indices.nbytes + t.data.nbytes for t in (b[2] for b in X.blocks if b[0] == "table"))
            rows_bytes = X.nbytes - tables + ((y.nbytes + (0 if by_item else ps.nbytes)) if y is not None else 0)
            self.h2d_bytes = int(tables + frac * rows_bytes + (ps.nbytes if (by_item and y is not None) else 0))
            return
        if by_item and y is not None:
            check(lib().rfm_factored_create_item_pscores(*common, ptr(ps), ps.shape[0], _capi.dtype_code(dtype),
                                                         byref(self.handle)))
        else:
            check(lib().rfm_factored_create(*common, ptr(ps), _capi.dtype_code(dtype), byref(self.handle)))
        if targets is not None:
            check(lib().rfm_csr_set_targets(self.handle, ptr(targets)))
        self.h2d_bytes = X.nbytes + ((y.nbytes + ps.nbytes) if y is not None else 0)

    def device_ptrs(self):
        """(user, item, ctx or None, targets) device addresses."""
        out = [c_void_p() for _ in range(4)]
        check(lib().rfm_rows_device_ptrs(self.handle, *[byref(p) for p in out]))
        return tuple(p.value for p in out)

    def finalize(self):
        check(lib().rfm_factored_finalize(self.handle))
