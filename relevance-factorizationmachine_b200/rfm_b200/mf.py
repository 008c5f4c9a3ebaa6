"""``LogisticMatrixFactorization`` -- drop-in for the reference class (``src/mf.py:16-216``).

Same dataclass fields in the same order (``estimator, n_epochs, n_factors, lr, batch_size,
seed, n_users, n_items, reg, alpha=4.0, evaluator=None``), same ``fit`` / ``predict`` and
parameter holders ``P / Q / b_u / b_i`` and the float ``b`` set by ``fit``. The per-sample
sequential SGD of ``src/mf.py:97-108`` runs on the device as a wavefront schedule
(``csrc/mf.cu``) and reproduces the reference's update order exactly.
"""
from __future__ import annotations

import weakref
from ctypes import byref, c_double, c_void_p
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np

from . import _capi
from ._capi import check, lib, ptr
from .base import EvalChain, PointwiseBaseRecommender
from .optimizer import SGD
from .sampler import LegacyBatchPrefetcher


class _MfDevice(_capi._Handle):
    _destroy = "rfm_mf_destroy"

    def __init__(self, ctx, n_users, n_items, n_factors, dtype):
        super().__init__()
        check(lib().rfm_mf_create(ctx.handle, n_users, n_items, n_factors, _capi.dtype_code(dtype),
                                  byref(self.handle)))


class PairRows(_capi._Handle):
    """Device copy of ``(N, 2)`` ``[user, item]`` rows (+ labels / pscores)."""

    _destroy = "rfm_pairs_destroy"

    def __init__(self, ctx, pairs, labels=None, pscores=None, dtype="float64"):
        super().__init__()
        pairs = np.asarray(pairs)
        if pairs.ndim != 2 or pairs.shape[1] != 2:
            raise ValueError("MF features must have shape (N, 2): [user, item]")
        pairs = _capi.as_array(pairs, np.int64)
        ps = None if pscores is None else _capi.as_array(pscores, np.float64)
        if labels is not None and (len(labels) != pairs.shape[0] or ps is None or ps.shape[0] != pairs.shape[0]):
            raise ValueError("labels/pscores must have one entry per row")
        y, targets = (None, None) if labels is None else _capi.integer_labels(labels, ps)
        check(lib().rfm_pairs_create(ctx.handle, pairs.shape[0], ptr(pairs), ptr(y), ptr(ps),
                                     _capi.dtype_code(dtype), byref(self.handle)))
        if targets is not None:
            check(lib().rfm_pairs_set_targets(self.handle, ptr(targets)))
        self.n_rows = pairs.shape[0]


@dataclass
class LogisticMatrixFactorization(PointwiseBaseRecommender):
    n_users: int
    n_items: int
    reg: float
    alpha: float = 4.0
    evaluator: Optional[object] = None
    # ---- extensions (defaults keep reference behaviour) ----
    dtype: str = "float64"
    device: int = 0
    progress: bool = False

    def __post_init__(self) -> None:
        _capi.dtype_code(self.dtype)
        np.random.seed(self.seed)                                   # same draw order as src/mf.py:37-62
        limit = self.alpha * np.sqrt(6 / self.n_factors)
        self.P = SGD(params=np.random.uniform(low=-limit, high=limit, size=(self.n_users, self.n_factors)),
                     lr=self.lr)
        self.Q = SGD(params=np.random.uniform(low=-limit, high=limit, size=(self.n_items, self.n_factors)),
                     lr=self.lr)
        self.b_u = SGD(params=np.random.normal(scale=0.001, size=self.n_users), lr=self.lr)
        self.b_i = SGD(params=np.random.normal(scale=0.001, size=self.n_items), lr=self.lr)
        if self.evaluator is not None:
            self.val_metrics = []
            self.model_name = "MF"
        self._ctx = None
        self._dev = None
        self._synced = None
        self._rows_cache: Dict[int, tuple] = {}
        self.last_fit_stats = {}

    def _context(self):
        if self._ctx is None:
            self._ctx = _capi.Context.default(self.device)
            self._dev = _MfDevice(self._ctx, self.n_users, self.n_items, self.n_factors, self.dtype)
        return self._ctx

    def _host_state(self):
        holders = (self.P, self.Q, self.b_u, self.b_i)
        return tuple((id(h.params), h.version) for h in holders) + (float(getattr(self, "b", 0.0)),)

    def sync_to_device(self, force: bool = False) -> None:
        self._context()
        if force or self._synced != self._host_state():
            P = _capi.as_array(self.P.params, np.float64)
            Q = _capi.as_array(self.Q.params, np.float64)
            bu = _capi.as_array(self.b_u.params, np.float64)
            bi = _capi.as_array(self.b_i.params, np.float64)
            if P.shape != (self.n_users, self.n_factors) or Q.shape != (self.n_items, self.n_factors):
                raise ValueError("parameter arrays changed shape")
            check(lib().rfm_mf_set_params(self._dev.handle, ptr(P), ptr(Q), ptr(bu), ptr(bi),
                                          float(getattr(self, "b", 0.0))))
            self._synced = self._host_state()

    def sync_to_host(self) -> None:
        P = np.empty((self.n_users, self.n_factors))
        Q = np.empty((self.n_items, self.n_factors))
        bu, bi = np.empty(self.n_users), np.empty(self.n_items)
        check(lib().rfm_mf_get_params(self._dev.handle, ptr(P), ptr(Q), ptr(bu), ptr(bi)))
        for holder, new in ((self.P, P), (self.Q, Q), (self.b_u, bu), (self.b_i, bi)):
            if (isinstance(holder.params, np.ndarray) and holder.params.shape == new.shape
                    and holder.params.dtype == np.float64 and holder.params.flags.writeable):
                holder.params[...] = new
            else:
                holder.params = new
        self._synced = self._host_state()

    def _rows(self, X, labels=None, pscores=None):
        """Device copy of (N, 2) rows, cached on weak references to X, labels and pscores (in-place edits are not
        seen: ``reset_rows_cache()``)."""
        key = (id(X), id(labels), id(pscores))
        hit = self._rows_cache.get(key)
        if hit is not None and _capi.refs_match(hit[0], X, labels, pscores):
            return hit[1]
        rows = PairRows(self._context(), X, labels, pscores, self.dtype)
        refs = _capi.weak_refs(X, labels, pscores) if isinstance(X, np.ndarray) else None
        if refs is not None:
            self._rows_cache[key] = (refs, rows)
            if len(self._rows_cache) > 8:
                self._rows_cache.pop(next(iter(self._rows_cache)))
        return rows

    def reset_rows_cache(self) -> None:
        self._rows_cache.clear()

    # ---- reference API -----------------------------------------------------------------------
    def fit(self, train, val) -> tuple:
        ctx = self._context()
        self.b = np.mean(train["labels"])                           # src/mf.py:84
        n_rows = np.asarray(train["features"]).shape[0]
        if self.batch_size > n_rows:
            raise ValueError("Cannot sample %d out of arrays with dim %d when replace is False"
                             % (self.batch_size, n_rows))
        train_rows = self._rows(train["features"], train["labels"], train["pscores"])
        val_rows = self._rows(val["features"], val["labels"], val["pscores"])
        self.sync_to_device()
        eval_rows = self._rows(self.evaluator.features[self.model_name]) if self.evaluator is not None else None
        chain = None
        if eval_rows is not None and EvalChain.supported(self.evaluator):
            chain = EvalChain(self.evaluator, self.estimator, eval_rows.n_rows, self.n_epochs, self.device)
        epochs = range(self.n_epochs)
        prefetch = LegacyBatchPrefetcher(n_rows, self.batch_size, epochs)
        it = epochs
        if self.progress:
            from tqdm import tqdm
            it = tqdm(epochs)
        train_loss, val_loss = [], []
        tl, vl = c_double(), c_double()
        launches0 = ctx.launch_count()
        try:
            for epoch in it:
                idx = prefetch.next()
                check(lib().rfm_mf_train_epoch(self._dev.handle, train_rows.handle, val_rows.handle, ptr(idx),
                                               self.batch_size, self.lr, self.reg, byref(tl), byref(vl)))
                train_loss.append(tl.value)
                val_loss.append(vl.value)
                if chain is not None:           # predict -> rank -> metric slot on the device (no read-back per epoch)
                    check(lib().rfm_mf_predict_dev(self._dev.handle, eval_rows.handle, c_void_p(chain.scores_ptr)))
                    chain.after_epoch(epoch)
                elif eval_rows is not None:
                    scores = np.empty(eval_rows.n_rows)
                    check(lib().rfm_mf_predict(self._dev.handle, eval_rows.handle, ptr(scores)))
                    self.val_metrics.append(self.evaluator.evaluate(y_scores=scores, estimator=self.estimator))
            if chain is not None:
                self.val_metrics.extend(chain.finish(self.n_epochs))
        finally:
            prefetch.close()
        self.last_fit_stats = {"gpu_launches": ctx.launch_count() - launches0}
        self.sync_to_host()
        return train_loss, val_loss

    def predict(self, X) -> np.ndarray:
        self.sync_to_device()
        rows = self._rows(X)
        out = np.empty(rows.n_rows)
        check(lib().rfm_mf_predict(self._dev.handle, rows.handle, ptr(out)))
        return out

    def logloss(self, data) -> float:
        self.sync_to_device()
        rows = self._rows(data["features"], data["labels"], data["pscores"])
        out = c_double()
        check(lib().rfm_mf_logloss(self._dev.handle, rows.handle, byref(out)))
        return out.value
