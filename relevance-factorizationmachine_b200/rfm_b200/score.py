"""Full-catalog scoring and exact per-user top-K on the device.

``score(u, i) = bias + alpha[u] + beta[i] + <A_u, C_i>``. The reference never scores the whole
user x item grid: ``predict`` scores the rows it is given (``src/fm.py:114-133``,
``src/mf.py:136-152``) and the evaluators rank those (``utils/evaluate.py:80-127``). This module is
the B200 path for "rank every item for every user": a bf16 tcgen05 GEMM prunes the catalog, the
survivors are re-scored in float64 and the result is proven exact per user (``csrc/score.cu``). It
equals ``predict`` on the Cartesian-product rows followed by the reference's per-user argsort.
"""
from __future__ import annotations

from ctypes import byref, c_int64, c_void_p

import numpy as np

from . import _capi
from ._capi import check, lib, ptr


def mf_factors(model):
    """(A, C, alpha, beta, bias) of a LogisticMatrixFactorization: P, Q, b_u, b_i, b (src/mf.py:165-170)."""
    return model.P(), model.Q(), model.b_u(), model.b_i(), float(getattr(model, "b", 0.0))


def fm_side(table, w, V):
    """Per-entity vector and scalar part of an FM side (user or item): for a CSR table whose row e holds
    that entity's feature columns/values,  vec_e = sum_j x_j v_j  and
    scal_e = sum_j x_j w_j + (||vec_e||^2 - sum_j x_j^2 ||v_j||^2) / 2   (src/fm.py:125-131 split by side)."""
    table = table.tocsr()
    vec = table.dot(V)
    q = np.asarray(table.power(2).dot((V ** 2).sum(axis=1))).ravel()
    scal = np.asarray(table.dot(w)).ravel() + 0.5 * ((vec ** 2).sum(axis=1) - q)
    return np.ascontiguousarray(vec), scal


def fm_factors(model, user_table, item_table):
    """(A, C, alpha, beta, bias) of a FactorizationMachines whose rows are [user features | item features].
    A per-pair context column with a fixed value belongs in ``user_table`` (same entry in every row)."""
    w, V = model.w(), model.V()
    A, alpha = fm_side(user_table, w, V)
    C, beta = fm_side(item_table, w, V)
    return A, C, alpha, beta, float(model.w0()[0])


class TopKScorer(_capi._Handle):
    _destroy = "rfm_topk_destroy"

    def __init__(self, A, C, alpha=None, beta=None, bias=0.0, device=0):
        super().__init__()
        A = _capi.as_array(A, np.float64)
        C = _capi.as_array(C, np.float64)
        if A.ndim != 2 or C.ndim != 2 or A.shape[1] != C.shape[1]:
            raise ValueError("A and C must be (n_users, k) and (n_items, k)")
        alpha = None if alpha is None else _capi.as_array(alpha, np.float64)
        beta = None if beta is None else _capi.as_array(beta, np.float64)
        if alpha is not None and alpha.shape != (A.shape[0],):
            raise ValueError("alpha must have one entry per user")
        if beta is not None and beta.shape != (C.shape[0],):
            raise ValueError("beta must have one entry per item")
        self.ctx = _capi.Context.default(device)
        self.n_users, self.n_items, self.k = A.shape[0], C.shape[0], A.shape[1]
        check(lib().rfm_topk_create(self.ctx.handle, self.n_users, self.n_items, self.k, byref(self.handle)))
        check(lib().rfm_topk_set_factors(self.handle, ptr(A), ptr(C), ptr(alpha), ptr(beta), float(bias)))
        self.last_stats = {}
        self._xchg = None           # (world, k_cap) once rfm_b200.dist.connect_scorer has mapped the peers' regions
        self._xchg_env = None

    def update(self, A, C, alpha=None, beta=None, bias=0.0):
        """New factors of the same shapes (the model moved on: another epoch, another estimator) without
        re-creating the device buffers, tensor maps and staging areas."""
        A = _capi.as_array(A, np.float64)
        C = _capi.as_array(C, np.float64)
        if A.shape != (self.n_users, self.k) or C.shape != (self.n_items, self.k):
            raise ValueError("update() keeps the shapes the scorer was created with")
        alpha = None if alpha is None else _capi.as_array(alpha, np.float64)
        beta = None if beta is None else _capi.as_array(beta, np.float64)
        check(lib().rfm_topk_set_factors(self.handle, ptr(A), ptr(C), ptr(alpha), ptr(beta), float(bias)))

    def close(self):
        """Collective when the scorer is connected to peers: no rank may unmap / free its exchange region while
        another still merges from it."""
        if self._xchg_env is not None and self.handle:
            env, self._xchg_env = self._xchg_env, None
            try:
                env.barrier()
            except Exception:
                pass
        super().close()

    def __del__(self):              # garbage collection is not a collective moment: no barrier here
        self._xchg_env = None
        super().__del__()

    def topk(self, K: int, mode: str = "tensor", item_range=None, copy: bool = True):
        """(items (n_users, K) int32, scores (n_users, K) float64): every user's K best items, best first;
        -1 / -inf pad when the catalog range holds fewer than K items. ``copy=False`` returns read-only views
        of the library's page-locked result buffers, valid until this scorer's next call (saves a host copy)."""
        if mode not in ("tensor", "exact"):
            raise ValueError("mode must be 'tensor' or 'exact'")
        begin, end = (0, self.n_items) if item_range is None else item_range
        stats = (c_int64 * 4)()
        if copy:
            items = np.empty((self.n_users, K), dtype=np.int32)
            scores = np.empty((self.n_users, K), dtype=np.float64)
            check(lib().rfm_topk_run(self.handle, K, 0 if mode == "tensor" else 1, begin, end, ptr(items),
                                     ptr(scores), stats))
        else:
            import ctypes
            check(lib().rfm_topk_run(self.handle, K, 0 if mode == "tensor" else 1, begin, end, None, None, stats))
            pi, ps = c_void_p(), c_void_p()
            check(lib().rfm_topk_result_host(self.handle, K, byref(pi), byref(ps)))
            n = self.n_users * K
            items = np.ctypeslib.as_array(ctypes.cast(pi, ctypes.POINTER(ctypes.c_int32)), shape=(n,)).reshape(self.n_users, K)
            scores = np.ctypeslib.as_array(ctypes.cast(ps, ctypes.POINTER(ctypes.c_double)), shape=(n,)).reshape(self.n_users, K)
            items.flags.writeable = False
            scores.flags.writeable = False
        self.last_stats = {"tensor_core_path": bool(stats[0]), "users_ranked_exactly": int(stats[1]),
                           "candidates": int(stats[2]), "sample_stride": int(stats[3])}
        return items, scores


def merge_topk(items_list, scores_list, K):
    """Merge per-shard top-K lists (item-sharded scoring) into the global top-K
    with the canonical order: score descending, larger item id first among exact ties."""
    items = np.concatenate(items_list, axis=1)
    scores = np.concatenate(scores_list, axis=1)
    order = np.lexsort((-items.astype(np.int64), -scores), axis=1)[:, :K]
    rows = np.arange(items.shape[0])[:, None]
    return items[rows, order], scores[rows, order]
