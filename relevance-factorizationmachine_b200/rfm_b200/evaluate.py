"""``TestEvaluator`` / ``ValEvaluator`` -- drop-ins for ``utils/evaluate.py:41-239``.

Constructor fields, ``evaluate`` signatures, return types and error behaviour follow the
reference; the per-user Python loop and the pandas group-by are replaced by one device call
(``rfm_ranker_evaluate``, ``csrc/rank.cu``). ``interaction_df`` may be a pandas DataFrame (as in
the reference) or any mapping of equally long columns ``user, item, label, pscore[, ones_pscore]``.

Tie rule: score descending, later row first among exactly equal scores
(``argsort(kind="stable")[::-1]``); the reference's default-sort order among exact ties is
NumPy-build dependent.
"""
from __future__ import annotations

import warnings
from abc import ABC, abstractmethod
from collections import defaultdict
from ctypes import byref, c_int64
from dataclasses import dataclass
from typing import Dict, Tuple, Union

import numpy as np

from . import _capi
from ._capi import RANK_COLS, RANK_NCOLS, check, lib, ptr
from .metrics import calc_ips_of_dcg_at_k, gini_from_counts, metric_candidates

METRIC_NAME_ERROR_MESSAGE = "metric_name must be in {}. metric_name: '{}'"


class _Ranker(_capi._Handle):
    _destroy = "rfm_ranker_destroy"

    def __init__(self, ctx, users, items, labels, pscores, n_items):
        super().__init__()
        users = _capi.as_array(users, np.int64)
        items = _capi.as_array(items, np.int64)
        labels = _capi.as_array(labels, np.float64)
        pscores = _capi.as_array(pscores, np.float64)
        check(lib().rfm_ranker_create(ctx.handle, users.shape[0], ptr(users), ptr(items), ptr(labels),
                                      ptr(pscores), n_items, byref(self.handle)))
        self.n_rows, self.n_items = users.shape[0], n_items
        n = c_int64()
        check(lib().rfm_ranker_num_users(self.handle, byref(n)))
        self.n_users = n.value

    def scores_ptr(self) -> int:
        """Device address of the ranker's score buffer (double [n_rows], frame order)."""
        from ctypes import c_void_p
        p = c_void_p()
        check(lib().rfm_ranker_scores_ptr_dev(self.handle, byref(p)))
        return p.value

    def evaluate_dev(self, K, slot: int, max_slots: int):
        """Rank the scores already in the device buffer; metric rows go to history slot ``slot`` on the device
        (asynchronous: no copy, no synchronisation)."""
        K = np.ascontiguousarray(K, dtype=np.int32)
        check(lib().rfm_ranker_evaluate_dev(self.handle, None, ptr(K), len(K), slot, max_slots))

    def read_slots(self, first: int, n: int, n_k: int) -> np.ndarray:
        out = np.zeros((n, n_k, RANK_NCOLS))
        check(lib().rfm_ranker_read_slots(self.handle, first, n, n_k, ptr(out)))
        return out

    def evaluate(self, scores, K, want_hits=False, want_top=False):
        scores = _capi.as_array(scores, np.float64)
        if scores.shape[0] != self.n_rows:
            raise ValueError("y_scores has %d entries, interaction_df has %d rows" % (scores.shape[0], self.n_rows))
        K = np.ascontiguousarray(K, dtype=np.int32)
        metrics = np.zeros((len(K), RANK_NCOLS))
        hits = np.zeros((len(K), self.n_items), dtype=np.int32) if want_hits else None
        top = np.full((self.n_users, int(K.max())), -1, dtype=np.int64) if want_top else None
        check(lib().rfm_ranker_evaluate(self.handle, ptr(scores), ptr(K), len(K), ptr(metrics), ptr(hits), ptr(top)))
        return metrics, hits, top


def _column(frame, name):
    col = frame[name]
    return col.to_numpy() if hasattr(col, "to_numpy") else np.asarray(col)


@dataclass
class _BaseEvaluator(ABC):
    interaction_df: object
    features: Dict[str, object]

    @abstractmethod
    def evaluate(self, *args, **kwargs):
        ...

    def _ranker(self, pscore_name="pscore", n_items=None, device=0):
        """Device copy of the frame grouped by user, rebuilt when the frame's columns are replaced or n_items
        changes (the reference regroups on every call, ``utils/evaluate.py:141-156``). The cache keys on the identity
        of the column arrays: replacing a column is seen, editing one IN PLACE is not -- call ``reset_cache()``."""
        cache = self.__dict__.setdefault("_rankers", {})
        cols = [np.asarray(_column(self.interaction_df, name)) for name in ("user", "item", "label", pscore_name)]
        items = cols[1]
        if n_items is None:
            n_items = int(items.max()) + 1 if items.size else 1
        stamp = tuple((c.__array_interface__["data"][0], c.shape[0], c.dtype.str) for c in cols) + (int(n_items),)
        hit = cache.get(pscore_name)
        if hit is None or hit[0] != stamp:
            if hit is not None:
                hit[1].close()
            cache[pscore_name] = (stamp, _Ranker(_capi.Context.default(device), cols[0], items, cols[2], cols[3],
                                                 n_items), cols)     # cols kept alive: the stamp holds their addresses
        return cache[pscore_name][1]

    def reset_cache(self):
        """Drop the device copies (after editing a column of ``interaction_df`` in place)."""
        for _, ranker, _cols in self.__dict__.pop("_rankers", {}).values():
            ranker.close()

    def _grouped(self, y_scores, pscore_name):
        """``{user: {"items", "labels", "y_scores", "pscores"}}`` -- what the reference's pandas group-by returns
        (``utils/evaluate.py:141-156, 223-239``): users ascending, a user's rows in frame order."""
        self._record_scores(y_scores)
        users = np.asarray(_column(self.interaction_df, "user"))
        order = np.argsort(users, kind="stable")
        su = users[order]
        starts = np.flatnonzero(np.r_[True, su[1:] != su[:-1]]) if len(su) else np.zeros(0, dtype=np.int64)
        ends = np.r_[starts[1:], len(order)]
        cols = {"items": "item", "labels": "label", "pscores": pscore_name}
        data = {k: np.asarray(_column(self.interaction_df, v))[order] for k, v in cols.items()}
        data["y_scores"] = np.asarray(y_scores)[order]
        return {su[b].item(): {"items": data["items"][b:e], "labels": data["labels"][b:e],
                               "y_scores": data["y_scores"][b:e], "pscores": data["pscores"][b:e]}
                for b, e in zip(starts, ends)}

    def _record_scores(self, y_scores):
        # the reference mutates interaction_df["y_score"] (evaluate.py:141, 223); keep that visible
        try:
            self.interaction_df["y_score"] = y_scores
        except Exception:
            pass


@dataclass
class TestEvaluator(_BaseEvaluator):
    K: Tuple[int]
    used_metrics: set
    n_items: int

    __test__ = False  # not a pytest class

    def __post_init__(self) -> None:
        self.metric_functions = {"ME": metric_candidates["ME"]}
        for metric_name in self.used_metrics:
            if metric_name not in metric_candidates:
                raise ValueError(METRIC_NAME_ERROR_MESSAGE.format(metric_candidates.keys(), metric_name))
            self.metric_functions[metric_name] = metric_candidates[metric_name]

    def evaluate(self, y_scores: np.ndarray) -> defaultdict:
        self._record_scores(y_scores)
        need_hits = any(m in self.metric_functions for m in ("CatalogCoverage", "Gini"))
        metrics, hits, _ = self._ranker(n_items=self.n_items).evaluate(y_scores, list(self.K), want_hits=need_hits)
        results = defaultdict(list)
        for j, _k in enumerate(self.K):
            row = metrics[j]
            kept = row[RANK_COLS["USERS"]]
            for name in self.metric_functions:
                if name == "CatalogCoverage":
                    results[name].append(int(row[RANK_COLS["COVERED"]]) / self.n_items)
                elif name == "Gini":
                    results[name].append(gini_from_counts(hits[j].astype(np.int64)))
                elif name == "ME":
                    n = row[RANK_COLS["ME_COUNT"]]
                    if n == 0:
                        warnings.warn("Mean of empty slice", RuntimeWarning)
                    results[name].append(row[RANK_COLS["ME_SUM"]] / n if n else np.nan)
                else:
                    col = {"DCG": "DCG_SUM", "Recall": "RECALL_SUM", "MAP": "MAP_SUM"}[name]
                    results[name].append(row[RANK_COLS[col]] / kept if kept else np.nan)
        return results

    def top_rows(self, y_scores: np.ndarray, k: int) -> np.ndarray:
        """(n_users, k) original row ids of every user's top-k (users ascending, -1 padded)."""
        _, _, top = self._ranker(n_items=self.n_items).evaluate(y_scores, [k], want_top=True)
        return top

    def _group_by_user_data(self, y_scores: np.ndarray) -> Dict[str, Dict[str, np.ndarray]]:
        """``utils/evaluate.py:129-156`` (host helper; ``evaluate`` itself groups once on the device)."""
        return self._grouped(y_scores, "pscore")


@dataclass
class ValEvaluator(_BaseEvaluator):
    k: int
    metric_name: str

    def __post_init__(self) -> None:
        if self.metric_name == "DCG":
            self.metric_func = calc_ips_of_dcg_at_k
        else:
            raise ValueError("You can use only DCG metric.")

    def evaluate(self, y_scores: np.ndarray, estimator: str) -> float:
        self._record_scores(y_scores)
        pscore_name = "pscore" if estimator == "IPS" else "ones_pscore"
        metrics, _, _ = self._ranker(pscore_name).evaluate(y_scores, [self.k])
        kept = metrics[0, RANK_COLS["USERS"]]
        return metrics[0, RANK_COLS["IPSDCG_SUM"]] / kept if kept else np.nan

    # ---- device-chained use inside fit(evaluator=...) ------------------------------
    def device_chain(self, estimator: str, device: int = 0):
        """The ranker a model's ``fit`` writes its per-epoch scores into (``rfm_*_predict_dev`` ->
        ``rfm_ranker_evaluate_dev``): the epoch loop of ``src/fm.py:104-110`` without a host round trip."""
        return self._ranker("pscore" if estimator == "IPS" else "ones_pscore", device=device)

    def chain_results(self, ranker, n_epochs: int) -> list:
        """IPS-DCG@k of every epoch (``np.mean`` over the kept users, ``utils/evaluate.py:207``) read back once."""
        rows = ranker.read_slots(0, n_epochs, 1)[:, 0, :]
        kept = rows[:, RANK_COLS["USERS"]]
        with np.errstate(invalid="ignore", divide="ignore"):
            vals = np.where(kept > 0, rows[:, RANK_COLS["IPSDCG_SUM"]] / kept, np.nan)
        return [float(v) for v in vals]

    def _group_by_user_data(self, y_scores: np.ndarray, estimator: str) -> Dict[str, Dict[str, np.ndarray]]:
        """``utils/evaluate.py:209-239`` (host helper; ``evaluate`` itself groups once on the device)."""
        return self._grouped(y_scores, "pscore" if estimator == "IPS" else "ones_pscore")


@dataclass
class FullCatalogEvaluator:
    """TestEvaluator's metrics when EVERY item of the catalog is a candidate for every user.

    The reference only ranks the rows of ``interaction_df`` (``utils/evaluate.py:80-127``); scoring the
    whole user x item grid is a new capability. This class is defined to equal
    ``TestEvaluator`` on the Cartesian-product frame whose label is the held-out label where one exists
    and 0 elsewhere, and whose pscore is the item's exposure: scores come from the device scorer
    (``rfm_b200.score.TopKScorer``: tcgen05 GEMM prune + exact float64 top-K), and the metric reductions
    run in the device ranker on the top-max(K) rows of every user.
    """

    interaction_df: object          # held-out rows with columns user, item, label
    item_pscores: np.ndarray        # (n_items,) exposure of every item (MeanExposure@K)
    K: Tuple[int]
    used_metrics: set
    n_users: int
    n_items: int

    def __post_init__(self) -> None:
        from scipy.sparse import csr_matrix
        self.metric_names = ["ME"]
        for metric_name in self.used_metrics:
            if metric_name not in metric_candidates:
                raise ValueError(METRIC_NAME_ERROR_MESSAGE.format(metric_candidates.keys(), metric_name))
            if metric_name != "ME":
                self.metric_names.append(metric_name)
        users = _column(self.interaction_df, "user").astype(np.int64)
        items = _column(self.interaction_df, "item").astype(np.int64)
        labels = _column(self.interaction_df, "label").astype(np.float64)
        self._labels = csr_matrix((labels, (users, items)), shape=(self.n_users, self.n_items))
        self._labels.sum_duplicates()
        self._labels.sort_indices()
        self._totals = np.asarray(self._labels.sum(axis=1)).ravel().astype(np.float64)
        self.item_pscores = _capi.as_array(self.item_pscores, np.float64)
        self.last_stats = {}
        self._dev = {}

    def _device_state(self, ctx):
        """Held-out labels (CSR by user) and item exposures, uploaded once per context."""
        key = id(ctx)
        if key not in self._dev:
            self._dev[key] = _CatalogEval(ctx, self.n_users, self.n_items, self._labels, self.item_pscores)
        return self._dev[key]

    def _results(self, metrics, hits) -> defaultdict:
        results = defaultdict(list)
        for j, _k in enumerate(self.K):
            row = metrics[j]
            kept = row[RANK_COLS["USERS"]]
            for name in self.metric_names:
                if name == "CatalogCoverage":
                    results[name].append(int(row[RANK_COLS["COVERED"]]) / self.n_items)
                elif name == "Gini":
                    results[name].append(gini_from_counts(hits[j].astype(np.int64)))
                elif name == "ME":
                    n = row[RANK_COLS["ME_COUNT"]]
                    results[name].append(row[RANK_COLS["ME_SUM"]] / n if n else np.nan)
                else:
                    col = {"DCG": "DCG_SUM", "Recall": "RECALL_SUM", "MAP": "MAP_SUM"}[name]
                    results[name].append(row[RANK_COLS[col]] / kept if kept else np.nan)
        return results

    def evaluate(self, scorer, mode: str = "tensor", env=None) -> defaultdict:
        """``scorer`` is a ``TopKScorer`` (see ``rfm_b200.score.fm_factors`` / ``mf_factors``). The ranked lists
        never leave the device: top-K -> label look-up -> metric reductions -> one small read-back
        (``rfm_catalog_eval_run``). With ``env`` (a ``rfm_b200.dist.DistEnv``) the catalog is item-sharded over
        the ranks, every rank reduces the users it owns after the merge, and the partial sums and per-item hit
        counts are all-reduced; every rank returns the same metrics."""
        from ctypes import c_int64, c_void_p
        k_max = int(max(self.K))
        if k_max > scorer.n_items:
            k_max = scorer.n_items
        state = self._device_state(scorer.ctx)
        K = np.ascontiguousarray([int(k) for k in self.K], dtype=np.int32)     # lists shorter than k: ME@k is nan
        need_hits = any(m in self.metric_names for m in ("CatalogCoverage", "Gini"))
        sharded = env is not None and env.world > 1
        if sharded:
            from . import dist as rdist
            rdist.connect_scorer(scorer, env, k_max)
            rng, stats = (c_int64 * 2)(), (c_int64 * 4)()
            check(lib().rfm_topk_run_sharded(scorer.handle, k_max, 0 if mode == "tensor" else 1, None, None, rng, stats))
            user_begin, n_rows = rng[0], rng[1] - rng[0]
        else:
            stats = (c_int64 * 4)()
            check(lib().rfm_topk_run(scorer.handle, k_max, 0 if mode == "tensor" else 1, 0, scorer.n_items, None, None,
                                     stats))
            user_begin, n_rows = 0, self.n_users
        scorer.last_stats = {"tensor_core_path": bool(stats[0]), "users_ranked_exactly": int(stats[1]),
                             "candidates": int(stats[2]), "sample_stride": int(stats[3])}
        self.last_stats = dict(scorer.last_stats)
        ip, sp = c_void_p(), c_void_p()
        check(lib().rfm_topk_result_ptr_dev(scorer.handle, byref(ip), byref(sp)))
        metrics = np.zeros((len(K), RANK_NCOLS))
        hits = np.zeros((len(K), self.n_items), dtype=np.int32) if (need_hits or sharded) else None
        check(lib().rfm_catalog_eval_run(state.handle, ip, k_max, user_begin, n_rows, ptr(K), len(K), ptr(metrics),
                                         ptr(hits)))
        if sharded:
            torch = env.torch
            dev = "cuda:%d" % env.device
            tm = torch.from_numpy(metrics).to(dev)
            th = torch.from_numpy(hits.astype(np.int64)).to(dev)
            env.dist.all_reduce(tm)                      # sums of per-user terms and user counts
            env.dist.all_reduce(th)                      # per-item hit counts (coverage = non-zeros, Gini)
            metrics = tm.cpu().numpy()
            hits = th.cpu().numpy()
            metrics[:, RANK_COLS["COVERED"]] = (hits != 0).sum(axis=1)
        return self._results(metrics, hits)


class _CatalogEval(_capi._Handle):
    _destroy = "rfm_catalog_eval_destroy"

    def __init__(self, ctx, n_users, n_items, labels_csr, item_pscores):
        super().__init__()
        indptr = _capi.as_array(labels_csr.indptr, np.int64)
        items = _capi.as_array(labels_csr.indices, np.int32)
        vals = _capi.as_array(labels_csr.data, np.float64)
        ps = _capi.as_array(item_pscores, np.float64)
        if ps.shape != (n_items,):
            raise ValueError("item_pscores must have one entry per item")
        check(lib().rfm_catalog_eval_create(ctx.handle, n_users, n_items, ptr(indptr), ptr(items), ptr(vals), ptr(ps),
                                            byref(self.handle)))
