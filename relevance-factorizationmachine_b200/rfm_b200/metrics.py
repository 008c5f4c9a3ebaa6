"""Ranking metric functions with the reference's names and signatures (``utils/metrics.py``).

These host functions take one user's score-sorted arrays, exactly like the reference's, and are
what ``metric_candidates`` exposes for API fidelity (``utils/metrics.py:169-178``). The
evaluators in ``rfm_b200.evaluate`` do NOT loop over users with them: they compute the same
quantities for all users at once on the device (``csrc/rank.cu``).
"""
from __future__ import annotations

from collections import Counter
from typing import Callable, Dict, List, Union

import numpy as np


def calc_average_precision_at_k(y_true_sorted_by_scores: np.ndarray, k: int) -> float:
    ap = 0.0
    if not np.sum(y_true_sorted_by_scores) == 0:
        hits = np.cumsum(y_true_sorted_by_scores[:k])
        for i in range(min(k, len(y_true_sorted_by_scores))):
            if y_true_sorted_by_scores[i] >= 1:
                ap += hits[i] / (i + 1)
    return ap


def calc_recall_at_k(y_true_sorted_by_scores: np.ndarray, k: int) -> float:
    total = np.sum(y_true_sorted_by_scores)
    return 0.0 if total == 0 else np.sum(y_true_sorted_by_scores[:k]) / total


def _discounts(n: int) -> np.ndarray:
    return np.log2(np.arange(1, n + 1) + 1)


def calc_ips_of_dcg_at_k(y_true_sorted_by_scores: np.ndarray, k: int,
                         pscores_sorted_by_scores: np.ndarray) -> float:
    if np.sum(y_true_sorted_by_scores) == 0:
        return np.nan
    tail = y_true_sorted_by_scores[1:k]
    den = pscores_sorted_by_scores[1:k] * _discounts(tail.shape[0])
    return y_true_sorted_by_scores[0] / pscores_sorted_by_scores[0] + np.sum(tail / den)


def calc_dcg_at_k(y_true_sorted_by_scores: np.ndarray, k: int) -> float:
    if np.sum(y_true_sorted_by_scores) == 0:
        return np.nan
    tail = y_true_sorted_by_scores[1:k]
    return y_true_sorted_by_scores[0] + np.sum(tail / _discounts(tail.shape[0]))


def return_exposure_at_k(pscores_sorted_by_scores: np.ndarray, k: int) -> Union[float, None]:
    return pscores_sorted_by_scores[k - 1] if len(pscores_sorted_by_scores) >= k else np.nan


def gini_from_counts(rec_freqs: np.ndarray) -> float:
    """Gini coefficient of per-item recommendation counts (``utils/metrics.py:143-149``)."""
    n_items = rec_freqs.shape[0]
    freqs = np.sort(rec_freqs, kind="stable")
    idx = np.arange(1, n_items + 1)
    return np.sum((2 * idx - n_items - 1) * freqs) / (n_items * np.sum(freqs))


def calc_gini_at_k(rec_items: List[int], n_items: int) -> float:
    counter = Counter(rec_items)
    return gini_from_counts(np.array([counter.get(i, 0) for i in range(n_items)]))


def calc_catalog_coverage_at_k(rec_items: np.ndarray, n_items: int) -> float:
    return len(set(rec_items)) / n_items


metric_candidates: Dict[str, Callable] = {
    "Recall": calc_recall_at_k,
    "MAP": calc_average_precision_at_k,
    "DCG": calc_dcg_at_k,
    "ME": return_exposure_at_k,
    "CatalogCoverage": calc_catalog_coverage_at_k,
    "Gini": calc_gini_at_k,
}
