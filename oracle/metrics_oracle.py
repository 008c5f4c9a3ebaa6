"""ORACLE (test infrastructure, not product code) -- CPU restatement of the ranking metrics
and the two evaluators.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this module.

Parity status: PINNED against ``tests/golden/coat_eval_*.npz`` (outputs of the unmodified
reference's ``TestEvaluator`` / ``ValEvaluator``).

Restated (reference file:line):

* ``dcg_at_k``      -- ``utils/metrics.py:83-107``  (nan when the user has no positive)
* ``ips_dcg_at_k``  -- ``utils/metrics.py:53-80``
* ``exposure_at_k`` -- ``utils/metrics.py:110-127`` (nan when the list is shorter than k)
* ``recall_at_k``   -- ``utils/metrics.py:32-50``;  ``ap_at_k`` -- ``utils/metrics.py:9-29``
* ``gini``          -- ``utils/metrics.py:130-149``; ``coverage`` -- ``utils/metrics.py:152-166``
* ``test_evaluate`` -- ``utils/evaluate.py:80-127`` (+ group-by ``:129-156``): users without a
  positive label are skipped for every metric; DCG/ME/Recall/MAP are nanmean over the kept
  users; coverage/Gini use the union / multiset of the kept users' top-k items.
* ``val_evaluate``  -- ``utils/evaluate.py:183-207``: plain mean of IPS-DCG@k over kept users.

Tie rule. The reference ranks with ``scores.argsort()[::-1]`` (``evaluate.py:93,197``), NumPy's
default unstable sort, whose order among EXACTLY equal scores depends on the NumPy build
(SURVEY.md F10). The canonical order used here and by the CUDA path is
``argsort(kind="stable")[::-1]``: score descending, and among equal scores the row that
comes LATER in the user's candidate list first. On tie-free inputs this is identical to the
reference; ``rank_equal_modulo_ties`` is the comparator for tie-heavy inputs.
"""
from __future__ import annotations

from collections import defaultdict

import numpy as np


def dcg_at_k(y_sorted, k):
    if np.sum(y_sorted) == 0:
        return np.nan
    mol = y_sorted[1:k]
    return float(y_sorted[0] + np.sum(mol / np.log2(np.arange(1, mol.shape[0] + 1) + 1)))


def ips_dcg_at_k(y_sorted, k, ps_sorted):
    if np.sum(y_sorted) == 0:
        return np.nan
    mol = y_sorted[1:k]
    den = ps_sorted[1:k] * np.log2(np.arange(1, mol.shape[0] + 1) + 1)
    return float(y_sorted[0] / ps_sorted[0] + np.sum(mol / den))


def exposure_at_k(ps_sorted, k):
    return float(ps_sorted[k - 1]) if len(ps_sorted) >= k else np.nan


def recall_at_k(y_sorted, k):
    tot = np.sum(y_sorted)
    return 0.0 if tot == 0 else float(np.sum(y_sorted[:k]) / tot)


def ap_at_k(y_sorted, k):
    ap = 0.0
    if np.sum(y_sorted) != 0:
        for i in range(min(k, len(y_sorted))):
            if y_sorted[i] >= 1:
                ap += np.sum(y_sorted[: i + 1]) / (i + 1)
    return float(ap)


def coverage(rec_items, n_items):
    return len(set(int(i) for i in rec_items)) / n_items


def gini(rec_items, n_items):
    freq = np.bincount(np.asarray(rec_items, dtype=np.int64), minlength=n_items)[:n_items]
    freq = np.sort(freq, kind="stable")
    idx = np.arange(1, n_items + 1)
    return float(np.sum((2 * idx - n_items - 1) * freq) / (n_items * np.sum(freq)))


def canonical_order(scores):
    return np.argsort(scores, kind="stable")[::-1]


def group_by_user(users):
    """pandas ``groupby("user")`` semantics: groups in ascending user id, rows inside a
    group in original order. Returns (unique_users, order, ptr)."""
    users = np.asarray(users)
    order = np.argsort(users, kind="stable")
    uniq, counts = np.unique(users, return_counts=True)
    ptr = np.zeros(uniq.shape[0] + 1, dtype=np.int64)
    np.cumsum(counts, out=ptr[1:])
    return uniq, order, ptr


def ranked_lists(frame, y_scores):
    """Yield (user, ranked row ids) for every user, canonical order."""
    uniq, order, ptr = group_by_user(frame["user"])
    y_scores = np.asarray(y_scores)
    for g, user in enumerate(uniq):
        rows = order[ptr[g]: ptr[g + 1]]
        yield user, rows[canonical_order(y_scores[rows])]


_PER_USER = {"DCG": dcg_at_k, "Recall": recall_at_k, "MAP": ap_at_k}


def test_evaluate(frame, y_scores, K, used_metrics, n_items):
    names = ["ME"] + [m for m in used_metrics if m != "ME"]
    per_user = defaultdict(lambda: defaultdict(list))
    for _, rows in ranked_lists(frame, y_scores):
        y = frame["label"][rows]
        if np.sum(y) == 0:
            continue
        ps, items = frame["pscore"][rows], frame["item"][rows]
        for k in K:
            for name in names:
                if name in ("CatalogCoverage", "Gini"):
                    per_user[name][k].extend(items[:k])
                elif name == "ME":
                    per_user[name][k].append(exposure_at_k(ps, k))
                else:
                    per_user[name][k].append(_PER_USER[name](y, k))
    out = defaultdict(list)
    for k in K:
        for name in names:
            if name == "CatalogCoverage":
                out[name].append(coverage(per_user[name][k], n_items))
            elif name == "Gini":
                out[name].append(gini(per_user[name][k], n_items))
            else:
                out[name].append(float(np.nanmean(per_user[name][k])))
    return out


def val_evaluate(frame, y_scores, k, estimator):
    ps_all = frame["pscore"] if estimator == "IPS" else frame["ones_pscore"]
    vals = []
    for _, rows in ranked_lists(frame, y_scores):
        y = frame["label"][rows]
        if np.sum(y) == 0:
            continue
        vals.append(ips_dcg_at_k(y, k, ps_all[rows]))
    return float(np.mean(vals))


test_evaluate.__test__ = False  # not a pytest test despite the name (mirrors TestEvaluator)


def rank_equal_modulo_ties(rows_a, rows_b, scores, k):
    """True when two top-k row lists agree except for permutations inside groups of
    exactly equal score (including a tie group cut by the k boundary)."""
    rows_a, rows_b = np.asarray(rows_a)[:k], np.asarray(rows_b)[:k]
    sa, sb = scores[rows_a], scores[rows_b]
    if not np.array_equal(sa, sb):
        return False
    boundary = sa[-1] if len(sa) else None
    for s in np.unique(sa):
        if s == boundary:
            continue
        if set(rows_a[sa == s]) != set(rows_b[sb == s]):
            return False
    return True
