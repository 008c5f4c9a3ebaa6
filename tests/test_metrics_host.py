"""Host metric functions (API fidelity with utils/metrics.py) against the oracle."""
import numpy as np

from oracle import metrics_oracle
from rfm_b200 import metrics


def test_per_user_functions_match_oracle():
    rng = np.random.default_rng(0)
    for n in (1, 2, 5, 9, 30):
        for _ in range(20):
            y = rng.integers(0, 2, n).astype(float)
            ps = rng.uniform(0.1, 1.0, n)
            for k in (1, 3, 5, 9):
                np.testing.assert_allclose(metrics.calc_dcg_at_k(y, k), metrics_oracle.dcg_at_k(y, k), equal_nan=True)
                np.testing.assert_allclose(metrics.calc_ips_of_dcg_at_k(y, k, ps),
                                           metrics_oracle.ips_dcg_at_k(y, k, ps), equal_nan=True)
                np.testing.assert_allclose(metrics.return_exposure_at_k(ps, k), metrics_oracle.exposure_at_k(ps, k),
                                           equal_nan=True)
                np.testing.assert_allclose(metrics.calc_recall_at_k(y, k), metrics_oracle.recall_at_k(y, k))
                np.testing.assert_allclose(metrics.calc_average_precision_at_k(y, k), metrics_oracle.ap_at_k(y, k))
    items = rng.integers(0, 50, 400).tolist()
    assert metrics.calc_catalog_coverage_at_k(items, 50) == metrics_oracle.coverage(items, 50)
    np.testing.assert_allclose(metrics.calc_gini_at_k(items, 50), metrics_oracle.gini(items, 50))
    assert set(metrics.metric_candidates) == {"Recall", "MAP", "DCG", "ME", "CatalogCoverage", "Gini"}


def test_optimizer_holder_contract():
    from rfm_b200.optimizer import SGD
    p = np.arange(6, dtype=float).reshape(3, 2)
    h = SGD(params=p, lr=0.5)
    assert h() is p and h(1)[0] == 2.0
    h.update(grad=np.ones((3, 2)), index=None)
    np.testing.assert_array_equal(p, np.arange(6).reshape(3, 2) - 0.5)
    h.update(grad=np.array([2.0, 4.0]), index=(np.array([0, 2]), 1))
    assert p[0, 1] == 0.5 - 1.0 and p[2, 1] == 4.5 - 2.0
    assert h.version == 2


def test_merge_topk_follows_the_canonical_order():
    """Host merge of per-shard top-K lists (item-sharded scoring): score descending, larger item id first among
    exact ties, -1 / -inf padding last."""
    from rfm_b200.score import merge_topk
    rng = np.random.default_rng(3)
    U, K, n_items = 40, 5, 60
    scores_full = rng.integers(0, 6, size=(U, n_items)).astype(float)         # many exact ties
    order = np.argsort(scores_full, axis=1, kind="stable")[:, ::-1]
    ref_items, ref_scores = order[:, :K], np.take_along_axis(scores_full, order[:, :K], axis=1)
    parts_i, parts_s = [], []
    for b, e in ((0, 17), (17, 20), (20, 60)):                                # one shard holds fewer than K items
        o = np.argsort(scores_full[:, b:e], axis=1, kind="stable")[:, ::-1][:, :K]
        it = (o + b).astype(np.int32)
        sc = np.take_along_axis(scores_full[:, b:e], o, axis=1)
        if it.shape[1] < K:
            it = np.concatenate([it, np.full((U, K - it.shape[1]), -1, np.int32)], axis=1)
            sc = np.concatenate([sc, np.full((U, K - sc.shape[1]), -np.inf)], axis=1)
        parts_i.append(it)
        parts_s.append(sc)
    items, scores = merge_topk(parts_i, parts_s, K)
    np.testing.assert_array_equal(items, ref_items)
    np.testing.assert_array_equal(scores, ref_scores)
