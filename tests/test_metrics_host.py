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
