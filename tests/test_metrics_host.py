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


def test_fm_grid_decomposition_equals_predict_on_cartesian_rows():
    """SURVEY.md Appendix A.4, the bridge from FM rows to the scoring GEMM: for rows x = [user features | item
    features], sigma(w0 + alpha_u + beta_i + <A_u, C_i>) with (A, C, alpha, beta, w0) = fm_factors(model, ...)
    equals FM predict (src/fm.py:114-133, here the pinned oracle) on the Cartesian-product rows. Host NumPy only."""
    import scipy.sparse as sp
    from oracle import fm_oracle
    from rfm_b200.fm import FactorizationMachines
    from rfm_b200.score import fm_factors
    rng = np.random.default_rng(11)
    n_users, n_items, n_uf, n_if, k = 13, 17, 20, 31, 8
    n = n_uf + n_if

    def side(rows, lo, width):          # a few real-valued entries per entity inside its side's column range
        cols = np.stack([rng.choice(width, size=4, replace=False) + lo for _ in range(rows)])
        vals = rng.normal(size=cols.shape)
        ptr_ = np.arange(0, rows * 4 + 1, 4)
        return sp.csr_matrix((vals.ravel(), cols.ravel(), ptr_), shape=(rows, n))

    user_table, item_table = side(n_users, 0, n_uf), side(n_items, n_uf, n_if)
    model = FactorizationMachines("IPS", 1, k, 0.1, 4, 12345, n)
    model.w0.params[:] = 0.3                                   # a non-zero global bias as well
    A, C, alpha, beta, bias = fm_factors(model, user_table, item_table)
    logits = bias + alpha[:, None] + beta[None, :] + A @ C.T
    u, i = np.divmod(np.arange(n_users * n_items), n_items)
    X = (user_table[u] + item_table[i]).tocsr()
    want = fm_oracle.fm_predict(X, model.w0(), model.w(), model.V()).reshape(n_users, n_items)
    np.testing.assert_allclose(1.0 / (1.0 + np.exp(-logits)), want, rtol=1e-12, atol=1e-15)


def test_group_by_user_data_matches_the_pandas_group_by():
    """TestEvaluator/ValEvaluator._group_by_user_data return what the reference's pandas expression returns
    (utils/evaluate.py:141-156, 223-239): users ascending, rows in frame order, the four keys, y_score recorded."""
    import pandas as pd
    from rfm_b200.evaluate import TestEvaluator, ValEvaluator
    rng = np.random.default_rng(2)
    n = 300
    df = pd.DataFrame({"user": rng.integers(0, 17, n), "item": rng.integers(0, 40, n), "label": rng.integers(0, 2, n),
                       "pscore": rng.uniform(0.1, 1, n), "ones_pscore": np.ones(n)})
    scores = rng.random(n)
    ref = df.assign(y_score=scores).groupby("user").agg(list).map(np.array)
    te = TestEvaluator(interaction_df=df.copy(), features={}, K=[1, 3], used_metrics={"DCG"}, n_items=40)
    ve = ValEvaluator(interaction_df=df.copy(), features={}, k=3, metric_name="DCG")
    for got, ps_name in ((te._group_by_user_data(scores), "pscore"), (ve._group_by_user_data(scores, "IPS"), "pscore"),
                         (ve._group_by_user_data(scores, "Naive"), "ones_pscore")):
        assert list(got) == list(ref.index)
        for user, data in got.items():
            assert set(data) == {"items", "labels", "y_scores", "pscores"}
            np.testing.assert_array_equal(data["items"], ref.loc[user, "item"])
            np.testing.assert_array_equal(data["labels"], ref.loc[user, "label"])
            np.testing.assert_array_equal(data["y_scores"], ref.loc[user, "y_score"])
            np.testing.assert_array_equal(data["pscores"], ref.loc[user, ps_name])
    np.testing.assert_array_equal(te.interaction_df["y_score"].to_numpy(), scores)


def test_fractional_labels_are_not_truncated():
    """The reference divides whatever `labels` holds (src/fm.py:80); the shim sends integer labels to the device and
    falls back to host-computed float64 targets when a label has a fractional part."""
    from rfm_b200 import _capi
    ps = np.array([0.5, 0.25, 1.0])
    y, t = _capi.integer_labels(np.array([1, 0, 1]), ps)
    assert y.dtype == np.int64 and t is None
    y, t = _capi.integer_labels(np.array([1.0, 0.0, 2.0]), ps)
    assert y.tolist() == [1, 0, 2] and t is None
    y, t = _capi.integer_labels(np.array([0.5, 0.0, 1.0]), ps)
    np.testing.assert_array_equal(t, np.array([0.5, 0.0, 1.0]) / ps)
