"""GPU parity tests of full-catalog scoring + top-K (tcgen05 GEMM prune, exact float64 result).

Bar: top-K item ids bit-exact against float64 NumPy ranking of the same scores with the canonical
tie rule; for FM, against the oracle's predict on the Cartesian-product rows."""
import numpy as np
import pytest

from oracle import fm_oracle

pytestmark = pytest.mark.gpu


def numpy_topk(A, C, alpha, beta, bias, K):
    S = ((A @ C.T + (0 if alpha is None else alpha[:, None])) + (0 if beta is None else beta[None, :])) + bias
    order = np.argsort(S, axis=1, kind="stable")[:, ::-1][:, :K]     # score desc, larger item id first on ties
    return order.astype(np.int32), np.take_along_axis(S, order, axis=1)


def check(A, C, alpha, beta, bias, K, mode="tensor", expect_tensor=True):
    from rfm_b200.score import TopKScorer
    sc = TopKScorer(A, C, alpha, beta, bias)
    items, scores = sc.topk(K, mode=mode)
    ref_items, ref_scores = numpy_topk(A, C, alpha, beta, bias, K)
    np.testing.assert_allclose(scores, ref_scores, rtol=1e-12, atol=1e-12)
    # ids must agree wherever the scores are not within rounding of a neighbour's
    same = items == ref_items
    if not same.all():
        gap = np.abs(np.diff(ref_scores, axis=1))
        near = np.zeros_like(same)
        near[:, :-1] |= gap < 1e-12
        near[:, 1:] |= gap < 1e-12
        assert (same | near).all(), "top-K ids differ beyond rounding-level ties"
    if mode == "tensor":
        assert sc.last_stats["tensor_core_path"] == expect_tensor
    return sc


@pytest.mark.parametrize("U,I,k,K", [(1411, 3327, 64, 9), (300, 1000, 16, 5), (129, 257, 64, 9), (128, 256, 128, 9),
                                     (50, 40, 33, 9), (2000, 5000, 64, 100), (700, 900, 100, 20)])
def test_mf_style_topk_matches_numpy(U, I, k, K):
    rng = np.random.default_rng(U + I + k)
    A = rng.normal(size=(U, k)) * 0.5
    C = rng.normal(size=(I, k)) * 0.5
    alpha = rng.normal(size=U) * 0.1
    beta = rng.normal(size=I) * 0.3
    sc = check(A, C, alpha, beta, 0.25, K)
    assert sc.last_stats["users_ranked_exactly"] <= U // 10      # pruning is proven for nearly every user
    check(A, C, alpha, beta, 0.25, K, mode="exact")


def test_item_terms_alone_rank_the_catalog():
    """A = 0: the approximate score is the item term added by the extra MMA step, so a wrong operand layout or
    split of beta shows up directly (the candidate sets would miss the true top items)."""
    rng = np.random.default_rng(4)
    U, I, k = 300, 5000, 64
    beta = rng.normal(size=I) * 3.0
    sc = check(np.zeros((U, k)), rng.normal(size=(I, k)), None, beta, 0.0, 9)
    assert sc.last_stats["users_ranked_exactly"] == 0
    assert sc.last_stats["candidates"] <= 40 * U            # thresholds are tight: a handful of items per user
    # a large common offset plus tiny differences: only the low-order parts of the split separate the items
    beta = 1000.0 + rng.normal(size=I) * 1e-3
    check(rng.normal(size=(U, k)) * 1e-4, rng.normal(size=(I, k)) * 1e-4, None, beta, 0.0, 9)


def test_user_chunking_and_extreme_k(monkeypatch):
    """Users are processed in chunks when the group maxima would exceed the scratch budget; K = 1 and the
    largest supported K; a catalog smaller than K (padding with -1 / -inf)."""
    rng = np.random.default_rng(9)
    U, I, k = 1500, 9000, 64
    A, C = rng.normal(size=(U, k)) * 0.4, rng.normal(size=(I, k)) * 0.4
    beta = rng.normal(size=I) * 0.2
    monkeypatch.setenv("RFM_SCORE_SCRATCH_MB", "1")            # 1 MiB: a few 128-user blocks per chunk
    sc = check(A, C, None, beta, 0.0, 9)
    assert sc.last_stats["users_ranked_exactly"] == 0
    check(A, C, None, beta, 0.0, 1)
    check(A[:300], C, None, beta, 0.0, 120)
    monkeypatch.delenv("RFM_SCORE_SCRATCH_MB")
    from rfm_b200.score import TopKScorer
    items, scores = TopKScorer(A[:10], C[:5], None, beta[:5], 0.0).topk(9)
    assert (items[:, 5:] == -1).all() and np.isneginf(scores[:, 5:]).all()
    ref_items, ref_scores = numpy_topk(A[:10], C[:5], None, beta[:5], 0.0, 5)
    np.testing.assert_array_equal(items[:, :5], ref_items)


def test_result_views_equal_copies():
    from rfm_b200.score import TopKScorer
    rng = np.random.default_rng(13)
    sc = TopKScorer(rng.normal(size=(200, 64)), rng.normal(size=(1500, 64)), None, rng.normal(size=1500), 0.0)
    items, scores = sc.topk(9)
    vi, vs = sc.topk(9, copy=False)
    np.testing.assert_array_equal(vi, items)
    np.testing.assert_array_equal(vs, scores)
    assert not vi.flags.writeable and not vs.flags.writeable


def test_k_wider_than_the_tensor_path_uses_exact_kernel():
    rng = np.random.default_rng(0)
    check(rng.normal(size=(40, 300)), rng.normal(size=(90, 300)), None, None, 0.0, 7, expect_tensor=False)


def test_near_ties_fall_back_to_exact_ranking():
    """Items that differ by less than bf16 resolution: the proof fails and the users are ranked exactly."""
    rng = np.random.default_rng(1)
    U, I, k, K = 200, 2000, 64, 9
    A = rng.normal(size=(U, k))
    base = rng.normal(size=(1, k))
    C = base + rng.normal(size=(I, k)) * 1e-5
    sc = check(A, C, None, None, 0.0, K)
    assert sc.last_stats["users_ranked_exactly"] > 0


def test_exact_ties_follow_the_canonical_rule():
    rng = np.random.default_rng(2)
    A = rng.integers(-2, 3, size=(64, 8)).astype(float)
    C = rng.integers(-2, 3, size=(500, 8)).astype(float)       # integer scores: many exact ties
    from rfm_b200.score import TopKScorer
    for mode in ("tensor", "exact"):
        items, scores = TopKScorer(A, C).topk(9, mode=mode)
        ref_items, ref_scores = numpy_topk(A, C, None, None, 0.0, 9)
        np.testing.assert_array_equal(items, ref_items)
        np.testing.assert_array_equal(scores, ref_scores)


def test_fm_decomposition_equals_predict_on_cartesian_rows():
    """FM full-grid ranking from the per-side decomposition == oracle predict on every (user, item) row."""
    from rfm_b200.fm import FactorizationMachines
    from rfm_b200.score import TopKScorer, fm_factors
    from rfm_b200.synth import make_coat_shaped, csr_from_tables
    from scipy.sparse import csr_matrix
    log = make_coat_shaped(seed=3, n_users=70, n_items=90, n_rated=10, n_test=6)
    m = FactorizationMachines("IPS", 3, 32, 1e-3, 200, 12345, log.n_features, alpha=0.3)
    m.fit(log.fm_train, log.fm_val)
    t = log.tables
    user_table = csr_matrix((t["u_val"], t["u_col"], t["u_ptr"]), shape=(log.n_users, log.n_features))
    item_table = csr_matrix((t["i_val"], t["i_col"], t["i_ptr"]), shape=(log.n_items, log.n_features))
    A, C, alpha, beta, bias = fm_factors(m, user_table, item_table)
    items, scores = TopKScorer(A, C, alpha, beta, bias).topk(9)
    uu, ii = np.meshgrid(np.arange(log.n_users), np.arange(log.n_items), indexing="ij")
    grid = csr_from_tables(uu.ravel(), ii.ravel(), t)
    logits = fm_oracle.fm_logits(grid, float(m.w0()[0]), m.w(), m.V()).reshape(log.n_users, log.n_items)
    order = np.argsort(logits, axis=1, kind="stable")[:, ::-1][:, :9]
    np.testing.assert_array_equal(items, order.astype(np.int32))
    np.testing.assert_allclose(scores, np.take_along_axis(logits, order, axis=1), rtol=1e-9, atol=1e-12)
    # and predict() of those rows is the sigmoid of the same numbers (src/fm.py:131)
    np.testing.assert_allclose(fm_oracle.sigmoid(scores[:, 0]),
                               m.predict(X=grid).reshape(log.n_users, log.n_items)[np.arange(log.n_users), items[:, 0]],
                               rtol=1e-9)


def test_item_sharded_merge_equals_single_pass():
    from rfm_b200.score import TopKScorer, merge_topk
    rng = np.random.default_rng(5)
    A, C = rng.normal(size=(300, 64)), rng.normal(size=(1200, 64))
    beta = rng.normal(size=1200)
    sc = TopKScorer(A, C, None, beta, 0.0)
    full_items, full_scores = sc.topk(9)
    parts = [sc.topk(9, item_range=(b, e)) for b, e in ((0, 300), (300, 900), (900, 1200))]
    items, scores = merge_topk([p[0] for p in parts], [p[1] for p in parts], 9)
    np.testing.assert_array_equal(items, full_items)
    np.testing.assert_array_equal(scores, full_scores)


def test_full_catalog_evaluator_equals_test_evaluator_on_cartesian_frame():
    """Metrics over the whole catalog == the oracle's TestEvaluator semantics on every (user, item) row."""
    from oracle import metrics_oracle
    from rfm_b200.evaluate import FullCatalogEvaluator
    from rfm_b200.score import TopKScorer
    rng = np.random.default_rng(8)
    U, I, k = 90, 400, 64
    A, C = rng.normal(size=(U, k)) * 0.4, rng.normal(size=(I, k)) * 0.4
    beta = rng.normal(size=I) * 0.2
    # held-out rows: ~12 per user, some users without any positive
    hu = np.repeat(np.arange(U), 12)
    hi = np.concatenate([rng.choice(I, 12, replace=False) for _ in range(U)])
    hl = (rng.random(hu.size) < 0.4).astype(np.int64)
    hl[hu % 9 == 0] = 0
    theta = rng.uniform(0.1, 1.0, size=I)
    K = [1, 3, 5, 9]
    used = {"DCG", "CatalogCoverage", "Recall", "MAP", "Gini"}
    ev = FullCatalogEvaluator({"user": hu, "item": hi, "label": hl}, theta, K, used, U, I)
    res = ev.evaluate(TopKScorer(A, C, None, beta, 0.1))
    # oracle on the Cartesian frame
    uu, ii = np.meshgrid(np.arange(U), np.arange(I), indexing="ij")
    lab = np.zeros((U, I), dtype=np.int64)
    lab[hu, hi] = hl
    frame = {"user": uu.ravel(), "item": ii.ravel(), "label": lab.ravel(), "pscore": theta[ii.ravel()],
             "ones_pscore": np.ones(U * I)}
    S = ((A @ C.T) + beta[None, :]) + 0.1
    ref = metrics_oracle.test_evaluate(frame, S.ravel(), K, used, I)
    for m in ["ME"] + sorted(used):
        np.testing.assert_allclose(res[m], ref[m], rtol=1e-12, err_msg=m)
    np.testing.assert_array_equal(np.array(res["CatalogCoverage"]), np.array(ref["CatalogCoverage"]))


def test_full_catalog_evaluator_large_k_and_short_catalog():
    """K up to 100 (the stress configuration's top-100), a catalog shorter than max(K) (ME@k is nan there), users
    without any held-out positive, duplicate held-out pairs (summed like scipy's csr_matrix)."""
    from oracle import metrics_oracle
    from rfm_b200.evaluate import FullCatalogEvaluator
    from rfm_b200.score import TopKScorer
    rng = np.random.default_rng(18)
    for U, I, K in ((70, 600, [1, 10, 50, 100]), (40, 30, [1, 9, 50])):
        k = 32
        A, C = rng.normal(size=(U, k)) * 0.4, rng.normal(size=(I, k)) * 0.4
        beta = rng.normal(size=I) * 0.2
        hu = np.repeat(np.arange(U), 8)
        hi = rng.integers(0, I, hu.size)                         # duplicates happen
        hl = (rng.random(hu.size) < 0.5).astype(np.int64)
        hl[hu % 5 == 0] = 0
        theta = rng.uniform(0.1, 1.0, size=I)
        used = {"DCG", "CatalogCoverage", "Recall", "MAP", "Gini"}
        ev = FullCatalogEvaluator({"user": hu, "item": hi, "label": hl}, theta, K, used, U, I)
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            res = ev.evaluate(TopKScorer(A, C, None, beta, 0.0))
        uu, ii = np.meshgrid(np.arange(U), np.arange(I), indexing="ij")
        lab = np.zeros((U, I))
        np.add.at(lab, (hu, hi), hl)
        frame = {"user": uu.ravel(), "item": ii.ravel(), "label": lab.ravel(), "pscore": theta[ii.ravel()],
                 "ones_pscore": np.ones(U * I)}
        S = (A @ C.T) + beta[None, :]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ref = metrics_oracle.test_evaluate(frame, S.ravel(), K, used, I)
        for m in ["ME"] + sorted(used):
            np.testing.assert_allclose(res[m], ref[m], rtol=1e-12, equal_nan=True, err_msg="%s U=%d I=%d" % (m, U, I))
