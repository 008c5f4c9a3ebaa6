"""GPU parity tests of the ranking evaluators (utils/evaluate.py, utils/metrics.py).

Bar: top-K row ids and coverage counts bit-exact; DCG / ME / Recall / MAP within 1e-5 relative
(held to 1e-12 here)."""
import numpy as np
import pytest

from conftest import load_golden, golden_frame
from oracle import metrics_oracle

pytestmark = pytest.mark.gpu
USED = {"DCG", "CatalogCoverage", "Recall", "MAP", "Gini"}


@pytest.mark.parametrize("name", ["coat_eval_tiefree", "coat_eval_mf", "coat_eval_ragged"])
def test_test_evaluator_matches_reference_golden(name):
    from rfm_b200.evaluate import TestEvaluator, ValEvaluator
    g = load_golden(name)
    frame, scores = golden_frame(g), g["scores"]
    K = [int(k) for k in g["K"]]
    te = TestEvaluator(interaction_df=frame, features={}, K=K, used_metrics=USED, n_items=int(g["n_items"]))
    with pytest.warns(RuntimeWarning) if np.isnan(g["test_ME"]).any() else _nullcontext():
        res = te.evaluate(scores)
    for m in ["ME"] + sorted(USED):
        np.testing.assert_allclose(res[m], g["test_" + m], rtol=1e-12, equal_nan=True, err_msg=m)
    np.testing.assert_array_equal(np.array(res["CatalogCoverage"]), g["test_CatalogCoverage"])   # exact
    # top-9 rows per user, bit-exact with the reference's own argsort()[::-1] on tie-free scores
    top = te.top_rows(scores, 9)
    np.testing.assert_array_equal(top, g["top_rows"])
    for est in ("IPS", "Naive"):
        for k in (3, 5):
            ve = ValEvaluator(interaction_df=frame, features={}, k=k, metric_name="DCG")
            np.testing.assert_allclose(ve.evaluate(scores, est), g["val_%s_%d" % (est, k)], rtol=1e-12)


class _nullcontext:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


def test_tie_heavy_scores_follow_canonical_rule():
    """Saturated sigmoids give exact duplicates; the device ranking must equal
    argsort(kind='stable')[::-1] and therefore the oracle, exactly."""
    from rfm_b200.evaluate import TestEvaluator
    rng = np.random.default_rng(11)
    n_users, per_user, n_items = 200, 37, 150
    users = np.repeat(np.arange(n_users), per_user)
    rng.shuffle(users)
    frame = {"user": users, "item": rng.integers(0, n_items, users.size),
             "label": rng.integers(0, 2, users.size), "pscore": rng.uniform(0.1, 1, users.size),
             "ones_pscore": np.ones(users.size)}
    scores = rng.integers(0, 6, users.size) / 5.0            # only 6 distinct values
    te = TestEvaluator(interaction_df=frame, features={}, K=[1, 5, 9], used_metrics=USED, n_items=n_items)
    res = te.evaluate(scores)
    ref = metrics_oracle.test_evaluate(frame, scores, [1, 5, 9], USED, n_items)
    for m in ["ME"] + sorted(USED):
        np.testing.assert_allclose(res[m], ref[m], rtol=1e-12, err_msg=m)
    top = te.top_rows(scores, 9)
    for row, (user, rows) in zip(top, metrics_oracle.ranked_lists(frame, scores)):
        np.testing.assert_array_equal(row, rows[:9])


def test_long_lists_and_large_k():
    """Full-grid style: every user ranks a few thousand candidates; K up to 100."""
    from rfm_b200.evaluate import TestEvaluator
    rng = np.random.default_rng(3)
    n_users, n_items = 40, 3327
    users = np.repeat(np.arange(n_users), n_items)
    items = np.tile(np.arange(n_items), n_users)
    frame = {"user": users, "item": items, "label": (rng.random(users.size) < 0.02).astype(np.int64),
             "pscore": rng.uniform(0.1, 1, users.size), "ones_pscore": np.ones(users.size)}
    scores = rng.random(users.size)
    K = [1, 10, 100]
    te = TestEvaluator(interaction_df=frame, features={}, K=K, used_metrics={"DCG", "CatalogCoverage", "Recall"},
                       n_items=n_items)
    res = te.evaluate(scores)
    ref = metrics_oracle.test_evaluate(frame, scores, K, {"DCG", "CatalogCoverage", "Recall"}, n_items)
    for m in res:
        np.testing.assert_allclose(res[m], ref[m], rtol=1e-12, err_msg=m)


def test_error_paths_match_reference():
    from rfm_b200.evaluate import TestEvaluator, ValEvaluator
    with pytest.raises(ValueError, match="metric_name must be in"):
        TestEvaluator(interaction_df={}, features={}, K=[1], used_metrics={"NDCG"}, n_items=3)
    with pytest.raises(ValueError, match="only DCG"):
        ValEvaluator(interaction_df={}, features={}, k=3, metric_name="Recall")


def test_pandas_frame_is_accepted_and_mutated_like_the_reference():
    import pandas as pd
    from rfm_b200.evaluate import ValEvaluator
    g = load_golden("coat_eval_tiefree")
    df = pd.DataFrame(golden_frame(g))
    ve = ValEvaluator(interaction_df=df, features={}, k=5, metric_name="DCG")
    np.testing.assert_allclose(ve.evaluate(g["scores"], "IPS"), g["val_IPS_5"], rtol=1e-12)
    np.testing.assert_array_equal(df["y_score"].to_numpy(), g["scores"])


def test_nan_and_signed_zero_scores_rank_like_numpy():
    """A diverged fit hands over NaN scores. NumPy's argsort puts NaN behind every number, so the reference's
    argsort()[::-1] ranks NaN FIRST; -0.0 ties with +0.0. The device order is total and equal to the oracle's."""
    from rfm_b200.evaluate import TestEvaluator
    rng = np.random.default_rng(5)
    n_users, per_user, n_items = 60, 23, 90
    users = np.repeat(np.arange(n_users), per_user)
    frame = {"user": users, "item": rng.integers(0, n_items, users.size),
             "label": rng.integers(0, 2, users.size), "pscore": rng.uniform(0.1, 1, users.size),
             "ones_pscore": np.ones(users.size)}
    scores = rng.normal(size=users.size)
    scores[rng.random(users.size) < 0.2] = np.nan
    scores[rng.random(users.size) < 0.1] = 0.0
    scores[rng.random(users.size) < 0.1] = -0.0
    scores[rng.random(users.size) < 0.05] = np.inf
    scores[rng.random(users.size) < 0.05] = -np.inf
    scores[:per_user] = np.nan                                   # a user whose every score is NaN
    te = TestEvaluator(interaction_df=frame, features={}, K=[1, 5, 9], used_metrics=USED, n_items=n_items)
    res = te.evaluate(scores)
    ref = metrics_oracle.test_evaluate(frame, scores, [1, 5, 9], USED, n_items)
    for m in ["ME"] + sorted(USED):
        np.testing.assert_allclose(res[m], ref[m], rtol=1e-12, err_msg=m)
    top = te.top_rows(scores, 9)
    for row, (user, rows) in zip(top, metrics_oracle.ranked_lists(frame, scores)):
        np.testing.assert_array_equal(row, rows[:9])


def test_long_tie_heavy_lists_use_the_selection_path():
    """Lists longer than the shared-memory sort (2,048 rows) with a handful of distinct scores: the two-level
    bisection (score key, then row position among ties) must pick exactly the canonical top-K."""
    from rfm_b200.evaluate import TestEvaluator
    rng = np.random.default_rng(8)
    n_users, per_user, n_items = 12, 5000, 6000
    users = np.repeat(np.arange(n_users), per_user)
    frame = {"user": users, "item": np.concatenate([rng.permutation(n_items)[:per_user] for _ in range(n_users)]),
             "label": (rng.random(users.size) < 0.01).astype(np.int64), "pscore": rng.uniform(0.1, 1, users.size),
             "ones_pscore": np.ones(users.size)}
    scores = rng.integers(0, 4, users.size) / 3.0
    scores[per_user: 2 * per_user] = 0.5                          # one user: every score equal
    K = [1, 10, 100, 128]
    te = TestEvaluator(interaction_df=frame, features={}, K=K, used_metrics=USED, n_items=n_items)
    res = te.evaluate(scores)
    ref = metrics_oracle.test_evaluate(frame, scores, K, USED, n_items)
    for m in ["ME"] + sorted(USED):
        np.testing.assert_allclose(res[m], ref[m], rtol=1e-12, err_msg=m)
    top = te.top_rows(scores, 128)
    for row, (user, rows) in zip(top, metrics_oracle.ranked_lists(frame, scores)):
        np.testing.assert_array_equal(row, rows[:128])


def test_replaced_frame_columns_rebuild_the_device_ranker():
    """The reference regroups interaction_df on every call; the device copy must follow a replaced column."""
    from rfm_b200.evaluate import ValEvaluator
    g = load_golden("coat_eval_tiefree")
    frame = dict(golden_frame(g))
    ve = ValEvaluator(interaction_df=frame, features={}, k=5, metric_name="DCG")
    first = ve.evaluate(g["scores"], "IPS")
    np.testing.assert_allclose(first, g["val_IPS_5"], rtol=1e-12)
    frame["label"] = 1 - frame["label"]                           # same length, other content
    ref = metrics_oracle.val_evaluate(frame, g["scores"], 5, "IPS")
    np.testing.assert_allclose(ve.evaluate(g["scores"], "IPS"), ref, rtol=1e-12)
