"""GPU parity tests of the MF path (src/mf.py) against the reference goldens and the oracle."""
import numpy as np
import pytest

from conftest import load_golden, golden_frame
from oracle import mf_oracle

pytestmark = pytest.mark.gpu


def _model(g, **kw):
    from rfm_b200.mf import LogisticMatrixFactorization
    return LogisticMatrixFactorization(estimator="IPS", n_epochs=int(g["n_epochs"]), n_factors=int(g["k"]),
                                       lr=float(g["lr"]), batch_size=int(g["B"]), seed=int(g["seed"]),
                                       n_users=int(g["n_users"]), n_items=int(g["n_items"]), reg=float(g["reg"]),
                                       alpha=float(g["alpha"]), **kw)


def _dicts(g):
    train = {"features": g["train_pairs"], "labels": g["train_labels"], "pscores": g["train_pscores"]}
    val = {"features": g["val_pairs"], "labels": g["val_labels"], "pscores": g["val_pscores"]}
    return train, val


@pytest.mark.parametrize("name", ["coat_mf_ips", "coat_mf_ips_alpha01"])
def test_fit_trajectory_matches_reference_golden(name):
    g = load_golden(name)
    train, val = _dicts(g)
    m = _model(g)
    np.testing.assert_array_equal(m.P(), g["P_init"])
    np.testing.assert_array_equal(m.b_i(), g["bi_init"])
    tl, vl = m.fit(train, val)
    assert m.b == float(g["b"])
    np.testing.assert_allclose(tl, g["train_loss"], rtol=1e-9)
    np.testing.assert_allclose(vl, g["val_loss"], rtol=1e-9)
    for mine, ref in ((m.P(), g["P"]), (m.Q(), g["Q"]), (m.b_u(), g["b_u"]), (m.b_i(), g["b_i"])):
        np.testing.assert_allclose(mine, ref, rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(m.predict(g["test_pairs"]), g["test_scores"], rtol=1e-9)


def test_val_metrics_every_epoch_match_reference():
    """evaluator hook inside fit (src/mf.py:126-132) with the device ranker."""
    from rfm_b200.evaluate import ValEvaluator
    g = load_golden("coat_mf_ips_alpha01")
    train, val = _dicts(g)
    ev = ValEvaluator(interaction_df=golden_frame(g), features={"MF": g["test_pairs"]}, k=5, metric_name="DCG")
    m = _model(g, evaluator=ev)
    m.fit(train, val)
    assert m.model_name == "MF"
    np.testing.assert_allclose(m.val_metrics, g["val_metrics"], rtol=1e-9)


def test_heavy_repeats_long_dependency_chains():
    """Few users/items and a big batch: hundreds of wavefront levels; equality with the strictly
    sequential oracle shows the schedule preserves the reference's update order."""
    rng = np.random.default_rng(5)
    U, I, N, B, k = 7, 5, 4000, 1500, 24
    pairs = np.stack([rng.integers(0, U, N), rng.integers(0, I, N)], axis=1).astype(np.int64)
    y = rng.integers(0, 2, N)
    ps = rng.uniform(0.3, 1.0, N)
    train = {"features": pairs, "labels": y, "pscores": ps}
    from rfm_b200.mf import LogisticMatrixFactorization
    m = LogisticMatrixFactorization("IPS", 3, k, 0.01, B, 9, U, I, 0.1, alpha=0.5)
    P, Q, bu, bi = m.P().copy(), m.Q().copy(), m.b_u().copy(), m.b_i().copy()
    tl, vl = m.fit(train, train)
    (P, Q, bu, bi, b), rtl, rvl = mf_oracle.mf_fit(train, train, 3, B, 0.01, 0.1, P, Q, bu, bi,
                                                   epoch_fn=mf_oracle.mf_epoch_py)
    np.testing.assert_allclose(tl, rtl, rtol=1e-9)
    np.testing.assert_allclose(vl, rvl, rtol=1e-9)
    np.testing.assert_allclose(m.P(), P, rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(m.Q(), Q, rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(m.b_u(), bu, rtol=1e-9, atol=1e-13)


def test_float32_mode_and_bad_ids():
    g = load_golden("coat_mf_ips_alpha01")
    train, val = _dicts(g)
    m = _model(g, dtype="float32")
    tl, vl = m.fit(train, val)
    np.testing.assert_allclose(tl, g["train_loss"], rtol=2e-5)
    bad = g["test_pairs"].copy()
    bad[0, 1] = int(g["n_items"]) + 3
    with pytest.raises(ValueError, match="out of range"):
        m.predict(bad)
