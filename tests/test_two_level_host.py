"""The algebra of the two-level FM step (csrc/two_level.cuh) on the CPU: per-entity aggregates forward and per-entity
sums backward give the predictions and the batch gradient of the reference arithmetic (fm_oracle, pinned to the goldens
of the unmodified reference) on the stacked matrix."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import fm_oracle, two_level_oracle


def _problem(seed, n_ctx, order):
    from rfm_b200.factored import FactoredFeatures
    rng = np.random.default_rng(seed)
    n_users, n_items, n = 37, 29, 900
    ut = sp.random(n_users, 11, density=0.35, format="csr", random_state=seed)
    ut = sp.vstack([ut[:30], sp.csr_matrix((7, 11))]).tocsr()             # entities without side features
    it = sp.random(n_items, 6, density=0.5, format="csr", random_state=seed + 1)
    users, items = rng.integers(0, n_users, n), rng.integers(0, n_items, n)
    named = {"idu": ("id", "user", n_users), "idi": ("id", "item", n_items), "tu": ("table", "user", ut),
             "ti": ("table", "item", it)}
    blocks = [named[k] for k in order]
    if n_ctx:
        ctx = rng.normal(size=(n, n_ctx))
        ctx[rng.random((n, n_ctx)) < 0.25] = 0.0                          # exact zeros are no entries
        blocks.insert(2, ("ctx", ctx))
    ff = FactoredFeatures(blocks, users, items)
    y = (rng.random(n) < 0.45).astype(np.float64)
    ps = rng.uniform(0.2, 1.0, n)
    return ff, y, ps, rng


@pytest.mark.parametrize("n_ctx", [0, 1, 3])
@pytest.mark.parametrize("order", [("idu", "idi", "tu", "ti"), ("tu", "idi", "ti", "idu")])
def test_two_level_gradient_and_predictions_equal_the_flat_arithmetic(n_ctx, order):
    ff, y, ps, rng = _problem(5 + n_ctx, n_ctx, order)
    X = ff.tocsr()
    n, k = X.shape[1], 12
    w0, w, V = np.array([0.3]), rng.normal(size=n) * 0.2, rng.normal(size=(n, k)) * 0.2
    rows = rng.choice(X.shape[0], 400, replace=False)
    g0, gw, gV = fm_oracle.fm_grad(X[rows], y[rows], ps[rows], w0, w, V)
    h0, hw, hV = two_level_oracle.two_level_grad(ff, rows, y[rows], ps[rows], w0, w, V)
    scale = np.abs(gV).max()
    assert scale > 0
    np.testing.assert_allclose(h0, g0, rtol=1e-12)
    np.testing.assert_allclose(hw, gw, rtol=1e-11, atol=1e-13 * np.abs(gw).max())
    np.testing.assert_allclose(hV, gV, rtol=1e-10, atol=1e-13 * scale)
    E, n_users, n_items, nc, _ = two_level_oracle.entity_lists(ff)
    Xv = two_level_oracle.virtual_rows(ff, n_users, n_items, nc, rows)
    p, _ = two_level_oracle.two_level_predict(Xv, w0, *two_level_oracle.entity_forward(E, w, V))
    np.testing.assert_allclose(p, fm_oracle.fm_predict(X[rows], w0, w, V), rtol=1e-12)
    # the stacked matrix IS virtual rows times entity lists
    assert abs(Xv.dot(E) - X[rows]).max() < 1e-15
