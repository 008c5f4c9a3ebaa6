"""GPU parity of the device-side click generator (csrc/synth.cu) against its specification
(oracle/clicks_oracle.py): ids and labels bit for bit, reals to 1e-12; shards agree with the whole; rows generated
on the device train exactly like the same rows uploaded from the host."""
import numpy as np
import pytest
import scipy.sparse as sp

from oracle import clicks_oracle as co

pytestmark = pytest.mark.gpu


def _blocks(model, rng):
    ut = sp.random(model.n_users, 6, density=0.4, format="csr", random_state=5)
    it = sp.random(model.n_items, 4, density=0.5, format="csr", random_state=6)
    return [("id", "user", model.n_users), ("id", "item", model.n_items), ("ctx", 2), ("table", "user", ut),
            ("table", "item", it)], ut, it


@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_generated_rows_match_the_specification(dtype):
    from rfm_b200 import _capi
    from rfm_b200.clicks import ClickModel, GeneratedRows
    model = ClickModel(3000, 5000, seed=21)
    blocks, _, _ = _blocks(model, None)
    ctx = _capi.Context.default(0)
    n, row0 = 200_000, 123_456_789_012
    rows = GeneratedRows(ctx, model, n, blocks, row0=row0, dtype=dtype, keep_labels=True)
    got = rows.download()
    ref = co.generate(model.seed, row0, n, model.user_cdf, model.item_cdf, model.item_exposure, model.pow_used,
                      n_ctx=2, n_hidden=model.n_hidden, hidden_scale=model.hidden_scale, noise_scale=model.noise_scale,
                      watch_shift=model.watch_shift, relevance_clip=model.relevance_clip)
    np.testing.assert_array_equal(got["users"], ref["users"])
    np.testing.assert_array_equal(got["items"], ref["items"])
    tol = 1e-12 if dtype == "float64" else 1e-6
    np.testing.assert_allclose(got["ctx"], ref["ctx"], rtol=tol, atol=tol)
    # a Bernoulli draw may only differ where its uniform sits within rounding of the probability
    near = np.abs(ref["u_relevance"] - ref["gamma"]) < 1e-12
    assert near.sum() <= 1
    np.testing.assert_array_equal(got["relevance"][~near], ref["relevance"][~near])
    np.testing.assert_array_equal(got["labels"][~near], ref["labels"][~near])
    np.testing.assert_allclose(got["targets"][~near], ref["targets"][~near], rtol=tol)
    # any shard of the log is the same bits
    part = GeneratedRows(ctx, model, 50_000, blocks, row0=row0 + 70_000, dtype=dtype, keep_labels=True).download()
    for key in ("users", "items", "ctx", "targets", "labels", "relevance"):
        np.testing.assert_array_equal(part[key], got[key][70_000:120_000])


def test_fit_on_generated_rows_equals_fit_on_the_uploaded_copy():
    from rfm_b200 import _capi
    from rfm_b200.clicks import ClickModel, GeneratedRows
    from rfm_b200.factored import FactoredFeatures
    from rfm_b200.fm import FactorizationMachines
    model = ClickModel(400, 600, seed=4)
    blocks, ut, it = _blocks(model, None)
    ctx = _capi.Context.default(0)
    train = GeneratedRows(ctx, model, 30_000, blocks, row0=0, keep_labels=True)
    val = GeneratedRows(ctx, model, 2_000, blocks, row0=30_000, keep_labels=True)
    n_features = train.shape[1]
    kw = dict(estimator="IPS", n_epochs=6, n_factors=24, lr=1e-4, batch_size=4096, seed=2, n_features=n_features,
              alpha=0.1, sampler="feistel")
    a = FactorizationMachines(**kw)
    la = a.fit({"features": train, "labels": None, "pscores": None}, {"features": val, "labels": None, "pscores": None})
    host = []
    for rows in (train, val):
        d = rows.download()
        ps = model.item_exposure[d["items"]] ** model.pow_used
        ff = FactoredFeatures([("id", "user", model.n_users), ("id", "item", model.n_items), ("ctx", d["ctx"]),
                               ("table", "user", ut), ("table", "item", it)], d["users"], d["items"])
        np.testing.assert_array_equal(d["targets"], d["labels"] / ps)       # the device divides like NumPy
        host.append({"features": ff, "labels": d["labels"], "pscores": ps})
    b = FactorizationMachines(**kw)
    lb = b.fit(*host)
    assert la == lb and np.all(np.isfinite(la[0]))
    np.testing.assert_array_equal(a.V(), b.V())
    np.testing.assert_array_equal(a.predict(X=val), b.predict(X=host[1]["features"]))
