"""GPU parity of the factored input format: with the flat step (step="flat") the row kernels
assemble x_t on the fly from [id | table row | context] blocks and must reproduce the hstacked-CSR path BIT FOR BIT
(same column order, same summation order), hence the reference goldens at 1e-9. The two-level step, which
step="auto" picks where it pays, has its own file (tests/test_two_level_gpu.py)."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import load_golden, golden_csr
from test_factored_host import kuairec_small_log

pytestmark = pytest.mark.gpu


def _log_and_golden(which):
    from rfm_b200.synth import make_coat_shaped
    if which == "coat":
        return make_coat_shaped(seed=2024), load_golden("coat_fm_ips_alpha01")
    return kuairec_small_log(), load_golden("kuairec_small_fm_ips")


def _factored(log, d):
    from rfm_b200.synth import factored_from_tables
    return {"features": factored_from_tables(log.tables, d["users"], d["items"], d["ctx"]), "labels": d["labels"],
            "pscores": d["pscores"]}


@pytest.mark.parametrize("which", ["coat", "kuairec"])
@pytest.mark.parametrize("dtype", ["float64", "float32"])
def test_fit_on_factored_rows_equals_csr_fit_and_the_reference(which, dtype):
    from rfm_b200.fm import FactorizationMachines
    from rfm_b200.synth import factored_from_tables
    log, g = _log_and_golden(which)
    kw = dict(estimator="IPS", n_epochs=int(g["n_epochs"]), n_factors=int(g["k"]), lr=float(g["lr"]),
              batch_size=int(g["B"]), seed=int(g["seed"]), n_features=log.n_features, alpha=float(g["alpha"]),
              dtype=dtype, step="flat")
    train = {"features": golden_csr(g, "train"), "labels": g["train_labels"], "pscores": g["train_pscores"]}
    val = {"features": golden_csr(g, "val"), "labels": g["val_labels"], "pscores": g["val_pscores"]}
    a = FactorizationMachines(**kw)
    tl_a, vl_a = a.fit(train, val)
    b = FactorizationMachines(**kw)
    ftrain, fval = _factored(log, log.fm_train), _factored(log, log.fm_val)
    tl_b, vl_b = b.fit(ftrain, fval)
    np.testing.assert_array_equal(tl_b, tl_a)
    np.testing.assert_array_equal(vl_b, vl_a)
    np.testing.assert_array_equal(b.V(), a.V())
    np.testing.assert_array_equal(b.w(), a.w())
    np.testing.assert_array_equal(b.w0(), a.w0())
    if dtype == "float64":
        np.testing.assert_allclose(tl_b, g["train_loss"], rtol=1e-9)
        np.testing.assert_allclose(vl_b, g["val_loss"], rtol=1e-9)
        np.testing.assert_allclose(b.V(), g["V"], rtol=1e-9, atol=1e-13)
        np.testing.assert_allclose(b.w(), g["w"], rtol=1e-9, atol=1e-13)
    t = log.test_frame
    ftest = factored_from_tables(log.tables, t["user"], t["item"], None if which == "coat" else np.zeros(t["user"].size))
    scores = b.predict(X=ftest)
    ref_test = golden_csr(g, "test")
    ref_test.eliminate_zeros()   # the synthetic test rows store the context value 0.0 explicitly; a stacked matrix would not
    np.testing.assert_array_equal(scores, a.predict(X=ref_test))
    if dtype == "float64":
        np.testing.assert_allclose(scores, g["test_scores"], rtol=1e-9)
    assert b.logloss(fval) == a.logloss(val)
    assert b.last_fit_stats["h2d_bytes_rows"] * 3 < a.last_fit_stats["h2d_bytes_rows"]


def test_device_sampler_and_row_subsets_agree_with_csr():
    from rfm_b200.fm import FactorizationMachines
    log, g = _log_and_golden("kuairec")
    kw = dict(estimator="IPS", n_epochs=6, n_factors=16, lr=1e-4, batch_size=1500, seed=3, n_features=log.n_features,
              alpha=0.1, sampler="feistel", step="flat")
    a, b = FactorizationMachines(**kw), FactorizationMachines(**kw)
    la = a.fit(log.fm_train, log.fm_val)
    lb = b.fit(_factored(log, log.fm_train), _factored(log, log.fm_val))
    assert la == lb
    np.testing.assert_array_equal(a.V(), b.V())
    ff = _factored(log, log.fm_train)["features"]
    np.testing.assert_array_equal(b.predict(X=ff[100:900]), a.predict(X=log.fm_train["features"][100:900]))


def test_ragged_tables_empty_rows_and_wide_context():
    """Rows of very different lengths (the packed triple layout instead of the fixed stride), entities with no side
    features at all, a 3-column context block with exact zeros, int32 ids, int8 labels."""
    from rfm_b200.factored import FactoredFeatures
    from rfm_b200.fm import FactorizationMachines
    rng = np.random.default_rng(12)
    n_users, n_items, n = 40, 55, 3000
    ut = sp.random(n_users, 30, density=0.3, format="csr", random_state=3)
    ut = sp.vstack([ut[:20], sp.csr_matrix((20, 30))]).tocsr()           # half the users: no side features
    it = sp.random(n_items, 9, density=0.15, format="csr", random_state=4)
    users, items = rng.integers(0, n_users, n).astype(np.int32), rng.integers(0, n_items, n).astype(np.int32)
    ctx = rng.normal(size=(n, 3))
    ctx[rng.random((n, 3)) < 0.2] = 0.0
    ff = FactoredFeatures([("table", "user", ut), ("ctx", ctx), ("id", "item", n_items), ("table", "item", it),
                           ("id", "user", n_users)], users, items)
    X = ff.tocsr()
    y = (rng.random(n) < 0.4).astype(np.int8)
    ps = rng.uniform(0.2, 1.0, n)
    val_sel = np.arange(0, n, 7)
    kw = dict(estimator="IPS", n_epochs=8, n_factors=20, lr=5e-4, batch_size=700, seed=5, n_features=X.shape[1],
              alpha=0.2, step="flat")
    a, b = FactorizationMachines(**kw), FactorizationMachines(**kw)
    la = a.fit({"features": X, "labels": y.astype(np.int64), "pscores": ps},
               {"features": X[val_sel], "labels": y[val_sel].astype(np.int64), "pscores": ps[val_sel]})
    lb = b.fit({"features": ff, "labels": y, "pscores": ps},
               {"features": ff[val_sel], "labels": y[val_sel], "pscores": ps[val_sel]})
    # scipy drops the context block's exact zeros and so does the on-the-fly assembly: same entries, same positions
    assert la == lb
    np.testing.assert_array_equal(a.V(), b.V())
    np.testing.assert_array_equal(a.predict(X=X), b.predict(X=ff))


def test_ids_out_of_range_are_refused():
    from rfm_b200 import _capi
    from rfm_b200.factored import FactoredFeatures, FactoredRows
    ff = FactoredFeatures([("id", "user", 5), ("id", "item", 4)], [0, 5], [0, 1])
    with pytest.raises(ValueError, match="user id"):
        FactoredRows(_capi.Context.default(0), ff)
    ff = FactoredFeatures([("id", "user", 5), ("table", "item", sp.identity(3, format="csr"))], [0, 1], [0, 3])
    with pytest.raises(ValueError, match="item id"):
        FactoredRows(_capi.Context.default(0), ff)


def test_mixing_csr_train_with_factored_val_is_refused():
    from rfm_b200.fm import FactorizationMachines
    log, g = _log_and_golden("coat")
    m = FactorizationMachines("IPS", 2, 8, 1e-3, 100, 1, log.n_features)
    with pytest.raises(ValueError, match="both be CSR or both be factored"):
        m.fit(log.fm_train, _factored(log, log.fm_val))


def test_per_item_pscores_equal_per_row_pscores():
    """The propensity is an item-level table in the reference's loaders (kuairec/loader.py:160-168); handing the table
    over (PerItem) instead of the gathered column gives the same bits."""
    from rfm_b200.factored import PerItem
    from rfm_b200.fm import FactorizationMachines
    log, g = _log_and_golden("kuairec")
    table = log.tables["item_pscore"]
    np.testing.assert_array_equal(table[log.fm_train["items"]], log.fm_train["pscores"])
    kw = dict(estimator="IPS", n_epochs=int(g["n_epochs"]), n_factors=int(g["k"]), lr=float(g["lr"]),
              batch_size=int(g["B"]), seed=int(g["seed"]), n_features=log.n_features, alpha=float(g["alpha"]),
              step="flat")
    a, b = FactorizationMachines(**kw), FactorizationMachines(**kw)
    ftrain, fval = _factored(log, log.fm_train), _factored(log, log.fm_val)
    la = a.fit(ftrain, fval)
    lb = b.fit(dict(ftrain, pscores=PerItem(table)), dict(fval, pscores=PerItem(table)))
    assert la == lb
    np.testing.assert_array_equal(a.V(), b.V())
    np.testing.assert_allclose(lb[0], g["train_loss"], rtol=1e-9)
    with pytest.raises(ValueError, match="item id"):
        FactorizationMachines(**kw).fit(dict(ftrain, pscores=PerItem(table[:10])), fval)


def test_materialized_on_device_equals_both_paths():
    """rfm_rows_materialize assembles the stacked CSR on the device from factored rows: fit / predict through it are
    bit-identical to the factored path and to the CSR uploaded from the host; it is what a long fit trains on."""
    from rfm_b200 import _capi
    from rfm_b200.factored import FactoredRows
    from rfm_b200.fm import FactorizationMachines
    log, g = _log_and_golden("kuairec")
    kw = dict(estimator="IPS", n_epochs=int(g["n_epochs"]), n_factors=int(g["k"]), lr=float(g["lr"]),
              batch_size=int(g["B"]), seed=int(g["seed"]), n_features=log.n_features, alpha=float(g["alpha"]),
              step="flat")
    ftrain, fval = _factored(log, log.fm_train), _factored(log, log.fm_val)
    runs = []
    for policy, data in (("never", (ftrain, fval)), ("always", (ftrain, fval)), ("never", (log.fm_train, log.fm_val))):
        m = FactorizationMachines(materialize=policy, **kw)
        runs.append((m.fit(*data), m.V().copy()))
    assert runs[0][0] == runs[1][0] == runs[2][0]
    np.testing.assert_array_equal(runs[0][1], runs[1][1])
    np.testing.assert_array_equal(runs[0][1], runs[2][1])
    np.testing.assert_allclose(runs[1][0][0], g["train_loss"], rtol=1e-9)
    # the assembled arrays themselves: row pointers, columns, values, targets == the host CSR's device copy
    import torch
    from rfm_b200.dist import _DeviceArray
    ctx = _capi.Context.default(0)
    X = log.fm_train["features"]
    host = _capi.CsrRows(ctx, X, log.fm_train["labels"], log.fm_train["pscores"])
    mat = _capi.MaterializedRows(FactoredRows(ctx, ftrain["features"], ftrain["labels"], ftrain["pscores"]))
    from ctypes import byref, c_void_p
    ptrs = []
    for rows in (host, mat):
        out = [c_void_p() for _ in range(4)]
        _capi.check(_capi.lib().rfm_csr_device_ptrs(rows.handle, *[byref(p) for p in out]))
        ptrs.append([p.value for p in out])
    for pa, pb, nbytes in zip(ptrs[0], ptrs[1], ((X.shape[0] + 1) * 8, X.nnz * 4, X.nnz * 8, X.shape[0] * 8)):
        ta = torch.as_tensor(_DeviceArray(pa, nbytes, "|u1"), device="cuda:0")
        tb = torch.as_tensor(_DeviceArray(pb, nbytes, "|u1"), device="cuda:0")
        assert torch.equal(ta, tb)
