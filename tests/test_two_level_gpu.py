"""GPU parity of the two-level FM step (csrc/two_level.cuh): per-entity aggregates forward, per-entity sums backward.
It computes the reference's update (src/fm.py:80-88, 135-187) with the sums associated per entity first, so it is held
to the goldens of the unmodified reference and to the CPU oracle at the same 1e-9 as the flat step, to the flat step
itself at 1e-11, and to bit-reproducibility run to run."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import load_golden
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
def test_two_level_fit_matches_the_reference_goldens_and_the_flat_step(which):
    from rfm_b200.fm import FactorizationMachines
    log, g = _log_and_golden(which)
    kw = dict(estimator="IPS", n_epochs=int(g["n_epochs"]), n_factors=int(g["k"]), lr=float(g["lr"]),
              batch_size=int(g["B"]), seed=int(g["seed"]), n_features=log.n_features, alpha=float(g["alpha"]))
    ftrain, fval = _factored(log, log.fm_train), _factored(log, log.fm_val)
    flat = FactorizationMachines(step="flat", **kw)
    tl_a, vl_a = flat.fit(ftrain, fval)
    assert flat.last_fit_stats["two_level"] is False
    runs = []
    for _ in range(2):
        m = FactorizationMachines(step="two_level", **kw)
        runs.append((m.fit(ftrain, fval), m.w0().copy(), m.w().copy(), m.V().copy()))
        assert m.last_fit_stats["two_level"] is True
    (tl_b, vl_b), w0, w, V = runs[0]
    assert runs[0][0] == runs[1][0]                                    # run to run: the same bits
    for a, b in zip(runs[0][1:], runs[1][1:]):
        np.testing.assert_array_equal(a, b)
    np.testing.assert_allclose(tl_b, g["train_loss"], rtol=1e-9)       # the unmodified reference
    np.testing.assert_allclose(vl_b, g["val_loss"], rtol=1e-9)
    np.testing.assert_allclose(V, g["V"], rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(w, g["w"], rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(w0, g["w0"], rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(tl_b, tl_a, rtol=1e-11)                 # the flat step: same sums, other association
    np.testing.assert_allclose(vl_b, vl_a, rtol=1e-11)
    np.testing.assert_allclose(V, flat.V(), rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(m.predict(X=fval["features"]), flat.predict(X=fval["features"]), rtol=1e-11)


def test_two_level_with_the_device_sampler_and_the_evaluator_chain():
    """Feistel batches, and the per-epoch evaluator hook (src/fm.py:104-110) running on the real parameters between
    two-level epochs: val_metrics against the golden's."""
    from rfm_b200.evaluate import ValEvaluator
    from rfm_b200.fm import FactorizationMachines
    log, g = _log_and_golden("kuairec")
    kw = dict(estimator="IPS", n_epochs=6, n_factors=16, lr=1e-4, batch_size=1500, seed=3, n_features=log.n_features,
              alpha=0.1, sampler="feistel")
    ftrain, fval = _factored(log, log.fm_train), _factored(log, log.fm_val)
    a, b = FactorizationMachines(step="flat", **kw), FactorizationMachines(step="two_level", **kw)
    la, lb = a.fit(ftrain, fval), b.fit(ftrain, fval)
    np.testing.assert_allclose(lb[0], la[0], rtol=1e-11)
    np.testing.assert_allclose(lb[1], la[1], rtol=1e-11)
    np.testing.assert_allclose(b.V(), a.V(), rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(b.w(), a.w(), rtol=1e-9, atol=1e-13)
    # the evaluator hook reads the REAL parameters between two-level epochs: the Coat golden's per-epoch val_metrics
    from conftest import golden_csr, golden_frame
    from rfm_b200.synth import make_coat_shaped
    log, g = make_coat_shaped(seed=2024), load_golden("coat_fm_ips_alpha01")
    ev = ValEvaluator(interaction_df=golden_frame(g), features={"FM": golden_csr(g, "test")}, k=5, metric_name="DCG")
    m = FactorizationMachines(estimator="IPS", n_epochs=int(g["n_epochs"]), n_factors=int(g["k"]), lr=float(g["lr"]),
                              batch_size=int(g["B"]), seed=int(g["seed"]), n_features=log.n_features,
                              alpha=float(g["alpha"]), evaluator=ev, step="two_level")
    tl, _ = m.fit(_factored(log, log.fm_train), _factored(log, log.fm_val))
    assert m.last_fit_stats["two_level"] is True
    np.testing.assert_allclose(tl, g["train_loss"], rtol=1e-9)
    np.testing.assert_allclose(m.val_metrics, g["val_metrics"], rtol=1e-9)


def test_ragged_tables_empty_entities_wide_context_and_foreign_val_tables():
    """Entities without side features, a 3-column context block with exact zeros, blocks in an unusual order, int32 ids
    and int8 labels; then val rows keyed by OTHER tables (their loss must come from their own tables, through the
    flat row pass on the real parameters)."""
    from rfm_b200.factored import FactoredFeatures
    from rfm_b200.fm import FactorizationMachines
    rng = np.random.default_rng(12)
    n_users, n_items, n = 40, 55, 3000
    ut = sp.random(n_users, 30, density=0.3, format="csr", random_state=3)
    ut = sp.vstack([ut[:20], sp.csr_matrix((20, 30))]).tocsr()
    it = sp.random(n_items, 9, density=0.15, format="csr", random_state=4)
    users, items = rng.integers(0, n_users, n).astype(np.int32), rng.integers(0, n_items, n).astype(np.int32)
    ctx = rng.normal(size=(n, 3))
    ctx[rng.random((n, 3)) < 0.2] = 0.0
    blocks = lambda user_table, c: [("table", "user", user_table), ("ctx", c), ("id", "item", n_items),
                                    ("table", "item", it), ("id", "user", n_users)]
    ff = FactoredFeatures(blocks(ut, ctx), users, items)
    X = ff.tocsr()
    y = (rng.random(n) < 0.4).astype(np.int8)
    ps = rng.uniform(0.2, 1.0, n)
    val_sel = np.arange(0, n, 7)
    kw = dict(estimator="IPS", n_epochs=8, n_factors=20, lr=5e-4, batch_size=700, seed=5, n_features=X.shape[1],
              alpha=0.2)
    a, b = FactorizationMachines(**kw), FactorizationMachines(step="two_level", **kw)
    la = a.fit({"features": X, "labels": y.astype(np.int64), "pscores": ps},
               {"features": X[val_sel], "labels": y[val_sel].astype(np.int64), "pscores": ps[val_sel]})
    lb = b.fit({"features": ff, "labels": y, "pscores": ps},
               {"features": ff[val_sel], "labels": y[val_sel], "pscores": ps[val_sel]})
    assert b.last_fit_stats["two_level"] is True
    np.testing.assert_allclose(lb[0], la[0], rtol=1e-11)
    np.testing.assert_allclose(lb[1], la[1], rtol=1e-11)
    np.testing.assert_allclose(b.V(), a.V(), rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(b.w(), a.w(), rtol=1e-9, atol=1e-13)
    # val rows keyed by a different user table
    ut2 = (ut * 2.0).tocsr()
    fv = FactoredFeatures(blocks(ut2, ctx[val_sel]), users[val_sel], items[val_sel])
    Xv = fv.tocsr()
    c, d = FactorizationMachines(**kw), FactorizationMachines(step="two_level", **kw)
    lc = c.fit({"features": X, "labels": y.astype(np.int64), "pscores": ps},
               {"features": Xv, "labels": y[val_sel].astype(np.int64), "pscores": ps[val_sel]})
    ld = d.fit({"features": ff, "labels": y, "pscores": ps},
               {"features": fv, "labels": y[val_sel], "pscores": ps[val_sel]})
    np.testing.assert_allclose(ld[0], lc[0], rtol=1e-11)
    np.testing.assert_allclose(ld[1], lc[1], rtol=1e-11)
    assert not np.allclose(lc[1], la[1], rtol=1e-6)        # the other table does change the val loss


@pytest.mark.parametrize("k", [65, 130, 300])
@pytest.mark.parametrize("n_ctx", [0, 1, 2])
def test_two_level_factor_counts_and_context_widths(k, n_ctx):
    """n_factors beyond one 64-wide chunk (16- and 32-lane row groups, the reference's own 300) with no, one (the lean
    row kernel + the dense-column CTAs of level 1) and two context columns (the generic row kernel, every context
    column in the sort): against the flat step on the stacked matrix."""
    from rfm_b200.factored import FactoredFeatures
    from rfm_b200.fm import FactorizationMachines
    rng = np.random.default_rng(100 * k + n_ctx)
    n_users, n_items, n = 60, 45, 2500
    ut = sp.random(n_users, 12, density=0.4, format="csr", random_state=1)
    it = sp.random(n_items, 7, density=0.5, format="csr", random_state=2)
    users, items = rng.integers(0, n_users, n), rng.integers(0, n_items, n)
    blocks = [("id", "user", n_users), ("id", "item", n_items)]
    if n_ctx:
        blocks.append(("ctx", rng.normal(size=(n, n_ctx))))
    blocks += [("table", "user", ut), ("table", "item", it)]
    ff = FactoredFeatures(blocks, users, items)
    X = ff.tocsr()
    y, ps = (rng.random(n) < 0.5).astype(np.int64), rng.uniform(0.3, 1.0, n)
    kw = dict(estimator="IPS", n_epochs=5, n_factors=k, lr=3e-4, batch_size=900, seed=k, n_features=X.shape[1], alpha=0.2)
    a, b = FactorizationMachines(**kw), FactorizationMachines(step="two_level", **kw)
    la = a.fit({"features": X, "labels": y, "pscores": ps}, {"features": X[:300], "labels": y[:300], "pscores": ps[:300]})
    lb = b.fit({"features": ff, "labels": y, "pscores": ps}, {"features": ff[:300], "labels": y[:300], "pscores": ps[:300]})
    assert b.last_fit_stats["two_level"] is True
    np.testing.assert_allclose(lb[0], la[0], rtol=1e-11)
    np.testing.assert_allclose(lb[1], la[1], rtol=1e-11)
    np.testing.assert_allclose(b.V(), a.V(), rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(b.w(), a.w(), rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(b.w0(), a.w0(), rtol=1e-9, atol=1e-13)


@pytest.mark.parametrize("opt", ["adam", "sgd_l2"])
def test_two_level_feeds_the_dense_optimizers(opt):
    """Adam / SGD + L2 take the batch gradient from the two-level passes (the gradient buffer the data-parallel step
    exchanges): against the flat step's."""
    from rfm_b200.fm import FactorizationMachines
    log, g = _log_and_golden("kuairec")
    kw = dict(estimator="IPS", n_epochs=5, n_factors=16, lr=1e-3, batch_size=2000, seed=9, n_features=log.n_features,
              alpha=0.1, sampler="feistel")
    kw.update(dict(optimizer="adam") if opt == "adam" else dict(l2=1e-3))
    ftrain, fval = _factored(log, log.fm_train), _factored(log, log.fm_val)
    a, b = FactorizationMachines(step="flat", **kw), FactorizationMachines(step="two_level", **kw)
    la, lb = a.fit(ftrain, fval), b.fit(ftrain, fval)
    assert b.last_fit_stats["two_level"] is True
    np.testing.assert_allclose(lb[0], la[0], rtol=1e-9)
    np.testing.assert_allclose(lb[1], la[1], rtol=1e-9)
    np.testing.assert_allclose(b.V(), a.V(), rtol=1e-7, atol=1e-12)   # Adam divides by sqrt(v): rounding is amplified


def test_float32_two_level_stays_within_the_north_star_tolerance():
    from rfm_b200.fm import FactorizationMachines
    log, g = _log_and_golden("kuairec")
    kw = dict(estimator="IPS", n_epochs=int(g["n_epochs"]), n_factors=int(g["k"]), lr=float(g["lr"]),
              batch_size=int(g["B"]), seed=int(g["seed"]), n_features=log.n_features, alpha=float(g["alpha"]),
              dtype="float32", step="two_level")
    m = FactorizationMachines(**kw)
    tl, vl = m.fit(_factored(log, log.fm_train), _factored(log, log.fm_val))
    np.testing.assert_allclose(tl, g["train_loss"], rtol=1e-5)
    np.testing.assert_allclose(vl, g["val_loss"], rtol=1e-5)


@pytest.fixture(scope="module")
def big_log():
    from rfm_b200.synth import make_kuairec_shaped
    return make_kuairec_shaped(seed=7, n_train=3_000_000, n_val=2000, build_mf=False, build_eval=False)


@pytest.mark.parametrize("sampler", ["legacy", "feistel"])
def test_bench_shape_two_level_epochs_match_the_oracle(big_log, sampler):
    """The benchmark's own shape (3 M of its 12 M rows, B = 65,536, k = 64): auto picks the two-level step there;
    three epochs against the CPU oracle's fm_fit on the STACKED matrix -- every loss and parameter at 1e-9."""
    from oracle import fm_oracle, sampler_oracle
    from rfm_b200.fm import FactorizationMachines
    B, K_FACTORS, LR, n_epochs = 65536, 64, 9e-6, 3
    m = FactorizationMachines("IPS", n_epochs, K_FACTORS, LR, B, 12345, big_log.n_features, sampler=sampler)
    w0, w, V = m.w0().copy(), m.w().copy(), m.V().copy()
    tl, vl = m.fit(_factored(big_log, big_log.fm_train), _factored(big_log, big_log.fm_val))
    assert m.last_fit_stats["two_level"] is True                      # step="auto": the cost model says it pays here
    pick = fm_oracle.legacy_batch if sampler == "legacy" else \
        (lambda n, b, e: sampler_oracle.feistel_batch(n, b, e, 12345))
    (rw0, rw, rV), rtl, rvl = fm_oracle.fm_fit(big_log.fm_train, big_log.fm_val, n_epochs, B, LR, w0, w, V, sampler=pick)
    np.testing.assert_allclose(tl, rtl, rtol=1e-9)
    np.testing.assert_allclose(vl, rvl, rtol=1e-9)
    np.testing.assert_allclose(m.w0(), rw0, rtol=1e-9)
    np.testing.assert_allclose(m.w(), rw, rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(m.V(), rV, rtol=1e-9, atol=1e-13)


def test_two_level_gradient_is_additive_over_slices(big_log):
    """What the data-parallel split relies on, through the two-level passes: gradient(slice a) + gradient(slice b) ==
    gradient(batch); and it equals the flat step's gradient buffer."""
    from ctypes import byref, c_int64, c_void_p
    from rfm_b200._capi import check, lib
    from rfm_b200.dist import device_tensor
    from rfm_b200.fm import FactorizationMachines, _FmTrainer
    B = 65536
    grads = {}
    for step in ("flat", "two_level"):
        m = FactorizationMachines("IPS", 4, 64, 9e-6, B, 12345, big_log.n_features, alpha=0.05, sampler="feistel")
        m._context()
        ftrain, fval = _factored(big_log, big_log.fm_train), _factored(big_log, big_log.fm_val)
        train = m._rows(ftrain["features"], ftrain["labels"], ftrain["pscores"])
        val = m._rows(fval["features"], fval["labels"], fval["pscores"])
        m.sync_to_device()
        trainer = _FmTrainer(m._dev, train, val, B, 8)
        if step == "two_level":
            assert trainer.set_two_level(1)
        n, p = c_int64(), c_void_p()
        check(lib().rfm_fm_grad_size(trainer.handle, byref(n)))
        check(lib().rfm_fm_grad_ptr_dev(trainer.handle, byref(p)))
        g = device_tensor(p.value, n.value, m.dtype, m.device)

        def grad(begin, count):
            check(lib().rfm_fm_grad_epoch_sampled(trainer.handle, 12345, 3, begin, count))
            m._ctx.synchronize()
            return g.clone().cpu().numpy()

        full = grad(0, B)
        cut = B // 3 + 5
        parts = grad(0, cut) + grad(cut, B - cut)
        scale = np.abs(full).max()
        np.testing.assert_allclose(parts, full, rtol=1e-9, atol=1e-12 * scale)
        grads[step] = full
        trainer.close()
    scale = np.abs(grads["flat"]).max()
    np.testing.assert_allclose(grads["two_level"], grads["flat"], rtol=1e-9, atol=1e-12 * scale)
