"""Size-independent properties at the benchmark's sizes (BASELINE.json configs[2] shape, B = 65,536, k = 64; the
scoring grid of bench.py), where a NumPy run of the reference arithmetic is no longer a few-second check:
additivity of the batch gradient (what the data-parallel split relies on), fused step == gradient + apply,
bit-reproducibility, and for the tensor-core top-K: sorted output, exact scores, nothing outside the list beats
the K-th (on a sample of users), item-sharding invariance."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

B, K_FACTORS, LR = 65536, 64, 9e-6


@pytest.fixture(scope="module")
def big_log():
    from rfm_b200.synth import make_kuairec_shaped
    return make_kuairec_shaped(seed=7, n_train=3_000_000, n_val=2000, build_mf=False, build_eval=False)


def _model_and_trainer(log, seed=12345):
    from rfm_b200.fm import FactorizationMachines, _FmTrainer
    m = FactorizationMachines("IPS", 4, K_FACTORS, LR, B, seed, log.n_features, alpha=0.05, sampler="feistel")
    m._context()
    train = m._rows(log.fm_train["features"], log.fm_train["labels"], log.fm_train["pscores"])
    val = m._rows(log.fm_val["features"], log.fm_val["labels"], log.fm_val["pscores"])
    m.sync_to_device()
    return m, _FmTrainer(m._dev, train, val, B, 8), train, val


def _grad_view(trainer, model):
    from ctypes import byref, c_int64, c_void_p
    from rfm_b200._capi import check, lib
    from rfm_b200.dist import device_tensor
    n, p = c_int64(), c_void_p()
    check(lib().rfm_fm_grad_size(trainer.handle, byref(n)))
    check(lib().rfm_fm_grad_ptr_dev(trainer.handle, byref(p)))
    return device_tensor(p.value, n.value, model.dtype, model.device)


def test_batch_gradient_is_additive_over_slices(big_log):
    from rfm_b200._capi import check, lib
    m, trainer, _, _ = _model_and_trainer(big_log)
    g = _grad_view(trainer, m)

    def grad(begin, count):
        check(lib().rfm_fm_grad_epoch_sampled(trainer.handle, 12345, 3, begin, count))
        m._ctx.synchronize()
        return g.clone().cpu().numpy()

    full = grad(0, B)
    cut = B // 3 + 5
    parts = grad(0, cut) + grad(cut, B - cut)
    scale = np.abs(full).max()
    assert scale > 0
    np.testing.assert_allclose(parts, full, rtol=1e-9, atol=1e-12 * scale)
    assert np.count_nonzero(full[4:4 + big_log.n_features]) > 0.9 * big_log.n_features   # nearly every column is touched


def test_fused_step_equals_gradient_plus_apply_and_is_reproducible(big_log):
    from rfm_b200._capi import check, lib
    outs = []
    for variant in ("fused", "fused", "split"):
        m, trainer, _, _ = _model_and_trainer(big_log)
        for epoch in range(3):
            if variant == "fused":
                check(lib().rfm_fm_train_epoch_sampled(trainer.handle, 12345, epoch, B, LR, epoch))
            else:
                check(lib().rfm_fm_grad_epoch_sampled(trainer.handle, 12345, epoch, 0, B))
                check(lib().rfm_fm_apply_grad(trainer.handle, LR))
        m.sync_to_host()
        outs.append((m.w0().copy(), m.w().copy(), m.V().copy()))
        trainer.close()
    for a, b in zip(outs[0], outs[1]):
        np.testing.assert_array_equal(a, b)                               # run-to-run: same bits
    for a, b in zip(outs[0], outs[2]):
        np.testing.assert_allclose(a, b, rtol=1e-10, atol=1e-15)          # two associations of the same sums
    assert np.abs(outs[0][2] - _model_and_trainer(big_log)[0].V()).max() > 0    # the step moved the parameters


def test_post_update_loss_equals_logloss_of_the_sampled_rows(big_log):
    from rfm_b200 import _capi
    from rfm_b200._capi import check, lib, ptr
    m, trainer, _, _ = _model_and_trainer(big_log)
    check(lib().rfm_fm_train_epoch_sampled(trainer.handle, 12345, 0, B, LR, 0))
    tl, vl = np.empty(1), np.empty(1)
    check(lib().rfm_fm_trainer_losses(trainer.handle, 0, 1, ptr(tl), ptr(vl)))
    rows = _capi.feistel_batch(big_log.fm_train["features"].shape[0], B, 0, 12345)
    batch = {"features": big_log.fm_train["features"][rows], "labels": big_log.fm_train["labels"][rows],
             "pscores": big_log.fm_train["pscores"][rows]}
    np.testing.assert_allclose(tl[0], m.logloss(batch), rtol=1e-11)
    np.testing.assert_allclose(vl[0], m.logloss(big_log.fm_val), rtol=1e-11)


@pytest.mark.parametrize("k,K", [(64, 9), (128, 9)])
def test_large_grid_topk_properties(k, K):
    from rfm_b200.score import TopKScorer, merge_topk
    rng = np.random.default_rng(k)
    U, I = 8192, 131072
    A, C, beta = rng.normal(size=(U, k)) * 0.3, rng.normal(size=(I, k)) * 0.3, rng.normal(size=I) * 0.2
    sc = TopKScorer(A, C, None, beta, 0.125)
    items, scores = sc.topk(K)
    assert sc.last_stats["tensor_core_path"] and sc.last_stats["users_ranked_exactly"] == 0
    assert (items >= 0).all() and (items < I).all()
    assert all(len(set(row)) == K for row in items[:256])                       # no item twice
    assert (np.diff(scores, axis=1) <= 0).all()                                  # sorted, best first
    exact = np.einsum("ukf,uf->uk", C[items], A) + beta[items] + 0.125           # every returned score is exact
    np.testing.assert_allclose(scores, exact, rtol=1e-12, atol=1e-12)
    sample = rng.choice(U, 48, replace=False)                                    # nothing outside the list beats the K-th
    S = A[sample] @ C.T + beta[None, :] + 0.125
    ref = np.argsort(S, axis=1, kind="stable")[:, ::-1][:, :K]
    np.testing.assert_array_equal(items[sample], ref.astype(np.int32))
    parts = [sc.topk(K, item_range=(b, e)) for b, e in ((0, 32768), (32768, 98304), (98304, I))]   # tile-aligned shards
    m_items, m_scores = merge_topk([p[0] for p in parts], [p[1] for p in parts], K)
    np.testing.assert_array_equal(m_items, items)
    np.testing.assert_array_equal(m_scores, scores)


@pytest.mark.parametrize("sampler", ["legacy", "feistel"])
def test_bench_shape_epochs_match_the_oracle(big_log, sampler):
    """The benchmark's own shape (3 M of its 12 M rows, B = 65,536, k = 64) for three epochs against the CPU oracle's
    fm_fit -- every loss and every parameter at 1e-9 -- with the reference's sampler and with the device sampler."""
    from oracle import fm_oracle, sampler_oracle
    from rfm_b200.fm import FactorizationMachines
    n_epochs = 3
    m = FactorizationMachines("IPS", n_epochs, K_FACTORS, LR, B, 12345, big_log.n_features, sampler=sampler)
    w0, w, V = m.w0().copy(), m.w().copy(), m.V().copy()
    tl, vl = m.fit(big_log.fm_train, big_log.fm_val)
    pick = fm_oracle.legacy_batch if sampler == "legacy" else \
        (lambda n, b, e: sampler_oracle.feistel_batch(n, b, e, 12345))
    (rw0, rw, rV), rtl, rvl = fm_oracle.fm_fit(big_log.fm_train, big_log.fm_val, n_epochs, B, LR, w0, w, V, sampler=pick)
    np.testing.assert_allclose(tl, rtl, rtol=1e-9)
    np.testing.assert_allclose(vl, rvl, rtol=1e-9)
    np.testing.assert_allclose(m.w0(), rw0, rtol=1e-9)
    np.testing.assert_allclose(m.w(), rw, rtol=1e-9, atol=1e-13)
    np.testing.assert_allclose(m.V(), rV, rtol=1e-9, atol=1e-13)
