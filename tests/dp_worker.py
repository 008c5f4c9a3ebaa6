"""torchrun worker for tests/test_dist_gpu.py: data-parallel FM fit vs the reference goldens."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "relevance-factorizationmachine_b200"), ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from conftest import load_golden, golden_csr  # noqa: E402


def main():
    from rfm_b200 import dist as rdist
    from rfm_b200.fm import FactorizationMachines
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    env = rdist.init(local_rank)
    for name, sampler in (("coat_fm_ips_alpha01", "legacy"), ("kuairec_small_fm_ips", "legacy"),
                          ("kuairec_small_fm_ips_alpha01", "feistel")):
        g = load_golden(name)
        train = {"features": golden_csr(g, "train"), "labels": g["train_labels"], "pscores": g["train_pscores"]}
        val = {"features": golden_csr(g, "val"), "labels": g["val_labels"], "pscores": g["val_pscores"]}
        kw = dict(estimator="IPS", n_epochs=int(g["n_epochs"]), n_factors=int(g["k"]), lr=float(g["lr"]),
                  batch_size=int(g["B"]), seed=int(g["seed"]), n_features=train["features"].shape[1],
                  alpha=float(g["alpha"]), sampler=sampler, device=local_rank)
        m = FactorizationMachines(distributed=env, **kw)
        tl, vl = m.fit(train, val)
        if sampler == "legacy":
            ref_tl, ref_vl, ref_V, ref_w = g["train_loss"], g["val_loss"], g["V"], g["w"]
        else:   # the fused single-GPU path with the same device sampler is the comparison
            s = FactorizationMachines(**kw)
            ref_tl, ref_vl = s.fit(train, val)
            ref_V, ref_w = s.V(), s.w()
        np.testing.assert_allclose(tl, ref_tl, rtol=1e-9, err_msg=name)
        np.testing.assert_allclose(vl, ref_vl, rtol=1e-9, err_msg=name)
        np.testing.assert_allclose(m.V(), ref_V, rtol=1e-9, atol=1e-13, err_msg=name)
        np.testing.assert_allclose(m.w(), ref_w, rtol=1e-9, atol=1e-13, err_msg=name)
        # every rank holds the same bits after the identical apply
        import torch
        mine = torch.from_numpy(m.V().copy()).cuda(local_rank)
        lo, hi = mine.clone(), mine.clone()
        env.dist.all_reduce(lo, op=env.dist.ReduceOp.MIN)
        env.dist.all_reduce(hi, op=env.dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), "ranks diverged"
    # sharded upload (1/G of the rows over PCIe per rank, the rest over NVLink) == the replicated upload, byte for byte
    from rfm_b200 import _capi
    from rfm_b200.dist import _DeviceArray, sharded_csr_rows
    import torch
    X = train["features"]
    ctx = m._context()
    for dtype in ("float64", "float32"):
        full = _capi.CsrRows(ctx, X, train["labels"], train["pscores"], dtype)
        part = sharded_csr_rows(ctx, X, train["labels"], train["pscores"], dtype, env)
        es = 8 if dtype == "float64" else 4
        sizes = ((X.shape[0] + 1) * 8, X.nnz * 4, X.nnz * es, X.shape[0] * es)
        for pa, pb, nbytes in zip(full.device_ptrs(), part.device_ptrs(), sizes):
            if nbytes:
                ta = torch.as_tensor(_DeviceArray(pa, nbytes, "|u1"), device="cuda:%d" % local_rank)
                tb = torch.as_tensor(_DeviceArray(pb, nbytes, "|u1"), device="cuda:%d" % local_rank)
                assert torch.equal(ta, tb), "sharded upload differs from the replicated one (%s)" % dtype
    os.environ["RFM_DP_UPLOAD_MIN_ROWS"] = "0"                # and a whole fit through it gives the same bits
    m2 = FactorizationMachines(distributed=env, **kw)
    tl2, vl2 = m2.fit(train, val)
    np.testing.assert_array_equal(tl2, tl)
    np.testing.assert_array_equal(m2.V(), m.V())
    os.environ["RFM_DP_UPLOAD_MIN_ROWS"] = "1000000"
    # row-sharded predict == single-process predict, bit for bit
    test_X = golden_csr(g, "test")
    np.testing.assert_array_equal(rdist.sharded_predict(m, test_X, env), m.predict(X=test_X))
    # item-sharded full-catalog scoring: merged per-rank lists == one pass over the whole catalog, bit for bit,
    # through the library's NVLink exchange (global thresholds + user-partitioned K-way merge) and through NCCL
    from rfm_b200.score import TopKScorer
    from rfm_b200.dist import slice_bounds
    rng = np.random.default_rng(21)
    cases = [(500, 3000, 64, 9, "normal"), (300, 5000, 128, 100, "normal"), (257, 700, 64, 20, "ties"),
             (130, 300, 32, 9, "normal"), (64, 2600, 200, 9, "normal")]       # last: k > 128 -> exact path per shard
    for n_u, n_i, k, K, kind in cases:
        A, C, beta = rng.normal(size=(n_u, k)) * 0.4, rng.normal(size=(n_i, k)) * 0.4, rng.normal(size=n_i) * 0.2
        if kind == "ties":                                       # few distinct scores: exact ties across shards
            A, C, beta = np.round(A), np.round(C), np.round(beta)
        sc = TopKScorer(A, C, None, beta, 0.0, device=local_rank)
        full_items, full_scores = sc.topk(K)
        for exchange in ("nvlink", "nccl"):
            items, scores = rdist.sharded_topk(sc, env, K, exchange=exchange)
            np.testing.assert_array_equal(items, full_items, err_msg="%s %s" % (exchange, (n_u, n_i, k, K)))
            np.testing.assert_array_equal(scores, full_scores)
        ub, ue, own_items, own_scores = rdist.sharded_topk(sc, env, K, gather=False, copy=False)
        assert (ub, ue) == slice_bounds(n_u, env.world, env.rank)
        np.testing.assert_array_equal(own_items, full_items[ub:ue])
        np.testing.assert_array_equal(own_scores, full_scores[ub:ue])
        items, scores = rdist.sharded_topk(sc, env, K, mode="exact")
        np.testing.assert_array_equal(items, full_items)
        if k <= 128:
            assert sc.last_stats["tensor_core_path"] is False    # the exact-mode call above
        sc.close()
    # full-catalog metrics with the catalog item-sharded: every rank reduces the users it owns after the merge, the
    # partial sums and per-item hit counts are all-reduced; equal to the single-GPU evaluation
    from rfm_b200.evaluate import FullCatalogEvaluator
    U, I, k = 300, 2100, 64
    A, C, beta = rng.normal(size=(U, k)) * 0.4, rng.normal(size=(I, k)) * 0.4, rng.normal(size=I) * 0.2
    hu = np.repeat(np.arange(U), 10)
    hi = np.concatenate([rng.choice(I, 10, replace=False) for _ in range(U)])
    hl = (rng.random(hu.size) < 0.4).astype(np.int64)
    hl[hu % 9 == 0] = 0
    theta = rng.uniform(0.1, 1.0, size=I)
    used = {"DCG", "CatalogCoverage", "Recall", "MAP", "Gini"}
    ev = FullCatalogEvaluator({"user": hu, "item": hi, "label": hl}, theta, [1, 3, 9, 20], used, U, I)
    sc = TopKScorer(A, C, None, beta, 0.1, device=local_rank)
    single = ev.evaluate(sc)
    multi = ev.evaluate(sc, env=env)
    for name in single:
        np.testing.assert_allclose(multi[name], single[name], rtol=1e-12, err_msg=name)
    np.testing.assert_array_equal(np.array(multi["CatalogCoverage"]), np.array(single["CatalogCoverage"]))
    sc.close()
    # data-parallel fit on factored rows == the same fit on the stacked CSR, bit for bit
    from rfm_b200.synth import factored_from_tables, make_kuairec_shaped
    kr = make_kuairec_shaped(seed=2025, n_users=300, n_items=400, n_train=6000, n_val=600, eval_users=60,
                             eval_items=200, eval_rows_per_user=20)
    fac = lambda d: {"features": factored_from_tables(kr.tables, d["users"], d["items"], d["ctx"]),
                     "labels": d["labels"], "pscores": d["pscores"]}
    kw2 = dict(estimator="IPS", n_epochs=5, n_factors=32, lr=1e-4, batch_size=2000, seed=1, n_features=kr.n_features,
               alpha=0.1, sampler="feistel", device=local_rank, distributed=env, step="flat")
    ma, mb = FactorizationMachines(**kw2), FactorizationMachines(**kw2)
    la, lb = ma.fit(kr.fm_train, kr.fm_val), mb.fit(fac(kr.fm_train), fac(kr.fm_val))
    assert la == lb
    np.testing.assert_array_equal(ma.V(), mb.V())
    # sharded upload of factored rows (1/G of the per-interaction arrays per rank over PCIe, NVLink broadcasts for the
    # rest) == the replicated upload, byte for byte; and a whole fit through it gives the same bits
    from rfm_b200.factored import FactoredRows, PerItem
    from rfm_b200.dist import sharded_factored_rows
    ftr = fac(kr.fm_train)
    for dtype in ("float64", "float32"):
        for ps_arg in (ftr["pscores"], PerItem(kr.tables["item_pscore"])):
            y8 = ftr["labels"].astype(np.int8)
            full = FactoredRows(mb._context(), ftr["features"], y8, ps_arg, dtype)
            part = sharded_factored_rows(mb._context(), ftr["features"], y8, ps_arg, dtype, env)
            es = 8 if dtype == "float64" else 4
            n_r = ftr["features"].shape[0]
            for pa, pb, nbytes in zip(full.device_ptrs(), part.device_ptrs(), (n_r * 4, n_r * 4, n_r * es * full.n_ctx, n_r * es)):
                if nbytes and pa and pb:
                    ta = torch.as_tensor(_DeviceArray(pa, nbytes, "|u1"), device="cuda:%d" % local_rank)
                    tb = torch.as_tensor(_DeviceArray(pb, nbytes, "|u1"), device="cuda:%d" % local_rank)
                    assert torch.equal(ta, tb), "sharded factored upload differs from the replicated one (%s)" % dtype
    os.environ["RFM_DP_UPLOAD_MIN_ROWS"] = "0"
    me = FactorizationMachines(**kw2)
    le = me.fit(fac(kr.fm_train), fac(kr.fm_val))
    os.environ["RFM_DP_UPLOAD_MIN_ROWS"] = "1000000"
    assert le == lb
    np.testing.assert_array_equal(me.V(), mb.V())
    # the two-level step (csrc/two_level.cuh) feeding the same exchange: against the flat data-parallel fit and
    # against the single-GPU two-level fit (two associations of the same sums), ranks bit-identical
    mc = FactorizationMachines(**dict(kw2, step="two_level"))
    lc = mc.fit(fac(kr.fm_train), fac(kr.fm_val))
    assert mc.last_fit_stats["two_level"] is True
    md = FactorizationMachines(**dict(kw2, step="two_level", distributed=None))
    ld = md.fit(fac(kr.fm_train), fac(kr.fm_val))
    for ref_l, ref_m in ((lb, mb), (ld, md)):
        np.testing.assert_allclose(lc[0], ref_l[0], rtol=1e-11)
        np.testing.assert_allclose(lc[1], ref_l[1], rtol=1e-11)
        np.testing.assert_allclose(mc.V(), ref_m.V(), rtol=1e-9, atol=1e-13)
        np.testing.assert_allclose(mc.w(), ref_m.w(), rtol=1e-9, atol=1e-13)
    mine = torch.from_numpy(mc.V().copy()).cuda(local_rank)
    lo, hi = mine.clone(), mine.clone()
    env.dist.all_reduce(lo, op=env.dist.ReduceOp.MIN)
    env.dist.all_reduce(hi, op=env.dist.ReduceOp.MAX)
    assert torch.equal(lo, hi), "ranks diverged (two-level)"
    if env.rank == 0:
        print("DP_OK world=%d exchange=%s" % (env.world, os.environ.get("RFM_DP_EXCHANGE", "nvlink")))
    env.shutdown()


if __name__ == "__main__":
    main()
