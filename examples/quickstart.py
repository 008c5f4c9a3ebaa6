"""Quickstart on a B200: the reference's workflow (main_coat.py:95-135) on synthetic Coat-shaped data, then the
full-catalog ranking the reference has no counterpart for. Run from the repository root after
`python __graft_entry__.py`:

    python examples/quickstart.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "relevance-factorizationmachine_b200"))

import numpy as np  # noqa: E402
from scipy.sparse import csr_matrix  # noqa: E402

# the reference's import paths resolve to the B200 implementation
from src.fm import FactorizationMachines  # noqa: E402
from src.mf import LogisticMatrixFactorization  # noqa: E402
from utils.evaluate import TestEvaluator  # noqa: E402

from rfm_b200.evaluate import FullCatalogEvaluator  # noqa: E402
from rfm_b200.score import TopKScorer, fm_factors, mf_factors  # noqa: E402
from rfm_b200.synth import factored_from_tables, make_coat_shaped  # noqa: E402


def main():
    log = make_coat_shaped(seed=12345)
    K = [1, 3, 5]
    used = {"DCG", "CatalogCoverage", "Recall"}
    evaluator = TestEvaluator(interaction_df=log.test_frame, features={}, K=K, used_metrics=used, n_items=log.n_items)

    # --- the reference's loop over models and estimators (main_coat.py:86-124), same constructor arguments ---
    for estimator in ("IPS", "Naive"):
        def pscores(d):
            return d["pscores"] if estimator == "IPS" else np.ones_like(d["pscores"])

        fm = FactorizationMachines(estimator=estimator, n_epochs=100, n_factors=32, lr=1e-3, batch_size=500,
                                   seed=12345, n_features=log.n_features, alpha=0.1)
        train = dict(log.fm_train, pscores=pscores(log.fm_train))
        val = dict(log.fm_val, pscores=pscores(log.fm_val))
        train_loss, val_loss = fm.fit(train, val)
        train_loss_fm = train_loss
        metrics = evaluator.evaluate(fm.predict(X=log.fm_test_features))
        print("FM  %-5s  loss %.4f -> %.4f   DCG@3 %.4f  coverage@3 %.3f" % (
            estimator, train_loss[0], train_loss[-1], metrics["DCG"][1], metrics["CatalogCoverage"][1]))

        mf = LogisticMatrixFactorization(estimator=estimator, n_epochs=100, n_factors=16, lr=1e-2, batch_size=500,
                                         seed=12345, n_users=log.n_users, n_items=log.n_items, reg=1e-4)
        mtrain = dict(log.mf_train, pscores=pscores(log.mf_train))
        mval = dict(log.mf_val, pscores=pscores(log.mf_val))
        train_loss, val_loss = mf.fit(mtrain, mval)
        metrics = evaluator.evaluate(mf.predict(X=log.mf_test_features))
        print("MF  %-5s  loss %.4f -> %.4f   DCG@3 %.4f  coverage@3 %.3f" % (
            estimator, train_loss[0], train_loss[-1], metrics["DCG"][1], metrics["CatalogCoverage"][1]))

    # --- the same FM fit with the rows handed over BEFORE scipy.sparse.hstack: the
    # preparer's blocks + one (user, item) pair per interaction; the two-level step then works per entity ---
    fac = lambda d: dict(d, features=factored_from_tables(log.tables, d["users"], d["items"], d["ctx"]))
    fm2 = FactorizationMachines(estimator="Naive", n_epochs=100, n_factors=32, lr=1e-3, batch_size=500, seed=12345,
                                n_features=log.n_features, alpha=0.1, step="two_level")
    loss2, _ = fm2.fit(dict(fac(log.fm_train), pscores=np.ones_like(log.fm_train["pscores"])),
                       dict(fac(log.fm_val), pscores=np.ones_like(log.fm_val["pscores"])))
    print("FM  Naive  factored rows, two-level step: loss %.4f -> %.4f  (stacked CSR above: -> %.4f); upload %d B vs %d B"
          % (loss2[0], loss2[-1], train_loss_fm[-1], fm2.last_fit_stats["h2d_bytes_rows"], fm.last_fit_stats["h2d_bytes_rows"]))

    # --- beyond the reference: rank the WHOLE catalog for every user on the tensor cores, exact float64 result ---
    t = log.tables
    user_table = csr_matrix((t["u_val"], t["u_col"], t["u_ptr"]), shape=(log.n_users, log.n_features))
    item_table = csr_matrix((t["i_val"], t["i_col"], t["i_ptr"]), shape=(log.n_items, log.n_features))
    for name, factors in (("FM", fm_factors(fm, user_table, item_table)), ("MF", mf_factors(mf))):
        scorer = TopKScorer(*factors)
        items, scores = scorer.topk(5)
        full = FullCatalogEvaluator(log.test_frame, np.ones(log.n_items), K, used, log.n_users, log.n_items)
        res = full.evaluate(scorer)
        print("%s full catalog: user 0 -> items %s; DCG@3 %.4f  coverage@3 %.3f  (%s)" % (
            name, items[0].tolist(), res["DCG"][1], res["CatalogCoverage"][1], scorer.last_stats))


if __name__ == "__main__":
    main()
