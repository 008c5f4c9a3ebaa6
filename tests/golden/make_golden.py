"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports ``src.fm``, ``src.mf`` and ``utils.evaluate`` from ``/root/reference``, runs them on
small seeded synthetic inputs, and stores inputs + outputs as ``.npz``. The reference's own
tests are empty stubs (``test/test_fm.py:15-16``), so these fixtures are the pin for the
oracle (``oracle/``) and, through it and directly, for the CUDA path.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("RFM_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(1, os.path.join(ROOT, "relevance-factorizationmachine_b200"))

import pandas as pd  # noqa: E402
from tqdm import tqdm  # noqa: E402,F401

from src.fm import FactorizationMachines  # noqa: E402  (reference)
from src.mf import LogisticMatrixFactorization  # noqa: E402  (reference)
from utils.evaluate import TestEvaluator, ValEvaluator  # noqa: E402  (reference)
from rfm_b200.synth import make_coat_shaped, make_kuairec_shaped  # noqa: E402

assert os.path.realpath(sys.modules["src.fm"].__file__).startswith(os.path.realpath(REF))


def csr_parts(X, prefix):
    X = X.tocsr()
    return {prefix + "_indptr": X.indptr.astype(np.int64), prefix + "_indices": X.indices.astype(np.int32),
            prefix + "_data": X.data.astype(np.float64), prefix + "_shape": np.array(X.shape, dtype=np.int64)}


def frame_df(frame):
    return pd.DataFrame({k: v for k, v in frame.items()})


def fm_case(name, log, k, B, lr, n_epochs, alpha, estimator="IPS", with_eval=True):
    ev = None
    if with_eval:
        ev = ValEvaluator(interaction_df=frame_df(log.test_frame), features={"FM": log.fm_test_features},
                          k=5, metric_name="DCG")
    train = dict(log.fm_train)
    if estimator == "Naive":
        train["pscores"] = np.ones_like(train["pscores"])
    m = FactorizationMachines(estimator=estimator, n_epochs=n_epochs, n_factors=k, lr=lr, batch_size=B,
                              seed=12345, n_features=log.n_features, alpha=alpha, evaluator=ev)
    init = dict(w0_init=m.w0().copy(), w_init=m.w().copy(), V_init=m.V().copy())
    tl, vl = m.fit(train, log.fm_val)
    out = dict(k=k, B=B, lr=lr, n_epochs=n_epochs, alpha=alpha, seed=12345,
               train_loss=np.array(tl), val_loss=np.array(vl),
               w0=m.w0().copy(), w=m.w().copy(), V=m.V().copy(),
               test_scores=m.predict(X=log.fm_test_features),
               val_metrics=np.array(m.val_metrics if with_eval else []), **init)
    out.update(csr_parts(train["features"], "train"))
    out.update(csr_parts(log.fm_val["features"], "val"))
    out.update(csr_parts(log.fm_test_features, "test"))
    out.update(train_labels=train["labels"], train_pscores=train["pscores"],
               val_labels=log.fm_val["labels"], val_pscores=log.fm_val["pscores"])
    out.update({"frame_" + c: v for c, v in log.test_frame.items()})
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "train_loss[-1]=%.6f val_loss[-1]=%.6f" % (tl[-1], vl[-1]))
    return m


def mf_case(name, log, k, B, lr, reg, n_epochs, alpha, estimator="IPS"):
    ev = ValEvaluator(interaction_df=frame_df(log.test_frame), features={"MF": log.mf_test_features},
                      k=5, metric_name="DCG")
    m = LogisticMatrixFactorization(estimator=estimator, n_epochs=n_epochs, n_factors=k, lr=lr, batch_size=B,
                                    seed=12345, n_users=log.n_users, n_items=log.n_items, reg=reg,
                                    alpha=alpha, evaluator=ev)
    init = dict(P_init=m.P().copy(), Q_init=m.Q().copy(), bu_init=m.b_u().copy(), bi_init=m.b_i().copy())
    tl, vl = m.fit(log.mf_train, log.mf_val)
    out = dict(k=k, B=B, lr=lr, reg=reg, n_epochs=n_epochs, alpha=alpha, seed=12345,
               n_users=log.n_users, n_items=log.n_items,
               train_loss=np.array(tl), val_loss=np.array(vl), P=m.P().copy(), Q=m.Q().copy(),
               b_u=m.b_u().copy(), b_i=m.b_i().copy(), b=float(m.b),
               test_scores=m.predict(log.mf_test_features), val_metrics=np.array(m.val_metrics),
               train_pairs=log.mf_train["features"], train_labels=log.mf_train["labels"],
               train_pscores=log.mf_train["pscores"], val_pairs=log.mf_val["features"],
               val_labels=log.mf_val["labels"], val_pscores=log.mf_val["pscores"],
               test_pairs=log.mf_test_features, **init)
    out.update({"frame_" + c: v for c, v in log.test_frame.items()})
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "train_loss[-1]=%.6f val_loss[-1]=%.6f" % (tl[-1], vl[-1]))
    return m


def eval_case(name, frame, scores, n_items, K=(1, 3, 5, 7, 9)):
    used = {"DCG", "CatalogCoverage", "Recall", "MAP", "Gini"}
    te = TestEvaluator(interaction_df=frame_df(frame), features={}, K=K, used_metrics=used, n_items=n_items)
    res = te.evaluate(scores)
    out = {"test_" + m: np.array(v, dtype=np.float64) for m, v in res.items()}
    for est in ("IPS", "Naive"):
        for k in (3, 5):
            ve = ValEvaluator(interaction_df=frame_df(frame), features={}, k=k, metric_name="DCG")
            out["val_%s_%d" % (est, k)] = np.float64(ve.evaluate(scores, est))
    # the reference's own per-user ranking (argsort()[::-1]) for the top-9, tie-free inputs only
    df = frame_df(frame)
    df["y_score"] = scores
    tops, users = [], []
    for user, g in df.groupby("user"):
        order = g["y_score"].to_numpy().argsort()[::-1][:9]
        rows = g.index.to_numpy()[order]
        tops.append(np.pad(rows, (0, 9 - len(rows)), constant_values=-1))
        users.append(user)
    out.update(K=np.array(K), n_items=n_items, scores=scores, top_rows=np.array(tops), top_users=np.array(users))
    out.update({"frame_" + c: v for c, v in frame.items()})
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, {m: np.round(v, 4).tolist() for m, v in res.items()})


def main():
    coat = make_coat_shaped(seed=2024)
    # C2: Coat-shaped IPS-FM, k=32; reference alpha (saturating logits) and a tame alpha
    m = fm_case("coat_fm_ips_alpha2", coat, k=32, B=500, lr=1e-4, n_epochs=40, alpha=2.0)
    fm_case("coat_fm_ips_alpha01", coat, k=32, B=500, lr=1e-3, n_epochs=40, alpha=0.1)
    fm_case("coat_fm_naive_alpha01", coat, k=32, B=500, lr=2e-3, n_epochs=20, alpha=0.1,
            estimator="Naive", with_eval=False)
    # C1: Coat-shaped IPS-MF, k=16
    mf = mf_case("coat_mf_ips", coat, k=16, B=500, lr=0.02, reg=0.5, n_epochs=40, alpha=4.0)
    mf_case("coat_mf_ips_alpha01", coat, k=16, B=500, lr=0.02, reg=0.01, n_epochs=40, alpha=0.1)
    # evaluators: tie-free scores (random) and the FM's own (tie-heavy at alpha=2) scores
    rng = np.random.default_rng(7)
    eval_case("coat_eval_tiefree", coat.test_frame, rng.random(coat.test_frame["user"].size), coat.n_items)
    eval_case("coat_eval_mf", coat.test_frame, mf.predict(coat.mf_test_features), coat.n_items)
    # ragged user lists (some shorter than k, some without positives)
    frame = {c: v[:900].copy() for c, v in coat.test_frame.items()}
    frame["label"][frame["user"] % 7 == 0] = 0
    eval_case("coat_eval_ragged", frame, rng.random(900), coat.n_items)

    # KuaiRec-shaped (real-valued columns, ragged rows), small enough for the live reference
    kr = make_kuairec_shaped(seed=2025, n_users=300, n_items=400, n_train=6000, n_val=600,
                             eval_users=60, eval_items=200, eval_rows_per_user=20)
    fm_case("kuairec_small_fm_ips", kr, k=64, B=2000, lr=9e-6, n_epochs=12, alpha=2.0)
    fm_case("kuairec_small_fm_ips_alpha01", kr, k=64, B=2000, lr=1e-4, n_epochs=12, alpha=0.1)

    # legacy sampler known answers: RandomState(epoch).shuffle(arange(N))[:B]
    from sklearn.utils import resample
    cases = [(10, 10, 0), (1, 1, 3), (2, 1, 5), (3660, 500, 0), (3660, 500, 39), (100003, 257, 123456),
             (65536, 64, 7), (65537, 64, 7), (1 << 20, 100, 2**31 + 5)]
    out = {}
    for N, B, ep in cases:
        out["N%d_B%d_e%d" % (N, B, ep)] = resample(np.arange(N), replace=False, n_samples=B, random_state=ep)
    np.savez_compressed(os.path.join(HERE, "legacy_sampler.npz"), **out)
    print("legacy_sampler", len(out), "cases")


if __name__ == "__main__":
    main()
