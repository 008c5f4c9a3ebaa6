"""The reference's experiment drivers, run against the drop-in classes through the reference's own import paths.

``main_coat.py:76-135`` (model x estimator loop, predict, TestEvaluator, random baseline) and
``utils/search_params.py:59-152`` (epoch search with a ValEvaluator hooked into ``fit``) need Hydra, the Coat files
and matplotlib to run as scripts; their loop bodies are executed here verbatim in structure -- same constructor keyword
arguments, same method calls, same result handling -- on a stub data loader with the Coat-shaped synthetic log, and the
outcome is compared with the same flow over the CPU oracle."""
import json

import numpy as np
import pandas as pd
import pytest

from oracle import fm_oracle, metrics_oracle, mf_oracle

pytestmark = pytest.mark.gpu

K = [1, 3, 5, 7, 9]
QUANTITATIVE_METRIC, QUALITATIVE_METRIC = "DCG", "CatalogCoverage"
ESTIMATORS, MODELS = ["IPS", "Naive"], ["FM", "MF"]


class StubLoader:
    """What utils/dataloader/coat/loadar.py exposes to the drivers."""

    def __init__(self):
        from rfm_b200.synth import make_coat_shaped
        log = make_coat_shaped(seed=11, n_users=80, n_items=90, n_rated=14, n_test=9)
        self.log, self.n_users, self.n_items = log, log.n_users, log.n_items
        self.test_df = pd.DataFrame(log.test_frame)
        self.val_df = pd.DataFrame(log.test_frame)
        feats = {"FM": log.fm_test_features, "MF": log.mf_test_features}
        self.test_evaluation_features = self.val_evaluation_features = feats

    def load(self, model_name, estimator):
        train, val = (self.log.fm_train, self.log.fm_val) if model_name == "FM" else (self.log.mf_train, self.log.mf_val)
        train, val = dict(train), dict(val)
        if estimator == "Naive":                     # coat/loadar.py hands ones as propensities to the naive estimator
            train["pscores"] = np.ones_like(train["pscores"])
            val["pscores"] = np.ones_like(val["pscores"])
        return train, val


MODEL_PARAMS = {"n_factors": 12, "batch_size": 300, "reg": 0.3,
                "lr": {"FM": {"IPS": 1e-3, "Naive": 2e-3}, "MF": {"IPS": 0.02, "Naive": 0.02}}}


def _oracle_model(model_name, estimator, n_epochs, dataloader, seed):
    train, val = dataloader.load(model_name, estimator)
    lr = MODEL_PARAMS["lr"][model_name][estimator]
    if model_name == "FM":
        w0, w, V = fm_oracle.fm_init(seed, train["features"].shape[1], MODEL_PARAMS["n_factors"])
        (w0, w, V), _, _ = fm_oracle.fm_fit(train, val, n_epochs, MODEL_PARAMS["batch_size"], lr, w0, w, V)
        return fm_oracle.fm_predict(dataloader.test_evaluation_features["FM"], w0, w, V)
    P, Q, bu, bi = mf_oracle.mf_init(seed, dataloader.n_users, dataloader.n_items, MODEL_PARAMS["n_factors"])
    (P, Q, bu, bi, b), _, _ = mf_oracle.mf_fit(train, val, n_epochs, MODEL_PARAMS["batch_size"], lr,
                                              MODEL_PARAMS["reg"], P, Q, bu, bi)
    return mf_oracle.mf_predict(dataloader.test_evaluation_features["MF"], P, Q, bu, bi, b)


def test_main_coat_loop_body(tmp_path):
    # the reference's import paths, resolved to this build by the alias packages
    from src.fm import FactorizationMachines as FM
    from src.mf import LogisticMatrixFactorization as MF
    from utils.evaluate import TestEvaluator
    seed = 12345
    dataloader = StubLoader()
    params_path = tmp_path
    for model_name in MODELS:
        for estimator in ESTIMATORS:
            with open(params_path / f"{model_name}_{estimator}.json", "w") as f:
                json.dump({"n_epochs": 7, "val_DCG": 0.0}, f)
    # ---- main_coat.py:76-135 ----
    evaluator = TestEvaluator(interaction_df=dataloader.test_df, features=dataloader.test_evaluation_features,
                              n_items=dataloader.n_items, used_metrics={QUANTITATIVE_METRIC, QUALITATIVE_METRIC}, K=K)
    metric_df = pd.DataFrame()
    for model_name in MODELS:
        for estimator in ESTIMATORS:
            base_name = f"{model_name}_{estimator}"
            train, val = dataloader.load(model_name=model_name, estimator=estimator)
            with open(params_path / f"{base_name}.json", "r") as f:
                search_results = json.load(f)
            if model_name == "FM":
                model = FM(estimator=estimator, n_epochs=search_results["n_epochs"], n_factors=MODEL_PARAMS["n_factors"],
                           n_features=train["features"].shape[1], lr=MODEL_PARAMS["lr"][model_name][estimator],
                           batch_size=MODEL_PARAMS["batch_size"], seed=seed)
            elif model_name == "MF":
                model = MF(estimator=estimator, n_epochs=search_results["n_epochs"], n_factors=MODEL_PARAMS["n_factors"],
                           n_users=dataloader.n_users, n_items=dataloader.n_items,
                           lr=MODEL_PARAMS["lr"][model_name][estimator], reg=MODEL_PARAMS["reg"],
                           batch_size=MODEL_PARAMS["batch_size"], seed=seed)
            _ = model.fit(train, val)
            test_pred_y = model.predict(X=evaluator.features[model_name])
            results = evaluator.evaluate(test_pred_y)
            for metric_name, values in results.items():
                metric_df[f"{base_name}_{metric_name}@K"] = values
    np.random.seed(seed)
    results = evaluator.evaluate(y_scores=np.random.uniform(0, 1, size=len(dataloader.test_df)))
    for metric_name, values in results.items():
        metric_df[f"Random_{metric_name}@K"] = values
    metric_df.to_csv(tmp_path / "metric.csv", index=False)
    # ---- the same flow over the CPU oracle ----
    frame = dataloader.log.test_frame
    used = {QUANTITATIVE_METRIC, QUALITATIVE_METRIC}
    for model_name in MODELS:
        for estimator in ESTIMATORS:
            scores = _oracle_model(model_name, estimator, 7, dataloader, seed)
            ref = metrics_oracle.test_evaluate(frame, scores, K, used, dataloader.n_items)
            for metric_name in ("DCG", "CatalogCoverage", "ME"):
                np.testing.assert_allclose(metric_df[f"{model_name}_{estimator}_{metric_name}@K"], ref[metric_name],
                                           rtol=1e-9, err_msg=f"{model_name}_{estimator}_{metric_name}")
    np.random.seed(seed)
    ref = metrics_oracle.test_evaluate(frame, np.random.uniform(0, 1, size=len(dataloader.test_df)), K, used,
                                       dataloader.n_items)
    for metric_name in ("DCG", "CatalogCoverage", "ME"):
        np.testing.assert_allclose(metric_df[f"Random_{metric_name}@K"], ref[metric_name], rtol=1e-12)
    assert set(metric_df.columns) == {f"{m}_{e}_{n}@K" for m in MODELS for e in ESTIMATORS
                                      for n in ("ME", "DCG", "CatalogCoverage")} | \
        {f"Random_{n}@K" for n in ("ME", "DCG", "CatalogCoverage")}


def test_search_params_loop_body(tmp_path):
    from src.fm import FactorizationMachines as FM
    from src.mf import LogisticMatrixFactorization as MF
    from utils.evaluate import ValEvaluator
    seed, max_epoch, k, used_metric = 12345, 9, 5, "DCG"
    dataloader = StubLoader()
    # ---- utils/search_params.py:59-152 ----
    evaluator = ValEvaluator(interaction_df=dataloader.val_df, features=dataloader.val_evaluation_features, k=k,
                             metric_name=used_metric)
    np.random.seed(seed)
    y_scores = np.random.uniform(0, 1, dataloader.val_df.shape[0])
    for estimator in ESTIMATORS:
        metric_value = evaluator.evaluate(y_scores=y_scores, estimator=estimator)
        np.testing.assert_allclose(metric_value, metrics_oracle.val_evaluate(dataloader.log.test_frame, y_scores, k,
                                                                              estimator), rtol=1e-12)
    for model_name in MODELS:
        for estimator in ESTIMATORS:
            train, val = dataloader.load(model_name=model_name, estimator=estimator)
            if model_name == "FM":
                model = FM(estimator=estimator, n_epochs=max_epoch, n_factors=MODEL_PARAMS["n_factors"],
                           n_features=train["features"].shape[1], lr=MODEL_PARAMS["lr"][model_name][estimator],
                           batch_size=MODEL_PARAMS["batch_size"], seed=seed, evaluator=evaluator)
            elif model_name == "MF":
                model = MF(estimator=estimator, n_epochs=max_epoch, n_factors=MODEL_PARAMS["n_factors"],
                           n_users=dataloader.n_users, n_items=dataloader.n_items,
                           lr=MODEL_PARAMS["lr"][model_name][estimator], reg=MODEL_PARAMS["reg"],
                           batch_size=MODEL_PARAMS["batch_size"], seed=seed, evaluator=evaluator)
            train_loss, val_loss = model.fit(train, val)
            best_epoch = int(np.argmax(model.val_metrics))
            metric_value = model.val_metrics[best_epoch]
            with open(tmp_path / f"{model_name}_{estimator}.json", "w") as f:
                json.dump({"n_epochs": best_epoch, f"val_{used_metric}": metric_value}, f)
            assert len(train_loss) == len(val_loss) == len(model.val_metrics) == max_epoch
            # the last epoch's metric equals the evaluator on the final model's predictions (the host flow)
            final = evaluator.evaluate(y_scores=model.predict(X=evaluator.features[model_name]), estimator=estimator)
            np.testing.assert_allclose(model.val_metrics[-1], final, rtol=1e-12)
            assert np.isfinite(metric_value)
