"""Base class of the point-wise recommenders (reference ``src/base.py:9-66``)."""
from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass

import numpy as np


@dataclass
class PointwiseBaseRecommender(ABC):
    estimator: str
    n_epochs: int
    n_factors: int
    lr: float
    batch_size: int
    seed: int

    @abstractmethod
    def fit(self, train, val) -> tuple:
        ...

    @abstractmethod
    def predict(self, **kwargs) -> np.ndarray:
        ...

    # The two helpers below exist for API fidelity (callers and subclasses of the reference may
    # use them on host arrays); the training loops evaluate the same formulas on the device.
    def _cross_entropy_loss(self, y_trues, y_scores, pscores, eps: float = 1e-8) -> float:
        r = np.asarray(y_trues) / np.asarray(pscores)
        return float(-np.sum(r * np.log(y_scores + eps) + (1 - r) * np.log(1 - y_scores + eps)) / len(y_trues))

    def _sigmoid(self, x):
        x = np.clip(x, -700, 700)
        return 1 / (1 + np.exp(-x))


class EvalChain:
    """Per-epoch evaluation inside ``fit(evaluator=...)`` (``src/fm.py:104-110``, ``src/mf.py:126-132``) without a
    host round trip: the model's predict kernel writes straight into the device ranker's score buffer, the ranker
    evaluates in place and parks the metric rows in a device history slot, and ``finish`` reads every epoch's value
    back once. Used when the evaluator is this package's ``ValEvaluator``; any other evaluator object gets the
    reference's host flow (predict -> ``evaluator.evaluate(y_scores=..., estimator=...)``)."""

    def __init__(self, evaluator, estimator: str, n_eval_rows: int, n_epochs: int, device: int):
        self.evaluator, self.n_epochs = evaluator, max(n_epochs, 1)
        self.ranker = evaluator.device_chain(estimator, device=device)
        if self.ranker.n_rows != n_eval_rows:
            raise ValueError("evaluator.features has %d rows, interaction_df has %d" % (n_eval_rows, self.ranker.n_rows))
        self.scores_ptr = self.ranker.scores_ptr()
        self.k = [int(evaluator.k)]

    @staticmethod
    def supported(evaluator) -> bool:
        return hasattr(evaluator, "device_chain") and hasattr(evaluator, "chain_results")

    def after_epoch(self, epoch: int) -> None:
        self.ranker.evaluate_dev(self.k, epoch, self.n_epochs)

    def finish(self, n_done: int) -> list:
        return self.evaluator.chain_results(self.ranker, n_done)
