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
