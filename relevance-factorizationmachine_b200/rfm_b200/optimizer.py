"""Parameter holders with the reference's optimizer interface (``utils/optimizer.py:10-64``).

``model.V`` etc. are instances of these: callables returning the host ndarray (or an indexed
view), with ``.params``, ``.lr`` and ``.update(grad, index)``. During ``fit`` the master copy
lives on the device; the holder's ``params`` array is refreshed in place when ``fit`` returns,
so code that reads ``model.V()`` afterwards sees exactly what the reference would hold.
Host-side ``update`` calls bump ``version`` so the next device use re-uploads.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass, field
from typing import Optional, Union

import numpy as np


@dataclass
class BaseOptimizer(ABC):
    params: np.ndarray
    lr: float
    version: int = field(default=0, repr=False, compare=False)

    @abstractmethod
    def update(self, grad: Union[float, np.ndarray], index: Optional[Union[int, tuple]]) -> None:
        ...

    def __call__(self, index=None) -> Union[np.ndarray, float]:
        if index is None:
            return self.params
        return self.params[index]


@dataclass
class SGD(BaseOptimizer):
    """``params[index] -= lr * grad`` (``utils/optimizer.py:52-64``)."""

    def update(self, grad, index) -> None:
        if index is None:
            self.params -= self.lr * grad
        else:
            self.params[index] -= self.lr * grad
        self.version += 1


@dataclass
class Adam(BaseOptimizer):
    """Adam (Kingma & Ba) behind the reference's holder contract; not in the reference, whose
    ``utils/optimizer.py`` ends with SGD at line 64. Specification: ``oracle/optimizer_oracle.py``.

    ``update(grad, index)`` advances only the entries selected by ``index`` (whole array if None):
    ``m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; params -= lr (m/(1-b1^t)) / (sqrt(v/(1-b2^t)) + eps)``
    with one step counter ``t`` per holder, incremented on every call. ``l2`` adds ``l2 * params`` to the
    gradient first. During ``FactorizationMachines.fit`` the same rule runs on the device
    (``rfm_fm_train_epoch_opt``) with moments that start at zero for that fit."""
    beta1: float = 0.9
    beta2: float = 0.999
    eps: float = 1e-8
    l2: float = 0.0
    t: int = field(default=0, repr=False, compare=False)
    m: Optional[np.ndarray] = field(default=None, repr=False, compare=False)
    v: Optional[np.ndarray] = field(default=None, repr=False, compare=False)

    def update(self, grad, index) -> None:
        if self.m is None:
            self.m = np.zeros_like(self.params, dtype=np.float64)
            self.v = np.zeros_like(self.params, dtype=np.float64)
        self.t += 1
        sel = slice(None) if index is None else index
        g = grad + self.l2 * self.params[sel]
        self.m[sel] = self.beta1 * self.m[sel] + (1 - self.beta1) * g
        self.v[sel] = self.beta2 * self.v[sel] + (1 - self.beta2) * (g * g)
        m_hat = self.m[sel] / (1 - self.beta1 ** self.t)
        v_hat = self.v[sel] / (1 - self.beta2 ** self.t)
        self.params[sel] -= self.lr * m_hat / (np.sqrt(v_hat) + self.eps)
        self.version += 1
