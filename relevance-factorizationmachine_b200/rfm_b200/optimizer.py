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
