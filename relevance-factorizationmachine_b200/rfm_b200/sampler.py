"""Host-side batch-index producers for ``fit``.

* ``legacy``  -- the reference's sampler, ``sklearn.utils.resample(replace=False,
  n_samples=B, random_state=epoch)`` (``src/fm.py:72-79``, ``src/mf.py:88-95``), reproduced bit
  for bit by the C++ MT19937 + legacy-shuffle replica in ``csrc/sampler.cu``. One shuffle is
  O(N) sequential work, but epochs are independent (seed == epoch), so a
  thread pool computes several epochs ahead; ctypes releases the GIL during the call.
* ``feistel`` -- perf mode; drawn on the device inside the train step, nothing to do here.
"""
from __future__ import annotations

import os
from collections import deque
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import _capi


class LegacyBatchPrefetcher:
    def __init__(self, n_rows: int, batch: int, epochs, n_threads: int = 0):
        if batch > n_rows:
            raise ValueError(
                "Cannot sample %d out of arrays with dim %d when replace is False" % (batch, n_rows))
        self.n_rows, self.batch = n_rows, batch
        self.epochs = iter(epochs)
        # small N: the shuffle is microseconds, threads only add latency
        self.n_threads = n_threads or (1 if n_rows < 200_000 else min(16, os.cpu_count() or 1))
        self.pool = ThreadPoolExecutor(self.n_threads) if self.n_threads > 1 else None
        self.pending = deque()
        self._scratch = [np.empty(n_rows, dtype=np.int32) for _ in range(self.n_threads)] if self.pool else None
        self._free = deque(range(self.n_threads)) if self.pool else None
        self._fill()

    def _job(self, epoch, slot):
        out = _capi.legacy_batch(self.n_rows, self.batch, epoch, self._scratch[slot])
        return out, slot

    def _fill(self):
        if not self.pool:
            return
        while self._free:
            try:
                epoch = next(self.epochs)
            except StopIteration:
                return
            slot = self._free.popleft()
            self.pending.append(self.pool.submit(self._job, epoch, slot))

    def next(self) -> np.ndarray:
        if not self.pool:
            return _capi.legacy_batch(self.n_rows, self.batch, next(self.epochs))
        out, slot = self.pending.popleft().result()
        self._free.append(slot)
        self._fill()
        return out

    def close(self):
        if self.pool:
            self.pool.shutdown(wait=True, cancel_futures=True)
            self.pool = None
